from stylemesh_b200.model.model import *  # noqa: F401,F403
from stylemesh_b200.model.model import (FusedTextureAdam, TextureOptimizationStyleTransferPipeline,  # noqa: F401
                                        find_pyramid_size, to_tensor_image)
