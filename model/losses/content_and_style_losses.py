from stylemesh_b200.model.losses.content_and_style_losses import *  # noqa: F401,F403
from stylemesh_b200.model.losses.content_and_style_losses import (VGG, ContentAndStyleLoss, GramMatrix,  # noqa: F401
                                                                   image_pyramid)
