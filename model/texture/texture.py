from stylemesh_b200.model.texture.texture import *  # noqa: F401,F403
from stylemesh_b200.model.texture.texture import HierarchicalNeuralTexture, NeuralTexture, to_image  # noqa: F401
