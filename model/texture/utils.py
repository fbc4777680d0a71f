from stylemesh_b200.model.texture.utils import *  # noqa: F401,F403
