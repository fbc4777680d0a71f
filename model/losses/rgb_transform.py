from stylemesh_b200.model.losses.rgb_transform import post, pre  # noqa: F401
