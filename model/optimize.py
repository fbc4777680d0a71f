from stylemesh_b200.model.optimize import *  # noqa: F401,F403
from stylemesh_b200.model.optimize import build_parser, main, register_datamodule  # noqa: F401

if __name__ == "__main__":
    main(build_parser().parse_args())
