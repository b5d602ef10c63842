from . import rotations, utils  # noqa: F401
