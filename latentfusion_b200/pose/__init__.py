from . import estimation, utils  # noqa: F401
