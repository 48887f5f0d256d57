from . import common  # noqa: F401
