from . import backend, ops  # noqa: F401
