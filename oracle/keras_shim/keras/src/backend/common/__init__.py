from . import keras_tensor  # noqa: F401
