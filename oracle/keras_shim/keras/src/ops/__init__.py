from ..backend.common.keras_tensor import KerasTensor  # noqa: F401
