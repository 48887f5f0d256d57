"""keras.ops -> oracle.keras_ops (restated Keras torch-backend primitives)."""
from oracle.keras_ops import *  # noqa: F401,F403
from oracle.keras_ops import max, sum  # noqa: F401,A004
