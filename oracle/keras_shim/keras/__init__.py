"""Minimal stand-in for the ``keras`` package (torch backend, host only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  It exists so that the UNMODIFIED reference
sources under /root/reference/src can be imported in the build container, where real Keras
is not installable (no network, no wheel).  Only what the reference's hot-path modules touch
is provided; primitive semantics come from oracle/keras_ops.py.
"""
import torch as _torch

from . import activations, backend, constraints, initializers, layers, ops, regularizers, src  # noqa: F401
from .layers import Layer, Sequential  # noqa: F401
from .src.backend.common.keras_tensor import KerasTensor  # noqa: F401

Variable = _torch.nn.Parameter
__version__ = "3.0.0+oracle.shim"
