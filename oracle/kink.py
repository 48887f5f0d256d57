"""Pinning the oracle's GEMM outputs for gradient comparisons through LeakyReLU (TEST INFRASTRUCTURE, see
oracle/__init__.py).

GATv2 applies LeakyReLU to z = h_i + h_j with h = x W (layers/gatv2_conv.py:241-284 of the reference); its derivative
jumps from ``negative_slope`` to 1 at z = 0.  Any two fp32 implementations of the GEMM (another summation order,
3xTF32 on the tensor cores) differ by ~1e-7..1e-6 relative, so among millions of (edge, channel) pairs a handful of
z change SIGN between the two - and each such flip moves a finite share of every gradient downstream (measured on
BASELINE's C2: one flip shifts dL/dW by 4e-5 of its scale).  That is a property of the reference's function, not of
either implementation, and it makes a 1e-5 gradient comparison at full size a coin toss.

``pinned_matmul(values)`` removes the ambiguity without loosening any tolerance: inside the context the k-th
``keras_ops.matmul`` of the oracle returns ``values[k]`` (the h the GPU computed - the GEMM itself is compared against
float64 at 1e-5 in tests/test_gpu_parity.py::test_linear_tensor_core_gemm) while its gradient still flows through the
oracle's own matmul.  z is then one fp32 addition of identical operands on both sides, the sign patterns coincide
bit for bit, and forward, input gradients and weight gradients are compared at the plain 1e-5 tolerance."""
from __future__ import annotations

from contextlib import contextmanager

import torch

from . import keras_ops


class _StraightThrough(torch.autograd.Function):
    """forward: ``values``; backward: identity onto ``computed`` (the oracle's own matmul result)."""

    @staticmethod
    def forward(ctx, computed, values):
        if tuple(computed.shape) != tuple(values.shape):
            raise ValueError(f"pinned value has shape {tuple(values.shape)}, the oracle computed {tuple(computed.shape)}")
        return values.to(computed.dtype).clone()

    @staticmethod
    def backward(ctx, g):
        return g, None


@contextmanager
def pinned_matmul(values):
    """Inside the context the k-th call of ``keras_ops.matmul`` returns ``values[k]`` (gradient: straight through to
    the oracle's own product).  Raises if the oracle multiplies more often than values were supplied; ``stats``
    (yielded) records the largest relative deviation between the pinned and the computed values."""
    orig = keras_ops.matmul
    stats = {"calls": 0, "max_rel_dev": 0.0}

    def pinned(a, b):
        k = stats["calls"]
        if k >= len(values):
            raise RuntimeError("pinned_matmul: the oracle called matmul more often than values were pinned")
        stats["calls"] = k + 1
        computed = orig(a, b)
        v = torch.as_tensor(values[k])
        dev = float((computed.detach().double() - v.double()).abs().max()) / (float(v.double().abs().max()) + 1e-300)
        stats["max_rel_dev"] = max(stats["max_rel_dev"], dev)
        return _StraightThrough.apply(computed, v)

    keras_ops.matmul = pinned
    try:
        yield stats
    finally:
        keras_ops.matmul = orig
