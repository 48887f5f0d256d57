"""Restatement of the Keras-3 torch-backend primitives used by the reference hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function mirrors the behaviour of
``keras.ops.<name>`` under ``KERAS_BACKEND=torch`` as published in
keras/src/backend/torch/numpy.py and keras/src/backend/torch/math.py (Keras >= 3.0, the
reference's only pin: /root/reference/pyproject.toml:33).  Call sites in the reference:

  take         layers/message_passing.py:195-196, layers/sage_conv.py:331-332,
               layers/gatv2_conv.py:245-246,299,308, utils/main.py:30-31, aggregators.py:208
  segment_sum  layers/aggregators.py:67,72,135,194,198,215, utils/main.py:24,
               layers/gatv2_conv.py:305,326
  segment_max  layers/aggregators.py:108,162,270, layers/gatv2_conv.py:298

Everything runs on the CPU in float32 exactly like the reference would with the torch
backend forced onto the host.
"""

from __future__ import annotations

import numpy as np
import torch

_FLOATX = "float32"


def floatx() -> str:
    return _FLOATX


_DTYPES = {
    "float32": torch.float32,
    "float64": torch.float64,
    "float16": torch.float16,
    "bfloat16": torch.bfloat16,
    "int32": torch.int32,
    "int64": torch.int64,
    "bool": torch.bool,
}


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    if isinstance(dtype, str):
        return _DTYPES[dtype]
    return _DTYPES[np.dtype(dtype).name]


def convert_to_tensor(x, dtype=None):
    """keras.ops.convert_to_tensor: numpy/python -> torch tensor (host here)."""
    if isinstance(x, torch.Tensor):
        return x.to(_dt(dtype)) if dtype is not None else x
    if hasattr(x, "value") and isinstance(getattr(x, "value"), torch.Tensor):
        x = x.value  # shim Variable
        return x.to(_dt(dtype)) if dtype is not None else x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    else:
        arr = np.asarray(x)
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)  # python floats take floatx
        t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.to(_dt(dtype)) if dtype is not None else t


def convert_to_numpy(x):
    return convert_to_tensor(x).detach().cpu().numpy()


def shape(x):
    return tuple(convert_to_tensor(x).shape)


def cast(x, dtype):
    return convert_to_tensor(x).to(_dt(dtype))


def take(x, indices, axis=None):
    """keras torch backend ``take``: int64 indices, negatives wrapped, 2-D/axis-0 goes
    through ``torch.nn.functional.embedding`` (so OOB raises IndexError on the CPU)."""
    x = convert_to_tensor(x)
    indices = convert_to_tensor(indices).long()
    if axis is None:
        return torch.take(x.reshape(-1), indices)
    dim = x.shape[axis]
    indices = torch.where(indices < 0, indices + dim, indices)
    if x.ndim == 2 and axis == 0:
        return torch.nn.functional.embedding(indices, x)
    flat = torch.index_select(x, axis, indices.reshape(-1))
    new_shape = tuple(x.shape[:axis]) + tuple(indices.shape) + tuple(x.shape[axis + 1:])
    return flat.reshape(new_shape)


def _segment_reduction(data, segment_ids, reduction, num_segments, sorted_=False):
    """keras/src/backend/torch/math.py::_segment_reduction_fn.

    ids are repeat_interleaved to the data's shape (int64), ids outside
    [0, num_segments) are redirected to an extra trailing row which is dropped,
    accumulation happens in float32 on a (num_segments + 1, ...) buffer via
    ``scatter_reduce(..., include_self=False)`` and the result is cast back.
    """
    data = convert_to_tensor(data)
    segment_ids = convert_to_tensor(segment_ids)
    num_repeats = int(np.prod(data.shape[1:])) if data.ndim > 1 else 1
    ids = segment_ids.repeat_interleave(num_repeats).view(*data.shape).long()
    num_segments = int(num_segments) if num_segments is not None else int(ids.max()) + 1
    ids = torch.where(ids >= 0, ids, num_segments)
    ids = torch.where(ids < num_segments, ids, num_segments)
    shp = (num_segments + 1,) + tuple(data.shape[1:])
    if reduction == "amax":
        result = torch.ones(*shp) * -float("inf")
    else:
        result = torch.zeros(*shp)
    result = result.scatter_reduce(0, ids, data.float(), reduction, include_self=False)
    result = result[:-1, ...]
    return result.type(data.dtype)


def segment_sum(data, segment_ids, num_segments=None, sorted=False):
    return _segment_reduction(data, segment_ids, "sum", num_segments, sorted)


def segment_max(data, segment_ids, num_segments=None, sorted=False):
    return _segment_reduction(data, segment_ids, "amax", num_segments, sorted)


# ---- element-wise / shape glue (thin wrappers; all follow torch semantics) -------------

def zeros(shape, dtype=None):
    return torch.zeros(tuple(shape), dtype=_dt(dtype or _FLOATX))


def ones(shape, dtype=None):
    return torch.ones(tuple(shape), dtype=_dt(dtype or _FLOATX))


def zeros_like(x, dtype=None):
    return torch.zeros_like(convert_to_tensor(x), dtype=_dt(dtype))


def ones_like(x, dtype=None):
    return torch.ones_like(convert_to_tensor(x), dtype=_dt(dtype))


def arange(start, stop=None, step=1, dtype=None):
    if stop is None:
        start, stop = 0, start
    return torch.arange(start, stop, step, dtype=_dt(dtype or "int32"))


def _bin(a, b):
    a_t = isinstance(a, torch.Tensor) or hasattr(a, "value")
    b_t = isinstance(b, torch.Tensor) or hasattr(b, "value")
    a = convert_to_tensor(a) if (a_t or not isinstance(a, (int, float))) else a
    b = convert_to_tensor(b) if (b_t or not isinstance(b, (int, float))) else b
    return a, b


def add(a, b):
    a, b = _bin(a, b)
    return a + b


def subtract(a, b):
    a, b = _bin(a, b)
    return a - b


def multiply(a, b):
    a, b = _bin(a, b)
    return a * b


def divide(a, b):
    a, b = _bin(a, b)
    return a / b


def maximum(a, b):
    a, b = _bin(a, b)
    if not isinstance(b, torch.Tensor):
        b = torch.tensor(b, dtype=a.dtype)
    if not isinstance(a, torch.Tensor):
        a = torch.tensor(a, dtype=b.dtype)
    return torch.maximum(a, b)


def power(a, b):
    a, b = _bin(a, b)
    return torch.pow(a, b)


def negative(x):
    return -convert_to_tensor(x)


def square(x):
    return torch.square(convert_to_tensor(x))


def sqrt(x):
    return torch.sqrt(convert_to_tensor(x))


def exp(x):
    return torch.exp(convert_to_tensor(x))


def isinf(x):
    return torch.isinf(convert_to_tensor(x))


def where(cond, a, b):
    a, b = _bin(a, b)
    return torch.where(convert_to_tensor(cond), a, b)


def expand_dims(x, axis):
    return torch.unsqueeze(convert_to_tensor(x), axis)


def reshape(x, newshape):
    return torch.reshape(convert_to_tensor(x), tuple(int(s) for s in newshape))


def transpose(x, axes=None):
    x = convert_to_tensor(x)
    if axes is None:
        return x.permute(*reversed(range(x.ndim)))
    return x.permute(*axes)


def stack(xs, axis=0):
    return torch.stack([convert_to_tensor(t) for t in xs], dim=axis)


def concatenate(xs, axis=0):
    return torch.cat([convert_to_tensor(t) for t in xs], dim=axis)


def matmul(a, b):
    return torch.matmul(convert_to_tensor(a), convert_to_tensor(b))


def sum(x, axis=None, keepdims=False):  # noqa: A001 - keras name
    x = convert_to_tensor(x)
    if axis is None:
        return torch.sum(x)
    return torch.sum(x, dim=axis, keepdim=keepdims)


def mean(x, axis=None, keepdims=False):
    x = convert_to_tensor(x)
    if axis is None:
        return torch.mean(x)
    return torch.mean(x, dim=axis, keepdim=keepdims)


def max(x, axis=None, keepdims=False):  # noqa: A001 - keras name
    x = convert_to_tensor(x)
    if axis is None:
        return torch.max(x)
    return torch.amax(x, dim=axis, keepdim=keepdims)


def leaky_relu(x, negative_slope=0.2):
    return torch.nn.functional.leaky_relu(convert_to_tensor(x), negative_slope=negative_slope)


def relu(x):
    return torch.relu(convert_to_tensor(x))


def normalize(x, axis=-1, order=2, epsilon=None):
    """keras.ops.normalize (torch backend): x / max(||x||_order, eps), eps = 1e-12
    (keras/src/backend/torch/nn.py::_l2_normalize / ops.normalize default epsilon)."""
    x = convert_to_tensor(x)
    eps = 1e-12 if epsilon is None else epsilon
    norm = torch.linalg.vector_norm(x, ord=order, dim=axis, keepdim=True)
    return x / torch.clamp(norm, min=eps)


def slice_update(inputs, start_indices, updates):
    inputs = convert_to_tensor(inputs).clone()
    updates = convert_to_tensor(updates)
    sl = tuple(slice(int(s), int(s) + int(n)) for s, n in zip(start_indices, updates.shape))
    inputs[sl] = updates
    return inputs
