"""Keras-facing plumbing.

Under Keras 3 with ``KERAS_BACKEND=torch`` the layers in this package subclass
``keras.layers.Layer`` (so they drop into ``keras.Model`` / ``fit`` unchanged).  Keras is not
installable in the build image (no network), so when ``import keras`` fails a small local
implementation of the same protocol (``add_weight`` / ``build`` / ``__call__`` / ``get_config`` /
``get_weights``) is used instead; weights are ``torch.nn.Parameter``s on the current CUDA device.
Only what the four conv layers and ``MessagePassing`` need is provided.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

HAVE_KERAS = False
try:  # pragma: no cover - Keras is absent in the build image
    if os.environ.get("KGB200_FORCE_SHIM", "0") != "1":
        import keras  # type: ignore

        if getattr(keras.backend, "backend", lambda: None)() == "torch" and "oracle.shim" not in getattr(keras, "__version__", ""):
            HAVE_KERAS = True
except Exception:  # noqa: BLE001
    HAVE_KERAS = False


def default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def to_device_tensor(x, dtype=None, what: str = "input") -> torch.Tensor:
    """numpy / list / tensor -> tensor on the current CUDA device (Keras' torch backend does the
    same in ``convert_to_tensor``).  Raises when no CUDA device exists: there is no CPU path."""
    if isinstance(x, torch.Tensor) and x.device.type == "meta":
        raise RuntimeError("symbolic (meta) tensors cannot be fed to the CUDA kernels; "
                           "use compute_output_shape for shape inference")
    if not torch.cuda.is_available():
        raise RuntimeError("keras_geometric_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if hasattr(x, "value") and isinstance(getattr(x, "value"), torch.Tensor):
        x = x.value
    if not isinstance(x, torch.Tensor):
        arr = np.asarray(x)
        if arr.dtype == np.float64 and dtype is None:
            arr = arr.astype(np.float32)
        x = torch.from_numpy(np.ascontiguousarray(arr))
    if not x.is_cuda:
        x = x.to(default_device())
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x


def value_of(v):
    """Parameter (local shim) or keras Variable -> torch tensor."""
    if v is None or isinstance(v, torch.Tensor):
        return v
    return v.value if hasattr(v, "value") else v


def apply_dense(layer, x: torch.Tensor) -> torch.Tensor:
    """``Dense`` forward on the hand-written tensor-core kernels (K8): ``act(x @ kernel + bias)`` with bias and ReLU
    in the GEMM epilogue.  Takes the layer's WEIGHTS, not its ``call``, so it is the same code whether ``layer`` is
    the local shim or a real ``keras.layers.Dense`` (whose own call would run cuBLAS through ``torch.matmul``)."""
    from . import ops
    kernel, bias = value_of(layer.kernel), value_of(getattr(layer, "bias", None))
    act = getattr(layer, "activation", None)
    name = getattr(act, "__name__", None) if act is not None else "linear"
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
    if name in ("relu", "linear"):
        y = ops.linear(x2, kernel, bias=bias, act=None if name == "linear" else "relu")
    else:
        y = act(ops.linear(x2, kernel, bias=bias))
    return y if x.dim() == 2 else y.reshape(*lead, y.shape[-1])


def apply_mlp(seq, x: torch.Tensor, training=None) -> torch.Tensor:
    """A ``Sequential`` of Dense / Dropout layers (GINConv.mlp, gin_conv.py:129-159) with every Dense on K8."""
    for lyr in seq.layers:
        if getattr(lyr, "kernel", None) is not None and hasattr(lyr, "units"):
            x = apply_dense(lyr, x)
        else:
            x = lyr(x, training=training)
    return x


if HAVE_KERAS:  # pragma: no cover
    from keras import activations, constraints, initializers, regularizers  # noqa: F401
    from keras.layers import Dense, Dropout, Layer  # noqa: F401
    from keras import Sequential  # noqa: F401
else:
    # ------------------------------------------------------------------ initializers
    class _Initializers:
        class Initializer:
            def __call__(self, shape, dtype=None):
                raise NotImplementedError

            def get_config(self):
                return {}

        class Zeros(Initializer):
            def __call__(self, shape, dtype=None):
                return torch.zeros(tuple(shape))

        class Ones(Initializer):
            def __call__(self, shape, dtype=None):
                return torch.ones(tuple(shape))

        class Constant(Initializer):
            def __init__(self, value=0.0):
                self.value = value

            def __call__(self, shape, dtype=None):
                return torch.full(tuple(shape), float(self.value))

            def get_config(self):
                return {"value": self.value}

        class GlorotUniform(Initializer):
            def __init__(self, seed=None):
                self.seed = seed

            def __call__(self, shape, dtype=None):
                shape = tuple(shape)
                if len(shape) == 0:
                    fan_in = fan_out = 1
                elif len(shape) == 1:
                    fan_in = fan_out = shape[0]
                elif len(shape) == 2:
                    fan_in, fan_out = shape
                else:
                    rf = int(np.prod(shape[:-2]))
                    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
                lim = math.sqrt(6.0 / max(1.0, fan_in + fan_out))
                gen = torch.Generator().manual_seed(self.seed) if self.seed is not None else None
                return (torch.rand(shape, generator=gen) * 2 - 1) * lim

            def get_config(self):
                return {"seed": self.seed}

        class HeUniform(GlorotUniform):
            def __call__(self, shape, dtype=None):
                shape = tuple(shape)
                fan_in = shape[0] if len(shape) >= 1 else 1
                lim = math.sqrt(6.0 / max(1.0, fan_in))
                gen = torch.Generator().manual_seed(self.seed) if self.seed is not None else None
                return (torch.rand(shape, generator=gen) * 2 - 1) * lim

        _BY_NAME = {}

        def get(self, identifier):
            if identifier is None:
                return None
            if isinstance(identifier, self.Initializer):
                return identifier
            if isinstance(identifier, str):
                try:
                    return self._BY_NAME[identifier.lower()]()
                except KeyError:
                    raise ValueError(f"Unknown initializer: {identifier}") from None
            if isinstance(identifier, dict):
                return self.deserialize(identifier)
            if callable(identifier):
                return identifier
            raise ValueError(f"Could not interpret initializer identifier: {identifier}")

        def serialize(self, init):
            if init is None:
                return None
            for k, v in self._BY_NAME.items():
                if type(init) is v:
                    return {"class_name": type(init).__name__, "config": init.get_config(), "registered_name": k}
            return init

        def deserialize(self, cfg):
            if cfg is None or isinstance(cfg, self.Initializer) or callable(cfg):
                return cfg
            if isinstance(cfg, str):
                return self.get(cfg)
            name = cfg.get("registered_name") or cfg["class_name"]
            for k, v in self._BY_NAME.items():
                if k == name.lower() or v.__name__ == cfg["class_name"]:
                    return v(**cfg.get("config", {}))
            raise ValueError(f"Unknown initializer config: {cfg}")

    initializers = _Initializers()
    _Initializers._BY_NAME.update({"zeros": _Initializers.Zeros, "ones": _Initializers.Ones,
                                   "constant": _Initializers.Constant,
                                   "glorot_uniform": _Initializers.GlorotUniform,
                                   "he_uniform": _Initializers.HeUniform})

    # ------------------------------------------------------------------ activations
    class _Activations:
        @staticmethod
        def linear(x):
            return x

        @staticmethod
        def relu(x):
            return torch.relu(x)

        @staticmethod
        def tanh(x):
            return torch.tanh(x)

        @staticmethod
        def sigmoid(x):
            return torch.sigmoid(x)

        @staticmethod
        def elu(x):
            return torch.nn.functional.elu(x)

        @staticmethod
        def softmax(x):
            return torch.softmax(x, dim=-1)

        _NAMES = ("linear", "relu", "tanh", "sigmoid", "elu", "softmax")

        def get(self, identifier):
            if identifier is None:
                return self.linear
            if callable(identifier):
                return identifier
            if isinstance(identifier, str) and identifier in self._NAMES:
                return getattr(self, identifier)
            raise ValueError(f"Could not interpret activation function identifier: {identifier}")

        def serialize(self, fn):
            if fn is None:
                return None
            return getattr(fn, "__name__", "linear")

        def deserialize(self, name):
            return self.get(name) if name is not None else None

    activations = _Activations()

    class _Passthrough:
        """regularizers / constraints: carried in the config, not applied by the local shim."""

        class Regularizer:
            pass

        Constraint = Regularizer

        @staticmethod
        def get(x):
            return x

        @staticmethod
        def serialize(x):
            return x

        @staticmethod
        def deserialize(x):
            return x

    regularizers = _Passthrough()
    constraints = _Passthrough()

    # ------------------------------------------------------------------ layers
    def _shape_of(x):
        if isinstance(x, (list, tuple)) and len(x) and not isinstance(x[0], (int, type(None))):
            return [_shape_of(t) for t in x]
        if hasattr(x, "shape"):
            return tuple(x.shape)
        return tuple(np.asarray(x).shape)

    _NAME_COUNTS: dict = {}

    class Layer:
        def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
            if kwargs:
                raise TypeError(f"Unrecognized keyword arguments passed to {type(self).__name__}: {sorted(kwargs)}")
            if name is None:
                base = type(self).__name__.lower()
                _NAME_COUNTS[base] = _NAME_COUNTS.get(base, 0) + 1
                name = base if _NAME_COUNTS[base] == 1 else f"{base}_{_NAME_COUNTS[base] - 1}"
            self.name = name
            self.trainable = trainable
            self.built = False
            self._own_weights: list = []
            self._dtype = dtype or "float32"

        @property
        def dtype(self):
            return self._dtype

        @property
        def compute_dtype(self):
            return self._dtype

        @property
        def variable_dtype(self):
            return self._dtype

        def add_weight(self, shape=None, initializer=None, name=None, trainable=True, dtype=None,
                       regularizer=None, constraint=None):
            init = initializers.get(initializer) if initializer is not None else initializers.get("zeros")
            data = init(tuple(shape)).to(torch.float32).to(default_device())
            p = torch.nn.Parameter(data, requires_grad=bool(trainable and self.trainable))
            p.keras_name = name
            self._own_weights.append(p)
            return p

        def build(self, input_shape):
            self.built = True

        def call(self, *args, **kwargs):
            raise NotImplementedError

        def __call__(self, *args, **kwargs):
            if not self.built:
                self.build(_shape_of(args[0]))
                self.built = True
            return self.call(*args, **kwargs)

        def _sublayers(self):
            seen = []
            for v in self.__dict__.values():
                if isinstance(v, Layer):
                    seen.append(v)
                elif isinstance(v, (list, tuple)):
                    seen.extend(t for t in v if isinstance(t, Layer))
            return seen

        @property
        def weights(self):
            ws = list(self._own_weights)
            for sub in self._sublayers():
                ws.extend(sub.weights)
            return ws

        @property
        def trainable_weights(self):
            return [w for w in self.weights if w.requires_grad]

        trainable_variables = trainable_weights

        def get_weights(self):
            return [w.detach().cpu().numpy() for w in self.weights]

        def set_weights(self, ws):
            mine = self.weights
            if len(ws) != len(mine):
                raise ValueError(f"expected {len(mine)} weight arrays, got {len(ws)}")
            with torch.no_grad():
                for p, w in zip(mine, ws):
                    w = torch.as_tensor(np.asarray(w), dtype=p.dtype)
                    if tuple(w.shape) != tuple(p.shape):
                        raise ValueError(f"weight shape mismatch {tuple(w.shape)} vs {tuple(p.shape)}")
                    p.copy_(w)

        def get_config(self):
            return {"name": self.name, "trainable": self.trainable, "dtype": self._dtype}

        @classmethod
        def from_config(cls, config):
            return cls(**config)

        def compute_output_shape(self, input_shape):
            return input_shape

    class Dense(Layer):
        def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                     bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                     kernel_constraint=None, bias_constraint=None, **kwargs):
            super().__init__(**kwargs)
            self.units = int(units)
            self.activation = activations.get(activation)
            self.use_bias = use_bias
            self.kernel_initializer = initializers.get(kernel_initializer)
            self.bias_initializer = initializers.get(bias_initializer)
            self.kernel = None
            self.bias = None

        def build(self, input_shape):
            self.kernel = self.add_weight((int(input_shape[-1]), self.units), self.kernel_initializer, "kernel")
            if self.use_bias:
                self.bias = self.add_weight((self.units,), self.bias_initializer, "bias")
            self.built = True

        def call(self, x, training=None):
            return apply_dense(self, x)

        def compute_output_shape(self, input_shape):
            return tuple(input_shape[:-1]) + (self.units,)

    class Dropout(Layer):
        def __init__(self, rate, seed=None, **kwargs):
            super().__init__(**kwargs)
            self.rate = float(rate)

        def call(self, x, training=None):
            if training and self.rate > 0:
                return torch.nn.functional.dropout(x, self.rate, training=True)
            return x

    class Sequential(Layer):
        def __init__(self, layers=None, name=None, **kwargs):
            super().__init__(name=name, **kwargs)
            self.layers = list(layers or [])

        def build(self, input_shape):
            shape = tuple(input_shape)
            for lyr in self.layers:
                if not lyr.built:
                    lyr.build(shape)
                    lyr.built = True
                shape = lyr.compute_output_shape(shape)
            self.built = True

        def call(self, x, training=None):
            for lyr in self.layers:
                x = lyr(x, training=training)
            return x
