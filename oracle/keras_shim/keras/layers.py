"""keras.layers subset: Layer, Dense, Dropout, Sequential (+ inert stubs)."""
import numpy as np
import torch

from . import activations as _act
from . import initializers as _init


def _shape_of(x):
    if isinstance(x, (list, tuple)) and len(x) and not isinstance(x[0], (int, type(None))):
        return [_shape_of(t) for t in x]
    if hasattr(x, "shape"):
        return tuple(x.shape)
    return tuple(np.asarray(x).shape)


def _to_tensor(x):
    if isinstance(x, (list, tuple)):
        return type(x)(_to_tensor(t) for t in x)
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x))
    return x


class Layer:
    def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
        if kwargs:
            raise TypeError(f"Unrecognized keyword arguments: {sorted(kwargs)}")
        self.name = name or type(self).__name__.lower()
        self.trainable = trainable
        self.built = False
        self._own_weights = []
        self._dtype = dtype or "float32"

    dtype = property(lambda self: self._dtype)
    compute_dtype = property(lambda self: self._dtype)

    def add_weight(self, shape=None, initializer=None, name=None, trainable=True, dtype=None,
                   regularizer=None, constraint=None):
        init = _init.get(initializer) if initializer is not None else _init.Zeros()
        p = torch.nn.Parameter(init(tuple(shape)).float(), requires_grad=bool(trainable))
        p.keras_name = name
        self._own_weights.append(p)
        return p

    def build(self, input_shape):
        self.built = True

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        # real Keras converts array-likes in the positional inputs to backend tensors first
        args = tuple(_to_tensor(a) for a in args)
        if not self.built:
            self.build(_shape_of(args[0]))
            self.built = True
        return self.call(*args, **kwargs)

    @property
    def weights(self):
        return list(self._own_weights)

    trainable_weights = weights

    def get_weights(self):
        return [w.detach().cpu().numpy() for w in self._own_weights]

    def set_weights(self, ws):
        for p, w in zip(self._own_weights, ws):
            with torch.no_grad():
                p.copy_(torch.as_tensor(np.asarray(w), dtype=p.dtype))

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
        self.activation = _act.get(activation)
        self.use_bias = use_bias
        self.kernel_initializer = _init.get(kernel_initializer)
        self.bias_initializer = _init.get(bias_initializer)
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        self.kernel = self.add_weight((int(input_shape[-1]), self.units), self.kernel_initializer, "kernel")
        if self.use_bias:
            self.bias = self.add_weight((self.units,), self.bias_initializer, "bias")
        self.built = True

    def call(self, x, training=None):
        y = torch.matmul(x, self.kernel)
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y)


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
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
            if isinstance(lyr, Dense):
                shape = shape[:-1] + (lyr.units,)
        self.built = True

    def call(self, x, training=None):
        for lyr in self.layers:
            x = lyr(x, training=training)
        return x

    @property
    def weights(self):
        return [w for lyr in self.layers for w in lyr.weights]


class _Inert(Layer):
    """Stands in for layers the hot path never instantiates (LSTMCell, ...)."""


def __getattr__(name):
    if name[:1].isupper():
        return _Inert
    raise AttributeError(name)
