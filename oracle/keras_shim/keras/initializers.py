import math

import torch


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


class GlorotUniform(Initializer):
    def __call__(self, shape, dtype=None):
        shape = tuple(shape)
        if len(shape) < 1:
            fan_in = fan_out = 1
        elif len(shape) == 1:
            fan_in = fan_out = shape[0]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:
            rf = 1
            for d in shape[:-2]:
                rf *= d
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        lim = math.sqrt(6.0 / max(1.0, (fan_in + fan_out)))
        return (torch.rand(shape) * 2 - 1) * lim


_BY_NAME = {"zeros": Zeros, "ones": Ones, "glorot_uniform": GlorotUniform, "constant": Constant}


def get(identifier):
    if identifier is None:
        return None
    if isinstance(identifier, Initializer):
        return identifier
    if isinstance(identifier, str):
        return _BY_NAME[identifier.lower()]()
    if isinstance(identifier, dict):
        return deserialize(identifier)
    if callable(identifier):
        return identifier
    raise ValueError(identifier)


def serialize(init):
    if init is None:
        return None
    for k, v in _BY_NAME.items():
        if type(init) is v:
            return {"class_name": k, "config": dict(getattr(init, "__dict__", {}))}
    return init


def deserialize(cfg):
    if cfg is None or isinstance(cfg, Initializer):
        return cfg
    if isinstance(cfg, str):
        return get(cfg)
    return _BY_NAME[cfg["class_name"]](**cfg.get("config", {}))
