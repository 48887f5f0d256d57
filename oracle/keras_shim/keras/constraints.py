class Constraint:  # pragma: no cover - structural stub
    pass


def get(x):
    return x


def serialize(x):
    return x


def deserialize(x):
    return x
