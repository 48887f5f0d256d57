from oracle.keras_ops import floatx  # noqa: F401


def epsilon():
    return 1e-7
