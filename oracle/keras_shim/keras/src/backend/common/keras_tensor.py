class KerasTensor:  # annotation-only stand-in
    pass
