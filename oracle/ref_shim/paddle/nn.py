"""paddle.nn of the NumPy stand-in: only the base class the reference's problem wrappers inherit from."""


class Layer:
    def __init__(self, *args, **kwargs):
        pass
