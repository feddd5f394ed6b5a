from . import _Anything


def __getattr__(name):
    return _Anything()
