from . import _Any


class Axes(_Any):
    pass


class Axis(_Any):
    pass


class Figure(_Any):
    pass


class Rectangle(_Any):
    pass


class FuncFormatter(_Any):
    pass


class EngFormatter(_Any):
    pass


def __getattr__(name):
    return _Any()
