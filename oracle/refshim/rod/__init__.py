"""Name-only stand-in for `rod` (the reference's URDF/SDF front end).  The goldens build the
reference's `ModelDescription` directly (tests/golden/make_goldens.py), so nothing here runs."""
import enum

from . import urdf  # noqa: F401


class _Unavailable:
    def __init__(self, *a, **k):
        raise NotImplementedError("rod is not available: build a ModelDescription directly")


class Model(_Unavailable): pass
class Sdf(_Unavailable): pass
class Link(_Unavailable): pass
class Joint(_Unavailable): pass
class Pose(_Unavailable): pass
class Inertia(_Unavailable): pass
class Box(_Unavailable): pass
class Sphere(_Unavailable): pass
class Cylinder(_Unavailable): pass
class Mesh(_Unavailable): pass
class Collision(_Unavailable): pass
class Frame(_Unavailable): pass


class FrameConvention(enum.Enum):
    Urdf = enum.auto()
    Sdf = enum.auto()
    World = enum.auto()
    Model = enum.auto()
