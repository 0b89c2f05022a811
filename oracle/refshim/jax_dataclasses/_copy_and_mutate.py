import dataclasses
import enum

from jax import _core


class _Mutability(enum.Enum):
    FROZEN = enum.auto()
    MUTABLE = enum.auto()
    MUTABLE_NO_VALIDATION = enum.auto()


def _mark_mutable(obj, mutable, visited):
    if id(obj) in visited:
        return
    visited.add(id(obj))
    if isinstance(obj, (list, tuple)):
        for c in obj:
            _mark_mutable(c, mutable, visited)
    elif isinstance(obj, dict):
        for c in obj.values():
            _mark_mutable(c, mutable, visited)
    elif dataclasses.is_dataclass(obj) and not isinstance(obj, type) and type(obj) in _core._REGISTRY:
        object.__setattr__(obj, "__mutability__", mutable)
        for f in dataclasses.fields(obj):
            try:
                _mark_mutable(getattr(obj, f.name), mutable, visited)
            except AttributeError:
                pass


def copy_and_mutate(pytree, validate=True):
    import contextlib

    @contextlib.contextmanager
    def ctx():
        out = _core.tree_map(lambda x: x, pytree)
        _mark_mutable(out, _Mutability.MUTABLE if validate else _Mutability.MUTABLE_NO_VALIDATION, set())
        yield out
        _mark_mutable(out, _Mutability.FROZEN, set())

    return ctx()
