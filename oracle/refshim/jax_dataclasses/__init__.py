"""Stand-in for `jax_dataclasses` (test infrastructure only, see ../README.md): frozen
dataclasses registered as pytrees, `Static[...]` annotations select the static fields."""
import dataclasses
import typing

from jax import _core

from . import _copy_and_mutate
from ._copy_and_mutate import _Mutability, copy_and_mutate  # noqa: F401

field = dataclasses.field
replace = dataclasses.replace
asdict = dataclasses.asdict


class Static:
    def __class_getitem__(cls, item):
        return typing.Annotated[item, "JDC_STATIC_MARKER"]


def static_field(*args, **kwargs):
    kwargs["metadata"] = {**kwargs.get("metadata", {}), "jdc_static": True}
    return dataclasses.field(*args, **kwargs)


def _setattr(self, name, value):
    if self.__mutability__ is _Mutability.FROZEN:
        raise dataclasses.FrozenInstanceError(
            f"Dataclass registered as pytree is immutable, cannot assign to field {name!r}")
    object.__setattr__(self, name, value)


def _delattr(self, name):
    if self.__mutability__ is _Mutability.FROZEN:
        raise dataclasses.FrozenInstanceError(f"cannot delete field {name!r}")
    object.__delattr__(self, name)


def pytree_dataclass(cls=None, **kwargs):
    kwargs.setdefault("frozen", True)
    assert kwargs["frozen"] is True

    def wrap(c):
        c = dataclasses.dataclass(c, **kwargs)
        c.__mutability__ = _Mutability.FROZEN
        c.__setattr__ = _setattr
        c.__delattr__ = _delattr
        dyn, sta = _core.dataclass_fields_split(c)

        def flatten(obj):
            return [getattr(obj, k) for k in dyn], tuple(getattr(obj, k) for k in sta)

        def unflatten(aux, children):
            obj = c.__new__(c)
            for k, v in zip(dyn, children):
                object.__setattr__(obj, k, v)
            for k, v in zip(sta, aux):
                object.__setattr__(obj, k, v)
            return obj

        _core.register_pytree_node(c, flatten, unflatten)
        c.__jdc_dynamic__, c.__jdc_static__ = tuple(dyn), tuple(sta)
        return c

    return wrap if cls is None else wrap(cls)
