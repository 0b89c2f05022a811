"""Core of the NumPy-backed jax stand-in (test infrastructure, see ../README.md):
the Array type, the pytree registry, vmap and the lax control-flow primitives."""

from __future__ import annotations

import dataclasses
import functools

import numpy as np


# --------------------------------------------------------------------------------------
# Array: an ndarray subclass with the `.at[idx].set/add/...` functional-update interface
# --------------------------------------------------------------------------------------
class _AtIndexer:
    __slots__ = ("arr",)

    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtRef(self.arr, idx)


def _np_index(idx):
    """Strip the Array subclass from index arrays (plain NumPy fancy indexing)."""
    if isinstance(idx, tuple):
        return tuple(_np_index(i) for i in idx)
    if isinstance(idx, np.ndarray):
        return np.asarray(idx)
    return idx


class _AtRef:
    __slots__ = ("arr", "idx")

    def __init__(self, arr, idx):
        self.arr = arr
        self.idx = _np_index(idx)

    def _copy(self):
        return np.array(self.arr, copy=True)  # JAX keeps the dtype of the updated array

    def set(self, v, **_):
        out = self._copy()
        out[self.idx] = np.asarray(v)
        return asarray(out)

    def add(self, v, **_):
        out = self._copy()
        np.add.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def subtract(self, v, **_):
        out = self._copy()
        np.subtract.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def multiply(self, v, **_):
        out = self._copy()
        np.multiply.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def divide(self, v, **_):
        out = self._copy()
        np.divide.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def min(self, v, **_):
        out = self._copy()
        np.minimum.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def max(self, v, **_):
        out = self._copy()
        np.maximum.at(out, self.idx, np.asarray(v))
        return asarray(out)

    def get(self, **_):
        return asarray(np.asarray(self.arr)[self.idx])

    def apply(self, fn):
        out = self._copy()
        out[self.idx] = np.asarray(fn(asarray(out[self.idx])))
        return asarray(out)


class Array(np.ndarray):
    """ndarray subclass standing in for jax.Array."""

    __array_priority__ = 100.0
    weak_type = False

    @property
    def at(self):
        return _AtIndexer(self)

    def block_until_ready(self):
        return self

    def __hash__(self):  # jax arrays are unhashable too, but dataclass eq/hash helpers may probe
        raise TypeError("unhashable type: 'Array'")

    def __getitem__(self, idx):
        out = np.ndarray.__getitem__(self, _np_index(idx))
        if not isinstance(out, np.ndarray):
            out = np.asarray(out).view(Array)  # 0-d array instead of a NumPy scalar, like JAX
        return out

    # comparisons / arithmetic producing NumPy scalars are turned back into 0-d Arrays
    def __array_wrap__(self, obj, context=None, return_scalar=False):
        out = np.asarray(obj).view(Array)
        return out

    def astype(self, dtype, *a, **k):
        if dtype is float:
            dtype = np.float64
        elif dtype is int:
            dtype = np.int64
        return np.ndarray.astype(self, dtype, *a, **k)

    def __bool__(self):
        return bool(np.asarray(self))

    def __index__(self):
        return int(np.asarray(self))

    def __int__(self):
        return int(np.asarray(self))

    def __float__(self):
        return float(np.asarray(self))


def asarray(x, dtype=None):
    if dtype is float:
        dtype = np.float64
    elif dtype is int:
        dtype = np.int64
    elif dtype is bool:
        dtype = np.bool_
    a = np.asarray(x, dtype=dtype)
    if a.dtype == object:
        raise TypeError(f"cannot convert {type(x)} to an array")
    return a.view(Array)


def wrap_out(x):
    if isinstance(x, np.ndarray):
        return x.view(Array)
    if isinstance(x, np.generic):
        return np.asarray(x).view(Array)
    if isinstance(x, tuple):
        if hasattr(x, "_fields"):
            return type(x)(*[wrap_out(e) for e in x])
        return tuple(wrap_out(e) for e in x)
    if isinstance(x, list):
        return [wrap_out(e) for e in x]
    return x


def wrap_fn(fn):
    @functools.wraps(fn)
    def inner(*args, **kwargs):
        if "dtype" in kwargs:
            if kwargs["dtype"] is float:
                kwargs["dtype"] = np.float64
            elif kwargs["dtype"] is int:
                kwargs["dtype"] = np.int64
        return wrap_out(fn(*args, **kwargs))

    return inner


# --------------------------------------------------------------------------------------
# pytrees
# --------------------------------------------------------------------------------------
_REGISTRY: dict[type, tuple] = {}


def register_pytree_node(cls, flatten, unflatten):
    _REGISTRY[cls] = (flatten, unflatten)


def register_pytree_node_class(cls):
    register_pytree_node(cls, lambda o: o.tree_flatten(), lambda aux, ch: cls.tree_unflatten(aux, ch))
    return cls


class PyTreeDef:
    def __init__(self, kind, aux, children):
        self.kind, self.aux, self.children = kind, aux, children

    @property
    def num_leaves(self):
        if self.kind == "leaf":
            return 1
        if self.kind == "none":
            return 0
        return sum(c.num_leaves for c in self.children)

    def _key(self):
        aux = self.aux
        try:
            hash(aux)
        except TypeError:
            aux = repr(aux)
        return (self.kind if not isinstance(self.kind, type) else self.kind.__qualname__, aux, tuple(c._key() for c in self.children))

    def __eq__(self, other):
        return isinstance(other, PyTreeDef) and self._key() == other._key()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        if self.kind == "leaf":
            return "*"
        if self.kind == "none":
            return "None"
        name = self.kind if isinstance(self.kind, str) else self.kind.__name__
        return f"{name}[{self.aux}]({', '.join(map(repr, self.children))})"

    def unflatten(self, leaves):
        it = iter(leaves)
        return _unflatten(self, it)


def _flatten(tree, leaves, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        leaves.append(tree)
        return PyTreeDef("leaf", None, ())
    if tree is None:
        return PyTreeDef("none", None, ())
    t = type(tree)
    if t in _REGISTRY:
        children, aux = _REGISTRY[t][0](tree)
        return PyTreeDef(t, aux, tuple(_flatten(c, leaves, is_leaf) for c in children))
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return PyTreeDef(t, "namedtuple", tuple(_flatten(c, leaves, is_leaf) for c in tree))
    if t is tuple:
        return PyTreeDef("tuple", None, tuple(_flatten(c, leaves, is_leaf) for c in tree))
    if t is list:
        return PyTreeDef("list", None, tuple(_flatten(c, leaves, is_leaf) for c in tree))
    if t is dict:
        keys = tuple(sorted(tree.keys()))
        return PyTreeDef("dict", keys, tuple(_flatten(tree[k], leaves, is_leaf) for k in keys))
    leaves.append(tree)
    return PyTreeDef("leaf", None, ())


def _unflatten(td, it):
    if td.kind == "leaf":
        return next(it)
    if td.kind == "none":
        return None
    ch = [_unflatten(c, it) for c in td.children]
    if td.kind == "tuple":
        return tuple(ch)
    if td.kind == "list":
        return list(ch)
    if td.kind == "dict":
        return dict(zip(td.aux, ch))
    if td.aux == "namedtuple" and td.kind not in _REGISTRY:
        return td.kind(*ch)
    return _REGISTRY[td.kind][1](td.aux, ch)


def tree_flatten(tree, is_leaf=None):
    leaves = []
    td = _flatten(tree, leaves, is_leaf)
    return leaves, td


def tree_unflatten(treedef, leaves):
    return treedef.unflatten(leaves)


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = []
    for r in rest:
        # `rest` trees may have the first tree as a prefix-compatible structure
        lv = _flatten_up_to(td, r)
        others.append(lv)
    return td.unflatten([f(*xs) for xs in zip(leaves, *others)])


def _flatten_up_to(td, tree):
    """Flatten `tree` down to the leaves of `td` (subtrees at leaf positions are kept whole)."""
    out = []

    def rec(d, t):
        if d.kind == "leaf":
            out.append(t)
            return
        if d.kind == "none":
            return
        if d.kind == "tuple" or d.kind == "list":
            ch = list(t)
        elif d.kind == "dict":
            ch = [t[k] for k in d.aux]
        elif d.kind in _REGISTRY:
            ch = list(_REGISTRY[d.kind][0](t)[0])
        else:
            ch = list(t)
        if len(ch) != len(d.children):
            raise ValueError("pytree structure mismatch")
        for c, x in zip(d.children, ch):
            rec(c, x)

    rec(td, tree)
    return out


def tree_reduce(f, tree, initializer=None):
    leaves = tree_leaves(tree)
    return functools.reduce(f, leaves) if initializer is None else functools.reduce(f, leaves, initializer)


def tree_all(tree):
    return all(tree_leaves(tree))


def ravel_pytree(tree):
    leaves, td = tree_flatten(tree)
    arrs = [np.asarray(leaf) for leaf in leaves]
    shapes = [a.shape for a in arrs]
    dtypes = [a.dtype for a in arrs]
    sizes = [a.size for a in arrs]
    flat = np.concatenate([a.ravel() for a in arrs]) if arrs else np.zeros(0)

    def unravel(v):
        v = np.asarray(v)
        out, k = [], 0
        for sh, dt, sz in zip(shapes, dtypes, sizes):
            out.append(asarray(v[k:k + sz].reshape(sh).astype(dt)))
            k += sz
        return td.unflatten(out)

    return asarray(flat), unravel


# --------------------------------------------------------------------------------------
# transformations
# --------------------------------------------------------------------------------------
def jit(fun=None, **_kw):
    if fun is None:
        return lambda f: f
    return fun


def _broadcast_prefix(prefix, tree):
    """Expand an in_axes prefix (int / None / tuple / list / dict) to one entry per leaf of `tree`."""
    leaves, td = tree_flatten(tree)
    if prefix is None or isinstance(prefix, int):
        return [prefix] * len(leaves)
    out = []

    def rec(p, t):
        if p is None or isinstance(p, int):
            out.extend([p] * len(tree_leaves(t)))
            return
        if isinstance(p, (tuple, list)):
            if len(p) != len(t):
                raise ValueError("in_axes prefix does not match the arguments")
            for pp, tt in zip(p, t):
                rec(pp, tt)
            return
        if isinstance(p, dict):
            for k in sorted(t.keys()):
                rec(p[k], t[k])
            return
        if type(p) in _REGISTRY:
            pch = _REGISTRY[type(p)][0](p)[0]
            tch = _REGISTRY[type(t)][0](t)[0]
            for pp, tt in zip(pch, tch):
                rec(pp, tt)
            return
        raise TypeError(f"unsupported in_axes entry {p!r}")

    rec(prefix, tree)
    return out


def vmap(fun, in_axes=0, out_axes=0, **_kw):
    """jax.vmap as a Python loop over the mapped axis (keyword arguments map over axis 0)."""

    @functools.wraps(fun)
    def mapped(*args, **kwargs):
        ia = in_axes
        if isinstance(ia, (tuple, list)) and len(ia) != len(args):
            raise ValueError(f"vmap in_axes {ia} does not match {len(args)} positional arguments")
        arg_leaves, arg_td = tree_flatten(tuple(args))
        arg_axes = _broadcast_prefix(tuple(ia) if isinstance(ia, (tuple, list)) else ia, tuple(args))
        kw_leaves, kw_td = tree_flatten(kwargs)
        kw_axes = [0] * len(kw_leaves)
        n = None
        for leaf, ax in zip(arg_leaves + kw_leaves, arg_axes + kw_axes):
            if ax is None:
                continue
            m = np.shape(leaf)[ax]
            if n is None:
                n = m
            elif n != m:
                raise ValueError(f"vmap got inconsistent sizes for the mapped axis: {n} vs {m}")
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")

        def take(leaf, ax, i):
            if ax is None:
                return leaf
            return asarray(np.take(np.asarray(leaf), i, axis=ax))

        outs = []
        for i in range(n):
            a_i = arg_td.unflatten([take(leaf, ax, i) for leaf, ax in zip(arg_leaves, arg_axes)])
            k_i = kw_td.unflatten([take(leaf, ax, i) for leaf, ax in zip(kw_leaves, kw_axes)])
            outs.append(fun(*a_i, **k_i))
        if n == 0:
            raise ValueError("vmap over an empty axis is not supported by the stand-in")
        out_leaves0, out_td = tree_flatten(outs[0])
        all_leaves = [tree_flatten(o)[0] for o in outs]
        oaxes = _broadcast_prefix(out_axes if not isinstance(out_axes, list) else tuple(out_axes), outs[0])
        stacked = []
        for j in range(len(out_leaves0)):
            col = [np.asarray(l[j]) for l in all_leaves]
            ax = oaxes[j]
            stacked.append(asarray(np.stack(col, axis=0 if ax is None else ax)) if ax is not None else asarray(col[0]))
        return out_td.unflatten(stacked)

    return mapped


class custom_jvp:
    def __init__(self, fun, nondiff_argnums=()):
        self.fun = fun
        functools.update_wrapper(self, fun)

    def defjvp(self, jvp):
        self.jvp = jvp
        return jvp

    def __call__(self, *a, **k):
        return self.fun(*a, **k)


# --------------------------------------------------------------------------------------
# lax
# --------------------------------------------------------------------------------------
def _to_bool(p):
    return bool(np.asarray(p))


class lax:
    @staticmethod
    def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
        if xs is None:
            n = int(length)
            xs_leaves, xs_td = [], None
        else:
            xs_leaves, xs_td = tree_flatten(xs)
            n = int(np.shape(xs_leaves[0])[0]) if xs_leaves else int(length)
        carry = init
        ys = []
        order = range(n - 1, -1, -1) if reverse else range(n)
        for i in order:
            x = None if xs_td is None else xs_td.unflatten([asarray(np.asarray(l)[i]) for l in xs_leaves])
            carry, y = f(carry, x)
            ys.append(y)
        if reverse:
            ys = ys[::-1]
        if n == 0:
            return carry, None
        y_leaves0, y_td = tree_flatten(ys[0])
        cols = [tree_flatten(y)[0] for y in ys]
        stacked = [asarray(np.stack([np.asarray(c[j]) for c in cols], axis=0)) for j in range(len(y_leaves0))]
        return carry, y_td.unflatten(stacked)

    @staticmethod
    def cond(pred, true_fun, false_fun, *operands, operand=None):
        if operand is not None and not operands:
            operands = (operand,)
        return true_fun(*operands) if _to_bool(pred) else false_fun(*operands)

    @staticmethod
    def switch(index, branches, *operands):
        i = int(np.clip(int(np.asarray(index)), 0, len(branches) - 1))
        return branches[i](*operands)

    @staticmethod
    def select(pred, on_true, on_false):
        return asarray(np.where(np.asarray(pred), np.asarray(on_true), np.asarray(on_false)))

    @staticmethod
    def select_n(which, *cases):
        w = np.asarray(which)
        if w.dtype == np.bool_:
            w = w.astype(np.int64)
        return asarray(np.choose(w, [np.asarray(c) for c in cases]))

    @staticmethod
    def while_loop(cond_fun, body_fun, init_val):
        val = init_val
        while _to_bool(cond_fun(val)):
            val = body_fun(val)
        return val

    @staticmethod
    def fori_loop(lower, upper, body_fun, init_val):
        val = init_val
        for i in range(int(lower), int(upper)):
            val = body_fun(i, val)
        return val

    @staticmethod
    def stop_gradient(x):
        return x

    @staticmethod
    def custom_linear_solve(matvec, b, solve, transpose_solve=None, symmetric=False, has_aux=False):
        """The solution of matvec(x) = b itself: the operator is materialised column by column and
        solved directly (minimum-norm where it is singular), instead of running the caller's
        iterative `solve` -- the only call site (rbda/contacts/relaxed_rigid.py:497-505) passes an
        L-BFGS loop over optax there, whose fixed point is this solution (README.md)."""
        rhs = np.asarray(b, dtype=float)
        n = rhs.shape[0]
        A = np.stack([np.asarray(matvec(asarray(np.eye(n)[:, k])), dtype=float) for k in range(n)], axis=1) if n else np.zeros((0, 0))
        x = np.linalg.lstsq(A, rhs, rcond=1e-13)[0] if n else rhs
        return (asarray(x), None) if has_aux else asarray(x)

    @staticmethod
    def dynamic_slice(operand, start_indices, slice_sizes):
        a = np.asarray(operand)
        idx = tuple(slice(int(np.clip(int(s), 0, a.shape[k] - n)), int(np.clip(int(s), 0, a.shape[k] - n)) + n)
                    for k, (s, n) in enumerate(zip(start_indices, slice_sizes)))
        return asarray(a[idx])

    @staticmethod
    def dynamic_update_slice(operand, update, start_indices):
        a = np.array(operand, copy=True)
        u = np.asarray(update)
        idx = tuple(slice(int(s), int(s) + n) for s, n in zip(start_indices, u.shape))
        a[idx] = u
        return asarray(a)


def dataclass_fields_split(cls):
    """(dynamic field names, static field names) of a pytree dataclass: a field is static if
    its annotation is wrapped in Static[...] (jax_dataclasses convention)."""
    dyn, sta = [], []
    ann = {}
    for klass in reversed(cls.__mro__):
        ann.update(getattr(klass, "__annotations__", {}))
    for f in dataclasses.fields(cls):
        a = ann.get(f.name, f.type)
        s = a if isinstance(a, str) else repr(a)
        is_static = ("Static[" in s) or ("JDC_STATIC_MARKER" in s) or bool(f.metadata.get("jdc_static", False))
        (sta if is_static else dyn).append(f.name)
    return dyn, sta
