from ._core import tree_all as all, tree_flatten as flatten, tree_leaves as leaves, tree_map as map  # noqa: A004,F401
from ._core import tree_reduce as reduce, tree_structure as structure, tree_unflatten as unflatten  # noqa: F401
