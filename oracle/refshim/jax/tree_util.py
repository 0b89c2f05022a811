from ._core import (PyTreeDef, register_pytree_node, register_pytree_node_class, tree_all, tree_flatten, tree_leaves,  # noqa: F401
                    tree_map, tree_reduce, tree_structure, tree_unflatten)
