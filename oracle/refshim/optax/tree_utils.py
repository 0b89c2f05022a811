def _unavailable(*_a, **_k):
    raise NotImplementedError("optax is not available in the stand-in")


tree_get = tree_norm = tree_l2_norm = _unavailable
