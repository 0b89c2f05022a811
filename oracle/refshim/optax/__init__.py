"""Name-only stand-in for `optax` (only the out-of-scope RelaxedRigid contact model uses it)."""
from . import tree_utils  # noqa: F401


class GradientTransformationExtraArgs: pass
class OptState: pass


def _unavailable(*_a, **_k):
    raise NotImplementedError("optax is not available in the stand-in")


lbfgs = value_and_grad_from_state = apply_updates = _unavailable
