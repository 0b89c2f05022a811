"""Parameter classes of the contact / actuation models read by the step
(``src/jaxsim/rbda/contacts/*.py``, ``src/jaxsim/rbda/actuation/common.py``)."""
from . import actuation, contacts  # noqa: F401
