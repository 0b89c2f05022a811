"""Contact-model selectors and parameters.

Only the data the kernel needs: the model classes are tags (the reference's classes carry
the algorithms, ``src/jaxsim/rbda/contacts/soft.py:126-444``; here the algorithm lives in the
CUDA kernel and the tag selects it, like the reference's ``Static`` field selects a trace).
"""

from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class SoftContactsParams:
    """``SoftContactsParams`` (``rbda/contacts/soft.py:24-123``)."""

    K: float = 1e6
    D: float = 2000.0
    mu: float = 0.5
    p: float = 0.5
    q: float = 0.5

    @classmethod
    def build(cls, *, K=1e6, D=2_000, mu=0.5, p=0.5, q=0.5, **kwargs) -> "SoftContactsParams":
        return cls(K=float(K), D=float(D), mu=float(mu), p=float(p), q=float(q))

    def valid(self) -> bool:
        return all(v >= 0.0 for v in (self.K, self.D, self.mu, self.p, self.q))


@dataclasses.dataclass(frozen=True)
class SoftContacts:
    """Tag for the Hunt/Crossley soft-contact model (``rbda/contacts/soft.py:126-444``)."""

    _parameters_class = SoftContactsParams

    @classmethod
    def build(cls, **kwargs) -> "SoftContacts":
        return cls()


@dataclasses.dataclass(frozen=True)
class RigidContacts:
    """Tag for ``rbda/contacts/rigid.py`` -- not implemented by the kernel yet: building a
    model with it raises ``NotImplementedError`` (DESIGN.md "out of scope")."""

    @classmethod
    def build(cls, **kwargs) -> "RigidContacts":
        return cls()
