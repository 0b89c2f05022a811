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
class RigidContactsParams:
    """``RigidContactsParams`` (``rbda/contacts/rigid.py:27-96``): friction coefficient and the
    Baumgarte gains of the contact constraint (both zero by default)."""

    mu: float = 0.5
    K: float = 0.0
    D: float = 0.0

    @classmethod
    def build(cls, *, mu=None, K=None, D=None, **kwargs) -> "RigidContactsParams":
        return cls(mu=float(0.5 if mu is None else mu), K=float(0.0 if K is None else K),
                   D=float(0.0 if D is None else D))

    def valid(self) -> bool:
        return all(v >= 0.0 for v in (self.mu, self.K, self.D))


@dataclasses.dataclass(frozen=True)
class RigidContacts:
    """Tag + static options of the rigid-contact model (``rbda/contacts/rigid.py:99-160``).

    ``solver_options`` is accepted for signature compatibility.  The reference forwards it to
    ``qpax.solve_qp`` (default ``solver_tol=1e-3``); the kernel's interior-point solver always
    iterates to the resolution of its arithmetic (DESIGN.md "rigid contacts"), i.e. it returns
    the optimum the reference's solver approximates."""

    regularization_delassus: float = 1e-6
    solver_options: tuple = (("solver_tol", 1e-3),)

    _parameters_class = RigidContactsParams

    @classmethod
    def build(cls, regularization_delassus=None, solver_options=None, **kwargs) -> "RigidContacts":
        opts = {"solver_tol": 1e-3} | (dict(solver_options) if solver_options is not None else {})
        try:
            hash(tuple(opts.values()))
        except TypeError as exc:
            raise ValueError("The values of the solver options must be hashable.") from exc
        return cls(regularization_delassus=float(1e-6 if regularization_delassus is None else regularization_delassus),
                   solver_options=tuple(opts.items()))


@dataclasses.dataclass(frozen=True)
class RelaxedRigidContactsParams:
    """``RelaxedRigidContactsParams`` (``rbda/contacts/relaxed_rigid.py:30-160``).  ``K`` and ``D``
    are accepted like in the reference, which derives the spring / damper of the reference
    acceleration from the time constant instead (``:585-591`` shadow them)."""

    time_constant: float = 0.02
    damping_coefficient: float = 1.0
    d_min: float = 0.9
    d_max: float = 0.95
    width: float = 0.001
    midpoint: float = 0.5
    power: float = 2.0
    K: float = 0.0
    D: float = 0.0
    mu: float = 0.005

    @classmethod
    def build(cls, *, time_constant=None, damping_coefficient=None, d_min=None, d_max=None, width=None, midpoint=None,
              power=None, K=None, D=None, mu=None, **kwargs) -> "RelaxedRigidContactsParams":
        given = dict(time_constant=time_constant, damping_coefficient=damping_coefficient, d_min=d_min, d_max=d_max,
                     width=width, midpoint=midpoint, power=power, K=K, D=D, mu=mu)
        return cls(**{k: float(v) for k, v in given.items() if v is not None})

    def valid(self) -> bool:
        return (self.time_constant > 0 and self.damping_coefficient > 0 and 0 < self.d_min <= self.d_max <= 1
                and self.width > 0 and 0 < self.midpoint < 1 and self.power > 0 and self.mu >= 0)


@dataclasses.dataclass(frozen=True)
class RelaxedRigidContacts:
    """Tag + static options of the relaxed-rigid contact model (``rbda/contacts/relaxed_rigid.py:163-281``).

    ``solver_options`` is accepted for signature compatibility: the reference minimises
    ``|A x + b|^2`` with L-BFGS (``tol``, ``maxiter``, ``memory_size``); the kernel solves the
    positive-definite system ``A x = -b`` on the active points directly, i.e. it returns the
    fixed point the reference's iteration approaches."""

    solver_options: tuple = (("tol", 1e-6), ("maxiter", 50), ("memory_size", 10))

    _parameters_class = RelaxedRigidContactsParams

    @classmethod
    def build(cls, solver_options=None, **kwargs) -> "RelaxedRigidContacts":
        opts = {"tol": 1e-6, "maxiter": 50, "memory_size": 10} | (dict(solver_options) if solver_options is not None else {})
        try:
            hash(tuple(opts.values()))
        except TypeError as exc:
            raise ValueError("The values of the solver options must be hashable.") from exc
        return cls(solver_options=tuple(opts.items()))
