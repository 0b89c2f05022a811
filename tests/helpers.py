"""Shared helpers for the parity tests: bridge between the product objects
(``jaxsim_b200.api``) and the oracle's (``oracle.jaxsim_oracle``)."""

from __future__ import annotations

import numpy as np

import jaxsim_b200.api as js
from jaxsim_b200 import models
from jaxsim_b200.rbda.contacts import RelaxedRigidContacts, RigidContacts, SoftContacts
from oracle import jaxsim_oracle as O

# north_star tolerances: 1e-5 rel (fp64) / 1e-3 rel (fp32)
RTOL = {"float64": 1e-5, "float32": 1e-3}


def build_model(name: str, **kw):
    return js.model.JaxSimModel.build_from_model_description(models.urdf(name), **kw)


def build_model_for_case(case: dict):
    """Product model of a golden-fixture case (tests/golden/cases.py)."""
    from jaxsim_b200.rbda.actuation import ActuationParams
    from jaxsim_b200.rbda.contacts import RelaxedRigidContactsParams, RigidContactsParams, SoftContactsParams

    kinds = {"soft": (SoftContacts, SoftContactsParams), "rigid": (RigidContacts, RigidContactsParams),
             "relaxed": (RelaxedRigidContacts, RelaxedRigidContactsParams)}
    cm_cls, cp_cls = kinds[case["contact"]]
    cm = cm_cls.build()
    cp = cp_cls.build(**case["contact_params"]) if case["contact_params"] else None
    ap = ActuationParams(**case["actuation"]) if case["actuation"] else None
    integ = {"semi_implicit_euler": js.model.IntegratorType.SemiImplicitEuler, "rk4": js.model.IntegratorType.RungeKutta4,
             "rk4fast": js.model.IntegratorType.RungeKutta4Fast}[case["integrator"]]
    model = build_model(case["model"], time_step=case["time_step"], contact_model=cm, contact_params=cp,
                        actuation_params=ap, integrator=integ)
    return with_constraints(model, case.get("constraints"))


def with_constraints(model, spec):
    """The model with the weld constraints ``[(frame_1, frame_2, K_P, K_D)]`` attached the way the reference's tests do
    (``tests/test_simulations.py:429-440``)."""
    if not spec:
        return model
    cmap = js.kin_dyn_parameters.ConstraintMap()
    for f1, f2, kp, kd in spec:
        cmap = cmap.add_constraint(model, js.frame.name_to_idx(model, frame_name=f1), js.frame.name_to_idx(model, frame_name=f2),
                                   js.kin_dyn_parameters.ConstraintType.Weld, K_P=kp, K_D=kd)
    with model.editable(validate=False) as model:
        model.kin_dyn_parameters.constraints = cmap
    return model


def oracle_model(model) -> O.OracleModel:
    prm = model.contact_params
    soft = isinstance(model.contact_model, SoftContacts)
    rigid = isinstance(model.contact_model, RigidContacts)
    relaxed = isinstance(model.contact_model, RelaxedRigidContacts)
    kw = dict(K=prm.K, D=prm.D, mu=prm.mu)
    if soft:
        kw.update(p=prm.p, q=prm.q)
    if rigid:
        kw.update(regularization_delassus=model.contact_model.regularization_delassus)
    if relaxed:
        kw.update(relaxed=dict(time_constant=prm.time_constant, damping_coefficient=prm.damping_coefficient, d_min=prm.d_min,
                               d_max=prm.d_max, width=prm.width, midpoint=prm.midpoint, power=prm.power))
    return O.OracleModel(
        kin_dyn_parameters=model.kin_dyn_parameters, floating_base=model.floating_base(),
        time_step=model.time_step, gravity=model.gravity, terrain_height=model.terrain.height(),
        contact_model="soft" if soft else ("rigid" if rigid else ("relaxed" if relaxed else "none")),
        torque_max=model.actuation_params.torque_max, omega_th=model.actuation_params.omega_th,
        omega_max=model.actuation_params.omega_max, enable_friction=model.actuation_params.enable_friction,
        **kw,
    )


def to_product(model, od: O.OracleData, dtype, device, velocity_representation=None):
    """OracleData (numpy) -> JaxSimModelData (torch on device) holding the SAME numbers."""
    import torch

    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=device)  # noqa: E731
    vr = js.common.VelRepr.Inertial if velocity_representation is None else velocity_representation
    cs = {}
    if od.tangential_deformation is not None and isinstance(model.contact_model, SoftContacts):
        cs["tangential_deformation"] = t(od.tangential_deformation)
    # velocities are given inertial-fixed: build in Inertial representation then relabel
    d = js.data.JaxSimModelData.build(
        model, base_position=t(od.base_position), base_quaternion=t(od.base_quaternion),
        joint_positions=t(od.joint_positions), joint_velocities=t(od.joint_velocities),
        base_linear_velocity=t(od.base_linear_velocity), base_angular_velocity=t(od.base_angular_velocity),
        contact_state=cs, velocity_representation=js.common.VelRepr.Inertial,
        batch_size=od.base_position.shape[0], dtype=dtype, device=device,
    )
    d.velocity_representation = vr
    return d


LEAVES = [
    ("joint_positions", "_joint_positions"), ("joint_velocities", "_joint_velocities"),
    ("base_quaternion", "_base_quaternion"), ("base_linear_velocity", "_base_linear_velocity"),
    ("base_angular_velocity", "_base_angular_velocity"), ("base_position", "_base_position"),
    ("base_transform", "_base_transform"), ("joint_transforms", "_joint_transforms"),
    ("link_transforms", "_link_transforms"), ("link_velocities", "_link_velocities"),
]


def rel_err(x: np.ndarray, ref: np.ndarray) -> float:
    """max |x - ref| / max(|ref|, tiny): relative to the scale of the leaf."""
    if ref.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(ref))), 1e-12)
    return float(np.max(np.abs(x.astype(np.float64) - ref.astype(np.float64)))) / scale


def _per_env(a: np.ndarray, batched: bool) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(a.shape[0], -1) if (batched and a.ndim >= 1) else a.reshape(1, -1)


def elementwise_violation(got: np.ndarray, ref: np.ndarray, rtol: float, batched: bool, floor: float = 0.0) -> float:
    """max over the entries of |got - ref| / (rtol * (|ref| + s_env)), where s_env is the largest magnitude of the
    leaf IN THE SAME ENVIRONMENT (at least ``floor``).  <= 1 means every entry satisfies

        |got - ref| <= rtol * |ref| + atol,     atol = rtol * s_env.

    Why atol is the per-environment leaf scale and not zero: every entry of a leaf (a rotation entry, a component
    of a velocity, a row of a 6x6 adjoint) is a sum of products of quantities as large as the biggest entries of
    that leaf for that environment, so its rounding error is eps * s_env whatever its own magnitude -- an entry
    that cancels to 1e-9 next to entries of 1 cannot be correct to 1e-3 of ITSELF in float32.  What the criterion
    no longer allows is a small environment hiding behind a large one elsewhere in the batch, or an
    error of rtol * (batch-wide maximum) on every entry."""
    g, r = _per_env(got, batched), _per_env(ref, batched)
    if r.size == 0:
        return 0.0
    s_env = np.maximum(np.max(np.abs(r), axis=1, keepdims=True), max(floor, 1e-300))
    return float(np.max(np.abs(g - r) / (rtol * (np.abs(r) + s_env))))


def compare_data(pd, od: O.OracleData, rtol: float, what="", floors: dict | None = None) -> dict:
    """Primary criterion: elementwise, per environment (``elementwise_violation``) on every leaf.
    ``floors`` maps a leaf name to a minimal scale (used where the expected value is ~0, e.g. velocities after a
    perfectly inelastic impact: the error is then measured against the scale of the same quantity before the
    step).  Returns the leaf-max-normalised relative errors (the round-1 metric) as a secondary report."""
    errs, viol = {}, {}
    batched = pd._base_quaternion.dim() == 2
    for oname, pname in LEAVES:
        ref = getattr(od, oname)
        got = getattr(pd, pname)
        if got is None:
            continue
        got = got.detach().cpu().numpy()
        fl = float(floors[oname]) if (floors and oname in floors) else 0.0
        e = rel_err(got, ref)
        if fl and ref.size:
            scale = max(float(np.max(np.abs(ref))), 1e-12)
            e = e * scale / max(scale, fl)
        errs[oname] = e
        viol[oname] = elementwise_violation(got, ref, rtol, batched, fl)
    if "tangential_deformation" in pd.contact_state and od.tangential_deformation is not None:
        ref = od.tangential_deformation
        got = pd.contact_state["tangential_deformation"].detach().cpu().numpy()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        if ref.size == 0:
            ref = got = np.zeros(1)
        scale = max(float(np.max(np.abs(ref))), 1e-6)
        errs["tangential_deformation"] = float(np.max(np.abs(got - ref))) / scale
        # The deformation state integrates the tangential velocity of the points (O(0.1) m/s) over dt = 1e-3 s: its
        # natural scale is 1e-4 m whatever an individual point currently holds (a point that has just touched the
        # ground holds ~1e-7 m), and its float32 error is set by the cancellation in delta = h - p_z (|p| ~ 1 m),
        # not by the point's own value.  Scale floor 1e-4 m: atol = 1e-7 m (float32) / 1e-9 m (float64).
        viol["tangential_deformation"] = elementwise_violation(got, ref, rtol, batched, 1e-4)
    bad = {k: v for k, v in viol.items() if not (v <= 1.0)}
    assert not bad, (f"{what}: entries outside |x - ref| <= {rtol} * (|ref| + per-environment leaf scale) by the factor "
                     f"{bad} (leaf-max relative errors: {errs})")
    return errs
