#!/usr/bin/env python
"""Generate the golden fixtures `tests/golden/*.npz` by RUNNING THE REFERENCE.

    python tests/golden/make_goldens.py [case_id ...]         (build container only)

The reference's own, unmodified sources (`/root/reference/src/jaxsim`) are imported over the
NumPy-backed stand-ins of its third-party dependencies (`oracle/refshim`, see its README for
what that does and does not change) and its public API is called exactly like a user would:

    model = JaxSimModel.build(model_description=...)               api/model.py:239
    data  = JaxSimModelData.build(model, ..., velocity_representation=...)   api/data.py:66
    out   = js.model.step(model, data, link_forces=..., joint_force_references=...)   api/model.py:2601
    js.model.forward_dynamics_aba / inverse_dynamics / free_floating_mass_matrix / js.ode.system_dynamics

one environment at a time (the reference batches with `jax.vmap`, which is the same function
applied per environment).  Inputs are drawn with this repo's seeded generator and STORED in
the fixture; outputs are whatever the reference returned.  Each fixture also carries the
reference's `KinDynParameters` of the model so that the URDF loader can be checked against
the reference's kinematic-graph code (lumping, BFS order, collidable points).

Fixtures are float64 (the reference's default precision).  `/root/reference` does not exist
on the GPU box: only the committed .npz files travel.
"""

from __future__ import annotations

import dataclasses
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests.golden import cases as C  # noqa: E402
from tests.golden import refenv  # noqa: E402

STATE_LEAVES = ["_joint_positions", "_joint_velocities", "_base_quaternion", "_base_linear_velocity",
                "_base_angular_velocity", "_base_position", "_base_transform", "_joint_transforms",
                "_link_transforms", "_link_velocities"]


def draw_inputs(case: dict):
    """Seeded inputs of a case (this repo's restatement of `random_model_data`, plus torques,
    contact state and external forces); returned as plain NumPy arrays."""
    from jaxsim_b200 import models
    from oracle import jaxsim_oracle as O
    from tests import helpers as H

    pm = H.build_model_for_case(case)
    om = H.oracle_model(pm)
    B = case["B"]
    if case["in_contact"] == "mixed":  # half of the batch airborne, half touching the ground
        a = O.random_model_data(om, B // 2, seed=case["seed"], in_contact=False)
        b = O.random_model_data(om, B - B // 2, seed=case["seed"] + 500, in_contact=True)
        cat = lambda f: np.concatenate([getattr(a, f), getattr(b, f)], axis=0)  # noqa: E731
        od = O.data_replace(om, cat("joint_positions"), cat("joint_velocities"), cat("base_quaternion"),
                            cat("base_linear_velocity"), cat("base_angular_velocity"), cat("base_position"))
    else:
        od = O.random_model_data(om, B, seed=case["seed"], in_contact=case["in_contact"])
    rng = np.random.Generator(np.random.Philox(1000 + case["seed"]))
    n, nL = om.dofs(), om.number_of_links()
    nc = len(np.asarray(pm.kin_dyn_parameters.contact_parameters.body))
    inp = dict(
        joint_positions=od.joint_positions, joint_velocities=od.joint_velocities, base_quaternion=od.base_quaternion,
        base_linear_velocity=od.base_linear_velocity, base_angular_velocity=od.base_angular_velocity,
        base_position=od.base_position,
        tau=10.0 * rng.uniform(0, 1, size=(B, n)) if case["tau"] else np.zeros((B, n)),
        tangential_deformation=(1e-4 * rng.uniform(-1, 1, size=(B, nc, 3)) if case["m"] else np.zeros((B, nc, 3))),
    )
    if case["joint_scale"] != 1.0:  # joint angles (and rates) drawn over their range, then scaled towards zero
        inp["joint_positions"] = case["joint_scale"] * inp["joint_positions"]
        inp["joint_velocities"] = case["joint_scale"] * inp["joint_velocities"]
    if case["round32"]:
        inp = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in inp.items()}
    if case["fext"]:
        inp["link_forces"] = rng.uniform(-10, 10, size=(B, nL, 6))
        if case["round32"]:
            inp["link_forces"] = inp["link_forces"].astype(np.float32).astype(np.float64)
    if case["rbda"]:
        inp["joint_accelerations"] = rng.uniform(-5, 5, size=(B, n))
        inp["base_acceleration"] = rng.uniform(-5, 5, size=(B, 6))
    return models.urdf(case["model"]), inp


def kin_dyn_arrays(rm) -> dict:
    """The reference model's parameters as arrays (api/kin_dyn_parameters.py:30-230)."""
    kd = rm.kin_dyn_parameters
    lp, jp, cp, jm = kd.link_parameters, kd.joint_parameters, kd.contact_parameters, kd.joint_model
    out = dict(
        kd_parent_array=np.asarray(kd.parent_array), kd_link_names=np.array(kd.link_names),
        kd_mass=np.asarray(lp.mass), kd_center_of_mass=np.asarray(lp.center_of_mass),
        kd_inertia_elements=np.asarray(lp.inertia_elements),
        kd_lam_H_pre=np.asarray(jm.λ_H_pre), kd_suc_H_i=np.asarray(jm.suc_H_i),
        kd_joint_axis=np.array([np.asarray(a.axis if hasattr(a, "axis") else a).reshape(3) for a in jm.joint_axis]) if len(jm.joint_axis) else np.zeros((0, 3)),
        kd_joint_types=np.asarray(jm.joint_types),
        kd_contact_body=np.asarray(cp.body), kd_contact_point=np.asarray(cp.point), kd_contact_enabled=np.asarray(cp.enabled),
        kd_floating_base=np.asarray(bool(rm.floating_base())),
    )
    fp = kd.frame_parameters
    if len(fp.name) > 0:  # api/kin_dyn_parameters.py:843-917
        out.update(kd_frame_names=np.array(fp.name), kd_frame_body=np.asarray(fp.body), kd_frame_transform=np.asarray(fp.transform))
    if jp is not None:
        out.update(kd_friction_static=np.asarray(jp.friction_static), kd_friction_viscous=np.asarray(jp.friction_viscous),
                   kd_position_limits_min=np.asarray(jp.position_limits_min), kd_position_limits_max=np.asarray(jp.position_limits_max),
                   kd_position_limit_spring=np.asarray(jp.position_limit_spring), kd_position_limit_damper=np.asarray(jp.position_limit_damper))
    return out


def run_case(case: dict) -> dict:
    jaxsim, js = refenv.load()
    VelRepr = jaxsim.VelRepr
    urdf_text, inp = draw_inputs(case)
    rm = refenv.reference_model(urdf_text, contact=case["contact"], contact_params=case["contact_params"],
                                time_step=case["time_step"], integrator=case["integrator"], actuation=case["actuation"])
    if case["constraints"]:  # attached like tests/test_simulations.py:429-440 does
        import jax.numpy as jnp

        cmap = js.kin_dyn_parameters.ConstraintMap()
        for f1, f2, kp, kd in case["constraints"]:
            cmap = cmap.add_constraint(
                model=rm, frame_idx_1=js.frame.name_to_idx(model=rm, frame_name=f1),
                frame_idx_2=js.frame.name_to_idx(model=rm, frame_name=f2),
                constraint_type=js.kin_dyn_parameters.ConstraintType.Weld,
                K_P=None if kp is None else jnp.array([kp]), K_D=None if kd is None else jnp.array([kd]))
        with rm.editable(validate=False) as rm:
            rm.kin_dyn_parameters.constraints = cmap
    vr = {"inertial": VelRepr.Inertial, "mixed": VelRepr.Mixed, "body": VelRepr.Body}[case["velrepr"]]
    B = case["B"]
    soft = case["contact"] == "soft"
    outs: dict[str, list] = {}

    def push(key, val):
        outs.setdefault(key, []).append(np.asarray(val, dtype=float))

    for e in range(B):
        cs = {"tangential_deformation": inp["tangential_deformation"][e]} if soft else None
        data = js.data.JaxSimModelData.build(
            model=rm, base_position=inp["base_position"][e], base_quaternion=inp["base_quaternion"][e],
            joint_positions=inp["joint_positions"][e], joint_velocities=inp["joint_velocities"][e],
            base_linear_velocity=inp["base_linear_velocity"][e], base_angular_velocity=inp["base_angular_velocity"][e],
            contact_state=cs, velocity_representation=VelRepr.Inertial)
        if vr is not VelRepr.Inertial:
            # same state, other representation of the inputs/outputs that are expressed in it
            data = dataclasses.replace(data, velocity_representation=vr)
        lf = inp["link_forces"][e] if case["fext"] else None
        tau = inp["tau"][e] if case["tau"] else None
        if len(rm.kin_dyn_parameters.contact_parameters.body) > 0:
            cpos, cvel = js.contact.collidable_point_kinematics(model=rm, data=data)
            push("cp_position", cpos)
            push("cp_velocity", cvel)
            push("links_in_contact", np.asarray(js.contact.in_contact(model=rm, data=data), dtype=float))
        else:
            push("cp_position", np.zeros((0, 3)))
            push("cp_velocity", np.zeros((0, 3)))
        if case["constraints"]:
            from jaxsim.rbda.kinematic_constraints import compute_constraint_wrenches

            # the wrench pairs of the input state with no other forces (the quantity the step adds to the link forces)
            push("constraint_wrenches_free", compute_constraint_wrenches(model=rm, data=data))
            fi = np.array([js.frame.name_to_idx(model=rm, frame_name=f) for c_ in case["constraints"] for f in c_[:2]])
            push("constraint_frame_transforms", np.stack([js.frame.transform(model=rm, data=data, frame_index=int(i)) for i in fi]))
        new = js.model.step(model=rm, data=data, link_forces=lf, joint_force_references=tau)
        for _ in range(case["rollout"] - 1):  # the outputs below are those of the LAST step
            new = js.model.step(model=rm, data=new, link_forces=lf, joint_force_references=tau)
        for leaf in STATE_LEAVES:
            push("out" + leaf, getattr(new, leaf))
        if soft:
            push("out_tangential_deformation", new.contact_state["tangential_deformation"])
        if case["rbda"]:
            for name, r in (("inertial", VelRepr.Inertial), ("mixed", VelRepr.Mixed), ("body", VelRepr.Body)):
                # JaxSimModelData.build with the base velocity GIVEN in representation r (api/data.py:66-202): what it
                # stores (always inertial-fixed) and what the base_velocity accessor returns (api/data.py:288-312)
                d_b = js.data.JaxSimModelData.build(
                    model=rm, base_position=inp["base_position"][e], base_quaternion=inp["base_quaternion"][e],
                    joint_positions=inp["joint_positions"][e], joint_velocities=inp["joint_velocities"][e],
                    base_linear_velocity=inp["base_linear_velocity"][e], base_angular_velocity=inp["base_angular_velocity"][e],
                    velocity_representation=r)
                push(f"build_{name}_stored_velocity", np.concatenate([np.asarray(d_b._base_linear_velocity).reshape(3), np.asarray(d_b._base_angular_velocity).reshape(3)]))
                push(f"build_{name}_link_velocities", d_b._link_velocities)
                push(f"base_velocity_{name}", dataclasses.replace(data, velocity_representation=r).base_velocity)
                # the same state; link forces / base acceleration are READ in representation r and the
                # base acceleration / base force / mass matrix are RETURNED in it
                d_r = dataclasses.replace(data, velocity_representation=r)
                vd, sdd = js.model.forward_dynamics_aba(model=rm, data=d_r, joint_forces=inp["tau"][e], link_forces=lf)
                push(f"aba_base_acceleration_{name}", vd)
                push(f"aba_joint_accelerations_{name}", sdd)
                fb, tj = js.model.inverse_dynamics(model=rm, data=d_r, joint_accelerations=inp["joint_accelerations"][e],
                                                   base_acceleration=inp["base_acceleration"][e], link_forces=lf)
                push(f"rnea_base_force_{name}", fb)
                push(f"rnea_joint_forces_{name}", tj)
                push(f"mass_matrix_{name}", js.model.free_floating_mass_matrix(model=rm, data=d_r))
                push(f"bias_forces_{name}", js.model.free_floating_bias_forces(model=rm, data=d_r))
                push(f"gravity_forces_{name}", js.model.free_floating_gravity_forces(model=rm, data=d_r))
                push(f"mass_matrix_inverse_{name}", js.model.free_floating_mass_matrix_inverse(model=rm, data=d_r))
            if soft and not case["fext"]:
                xdot = js.ode.system_dynamics(model=rm, data=data, link_forces=None, joint_torques=inp["tau"][e])
                for k in ("base_position", "base_quaternion", "joint_positions", "base_linear_velocity", "base_angular_velocity", "joint_velocities"):
                    push("ode_" + k, xdot[k])
                push("ode_tangential_deformation", xdot["contact_state"]["tangential_deformation"])
    # estimate_good_contact_parameters (api/contact.py:155-211): defaults, and the arguments of the
    # reference's soft-contact rest test (tests/test_simulations.py:205-211)
    for tag, kw in (("default", {}), ("tuned", dict(number_of_active_collidable_points_steady_state=4,
                                                     static_friction_coefficient=1.0, damping_ratio=1.0, max_penetration=0.001))):
        prm = js.contact.estimate_good_contact_parameters(model=rm, **kw)
        outs[f"egcp_{tag}"] = [np.array([float(prm.K), float(prm.D), float(prm.mu)])]
    out = {k: np.stack(v) for k, v in outs.items()}
    out.update({"in_" + k: np.asarray(v) for k, v in inp.items()})
    out.update(kin_dyn_arrays(rm))
    spec = {k: v for k, v in case.items()}
    spec["urdf_sha256_16"] = C.urdf_digest(urdf_text)
    spec["reference"] = "ami-iit/jaxsim sources at /root/reference executed over oracle/refshim (NumPy stand-ins), float64"
    out["spec"] = np.array(json.dumps(spec))
    return out


def main():
    want = sys.argv[1:]
    for case in C.all_cases():
        if want and case["id"] not in want:
            continue
        t0 = time.time()
        out = run_case(case)
        np.savez_compressed(C.fixture_path(case["id"]), **out)
        print(f"{case['id']:32s} {time.time() - t0:6.1f}s  {C.fixture_path(case['id']).stat().st_size / 1024:7.1f} KiB", flush=True)


if __name__ == "__main__":
    main()
