"""The golden-fixture case table, shared by the generator (`make_goldens.py`, which runs the
reference) and by the tests that consume the fixtures (`tests/test_reference_goldens.py`,
`tests/test_gpu_reference_goldens.py`).

A case names a model of `jaxsim_b200.models`, a contact model and its parameters, an
integrator, and how the seeded inputs are drawn.  The inputs themselves are stored in the
fixture next to the reference's outputs, so the tests never regenerate them.
"""

from __future__ import annotations

import hashlib
import pathlib

GOLDEN_DIR = pathlib.Path(__file__).resolve().parent

_ICUB = dict(model="icub_like", B=3)

CASES = [
    # ---- soft contacts, semi-implicit Euler (BASELINE configs[0], [1])
    dict(id="pendulum_soft", model="pendulum", B=4, seed=0, tau=True, rbda=True),
    dict(id="double_pendulum_soft", model="double_pendulum", B=3, seed=1, tau=True),
    dict(id="cartpole_soft", model="cartpole", B=3, seed=2, tau=True, rbda=True),
    dict(id="box_soft_air", model="box", B=3, seed=3),
    dict(id="box_soft_contact", model="box", B=4, seed=4, in_contact=True, m=True, rbda=True),
    dict(id="box_soft_flat", model="box", B=3, seed=5, in_contact="flat", m=True),
    dict(id="sphere_soft_contact", model="sphere", B=2, seed=6, in_contact=True, m=True),
    dict(id="icub_soft_air", seed=7, tau=True, rbda=True, **_ICUB),
    dict(id="icub_soft_contact", seed=8, in_contact=True, tau=True, m=True, rbda=True, **_ICUB),
    dict(id="icub_soft_flat", seed=9, in_contact="flat", tau=True, m=True, **_ICUB),
    dict(id="icub_soft_params", seed=10, in_contact=True, tau=True, m=True,
         contact_params=dict(K=2e5, D=500.0, mu=0.8, p=0.7, q=1.3), **_ICUB),
    dict(id="icub_soft_nofriction", seed=11, in_contact=True, tau=True,
         actuation=dict(torque_max=40.0, omega_th=0.3, omega_max=0.9, enable_friction=False), **_ICUB),
    dict(id="icub_soft_fext_inertial", rbda=True, seed=12, in_contact=True, tau=True, fext=True, velrepr="inertial", **_ICUB),
    dict(id="icub_soft_fext_mixed", rbda=True, seed=13, in_contact=True, tau=True, fext=True, velrepr="mixed", **_ICUB),
    dict(id="icub_soft_fext_body", rbda=True, seed=14, in_contact=True, tau=True, fext=True, velrepr="body", **_ICUB),
    dict(id="ergocub_soft_contact", model="ergocub_like", B=2, seed=15, in_contact=True, tau=True, m=True),
    # ---- a model given in SDF (jaxsim_b200/models: POSED_SDF): posed base link, joint posed in the model frame, a link
    #      offset from its joint (non-identity successor transform), <limit> stiffness / dissipation (limit spring / damper)
    dict(id="posed_sdf_soft_contact", model="posed_sdf", B=6, seed=52, in_contact=True, tau=True, m=True, rbda=True, round32=True),
    # ---- RK4
    dict(id="icub_rk4_contact", seed=16, in_contact=True, tau=True, m=True, integrator="rk4", **_ICUB),
    dict(id="box_rk4_contact", model="box", B=3, seed=17, in_contact=True, m=True, integrator="rk4"),
    dict(id="icub_rk4fast_contact", seed=43, in_contact=True, tau=True, m=True, integrator="rk4fast", **_ICUB),
    dict(id="box_rk4fast_flat", model="box", B=3, seed=44, in_contact="flat", m=True, integrator="rk4fast"),
    # ---- rigid contacts (BASELINE configs[2])
    dict(id="box_rigid_air", model="box", B=2, seed=18, contact="rigid"),
    dict(id="box_rigid_contact", model="box", B=4, seed=19, contact="rigid", in_contact=True),
    dict(id="box_rigid_flat", model="box", B=4, seed=20, contact="rigid", in_contact="flat"),
    dict(id="box_rigid_flat_baumgarte", model="box", B=3, seed=21, contact="rigid", in_contact="flat",
         contact_params=dict(mu=0.7, K=1e4, D=20.0)),
    dict(id="icub_rigid_contact", seed=22, contact="rigid", in_contact=True, tau=True, **_ICUB),
    dict(id="icub_rigid_flat", seed=23, contact="rigid", in_contact="flat", tau=True, model="icub_like", B=2),
    dict(id="ergocub_rigid_flat", model="ergocub_like", B=2, seed=24, contact="rigid", in_contact="flat", tau=True),
    # ---- relaxed-rigid contacts (the contact model of the reference's own step benchmark)
    dict(id="box_relaxed_air", model="box", B=2, seed=25, contact="relaxed"),
    dict(id="box_relaxed_contact", model="box", B=4, seed=26, contact="relaxed", in_contact=True),
    dict(id="box_relaxed_flat", model="box", B=4, seed=27, contact="relaxed", in_contact="flat"),
    dict(id="box_relaxed_flat_params", model="box", B=3, seed=28, contact="relaxed", in_contact="flat",
         contact_params=dict(time_constant=0.01, damping_coefficient=0.8, d_min=0.7, d_max=0.9, width=0.004, midpoint=0.4,
                             power=3.0, mu=0.5)),
    dict(id="sphere_relaxed_contact", model="sphere", B=2, seed=29, contact="relaxed", in_contact=True),
    dict(id="icub_relaxed_contact", seed=30, contact="relaxed", in_contact=True, tau=True, **_ICUB),
    dict(id="icub_relaxed_flat", seed=31, contact="relaxed", in_contact="flat", tau=True, model="icub_like", B=2),
    dict(id="icub_relaxed_fext_mixed", seed=32, contact="relaxed", in_contact="flat", tau=True, fext=True, velrepr="mixed",
         model="icub_like", B=2),
    dict(id="ergocub_relaxed_flat", model="ergocub_like", B=2, seed=33, contact="relaxed", in_contact="flat", tau=True),
    # ---- short rollouts: the reference's `step` applied `rollout` times in a row (state carried through its own
    #      JaxSimModelData, incl. the contact state and -- for RigidContacts -- the stale cached link velocities)
    dict(id="box_soft_rollout", model="box", B=3, seed=34, in_contact=True, m=True, rollout=5),
    dict(id="icub_soft_rollout", seed=35, in_contact=True, tau=True, m=True, rollout=5, model="icub_like", B=2),
    dict(id="box_rigid_rollout", model="box", B=3, seed=36, contact="rigid", in_contact="flat", rollout=5),
    dict(id="icub_rigid_rollout", seed=37, contact="rigid", in_contact="flat", tau=True, rollout=4, model="icub_like", B=2),
    dict(id="icub_relaxed_rollout", seed=38, contact="relaxed", in_contact="flat", tau=True, rollout=4, model="icub_like", B=2),
    # ---- larger batches (goldens had 2-4 environments each) with inputs ROUNDED TO FLOAT32, so that the
    #      float32 kernel and the float64 reference start from exactly the same numbers
    dict(id="icub_soft_contact_b32", model="icub_like", B=32, seed=39, in_contact=True, tau=True, m=True, round32=True),
    dict(id="icub_soft_mixed_b32", model="icub_like", B=32, seed=40, in_contact="mixed", tau=True, m=True, round32=True),
    dict(id="ergocub_soft_flat_b32", model="ergocub_like", B=32, seed=41, in_contact="flat", tau=True, m=True, round32=True),
    dict(id="icub_rigid_flat_b16", model="icub_like", B=16, seed=42, contact="rigid", in_contact="flat", tau=True, round32=True),
    # ---- weld constraints (rbda/kinematic_constraints.py): a floating closed linkage whose loop is
    #      closed by one weld between two frames, like the reference's 4-bar test (tests/test_simulations.py:549-612).
    #      constraints = [(frame_1, frame_2, K_P or None, K_D or None)]
    #      The first two draw the joint angles over their whole range: the loop is wide open (gaps of ~0.5 m) and the
    #      solve returns wrench pairs of 1e6-1e7 N that cancel to ~1e3 N on the mechanism (three of the six weld directions
    #      of a planar loop are held by the 1e-3 regulariser alone).  No float32 link-force array can carry that
    #      cancellation -- the reference in float32 could not either -- so these two are float64-only (fp32=False); the
    #      `closed` cases scale the joint angles down to a nearly closed loop, the regime a weld runs in.
    dict(id="four_bar_weld", model="four_bar", B=4, seed=45, tau=True, round32=True, fp32=False,
         constraints=[("tip_a_frame", "tip_b_frame", None, None)]),
    dict(id="four_bar_weld_gains_fext", model="four_bar", B=4, seed=46, tau=True, fext=True, velrepr="mixed", round32=True, fp32=False,
         constraints=[("tip_a_frame", "tip_b_frame", 1e4, 50.0)]),
    dict(id="four_bar_weld_closed", model="four_bar", B=8, seed=49, tau=True, round32=True, joint_scale=0.02,
         constraints=[("tip_a_frame", "tip_b_frame", None, None)]),
    dict(id="four_bar_weld_closed_contact_fext", model="four_bar", B=8, seed=50, tau=True, fext=True, velrepr="body", round32=True,
         joint_scale=0.02, in_contact=True, m=True, constraints=[("tip_a_frame", "tip_b_frame", 1e4, None)]),
    dict(id="four_bar_fixed_weld", model="four_bar_fixed", B=4, seed=51, tau=True, round32=True, joint_scale=0.02,
         constraints=[("tip_a_frame", "tip_b_frame", None, None)]),
    dict(id="four_bar_weld_rollout", model="four_bar", B=2, seed=47, rollout=6, round32=True,
         constraints=[("tip_a_frame", "tip_b_frame", 1e4, None)]),
    dict(id="four_bar_weld_contact", model="four_bar", B=4, seed=48, in_contact=True, m=True, tau=True, round32=True,
         constraints=[("tip_a_frame", "tip_b_frame", None, None)]),
]

DEFAULTS = dict(contact="soft", contact_params=None, actuation=None, integrator="semi_implicit_euler", in_contact=False,
                tau=False, m=False, fext=False, velrepr="inertial", rbda=False, time_step=1e-3, rollout=1, round32=False, constraints=None, fp32=True,
                joint_scale=1.0)


def case(cid: str) -> dict:
    for c in CASES:
        if c["id"] == cid:
            return {**DEFAULTS, **c}
    raise KeyError(cid)


def all_cases() -> list[dict]:
    return [{**DEFAULTS, **c} for c in CASES]


def fixture_path(cid: str) -> pathlib.Path:
    return GOLDEN_DIR / f"{cid}.npz"


def urdf_digest(urdf_text: str) -> str:
    return hashlib.sha256(urdf_text.encode()).hexdigest()[:16]
