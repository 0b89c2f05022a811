"""In-repo synthetic URDF models.

The reference fetches iCub / ErgoCub from ``icub_models`` / ``robot_descriptions``
(``README.md:42-48``, ``tests/conftest.py:277-316``); neither is available offline, so the
benchmark and parity tests use stand-ins with the same topology and DoF count, generated
here as URDF *text* and parsed by :mod:`jaxsim_b200.parsers.urdf` exactly like a file on
disk would be.  The primitive models restate the recipes of the reference's fixtures:
box 0.3x0.2x0.1 m / 1 kg (``tests/conftest.py:207-243``), sphere r=0.1 m / 1 kg
(``:246-274``), single pendulum (``:370-476``).

Swap in the real URDFs with ``JaxSimModel.build_from_model_description(path)`` when they
are available -- nothing downstream depends on these particular numbers.
"""

from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ tiny URDF builder
def _f(v) -> str:
    return " ".join(f"{float(x):.10g}" for x in np.atleast_1d(v))


def box_inertia(m, size):
    x, y, z = size
    return m / 12 * np.array([y * y + z * z, x * x + z * z, x * x + y * y])


def sphere_inertia(m, r):
    return 2 / 5 * m * r * r * np.ones(3)


def cylinder_inertia(m, r, h):
    # axis along z
    return np.array([m / 12 * (3 * r * r + h * h), m / 12 * (3 * r * r + h * h), m / 2 * r * r])


class UrdfBuilder:
    def __init__(self, name: str):
        self.name = name
        self.parts: list[str] = []

    def link(self, name, mass, diag_inertia, com=(0, 0, 0), com_rpy=(0, 0, 0), products=(0, 0, 0), collisions=()):
        ixx, iyy, izz = (float(v) for v in diag_inertia)
        ixy, ixz, iyz = (float(v) for v in products)
        s = [f'  <link name="{name}">']
        s.append("    <inertial>")
        s.append(f'      <origin xyz="{_f(com)}" rpy="{_f(com_rpy)}"/>')
        s.append(f'      <mass value="{float(mass):.10g}"/>')
        s.append(
            f'      <inertia ixx="{ixx:.10g}" ixy="{ixy:.10g}" ixz="{ixz:.10g}" '
            f'iyy="{iyy:.10g}" iyz="{iyz:.10g}" izz="{izz:.10g}"/>'
        )
        s.append("    </inertial>")
        for kind, xyz, rpy, prm in collisions:
            s.append("    <collision>")
            s.append(f'      <origin xyz="{_f(xyz)}" rpy="{_f(rpy)}"/>')
            if kind == "box":
                s.append(f'      <geometry><box size="{_f(prm)}"/></geometry>')
            elif kind == "sphere":
                s.append(f'      <geometry><sphere radius="{float(prm):.10g}"/></geometry>')
            else:
                raise ValueError(kind)
            s.append("    </collision>")
        s.append("  </link>")
        self.parts.append("\n".join(s))
        return self

    def massless_link(self, name):
        self.parts.append(f'  <link name="{name}"/>')
        return self

    def joint(self, name, jtype, parent, child, xyz=(0, 0, 0), rpy=(0, 0, 0), axis=(0, 0, 1),
              limit=None, damping=0.0, friction=0.0, effort=1000.0, velocity=100.0):
        s = [f'  <joint name="{name}" type="{jtype}">']
        s.append(f'    <origin xyz="{_f(xyz)}" rpy="{_f(rpy)}"/>')
        s.append(f'    <parent link="{parent}"/>')
        s.append(f'    <child link="{child}"/>')
        if jtype != "fixed":
            s.append(f'    <axis xyz="{_f(axis)}"/>')
            if limit is not None:
                s.append(
                    f'    <limit lower="{float(limit[0]):.10g}" upper="{float(limit[1]):.10g}" '
                    f'effort="{effort:.10g}" velocity="{velocity:.10g}"/>'
                )
            elif jtype != "continuous":
                raise ValueError("revolute/prismatic joints need limits")
            if damping != 0.0 or friction != 0.0:
                s.append(f'    <dynamics damping="{float(damping):.10g}" friction="{float(friction):.10g}"/>')
        s.append("  </joint>")
        self.parts.append("\n".join(s))
        return self

    def urdf(self) -> str:
        return f'<?xml version="1.0"?>\n<robot name="{self.name}">\n' + "\n".join(self.parts) + "\n</robot>\n"


# ------------------------------------------------------------------------- primitives
def box_urdf(size=(0.3, 0.2, 0.1), mass=1.0) -> str:
    """Floating box with 8 collidable corners (``tests/conftest.py:207-243``)."""
    b = UrdfBuilder("box")
    b.link("box_link", mass, box_inertia(mass, size), collisions=[("box", (0, 0, 0), (0, 0, 0), size)])
    return b.urdf()


def sphere_urdf(radius=0.1, mass=1.0) -> str:
    """Floating sphere with the 50-point Fibonacci lattice (``tests/conftest.py:246-274``)."""
    b = UrdfBuilder("sphere")
    b.link("sphere_link", mass, sphere_inertia(mass, radius), collisions=[("sphere", (0, 0, 0), (0, 0, 0), radius)])
    return b.urdf()


def pendulum_urdf(mount_height=1.0, fixed_base=True) -> str:
    """BASELINE config 0: box-and-sphere 2-link pendulum.

    A box "base" (fixed to the world at ``mount_height``) carrying, through one revolute
    joint about x, an arm whose bob is a sphere (lumped through a fixed joint, so the
    fixed-joint lumping path of the loader is exercised).
    """
    b = UrdfBuilder("pendulum")
    b.link("base", 1.0, box_inertia(1.0, (0.1, 0.1, 0.1)), collisions=[("box", (0, 0, 0), (0, 0, 0), (0.1, 0.1, 0.1))])
    b.link("arm", 0.5, cylinder_inertia(0.5, 0.02, 0.5), com=(0, 0, -0.25))
    b.link("bob", 1.0, sphere_inertia(1.0, 0.05), collisions=[("sphere", (0, 0, 0), (0, 0, 0), 0.05)])
    if fixed_base:
        b.massless_link("world")
        b.joint("world_to_base", "fixed", "world", "base", xyz=(0, 0, mount_height))
    b.joint("pivot", "continuous", "base", "arm", xyz=(0, 0.06, 0), axis=(1, 0, 0), damping=0.0)
    b.joint("arm_to_bob", "fixed", "arm", "bob", xyz=(0, 0, -0.5))
    return b.urdf()


def double_pendulum_urdf() -> str:
    """Fixed-base planar double pendulum (two revolute joints about y)."""
    b = UrdfBuilder("double_pendulum")
    b.massless_link("world")
    b.link("support", 2.0, box_inertia(2.0, (0.1, 0.1, 0.1)))
    b.link("upper", 1.0, cylinder_inertia(1.0, 0.03, 0.6), com=(0, 0, -0.3))
    b.link("lower", 0.7, cylinder_inertia(0.7, 0.03, 0.5), com=(0, 0, -0.25))
    b.joint("world_to_support", "fixed", "world", "support", xyz=(0, 0, 1.5))
    b.joint("shoulder", "revolute", "support", "upper", xyz=(0, 0.05, 0), axis=(0, 1, 0), limit=(-3.0, 3.0), damping=0.05)
    b.joint("elbow", "revolute", "upper", "lower", xyz=(0, 0.05, -0.6), axis=(0, 1, 0), limit=(-2.5, 2.5), damping=0.02, friction=0.01)
    return b.urdf()


def cartpole_urdf() -> str:
    """Fixed-base rail + prismatic cart + revolute pole (same topology as the reference's
    ``examples/assets/cartpole.urdf``, numbers are this repo's own)."""
    b = UrdfBuilder("cartpole")
    b.massless_link("world")
    b.link("rail", 5.0, box_inertia(5.0, (0.05, 4.0, 0.05)))
    b.link("cart", 1.0, box_inertia(1.0, (0.2, 0.3, 0.1)))
    b.link("pole", 0.3, cylinder_inertia(0.3, 0.02, 0.8), com=(0, 0, 0.4))
    b.joint("world_to_rail", "fixed", "world", "rail", xyz=(0, 0, 0.8))
    b.joint("linear", "prismatic", "rail", "cart", axis=(0, 1, 0), limit=(-1.9, 1.9), damping=0.1)
    b.joint("pivot", "continuous", "cart", "pole", xyz=(0.12, 0, 0), axis=(1, 0, 0))
    return b.urdf()


def four_bar_urdf(fixed_base: bool = False) -> str:
    """Floating-base planar linkage opened at its coupler: a base bar, two cranks and the two halves of the coupler,
    each half ending in a massless frame link.  Welding `tip_a_frame` to `tip_b_frame` closes the loop (same topology as the
    reference's tests/assets/4_bar_opened.urdf, rbda/kinematic_constraints.py; dimensions and masses are this repo's own)."""
    b = UrdfBuilder("four_bar_fixed" if fixed_base else "four_bar")
    L, H, W = 0.6, 0.35, 0.08
    if fixed_base:  # the ground bar bolted to the world, like the reference's fixed-base constraint tests
        b.massless_link("world")
        b.joint("world_to_ground_bar", "fixed", "world", "ground_bar", xyz=(0, 0, 0.5))
    b.link("ground_bar", 1.2, box_inertia(1.2, (W, L, W)), collisions=[("box", (0, 0, 0), (0, 0, 0), (W, L, W))])
    b.link("crank_a", 0.6, box_inertia(0.6, (W, W, H)), com=(0, 0, H / 2))
    b.link("crank_b", 0.6, box_inertia(0.6, (W, W, H)), com=(0, 0, H / 2))
    b.link("coupler_a", 0.4, box_inertia(0.4, (W, L / 2, W)), com=(0, L / 4, 0))
    b.link("coupler_b", 0.4, box_inertia(0.4, (W, L / 2, W)), com=(0, -L / 4, 0))
    b.massless_link("tip_a_frame")
    b.massless_link("tip_b_frame")
    b.joint("pivot_a", "revolute", "ground_bar", "crank_a", xyz=(0, -L / 2, 0), axis=(1, 0, 0), limit=(-1.5, 1.5), damping=0.02)
    b.joint("pivot_b", "revolute", "ground_bar", "crank_b", xyz=(0, L / 2, 0), axis=(1, 0, 0), limit=(-1.5, 1.5), damping=0.02)
    b.joint("elbow_a", "revolute", "crank_a", "coupler_a", xyz=(0, 0, H), axis=(1, 0, 0), limit=(-2.5, 2.5))
    b.joint("elbow_b", "revolute", "crank_b", "coupler_b", xyz=(0, 0, H), axis=(1, 0, 0), limit=(-2.5, 2.5))
    b.joint("tip_a_fix", "fixed", "coupler_a", "tip_a_frame", xyz=(0, L / 2, 0))
    b.joint("tip_b_fix", "fixed", "coupler_b", "tip_b_frame", xyz=(0, -L / 2, 0), rpy=(0, 0, 0))
    return b.urdf()


# ---------------------------------------------------------------------- humanoids
def _leg(b: UrdfBuilder, side: str, parent: str, sgn: float, foot_boxes: int = 1):
    p = side + "_"
    # hip pitch / roll / yaw, knee, ankle pitch / roll  (6 DoF)
    b.link(p + "hip_1", 0.75, box_inertia(0.75, (0.06, 0.06, 0.06)), com=(0, 0.01 * sgn, 0))
    b.link(p + "hip_2", 0.95, box_inertia(0.95, (0.07, 0.07, 0.07)), com=(0, 0, -0.02))
    b.link(p + "upper_leg", 2.2, cylinder_inertia(2.2, 0.05, 0.24), com=(0.002, 0.003 * sgn, -0.12), products=(1e-4, -2e-4, 5e-5))
    b.link(p + "lower_leg", 1.3, cylinder_inertia(1.3, 0.04, 0.22), com=(-0.003, 0, -0.11))
    b.link(p + "ankle_1", 0.65, box_inertia(0.65, (0.05, 0.05, 0.05)))
    cols = []
    if foot_boxes == 1:
        cols = [("box", (0.03, 0, -0.045), (0, 0, 0), (0.16, 0.07, 0.03))]
    else:
        cols = [
            ("box", (0.075, 0, -0.045), (0, 0, 0), (0.09, 0.07, 0.03)),
            ("box", (-0.025, 0, -0.045), (0, 0, 0), (0.07, 0.07, 0.03)),
        ]
    b.link(p + "foot", 0.6, box_inertia(0.6, (0.16, 0.07, 0.03)), com=(0.03, 0, -0.04), collisions=cols)
    b.joint(p + "hip_pitch", "revolute", parent, p + "hip_1", xyz=(0, 0.068 * sgn, -0.05), rpy=(0, 0, 0), axis=(0, 1, 0), limit=(-0.77, 2.3), damping=0.1)
    b.joint(p + "hip_roll", "revolute", p + "hip_1", p + "hip_2", xyz=(0, 0, 0), rpy=(0.05 * sgn, 0, 0), axis=(1, 0, 0), limit=(-0.3, 2.0) if sgn > 0 else (-2.0, 0.3), damping=0.1)
    b.joint(p + "hip_yaw", "revolute", p + "hip_2", p + "upper_leg", xyz=(0, 0, -0.06), axis=(0, 0, 1), limit=(-1.4, 1.4), damping=0.1, friction=0.02)
    b.joint(p + "knee", "revolute", p + "upper_leg", p + "lower_leg", xyz=(0, 0, -0.24), rpy=(0, 0.02, 0), axis=(0, 1, 0), limit=(-2.2, 0.02), damping=0.1)
    b.joint(p + "ankle_pitch", "revolute", p + "lower_leg", p + "ankle_1", xyz=(0, 0, -0.22), axis=(0, 1, 0), limit=(-0.75, 0.75), damping=0.06, friction=0.02)
    b.joint(p + "ankle_roll", "revolute", p + "ankle_1", p + "foot", xyz=(0, 0, 0), axis=(1, 0, 0), limit=(-0.42, 0.42), damping=0.06)


def _arm(b: UrdfBuilder, side: str, parent: str, sgn: float, wrist: bool = False, fingers: int = 0):
    p = side + "_"
    b.link(p + "shoulder_1", 0.48, box_inertia(0.48, (0.05, 0.05, 0.05)))
    b.link(p + "shoulder_2", 0.2, box_inertia(0.2, (0.04, 0.04, 0.04)))
    b.link(p + "upper_arm", 1.1, cylinder_inertia(1.1, 0.035, 0.15), com=(0, 0.005 * sgn, -0.075), products=(2e-5, 0, -1e-5))
    b.joint(p + "shoulder_pitch", "revolute", parent, p + "shoulder_1", xyz=(0.0, 0.11 * sgn, 0.14), rpy=(0.26 * sgn, 0, 0), axis=(0, 1, 0), limit=(-1.65, 0.2), damping=0.06)
    b.joint(p + "shoulder_roll", "revolute", p + "shoulder_1", p + "shoulder_2", xyz=(0, 0, 0), axis=(1, 0, 0), limit=(0.0, 2.8) if sgn > 0 else (-2.8, 0.0), damping=0.06)
    b.joint(p + "shoulder_yaw", "revolute", p + "shoulder_2", p + "upper_arm", xyz=(0, 0, -0.02), axis=(0, 0, 1), limit=(-0.65, 1.4), damping=0.06, friction=0.01)
    if not wrist:
        # forearm + hand lumped in one body (iCub 23-DoF joint list, README.md:50-55)
        b.link(p + "forearm", 0.95, cylinder_inertia(0.95, 0.03, 0.2), com=(0, 0, -0.09))
        b.joint(p + "elbow", "revolute", p + "upper_arm", p + "forearm", xyz=(0, 0, -0.152), rpy=(0, 0, 0.1 * sgn), axis=(0, 1, 0), limit=(0.09, 1.85), damping=0.06)
        return
    b.link(p + "forearm", 0.6, cylinder_inertia(0.6, 0.03, 0.14), com=(0, 0, -0.07))
    b.joint(p + "elbow", "revolute", p + "upper_arm", p + "forearm", xyz=(0, 0, -0.152), rpy=(0, 0, 0.1 * sgn), axis=(0, 1, 0), limit=(0.09, 1.85), damping=0.06)
    b.link(p + "wrist_1", 0.15, box_inertia(0.15, (0.03, 0.03, 0.03)))
    b.link(p + "wrist_2", 0.1, box_inertia(0.1, (0.03, 0.03, 0.03)))
    b.link(p + "hand", 0.35, box_inertia(0.35, (0.08, 0.03, 0.09)), com=(0, 0, -0.04))
    b.joint(p + "wrist_yaw", "revolute", p + "forearm", p + "wrist_1", xyz=(0, 0, -0.14), axis=(0, 0, 1), limit=(-1.5, 1.5), damping=0.02)
    b.joint(p + "wrist_roll", "revolute", p + "wrist_1", p + "wrist_2", axis=(1, 0, 0), limit=(-0.5, 0.5), damping=0.02)
    b.joint(p + "wrist_pitch", "revolute", p + "wrist_2", p + "hand", axis=(0, 1, 0), limit=(-0.7, 0.7), damping=0.02)
    names = ["thumb", "index", "middle", "ring", "pinkie"]
    k = 0
    for fi, fn in enumerate(names):
        if k >= fingers:
            break
        prox = p + fn + "_prox"
        b.link(prox, 0.02, box_inertia(0.02, (0.012, 0.012, 0.035)), com=(0, 0, -0.017))
        b.joint(p + fn + "_add" if fn == "thumb" else p + fn + "_prox_j", "revolute", p + "hand", prox,
                xyz=(0.03 - 0.015 * fi, 0.0, -0.085), axis=(0, 0, 1) if fn == "thumb" else (1, 0, 0), limit=(0.0, 1.5), damping=0.002)
        k += 1
        if k >= fingers:
            break
        if fn in ("thumb", "index", "middle"):
            dist = p + fn + "_dist"
            b.link(dist, 0.012, box_inertia(0.012, (0.01, 0.01, 0.03)), com=(0, 0, -0.015))
            b.joint(p + fn + "_dist_j", "revolute", prox, dist, xyz=(0, 0, -0.035), axis=(1, 0, 0), limit=(0.0, 1.6), damping=0.002)
            k += 1


def icub_like_urdf() -> str:
    """23-DoF floating-base humanoid with the iCub joint list of ``README.md:50-55``
    (torso 3, arms 4+4, legs 6+6), 24 links after lumping, two box feet -> nc = 16."""
    b = UrdfBuilder("icub_like")
    b.link("root_link", 4.7, box_inertia(4.7, (0.12, 0.18, 0.12)), com=(0.0, 0, -0.02), products=(1e-4, 2e-4, 0))
    b.link("torso_1", 0.6, box_inertia(0.6, (0.06, 0.06, 0.06)))
    b.link("torso_2", 0.5, box_inertia(0.5, (0.05, 0.05, 0.05)))
    b.link("chest", 6.0, box_inertia(6.0, (0.13, 0.2, 0.2)), com=(-0.005, 0, 0.1), products=(0, 3e-4, 0))
    # head rigidly attached to the chest: exercises fixed-joint lumping in the loader
    b.link("head", 1.8, sphere_inertia(1.8, 0.08), com=(0.01, 0, 0.05))
    b.joint("torso_pitch", "revolute", "root_link", "torso_1", xyz=(0, 0, 0.05), axis=(0, 1, 0), limit=(-0.38, 1.22), damping=0.1)
    b.joint("torso_roll", "revolute", "torso_1", "torso_2", xyz=(0, 0, 0.0), axis=(1, 0, 0), limit=(-0.68, 0.68), damping=0.1)
    b.joint("torso_yaw", "revolute", "torso_2", "chest", xyz=(0, 0, 0.03), rpy=(0, 0, 0), axis=(0, 0, 1), limit=(-0.87, 0.87), damping=0.1, friction=0.02)
    b.joint("neck_fixed", "fixed", "chest", "head", xyz=(0, 0, 0.24))
    _arm(b, "l", "chest", +1.0)
    _arm(b, "r", "chest", -1.0)
    _leg(b, "l", "root_link", +1.0)
    _leg(b, "r", "root_link", -1.0)
    return b.urdf()


def ergocub_like_urdf() -> str:
    """~50-DoF floating-base humanoid (BASELINE config 2 stand-in): torso 3, neck 4,
    arms 7+7 (shoulder 3, elbow, wrist 3), hands 7+7, legs 6+6 -> 47 DoF... plus
    camera tilt and two extra finger joints = 51 DoF; four foot boxes -> nc = 32."""
    b = UrdfBuilder("ergocub_like")
    b.link("root_link", 6.5, box_inertia(6.5, (0.15, 0.22, 0.14)), com=(0.0, 0, -0.02))
    b.link("torso_1", 0.9, box_inertia(0.9, (0.07, 0.07, 0.07)))
    b.link("torso_2", 0.8, box_inertia(0.8, (0.06, 0.06, 0.06)))
    b.link("chest", 9.0, box_inertia(9.0, (0.16, 0.26, 0.26)), com=(-0.005, 0, 0.13))
    b.joint("torso_pitch", "revolute", "root_link", "torso_1", xyz=(0, 0, 0.06), axis=(0, 1, 0), limit=(-0.3, 0.8), damping=0.1)
    b.joint("torso_roll", "revolute", "torso_1", "torso_2", axis=(1, 0, 0), limit=(-0.4, 0.4), damping=0.1)
    b.joint("torso_yaw", "revolute", "torso_2", "chest", xyz=(0, 0, 0.04), axis=(0, 0, 1), limit=(-0.7, 0.7), damping=0.1)
    b.link("neck_1", 0.3, box_inertia(0.3, (0.04, 0.04, 0.04)))
    b.link("neck_2", 0.25, box_inertia(0.25, (0.04, 0.04, 0.04)))
    b.link("head", 2.0, sphere_inertia(2.0, 0.09), com=(0.01, 0, 0.06))
    b.link("camera", 0.1, box_inertia(0.1, (0.03, 0.08, 0.03)))
    b.joint("neck_pitch", "revolute", "chest", "neck_1", xyz=(0, 0, 0.3), axis=(0, 1, 0), limit=(-0.5, 0.4), damping=0.03)
    b.joint("neck_roll", "revolute", "neck_1", "neck_2", axis=(1, 0, 0), limit=(-0.4, 0.4), damping=0.03)
    b.joint("neck_yaw", "revolute", "neck_2", "head", xyz=(0, 0, 0.03), axis=(0, 0, 1), limit=(-0.9, 0.9), damping=0.03)
    b.joint("camera_tilt", "revolute", "head", "camera", xyz=(0.07, 0, 0.08), axis=(0, 1, 0), limit=(-0.5, 0.5), damping=0.005)
    _arm(b, "l", "chest", +1.0, wrist=True, fingers=8)
    _arm(b, "r", "chest", -1.0, wrist=True, fingers=8)
    _leg(b, "l", "root_link", +1.0, foot_boxes=2)
    _leg(b, "r", "root_link", -1.0, foot_boxes=2)
    return b.urdf()


# An SDF document that exercises the pose semantics parsers/sdf.py resolves: a floating base whose link frame is posed
# in the model frame, a joint posed in the model frame (not in its child), a child link offset from its joint
# (non-identity successor transform), an explicit <frame> attached to a link, <limit> stiffness / dissipation.
POSED_SDF = """<?xml version="1.0"?>
<sdf version="1.9"><model name="posed">
  <link name="trunk"><pose>0.1 0.2 0.3 0.1 -0.2 0.3</pose>
    <inertial><pose>0.01 0 0.02 0 0.1 0</pose><mass>3</mass><inertia><ixx>0.03</ixx><iyy>0.04</iyy><izz>0.05</izz><ixy>0.001</ixy><ixz>0</ixz><iyz>0.002</iyz></inertia></inertial>
    <collision name="c0"><pose>0 0 -0.05 0 0 0.2</pose><geometry><box><size>0.2 0.1 0.05</size></box></geometry></collision></link>
  <joint name="hip" type="revolute"><pose relative_to="__model__">0.1 0.3 0.3 0.2 0 0</pose><parent>trunk</parent><child>thigh</child>
    <axis><xyz>0 1 0</xyz><limit><lower>-1</lower><upper>1.5</upper><stiffness>50</stiffness><dissipation>2</dissipation></limit>
      <dynamics><damping>0.1</damping><friction>0.05</friction></dynamics></axis></joint>
  <link name="thigh"><pose relative_to="hip">0 0 0 0 0 0</pose>
    <inertial><pose>0 0 -0.1 0 0 0</pose><mass>1</mass><inertia><ixx>0.01</ixx><iyy>0.01</iyy><izz>0.002</izz></inertia></inertial></link>
  <joint name="knee" type="revolute"><pose>0 0 0.05 0 0 0</pose><parent>thigh</parent><child>shin</child><axis><xyz>1 0 0</xyz></axis></joint>
  <link name="shin"><pose relative_to="thigh">0.02 0 -0.3 0 0.1 0</pose>
    <inertial><pose>0 0 -0.1 0 0 0</pose><mass>0.5</mass><inertia><ixx>0.004</ixx><iyy>0.004</iyy><izz>0.001</izz></inertia></inertial>
    <collision name="c1"><pose>0 0 -0.15 0 0 0</pose><geometry><sphere><radius>0.03</radius></sphere></geometry></collision></link>
  <joint name="slide" type="prismatic"><parent>trunk</parent><child>arm</child><axis><xyz>1 0 0</xyz><limit><lower>-0.2</lower><upper>0.3</upper></limit></axis></joint>
  <link name="arm"><pose relative_to="trunk">0 -0.1 0.1 0 0 0.5</pose>
    <inertial><mass>0.8</mass><inertia><ixx>0.002</ixx><iyy>0.008</iyy><izz>0.008</izz></inertia></inertial></link>
  <frame name="camera" attached_to="trunk"><pose>0.05 0 0.1 0 0.3 0</pose></frame>
</model></sdf>"""


MODELS = {
    "box": box_urdf,
    "sphere": sphere_urdf,
    "pendulum": pendulum_urdf,
    "double_pendulum": double_pendulum_urdf,
    "cartpole": cartpole_urdf,
    "icub_like": icub_like_urdf,
    "ergocub_like": ergocub_like_urdf,
    "four_bar": four_bar_urdf,
    "four_bar_fixed": lambda: four_bar_urdf(fixed_base=True),
    "posed_sdf": lambda: POSED_SDF,
}


def urdf(name: str) -> str:
    """The description of a built-in model: URDF text, or SDF text for ``posed_sdf`` (the loader tells them apart)."""
    return MODELS[name]()
