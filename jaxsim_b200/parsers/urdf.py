"""Minimal URDF -> :class:`KinDynParameters` loader (host side, NumPy).

The reference builds its model through ``rod`` + ``sdformat`` (``gz sdf``), neither of
which exists offline, so this module restates the *observable result* of that pipeline
for URDF inputs, citing the reference logic it reproduces:

* links with ``mass <= 0`` are dropped / become frames
  (``parsers/rod/parser.py:110-139``);
* a fixed joint whose parent is ``world`` makes the model fixed-base
  (``parsers/rod/parser.py:147-197``); its origin ends up in ``suc_H_i[0]``
  (``math/joint_model.py:78-83``);
* the other fixed joints are removed by lumping the child into the parent,
  ``M_parent += X^T M_child X`` with ``X = Ad(parent_H_child^-1)``
  (``parsers/descriptions/link.py:86-115``, ``parsers/kinematic_graph.py:379-611``);
* link index = BFS from the base link with children sorted by name
  (``parsers/kinematic_graph.py:669-709``); joint index = child link index;
* URDF frame convention: ``lam_H_pre[i] = <origin>`` of joint *i* w.r.t. the parent
  link and ``suc_H_i[i] = I`` (``math/joint_model.py:92-98``);
* link 6D inertia ``M_L = X^T M_CoM X`` with ``X = Ad(L_H_CoM^-1)``
  (``parsers/rod/utils.py:21-66``);
* box collision -> 8 corners (bottom 4 then top 4), sphere -> Fibonacci lattice of
  ``JAXSIM_COLLISION_SPHERE_POINTS`` (50) points; ``JAXSIM_COLLISION_USE_BOTTOM_ONLY``
  honoured (``parsers/rod/utils.py:102-225``); meshes/cylinders are ignored
  (``parsers/rod/parser.py:334-357``);
* joint parameters (``parsers/rod/parser.py:234-277``): friction_static <-
  ``<dynamics friction>``, friction_viscous <- ``<dynamics damping>``, limits default to
  +-finfo.max, limit spring/damper default to the ``JAXSIM_JOINT_POSITION_LIMIT_*`` env
  vars (0).
"""

from __future__ import annotations

import dataclasses
import os
import pathlib
import xml.etree.ElementTree as ET

import numpy as np

from jaxsim_b200.api.kin_dyn_parameters import (
    ContactParameters,
    FrameParameters,
    JointModel,
    JointParameters,
    JointType,
    KinDynParameters,
    LinkParameters,
)


# ----------------------------------------------------------------------------- helpers
def _rpy_to_R(rpy) -> np.ndarray:
    r, p, y = (float(a) for a in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _origin_to_H(elem) -> np.ndarray:
    H = np.eye(4)
    if elem is None:
        return H
    xyz = [float(v) for v in elem.get("xyz", "0 0 0").split()]
    rpy = [float(v) for v in elem.get("rpy", "0 0 0").split()]
    H[0:3, 0:3] = _rpy_to_R(rpy)
    H[0:3, 3] = xyz
    return H


def _wedge(v) -> np.ndarray:
    x, y, z = np.asarray(v, dtype=float).reshape(3)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def _adjoint_inverse(H: np.ndarray) -> np.ndarray:
    """``Adjoint.from_rotation_and_translation(..., inverse=True)``
    (``math/adjoint.py:99-105``): B_X_A of A_H_B."""
    R, p = H[0:3, 0:3], H[0:3, 3]
    X = np.zeros((6, 6))
    X[0:3, 0:3] = R.T
    X[0:3, 3:6] = -R.T @ _wedge(p)
    X[3:6, 3:6] = R.T
    return X


def _sixd_inertia(m: float, com: np.ndarray, I: np.ndarray) -> np.ndarray:
    """``Inertia.to_sixd`` (``math/inertia.py:14-41``)."""
    c = _wedge(com)
    M = np.zeros((6, 6))
    M[0:3, 0:3] = m * np.eye(3)
    M[0:3, 3:6] = m * c.T
    M[3:6, 0:3] = m * c
    M[3:6, 3:6] = I + m * c @ c.T
    return M


# ------------------------------------------------------------------------ description
@dataclasses.dataclass
class _Link:
    name: str
    mass: float
    inertia: np.ndarray  # 6x6 in link frame
    collisions: list  # list of (kind, H, params)
    children: list = dataclasses.field(default_factory=list)
    parent: "_Link | None" = None
    index: int = -1


@dataclasses.dataclass
class _Joint:
    name: str
    jtype: int
    parent: str
    child: str
    pose: np.ndarray  # parent_H_joint (== parent_H_child at s=0 in URDF)
    axis: np.ndarray
    position_limit: tuple[float, float]
    friction_static: float
    friction_viscous: float
    position_limit_damper: float
    position_limit_spring: float
    index: int = -1
    suc: np.ndarray = dataclasses.field(default_factory=lambda: np.eye(4))  # joint_H_child (identity in a URDF; SDF: parsers/sdf.py)


def _box_points(size, H) -> np.ndarray:
    """``create_box_collision`` (``parsers/rod/utils.py:102-155``)."""
    x, y, z = size
    center = np.array([x / 2, y / 2, z / 2])
    bottom = np.array([[0, 0, 0], [x, 0, 0], [x, y, 0], [0, y, 0]], dtype=float)
    use_top = os.environ.get("JAXSIM_COLLISION_USE_BOTTOM_ONLY", "0").lower() in {"false", "0"}
    top = np.array([[0, 0, z], [x, 0, z], [x, y, z], [0, y, z]], dtype=float)
    corners = (np.vstack([bottom, top]) if use_top else bottom) - center
    return corners @ H[0:3, 0:3].T + H[0:3, 3]


def _sphere_points(radius, H) -> np.ndarray:
    """``create_sphere_collision`` (``parsers/rod/utils.py:158-225``)."""
    samples = int(os.getenv("JAXSIM_COLLISION_SPHERE_POINTS", "50"))
    phi = np.pi * (3.0 - np.sqrt(5.0))
    pts = []
    for i in range(samples):
        y = 1 - 2 * i / (samples - 1)
        rad = np.sqrt(1 - y**2)
        pts.append([np.cos(phi * i) * rad, y, np.sin(phi * i) * rad])
    pts = np.array(pts)
    if os.environ.get("JAXSIM_COLLISION_USE_BOTTOM_ONLY", "0").lower() in {"true", "1"}:
        pts = pts[pts[:, 2] <= 0]
    pts = radius * pts
    return pts @ H[0:3, 0:3].T + H[0:3, 3]


def _root_tag(xml_text: str) -> str:
    return ET.fromstring(xml_text).tag


def _parse(xml_text: str):
    root = ET.fromstring(xml_text)
    if root.tag != "robot":
        raise ValueError("Expected a URDF (<robot>) description; SDF documents go through parsers/sdf.py")
    name = root.get("name", "model")

    links: dict[str, _Link] = {}
    seq = 0  # position of a collision shape in the file: the reference keeps this order (parsers/rod/parser.py:303-345)
    for le in root.findall("link"):
        lname = le.get("name")
        ine = le.find("inertial")
        if ine is None:
            mass, M = 0.0, np.zeros((6, 6))
        else:
            mass = float(ine.find("mass").get("value"))
            ie = ine.find("inertia")
            g = lambda k: float(ie.get(k, "0"))  # noqa: E731
            I_com = np.array(
                [[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]]
            )
            M_com = _sixd_inertia(mass, np.zeros(3), I_com)
            X = _adjoint_inverse(_origin_to_H(ine.find("origin")))  # CoM_X_L
            M = X.T @ M_com @ X
        cols = []
        for ce in le.findall("collision"):
            H = _origin_to_H(ce.find("origin"))
            geo = ce.find("geometry")
            if geo is None:
                continue
            if geo.find("box") is not None:
                size = [float(v) for v in geo.find("box").get("size").split()]
                cols.append(("box", H, size, seq))
                seq += 1
            elif geo.find("sphere") is not None:
                cols.append(("sphere", H, float(geo.find("sphere").get("radius")), seq))
                seq += 1
            # cylinder / capsule / mesh: not supported by the reference defaults either.
        links[lname] = _Link(name=lname, mass=mass, inertia=M, collisions=cols)

    joints: list[_Joint] = []
    for je in root.findall("joint"):
        jt = je.get("type")
        if jt in ("revolute", "continuous"):
            jtype = JointType.Revolute
        elif jt == "prismatic":
            jtype = JointType.Prismatic
        elif jt == "fixed":
            jtype = JointType.Fixed
        else:
            raise ValueError(f"Joint type '{jt}' not supported")
        axis_e = je.find("axis")
        axis = np.array([float(v) for v in axis_e.get("xyz").split()]) if axis_e is not None else np.array([1.0, 0, 0])
        if jtype != JointType.Fixed:
            axis = axis / np.linalg.norm(axis)
        lim = je.find("limit")
        fmax = float(np.finfo(float).max)
        lo = float(lim.get("lower")) if (lim is not None and lim.get("lower") is not None and jt != "continuous") else -fmax
        hi = float(lim.get("upper")) if (lim is not None and lim.get("upper") is not None and jt != "continuous") else fmax
        dyn = je.find("dynamics")
        joints.append(
            _Joint(
                name=je.get("name"),
                jtype=int(jtype),
                parent=je.find("parent").get("link"),
                child=je.find("child").get("link"),
                pose=_origin_to_H(je.find("origin")),
                axis=axis,
                position_limit=(lo, hi),
                friction_static=float(dyn.get("friction", "0")) if dyn is not None else 0.0,
                friction_viscous=float(dyn.get("damping", "0")) if dyn is not None else 0.0,
                position_limit_damper=float(os.environ.get("JAXSIM_JOINT_POSITION_LIMIT_DAMPER", 0.0)),
                position_limit_spring=float(os.environ.get("JAXSIM_JOINT_POSITION_LIMIT_SPRING", 0.0)),
            )
        )
    return name, links, joints


def build_kin_dyn_parameters(model_description: str | pathlib.Path) -> tuple[str, KinDynParameters, bool]:
    """Parse a URDF (path or XML string) into ``(name, KinDynParameters, floating_base)``."""

    text = str(model_description)
    if not text.lstrip().startswith("<"):
        text = pathlib.Path(model_description).read_text()
    frames0: list[tuple[str, str, np.ndarray]] = []
    base_link_pose = np.eye(4)  # pose of the base link in the model frame (SDF only; a URDF root link IS the model frame)
    if _root_tag(text) == "sdf":
        from . import sdf as _sdf

        name, links, joints, frames0, link_poses = _sdf.parse(text)
    else:
        link_poses = None
        name, links, joints = _parse(text)

    # ---- fixed-base detection (parsers/rod/parser.py:147-197)
    world_joints = [j for j in joints if j.parent == "world"]
    if "world" in links:
        del links["world"]
    floating_base = len(world_joints) == 0
    base_pose = np.eye(4)
    base_name = None
    if not floating_base:
        if len(world_joints) != 1 or world_joints[0].jtype != JointType.Fixed:
            raise ValueError("Found more/less than one fixed joint connecting the model to the world")
        base_name = world_joints[0].child
        base_pose = world_joints[0].pose @ world_joints[0].suc  # parser.py:192-197: joint pose @ pose of the base link
        joints = [j for j in joints if j.parent != "world"]

    # ---- lump the children of fixed joints into their parents, leaves first
    #      (parsers/kinematic_graph.py:379-611, parsers/descriptions/link.py:86-115).
    frames: list[tuple[str, str, np.ndarray]] = list(frames0)  # (frame name, parent link, parent_H_frame)
    while True:
        fixed = [j for j in joints if j.jtype == JointType.Fixed]
        if not fixed:
            break
        # pick a fixed joint whose child has no fixed-joint descendants left to process first
        parents_of_fixed = {j.parent for j in fixed}
        j = next((f for f in fixed if f.child not in parents_of_fixed), fixed[0])
        parent, child = links[j.parent], links[j.child]
        p_H_c = j.pose @ j.suc
        if child.mass > 0:
            X = _adjoint_inverse(p_H_c)  # c_X_p
            parent.inertia = parent.inertia + X.T @ child.inertia @ X
            parent.mass = parent.mass + child.mass
        for kind, H, prm, sq in child.collisions:
            parent.collisions.append((kind, p_H_c @ H, prm, sq))
        # joints whose parent link was removed are re-expressed in the lumped parent
        for jj in joints:
            if jj is not j and jj.parent == child.name:
                jj.parent = parent.name
                jj.pose = p_H_c @ jj.pose
        # frames attached to the removed link move too; the removed link becomes a frame
        frames = [(fn, parent.name, p_H_c @ fH) if fp == child.name else (fn, fp, fH) for fn, fp, fH in frames]
        frames.append((child.name, parent.name, p_H_c))
        joints.remove(j)
        del links[child.name]

    # ---- links with mass <= 0 cannot be simulated (parsers/rod/parser.py:110-119)
    for l in list(links.values()):
        if l.mass <= 0:
            raise ValueError(f"Link '{l.name}' has zero mass and is connected by a movable joint")

    # ---- tree
    for j in joints:
        links[j.child].parent = links[j.parent]
        links[j.parent].children.append(links[j.child])
    roots = [l for l in links.values() if l.parent is None]
    if len(roots) != 1:
        raise ValueError(f"The model must have exactly one root link, found {[r.name for r in roots]}")
    root = roots[0]
    if base_name is not None and base_name not in (root.name,):
        # the world joint pointed to a link that was lumped away -> it must be the root now
        pass

    # BFS with children sorted by name (parsers/kinematic_graph.py:669-709).
    order: list[_Link] = []
    queue = [root]
    while queue:
        l = queue.pop(0)
        l.index = len(order)
        order.append(l)
        queue.extend(sorted(l.children, key=lambda c: c.name))
    nL = len(order)
    joint_of_child = {j.child: j for j in joints}
    for j in joints:
        j.index = links[j.child].index
    ordered_joints = sorted(joints, key=lambda j: j.index)
    n = len(ordered_joints)
    assert n == nL - 1

    parent_array = np.array([-1] + [links[joint_of_child[l.name].parent].index for l in order[1:]], dtype=np.int64)

    lam_H_pre = np.tile(np.eye(4), (nL, 1, 1))
    suc_H_i = np.tile(np.eye(4), (nL, 1, 1))
    suc_H_i[0] = base_pose  # math/joint_model.py:78-83 (+ parser.py:192-197 for fixed base)
    if floating_base and link_poses is not None:
        suc_H_i[0] = link_poses[root.name]  # SDF: pose of the base link w.r.t. the model frame (joint_model.py:74-78)
    S = np.zeros((nL, 6))
    axes = np.zeros((n, 3))
    for j in ordered_joints:
        lam_H_pre[j.index] = j.pose
        suc_H_i[j.index] = j.suc  # math/joint_model.py:96-98
        axes[j.index - 1] = j.axis
        if j.jtype == JointType.Revolute:
            S[j.index, 3:6] = j.axis
        else:
            S[j.index, 0:3] = j.axis

    lp = LinkParameters.from_spatial_inertias(np.stack([l.inertia for l in order]))

    def arr(f):
        return np.array([f(j) for j in ordered_joints], dtype=float)

    jp = JointParameters(
        friction_static=arr(lambda j: j.friction_static),
        friction_viscous=arr(lambda j: j.friction_viscous),
        position_limits_min=arr(lambda j: min(j.position_limit)),
        position_limits_max=arr(lambda j: max(j.position_limit)),
        position_limit_spring=arr(lambda j: j.position_limit_spring),
        position_limit_damper=arr(lambda j: j.position_limit_damper),
    )

    # collidable points: ONE list of shapes in the order of the file; the shapes of lumped links keep their slot and
    # are only re-parented / transformed (parsers/rod/parser.py:303-345, parsers/descriptions/model.py:88-138,
    # api/kin_dyn_parameters.py:811-835) -- not grouped by link index
    bodies, points = [], []
    shapes = sorted(((sq, l.index, kind, H, prm) for l in order for kind, H, prm, sq in l.collisions), key=lambda t: t[0])
    for _, lidx, kind, H, prm in shapes:
        P = _box_points(prm, H) if kind == "box" else _sphere_points(prm, H)
        points.append(P)
        bodies.extend([lidx] * P.shape[0])
    cp = (
        ContactParameters(body=tuple(bodies), point=np.vstack(points), enabled=tuple(True for _ in bodies))
        if bodies
        else ContactParameters()
    )

    frames = sorted(frames, key=lambda f: f[0])
    fp = FrameParameters(
        name=tuple(f[0] for f in frames),
        body=tuple(links[f[1]].index for f in frames),
        transform=np.stack([f[2] for f in frames]) if frames else np.zeros((0, 4, 4)),
    )

    jm = JointModel(
        lam_H_pre=lam_H_pre,
        suc_H_i=suc_H_i,
        joint_dofs=tuple([6 if floating_base else 0] + [1] * n),
        joint_names=tuple(["world_to_base"] + [j.name for j in ordered_joints]),
        joint_types=tuple([int(JointType.Fixed)] + [j.jtype for j in ordered_joints]),
        joint_axis=axes,
    )

    kd = KinDynParameters(
        link_names=tuple(l.name for l in order),
        parent_array=parent_array,
        motion_subspaces=S,
        link_parameters=lp,
        joint_model=jm,
        joint_parameters=jp,
        contact_parameters=cp,
        frame_parameters=fp,
    )
    return name, kd, floating_base
