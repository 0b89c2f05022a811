"""SDF (``<sdf><model>``) front end of the loader: pose semantics -> the joint / link description that
``parsers/urdf.py: build_kin_dyn_parameters`` turns into :class:`KinDynParameters`.

The reference reads SDF through ``rod`` and then asks it for the URDF frame convention
(``parsers/rod/parser.py:83-87``: ``switch_frame_convention(FrameConvention.Urdf)``): every joint pose expressed in its
parent link, every link pose expressed in its parent joint, frames in the link they are attached to.  ``rod`` is not
available offline, so this module resolves the poses itself, from the SDF 1.7+ rules:

* ``<pose relative_to="F">x y z roll pitch yaw</pose>`` is expressed in frame ``F``; the default ``F`` is the model
  frame for a ``<link>``, the CHILD link for a ``<joint>``, the ``attached_to`` frame (else the model) for a
  ``<frame>``, and the owning link for ``<inertial>`` / ``<collision>`` (the reference reads those two poses raw,
  ``parsers/rod/utils.py:37-38,120-121``: any other ``relative_to`` there raises here);
* any link, joint or frame name is a frame; ``__model__`` and ``world`` are the model frame.

With ``M_H_X`` the pose of ``X`` in the model frame the joint model of the reference (``math/joint_model.py:74-98``)
is ``lam_H_pre = (M_H_parent)^-1 M_H_joint``, ``suc_H_i = (M_H_joint)^-1 M_H_child`` and, for a floating base,
``suc_H_i[0] = M_H_base``; a fixed joint from ``world`` makes the model fixed-base with
``suc_H_i[0] = pose(joint) @ pose(base link w.r.t. the joint)`` (``parser.py:147-197``).  The axis is the raw
``<axis><xyz>`` (``parser.py:224-231``; ``expressed_in`` is not read by the reference either), joint parameters follow
``parser.py:234-277`` (``<limit><stiffness>/<dissipation>`` are the position-limit spring / damper).  Inertials and
box / sphere collisions are turned into 6D inertias and collidable points exactly like the URDF path.
"""

from __future__ import annotations

import os
import xml.etree.ElementTree as ET

import numpy as np

from jaxsim_b200.api.kin_dyn_parameters import JointType

from . import urdf as _u


def _pose(elem):
    """(4x4 transform, relative_to or None) of the ``<pose>`` child of ``elem``."""
    H = np.eye(4)
    if elem is None:
        return H, None
    pe = elem.find("pose")
    if pe is None:
        return H, None
    v = [float(x) for x in (pe.text or "").split()]
    if len(v) != 6:
        if len(v) == 0:
            return H, pe.get("relative_to") or None
        raise ValueError(f"<pose> needs 6 numbers, got {pe.text!r}")
    H[0:3, 0:3] = _u._rpy_to_R(v[3:6])
    H[0:3, 3] = v[0:3]
    return H, (pe.get("relative_to") or None)


def _text(elem, path, default=None):
    e = elem.find(path) if elem is not None else None
    return e.text.strip() if (e is not None and e.text is not None) else default


def parse(xml_text: str):
    """-> ``(name, links, joints, frames, link_poses)`` with ``links`` / ``joints`` in the loader's URDF-convention
    records (``_Joint.pose`` = parent_H_joint, ``_Joint.suc`` = joint_H_child), ``frames`` = ``(name, link,
    link_H_frame)`` and ``link_poses[name]`` = pose of the link in the model frame."""
    root = ET.fromstring(xml_text)
    if root.tag != "sdf":
        raise ValueError("expected an <sdf> document")
    models = root.findall("model")
    if len(models) != 1:
        raise ValueError(f"expected exactly one <model>, found {len(models)}")  # parser.py:67-75 (model_name selection not needed here)
    me = models[0]
    name = me.get("name", "model")

    # ---- every named frame with its raw pose and the frame it is expressed in
    raw: dict[str, tuple[np.ndarray, str]] = {}
    for le in me.findall("link"):
        H, rel = _pose(le)
        raw[le.get("name")] = (H, rel or "__model__")
    for je in me.findall("joint"):
        H, rel = _pose(je)
        raw[je.get("name")] = (H, rel or _text(je, "child"))
    for fe in me.findall("frame"):
        H, rel = _pose(fe)
        raw[fe.get("name")] = (H, rel or fe.get("attached_to") or "__model__")

    resolved: dict[str, np.ndarray] = {"__model__": np.eye(4), "world": np.eye(4)}

    def M_H(frame: str, stack=()) -> np.ndarray:
        if frame in resolved:
            return resolved[frame]
        if frame not in raw:
            raise ValueError(f"pose expressed in an unknown frame '{frame}'")
        if frame in stack:
            raise ValueError(f"cyclic relative_to chain through '{frame}'")
        H, rel = raw[frame]
        resolved[frame] = M_H(rel, stack + (frame,)) @ H
        return resolved[frame]

    # ---- links: 6D inertia about the link frame, collision shapes (same conventions as the URDF path)
    links: dict[str, _u._Link] = {}
    seq = 0
    for le in me.findall("link"):
        lname = le.get("name")
        ine = le.find("inertial")
        mass, M = 0.0, np.zeros((6, 6))
        if ine is not None:
            mass = float(_text(ine, "mass", "0"))
            g = lambda k: float(_text(ine, f"inertia/{k}", "0"))  # noqa: E731
            I_com = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            L_H_com, rel = _pose(ine)
            if rel not in (None, lname):  # parsers/rod/utils.py:37-38 reads the raw pose: only the owning link is a defined meaning
                raise NotImplementedError(f"<inertial> of link '{lname}' posed relative_to '{rel}'")
            X = _u._adjoint_inverse(L_H_com)  # CoM_X_L (parsers/rod/utils.py:21-66)
            M = X.T @ _u._sixd_inertia(mass, np.zeros(3), I_com) @ X
        cols = []
        for ce in le.findall("collision"):
            L_H_c, rel = _pose(ce)
            if rel not in (None, lname):  # parsers/rod/utils.py:120-121, 178-179 read the raw pose
                raise NotImplementedError(f"<collision> of link '{lname}' posed relative_to '{rel}'")
            geo = ce.find("geometry")
            if geo is None:
                continue
            if geo.find("box") is not None:
                cols.append(("box", L_H_c, [float(v) for v in _text(geo, "box/size").split()], seq))
                seq += 1
            elif geo.find("sphere") is not None:
                cols.append(("sphere", L_H_c, float(_text(geo, "sphere/radius")), seq))
                seq += 1
            # cylinder / capsule / mesh: ignored like in the reference defaults (parser.py:334-357)
        links[lname] = _u._Link(name=lname, mass=mass, inertia=M, collisions=cols)

    # ---- joints
    joints: list[_u._Joint] = []
    fmax = float(np.finfo(float).max)
    for je in me.findall("joint"):
        jt = je.get("type")
        if jt in ("revolute", "continuous"):
            jtype = JointType.Revolute
        elif jt == "prismatic":
            jtype = JointType.Prismatic
        elif jt == "fixed":
            jtype = JointType.Fixed
        else:
            raise ValueError(f"Joint type '{jt}' not supported")
        parent, child = _text(je, "parent"), _text(je, "child")
        ax = je.find("axis")
        xyz = _text(ax, "xyz")
        axis = np.array([float(v) for v in xyz.split()]) if xyz is not None else np.array([1.0, 0.0, 0.0])
        if jtype != JointType.Fixed:
            if xyz is None:
                raise ValueError("Failed to read axis xyz data")  # parsers/rod/utils.py:86-87
            axis = axis / np.linalg.norm(axis)
        f = lambda path: (float(_text(ax, path)) if _text(ax, path) is not None else None)  # noqa: E731
        lo, hi = f("limit/lower"), f("limit/upper")
        damper, spring = f("limit/dissipation"), f("limit/stiffness")
        M_H_J = M_H(je.get("name"))
        M_H_P = M_H(parent) if parent != "world" else np.eye(4)
        joints.append(
            _u._Joint(
                name=je.get("name"),
                jtype=int(jtype),
                parent=parent,
                child=child,
                pose=np.linalg.inv(M_H_P) @ M_H_J,
                axis=axis,
                position_limit=(-fmax if lo is None else lo, fmax if hi is None else hi),
                friction_static=f("dynamics/friction") or 0.0,
                friction_viscous=f("dynamics/damping") or 0.0,
                position_limit_damper=float(os.environ.get("JAXSIM_JOINT_POSITION_LIMIT_DAMPER", 0.0)) if damper is None else damper,
                position_limit_spring=float(os.environ.get("JAXSIM_JOINT_POSITION_LIMIT_SPRING", 0.0)) if spring is None else spring,
                suc=np.linalg.inv(M_H_J) @ M_H(child),
            )
        )

    # ---- explicit <frame>s, expressed in the LINK they are (transitively) attached to (parser.py:121-139)
    frames: list[tuple[str, str, np.ndarray]] = []
    attached = {fe.get("name"): fe.get("attached_to") for fe in me.findall("frame")}
    joint_child = {je.get("name"): _text(je, "child") for je in me.findall("joint")}
    for fname, target in attached.items():
        seen = set()
        while target is not None and target not in links:
            if target in seen:
                raise ValueError(f"cyclic attached_to chain through '{target}'")
            seen.add(target)
            target = attached.get(target, joint_child.get(target))  # a frame attached to a frame / to a joint (= its child link)
        if target is None:
            continue  # attached to the model frame: not a link frame, the reference drops it (parser.py:137-138)
        frames.append((fname, target, np.linalg.inv(M_H(target)) @ M_H(fname)))

    link_poses = {lname: M_H(lname) for lname in links}
    return name, links, joints, frames, link_poses
