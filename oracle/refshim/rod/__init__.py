"""Stand-in for `rod` (the reference's URDF/SDF front end, `pyproject.toml: rod >= 0.3.3`), far enough for the
reference's OWN `jaxsim.parsers.rod.parser.extract_model_data` / `build_model_description` (parser.py:36-420) and
`jaxsim.parsers.rod.utils` (inertial -> 6D inertia :21-66, joint types :69-101, box / sphere collisions -> collidable
points :104-225) to run UNMODIFIED on a URDF string.

What the real stack does and this file restates: `rod.Sdf.load(urdf)` converts the URDF to SDF with sdformat and hands
out a tree of dataclasses (`Model`, `Link`, `Inertial`, `Joint`, `Axis`, `Collision`, `Geometry`, `Pose`, ...); after
`switch_frame_convention(FrameConvention.Urdf)` every joint pose is expressed in its parent link and every link pose is
the identity w.r.t. its parent joint -- which is what a URDF says in the first place, so this stand-in builds that tree
straight from the URDF elements:

    <link><inertial><origin xyz rpy/><mass value/><inertia ixx.. /></inertial>
          <collision><origin/><geometry><box size/>|<sphere radius/></geometry></collision></link>
    <joint type><origin/><parent link/><child link/><axis xyz/><limit lower upper effort velocity/>
           <dynamics damping friction/></joint>

Fixed joints are kept as joints (sdformat's `preserveFixedJoint`); the reference's `ModelDescription.build_model_from`
lumps them (parsers/kinematic_graph.py:379-611).  Only the attributes those two reference files read are provided.
Test infrastructure (fixture generation in the build container), never imported by the product.
"""

from __future__ import annotations

import dataclasses
import enum
import pathlib
import xml.etree.ElementTree as ET

import numpy as np

from . import urdf  # noqa: F401
from . import utils  # noqa: F401


class FrameConvention(enum.Enum):
    Urdf = enum.auto()
    Sdf = enum.auto()
    World = enum.auto()
    Model = enum.auto()


def _floats(text, n=None, default=None):
    if text is None:
        return default
    v = [float(x) for x in text.split()]
    assert n is None or len(v) == n, (text, n)
    return v


@dataclasses.dataclass
class Pose:
    """`<pose relative_to=...>x y z roll pitch yaw</pose>`; URDF `<origin xyz rpy>` has the same fixed-axis XYZ meaning:
    R = Rz(yaw) Ry(pitch) Rx(roll)."""

    pose: list
    relative_to: str | None = None

    def transform(self) -> np.ndarray:
        x, y, z, r, p, yw = self.pose
        cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(yw), np.sin(yw)
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        H = np.eye(4)
        H[0:3, 0:3] = Rz @ Ry @ Rx
        H[0:3, 3] = [x, y, z]
        return H


def _origin(elem, relative_to=None) -> Pose | None:
    o = None if elem is None else elem.find("origin")
    if o is None:
        return None
    return Pose(pose=_floats(o.get("xyz"), 3, [0.0, 0.0, 0.0]) + _floats(o.get("rpy"), 3, [0.0, 0.0, 0.0]), relative_to=relative_to)


@dataclasses.dataclass
class Inertia:
    ixx: float = 0.0
    iyy: float = 0.0
    izz: float = 0.0
    ixy: float | None = None
    ixz: float | None = None
    iyz: float | None = None


@dataclasses.dataclass
class Inertial:
    mass: float = 0.0
    inertia: Inertia = dataclasses.field(default_factory=Inertia)
    pose: Pose | None = None


@dataclasses.dataclass
class Box:
    size: list


@dataclasses.dataclass
class Sphere:
    radius: float


@dataclasses.dataclass
class Cylinder:
    radius: float
    length: float


@dataclasses.dataclass
class Mesh:
    uri: str
    scale: list | None = None


@dataclasses.dataclass
class Geometry:
    box: Box | None = None
    sphere: Sphere | None = None
    cylinder: Cylinder | None = None
    mesh: Mesh | None = None


@dataclasses.dataclass
class Collision:
    name: str
    geometry: Geometry
    pose: Pose | None = None


@dataclasses.dataclass
class Link:
    name: str
    inertial: Inertial
    pose: Pose | None = None
    collision: list = dataclasses.field(default_factory=list)

    def collisions(self) -> list:
        return list(self.collision)


@dataclasses.dataclass
class Xyz:
    xyz: list


@dataclasses.dataclass
class Limit:
    lower: float | None = None
    upper: float | None = None
    effort: float | None = None
    velocity: float | None = None
    stiffness: float | None = None
    dissipation: float | None = None


@dataclasses.dataclass
class Dynamics:
    damping: float | None = None
    friction: float | None = None


@dataclasses.dataclass
class Axis:
    xyz: Xyz | None = None
    limit: Limit | None = None
    dynamics: Dynamics | None = None


@dataclasses.dataclass
class Joint:
    name: str
    type: str
    parent: str
    child: str
    pose: Pose | None = None
    axis: Axis | None = None


@dataclasses.dataclass
class Frame:
    name: str
    attached_to: str
    pose: Pose | None = None


@dataclasses.dataclass
class Model:
    name: str
    link: list
    joint: list
    frame: list = dataclasses.field(default_factory=list)
    pose: Pose | None = None
    canonical_link: str | None = None
    sdf_convention: bool = False  # poses as written in an SDF file (switch_frame_convention re-expresses them)

    def links(self) -> list:
        return list(self.link)

    def joints(self) -> list:
        return list(self.joint)

    def frames(self) -> list:
        return list(self.frame)

    def is_fixed_base(self) -> bool:
        return any(j.parent == "world" for j in self.joint)

    def get_canonical_link(self) -> str:
        if self.canonical_link is not None:
            return self.canonical_link
        if self.is_fixed_base():
            return next(j.child for j in self.joint if j.parent == "world")
        children = {j.child for j in self.joint}
        roots = [l.name for l in self.link if l.name not in children]
        assert len(roots) == 1, roots
        return roots[0]

    def switch_frame_convention(self, frame_convention: FrameConvention, explicit_frames: bool = True, **_) -> None:
        if frame_convention is not FrameConvention.Urdf:
            raise NotImplementedError("the stand-in only converts to the URDF frame convention")
        if not self.sdf_convention:
            return  # built from a URDF: joint poses already in the parent link, identity link poses
        # rod re-expresses every pose without moving any frame: joints in their parent link, links in their parent joint
        # (the canonical link in the model frame), frames in the link they are attached to.  Poses in the model frame
        # first, from the SDF 1.7+ defaults: a link pose is given in the model frame, a joint pose in its CHILD link,
        # a frame pose in what it is attached to (else the model).
        raw = {}
        for l in self.link:
            raw[l.name] = (l.pose.transform() if l.pose is not None else np.eye(4), (l.pose.relative_to if l.pose is not None else None) or "__model__")
        for j in self.joint:
            raw[j.name] = (j.pose.transform() if j.pose is not None else np.eye(4), (j.pose.relative_to if j.pose is not None else None) or j.child)
        for f in self.frame:
            raw[f.name] = (f.pose.transform() if f.pose is not None else np.eye(4),
                           (f.pose.relative_to if f.pose is not None else None) or f.attached_to or "__model__")
        done = {"__model__": np.eye(4), "world": np.eye(4)}

        def in_model(name, depth=0):
            if name not in done:
                assert depth < 64, f"cyclic relative_to chain at {name}"
                H, rel = raw[name]
                done[name] = in_model(rel, depth + 1) @ H
            return done[name]

        link_names = {l.name for l in self.link}
        parent_joint = {j.child: j for j in self.joint}
        canonical = self.get_canonical_link()
        for j in self.joint:
            if j.parent == "world":
                j.pose = _pose_from_transform(in_model(j.name), relative_to="__model__")
            else:
                j.pose = _pose_from_transform(np.linalg.inv(in_model(j.parent)) @ in_model(j.name), relative_to=j.parent)
        for l in self.link:
            if l.name in parent_joint:
                pj = parent_joint[l.name]
                l.pose = _pose_from_transform(np.linalg.inv(in_model(pj.name)) @ in_model(l.name), relative_to=pj.name)
            else:
                assert l.name == canonical, (l.name, canonical)
                l.pose = _pose_from_transform(in_model(l.name), relative_to="__model__")
        joint_child = {j.name: j.child for j in self.joint}
        frames_by_name = {f.name: f for f in self.frame}
        for f in self.frame:
            target = f.attached_to
            while target is not None and target not in link_names:  # attached to a frame / a joint: follow to a link
                target = frames_by_name[target].attached_to if target in frames_by_name else joint_child.get(target)
            if target is None:
                continue
            f.pose = _pose_from_transform(np.linalg.inv(in_model(target)) @ in_model(f.name), relative_to=target)
            f.attached_to = target
        self.sdf_convention = False


def _link_from_urdf(e) -> Link:
    ine = e.find("inertial")
    inertial = Inertial()
    if ine is not None:
        m = ine.find("mass")
        I = ine.find("inertia")
        g = (lambda k: float(I.get(k)) if (I is not None and I.get(k) is not None) else None)  # noqa: E731
        inertial = Inertial(
            mass=float(m.get("value")) if m is not None else 0.0,
            inertia=Inertia(ixx=g("ixx") or 0.0, iyy=g("iyy") or 0.0, izz=g("izz") or 0.0, ixy=g("ixy"), ixz=g("ixz"), iyz=g("iyz")),
            pose=_origin(ine),
        )
    cols = []
    for k, c in enumerate(e.findall("collision")):
        ge = c.find("geometry")
        geo = Geometry()
        if ge is not None:
            if ge.find("box") is not None:
                geo.box = Box(size=_floats(ge.find("box").get("size"), 3))
            elif ge.find("sphere") is not None:
                geo.sphere = Sphere(radius=float(ge.find("sphere").get("radius")))
            elif ge.find("cylinder") is not None:
                geo.cylinder = Cylinder(radius=float(ge.find("cylinder").get("radius")), length=float(ge.find("cylinder").get("length")))
            elif ge.find("mesh") is not None:
                geo.mesh = Mesh(uri=ge.find("mesh").get("filename"), scale=_floats(ge.find("mesh").get("scale")))
        cols.append(Collision(name=c.get("name") or f"{e.get('name')}_collision_{k}", geometry=geo, pose=_origin(c)))
    return Link(name=e.get("name"), inertial=inertial, pose=None, collision=cols)


def _joint_from_urdf(e) -> Joint:
    jt = e.get("type")
    parent, child = e.find("parent").get("link"), e.find("child").get("link")
    # a fixed joint carries the URDF default axis (1 0 0): the reference hashes joint descriptions and a None axis is
    # not hashable (parsers/descriptions/joint.py:108-114); its value is never used (utils.py:85-86 returns Fixed first)
    axis = Axis(xyz=Xyz(xyz=[1.0, 0.0, 0.0])) if jt == "fixed" else None
    if jt in ("revolute", "continuous", "prismatic"):
        ax = e.find("axis")
        lim, dyn = e.find("limit"), e.find("dynamics")
        f = (lambda el, k: float(el.get(k)) if (el is not None and el.get(k) is not None) else None)  # noqa: E731
        # a continuous joint has no position limits (URDF); <limit effort velocity> may still be present
        limit = Limit(lower=None if jt == "continuous" else f(lim, "lower"), upper=None if jt == "continuous" else f(lim, "upper"),
                      effort=f(lim, "effort"), velocity=f(lim, "velocity")) if lim is not None else None
        dynamics = Dynamics(damping=f(dyn, "damping"), friction=f(dyn, "friction")) if dyn is not None else None
        axis = Axis(xyz=Xyz(xyz=_floats(ax.get("xyz"), 3) if ax is not None else [1.0, 0.0, 0.0]), limit=limit, dynamics=dynamics)
    # joint pose = <origin>, expressed in the parent link (the model frame for joints attached to the world)
    pose = _origin(e, relative_to=None if parent == "world" else parent) or Pose(pose=[0.0] * 6, relative_to=None if parent == "world" else parent)
    return Joint(name=e.get("name"), type=jt, parent=parent, child=child, pose=pose, axis=axis)


class Sdf:
    def __init__(self, model=None, version: str = "1.7"):
        self.model = model if isinstance(model, list) else ([model] if model is not None else [])
        self.version = version

    def models(self) -> list:
        return list(self.model)

    @staticmethod
    def load(sdf, is_urdf: bool | None = None) -> "Sdf":
        text = sdf
        if isinstance(sdf, pathlib.Path) or (isinstance(sdf, str) and len(sdf) < 1024 and "<" not in sdf):
            text = pathlib.Path(sdf).read_text()
        root = ET.fromstring(text)
        if root.tag == "sdf":
            return Sdf(model=[_model_from_sdf(m) for m in root.findall("model")], version=root.get("version", "1.7"))
        if root.tag != "robot":
            raise NotImplementedError("the rod stand-in reads URDF (<robot>) and SDF (<sdf>) documents")
        links = [_link_from_urdf(e) for e in root.findall("link") if e.get("name") != "world"]
        joints = [_joint_from_urdf(e) for e in root.findall("joint")]
        # sdformat turns a massless link that hangs on a FIXED joint into a frame attached to the parent link, with the
        # joint's pose (what `model.frame_names()` lists in the reference, e.g. the `*_frame` links of
        # tests/assets/4_bar_opened.urdf); frames of frames are re-attached to the first real link
        frames = []
        massless = {l.name for l in links if not (l.inertial.mass > 0)}
        changed = True
        while changed:
            changed = False
            for j in list(joints):
                if j.type == "fixed" and j.child in massless and j.parent != "world" and not any(jj.parent == j.child for jj in joints):
                    frames.append(Frame(name=j.child, attached_to=j.parent, pose=Pose(pose=list(j.pose.pose), relative_to=j.parent)))
                    joints.remove(j)
                    links = [l for l in links if l.name != j.child]
                    changed = True
        real = {l.name for l in links}
        for f in frames:  # a frame attached to another frame: compose the poses down to a real link
            while f.attached_to not in real:
                parent = next(g for g in frames if g.name == f.attached_to)
                H = parent.pose.transform() @ f.pose.transform()
                f.attached_to = parent.attached_to
                f.pose = _pose_from_transform(H, relative_to=parent.attached_to)
        return Sdf(model=Model(name=root.get("name"), link=links, joint=joints, frame=frames))


def _sdf_pose(e) -> Pose | None:
    pe = None if e is None else e.find("pose")
    if pe is None:
        return None
    return Pose(pose=_floats(pe.text, None, None) or [0.0] * 6, relative_to=pe.get("relative_to") or None)


def _sdf_float(e, path):
    x = None if e is None else e.find(path)
    return float(x.text) if (x is not None and x.text is not None and x.text.strip()) else None


def _model_from_sdf(me) -> Model:
    """rod's dataclass tree straight from the SDF elements (rod is SDF-native: element names == attribute names)."""
    links = []
    for le in me.findall("link"):
        ine = le.find("inertial")
        inertial = Inertial()
        if ine is not None:
            g = lambda k: _sdf_float(ine, f"inertia/{k}")  # noqa: E731
            inertial = Inertial(mass=_sdf_float(ine, "mass") or 0.0,
                                inertia=Inertia(ixx=g("ixx") or 0.0, iyy=g("iyy") or 0.0, izz=g("izz") or 0.0, ixy=g("ixy"), ixz=g("ixz"), iyz=g("iyz")),
                                pose=_sdf_pose(ine))
        cols = []
        for k, c in enumerate(le.findall("collision")):
            ge, geo = c.find("geometry"), Geometry()
            if ge is not None:
                if ge.find("box") is not None:
                    geo.box = Box(size=_floats(ge.find("box/size").text, 3))
                elif ge.find("sphere") is not None:
                    geo.sphere = Sphere(radius=float(ge.find("sphere/radius").text))
                elif ge.find("cylinder") is not None:
                    geo.cylinder = Cylinder(radius=float(ge.find("cylinder/radius").text), length=float(ge.find("cylinder/length").text))
                elif ge.find("mesh") is not None:
                    geo.mesh = Mesh(uri=ge.find("mesh/uri").text, scale=_floats(ge.find("mesh/scale").text) if ge.find("mesh/scale") is not None else None)
            cols.append(Collision(name=c.get("name") or f"{le.get('name')}_collision_{k}", geometry=geo, pose=_sdf_pose(c)))
        links.append(Link(name=le.get("name"), inertial=inertial, pose=_sdf_pose(le), collision=cols))
    joints = []
    for je in me.findall("joint"):
        ax, axis = je.find("axis"), None
        if ax is not None:
            xyz = ax.find("xyz")
            lim, dyn = ax.find("limit"), ax.find("dynamics")
            axis = Axis(xyz=Xyz(xyz=_floats(xyz.text, 3)) if xyz is not None else None,
                        limit=Limit(lower=_sdf_float(lim, "lower"), upper=_sdf_float(lim, "upper"), effort=_sdf_float(lim, "effort"),
                                    velocity=_sdf_float(lim, "velocity"), stiffness=_sdf_float(lim, "stiffness"),
                                    dissipation=_sdf_float(lim, "dissipation")) if lim is not None else None,
                        dynamics=Dynamics(damping=_sdf_float(dyn, "damping"), friction=_sdf_float(dyn, "friction")) if dyn is not None else None)
        elif je.get("type") == "fixed":
            axis = Axis(xyz=Xyz(xyz=[1.0, 0.0, 0.0]))  # hashable stand-in, never used (see _joint_from_urdf)
        joints.append(Joint(name=je.get("name"), type=je.get("type"), parent=je.find("parent").text.strip(), child=je.find("child").text.strip(),
                            pose=_sdf_pose(je), axis=axis))
    frames = [Frame(name=fe.get("name"), attached_to=fe.get("attached_to"), pose=_sdf_pose(fe)) for fe in me.findall("frame")]
    return Model(name=me.get("name"), link=links, joint=joints, frame=frames, pose=_sdf_pose(me), canonical_link=me.get("canonical_link"),
                 sdf_convention=True)


def _pose_from_transform(H, relative_to=None) -> Pose:
    R = H[0:3, 0:3]
    pitch = np.arcsin(-np.clip(R[2, 0], -1.0, 1.0))
    roll = np.arctan2(R[2, 1], R[2, 2])
    yaw = np.arctan2(R[1, 0], R[0, 0])
    return Pose(pose=[float(H[0, 3]), float(H[1, 3]), float(H[2, 3]), float(roll), float(pitch), float(yaw)], relative_to=relative_to)
