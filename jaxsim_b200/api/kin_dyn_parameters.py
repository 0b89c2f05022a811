"""Host-side kinematic/dynamic parameters of a model (NumPy, fp64).

Mirrors the *data contract* of the reference's ``KinDynParameters``
(``src/jaxsim/api/kin_dyn_parameters.py:21-63``) -- same field names, same
indexing conventions -- so that the model blob uploaded to the GPU is a
field-for-field image of what ``jaxsim.api.model.step`` consumes:

* link index = BFS order from the base link, children sorted by name
  (``parsers/kinematic_graph.py:128-134,669-709``);
* joint index = child link index, DoF index = joint index - 1
  (``parsers/kinematic_graph.py:158-171``, ``rbda/aba.py:133``);
* ``parent_array[0] == -1`` (``api/kin_dyn_parameters.py:193-198``);
* 6D quantities are ``[linear; angular]``.

This is host code executed once per model; nothing here is on the hot path.
"""

from __future__ import annotations

import dataclasses
import enum

import numpy as np


class JointType(enum.IntEnum):
    """Joint types (``parsers/descriptions/joint.py`` ``JointType``)."""

    Fixed = 0
    Revolute = 1
    Prismatic = 2


@dataclasses.dataclass
class LinkParameters:
    """``api/kin_dyn_parameters.py:574-598``: mass, CoM (link frame) and the six
    upper-triangular elements of the inertia tensor *at the CoM* in link axes."""

    mass: np.ndarray  # (nL,)
    center_of_mass: np.ndarray  # (nL, 3)
    inertia_elements: np.ndarray  # (nL, 6)  triu order xx, xy, xz, yy, yz, zz

    @staticmethod
    def from_spatial_inertias(M: np.ndarray) -> "LinkParameters":
        """``LinkParameters.build_from_spatial_inertia`` (``:600-626``) +
        ``Inertia.to_params`` (``math/inertia.py:43-63``)."""
        M = np.asarray(M, dtype=float).reshape(-1, 6, 6)
        m = np.trace(M[:, 0:3, 0:3], axis1=1, axis2=2) / 3.0
        mC = M[:, 3:6, 0:3]
        # vee(mC) / m
        c = 0.5 * np.stack(
            [mC[:, 2, 1] - mC[:, 1, 2], mC[:, 0, 2] - mC[:, 2, 0], mC[:, 1, 0] - mC[:, 0, 1]],
            axis=-1,
        ) / m[:, None]
        I = M[:, 3:6, 3:6] - np.einsum("bij,bkj->bik", mC, mC) / m[:, None, None]
        iu = np.triu_indices(3)
        return LinkParameters(mass=m, center_of_mass=c, inertia_elements=I[:, iu[0], iu[1]])

    def inertia_tensors(self) -> np.ndarray:
        """``LinkParameters.unflatten_inertia_tensor`` (``:747-762``)."""
        nL = self.mass.shape[0]
        I = np.zeros((nL, 3, 3))
        iu = np.triu_indices(3)
        I[:, iu[0], iu[1]] = self.inertia_elements
        I[:, iu[1], iu[0]] = self.inertia_elements
        return I

    def spatial_inertias(self) -> np.ndarray:
        """``Inertia.to_sixd`` (``math/inertia.py:14-41``) for every link."""
        nL = self.mass.shape[0]
        I = self.inertia_tensors()
        M = np.zeros((nL, 6, 6))
        for i in range(nL):
            c = _wedge(self.center_of_mass[i])
            m = self.mass[i]
            M[i, 0:3, 0:3] = m * np.eye(3)
            M[i, 0:3, 3:6] = m * c.T
            M[i, 3:6, 0:3] = m * c
            M[i, 3:6, 3:6] = I[i] + m * c @ c.T
        return M


@dataclasses.dataclass
class JointParameters:
    """``api/kin_dyn_parameters.py:502-571`` (all arrays have shape ``(n,)``)."""

    friction_static: np.ndarray
    friction_viscous: np.ndarray
    position_limits_min: np.ndarray
    position_limits_max: np.ndarray
    position_limit_spring: np.ndarray
    position_limit_damper: np.ndarray


@dataclasses.dataclass
class JointModel:
    """``math/joint_model.py:16-113``: fixed transforms around each joint.

    Entry 0 is the dummy world->base joint: ``lam_H_pre[0] = I`` and
    ``suc_H_i[0]`` stores the optional root->base-link pose (``:78-83``).
    """

    lam_H_pre: np.ndarray  # (nL, 4, 4)  parent link -> predecessor frame
    suc_H_i: np.ndarray  # (nL, 4, 4)  successor frame -> child link
    joint_dofs: tuple[int, ...]  # joint_dofs[0] is 6 (floating) or 0 (fixed base)
    joint_names: tuple[str, ...]  # joint_names[0] == "world_to_base"
    joint_types: tuple[int, ...]  # joint_types[0] == JointType.Fixed
    joint_axis: np.ndarray  # (n, 3)


@dataclasses.dataclass
class ContactParameters:
    """``api/kin_dyn_parameters.py:765-840``."""

    body: tuple[int, ...] = ()
    point: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0, 3)))
    enabled: tuple[bool, ...] = ()

    @property
    def indices_of_enabled_collidable_points(self) -> np.ndarray:
        return np.where(np.array(self.enabled, dtype=bool))[0]


@dataclasses.dataclass
class FrameParameters:
    """``api/kin_dyn_parameters.py:843-917`` (names, parent link, L_H_F)."""

    name: tuple[str, ...] = ()
    body: tuple[int, ...] = ()
    transform: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros((0, 4, 4)))


class ConstraintType:
    """``api/kin_dyn_parameters.py:1248-1255``: the reference implements the weld constraint only."""

    Weld = 0


@dataclasses.dataclass(frozen=True)
class ConstraintMap:
    """Kinematic constraints between pairs of frames (``api/kin_dyn_parameters.py:1258-1350``).  Immutable like the
    reference's: ``add_constraint`` returns a new map.  Frame indices count after the links (``api/frame.py``)."""

    frame_idxs_1: tuple[int, ...] = ()
    frame_idxs_2: tuple[int, ...] = ()
    constraint_types: tuple[int, ...] = ()
    K_P: tuple[float, ...] = ()
    K_D: tuple[float, ...] = ()
    parent_link_idxs_1: tuple[int, ...] = ()
    parent_link_idxs_2: tuple[int, ...] = ()

    def add_constraint(self, model, frame_idx_1: int, frame_idx_2: int, constraint_type: int,
                       K_P: float | None = None, K_D: float | None = None) -> "ConstraintMap":
        """``ConstraintMap.add_constraint`` (``:1287-1350``): Baumgarte gains default to K_P = 1000 and
        K_D = 2 sqrt(K_P) (critical damping)."""
        from . import frame as _frame

        if constraint_type != ConstraintType.Weld:
            raise NotImplementedError("only ConstraintType.Weld exists in the reference (kin_dyn_parameters.py:1253-1255)")
        K_P = 1000.0 if K_P is None else float(K_P)
        K_D = 2.0 * float(np.sqrt(K_P)) if K_D is None else float(K_D)
        l1 = _frame.idx_of_parent_link(model, frame_index=frame_idx_1)
        l2 = _frame.idx_of_parent_link(model, frame_index=frame_idx_2)
        return ConstraintMap(
            frame_idxs_1=self.frame_idxs_1 + (int(frame_idx_1),), frame_idxs_2=self.frame_idxs_2 + (int(frame_idx_2),),
            constraint_types=self.constraint_types + (int(constraint_type),),
            K_P=self.K_P + (K_P,), K_D=self.K_D + (K_D,),
            parent_link_idxs_1=self.parent_link_idxs_1 + (int(l1),), parent_link_idxs_2=self.parent_link_idxs_2 + (int(l2),),
        )

    def __len__(self) -> int:
        return len(self.frame_idxs_1)


@dataclasses.dataclass
class KinDynParameters:
    """``api/kin_dyn_parameters.py:21-63``."""

    link_names: tuple[str, ...]
    parent_array: np.ndarray  # (nL,) int, parent_array[0] = -1
    motion_subspaces: np.ndarray  # (nL, 6), row 0 zeros
    link_parameters: LinkParameters
    joint_model: JointModel
    joint_parameters: JointParameters
    contact_parameters: ContactParameters
    frame_parameters: FrameParameters = dataclasses.field(default_factory=FrameParameters)
    # kinematic constraints (``:62-63``); None / empty = unconstrained
    constraints: ConstraintMap | None = None

    def number_of_links(self) -> int:
        return len(self.link_names)

    def number_of_joints(self) -> int:
        return len(self.joint_model.joint_names) - 1

    def number_of_frames(self) -> int:
        return len(self.frame_parameters.name)

    @property
    def support_body_array_bool(self) -> np.ndarray:
        """kappa_b(i) of ``api/kin_dyn_parameters.py:200-237``."""
        nL = self.number_of_links()
        k = np.zeros((nL, nL), dtype=bool)
        for i in range(nL):
            j = i
            while j >= 0:
                k[i, j] = True
                j = int(self.parent_array[j])
        return k

    def levels(self) -> list[list[int]]:
        """Links grouped by tree depth (level 0 = base link)."""
        depth = np.zeros(self.number_of_links(), dtype=int)
        for i in range(1, self.number_of_links()):
            depth[i] = depth[self.parent_array[i]] + 1
        return [list(np.where(depth == d)[0]) for d in range(depth.max() + 1)]


def _wedge(v: np.ndarray) -> np.ndarray:
    x, y, z = np.asarray(v, dtype=float).reshape(3)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])
