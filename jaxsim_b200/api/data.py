"""``JaxSimModelData``: the state of (a batch of) environments as torch device tensors.

Mirrors ``src/jaxsim/api/data.py``: same leaf names (``:47-63``), ``build`` (``:66-202``),
``zero`` (``:204-222``), accessors (``:228-349``), ``replace`` (``:406-523``) and
``random_model_data`` (``:552-682``).  Batching is a leading axis on every leaf (what
``jax.vmap(..., in_axes=(None, 0))`` gives the reference); unbatched leaves are accepted too.
The cached transforms are computed by the ``b200sim_fk`` kernel.  Objects are immutable
in spirit: every operation returns a new object (``utils/jaxsim_dataclass.py:313-334``).
"""

from __future__ import annotations

import contextlib
import ctypes as C
import dataclasses
import math

import numpy as np
import torch

from jaxsim_b200 import _lib

from .common import VelRepr, inertial_to_other_representation, other_representation_to_inertial


@dataclasses.dataclass
class JaxSimModelData:
    """``src/jaxsim/api/data.py:26-63``."""

    velocity_representation: VelRepr = VelRepr.Mixed

    # Joint state
    _joint_positions: torch.Tensor = None
    _joint_velocities: torch.Tensor = None
    # Base state (velocities ALWAYS inertial-fixed, api/data.py:36-39)
    _base_quaternion: torch.Tensor = None
    _base_linear_velocity: torch.Tensor = None
    _base_angular_velocity: torch.Tensor = None
    _base_position: torch.Tensor = None
    # Cached computations
    _base_transform: torch.Tensor = dataclasses.field(repr=False, default=None)
    _joint_transforms: torch.Tensor = dataclasses.field(repr=False, default=None)
    _link_transforms: torch.Tensor = dataclasses.field(repr=False, default=None)
    _link_velocities: torch.Tensor = dataclasses.field(repr=False, default=None)
    # Extended state of the contact model
    contact_state: dict = dataclasses.field(default_factory=dict)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def build(
        model,
        base_position=None,
        base_quaternion=None,
        joint_positions=None,
        base_linear_velocity=None,
        base_angular_velocity=None,
        joint_velocities=None,
        contact_state: dict | None = None,
        velocity_representation: VelRepr = VelRepr.Mixed,
        *,
        batch_size: int | None = None,
        dtype: torch.dtype = torch.float64,
        device: torch.device | str = "cuda",
    ) -> "JaxSimModelData":
        """``JaxSimModelData.build`` (``api/data.py:66-202``).  Base velocities are given in
        ``velocity_representation`` and stored inertial-fixed (``:151-156``)."""
        from jaxsim_b200.rbda.contacts import SoftContacts

        device = torch.device(device)
        n = model.dofs()
        given = [base_position, base_quaternion, joint_positions, base_linear_velocity, base_angular_velocity, joint_velocities]
        widths = [3, 4, n, 3, 3, n]
        tens = [None if g is None else torch.as_tensor(g, dtype=dtype, device=device) for g in given]
        batched_B = [t.shape[0] for t in tens if t is not None and t.dim() == 2]
        B = batch_size if batch_size is not None else (batched_B[0] if batched_B else None)
        unbatched = B is None
        Bn = 1 if unbatched else B

        def canon(t, w, default):
            if t is None:
                t = torch.as_tensor(default, dtype=dtype, device=device)
            if t.dim() == 1:
                t = t.unsqueeze(0).expand(Bn, w)
            if t.shape != (Bn, w):
                raise ValueError(t.shape, (Bn, w))
            return t.contiguous()

        p = canon(tens[0], 3, [0.0, 0.0, 0.0])
        q = canon(tens[1], 4, [1.0, 0.0, 0.0, 0.0])
        s = canon(tens[2], n, [0.0] * n)
        vl = canon(tens[3], 3, [0.0, 0.0, 0.0])
        om = canon(tens[4], 3, [0.0, 0.0, 0.0])
        sd = canon(tens[5], n, [0.0] * n)

        if velocity_representation != VelRepr.Inertial:
            W_H_B = _transform_from_quat_pos(q, p)
            v6 = other_representation_to_inertial(
                torch.cat([vl, om], dim=-1), velocity_representation, W_H_B, is_force=False
            )
            vl, om = v6[:, 0:3].contiguous(), v6[:, 3:6].contiguous()

        cs = dict(contact_state or {})
        if isinstance(model.contact_model, SoftContacts):
            nc = model.number_of_collidable_points()
            if "tangential_deformation" not in cs:
                cs["tangential_deformation"] = torch.zeros(Bn, nc, 3, dtype=dtype, device=device)
            else:
                t = torch.as_tensor(cs["tangential_deformation"], dtype=dtype, device=device)
                cs["tangential_deformation"] = (t if t.dim() == 3 else t.unsqueeze(0).expand(Bn, nc, 3)).contiguous()

        out = _with_caches(model, velocity_representation, s, sd, q, vl, om, p, cs, normalise_q=False)
        return out._squeezed() if unbatched else out

    @staticmethod
    def zero(model, velocity_representation: VelRepr = VelRepr.Mixed, **kw) -> "JaxSimModelData":
        """``api/data.py:204-222``."""
        return JaxSimModelData.build(model=model, velocity_representation=velocity_representation, **kw)

    # ------------------------------------------------------------------ accessors
    @property
    def joint_positions(self):
        return self._joint_positions

    @property
    def joint_velocities(self):
        return self._joint_velocities

    @property
    def base_quaternion(self):
        return self._base_quaternion

    @property
    def base_position(self):
        return self._base_position

    @property
    def base_orientation(self):
        """Normalised quaternion (``api/data.py:267-286``)."""
        q = self._base_quaternion
        norm = torch.linalg.norm(q, dim=-1, keepdim=True)
        return q / (norm + torch.finfo(q.dtype).eps * (norm == 0))

    @property
    def base_velocity(self):
        """6D base velocity in the active representation (``api/data.py:288-312``)."""
        W_v = torch.cat([self._base_linear_velocity, self._base_angular_velocity], dim=-1)
        return inertial_to_other_representation(W_v, self.velocity_representation, self.base_transform, is_force=False)

    @property
    def generalized_position(self):
        return self.base_transform, self.joint_positions

    @property
    def generalized_velocity(self):
        return torch.cat([self.base_velocity, self.joint_velocities], dim=-1)

    def _cache(self, name):
        v = getattr(self, name)
        if v is None:
            raise RuntimeError(
                f"{name} was not materialised (step(..., update_caches=False)); call data.replace(model) to compute it"
            )
        return v

    @property
    def base_transform(self):
        return self._cache("_base_transform")

    @property
    def joint_transforms(self):
        return self._cache("_joint_transforms")

    @property
    def link_transforms(self):
        return self._cache("_link_transforms")

    @property
    def link_velocities(self):
        return self._cache("_link_velocities")

    def batch_size(self) -> int | None:
        return None if self._base_quaternion.dim() == 1 else self._base_quaternion.shape[0]

    # ------------------------------------------------------------------ representation switch
    @contextlib.contextmanager
    def switch_velocity_representation(self, velocity_representation: VelRepr):
        """``api/common.py:60-98`` (the state is repr-independent; only accessors change)."""
        original = self.velocity_representation
        try:
            self.velocity_representation = velocity_representation
            yield self
        finally:
            self.velocity_representation = original

    # ------------------------------------------------------------------ replace
    def replace(
        self,
        model,
        joint_positions=None,
        joint_velocities=None,
        base_quaternion=None,
        base_linear_velocity=None,
        base_angular_velocity=None,
        base_position=None,
        *,
        contact_state: dict | None = None,
        validate: bool = False,
        _inertial_base_velocity: tuple | None = None,
    ) -> "JaxSimModelData":
        """``JaxSimModelData.replace`` (``api/data.py:406-523``): new object with the given
        leaves replaced, quaternion normalised and all caches recomputed."""
        ref = self._base_quaternion
        unbatched = ref.dim() == 1
        dt, dev = ref.dtype, ref.device

        def pick(new, old):
            t = old if new is None else torch.as_tensor(new, dtype=dt, device=dev)
            return (t.unsqueeze(0) if t.dim() == 1 else t).contiguous()

        s = pick(joint_positions, self._joint_positions)
        sd = pick(joint_velocities, self._joint_velocities)
        q = pick(base_quaternion, self._base_quaternion)
        p = pick(base_position, self._base_position)
        if _inertial_base_velocity is not None:
            vl = pick(_inertial_base_velocity[0], None)
            om = pick(_inertial_base_velocity[1], None)
        elif base_linear_velocity is None and base_angular_velocity is None:
            vl = pick(None, self._base_linear_velocity)
            om = pick(None, self._base_angular_velocity)
        else:
            bv = self.base_velocity if (base_linear_velocity is None or base_angular_velocity is None) else None
            bl = pick(base_linear_velocity, None if bv is None else bv[..., 0:3])
            ba = pick(base_angular_velocity, None if bv is None else bv[..., 3:6])
            norm = torch.linalg.norm(q, dim=-1, keepdim=True)
            qn = q / torch.where(norm == 0, torch.ones_like(norm), norm)
            W_H_B = _transform_from_quat_pos(qn, p)
            v6 = other_representation_to_inertial(torch.cat([bl, ba], -1), self.velocity_representation, W_H_B, is_force=False)
            vl, om = v6[:, 0:3].contiguous(), v6[:, 3:6].contiguous()
        cs = dict(self.contact_state if contact_state is None else contact_state)
        out = _with_caches(model, self.velocity_representation, s, sd, q, vl, om, p, cs, normalise_q=True)
        return out._squeezed() if unbatched else out

    def reset_base_position(self, model, base_position):
        return self.replace(model=model, base_position=base_position)

    def reset_base_quaternion(self, model, base_quaternion):
        return self.replace(model=model, base_quaternion=base_quaternion)

    def reset_joint_positions(self, model, joint_positions):
        return self.replace(model=model, joint_positions=joint_positions)

    def reset_joint_velocities(self, model, joint_velocities):
        return self.replace(model=model, joint_velocities=joint_velocities)

    def valid(self, model) -> bool:
        """``api/data.py:525-549`` (per-environment shapes)."""
        n = model.dofs()
        return (
            self._joint_positions.shape[-1] == n
            and self._joint_velocities.shape[-1] == n
            and self._base_position.shape[-1] == 3
            and self._base_quaternion.shape[-1] == 4
            and self._base_linear_velocity.shape[-1] == 3
            and self._base_angular_velocity.shape[-1] == 3
        )

    def copy(self) -> "JaxSimModelData":
        return dataclasses.replace(self, contact_state=dict(self.contact_state))

    def _squeezed(self) -> "JaxSimModelData":
        sq = lambda t: None if t is None else t.squeeze(0)  # noqa: E731
        return JaxSimModelData(
            velocity_representation=self.velocity_representation,
            _joint_positions=sq(self._joint_positions), _joint_velocities=sq(self._joint_velocities),
            _base_quaternion=sq(self._base_quaternion), _base_linear_velocity=sq(self._base_linear_velocity),
            _base_angular_velocity=sq(self._base_angular_velocity), _base_position=sq(self._base_position),
            _base_transform=sq(self._base_transform), _joint_transforms=sq(self._joint_transforms),
            _link_transforms=sq(self._link_transforms), _link_velocities=sq(self._link_velocities),
            contact_state={k: sq(v) for k, v in self.contact_state.items()},
        )

    # state leaves as one dict (checkpointing = torch.save of this)
    def state_dict(self) -> dict:
        d = {
            "joint_positions": self._joint_positions, "joint_velocities": self._joint_velocities,
            "base_quaternion": self._base_quaternion, "base_linear_velocity": self._base_linear_velocity,
            "base_angular_velocity": self._base_angular_velocity, "base_position": self._base_position,
        }
        d.update({f"contact_state.{k}": v for k, v in self.contact_state.items()})
        return d


def _map_leaves(data: "JaxSimModelData", f) -> "JaxSimModelData":
    """A data object with ``f`` applied to every tensor leaf (None leaves stay None)."""
    g = lambda t: None if t is None else f(t)  # noqa: E731
    return JaxSimModelData(
        velocity_representation=data.velocity_representation,
        _joint_positions=g(data._joint_positions), _joint_velocities=g(data._joint_velocities),
        _base_quaternion=g(data._base_quaternion), _base_linear_velocity=g(data._base_linear_velocity),
        _base_angular_velocity=g(data._base_angular_velocity), _base_position=g(data._base_position),
        _base_transform=g(data._base_transform), _joint_transforms=g(data._joint_transforms),
        _link_transforms=g(data._link_transforms), _link_velocities=g(data._link_velocities),
        contact_state={k: g(v) for k, v in data.contact_state.items()},
    )


def _transform_from_quat_pos(q: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """``Transform.from_quaternion_and_translation`` (``math/transform.py:14-56``)."""
    nsq = (q * q).sum(-1)
    k = 2.0 / nsq
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    H = torch.zeros(q.shape[:-1] + (4, 4), dtype=q.dtype, device=q.device)
    H[..., 0, 0] = 1 - (y * y + z * z) * k
    H[..., 0, 1] = (x * y - w * z) * k
    H[..., 0, 2] = (x * z + w * y) * k
    H[..., 1, 0] = (x * y + w * z) * k
    H[..., 1, 1] = 1 - (x * x + z * z) * k
    H[..., 1, 2] = (y * z - w * x) * k
    H[..., 2, 0] = (x * z - w * y) * k
    H[..., 2, 1] = (y * z + w * x) * k
    H[..., 2, 2] = 1 - (x * x + y * y) * k
    H[..., 0:3, 3] = p
    H[..., 3, 3] = 1.0
    return H


def _with_caches(model, repr_, s, sd, q, vl, om, p, cs, *, normalise_q: bool) -> JaxSimModelData:
    """Run ``b200sim_fk`` on batched leaves and assemble a batched data object."""
    from .model import _dtype_code, _ptr, _stream_ptr

    dev, dtype = q.device, q.dtype
    dm = model.device_model(dev)
    B, nL = q.shape[0], model.number_of_links()
    new = lambda *shape: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
    q_o = new(B, 4)
    W_H_B, iXl, W_H_L, W_v = new(B, 4, 4), new(B, nL, 6, 6), new(B, nL, 4, 4), new(B, nL, 6)
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_fk(
            dm.handle, _dtype_code(dtype), B, _ptr(s), _ptr(sd), _ptr(q), _ptr(vl), _ptr(om), _ptr(p),
            _ptr(q_o), _ptr(W_H_B), _ptr(iXl), _ptr(W_H_L), _ptr(W_v), _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_fk")
    return JaxSimModelData(
        velocity_representation=repr_, _joint_positions=s, _joint_velocities=sd,
        _base_quaternion=q_o if normalise_q else q, _base_linear_velocity=vl, _base_angular_velocity=om,
        _base_position=p, _base_transform=W_H_B, _joint_transforms=iXl, _link_transforms=W_H_L,
        _link_velocities=W_v, contact_state=cs,
    )


# =============================================================================
# random_model_data (api/data.py:552-682) + random_joint_positions (api/joint.py:184-277)
# =============================================================================


def joint_position_sampling_bounds(model) -> tuple[np.ndarray, np.ndarray]:
    """The sampling window of ``random_joint_positions`` (``api/joint.py:184-277``)."""
    jp = model.kin_dyn_parameters.joint_parameters
    jt = np.array(model.kin_dyn_parameters.joint_model.joint_types[1:])
    s_min, s_max = jp.position_limits_min.copy(), jp.position_limits_max.copy()
    pi = math.pi
    with np.errstate(over="ignore"):
        full = np.logical_and(jt == 1, s_max - s_min >= 2 * pi)
    s_min = np.where(np.logical_and(full, np.logical_and(s_min <= -pi, s_max >= pi)), -pi, s_min)
    s_max = np.where(np.logical_and(full, np.logical_and(s_min <= -pi, s_max >= pi)), pi, s_max)
    s_min = np.where(np.logical_and(full, s_max < pi), s_max - 2 * pi, s_min)
    s_max = np.where(np.logical_and(full, s_min > -pi), s_min + 2 * pi, s_max)
    s_min = np.where(np.abs(s_min) < 1e30, s_min, -1.0)
    s_max = np.where(np.abs(s_max) < 1e30, s_max, 1.0)
    return s_min, s_max


def random_model_data(
    model,
    *,
    batch_size: int | None = None,
    seed: int = 0,
    velocity_representation: VelRepr | None = None,
    base_pos_bounds=((-1, -1, 0.5), 1.0),
    base_rpy_bounds=(-math.pi, math.pi),
    joint_pos_bounds=None,
    base_vel_lin_bounds=(-1.0, 1.0),
    base_vel_ang_bounds=(-1.0, 1.0),
    joint_vel_bounds=(-1.0, 1.0),
    dtype: torch.dtype = torch.float64,
    device: torch.device | str = "cuda",
) -> JaxSimModelData:
    """``random_model_data`` (``api/data.py:552-682``) with the same distributions, drawn
    from a seeded torch generator on the device (the JAX threefry stream of ``key`` cannot
    be reproduced without JAX; benchmarks only need the distribution)."""
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(int(seed))
    B = 1 if batch_size is None else batch_size
    n = model.dofs()

    def uni(shape, lo, hi):
        lo = torch.as_tensor(lo, dtype=torch.float64, device=device)
        hi = torch.as_tensor(hi, dtype=torch.float64, device=device)
        return lo + (hi - lo) * torch.rand(shape, generator=gen, dtype=torch.float64, device=device)

    p = uni((B, 3), base_pos_bounds[0], base_pos_bounds[1])
    rpy = uni((B, 3), base_rpy_bounds[0], base_rpy_bounds[1])
    # scipy Rotation.from_euler("XYZ") = intrinsic: q = qx * qy * qz
    h = rpy / 2
    z, o = torch.zeros(B, dtype=torch.float64, device=device), None
    qx = torch.stack([h[:, 0].cos(), h[:, 0].sin(), z, z], -1)
    qy = torch.stack([h[:, 1].cos(), z, h[:, 1].sin(), z], -1)
    qz = torch.stack([h[:, 2].cos(), z, z, h[:, 2].sin()], -1)
    q = _qmul(_qmul(qx, qy), qz)
    s = sd = vl = om = None
    if n > 0:
        if joint_pos_bounds is None:
            lo, hi = joint_position_sampling_bounds(model)
        else:
            lo, hi = joint_pos_bounds
        s = uni((B, n), lo, hi)
        sd = uni((B, n), joint_vel_bounds[0], joint_vel_bounds[1])
    if model.floating_base():
        vl = uni((B, 3), base_vel_lin_bounds[0], base_vel_lin_bounds[1])
        om = uni((B, 3), base_vel_ang_bounds[0], base_vel_ang_bounds[1])
    kw = {} if velocity_representation is None else {"velocity_representation": velocity_representation}
    out = JaxSimModelData.build(
        model, base_position=p, base_quaternion=q, joint_positions=s, joint_velocities=sd,
        base_linear_velocity=vl, base_angular_velocity=om, batch_size=B, dtype=dtype, device=device, **kw,
    )
    return out._squeezed() if batch_size is None else out


def _qmul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        -1,
    )
