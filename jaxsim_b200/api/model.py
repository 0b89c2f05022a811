"""``JaxSimModel`` and the drop-in ``step`` (host side).

Mirrors the public surface of ``src/jaxsim/api/model.py`` for the hot path only:

* ``JaxSimModel`` (``:46-90``) with the same field names and builders (``:128-330``);
* ``step(model, data, *, link_forces=None, joint_force_references=None)`` (``:2601-2681``);
* ``forward_dynamics_aba`` (``:1269-1406``).

Everything numerical happens in ``libb200sim.so`` (``csrc/``) through the C ABI of
``include/b200sim.h``; torch tensors are only the device-buffer type.  There is no CPU path:
calling these functions without the CUDA library or with CPU tensors raises.
"""

from __future__ import annotations

import contextlib
import copy
import ctypes as C
import dataclasses
import enum
import math
import pathlib
import weakref

import numpy as np
import torch

from jaxsim_b200 import _lib
from jaxsim_b200.parsers.urdf import build_kin_dyn_parameters
from jaxsim_b200.rbda.actuation import ActuationParams
from jaxsim_b200.rbda.contacts import (RelaxedRigidContacts, RelaxedRigidContactsParams, RigidContacts,
                                        RigidContactsParams, SoftContacts, SoftContactsParams)
from jaxsim_b200.terrain import FlatTerrain

from . import data as _data
from .common import VelRepr, inertial_to_other_representation, other_representation_to_inertial
from .kin_dyn_parameters import KinDynParameters

STANDARD_GRAVITY = 9.81  # src/jaxsim/math/__init__.py:14
_NULLCTX = contextlib.nullcontext()


class IntegratorType(enum.IntEnum):
    """``src/jaxsim/api/model.py:33-43``. Only SemiImplicitEuler is on the hot path."""

    SemiImplicitEuler = enum.auto()
    RungeKutta4 = enum.auto()
    RungeKutta4Fast = enum.auto()


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class _DeviceModel:
    """Owns one ``B200SimModel*`` (one per CUDA device)."""

    def __init__(self, model: "JaxSimModel", device_index: int):
        lib = _lib.load()
        kd = model.kin_dyn_parameters
        nL, n = kd.number_of_links(), kd.number_of_joints()
        cp = kd.contact_parameters
        nc = len(cp.body)
        keep = {}

        def dp(name, arr):
            keep[name] = _np(arr, np.float64)
            return keep[name].ctypes.data_as(_lib.c_dp)

        def ip(name, arr):
            keep[name] = _np(arr, np.int32)
            return keep[name].ctypes.data_as(_lib.c_ip)

        axis = np.zeros((nL, 3))
        if n > 0:
            axis[1:] = kd.joint_model.joint_axis
        reg = 1e-6
        rx = RelaxedRigidContactsParams()
        if (model.contact_model is not None and model.contact_params is not None
                and not isinstance(model.contact_params, model.contact_model._parameters_class)):
            raise TypeError(
                f"contact_params is a {type(model.contact_params).__name__} but {type(model.contact_model).__name__} "
                f"takes {model.contact_model._parameters_class.__name__}"
            )
        prm = model.contact_params if isinstance(model.contact_params, SoftContactsParams) else SoftContactsParams()
        if model.contact_model is None or nc == 0:
            cm = 0
        elif isinstance(model.contact_model, SoftContacts):
            cm = 1
        elif isinstance(model.contact_model, RigidContacts):
            cm = 2
            rp = model.contact_params if isinstance(model.contact_params, RigidContactsParams) else RigidContactsParams()
            # the descriptor's soft_K / soft_D / soft_mu slots carry the rigid K / D / mu (include/b200sim.h)
            prm = SoftContactsParams(K=rp.K, D=rp.D, mu=rp.mu)
            reg = float(model.contact_model.regularization_delassus)
        elif isinstance(model.contact_model, RelaxedRigidContacts):
            cm = 3
            rx = model.contact_params if isinstance(model.contact_params, RelaxedRigidContactsParams) else RelaxedRigidContactsParams()
            prm = SoftContactsParams(K=rx.K, D=rx.D, mu=rx.mu)  # the friction coefficient travels in soft_mu
        else:
            raise NotImplementedError(f"contact model {type(model.contact_model).__name__} is not implemented")
        jp = kd.joint_parameters
        d = _lib.B200SimModelDesc(
            abi_version=_lib.ABI_VERSION, n_links=nL, n_dofs=n, n_points=nc,
            floating_base=int(model.floating_base()), contact_model=cm,
            enable_friction=int(model.actuation_params.enable_friction), reserved0=0,
            parent=ip("parent", kd.parent_array), joint_type=ip("jt", kd.joint_model.joint_types),
            lam_H_pre=dp("lam", kd.joint_model.lam_H_pre), suc_H_i=dp("suc", kd.joint_model.suc_H_i),
            joint_axis=dp("axis", axis),
            link_mass=dp("mass", kd.link_parameters.mass), link_com=dp("com", kd.link_parameters.center_of_mass),
            link_inertia=dp("inertia", kd.link_parameters.inertia_elements),
            friction_static=dp("kc", jp.friction_static), friction_viscous=dp("kv", jp.friction_viscous),
            position_limits_min=dp("smin", jp.position_limits_min), position_limits_max=dp("smax", jp.position_limits_max),
            position_limit_spring=dp("ks", jp.position_limit_spring), position_limit_damper=dp("kd", jp.position_limit_damper),
            point_body=ip("pb", np.array(cp.body, dtype=np.int32)),
            point_position=dp("pp", np.asarray(cp.point, dtype=float).reshape(-1, 3) if nc else np.zeros((0, 3))),
            point_enabled=ip("pe", np.array(cp.enabled, dtype=np.int32)),
            time_step=float(model.time_step), gravity=float(model.gravity),
            terrain_height=float(model.terrain.height()),
            soft_K=prm.K, soft_D=prm.D, soft_mu=prm.mu, soft_p=prm.p, soft_q=prm.q,
            torque_max=model.actuation_params.torque_max, omega_th=model.actuation_params.omega_th,
            omega_max=model.actuation_params.omega_max, rigid_regularization=reg,
            relaxed_time_constant=rx.time_constant, relaxed_damping_coefficient=rx.damping_coefficient,
            relaxed_d_min=rx.d_min, relaxed_d_max=rx.d_max, relaxed_width=rx.width, relaxed_midpoint=rx.midpoint,
            relaxed_power=rx.power,
        )
        handle = C.c_void_p()
        _lib.check(lib.b200sim_model_create(C.byref(d), int(device_index), C.byref(handle)), "b200sim_model_create")
        self.handle = handle
        self.device_index = device_index
        self._finalizer = weakref.finalize(self, lib.b200sim_model_destroy, handle)


@dataclasses.dataclass(eq=False)
class JaxSimModel:
    """``src/jaxsim/api/model.py:46-90`` (host container; the device blob is created lazily
    per CUDA device on first use)."""

    model_name: str
    time_step: float = 0.001
    terrain: FlatTerrain = dataclasses.field(default_factory=FlatTerrain.build)
    gravity: float = -STANDARD_GRAVITY
    contact_model: object | None = None
    contact_params: object | None = None
    actuation_params: ActuationParams | None = None
    kin_dyn_parameters: KinDynParameters | None = None
    integrator: IntegratorType = IntegratorType.SemiImplicitEuler
    built_from: object | None = None
    _floating_base: bool = True
    _devices: dict = dataclasses.field(default_factory=dict, repr=False, init=False, compare=False)
    _tuning: tuple = (0, 0)
    _options: int | None = None

    # Fields baked into the device blob (b200sim_model_create).  The reference's models are immutable pytrees
    # edited through `model.editable()` / `replace` (utils/jaxsim_dataclass.py:22-27,313-334); here the host
    # container is a plain dataclass, so assigning any of these drops the cached device models and the next step
    # uploads a blob with the new values.  `dataclasses.replace` builds a new instance with its own (empty) cache.
    _PHYSICS_FIELDS = frozenset({
        "time_step", "terrain", "gravity", "contact_model", "contact_params", "actuation_params",
        "kin_dyn_parameters", "_floating_base",
    })

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in JaxSimModel._PHYSICS_FIELDS:
            devs = self.__dict__.get("_devices")
            if devs:
                devs.clear()

    def replace(self, **changes) -> "JaxSimModel":
        """A copy with some fields replaced (``replace`` of the reference's dataclasses); the copy creates its
        own device models lazily, with the new values."""
        return dataclasses.replace(self, **changes)

    @contextlib.contextmanager
    def editable(self, validate: bool = True):
        """``with model.editable(validate=False) as model:`` of the reference (``utils/jaxsim_dataclass.py:22-27``):
        yields a copy whose fields -- and whose ``kin_dyn_parameters`` container -- can be assigned without touching the
        original (e.g. ``model.kin_dyn_parameters.constraints = ...``, ``tests/test_simulations.py:438-440``)."""
        yield dataclasses.replace(self, kin_dyn_parameters=copy.copy(self.kin_dyn_parameters))

    # ------------------------------------------------------------------ builders
    @classmethod
    def build_from_model_description(
        cls,
        model_description: str | pathlib.Path,
        *,
        model_name: str | None = None,
        time_step: float | None = None,
        terrain: FlatTerrain | None = None,
        contact_model=None,
        contact_params=None,
        actuation_params: ActuationParams | None = None,
        integrator: IntegratorType | None = None,
        is_urdf: bool | None = None,
        considered_joints=None,
        gravity: float = STANDARD_GRAVITY,
    ) -> "JaxSimModel":
        """``JaxSimModel.build_from_model_description`` (``api/model.py:128-223``): URDF path
        or XML string.  ``gravity`` is passed positive and stored negated (``:206``)."""
        if considered_joints is not None:
            raise NotImplementedError("model reduction (api/model.py:807-878) is out of scope")
        name, kd, floating = build_kin_dyn_parameters(model_description)
        return cls.build(
            kd, floating_base=floating, model_name=model_name or name, time_step=time_step, terrain=terrain,
            contact_model=contact_model, contact_params=contact_params, actuation_params=actuation_params,
            integrator=integrator, gravity=-gravity, built_from=model_description,
        )

    @classmethod
    def build(
        cls,
        kin_dyn_parameters: KinDynParameters,
        *,
        floating_base: bool,
        model_name: str = "model",
        time_step: float | None = None,
        terrain: FlatTerrain | None = None,
        contact_model=None,
        contact_params=None,
        actuation_params: ActuationParams | None = None,
        integrator: IntegratorType | None = None,
        gravity: float = -STANDARD_GRAVITY,
        built_from=None,
    ) -> "JaxSimModel":
        """``JaxSimModel.build`` (``api/model.py:224-330``): defaults are SoftContacts with
        default parameters, default ActuationParams, SemiImplicitEuler, flat terrain at 0."""
        contact_model = contact_model if contact_model is not None else SoftContacts.build()
        if contact_params is None:
            contact_params = contact_model._parameters_class()
        elif not isinstance(contact_params, contact_model._parameters_class):
            raise TypeError(
                f"contact_params is a {type(contact_params).__name__} but {type(contact_model).__name__} takes "
                f"{contact_model._parameters_class.__name__}"
            )
        integrator = integrator if integrator is not None else IntegratorType.SemiImplicitEuler
        return cls(
            model_name=model_name,
            time_step=float(time_step) if time_step is not None else 0.001,
            terrain=terrain if terrain is not None else FlatTerrain.build(),
            gravity=float(gravity),
            contact_model=contact_model,
            contact_params=contact_params,
            actuation_params=actuation_params if actuation_params is not None else ActuationParams(),
            kin_dyn_parameters=kin_dyn_parameters,
            integrator=integrator,
            built_from=built_from,
            _floating_base=bool(floating_base),
        )

    # ------------------------------------------------------------------ properties
    def name(self) -> str:
        return self.model_name

    def number_of_links(self) -> int:
        return self.kin_dyn_parameters.number_of_links()

    def number_of_joints(self) -> int:
        return self.kin_dyn_parameters.number_of_joints()

    def dofs(self) -> int:
        return self.kin_dyn_parameters.number_of_joints()

    def floating_base(self) -> bool:
        return self._floating_base

    def base_link(self) -> str:
        return self.kin_dyn_parameters.link_names[0]

    def link_names(self) -> tuple[str, ...]:
        return self.kin_dyn_parameters.link_names

    def joint_names(self) -> tuple[str, ...]:
        return self.kin_dyn_parameters.joint_model.joint_names[1:]

    def number_of_collidable_points(self) -> int:
        return len(self.kin_dyn_parameters.contact_parameters.body)

    # ------------------------------------------------------------------ device side
    def device_model(self, device: torch.device) -> _DeviceModel:
        if device.type != "cuda":
            raise RuntimeError(
                "jaxsim_b200 runs on CUDA devices only (no CPU fallback): move the data to a B200 first"
            )
        idx = device.index if device.index is not None else torch.cuda.current_device()
        dm = self._devices.get(idx)
        if dm is None:
            dm = _DeviceModel(self, idx)
            if self._tuning != (0, 0):
                _lib.check(_lib.load().b200sim_model_set_tuning(dm.handle, *self._tuning), "set_tuning")
            if self._options is not None:
                _lib.check(_lib.load().b200sim_model_set_options(dm.handle, self._options), "set_options")
            self._devices[idx] = dm
        return dm

    def set_tuning(self, lanes_per_env: int = 0, envs_per_block: int = 0) -> None:
        """Performance knobs (never change results): see ``b200sim_model_set_tuning``."""
        self._tuning = (int(lanes_per_env), int(envs_per_block))
        for dm in self._devices.values():
            _lib.check(_lib.load().b200sim_model_set_tuning(dm.handle, *self._tuning), "set_tuning")

    def set_options(self, tma_store: bool = True, rigid_qp_f32: bool = False, generic_kernel: bool = False,
                    bulk_in: bool = False, pdl: bool = True, step_v1: bool = False, no_bulk_in: bool = False,
                    rigid_mono: bool = False) -> None:
        """Implementation switches: ``b200sim_model_set_options``.  ``rigid_qp_f32`` makes
        float32 rigid-contact steps solve the contact QP / impact system in float32 too
        (default float64; forces then only good to ~1e-3 relative, smaller workspace).
        ``generic_kernel`` disables the step-kernel instance specialised for floating-base
        soft-contact models (same results; A/B timing switch).  ``step_v1`` selects the
        first-generation specialised kernel instead of the second-generation one
        (``b200sim_step2.cuh``); ``bulk_in`` / ``tma_store`` only concern that older kernel,
        ``no_bulk_in`` makes the new one read the input caches with per-link ``cp.async``."""
        self._options = ((_lib.OPT_TMA_STORE if tma_store else 0) | (_lib.OPT_RIGID_QP_F32 if rigid_qp_f32 else 0)
                         | (_lib.OPT_GENERIC_KERNEL if generic_kernel else 0) | (_lib.OPT_BULK_IN if bulk_in else 0) | (0 if pdl else _lib.OPT_NO_PDL)
                         | (_lib.OPT_STEP_V1 if step_v1 else 0) | (_lib.OPT_NO_BULK_IN if no_bulk_in else 0)
                         | (_lib.OPT_RIGID_MONO if rigid_mono else 0))
        for dm in self._devices.values():
            _lib.check(_lib.load().b200sim_model_set_options(dm.handle, self._options), "set_options")

    def launch_geometry(self, batch: int, dtype: torch.dtype, device: torch.device) -> dict:
        dm = self.device_model(device)
        out = [C.c_int32() for _ in range(4)]
        _lib.check(
            _lib.load().b200sim_model_query(dm.handle, _dtype_code(dtype), int(batch), *[C.byref(o) for o in out]),
            "b200sim_model_query",
        )
        return dict(lanes_per_env=out[0].value, envs_per_block=out[1].value, grid=out[2].value, smem_bytes=out[3].value)


def _dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return 0
    if dtype == torch.float64:
        return 1
    raise TypeError(f"unsupported dtype {dtype}: the step computes in float32 or float64")


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _batched(t: torch.Tensor, unbatched_ndim: int) -> torch.Tensor:
    return t if t.dim() > unbatched_ndim else t.unsqueeze(0)


# =============================================================================
# step
# =============================================================================


def _alloc_outputs(model, B, dtype, dev, update_caches, soft):
    """One allocation for all the outputs of a step, carved into the leaves."""
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()
    # every block is a multiple of 4 elements so that each leaf stays 16-byte aligned
    r4 = lambda k: (k + 3) & ~3  # noqa: E731
    sizes = [("s", (B, n)), ("sd", (B, n)), ("q", (B, 4)), ("vl", (B, 3)), ("om", (B, 3)), ("p", (B, 3))]
    if soft:
        sizes.append(("m", (B, nc, 3)))
    if update_caches:
        sizes += [("W_H_B", (B, 4, 4)), ("iXl", (B, nL, 6, 6)), ("W_H_L", (B, nL, 4, 4)), ("W_v", (B, nL, 6))]
    total = sum(r4(math.prod(shape)) for _, shape in sizes)
    flat = torch.empty(total, dtype=dtype, device=dev)
    out, o = {}, 0
    for name, shape in sizes:
        k = math.prod(shape)
        out[name] = flat[o : o + k].view(shape)
        o += r4(k)
    return out


def _links_follow_aba_chain(model) -> bool:
    """True where the FK link poses coincide with the ABA chain poses (floating base, suc_H_i[0] = I: every URDF
    floating-base model): the kernels can then re-express Body / Mixed link forces themselves."""
    H0 = np.asarray(model.kin_dyn_parameters.joint_model.suc_H_i)[0]
    return bool(model.floating_base()) and bool(np.array_equal(H0, np.eye(4)))


def _link_forces_with_constraints(model, data, n_steps, link_forces, joint_force_references):
    """Inertial-fixed link forces of ONE step of a model with weld constraints: the caller's forces plus the constraint
    wrenches (``api/ode.py:60-107``).  The wrenches depend on the state, so a fused multi-step launch cannot carry them."""
    from jaxsim_b200.rbda import kinematic_constraints as _kc

    from . import contact as _contact
    from . import ode as _ode

    if n_steps != 1:
        raise NotImplementedError("kinematic constraints are re-solved at every step: use step() (step_n loops for you)")
    nc = model.number_of_collidable_points()
    soft = isinstance(model.contact_model, SoftContacts)
    if nc > 0 and not soft:
        raise NotImplementedError("kinematic constraints with RigidContacts / RelaxedRigidContacts collidable points: the "
                                  "contact forces of those models are solved inside the step kernel")
    dtype, dev = data._joint_positions.dtype, data._joint_positions.device
    nL = model.number_of_links()
    # The solve is ill-conditioned by construction: the regulariser (1e-3) sits next to Delassus eigenvalues of O(1/mass),
    # and a planar loop leaves half of the 6 weld directions to the regulariser alone -- directions that still act on
    # the base through the body/mixed mix of the reference's Jacobian.  Rounding of the right-hand side is amplified
    # by ~1e3-1e4, so everything that feeds it (kinematics, contact forces, ABA, mass matrix) is evaluated in float64
    # whatever the precision of the state; only the resulting link forces are rounded to the state's dtype.
    d = _kc.float64_data(model, data)
    lead = d._base_quaternion.shape[:-1]
    if link_forces is None:
        W_f = torch.zeros(lead + (nL, 6), dtype=torch.float64, device=dev)
    else:
        W_f = torch.as_tensor(link_forces, dtype=dtype, device=dev)
        if W_f.shape[-2:] != (nL, 6) or W_f.dim() != data._base_quaternion.dim() + 1:
            raise ValueError(W_f.shape, (nL, 6))
        W_f = other_representation_to_inertial(W_f.to(torch.float64).reshape(lead + (nL, 6)), data.velocity_representation,
                                               d.link_transforms, is_force=True)
    tau = None if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=torch.float64, device=dev)
    tau_total = _ode.compute_resultant_torques(model, d, joint_force_references=None if tau is None else tau.reshape(lead + (-1,)))
    W_f_terrain = 0
    if nc > 0:
        W_f_terrain, _ = _contact.soft_link_contact_forces(model, d)
    W_f = W_f + _kc.constraint_link_forces(model, d, joint_torques=tau_total, link_forces_inertial=W_f + W_f_terrain)
    return W_f.to(dtype).reshape(data._base_quaternion.shape[:-1] + (nL, 6))


def _step_impl(model, data, n_steps, link_forces, joint_force_references, update_caches, out, use_input_caches=True,
               status_flags=None):
    s = data._joint_positions
    unbatched = s.dim() == 1
    dev = s.device
    dm = model.device_model(dev)
    dtype = s.dtype
    code = _dtype_code(dtype)
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()

    s = _batched(data._joint_positions, 1).contiguous()
    sd = _batched(data._joint_velocities, 1).contiguous()
    q = _batched(data._base_quaternion, 1).contiguous()
    vl = _batched(data._base_linear_velocity, 1).contiguous()
    om = _batched(data._base_angular_velocity, 1).contiguous()
    p = _batched(data._base_position, 1).contiguous()
    B = q.shape[0]
    if s.shape != (B, n) or sd.shape != (B, n):
        raise ValueError((s.shape, sd.shape), (B, n))  # rbda/utils.py:102-133
    for t, w in ((q, 4), (vl, 3), (om, 3), (p, 3)):
        if t.shape != (B, w):
            raise ValueError(t.shape, (B, w))

    m = data.contact_state.get("tangential_deformation") if data.contact_state else None
    if m is not None:
        m = _batched(m, 2).contiguous()
        if m.shape != (B, nc, 3):
            raise ValueError(m.shape, (B, nc, 3))

    tau, tau_stride = None, 0
    if joint_force_references is not None:
        tau = torch.as_tensor(joint_force_references, dtype=dtype, device=dev)
        if tau.dim() == 3:  # (T, B, n): one row of references per step
            if tau.shape != (n_steps, B, n):
                raise ValueError(tau.shape, (n_steps, B, n))
            tau_stride = B * n
        else:
            tau = _batched(tau, 1)
            if tau.shape != (B, n):
                raise ValueError(tau.shape, (B, n))
        tau = tau.contiguous()

    forces_are_inertial = data.velocity_representation == VelRepr.Inertial
    cmap = model.kin_dyn_parameters.constraints
    if cmap is not None and len(cmap) > 0:
        # api/ode.py:75-107: the weld-constraint wrenches, solved against external + contact forces, join the link forces
        link_forces = _link_forces_with_constraints(model, data, n_steps, link_forces, joint_force_references)
        forces_are_inertial = True

    fext, fext_stride, fext_repr = None, 0, _lib.REPR_INERTIAL
    if link_forces is not None:
        O_f = torch.as_tensor(link_forces, dtype=dtype, device=dev)
        per_step = O_f.dim() == 4
        O_f = O_f if per_step else _batched(O_f, 2)
        if O_f.shape[-3:] != (B, nL, 6) or (per_step and O_f.shape[0] != n_steps):
            raise ValueError(O_f.shape, (B, nL, 6))
        if not forces_are_inertial:
            # api/model.py:2641-2646: expressed in data.velocity_representation, re-expressed with the link
            # transforms of every step.  The kernels do that themselves (b200sim_step_n_ex) wherever the link poses
            # are the ABA chain poses; otherwise (fixed base with an offset mount, SDF-posed base) a torch shim converts
            # with the transforms of the input state, which is only right for a single step.
            if _links_follow_aba_chain(model):
                fext_repr = _lib.REPR_BODY if data.velocity_representation == VelRepr.Body else _lib.REPR_MIXED
            else:
                if n_steps > 1:
                    raise NotImplementedError(
                        "Body / Mixed link_forces of a multi-step launch on a model whose base link pose is offset "
                        "from the chain root: call step() per step, or give the forces in VelRepr.Inertial")
                O_f = other_representation_to_inertial(
                    O_f, data.velocity_representation, _batched(data.link_transforms, 3), is_force=True
                )
        fext = O_f.contiguous()
        fext_stride = B * nL * 6 if per_step else 0

    soft = isinstance(model.contact_model, SoftContacts)
    if out is None:
        o = _alloc_outputs(model, B, dtype, dev, update_caches, soft)
    else:
        # reuse the buffers of an existing (batched) data object: no allocation at all
        o = {"s": out._joint_positions, "sd": out._joint_velocities, "q": out._base_quaternion,
             "vl": out._base_linear_velocity, "om": out._base_angular_velocity, "p": out._base_position}
        if soft:
            o["m"] = out.contact_state["tangential_deformation"]
        if update_caches:
            # caches left as None in `out` are not materialised
            for key, buf in (("W_H_B", out._base_transform), ("iXl", out._joint_transforms),
                             ("W_H_L", out._link_transforms), ("W_v", out._link_velocities)):
                if buf is not None:
                    o[key] = buf
        if o["q"].shape != (B, 4) or o["q"].dtype != dtype or o["q"].device != dev:
            raise ValueError("`out` does not match the batch/dtype/device of `data`")
        if any(v is None or not v.is_contiguous() for v in o.values()):
            raise ValueError("`out` must hold contiguous buffers for every requested leaf")

    # the cached kinematics of the input state spare the kernel a sincos + FK pass
    Hin = Vin = None
    if isinstance(model.contact_model, (RigidContacts, RelaxedRigidContacts)) and nc > 0:
        # RigidContacts READS the cached link velocities on purpose: after an impact they are the
        # pre-impact ones (rbda/contacts/rigid.py:429-434 never refreshes them) and the reference's
        # penetration rate / Jacobian-derivative term use them (api/contact.py:39-43,470-477)
        if n_steps != 1 and not update_caches:
            raise NotImplementedError("step_n with RigidContacts / RelaxedRigidContacts needs update_caches=True "
                                      "(every step reads the cached link transforms / velocities of the previous one)")
        if use_input_caches and data._link_transforms is not None and data._link_velocities is not None:
            Hin = _batched(data._link_transforms, 3).to(dtype).contiguous()
            Vin = _batched(data._link_velocities, 2).to(dtype).contiguous()
            if Hin.data_ptr() % 16 or Vin.data_ptr() % 16:
                Hin, Vin = Hin.clone(), Vin.clone()
    elif use_input_caches and not unbatched and data._link_transforms is not None and data._link_velocities is not None:
        Hin, Vin = data._link_transforms, data._link_velocities
        if not (Hin.is_contiguous() and Vin.is_contiguous() and Hin.dtype == dtype and Vin.dtype == dtype
                and Hin.shape == (B, nL, 4, 4) and Vin.shape == (B, nL, 6) and Hin.data_ptr() % 16 == 0 and Vin.data_ptr() % 16 == 0):
            Hin = Vin = None
    g = o.get
    if torch.cuda.current_device() != dm.device_index:
        ctx = torch.cuda.device(dev)
    else:
        ctx = _NULLCTX
    args = (
        dm.handle, code, B, int(n_steps),
        _ptr(s), _ptr(sd), _ptr(q), _ptr(vl), _ptr(om), _ptr(p), _ptr(m), _ptr(tau), tau_stride,
        _ptr(fext), fext_stride, _ptr(Hin), _ptr(Vin),
        _ptr(o["s"]), _ptr(o["sd"]), _ptr(o["q"]), _ptr(o["vl"]), _ptr(o["om"]), _ptr(o["p"]), _ptr(g("m")),
        _ptr(g("W_H_B")), _ptr(g("iXl")), _ptr(g("W_H_L")), _ptr(g("W_v")),
    )
    if status_flags is not None:
        if (status_flags.dtype != torch.int32 or status_flags.shape != (B,) or status_flags.device != dev
                or not status_flags.is_contiguous()):
            raise ValueError("status_flags must be a contiguous int32 tensor of shape (B,) on the device of `data`")
        if o["q"].data_ptr() == q.data_ptr():
            raise ValueError("status_flags describe the input quaternion too: not available for an in-place step")
    with ctx:
        if status_flags is None and fext_repr == _lib.REPR_INERTIAL:
            rc = _lib.load().b200sim_step_n(*args, _stream_ptr(dev))
        else:
            rc = _lib.load().b200sim_step_n_ex(*args, int(fext_repr), _ptr(status_flags), _stream_ptr(dev))
    _lib.check(rc, "b200sim_step_n")

    if out is not None:
        out.velocity_representation = data.velocity_representation
        if not update_caches:
            out._base_transform = out._joint_transforms = out._link_transforms = out._link_velocities = None
        return out
    sq = (lambda t: t.squeeze(0) if t is not None else None) if unbatched else (lambda t: t)
    contact_state = dict(data.contact_state) if data.contact_state else {}
    if soft:
        contact_state["tangential_deformation"] = sq(o["m"])
    return _data.JaxSimModelData(
        velocity_representation=data.velocity_representation,
        _joint_positions=sq(o["s"]), _joint_velocities=sq(o["sd"]), _base_quaternion=sq(o["q"]),
        _base_linear_velocity=sq(o["vl"]), _base_angular_velocity=sq(o["om"]), _base_position=sq(o["p"]),
        _base_transform=sq(g("W_H_B")), _joint_transforms=sq(g("iXl")), _link_transforms=sq(g("W_H_L")),
        _link_velocities=sq(g("W_v")), contact_state=contact_state,
    )


def step(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    *,
    link_forces: torch.Tensor | None = None,
    joint_force_references: torch.Tensor | None = None,
    update_caches: bool = True,
    out: "_data.JaxSimModelData | None" = None,
    use_input_caches: bool = True,
    status_flags: torch.Tensor | None = None,
) -> "_data.JaxSimModelData":
    """Perform a simulation step: drop-in for ``jaxsim.api.model.step``
    (``src/jaxsim/api/model.py:2601-2681``), batched over the leading axis of ``data``.

    Args:
        model: the model.
        data: the (batched) state.
        link_forces: 6D forces on the links, ``(B, nL, 6)`` (or ``(nL, 6)``), expressed in
            ``data.velocity_representation`` like in the reference (``:2617-2618``).
        joint_force_references: ``(B, n)`` joint force references.
        update_caches: if False, the cached transforms of the returned data are not
            materialised (rollout mode, "B_min" of SURVEY.md 8d); accessing them raises.
        out: optional data object whose buffers receive the result (like NumPy's ``out=``);
            avoids every allocation and makes the call CUDA-graph capturable.  ``out`` may
            be ``data`` itself (in-place step).
        use_input_caches: read the cached link transforms / velocities of ``data`` (like the
            reference's contact code does) instead of recomputing the kinematics of the
            input state; they are consistent with the state for every object this API
            produces.  Set False for data whose private leaves were edited by hand.
        status_flags: optional ``(B,)`` int32 device tensor that receives per-environment flags
            (``jaxsim_b200._lib.STATUS_*``: NaN / non-unit input quaternion, non-finite output, contact QP not
            converged) -- the conditions the reference raises as exceptions only under
            ``JAXSIM_ENABLE_EXCEPTIONS`` (``rbda/utils.py:136-146``) or drops (``rigid.py:359-362``).

    Returns:
        The new ``JaxSimModelData`` (same velocity representation; new tensors unless
        ``out`` is given: the input is not modified, like the reference's immutable pytrees).
    """
    if model.integrator in (IntegratorType.RungeKutta4, IntegratorType.RungeKutta4Fast):
        from .integrators import step_rk4

        cmap = model.kin_dyn_parameters.constraints
        if cmap is not None and len(cmap) > 0:
            raise NotImplementedError("kinematic constraints are wired into the SemiImplicitEuler step only")

        return step_rk4(model, data, link_forces=link_forces, joint_force_references=joint_force_references)
    return _step_impl(model, data, 1, link_forces, joint_force_references, update_caches, out, use_input_caches, status_flags)


def step_n(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    n_steps: int,
    *,
    link_forces: torch.Tensor | None = None,
    joint_force_references: torch.Tensor | None = None,
    update_caches: bool = True,
    out: "_data.JaxSimModelData | None" = None,
) -> "_data.JaxSimModelData":
    """``n_steps`` consecutive ``step`` calls fused in ONE kernel launch -- the user loop
    ``for _ in range(T): data = js.model.step(model, data, ...)`` (``README.md:80-84``) with the
    state kept on chip between steps.  ``joint_force_references`` is
    ``(B, n)`` (held constant) or ``(n_steps, B, n)``; ``link_forces`` likewise
    ``(B, nL, 6)`` or ``(n_steps, B, nL, 6)``.  Returns the data after the last step;
    results are identical to calling ``step`` ``n_steps`` times."""
    if n_steps < 1:
        raise ValueError(n_steps)
    if model.integrator != IntegratorType.SemiImplicitEuler:
        raise NotImplementedError("step_n fuses SemiImplicitEuler steps only")
    cmap = model.kin_dyn_parameters.constraints
    if cmap is not None and len(cmap) > 0 and n_steps > 1:
        # the constraint wrenches are a function of every intermediate state (api/ode.py:83-88): one launch per step
        tau, lf = joint_force_references, link_forces
        per_tau = tau is not None and torch.as_tensor(tau).dim() == 3
        per_lf = lf is not None and torch.as_tensor(lf).dim() == 4
        for k in range(int(n_steps)):
            last = k == n_steps - 1
            data = _step_impl(model, data, 1, lf[k] if per_lf else lf, tau[k] if per_tau else tau, True,
                              out if last else None)
        if not update_caches:
            data._base_transform = data._joint_transforms = data._link_transforms = data._link_velocities = None
        return data
    return _step_impl(model, data, int(n_steps), link_forces, joint_force_references, update_caches, out)


def rollout(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    n_steps: int,
    *,
    link_forces: torch.Tensor | None = None,
    joint_force_references: torch.Tensor | None = None,
    record_every: int = 1,
) -> tuple["_data.JaxSimModelData", "_data.JaxSimModelData"]:
    """``n_steps`` steps with the state recorded every ``record_every`` steps (the ``lax.scan``
    users write around ``step``, with trajectory subsampling): returns ``(final, trajectory)`` where the leaves of
    ``trajectory`` (state and contact state, no caches) carry a leading axis of ``n_steps // record_every`` samples,
    sample ``t`` being the state after ``(t + 1) * record_every`` steps.  Each sample is ONE fused ``step_n`` launch that
    writes straight into its slice of the trajectory buffers (no copies); ``joint_force_references`` / ``link_forces``
    are held constant or given per step (``(n_steps, B, ...)``) like in ``step_n``.  ``final`` is the last sample with
    its caches.  Batched data only."""
    k = int(record_every)
    if n_steps < 1 or k < 1 or n_steps % k:
        raise ValueError(f"n_steps = {n_steps} must be a positive multiple of record_every = {record_every}")
    q = data._base_quaternion
    if q.dim() != 2:
        raise ValueError("rollout needs batched data")
    T = int(n_steps) // k
    B, dtype, dev = q.shape[0], q.dtype, q.device
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()
    soft = isinstance(model.contact_model, SoftContacts)
    new = lambda *shape: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
    traj = {"s": new(T, B, n), "sd": new(T, B, n), "q": new(T, B, 4), "vl": new(T, B, 3), "om": new(T, B, 3), "p": new(T, B, 3)}
    if soft:
        traj["m"] = new(T, B, nc, 3)
    # the caches travel from sample to sample (the rigid contact models read them) in two alternating sets
    caches = [(new(B, 4, 4), new(B, nL, 6, 6), new(B, nL, 4, 4), new(B, nL, 6)) for _ in range(min(T, 2))]
    tau = None if joint_force_references is None else torch.as_tensor(joint_force_references, dtype=dtype, device=dev)
    lf = None if link_forces is None else torch.as_tensor(link_forces, dtype=dtype, device=dev)
    per_tau, per_lf = tau is not None and tau.dim() == 3, lf is not None and lf.dim() == 4
    if (per_tau and tau.shape[0] != n_steps) or (per_lf and lf.shape[0] != n_steps):
        raise ValueError("per-step references / forces need a leading axis of n_steps")
    fused = model.integrator == IntegratorType.SemiImplicitEuler
    cur = data
    for t in range(T):
        W_H_B, iXl, W_H_L, W_v = caches[t % 2]
        out = _data.JaxSimModelData(
            velocity_representation=data.velocity_representation,
            _joint_positions=traj["s"][t], _joint_velocities=traj["sd"][t], _base_quaternion=traj["q"][t],
            _base_linear_velocity=traj["vl"][t], _base_angular_velocity=traj["om"][t], _base_position=traj["p"][t],
            _base_transform=W_H_B, _joint_transforms=iXl, _link_transforms=W_H_L, _link_velocities=W_v,
            contact_state={"tangential_deformation": traj["m"][t]} if soft else dict(cur.contact_state or {}),
        )
        tau_t = tau[t * k:(t + 1) * k] if per_tau else tau
        lf_t = lf[t * k:(t + 1) * k] if per_lf else lf
        if fused:
            if k == 1:
                tau_t = tau_t[0] if per_tau else tau_t
                lf_t = lf_t[0] if per_lf else lf_t
            cur = step_n(model, cur, k, link_forces=lf_t, joint_force_references=tau_t, update_caches=True, out=out)
        else:  # RungeKutta4 / RungeKutta4Fast: one ABI call per step, the sample copied into its slice
            for j in range(k):
                cur = step(model, cur, link_forces=lf_t[j] if per_lf else lf_t, joint_force_references=tau_t[j] if per_tau else tau_t)
            for name, leaf in (("s", "_joint_positions"), ("sd", "_joint_velocities"), ("q", "_base_quaternion"),
                               ("vl", "_base_linear_velocity"), ("om", "_base_angular_velocity"), ("p", "_base_position")):
                traj[name][t].copy_(getattr(cur, leaf))
            if soft:
                traj["m"][t].copy_(cur.contact_state["tangential_deformation"])
    trajectory = _data.JaxSimModelData(
        velocity_representation=data.velocity_representation,
        _joint_positions=traj["s"], _joint_velocities=traj["sd"], _base_quaternion=traj["q"], _base_linear_velocity=traj["vl"],
        _base_angular_velocity=traj["om"], _base_position=traj["p"], _base_transform=None, _joint_transforms=None,
        _link_transforms=None, _link_velocities=None, contact_state={"tangential_deformation": traj["m"]} if soft else {},
    )
    return cur, trajectory


def forward_dynamics_aba(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    *,
    joint_forces: torch.Tensor | None = None,
    link_forces: torch.Tensor | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """``js.model.forward_dynamics_aba`` (``src/jaxsim/api/model.py:1269-1406``): returns the
    base acceleration ``(B, 6)`` in ``data.velocity_representation`` and the joint
    accelerations ``(B, n)``.  ``link_forces`` are expressed in that representation too
    (``:1315-1321``).  The kernel works inertial-fixed; the Body/Mixed ``to_active``
    conversion of ``:1356-1404`` is a few batched torch ops around it."""
    s = _batched(data._joint_positions, 1).contiguous()
    dev, dtype = s.device, s.dtype
    dm = model.device_model(dev)
    nL, n = model.number_of_links(), model.dofs()
    sd = _batched(data._joint_velocities, 1).contiguous()
    q = _batched(data._base_quaternion, 1).contiguous()
    vl = _batched(data._base_linear_velocity, 1).contiguous()
    om = _batched(data._base_angular_velocity, 1).contiguous()
    p = _batched(data._base_position, 1).contiguous()
    B = q.shape[0]
    tau = None if joint_forces is None else _batched(torch.as_tensor(joint_forces, dtype=dtype, device=dev), 1).contiguous()
    fext = None if link_forces is None else _batched(torch.as_tensor(link_forces, dtype=dtype, device=dev), 2).contiguous()
    if tau is not None and tau.shape != (B, n):
        raise ValueError(tau.shape, (B, n))
    if fext is not None and fext.shape != (B, nL, 6):
        raise ValueError(fext.shape, (B, nL, 6))
    vr = data.velocity_representation
    if fext is not None and vr != VelRepr.Inertial:
        fext = other_representation_to_inertial(fext, vr, _batched(data.link_transforms, 3), is_force=True).contiguous()
    avd = torch.empty(B, 6, dtype=dtype, device=dev)
    sdd = torch.empty(B, n, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_aba(
            dm.handle, _dtype_code(dtype), B, _ptr(s), _ptr(sd), _ptr(q), _ptr(vl), _ptr(om), _ptr(p),
            _ptr(tau), _ptr(fext), _ptr(avd), _ptr(sdd), _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_aba")
    if vr != VelRepr.Inertial:
        avd = _base_acceleration_to_active(model, data, avd, vl, om)
    if data._joint_positions.dim() == 1:
        return avd.squeeze(0), sdd.squeeze(0)
    return avd, sdd


def _frame_C(data, vl, om):
    """``W_H_C`` and ``W_v_WC`` of the active representation (``api/model.py:1373-1391``):
    C = B for Body, C = B[W] for Mixed."""
    W_H_B = _batched(data.base_transform, 2)
    W_v_WB = torch.cat([vl, om], dim=-1)
    if data.velocity_representation == VelRepr.Body:
        return W_H_B, W_v_WB, W_v_WB
    W_H_C = W_H_B.clone()
    W_H_C[..., 0:3, 0:3] = torch.eye(3, dtype=W_H_B.dtype, device=W_H_B.device)
    # linear velocity of the base origin: the mixed base velocity (api/data.py:288-312)
    W_pd_B = vl + torch.linalg.cross(om, W_H_B[..., 0:3, 3])
    W_v_WC = torch.cat([W_pd_B, torch.zeros_like(om)], dim=-1)
    return W_H_C, W_v_WC, W_v_WB


def _cross_vx(v: torch.Tensor) -> torch.Tensor:
    """``Cross.vx`` (``math/cross.py:10-27``): [[S(w), S(v)], [0, S(w)]], batched."""
    from .common import _wedge

    X = torch.zeros(v.shape[:-1] + (6, 6), dtype=v.dtype, device=v.device)
    X[..., 0:3, 0:3] = _wedge(v[..., 3:6])
    X[..., 0:3, 3:6] = _wedge(v[..., 0:3])
    X[..., 3:6, 3:6] = _wedge(v[..., 3:6])
    return X


def _base_acceleration_to_active(model, data, W_vd_WB, vl, om):
    """``to_active`` (``api/model.py:1356-1404``): C_X_W (W_vd_WB - W_v_WC x W_v_WB)."""
    from .common import adjoint_from_transform

    if not model.floating_base():
        return torch.zeros_like(W_vd_WB)
    W_H_C, W_v_WC, W_v_WB = _frame_C(data, vl, om)
    C_X_W = adjoint_from_transform(W_H_C, inverse=True)
    return torch.einsum("...ij,...j->...i", C_X_W, W_vd_WB - torch.einsum("...ij,...j->...i", _cross_vx(W_v_WC), W_v_WB))


def inverse_dynamics(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    *,
    joint_accelerations: torch.Tensor | None = None,
    base_acceleration: torch.Tensor | None = None,
    link_forces: torch.Tensor | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """``js.model.inverse_dynamics`` (``src/jaxsim/api/model.py:1746-1894``) -> vmapped
    ``rbda.rnea`` (``rbda/rnea.py:12-238``): returns the 6D base force ``(B, 6)`` in
    ``data.velocity_representation`` and the joint forces ``(B, n)``.  ``base_acceleration``
    and ``link_forces`` are expressed in that representation (``to_inertial``, ``:1800-1845``)."""
    s = _batched(data._joint_positions, 1).contiguous()
    dev, dtype = s.device, s.dtype
    dm = model.device_model(dev)
    nL, n = model.number_of_links(), model.dofs()
    sd = _batched(data._joint_velocities, 1).contiguous()
    q = _batched(data._base_quaternion, 1).contiguous()
    vl = _batched(data._base_linear_velocity, 1).contiguous()
    om = _batched(data._base_angular_velocity, 1).contiguous()
    p = _batched(data._base_position, 1).contiguous()
    B = q.shape[0]

    def opt(x, shape, nd):
        if x is None:
            return None
        x = _batched(torch.as_tensor(x, dtype=dtype, device=dev), nd).contiguous()
        if x.shape != shape:
            raise ValueError(x.shape, shape)
        return x

    sdd = opt(joint_accelerations, (B, n), 1)
    avd = opt(base_acceleration, (B, 6), 1)
    fext = opt(link_forces, (B, nL, 6), 2)
    vr = data.velocity_representation
    if vr != VelRepr.Inertial:
        from .common import adjoint_from_transform

        if fext is not None:
            fext = other_representation_to_inertial(fext, vr, _batched(data.link_transforms, 3), is_force=True).contiguous()
        # to_inertial (:1800-1845): W_X_C (C_vd_WB + C_v_WC x C_v_WB)
        W_H_C, W_v_WC, W_v_WB = _frame_C(data, vl, om)
        C_X_W = adjoint_from_transform(W_H_C, inverse=True)
        mv = lambda X, v: torch.einsum("...ij,...j->...i", X, v)  # noqa: E731
        C_v_WB = _batched(data.base_velocity, 1)
        acc = avd if avd is not None else torch.zeros(B, 6, dtype=dtype, device=dev)
        avd = mv(adjoint_from_transform(W_H_C), acc + mv(_cross_vx(mv(C_X_W, W_v_WC)), C_v_WB)).contiguous()
    W_f = torch.empty(B, 6, dtype=dtype, device=dev)
    tau = torch.empty(B, n, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_rnea(
            dm.handle, _dtype_code(dtype), B, _ptr(s), _ptr(sd), _ptr(q), _ptr(vl), _ptr(om), _ptr(p),
            _ptr(avd), _ptr(sdd), _ptr(fext), _ptr(W_f), _ptr(tau), _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_rnea")
    if vr != VelRepr.Inertial:
        W_f = inertial_to_other_representation(W_f, vr, _batched(data.base_transform, 2), is_force=True)
    if data._joint_positions.dim() == 1:
        return W_f.squeeze(0), tau.squeeze(0)
    return W_f, tau


def free_floating_mass_matrix(model: JaxSimModel, data: "_data.JaxSimModelData") -> torch.Tensor:
    """``js.model.free_floating_mass_matrix`` (``src/jaxsim/api/model.py:1556-1592``):
    CRBA (``rbda/crba.py:10-170``) in body-fixed representation, then ``_transform_M_block``
    (``:1527-1553``) into the active representation."""
    from .common import adjoint_from_transform

    s = _batched(data._joint_positions, 1).contiguous()
    dev, dtype = s.device, s.dtype
    dm = model.device_model(dev)
    n = model.dofs()
    B = s.shape[0] if n > 0 else _batched(data._base_quaternion, 1).shape[0]
    M = torch.empty(B, 6 + n, 6 + n, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_crba(dm.handle, _dtype_code(dtype), B, _ptr(s), _ptr(M), _stream_ptr(dev))
    _lib.check(rc, "b200sim_crba")
    vr = data.velocity_representation
    if vr != VelRepr.Body:
        H = _batched(data.base_transform, 2)
        if vr == VelRepr.Mixed:
            H = H.clone()
            H[..., 0:3, 3] = 0
        X = adjoint_from_transform(H, inverse=True)  # B_X_W or B_X_BW
        Xt = X.transpose(-1, -2)
        Mt = torch.empty_like(M)
        Mt[:, :6, :6] = Xt @ M[:, :6, :6] @ X
        Mt[:, :6, 6:] = Xt @ M[:, :6, 6:]
        Mt[:, 6:, :6] = M[:, 6:, :6] @ X
        Mt[:, 6:, 6:] = M[:, 6:, 6:]
        M = Mt
    return M.squeeze(0) if data._joint_positions.dim() == 1 else M


def free_floating_bias_forces(model: JaxSimModel, data: "_data.JaxSimModelData") -> torch.Tensor:
    """``js.model.free_floating_bias_forces`` (``src/jaxsim/api/model.py:1934-1979``): h(q, nu) =
    inverse dynamics with zero accelerations and no external forces, ``(B, 6+n)`` in the data's
    velocity representation."""
    f, tau = inverse_dynamics(model, data)
    return torch.cat([f, tau], dim=-1)


def free_floating_gravity_forces(model: JaxSimModel, data: "_data.JaxSimModelData") -> torch.Tensor:
    """``js.model.free_floating_gravity_forces`` (``:1896-1931``): g(q) = h(q, 0)."""
    d0 = data.copy()
    for leaf in ("_joint_velocities", "_base_linear_velocity", "_base_angular_velocity"):
        setattr(d0, leaf, torch.zeros_like(getattr(data, leaf)))
    f, tau = inverse_dynamics(model, d0)
    return torch.cat([f, tau], dim=-1)


def free_floating_mass_matrix_inverse(model: JaxSimModel, data: "_data.JaxSimModelData") -> torch.Tensor:
    """``js.model.free_floating_mass_matrix_inverse`` (``:1595-1631``).  The reference propagates
    articulated inertias (``rbda/mass_inverse.py``); M is symmetric positive definite, so this is
    ``inv(free_floating_mass_matrix)`` up to rounding (one batched inverse) -- of the full
    ``(6+n, 6+n)`` matrix also for fixed-base models, like the reference."""
    return torch.linalg.inv(free_floating_mass_matrix(model, data))


def total_mass(model: JaxSimModel) -> float:
    """``js.model.total_mass`` (``:1023-1035``)."""
    return float(np.asarray(model.kin_dyn_parameters.link_parameters.mass).sum())


_JVP_LEAVES = (
    ("joint_positions", "_joint_positions"), ("joint_velocities", "_joint_velocities"),
    ("base_quaternion", "_base_quaternion"), ("base_linear_velocity", "_base_linear_velocity"),
    ("base_angular_velocity", "_base_angular_velocity"), ("base_position", "_base_position"),
)


def step_jvp(
    model: JaxSimModel,
    data: "_data.JaxSimModelData",
    tangents: dict,
    *,
    joint_force_references: torch.Tensor | None = None,
    n_steps: int = 1,
    update_caches: bool = True,
    mass_direction_period: int = 0,
    mass_direction_first_link: int = 0,
) -> tuple["_data.JaxSimModelData", "_data.JaxSimModelData"]:
    """Forward-mode derivative of ``step`` (BASELINE config 5): the counterpart of
    ``jax.jvp(lambda theta: js.model.step(model(theta), data(theta)), ...)`` checked by the
    reference in ``tests/test_automatic_differentiation.py:346-420``.

    ``tangents`` maps input names to tangent arrays: any of ``joint_positions``,
    ``joint_velocities``, ``base_quaternion``, ``base_linear_velocity``,
    ``base_angular_velocity``, ``base_position`` (inertial-fixed, shapes of the leaves),
    ``tangential_deformation`` ``(B, nc, 3)``, ``joint_force_references`` ``(B, n)`` and
    ``link_masses`` ``(nL,)`` -- the ``LinkParameters.mass`` leaf, entering through
    ``Inertia.to_sixd`` with CoM and CoM-inertia held fixed.
    Missing entries are zero.  float64, batched data, ``VelRepr`` is irrelevant (no link
    forces).  Returns ``(data_out, tangent_out)``: the stepped data and a data object whose
    leaves (state, contact state and caches) hold the directional derivatives.  ``update_caches=False`` skips the
    caches of both (the kernel then neither computes nor stores the kinematics of the new state).  With
    ``mass_direction_period = P > 0`` the batch is read as replicas of ``P`` environments and replica ``r`` takes the
    ``link_masses`` direction of link ``mass_direction_first_link + r`` alone (``b200sim_step_jvp_ex``: many mass columns
    in one launch)."""
    q = data._base_quaternion
    if q.dim() != 2 or q.dtype != torch.float64:
        raise ValueError("step_jvp needs batched float64 data")
    unknown = set(tangents) - {k for k, _ in _JVP_LEAVES} - {"tangential_deformation", "joint_force_references", "link_masses"}
    if unknown:
        raise KeyError(f"unknown tangent inputs: {sorted(unknown)}")
    dev = q.device
    dm = model.device_model(dev)
    B = q.shape[0]
    nL, n, nc = model.number_of_links(), model.dofs(), model.number_of_collidable_points()
    soft = isinstance(model.contact_model, SoftContacts)

    def pack(val, tan):
        tan = torch.zeros_like(val) if tan is None else torch.as_tensor(tan, dtype=torch.float64, device=dev).expand_as(val)
        return torch.stack([val, tan], dim=-1).contiguous()

    ins = [pack(getattr(data, leaf), tangents.get(name)) for name, leaf in _JVP_LEAVES]
    m = data.contact_state.get("tangential_deformation") if data.contact_state else None
    if m is None and soft:
        m = torch.zeros(B, nc, 3, dtype=torch.float64, device=dev)
    m_in = pack(m, tangents.get("tangential_deformation")) if m is not None else None
    tau_in = None
    if joint_force_references is not None or "joint_force_references" in tangents:
        tv = torch.zeros(B, n, dtype=torch.float64, device=dev) if joint_force_references is None else torch.as_tensor(
            joint_force_references, dtype=torch.float64, device=dev).expand(B, n)
        tau_in = pack(tv, tangents.get("joint_force_references"))
    dmass = None
    if "link_masses" in tangents:
        dmass = np.ascontiguousarray(torch.as_tensor(tangents["link_masses"]).detach().cpu().numpy(), dtype=np.float64)
        if dmass.shape != (nL,):
            raise ValueError(dmass.shape, (nL,))
    new = lambda *shape: torch.empty(*shape, 2, dtype=torch.float64, device=dev)  # noqa: E731
    outs = [new(B, n), new(B, n), new(B, 4), new(B, 3), new(B, 3), new(B, 3)]
    m_o = new(B, nc, 3) if soft else None
    caches = [new(B, 4, 4), new(B, nL, 6, 6), new(B, nL, 4, 4), new(B, nL, 6)] if update_caches else [None] * 4
    with torch.cuda.device(dev):
        rc = _lib.load().b200sim_step_jvp_ex(
            dm.handle, B, int(n_steps), None if dmass is None else dmass.ctypes.data_as(_lib.c_dp),
            int(mass_direction_period), int(mass_direction_first_link),
            *[_ptr(t) for t in ins], _ptr(m_in), _ptr(tau_in),
            *[_ptr(t) for t in outs], _ptr(m_o), *[_ptr(t) for t in caches], _stream_ptr(dev),
        )
    _lib.check(rc, "b200sim_step_jvp")

    def split(idx):
        mk = lambda t: None if t is None else t[..., idx].contiguous()  # noqa: E731
        cs = {"tangential_deformation": mk(m_o)} if soft else {}
        return _data.JaxSimModelData(
            velocity_representation=data.velocity_representation,
            _joint_positions=mk(outs[0]), _joint_velocities=mk(outs[1]), _base_quaternion=mk(outs[2]),
            _base_linear_velocity=mk(outs[3]), _base_angular_velocity=mk(outs[4]), _base_position=mk(outs[5]),
            _base_transform=mk(caches[0]), _joint_transforms=mk(caches[1]), _link_transforms=mk(caches[2]),
            _link_velocities=mk(caches[3]), contact_state=cs,
        )

    return split(0), split(1)
