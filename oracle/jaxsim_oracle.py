"""CPU oracle: a literal NumPy restatement of the reference's ``jaxsim.api.model.step`` path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``jaxsim_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

Every function follows one reference function (cited ``file:line`` relative to
``/root/reference/src/jaxsim``) *as written there*: dense 6x6 spatial algebra in
body-fixed link frames, ``[linear; angular]`` ordering, the same order of operations --
deliberately NOT the world-aligned/structured formulation the CUDA kernels use, so that
agreement between the two is evidence, not tautology.  The only liberties taken:

* every array carries a leading batch axis ``B`` (the reference gets it from ``jax.vmap``);
* ``lax.scan`` loops are Python ``for`` loops over links;
* ``jaxlie`` (third-party, ``>=1.3.0``, not vendored, unpinned) is restated from its published
  formulas at the reference's call sites: ``SO3(wxyz).as_matrix()`` (quaternion -> DCM with
  the ``2/|q|^2`` scaling), ``SE3.adjoint() = [[R, S(p)R],[0, R]]``,
  ``SE3.inverse() = (R^T, -R^T p)``.  ``SE3.from_matrix(H)`` followed by
  ``.adjoint()/.as_matrix()`` is treated as the identity round trip it is mathematically
  (jaxlie goes through a quaternion; the difference is rounding-level);
* ``dtype`` selects float64 (reference default, ``__init__.py:19,35``) or float32
  (``JAX_ENABLE_X64=0``), incl. ``finfo(dtype).eps`` where the reference uses ``finfo(float)``.

PARITY PIN STATUS: pinned by the reference's own code.  JAX is not installable offline, but
the UNMODIFIED reference sources are executed over NumPy stand-ins of their third-party
imports (``oracle/refshim``, ``tests/golden/make_goldens.py``); the committed fixtures
``tests/golden/*.npz`` hold the reference's outputs and ``tests/test_reference_goldens.py``
checks every leaf of this oracle's ``step`` / ``step_rk4`` / ``aba`` / ``rnea`` / ``crba`` /
``system_dynamics`` against them at 1e-9 relative (float64).  In addition the reference's
*known-answer and invariant* tests are re-run against this restatement in
``tests/test_oracle_pins.py``: ABA == CRB forward dynamics and RNEA(ABA(tau)) == tau
(``tests/test_api_model.py:495-577``), balanced box (``tests/test_simulations.py:15-85``),
ballistic box (``:88-167``), soft-contact rest height (``:194-242``), joint limits
(``:347-401``), torque-speed curve (``tests/test_actuation.py:11-48``), plus analytic
pendulum/free-fall answers.  Not covered by the pin: XLA's float32 rounding / fusion (the
fixtures are float64, evaluated eagerly).
"""

from __future__ import annotations

import dataclasses

import numpy as np

# =============================================================================
# math/*  (batched: every input has leading dims "...")
# =============================================================================


def wedge(v):
    """``Skew.wedge`` (``math/skew.py:12-37``)."""
    v = np.asarray(v)
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    o = np.zeros_like(x)
    return np.stack(
        [np.stack([o, -z, y], -1), np.stack([z, o, -x], -1), np.stack([-y, x, o], -1)], -2
    )


def vee(m):
    """``Skew.vee`` (``math/skew.py:39-58``)."""
    return 0.5 * np.stack(
        [m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]], -1
    )


def safe_norm(a, axis=-1, keepdims=False):
    """``safe_norm`` forward pass (``math/utils.py:7-21``) = plain norm."""
    return np.linalg.norm(a, axis=axis, keepdims=keepdims)


def quat_to_dcm(q):
    """``jaxlie.SO3(wxyz=q).as_matrix()`` (call sites ``rbda/aba.py:79-86``,
    ``math/quaternion.py:52``, ``math/transform.py:42-49``)."""
    q = np.asarray(q)
    norm_sq = np.sum(q * q, axis=-1)
    q2 = q[..., :, None] * q[..., None, :] * (2.0 / norm_sq)[..., None, None]
    one = np.ones_like(norm_sq)
    R = np.stack(
        [
            np.stack([one - q2[..., 2, 2] - q2[..., 3, 3], q2[..., 1, 2] - q2[..., 3, 0], q2[..., 1, 3] + q2[..., 2, 0]], -1),
            np.stack([q2[..., 1, 2] + q2[..., 3, 0], one - q2[..., 1, 1] - q2[..., 3, 3], q2[..., 2, 3] - q2[..., 1, 0]], -1),
            np.stack([q2[..., 1, 3] - q2[..., 2, 0], q2[..., 2, 3] + q2[..., 1, 0], one - q2[..., 1, 1] - q2[..., 2, 2]], -1),
        ],
        -2,
    )
    return R.astype(q.dtype)


def transform_from_quat_pos(q, p):
    """``Transform.from_quaternion_and_translation`` (``math/transform.py:14-56``)."""
    R = quat_to_dcm(q)
    H = np.zeros(R.shape[:-2] + (4, 4), dtype=R.dtype)
    H[..., 0:3, 0:3] = R
    H[..., 0:3, 3] = p
    H[..., 3, 3] = 1.0
    return H


def adjoint_from_Rp(R, p, inverse=False):
    """``Adjoint.from_rotation_and_translation`` (``math/adjoint.py:66-107``)."""
    X = np.zeros(R.shape[:-2] + (6, 6), dtype=R.dtype)
    if not inverse:
        X[..., 0:3, 0:3] = R
        X[..., 0:3, 3:6] = wedge(p) @ R
        X[..., 3:6, 3:6] = R
    else:
        Rt = np.swapaxes(R, -1, -2)
        X[..., 0:3, 0:3] = Rt
        X[..., 0:3, 3:6] = -Rt @ wedge(p)
        X[..., 3:6, 3:6] = Rt
    return X


def adjoint_from_transform(H, inverse=False):
    """``Adjoint.from_transform`` (``math/adjoint.py:46-64``): jaxlie SE3(H).adjoint() or
    SE3(H).inverse().adjoint()."""
    return adjoint_from_Rp(H[..., 0:3, 0:3], H[..., 0:3, 3], inverse=inverse)


def adjoint_inverse(X):
    """``Adjoint.inverse`` (``math/adjoint.py:135-160``)."""
    Rt = np.swapaxes(X[..., 0:3, 0:3], -1, -2)
    T = X[..., 0:3, 3:6]
    out = np.zeros_like(X)
    out[..., 0:3, 0:3] = Rt
    out[..., 0:3, 3:6] = -Rt @ T @ Rt
    out[..., 3:6, 3:6] = Rt
    return out


def adjoint_to_transform(X):
    """``Adjoint.to_transform`` (``math/adjoint.py:109-133``)."""
    R = X[..., 0:3, 0:3]
    oxR = X[..., 0:3, 3:6]
    H = np.zeros(X.shape[:-2] + (4, 4), dtype=X.dtype)
    H[..., 0:3, 0:3] = R
    H[..., 0:3, 3] = vee(oxR @ np.swapaxes(R, -1, -2))
    H[..., 3, 3] = 1.0
    return H


def cross_vx(v6):
    """``Cross.vx`` (``math/cross.py:13-41``)."""
    v, w = v6[..., 0:3], v6[..., 3:6]
    X = np.zeros(v6.shape[:-1] + (6, 6), dtype=v6.dtype)
    X[..., 0:3, 0:3] = wedge(w)
    X[..., 0:3, 3:6] = wedge(v)
    X[..., 3:6, 3:6] = wedge(w)
    return X


def cross_vx_star(v6):
    """``Cross.vx_star`` (``math/cross.py:43-58``)."""
    return -np.swapaxes(cross_vx(v6), -1, -2)


def inertia_to_sixd(mass, com, I):
    """``Inertia.to_sixd`` (``math/inertia.py:14-41``)."""
    c = wedge(com)
    M = np.zeros(I.shape[:-2] + (6, 6), dtype=I.dtype)
    eye = np.eye(3, dtype=I.dtype)
    m = np.asarray(mass)[..., None, None]
    M[..., 0:3, 0:3] = m * eye
    M[..., 0:3, 3:6] = m * np.swapaxes(c, -1, -2)
    M[..., 3:6, 0:3] = m * c
    M[..., 3:6, 3:6] = I + m * (c @ np.swapaxes(c, -1, -2))
    return M


def rotation_from_axis_angle(vector):
    """``Rotation.from_axis_angle`` (``math/rotation.py:58-84``)."""
    theta = safe_norm(vector)
    s, c = np.sin(theta), np.cos(theta)
    c1 = 2 * np.sin(theta / 2.0) ** 2
    safe_theta = np.where(theta == 0, 1.0, theta)
    u = vector / safe_theta[..., None]
    eye = np.eye(3, dtype=vector.dtype)
    R = c[..., None, None] * eye - s[..., None, None] * wedge(u) + c1[..., None, None] * (u[..., :, None] * u[..., None, :])
    return np.swapaxes(R, -1, -2)


def quaternion_derivative(q, omega, K=0.1):
    """``Quaternion.derivative`` with ``omega_in_body_fixed=False``
    (``math/quaternion.py:68-132``)."""
    qw, qx, qy, qz = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    Q = np.stack(
        [
            np.stack([qw, -qx, -qy, -qz], -1),
            np.stack([qx, qw, qz, -qy], -1),
            np.stack([qy, -qz, qw, qx], -1),
            np.stack([qz, qy, -qx, qw], -1),
        ],
        -2,
    )
    norm_w = safe_norm(omega)
    vec = np.concatenate([(K * norm_w * (1 - safe_norm(q)))[..., None], omega], axis=-1)
    return 0.5 * np.einsum("...ij,...j->...i", Q, vec)


# =============================================================================
# model container used by the oracle (plain data; mirrors JaxSimModel static fields)
# =============================================================================


@dataclasses.dataclass
class OracleModel:
    """The subset of ``JaxSimModel`` (``api/model.py:46-90``) the step reads."""

    kin_dyn_parameters: object  # jaxsim_b200.api.kin_dyn_parameters.KinDynParameters
    floating_base: bool
    time_step: float = 0.001
    gravity: float = -9.81  # model.gravity (api/model.py:62,206): NEGATIVE
    terrain_height: float = 0.0  # FlatTerrain (terrain/terrain.py:66-113)
    contact_model: str = "soft"  # "soft" | "rigid" | "relaxed" (oracle/rigid_oracle.py) | "none"
    # RelaxedRigidContactsParams (rbda/contacts/relaxed_rigid.py:30-82); mu above is shared
    relaxed: dict = dataclasses.field(default_factory=lambda: dict(
        time_constant=0.02, damping_coefficient=1.0, d_min=0.9, d_max=0.95, width=0.001, midpoint=0.5, power=2.0))
    regularization_delassus: float = 1e-6  # RigidContacts (rbda/contacts/rigid.py:99-101)
    # SoftContactsParams (rbda/contacts/soft.py:24-46)
    K: float = 1e6
    D: float = 2000.0
    mu: float = 0.5
    p: float = 0.5
    q: float = 0.5
    # ActuationParams (rbda/actuation/common.py:10-19)
    torque_max: float = 3000.0
    omega_th: float = 30.0
    omega_max: float = 100.0
    enable_friction: bool = True

    def number_of_links(self):
        return self.kin_dyn_parameters.number_of_links()

    def dofs(self):
        return self.kin_dyn_parameters.number_of_joints()


def link_spatial_inertia_matrices(model: OracleModel, dtype):
    """``js.model.link_spatial_inertia_matrices`` (``api/model.py:902-917``)."""
    lp = model.kin_dyn_parameters.link_parameters
    return inertia_to_sixd(
        lp.mass.astype(dtype), lp.center_of_mass.astype(dtype), lp.inertia_tensors().astype(dtype)
    )


# =============================================================================
# api/kin_dyn_parameters.py:396-451 + math/joint_model.py:146-200
# =============================================================================


def joint_transforms(model: OracleModel, s, W_H_B):
    """``KinDynParameters.joint_transforms``: (B,nL,6,6) adjoints i_X_lambda(i)."""
    kd = model.kin_dyn_parameters
    jm = kd.joint_model
    dtype = s.dtype
    B = s.shape[0]
    nL = kd.number_of_links()
    out = np.zeros((B, nL, 6, 6), dtype=dtype)
    eye4 = np.eye(4, dtype=dtype)
    for i in range(nL):
        if i == 0:
            lam_H_pre = eye4
            pre_H_suc = W_H_B
        else:
            lam_H_pre = jm.lam_H_pre[i].astype(dtype)
            jt = jm.joint_types[i]
            axis = jm.joint_axis[i - 1].astype(dtype)
            pre_H_suc = np.tile(eye4, (B, 1, 1))
            if jt == 1:  # revolute: supported_joint_motion.compute_R
                pre_H_suc[:, 0:3, 0:3] = rotation_from_axis_angle(s[:, i - 1, None] * axis[None, :])
            elif jt == 2:  # prismatic: compute_P
                pre_H_suc[:, 0:3, 3] = s[:, i - 1, None] * axis[None, :]
        suc_H_i = jm.suc_H_i[i].astype(dtype)
        H = lam_H_pre @ pre_H_suc @ suc_H_i
        out[:, i] = adjoint_from_transform(H, inverse=True)
    return out


# =============================================================================
# rbda/forward_kinematics.py:12-113
# =============================================================================


def forward_kinematics_model(model: OracleModel, p, q, s, v_lin, omega, sd):
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL = kd.number_of_links()
    dtype = s.dtype
    B = s.shape[0]
    W_H_B = transform_from_quat_pos(q, p)
    i_X_lam = joint_transforms(model, s, W_H_B)
    W_X_i = np.zeros((B, nL, 6, 6), dtype=dtype)
    W_X_i[:, 0] = adjoint_inverse(i_X_lam[:, 0])
    W_v = np.zeros((B, nL, 6), dtype=dtype)
    W_v[:, 0] = np.concatenate([v_lin, omega], axis=-1)
    S = kd.motion_subspaces.astype(dtype)
    for i in range(1, nL):
        lam_X_i = adjoint_inverse(i_X_lam[:, i])
        W_X_i[:, i] = W_X_i[:, lam[i]] @ lam_X_i
        W_v[:, i] = W_v[:, lam[i]] + np.einsum("bij,bj->bi", W_X_i[:, i], S[i][None, :] * sd[:, i - 1, None])
    return adjoint_to_transform(W_X_i), W_v


# =============================================================================
# api/data.py  (state container + cached quantities)
# =============================================================================


@dataclasses.dataclass
class OracleData:
    """Leaves of ``JaxSimModelData`` (``api/data.py:47-63``), batched."""

    joint_positions: np.ndarray  # (B,n)
    joint_velocities: np.ndarray  # (B,n)
    base_quaternion: np.ndarray  # (B,4) wxyz
    base_linear_velocity: np.ndarray  # (B,3) inertial-fixed
    base_angular_velocity: np.ndarray  # (B,3)
    base_position: np.ndarray  # (B,3)
    base_transform: np.ndarray = None  # (B,4,4)
    joint_transforms: np.ndarray = None  # (B,nL,6,6)
    link_transforms: np.ndarray = None  # (B,nL,4,4)
    link_velocities: np.ndarray = None  # (B,nL,6)
    tangential_deformation: np.ndarray = None  # (B,nc_total,3)

    @property
    def base_orientation(self):
        """``JaxSimModelData.base_orientation`` (``api/data.py:267-286``)."""
        q = self.base_quaternion
        norm = safe_norm(q, axis=-1, keepdims=True)
        return q / (norm + np.finfo(q.dtype).eps * (norm == 0))


def data_replace(model: OracleModel, s, sd, q, v_lin, omega, p, m=None) -> OracleData:
    """``JaxSimModelData.replace(model=...)`` (``api/data.py:406-523``): normalise the
    quaternion and recompute all caches (also what ``build`` does, ``:66-202``)."""
    norm = safe_norm(q, axis=-1, keepdims=True)
    q = q / np.where(norm == 0, 1.0, norm)
    W_H_B = transform_from_quat_pos(q, p)
    i_X_lam = joint_transforms(model, s, W_H_B)
    W_H_L, W_v_WL = forward_kinematics_model(model, p, q, s, v_lin, omega, sd)
    if m is None:
        nc = len(model.kin_dyn_parameters.contact_parameters.body)
        m = np.zeros((s.shape[0], nc, 3), dtype=s.dtype)
    return OracleData(
        joint_positions=s, joint_velocities=sd, base_quaternion=q, base_linear_velocity=v_lin,
        base_angular_velocity=omega, base_position=p, base_transform=W_H_B, joint_transforms=i_X_lam,
        link_transforms=W_H_L, link_velocities=W_v_WL, tangential_deformation=m,
    )


# =============================================================================
# contacts
# =============================================================================


def collidable_points_pos_vel(model: OracleModel, W_H_L, W_v_WL):
    """``rbda/collidable_points.py:9-65`` (enabled points only)."""
    cp = model.kin_dyn_parameters.contact_parameters
    idx = cp.indices_of_enabled_collidable_points
    body = np.array(cp.body, dtype=int)[idx]
    L_p = cp.point[idx].astype(W_H_L.dtype)
    B = W_H_L.shape[0]
    ph = np.concatenate([L_p, np.ones((len(idx), 1), dtype=L_p.dtype)], axis=-1)
    W_p = np.einsum("bcij,cj->bci", W_H_L[:, body], ph)[..., 0:3]
    v = W_v_WL[:, body]
    # [I, -S(p)] @ v
    W_pd = v[..., 0:3] - np.einsum("bcij,bcj->bci", wedge(W_p), v[..., 3:6])
    return W_p, W_pd


def compute_penetration_data(model: OracleModel, p, v):
    """``rbda/contacts/common.py:25-63`` with ``FlatTerrain`` (``terrain/terrain.py:66-113``)."""
    dtype = p.dtype
    n_hat = np.zeros_like(p)
    n_hat[..., 2] = 1.0
    h = np.zeros_like(p)
    h[..., 2] = dtype.type(model.terrain_height) - p[..., 2]
    delta = np.maximum(0.0, np.sum(h * n_hat, axis=-1))
    delta_dot = -np.sum(v * n_hat, axis=-1)
    delta_dot = np.where(delta > 0, delta_dot, 0.0)
    return delta.astype(dtype), delta_dot.astype(dtype), n_hat


def hunt_crossley_contact_model(model: OracleModel, W_p_C, W_pd_C, m):
    """``SoftContacts.hunt_crossley_contact_model`` (``rbda/contacts/soft.py:195-339``)."""
    dtype = W_p_C.dtype
    K, D, mu = dtype.type(model.K), dtype.type(model.D), dtype.type(model.mu)
    pexp, qexp = dtype.type(model.p), dtype.type(model.q)
    delta, delta_dot, n_hat = compute_penetration_data(model, W_p_C, W_pd_C)
    eps = np.finfo(dtype).eps
    dp = np.power(delta + eps, pexp)
    dq = np.power(delta + eps, qexp)
    force_normal_mag = (K * dp) * delta + (D * dq) * delta_dot
    force_normal_mag = np.maximum(0.0, force_normal_mag)
    f_normal = force_normal_mag[..., None] * n_hat
    dot = lambda a, b: np.sum(a * b, axis=-1)  # noqa: E731
    v_tangential = W_pd_C - dot(W_pd_C, n_hat)[..., None] * n_hat
    m_normal = dot(m, n_hat)[..., None] * n_hat
    m_tangential = m - dot(m, n_hat)[..., None] * n_hat
    f_tangential = -((K * dp)[..., None] * m_tangential + (D * dq)[..., None] * v_tangential)
    sticking = np.logical_or(delta <= 0, dot(f_tangential, f_tangential) <= (mu * force_normal_mag) ** 2)
    norm = safe_norm(f_tangential)
    f_tangential_direction = f_tangential / (norm + np.finfo(dtype).eps * (norm == 0))[..., None]
    f_tangential = np.where(
        sticking[..., None],
        f_tangential,
        np.minimum(mu * force_normal_mag, norm)[..., None] * f_tangential_direction,
    )
    f_tangential = np.where((delta <= 0)[..., None], 0.0, f_tangential)
    m_dot_no_contact = -(K / D) * m
    m_dot_sticking = v_tangential - (K / D) * m_normal
    m_dot_slipping = -(f_tangential + (K * dp)[..., None] * m_tangential) / (D * dq)[..., None]
    contact_status = sticking.astype(int) + (delta <= 0).astype(int)
    m_dot = np.where(
        (contact_status == 0)[..., None], m_dot_slipping,
        np.where((contact_status == 1)[..., None], m_dot_sticking, m_dot_no_contact),
    )
    return (f_normal + f_tangential).astype(dtype), m_dot.astype(dtype)


def soft_compute_contact_forces(model: OracleModel, data: OracleData):
    """``SoftContacts.compute_contact_forces`` (``rbda/contacts/soft.py:390-444``) +
    ``compute_contact_force`` (``:341-388``)."""
    cp = model.kin_dyn_parameters.contact_parameters
    idx = cp.indices_of_enabled_collidable_points
    W_p_C, W_pd_C = collidable_points_pos_vel(model, data.link_transforms, data.link_velocities)
    m = data.tangential_deformation
    m_enabled = m[:, idx]
    m_dot = np.zeros_like(m)
    CW_fl, m_dot_enabled = hunt_crossley_contact_model(model, W_p_C, W_pd_C, m_enabled)
    # W_Xf_CW = Adjoint(translation=p, inverse=True).T ; W_f = W_Xf_CW @ [CW_fl; 0]
    CW_f = np.concatenate([CW_fl, np.zeros_like(CW_fl)], axis=-1)
    eye = np.broadcast_to(np.eye(3, dtype=W_p_C.dtype), W_p_C.shape[:-1] + (3, 3))
    W_Xf_CW = np.swapaxes(adjoint_from_Rp(eye, W_p_C, inverse=True), -1, -2)
    W_f = np.einsum("bcij,bcj->bci", W_Xf_CW, CW_f)
    m_dot[:, idx] = m_dot_enabled
    return W_f, m_dot


def link_forces_from_contact_forces(model: OracleModel, W_f_C):
    """``api/contact.py:557-603``: ``mask.T @ W_f_C``."""
    cp = model.kin_dyn_parameters.contact_parameters
    idx = cp.indices_of_enabled_collidable_points
    body = np.array(cp.body, dtype=int)[idx]
    mask = (body[:, None] == np.arange(model.number_of_links())[None, :]).astype(W_f_C.dtype)
    return np.einsum("cl,bck->blk", mask, W_f_C)


# =============================================================================
# api/actuation_model.py:7-126
# =============================================================================


def tn_curve_fn(model: OracleModel, sd):
    dt = sd.dtype.type
    tau_max, w_th, w_max = dt(model.torque_max), dt(model.omega_th), dt(model.omega_max)
    abs_vel = np.abs(sd)
    return np.where(
        abs_vel <= w_th, tau_max,
        np.where(abs_vel <= w_max, tau_max * (1 - (abs_vel - w_th) / (w_max - w_th)), 0.0),
    ).astype(sd.dtype)


def compute_resultant_torques(model: OracleModel, s, sd, tau_ref):
    jp = model.kin_dyn_parameters.joint_parameters
    dtype = s.dtype
    tau_position_limit = np.zeros_like(tau_ref)
    if model.dofs() > 0:
        k_j = jp.position_limit_spring.astype(dtype)
        d_j = jp.position_limit_damper.astype(dtype)
        lower = np.clip(s - jp.position_limits_min.astype(dtype), None, 0.0)
        upper = np.clip(s - jp.position_limits_max.astype(dtype), 0.0, None)
        tau_position_limit = tau_position_limit - k_j * (lower + upper)
        # (jnp.positive(tau) * diag(d_j)) @ sd  ==  tau * d_j * sd   (actuation_model.py:64-66)
        tau_position_limit = tau_position_limit - (+tau_position_limit) * d_j * sd
    tau_friction = np.zeros_like(tau_ref)
    if model.dofs() > 0 and model.enable_friction:
        kc = jp.friction_static.astype(dtype)
        kv = jp.friction_viscous.astype(dtype)
        tau_friction = -(kc * np.sign(sd) + kv * sd)
    tau_total = tau_ref + tau_friction + tau_position_limit
    tau_lim = tn_curve_fn(model, sd)
    return np.clip(tau_total, -tau_lim, tau_lim).astype(dtype)


# =============================================================================
# rbda/aba.py:12-292
# =============================================================================


def aba(model: OracleModel, p, q, s, v_lin, omega, sd, tau, W_f):
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL = kd.number_of_links()
    dtype = s.dtype
    B = s.shape[0]
    mv = lambda A, x: np.einsum("bij,bj->bi", A, x)  # noqa: E731

    W_g = np.zeros(6, dtype=dtype)
    W_g[2] = model.gravity
    W_v_WB = np.concatenate([v_lin, omega], axis=-1)
    M = link_spatial_inertia_matrices(model, dtype)
    R = quat_to_dcm(q)
    W_X_B = adjoint_from_Rp(R, p)
    B_X_W = adjoint_from_Rp(R, p, inverse=True)
    W_H_B = transform_from_quat_pos(q, p)
    i_X_lam = joint_transforms(model, s, W_H_B)
    S = kd.motion_subspaces.astype(dtype)

    v = np.zeros((B, nL, 6), dtype=dtype)
    c = np.zeros((B, nL, 6), dtype=dtype)
    pA = np.zeros((B, nL, 6), dtype=dtype)
    MA = np.zeros((B, nL, 6, 6), dtype=dtype)
    i_X_0 = np.zeros((B, nL, 6, 6), dtype=dtype)
    i_X_0[:, 0] = np.eye(6, dtype=dtype)

    if model.floating_base:
        v[:, 0] = mv(B_X_W, W_v_WB)
        MA[:, 0] = M[0]
        pA[:, 0] = mv(cross_vx_star(v[:, 0]) @ MA[:, 0], v[:, 0]) - mv(np.swapaxes(W_X_B, -1, -2), W_f[:, 0])

    # Pass 1 (:131-171)
    for i in range(1, nL):
        ii = i - 1
        vJ = S[i][None, :] * sd[:, ii, None]
        v[:, i] = mv(i_X_lam[:, i], v[:, lam[i]]) + vJ
        c[:, i] = mv(cross_vx(v[:, i]), vJ)
        MA[:, i] = M[i]
        i_X_0[:, i] = i_X_lam[:, i] @ i_X_0[:, lam[i]]
        i_Xf_W = np.swapaxes(adjoint_inverse(i_X_0[:, i] @ B_X_W), -1, -2)
        pA[:, i] = mv(cross_vx_star(v[:, i]) @ M[i][None], v[:, i]) - mv(i_Xf_W, W_f[:, i])

    # Pass 2 (:184-234)
    U = np.zeros((B, nL, 6), dtype=dtype)
    d = np.zeros((B, nL), dtype=dtype)
    u = np.zeros((B, nL), dtype=dtype)
    for i in range(nL - 1, 0, -1):
        ii = i - 1
        U[:, i] = mv(MA[:, i], np.broadcast_to(S[i], (B, 6)))
        d[:, i] = U[:, i] @ S[i]
        u[:, i] = tau[:, ii] - pA[:, i] @ S[i]
        Ma = MA[:, i] - (U[:, i] / d[:, i, None])[:, :, None] * U[:, i][:, None, :]
        pa = pA[:, i] + mv(Ma, c[:, i]) + U[:, i] * (u[:, i] / d[:, i])[:, None]
        if lam[i] != 0 or model.floating_base:
            Xt = np.swapaxes(i_X_lam[:, i], -1, -2)
            MA[:, lam[i]] = MA[:, lam[i]] + Xt @ Ma @ i_X_lam[:, i]
            pA[:, lam[i]] = pA[:, lam[i]] + mv(Xt, pa)

    # Pass 3 (:240-277)
    if model.floating_base:
        a0 = np.linalg.solve(-MA[:, 0], pA[:, 0][..., None])[..., 0]
    else:
        a0 = -mv(B_X_W, np.broadcast_to(W_g, (B, 6)))
    sdd = np.zeros_like(s)
    a = np.zeros_like(v)
    a[:, 0] = a0
    for i in range(1, nL):
        ii = i - 1
        a_i = mv(i_X_lam[:, i], a[:, lam[i]]) + c[:, i]
        sdd[:, ii] = (u[:, i] - np.sum(U[:, i] * a_i, axis=-1)) / d[:, i]
        a[:, i] = a_i + S[i][None, :] * sdd[:, ii, None]

    if model.floating_base:
        W_a_WB = mv(W_X_B, a[:, 0]) + W_g
    else:
        W_a_WB = np.zeros((B, 6), dtype=dtype)
    return W_a_WB.astype(dtype), sdd.astype(dtype)


# =============================================================================
# api/ode.py:16-131, api/integrators.py:14-88, api/model.py:2601-2681
# =============================================================================


def system_acceleration(model: OracleModel, data: OracleData, W_f_L_external, tau_total):
    """``ode.system_acceleration`` in ``VelRepr.Inertial`` (no constraints: ``:83-107`` is a
    no-op when the ConstraintMap is empty, ``rbda/kinematic_constraints.py:215-216``)."""
    f_L = W_f_L_external
    W_f_L_terrain = np.zeros_like(f_L)
    m_dot = np.zeros_like(data.tangential_deformation)
    nc = len(model.kin_dyn_parameters.contact_parameters.body)
    if nc > 0 and model.contact_model == "soft":
        W_f_C, m_dot = soft_compute_contact_forces(model, data)
        W_f_L_terrain = link_forces_from_contact_forces(model, W_f_C)
    W_f_L_total = f_L + W_f_L_terrain
    W_vd_WB, sdd = aba(
        model, data.base_position, data.base_orientation, data.joint_positions,
        data.base_linear_velocity, data.base_angular_velocity, data.joint_velocities,
        tau_total, W_f_L_total,
    )
    return W_vd_WB, sdd, m_dot


def semi_implicit_euler_integration(model: OracleModel, data: OracleData, W_f_L_external, tau_total) -> OracleData:
    dtype = data.joint_positions.dtype
    W_vd_WB, sdd, m_dot = system_acceleration(model, data, W_f_L_external, tau_total)
    dt = dtype.type(model.time_step)
    nu = np.concatenate([data.base_linear_velocity, data.base_angular_velocity, data.joint_velocities], axis=-1)
    nu_new = nu + dt * np.concatenate([W_vd_WB, sdd], axis=-1)
    W_v_B = nu_new[:, 0:6]
    sd = nu_new[:, 6:]
    W_w_WB = nu_new[:, 3:6]
    W_pd_B = nu_new[:, 0:3] + np.einsum("bij,bj->bi", wedge(W_w_WB), data.base_position)
    W_Qd_B = quaternion_derivative(data.base_orientation, W_w_WB)
    W_p_B = data.base_position + dt * W_pd_B
    W_Q_B = data.base_orientation + dt * W_Qd_B
    qn = safe_norm(W_Q_B, axis=-1)
    W_Q_B = W_Q_B / np.where(qn == 0, 1.0, qn)[:, None]
    s = data.joint_positions + dt * sd
    m = data.tangential_deformation + dt * m_dot
    return data_replace(
        model, s.astype(dtype), sd.astype(dtype), W_Q_B.astype(dtype), W_v_B[:, 0:3].astype(dtype),
        W_w_WB.astype(dtype), W_p_B.astype(dtype), m.astype(dtype),
    )


def step(model: OracleModel, data: OracleData, link_forces_inertial=None, joint_force_references=None) -> OracleData:
    """``js.model.step`` (``api/model.py:2601-2681``) for SoftContacts / no contacts +
    SemiImplicitEuler, with ``link_forces`` already in inertial-fixed representation
    (the only representation-dependent input)."""
    B = data.joint_positions.shape[0]
    dtype = data.joint_positions.dtype
    nL, n = model.number_of_links(), model.dofs()
    W_f = np.zeros((B, nL, 6), dtype=dtype) if link_forces_inertial is None else np.asarray(link_forces_inertial, dtype=dtype)
    tau_ref = np.zeros((B, n), dtype=dtype) if joint_force_references is None else np.asarray(joint_force_references, dtype=dtype)
    tau_total = compute_resultant_torques(model, data.joint_positions, data.joint_velocities, tau_ref)
    data_tf = semi_implicit_euler_integration(model, data, W_f, tau_total)
    # SoftContacts.update_velocity_after_impact is the identity (soft.py:179-193)
    return data_tf


def other_representation_to_inertial(array, representation: str, W_H_O, is_force: bool):
    """``api/common.py:160-222``. ``representation`` in {"inertial","body","mixed"}."""
    if representation == "inertial":
        return array
    H = W_H_O.copy()
    if representation == "mixed":
        H[..., 0:3, 0:3] = np.eye(3, dtype=H.dtype)
    if not is_force:
        X = adjoint_from_transform(H)
    else:
        X = np.swapaxes(adjoint_from_transform(H, inverse=True), -1, -2)
    return np.einsum("...ij,...j->...i", X, array)


def inertial_to_other_representation(array, representation: str, W_H_O, is_force: bool):
    """``api/common.py:100-158``."""
    if representation == "inertial":
        return array
    H = W_H_O.copy()
    if representation == "mixed":
        H[..., 0:3, 0:3] = np.eye(3, dtype=H.dtype)
    if not is_force:
        X = adjoint_from_transform(H, inverse=True)
    else:
        X = np.swapaxes(adjoint_from_transform(H), -1, -2)
    return np.einsum("...ij,...j->...i", X, array)


# =============================================================================
# rbda/rnea.py, rbda/crba.py  (used to PIN the oracle: ABA == CRB, RNEA o ABA == id)
# =============================================================================


def rnea(model: OracleModel, p, q, s, v_lin, omega, sd, W_vd_WB, sdd, W_f):
    """``rbda/rnea.py:12-238`` -> (W_f_B (B,6), tau (B,n))."""
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL = kd.number_of_links()
    dtype = s.dtype
    B = s.shape[0]
    mv = lambda A, x: np.einsum("bij,bj->bi", A, x)  # noqa: E731
    W_g = np.zeros(6, dtype=dtype)
    W_g[2] = model.gravity
    W_v_WB = np.concatenate([v_lin, omega], axis=-1)
    M = link_spatial_inertia_matrices(model, dtype)
    R = quat_to_dcm(q)
    W_X_B = adjoint_from_Rp(R, p)
    B_X_W = adjoint_from_Rp(R, p, inverse=True)
    W_H_B = transform_from_quat_pos(q, p)
    i_X_lam = joint_transforms(model, s, W_H_B)
    S = kd.motion_subspaces.astype(dtype)

    v = np.zeros((B, nL, 6), dtype=dtype)
    a = np.zeros((B, nL, 6), dtype=dtype)
    f = np.zeros((B, nL, 6), dtype=dtype)
    i_X_0 = np.zeros((B, nL, 6, 6), dtype=dtype)
    i_X_0[:, 0] = np.eye(6, dtype=dtype)
    # base (rnea.py:108-135)
    a[:, 0] = -mv(B_X_W, np.broadcast_to(W_g, (B, 6)))
    if model.floating_base:
        v[:, 0] = mv(B_X_W, W_v_WB)
        a[:, 0] = mv(B_X_W, W_vd_WB - W_g)
        f[:, 0] = (
            np.einsum("ij,bj->bi", M[0], a[:, 0])
            + mv(cross_vx_star(v[:, 0]) @ M[0][None], v[:, 0])
            - mv(np.swapaxes(W_X_B, -1, -2), W_f[:, 0])
        )
    for i in range(1, nL):
        ii = i - 1
        vJ = S[i][None, :] * sd[:, ii, None]
        v[:, i] = mv(i_X_lam[:, i], v[:, lam[i]]) + vJ
        a[:, i] = mv(i_X_lam[:, i], a[:, lam[i]]) + S[i][None, :] * sdd[:, ii, None] + mv(cross_vx(v[:, i]), vJ)
        i_X_0[:, i] = i_X_lam[:, i] @ i_X_0[:, lam[i]]
        i_Xf_W = np.swapaxes(adjoint_inverse(i_X_0[:, i] @ B_X_W), -1, -2)
        f[:, i] = (
            np.einsum("ij,bj->bi", M[i], a[:, i])
            + mv(cross_vx_star(v[:, i]) @ M[i][None], v[:, i])
            - mv(i_Xf_W, W_f[:, i])
        )
    tau = np.zeros_like(s)
    for i in range(nL - 1, 0, -1):
        ii = i - 1
        tau[:, ii] = f[:, i] @ S[i]
        if lam[i] != 0 or model.floating_base:
            f[:, lam[i]] = f[:, lam[i]] + mv(np.swapaxes(i_X_lam[:, i], -1, -2), f[:, i])
    W_f0 = mv(np.swapaxes(B_X_W, -1, -2), f[:, 0])
    return W_f0.astype(dtype), tau.astype(dtype)


def crba(model: OracleModel, s):
    """``rbda/crba.py:10-170``: free-floating mass matrix in body-fixed representation,
    shape (B, 6+n, 6+n) (rows/cols 0:6 are zero-coupled for fixed-base use)."""
    kd = model.kin_dyn_parameters
    lam = kd.parent_array
    nL = kd.number_of_links()
    n = model.dofs()
    dtype = s.dtype
    B = s.shape[0]
    eye4 = np.broadcast_to(np.eye(4, dtype=dtype), (B, 4, 4))
    i_X_lam = joint_transforms(model, s, eye4)
    S = kd.motion_subspaces.astype(dtype)
    M = link_spatial_inertia_matrices(model, dtype)
    Mc = np.broadcast_to(M, (B, nL, 6, 6)).copy()
    for i in range(nL - 1, 0, -1):
        Xt = np.swapaxes(i_X_lam[:, i], -1, -2)
        Mc[:, lam[i]] = Mc[:, lam[i]] + Xt @ Mc[:, i] @ i_X_lam[:, i]
    Mm = np.zeros((B, 6 + n, 6 + n), dtype=dtype)
    Mm[:, 0:6, 0:6] = Mc[:, 0]
    for i in range(1, nL):
        ii = i - 1
        Fi = np.einsum("bij,j->bi", Mc[:, i], S[i])
        Mm[:, 6 + ii, 6 + ii] = Fi @ S[i]
        j = i
        while True:
            Fi = np.einsum("bji,bj->bi", i_X_lam[:, j], Fi)  # X^T F
            j = lam[j]
            if j == 0:
                break
            jj = j - 1
            Mm[:, 6 + ii, 6 + jj] = Fi @ S[j]
            Mm[:, 6 + jj, 6 + ii] = Mm[:, 6 + ii, 6 + jj]
        Mm[:, 0:6, 6 + ii] = Fi
        Mm[:, 6 + ii, 0:6] = Fi
    return Mm


def forward_dynamics_crb(model: OracleModel, p, q, s, v_lin, omega, sd, tau, W_f):
    """``js.model.forward_dynamics_crb`` (``api/model.py:1409-1498``) restated in
    inertial-fixed I/O: solve M(q) nu_dot = S^T tau - h + J^T f with M from CRBA and the
    bias h from RNEA at zero acceleration.  Returns (W_vd_WB, sdd)."""
    dtype = s.dtype
    B, n = s.shape
    R = quat_to_dcm(q)
    W_X_B = adjoint_from_Rp(R, p)
    B_X_W = adjoint_from_Rp(R, p, inverse=True)
    M_B = crba(model, s)
    # generalized bias forces (includes gravity and external forces) in body-fixed repr:
    # run RNEA with body-fixed base acceleration zero.  RNEA takes the inertial-fixed
    # apparent acceleration; a zero *body-fixed* acceleration B_vd = 0 corresponds to
    # W_vd = W_X_B @ 0 + vx(W_v) ... (see below) -- we instead call the RNEA internals via
    # the identity W_vd_WB = W_X_B a0 + W_g used by ABA (aba.py:284-288): RNEA converts with
    # a0 = B_X_W (W_vd - W_g), so passing W_vd = W_g yields a0 = 0 exactly.
    W_g = np.zeros((B, 6), dtype=dtype)
    W_g[:, 2] = model.gravity
    W_f0, tau_h = rnea(model, p, q, s, v_lin, omega, sd, W_g, np.zeros_like(s), W_f)
    # f[0] in body frame = B_X_W^-T ... rnea returns W_f0 = B_X_W^T f0  ->  f0 = W_X_B^T W_f0
    f0 = np.einsum("bji,bj->bi", W_X_B, W_f0)
    if model.floating_base:
        rhs = np.concatenate([-f0, tau - tau_h], axis=-1)
        acc = np.linalg.solve(M_B, rhs[..., None])[..., 0]
        a0 = acc[:, 0:6]
        sdd = acc[:, 6:]
        W_vd = np.einsum("bij,bj->bi", W_X_B, a0) + W_g
    else:
        sdd = np.linalg.solve(M_B[:, 6:, 6:], (tau - tau_h)[..., None])[..., 0] if n > 0 else np.zeros_like(s)
        W_vd = np.zeros((B, 6), dtype=dtype)
    return W_vd, sdd


# =============================================================================
# synthetic inputs: api/data.py:552-682 + api/joint.py:184-277 (NumPy Philox stream)
# =============================================================================


def random_joint_positions(model: OracleModel, rng: np.random.Generator, B: int, dtype=np.float64):
    jp = model.kin_dyn_parameters.joint_parameters
    jt = np.array(model.kin_dyn_parameters.joint_model.joint_types[1:])
    s_min = jp.position_limits_min.copy()
    s_max = jp.position_limits_max.copy()
    pi = np.pi
    with np.errstate(over="ignore"):
        full = np.logical_and(jt == 1, s_max - s_min >= 2 * pi)
    both = np.logical_and(full, np.logical_and(s_min <= -pi, s_max >= pi))
    s_min = np.where(both, -pi, s_min)
    both2 = np.logical_and(full, np.logical_and(s_min <= -pi, s_max >= pi))
    s_max = np.where(both2, pi, s_max)
    s_min = np.where(np.logical_and(full, s_max < pi), s_max - 2 * pi, s_min)
    s_max = np.where(np.logical_and(full, s_min > -pi), s_min + 2 * pi, s_max)
    # prismatic / unbounded non-revolute joints: keep a finite window
    s_min = np.where(np.isfinite(s_min) & (np.abs(s_min) < 1e30), s_min, -1.0)
    s_max = np.where(np.isfinite(s_max) & (np.abs(s_max) < 1e30), s_max, 1.0)
    return rng.uniform(s_min, s_max, size=(B, len(s_min))).astype(dtype)


def random_model_data(model: OracleModel, B: int, seed: int = 0, dtype=np.float64,
                      base_pos_bounds=((-1, -1, 0.5), (1, 1, 1.0)), in_contact: bool = False) -> OracleData:
    """``random_model_data`` (``api/data.py:552-682``): same distributions, NumPy Philox
    stream (JAX's threefry stream cannot be reproduced without JAX).  ``in_contact=True``
    is this repo's second distribution: the base is lowered so that the
    lowest collidable point penetrates the ground by up to 5 mm; ``in_contact="flat"`` also
    levels the base and zeroes the joints (+-1e-3) so that several points touch at once."""
    rng = np.random.Generator(np.random.Philox(seed))
    n = model.dofs()
    p = rng.uniform(np.array(base_pos_bounds[0], float), np.array(base_pos_bounds[1], float), size=(B, 3))
    rpy = rng.uniform(-np.pi, np.pi, size=(B, 3))
    s = random_joint_positions(model, rng, B) if n > 0 else np.zeros((B, 0))
    sd = rng.uniform(-1, 1, size=(B, n))
    if model.floating_base:
        v = rng.uniform(-1, 1, size=(B, 3))
        w = rng.uniform(-1, 1, size=(B, 3))
        q = _quat_from_euler_xyz_intrinsic(rpy)
    else:
        v = np.zeros((B, 3))
        w = np.zeros((B, 3))
        p = np.zeros((B, 3))
        q = np.tile(np.array([1.0, 0, 0, 0]), (B, 1))
    if in_contact and model.floating_base and len(model.kin_dyn_parameters.contact_parameters.body) > 0:
        # moderate attitude so that feet/corners point down, then drop to touch the ground
        rpy = rng.uniform(-0.3, 0.3, size=(B, 3))
        lo = 0.0
        if in_contact == "flat":
            # near the zero pose with a level base: whole faces / soles touch the ground, so
            # that several collidable points are active at once (rigid-contact parity cases)
            rpy = rng.uniform(-1e-3, 1e-3, size=(B, 3))
            jp = model.kin_dyn_parameters.joint_parameters
            s = np.clip(rng.uniform(-1e-3, 1e-3, size=(B, n)), jp.position_limits_min, jp.position_limits_max) if n > 0 else s
            lo = 0.002
        q = _quat_from_euler_xyz_intrinsic(rpy)
        d0 = data_replace(model, s, sd, q, v, w, p)
        W_p_C, _ = collidable_points_pos_vel(model, d0.link_transforms, d0.link_velocities)
        zmin = W_p_C[..., 2].min(axis=1)
        p = p.copy()
        p[:, 2] += model.terrain_height - zmin - rng.uniform(lo, 0.005, size=B)
        v = 0.1 * v
        w = 0.1 * w
        sd = 0.1 * sd
    cast = lambda a: np.ascontiguousarray(a, dtype=dtype)  # noqa: E731
    return data_replace(model, cast(s), cast(sd), cast(q), cast(v), cast(w), cast(p))


def _quat_from_euler_xyz_intrinsic(angles):
    """scipy ``Rotation.from_euler("XYZ", angles).as_quat()`` -> wxyz
    (``api/data.py:626-633``): intrinsic rotations R = Rx(a) Ry(b) Rz(c)."""
    a, b, c = angles[:, 0] / 2, angles[:, 1] / 2, angles[:, 2] / 2
    qx = np.stack([np.cos(a), np.sin(a), 0 * a, 0 * a], -1)
    qy = np.stack([np.cos(b), 0 * b, np.sin(b), 0 * b], -1)
    qz = np.stack([np.cos(c), 0 * c, 0 * c, np.sin(c)], -1)
    return _qmul(_qmul(qx, qy), qz)


def _qmul(a, b):
    aw, ax, ay, az = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bw, bx, by, bz = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    return np.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        -1,
    )


# =============================================================================
# api/ode.py:134-225 + api/integrators.py:91-156  (RK4)
# =============================================================================


def system_dynamics(model: OracleModel, data: OracleData, W_f_L_external, tau_total) -> dict:
    """``ode.system_dynamics`` (``api/ode.py:174-225``) in inertial-fixed representation,
    with ``system_position_dynamics`` (``:134-171``, Baumgarte K = 1.0)."""
    W_vd_WB, sdd, m_dot = system_acceleration(model, data, W_f_L_external, tau_total)
    W_w = data.base_angular_velocity
    W_pd_B = data.base_linear_velocity + np.einsum("bij,bj->bi", wedge(W_w), data.base_position)
    W_Qd_B = quaternion_derivative(data.base_orientation, W_w, K=1.0)
    return dict(
        base_position=W_pd_B, base_quaternion=W_Qd_B, joint_positions=data.joint_velocities,
        base_linear_velocity=W_vd_WB[:, 0:3], base_angular_velocity=W_vd_WB[:, 3:6], joint_velocities=sdd,
        tangential_deformation=m_dot,
    )


def rk4_integration(model: OracleModel, data: OracleData, W_f_L_external, tau_total) -> OracleData:
    """``rk4_integration`` (``api/integrators.py:91-156``)."""
    dtype = data.joint_positions.dtype
    dt = dtype.type(model.time_step)
    qn = safe_norm(data.base_quaternion, axis=-1)
    q0 = data.base_quaternion / np.where(qn == 0, 1.0, qn)[:, None]
    x0 = dict(
        base_position=data.base_position, base_quaternion=q0, joint_positions=data.joint_positions,
        base_linear_velocity=data.base_linear_velocity, base_angular_velocity=data.base_angular_velocity,
        joint_velocities=data.joint_velocities, tangential_deformation=data.tangential_deformation,
    )

    def f(x):
        d = data_replace(model, x["joint_positions"], x["joint_velocities"], x["base_quaternion"],
                         x["base_linear_velocity"], x["base_angular_velocity"], x["base_position"],
                         x["tangential_deformation"])
        return system_dynamics(model, d, W_f_L_external, tau_total)

    mid = lambda x, k: {n: x[n] + (0.5 * dt) * k[n] for n in x}  # noqa: E731
    fin = lambda x, k: {n: x[n] + dt * k[n] for n in x}  # noqa: E731
    k1 = f(x0)
    k2 = f(mid(x0, k1))
    k3 = f(mid(x0, k2))
    k4 = f(fin(x0, k3))
    dxdt = {n: (k1[n] + 2 * k2[n] + 2 * k3[n] + k4[n]) / 6 for n in x0}
    xf = fin(x0, dxdt)
    return data_replace(model, xf["joint_positions"], xf["joint_velocities"], xf["base_quaternion"],
                        xf["base_linear_velocity"], xf["base_angular_velocity"], xf["base_position"],
                        xf["tangential_deformation"])


def step_rk4(model: OracleModel, data: OracleData, link_forces_inertial=None, joint_force_references=None) -> OracleData:
    """``js.model.step`` with ``IntegratorType.RungeKutta4`` (``api/model.py:2601-2681``)."""
    B = data.joint_positions.shape[0]
    dtype = data.joint_positions.dtype
    nL, n = model.number_of_links(), model.dofs()
    W_f = np.zeros((B, nL, 6), dtype=dtype) if link_forces_inertial is None else np.asarray(link_forces_inertial, dtype=dtype)
    tau_ref = np.zeros((B, n), dtype=dtype) if joint_force_references is None else np.asarray(joint_force_references, dtype=dtype)
    tau_total = compute_resultant_torques(model, data.joint_positions, data.joint_velocities, tau_ref)
    return rk4_integration(model, data, W_f, tau_total)


# =============================================================================
# api/integrators.py:159-263  (RK4Fast)
# =============================================================================


def rk4fast_integration(model: OracleModel, data: OracleData, W_f_L_external, tau_total) -> OracleData:
    """``rk4fast_integration`` (``api/integrators.py:159-263``), literally, quirks included:

    * the contact forces are evaluated ONCE, at the initial state (``:175-188``), and held over the four stages;
    * the contact state that enters the integration is ``update_contact_state(contact_state_derivative)``, which for
      SoftContacts returns the DERIVATIVE m_dot under the key ``tangential_deformation`` (``soft.py:164-176``), and the
      stages return the stage's contact state itself as its "derivative" (``:218-219``);
    * the position derivatives come from ``system_position_dynamics(data=data)`` -- the ORIGINAL data, not the stage's
      (``:206-209``) -- so positions advance with the initial velocities.
    Only models with collidable points reach the end of the reference function (``W_f_L_terrain`` is undefined
    otherwise, ``:175-190``)."""
    dtype = data.joint_positions.dtype
    dt = dtype.type(model.time_step)
    nc = len(model.kin_dyn_parameters.contact_parameters.body)
    if nc == 0 or model.contact_model != "soft":
        raise NotImplementedError("rk4fast_integration: the reference only runs with collidable points; oracle: SoftContacts")
    W_f_C, m_dot = soft_compute_contact_forces(model, data)
    W_f_L_total = W_f_L_external + link_forces_from_contact_forces(model, W_f_C)
    W_w = data.base_angular_velocity
    W_pd_B = data.base_linear_velocity + np.einsum("bij,bj->bi", wedge(W_w), data.base_position)
    W_Qd_B = quaternion_derivative(data.base_orientation, W_w, K=1.0)
    sd0 = data.joint_velocities

    def f(x):
        d = data_replace(model, x["joint_positions"], x["joint_velocities"], x["base_quaternion"],
                         x["base_linear_velocity"], x["base_angular_velocity"], x["base_position"],
                         x["tangential_deformation"])
        W_vd_WB, sdd = aba(model, d.base_position, d.base_orientation, d.joint_positions, d.base_linear_velocity,
                           d.base_angular_velocity, d.joint_velocities, tau_total, W_f_L_total)
        return dict(base_position=W_pd_B, base_quaternion=W_Qd_B, joint_positions=sd0,
                    base_linear_velocity=W_vd_WB[:, 0:3], base_angular_velocity=W_vd_WB[:, 3:6], joint_velocities=sdd,
                    tangential_deformation=x["tangential_deformation"])

    qn = safe_norm(data.base_quaternion, axis=-1)
    q0 = data.base_quaternion / np.where(qn == 0, 1.0, qn)[:, None]
    x0 = dict(
        base_position=data.base_position, base_quaternion=q0, joint_positions=data.joint_positions,
        base_linear_velocity=data.base_linear_velocity, base_angular_velocity=data.base_angular_velocity,
        joint_velocities=data.joint_velocities, tangential_deformation=m_dot,
    )
    mid = lambda x, k: {n: x[n] + (0.5 * dt) * k[n] for n in x}  # noqa: E731
    fin = lambda x, k: {n: x[n] + dt * k[n] for n in x}  # noqa: E731
    k1 = f(x0)
    k2 = f(mid(x0, k1))
    k3 = f(mid(x0, k2))
    k4 = f(fin(x0, k3))
    dxdt = {n: (k1[n] + 2 * k2[n] + 2 * k3[n] + k4[n]) / 6 for n in x0}
    xf = fin(x0, dxdt)
    return data_replace(model, xf["joint_positions"], xf["joint_velocities"], xf["base_quaternion"],
                        xf["base_linear_velocity"], xf["base_angular_velocity"], xf["base_position"],
                        xf["tangential_deformation"])


def step_rk4fast(model: OracleModel, data: OracleData, link_forces_inertial=None, joint_force_references=None) -> OracleData:
    """``js.model.step`` with ``IntegratorType.RungeKutta4Fast`` (``api/model.py:2601-2681``)."""
    B = data.joint_positions.shape[0]
    dtype = data.joint_positions.dtype
    nL, n = model.number_of_links(), model.dofs()
    W_f = np.zeros((B, nL, 6), dtype=dtype) if link_forces_inertial is None else np.asarray(link_forces_inertial, dtype=dtype)
    tau_ref = np.zeros((B, n), dtype=dtype) if joint_force_references is None else np.asarray(joint_force_references, dtype=dtype)
    tau_total = compute_resultant_torques(model, data.joint_positions, data.joint_velocities, tau_ref)
    return rk4fast_integration(model, data, W_f, tau_total)
