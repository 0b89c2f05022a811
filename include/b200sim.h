/*
 * b200sim.h -- C ABI of the B200-native batched rigid-body simulation step.
 *
 * This is the drop-in boundary for ONE hot path of ami-iit/jaxsim: the vmapped
 * `jaxsim.api.model.step` (reference: src/jaxsim/api/model.py:2601-2681) and the functions
 * under it.  The reference has no FFI of its own (it is a jitted Python function), so these
 * entry points are what a maintainer would bind from Python (ctypes stub: INTEGRATION.md)
 * in place of `jax.jit(jax.vmap(step, in_axes=(None, 0)))`.
 *
 * Conventions (identical to the reference):
 *   - all batched arrays are row-major with a leading batch axis B, contiguous,
 *     element type float (dtype 0) or double (dtype 1); pointers are DEVICE pointers;
 *   - quaternions are wxyz; 6D vectors are [linear; angular];
 *   - base/link velocities and link forces are in INERTIAL-FIXED representation
 *     (api/data.py:36-39); forces are [f; moment about the world origin];
 *   - DoF j (0-based) drives link j+1; link 0 is the base (rbda/aba.py:133).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream); the caller owns all buffers; nothing is allocated per call.
 * Return value: 0 ok; <0 invalid argument (B200SIM_E_*); >0 a cudaError_t.
 * Thread safety: a B200SimModel is immutable after create (except the explicit update_*
 * calls) and may be used concurrently from several host threads / streams.
 */
#ifndef B200SIM_H
#define B200SIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SIM_ABI_VERSION 3

#define B200SIM_DTYPE_F32 0
#define B200SIM_DTYPE_F64 1

#define B200SIM_CONTACT_NONE 0
#define B200SIM_CONTACT_SOFT 1 /* rbda/contacts/soft.py */
#define B200SIM_CONTACT_RIGID 2 /* rbda/contacts/rigid.py (floating base; enabled points a prefix; b200sim_step_n with nsteps > 1 runs
                                    nsteps cascades on the stream and needs the W_H_L / W_v_WL outputs, which feed the next step) */
#define B200SIM_CONTACT_RELAXED_RIGID 3 /* rbda/contacts/relaxed_rigid.py (same restrictions as RIGID) */
/* A RIGID / RELAXED_RIGID step is a cascade of launches on the caller's stream (csrc/b200sim.cu: launch_rigid):
 * the fused step kernel for the environments without contact, then one warp per environment in contact.  RIGID
 * (default): assemble / rigid_qp_kernel / resume launches for up to 12 active points, the full-size level on a side
 * stream that forks from and joins the caller's stream through events -- the call stays stream-ordered and CUDA-graph
 * capturable.  Its work lists and contact-QP records live in a per-(model, stream) scratch block (5.9 KB per
 * environment of the batch for float32 data, 20 KB for float64, where the records also carry the assembling launch's
 * workspace; above 2 GiB the contact QP stays inside the rigid kernel): the first call on a stream (or with a larger
 * batch) must not be inside a stream capture (B200SIM_E_UNSUPPORTED). */

#define B200SIM_E_INVALID (-1)     /* NULL / negative size / bad dtype */
#define B200SIM_E_UNSUPPORTED (-2) /* valid in the reference, not implemented here */
#define B200SIM_E_TOO_LARGE (-3)   /* model does not fit the shared-memory workspace */

/* Host description of a model: a field-for-field image of the reference's
 * KinDynParameters (src/jaxsim/api/kin_dyn_parameters.py:21-63) + the Static fields of
 * JaxSimModel (src/jaxsim/api/model.py:46-90) the step reads.  All pointers are HOST
 * pointers to float64 / int32 arrays; they are copied by b200sim_model_create. */
typedef struct B200SimModelDesc {
  int32_t abi_version; /* = B200SIM_ABI_VERSION */
  int32_t n_links;     /* nL */
  int32_t n_dofs;      /* n  (= nL - 1) */
  int32_t n_points;    /* collidable points (enabled or not), contact_parameters.body */
  int32_t floating_base; /* joint_dofs[0] == 6 (api/model.py:722-730) */
  int32_t contact_model; /* B200SIM_CONTACT_* */
  int32_t enable_friction; /* ActuationParams.enable_friction (rbda/actuation/common.py:19) */
  int32_t reserved0;

  const int32_t *parent;     /* [nL]   parent_array, parent[0] = -1 */
  const int32_t *joint_type; /* [nL]   0 fixed (index 0 only), 1 revolute, 2 prismatic */
  const double *lam_H_pre;   /* [nL,4,4] JointModel.lam_H_pre (math/joint_model.py:16-44) */
  const double *suc_H_i;     /* [nL,4,4] JointModel.suc_H_i */
  const double *joint_axis;  /* [nL,3]  row 0 unused; axis of joint i in its own frame */
  const double *link_mass;   /* [nL]    LinkParameters.mass */
  const double *link_com;    /* [nL,3]  LinkParameters.center_of_mass (link frame) */
  const double *link_inertia;/* [nL,6]  LinkParameters.inertia_elements (triu, at the CoM) */
  /* JointParameters (api/kin_dyn_parameters.py:502-571), all [n] */
  const double *friction_static;
  const double *friction_viscous;
  const double *position_limits_min;
  const double *position_limits_max;
  const double *position_limit_spring;
  const double *position_limit_damper;
  /* ContactParameters (api/kin_dyn_parameters.py:765-840) */
  const int32_t *point_body;    /* [nc] */
  const double *point_position; /* [nc,3] L_p_C */
  const int32_t *point_enabled; /* [nc] 0/1 */

  double time_step;      /* JaxSimModel.time_step */
  double gravity;        /* JaxSimModel.gravity: z acceleration, NEGATIVE (-9.81) */
  double terrain_height; /* FlatTerrain._height (terrain/terrain.py:66-113) */
  /* SoftContactsParams (rbda/contacts/soft.py:24-46).  With B200SIM_CONTACT_RIGID, soft_K /
   * soft_D / soft_mu carry RigidContactsParams.K / D / mu (Baumgarte gains and friction
   * coefficient, rbda/contacts/rigid.py:27-42) and soft_p / soft_q are ignored. */
  double soft_K, soft_D, soft_mu, soft_p, soft_q;
  /* ActuationParams (rbda/actuation/common.py:10-19) */
  double torque_max, omega_th, omega_max;
  /* RigidContacts.regularization_delassus (rbda/contacts/rigid.py:99-101), default 1e-6 */
  double rigid_regularization;
  /* RelaxedRigidContactsParams (rbda/contacts/relaxed_rigid.py:30-82); the friction coefficient
   * travels in soft_mu.  Ignored unless contact_model == B200SIM_CONTACT_RELAXED_RIGID. */
  double relaxed_time_constant, relaxed_damping_coefficient, relaxed_d_min, relaxed_d_max, relaxed_width,
      relaxed_midpoint, relaxed_power;
} B200SimModelDesc;

typedef struct B200SimModel B200SimModel;

/* Upload a model to `device` (CUDA ordinal).  Replaces building the KinDynParameters
 * pytree leaves on device (api/model.py:224-330). */
int b200sim_model_create(const B200SimModelDesc *desc, int device, B200SimModel **out);
void b200sim_model_destroy(B200SimModel *model);

/* Replace the inertial parameters of every link (host float64 arrays, same layout as the
 * descriptor).  Counterpart of KinDynParameters.set_link_mass / update_hw_parameters
 * feeding new LinkParameters (api/kin_dyn_parameters.py:453-499). Synchronous. */
int b200sim_model_update_link_params(B200SimModel *model, const double *link_mass,
                                     const double *link_com, const double *link_inertia);

/* Tuning knobs (0 = automatic): lanes per environment (1,2,4,8,16,32) and environments per
 * thread block.  Only affects performance, never results. */
int b200sim_model_set_tuning(B200SimModel *model, int lanes_per_env, int envs_per_block);

/* Implementation options (bit mask).  Default: B200SIM_OPT_TMA_STORE.
 *   B200SIM_OPT_TMA_STORE: the (B,nL,6,6) joint-transform cache leaves shared memory through
 *   the TMA engine (cp.async.bulk) instead of 128-bit stores from registers (same results).
 *   B200SIM_OPT_RIGID_QP_F32: float32 rigid-contact steps also solve the contact QP / impact
 *   system in float32.  Default is float64 for those two solves (the Delassus regularisation
 *   1e-6 sits at float32 resolution): in float32 the contact forces are only good to ~1e-3
 *   relative -- the accuracy the reference itself asks of qpax (solver_tol=1e-3) -- but the
 *   workspace per environment is smaller and the step faster. */
#define B200SIM_OPT_TMA_STORE 1
#define B200SIM_OPT_RIGID_QP_F32 2
/*   B200SIM_OPT_GENERIC_KERNEL: always launch the generic step-kernel instance, also where
 *   the instance specialised for floating-base soft-contact steps applies (same results;
 *   diagnostic / A-B timing switch). */
#define B200SIM_OPT_GENERIC_KERNEL 4
/*   B200SIM_OPT_BULK_IN: read the cached kinematics of the input state with two cp.async.bulk
 *   (TMA, mbarrier completion) per environment instead of six cp.async per link (same
 *   results).  Off by default: measured neutral at batch 4096 / 65536 with inputs from HBM,
 *   ~1 us faster per step when the inputs are L2-resident, 5 % slower at batch 8192
 *   (profiles/r01_bulk_in_ab.log). */
#define B200SIM_OPT_BULK_IN 8
/*   B200SIM_OPT_NO_PDL: launch the specialised step kernel as an ordinary kernel instead of with
 *   programmatic stream serialization (the launch set-up of step k+1 then no longer overlaps
 *   the execution of step k; same results; diagnostic). */
#define B200SIM_OPT_NO_PDL 16
/*   B200SIM_OPT_STEP_V1: launch the first-generation specialised step kernel (60-word link records,
 *   parents gather their children in ABA pass 2) instead of the second-generation one (44-word records,
 *   children add into the parent; b200sim_step2.cuh).  Same results up to rounding; A-B timing switch. */
#define B200SIM_OPT_STEP_V1 32
/*   B200SIM_OPT_NO_BULK_IN: the second-generation kernel reads the cached kinematics of the input state
 *   with cp.async (LDGSTS) per link even when the rows qualify for cp.async.bulk (diagnostic). */
#define B200SIM_OPT_NO_BULK_IN 64
/*   B200SIM_OPT_RIGID_MONO: RigidContacts steps solve the contact QP inside the rigid kernel (one launch per
 *   cascade level) instead of the default assemble / solve / resume launches of level 1 (A/B timing). */
#define B200SIM_OPT_RIGID_MONO 128
int b200sim_model_set_options(B200SimModel *model, int32_t options);

/* Query sizes / launch geometry chosen for a batch (for benchmarks and tests). */
int b200sim_model_query(const B200SimModel *model, int dtype, int64_t B, int32_t *lanes_per_env,
                        int32_t *envs_per_block, int32_t *grid, int32_t *smem_bytes);

/* One simulation step for B environments == jax.vmap(js.model.step, in_axes=(None, 0))
 * with SoftContacts (or no collidable points) and SemiImplicitEuler:
 *   actuation model (api/actuation_model.py:7-126) -> collidable-point kinematics and
 *   Hunt/Crossley forces (rbda/collidable_points.py, rbda/contacts/soft.py:195-444) ->
 *   ABA (rbda/aba.py:12-292) -> semi-implicit Euler (api/integrators.py:14-88) ->
 *   cache refresh = joint transforms + FK (api/data.py:406-523, rbda/forward_kinematics.py).
 *
 * in : s, sd (B,n); q_wxyz (B,4); v_lin, omega, p (B,3); m_tan (B,nc,3) or NULL (== zeros);
 *      tau_ref (B,n) joint_force_references or NULL; f_ext (B,nL,6) inertial-fixed link
 *      forces or NULL.  The link transforms/velocities the reference reads from the input
 *      data's caches are recomputed from the state (they are functions of it).
 * out: *_o new state leaves (may alias the inputs); m_tan_o (B,nc,3) or NULL;
 *      caches (each may be NULL to skip): W_H_B (B,4,4) _base_transform,
 *      i_X_lam (B,nL,6,6) _joint_transforms, W_H_L (B,nL,4,4) _link_transforms,
 *      W_v_WL (B,nL,6) _link_velocities.
 */
int b200sim_step(const B200SimModel *model, int dtype, int64_t B,
                 const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                 const void *omega, const void *p, const void *m_tan,
                 const void *tau_ref, const void *f_ext,
                 void *s_o, void *sd_o, void *q_o, void *v_lin_o, void *omega_o, void *p_o,
                 void *m_tan_o,
                 void *W_H_B, void *i_X_lam, void *W_H_L, void *W_v_WL,
                 void *stream);

/* `nsteps` consecutive steps in ONE launch (the caller's `for _ in range(T): data =
 * step(model, data, ...)` loop, README.md:80-84): the state stays on chip between steps,
 * only the joint force references are read per step.  tau_ref is (nsteps,B,n) with
 * `tau_step_stride` elements between consecutive steps (0: the same (B,n) block every
 * step); likewise f_ext with `fext_step_stride`.  Outputs are those of the LAST step.
 * W_H_L_in (B,nL,4,4) / W_v_WL_in (B,nL,6): the cached link transforms and velocities OF THE
 * INPUT STATE (JaxSimModelData._link_transforms/_link_velocities, which the reference's
 * contact code reads, api/contact.py:39-43), both or neither (NULL): given, the kernel skips
 * the joint transforms + FK of the input state.  They must be consistent with the state.
 * b200sim_step == b200sim_step_n with nsteps = 1 and no input caches.  All cache pointers
 * must be 16-byte aligned. */
int b200sim_step_n(const B200SimModel *model, int dtype, int64_t B, int32_t nsteps,
                   const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                   const void *omega, const void *p, const void *m_tan,
                   const void *tau_ref, int64_t tau_step_stride,
                   const void *f_ext, int64_t fext_step_stride,
                   const void *W_H_L_in, const void *W_v_WL_in,
                   void *s_o, void *sd_o, void *q_o, void *v_lin_o, void *omega_o, void *p_o,
                   void *m_tan_o,
                   void *W_H_B, void *i_X_lam, void *W_H_L, void *W_v_WL,
                   void *stream);

/* Per-environment status flags.  The reference can only raise these conditions as host exceptions, and only when
 * JAXSIM_ENABLE_EXCEPTIONS is set (rbda/utils.py:136-146, exceptions.py:43-48); qpax's convergence flag is dropped
 * (rbda/contacts/rigid.py:359-362).  Here they are evaluated on the device:
 *   QUATERNION_NAN       the stored base quaternion of the INPUT state contains a NaN;
 *   QUATERNION_NOT_UNIT  its squared norm differs from 1 beyond jnp.allclose's defaults (rtol 1e-5, atol 1e-8);
 *   NON_FINITE           a state leaf of the OUTPUT (s, sd, q, v_lin, omega, p) is NaN or infinite;
 *   QP_NOT_CONVERGED     the contact QP of a RigidContacts step left its iteration without meeting the tolerance
 *                        (the best iterate was used). */
#define B200SIM_STATUS_QUATERNION_NAN 1
#define B200SIM_STATUS_QUATERNION_NOT_UNIT 2
#define B200SIM_STATUS_NON_FINITE 4
#define B200SIM_STATUS_QP_NOT_CONVERGED 8
/* b200sim_step_n plus `status_flags`: (B,) int32 on the device, overwritten with the flags of every environment
 * (0 = nothing to report).  One small extra launch on the same stream; NULL falls back to b200sim_step_n.  The
 * output quaternion buffer must not alias the input one. */
int b200sim_step_n_status(const B200SimModel *model, int dtype, int64_t B, int32_t nsteps, const void *s,
                          const void *sd, const void *q_wxyz, const void *v_lin, const void *omega, const void *p,
                          const void *m_tan, const void *tau, int64_t tau_step_stride, const void *f_ext_inertial,
                          int64_t f_ext_step_stride, const void *W_H_L_in, const void *W_v_in, void *s_o, void *sd_o,
                          void *q_o, void *v_lin_o, void *omega_o, void *p_o, void *m_tan_o, void *W_H_B, void *i_X_lam,
                          void *W_H_L, void *W_v_WL, int32_t *status_flags, void *stream);

/* b200sim_step_n with both extras.  `f_ext_representation`: how `f_ext_inertial` is expressed -- the reference's
 * `link_forces` argument is given in `data.velocity_representation` (api/model.py:2617-2618) and re-expressed with the
 * link transforms of EVERY step (api/model.py:2641-2646, api/common.py:160-222):
 *   B200SIM_REPR_INERTIAL  inertial-fixed (what b200sim_step / b200sim_step_n expect);
 *   B200SIM_REPR_BODY      body-fixed: frame of the link;
 *   B200SIM_REPR_MIXED     mixed: origin of the link, world axes.
 * The conversion happens inside the kernels with the link poses of the current step, so a fused rollout applies the
 * forces exactly like repeated `step` calls.  B200SIM_E_UNSUPPORTED for body-fixed / mixed forces on models whose link
 * poses differ from the ABA chain (fixed base with an offset mount, SDF-posed base link). */
#define B200SIM_REPR_INERTIAL 0
#define B200SIM_REPR_BODY 1
#define B200SIM_REPR_MIXED 2
int b200sim_step_n_ex(const B200SimModel *model, int dtype, int64_t B, int32_t nsteps, const void *s, const void *sd,
                      const void *q_wxyz, const void *v_lin, const void *omega, const void *p, const void *m_tan,
                      const void *tau, int64_t tau_step_stride, const void *f_ext, int64_t f_ext_step_stride,
                      const void *W_H_L_in, const void *W_v_in, void *s_o, void *sd_o, void *q_o, void *v_lin_o,
                      void *omega_o, void *p_o, void *m_tan_o, void *W_H_B, void *i_X_lam, void *W_H_L, void *W_v_WL,
                      int32_t f_ext_representation, int32_t *status_flags, void *stream);

/* Cache computation only (JaxSimModelData.build / .replace, api/data.py:66-202,406-523):
 * normalises q (written to q_o if not NULL) and fills the requested caches. */
int b200sim_fk(const B200SimModel *model, int dtype, int64_t B,
               const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
               const void *omega, const void *p, void *q_o,
               void *W_H_B, void *i_X_lam, void *W_H_L, void *W_v_WL, void *stream);

/* Forward dynamics == vmapped rbda.aba (rbda/aba.py:12-292) with the given joint forces tau
 * (B,n) or NULL and inertial-fixed link forces f_ext (B,nL,6) or NULL (no actuation model,
 * no contacts).  out: W_vd_WB (B,6) inertial-fixed base acceleration, sdd (B,n). */
int b200sim_aba(const B200SimModel *model, int dtype, int64_t B,
                const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                const void *omega, const void *p, const void *tau, const void *f_ext,
                void *W_vd_WB, void *sdd, void *stream);

/* State derivative == vmapped ode.system_dynamics (api/ode.py:174-225) in inertial-fixed
 * representation: soft-contact forces of the given state, ABA with the given RESULTANT joint
 * torques tau (B,n) or NULL (no actuation model here, like the reference's integrators
 * receive joint_torques), inertial-fixed link forces f_ext or NULL.
 * out: pd (B,3) base position derivative, qd (B,4) quaternion derivative (Baumgarte K = 1,
 * api/ode.py:134-171), W_vd (B,6) base acceleration, sdd (B,n), md (B,nc,3) derivative of
 * the tangential deformation (any of pd, qd, md may be NULL).  Building block of the
 * RK4 integrators (api/integrators.py:91-263). */
int b200sim_dynamics(const B200SimModel *model, int dtype, int64_t B,
                     const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                     const void *omega, const void *p, const void *m_tan, const void *tau,
                     const void *f_ext, void *pd, void *qd, void *W_vd, void *sdd, void *md,
                     void *stream);

/* One RungeKutta4 step == vmapped jaxsim.api.model.step with IntegratorType.RungeKutta4
 * (api/model.py:2601-2681 -> api/integrators.py:91-156): the actuation model on the input state
 * (api/actuation_model.py:7-126; tau_ref (B,n) or NULL), four evaluations of ode.system_dynamics
 * (api/ode.py:174-225; contact state integrated with the rest of the state) on x0, x0 + dt/2 k1,
 * x0 + dt/2 k2, x0 + dt k3, the (1,2,2,1)/6 combination, then data.replace (api/data.py:441-447:
 * quaternion normalised, caches of the new state).  Five launches of the dynamics / FK kernel and
 * five elementwise ones, all on `stream`, no host round trip; the stage buffers live in a per-
 * (model, stream) scratch block, so the first call on a stream (or with a larger batch) must not be
 * inside a stream capture.  f_ext (B,nL,6) inertial-fixed or NULL, held over the stages like the
 * reference does.  Outputs as b200sim_step (the four cache pointers may be NULL); out may alias
 * in.  RigidContacts / RelaxedRigidContacts with collidable points: B200SIM_E_UNSUPPORTED. */
int b200sim_step_rk4(B200SimModel *model, int dtype, int64_t B,
                     const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                     const void *omega, const void *p, const void *m_tan, const void *tau_ref,
                     const void *f_ext,
                     void *s_o, void *sd_o, void *q_o, void *v_lin_o, void *omega_o, void *p_o,
                     void *m_o, void *W_H_B, void *joint_X, void *W_H_L, void *W_v_WL,
                     void *stream);

/* Inverse dynamics == vmapped rbda.rnea (rbda/rnea.py:12-238).  in: state as above,
 * W_vd_WB (B,6) inertial-fixed base acceleration or NULL (zeros), sdd (B,n) or NULL,
 * f_ext (B,nL,6) inertial-fixed link forces or NULL.  out: W_f_B (B,6) the inertial-fixed
 * 6D force on the base, tau (B,n). */
int b200sim_rnea(const B200SimModel *model, int dtype, int64_t B,
                 const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                 const void *omega, const void *p, const void *W_vd_WB, const void *sdd,
                 const void *f_ext, void *W_f_B, void *tau, void *stream);

/* Free-floating mass matrix == vmapped rbda.crba (rbda/crba.py:10-170), body-fixed
 * representation.  in: s (B,n).  out: M (B,6+n,6+n) (fully written). */
int b200sim_crba(const B200SimModel *model, int dtype, int64_t B, const void *s, void *M,
                 void *stream);

/* Forward-mode derivative of `nsteps` steps (BASELINE config 5: d(step)/d(link masses,
 * joint positions, ...)): == jax.jvp(step) of the reference
 * (tests/test_automatic_differentiation.py:346-420).  float64 only.  Every batched array is
 * the primal array with one extra trailing axis of size 2: [..., 0] = value, [..., 1] =
 * tangent (e.g. s is (B,n,2)).  link_mass_tangent: HOST (nL) direction in the space of
 * LinkParameters.mass (api/kin_dyn_parameters.py:596) or NULL; it enters through
 * Inertia.to_sixd (math/inertia.py:32-39) with the CoM and the CoM inertia held fixed.
 * Outputs: value and tangent of every output leaf of b200sim_step (same NULL rules: with the four
 * cache pointers NULL the kinematics of the new state are neither computed nor stored).
 * Several directions of per-environment inputs = one launch over replicas of the batch.
 * The (value, tangent) image of the model constants is rewritten -- ordered on `stream`, no host
 * synchronisation -- only when a mass direction is given or the previous call had one; JVPs of
 * one model belong on one stream.  Not thread-safe per model. */
int b200sim_step_jvp(B200SimModel *model, int64_t B, int32_t nsteps,
                     const double *link_mass_tangent,
                     const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                     const void *omega, const void *p, const void *m_tan, const void *tau_ref,
                     void *s_o, void *sd_o, void *q_o, void *v_lin_o, void *omega_o, void *p_o,
                     void *m_tan_o,
                     void *W_H_B, void *i_X_lam, void *W_H_L, void *W_v_WL, void *stream);

/* Library / build information: "b200sim <abi> sm_100a ..." */
/* b200sim_step_jvp with SEVERAL mass directions in one launch: with mass_direction_period = P > 0 the batch is read as
 * B / P replicas of P environments and replica r differentiates w.r.t. link_mass_tangent[k] * (mass of link k),
 * k = mass_direction_first_link + r, alone (the entries of link_mass_tangent scale the directions; pass ones for unit
 * columns).  Replicas whose link index falls outside [0, nL) carry no mass direction.  P = 0: b200sim_step_jvp. */
int b200sim_step_jvp_ex(B200SimModel *model, int64_t B, int32_t nsteps,
                        const double *link_mass_tangent, int64_t mass_direction_period,
                        int32_t mass_direction_first_link,
                        const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                        const void *omega, const void *p, const void *m_tan, const void *tau,
                        void *s_o, void *sd_o, void *q_o, void *v_lin_o, void *omega_o, void *p_o,
                        void *m_o, void *W_H_B, void *joint_X, void *W_H_L, void *W_v_WL,
                        void *stream);

/* Gradient of the scalar <cotangent, step(x, theta)> with respect to the joint positions (per environment) and the link
 * masses (per environment; sum over the batch for a shared-parameter loss): what jax.grad / jax.vjp of the reference's
 * step returns for those arguments (tests/test_automatic_differentiation.py:346-420).  float64.  in: the primal state and
 * joint force references as for b200sim_step (NULL rules alike), the cotangents ct_* of the new state leaves in the
 * shapes of the leaves (any may be NULL = zero).  out: grad_s (B,n) and / or grad_link_mass (B,nL); either may be NULL.
 * Implementation: forward-mode columns contracted on the device -- the n joint directions run as replicas of the batch
 * in one launch of the forward-mode step kernel, each mass direction as one launch -- all on `stream`; the dual buffers
 * live in a per-(model, stream) scratch block (first call / larger batch: not inside a stream capture).  Cost: about
 * (n + nL) forward-mode steps.  RigidContacts / RelaxedRigidContacts with collidable points: B200SIM_E_UNSUPPORTED. */
int b200sim_step_vjp(B200SimModel *model, int64_t B,
                     const void *s, const void *sd, const void *q_wxyz, const void *v_lin,
                     const void *omega, const void *p, const void *m_tan, const void *tau,
                     const void *ct_s, const void *ct_sd, const void *ct_q, const void *ct_v_lin,
                     const void *ct_omega, const void *ct_p, const void *ct_m_tan,
                     void *grad_s, void *grad_link_mass, void *stream);

const char *b200sim_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SIM_H */
