// b200sim_kernels.cuh -- fused batched rigid-body step kernel for sm_100a.
//
// One launch performs, for every environment, everything `jaxsim.api.model.step` does
// (reference call stack: SURVEY.md section 3.1), optionally for several consecutive steps:
//   joint transforms -> FK -> [ collidable points + Hunt/Crossley -> actuation -> ABA passes
//   1/2/3 -> semi-implicit Euler -> joint transforms + FK of the new state ] x nsteps -> caches.
//
// Mapping:  G lanes of one warp cooperate on one environment (G in
// {1,2,4,8,16,32}); the per-link state of the environment lives in a shared-memory record
// (REC words per link) for the whole launch; phases that are independent per link / per
// collidable point / per DoF stride over them with the G lanes; the sequential tree
// recursions advance one tree LEVEL at a time with the lanes spread over the links of the
// level, separated by __syncwarp().  Inputs are pulled from HBM with one burst of
// cp.async (LDGSTS) per environment; the big cache output (the 6x6 joint adjoints) leaves
// through the TMA engine (cp.async.bulk shared->global), the rest through 128-bit stores.
//
// Formulation: the reference runs ABA in body-fixed link coordinates with dense 6x6
// adjoints (rbda/aba.py).  Here every link's spatial quantities are expressed in the frame
// F_i = (origin of link i, WORLD axes).  Plucker transforms between neighbouring links
// are then pure translations by r_i = p_i - p_parent, so the 6x6 congruence
// X^T M X of pass 2 (two dense 6x6x6 products in the reference, rbda/aba.py:205-213)
// collapses to ~70 flops, and the link inertias are rotated to world axes once per step
// outside the sequential chain.  The equations are the same ABA (a change of coordinates
// per link), so joint accelerations are identical up to rounding; the base acceleration is
// converted back to the reference's inertial-fixed representation at the end.
//
// Articulated inertia layout: I = [[A, B],[B^T, D]] acting on [lin; ang];
// A, D symmetric stored as 6 (xx,xy,xz,yy,yz,zz), B full row-major 9.
#pragma once

#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200sim_dual.cuh"

namespace b200sim {

// ------------------------------------------------------------------------------------
// layouts
// ------------------------------------------------------------------------------------
constexpr int REC = 60;   // workspace words per link (60 = 4*15: conflict-free 128-bit rows)
constexpr int CREC = 60;  // model-constant words per link (60 = 4*15)

// workspace record.  Words 0..26 hold (R,p) while kinematics are needed and the
// articulated inertia + bias force during ABA; words 12..47 double as the staging area of
// the link's 6x6 joint adjoint when the caches are written.
constexpr int O_R = 0;      // 9  world rotation of the link (ABA chain)
constexpr int O_P = 9;      // 3  world position of the link origin
constexpr int O_IA = 0;     // 6  A (sym)
constexpr int O_IB = 6;     // 9  B
constexpr int O_ID = 15;    // 6  D (sym)
constexpr int O_PA = 21;    // 6  articulated bias force
constexpr int O_C = 28;     // 6  velocity-product term c_i
constexpr int O_U = 34;     // 6  U_i
constexpr int O_DINV = 40;  // 1  1/d_i
constexpr int O_UU = 41;    // 1  u_i
constexpr int O_AX = 42;    // 3  joint axis in world axes
constexpr int O_RR = 45;    // 3  r_i = p_i - p_parent (world axes)
constexpr int O_X = 12;     // 36 staging of i_X_lambda(i) (output phase only)
constexpr int O_V = 48;     // 6  spatial velocity in F_i, later spatial acceleration
constexpr int O_SD = 54;    // 1  joint velocity
constexpr int O_S = 55;     // 1  joint position
constexpr int O_TAU = 56;   // 1  resultant joint torque
constexpr int O_SDD = 57;   // 1  joint acceleration
constexpr int O_TREF = 58;  // 1  joint force reference of the current step
// Compact layout of the FINAL phase of the last step (joint transforms + FK of the new state +
// cache outputs), used by the specialised instance when nL <= 4*G + 1.  The ABA scratch is dead
// by then, so the environment's workspace is re-interpreted as
//   [ nL x FS words: R(9) p(3) v(6) sd(1) pad ]  [ nL x 36 words: the (nL,6,6) joint adjoints ]
// The second block is contiguous exactly like the environment's slice of the (B,nL,6,6) output,
// so it leaves with ONE cp.async.bulk per environment instead of one per link (UBLKCP takes
// uniform-register operands: per-lane bulk copies are serialised lane by lane).  FS = 20 and 36
// keep 128-bit accesses of consecutive links on distinct banks.
constexpr int FS = 20;
constexpr int F_R = 0, F_P = 9, F_V = 12, F_SD = 18;
// per collidable point: contact force (3), lever arm (3), tangential deformation (3)
constexpr int PTREC = 9;
constexpr int PT_F = 0, PT_LEV = 3, PT_M = 6;

// model constant record: R_rel(s) = M0 + cos(s) M1 + sin(s) M2, t_rel(s) = TPRE + s RA
// (revolute: M0 = Rpre a a^T, M1 = Rpre - M0, M2 = Rpre S(a), RA = 0;
//  prismatic: M0 = Rpre, M1 = M2 = 0, RA = Rpre a)
constexpr int C_M0 = 0;     // 9
constexpr int C_M1 = 9;     // 9
constexpr int C_M2 = 18;    // 9
constexpr int C_TPRE = 27;  // 3
constexpr int C_RA = 30;    // 3
constexpr int C_AXIS = 33;  // 3 joint axis in link coordinates (motion subspace)
constexpr int C_MASS = 36;  // 1
constexpr int C_COM = 37;   // 3
constexpr int C_DL = 40;    // 6 I_c + m S(c)S(c)^T in link axes (sym)
constexpr int C_KS = 46;    // position_limit_spring
constexpr int C_KD = 47;    // position_limit_damper
constexpr int C_SMIN = 48;
constexpr int C_SMAX = 49;
constexpr int C_KC = 50;    // friction_static
constexpr int C_KV = 51;    // friction_viscous
constexpr int C_PAX = 52;   // 3 joint axis in PARENT-link coordinates: R_rel(s) a = R_pre a for every s

enum Mode : int { MODE_STEP = 0, MODE_FK = 1, MODE_ABA = 2, MODE_DYN = 3 };
enum Flags : int {
  F_SUC_NONID = 1,   // some suc_H_i[i>=1] is not the identity
  F_GENERIC_FK = 2,  // suc_H_i[0] != I or fixed base: FK poses differ from the ABA chain
  F_SQRT_P = 4,      // soft_p == 0.5
  F_SQRT_Q = 8,      // soft_q == 0.5
  F_TMA_STORE = 16,  // joint adjoints leave through cp.async.bulk (TMA) instead of STG.128
  F_BULK_IN = 32,    // the cached input kinematics may come in with cp.async.bulk (set per launch by the host
                     // when the (B,nL,4,4) / (B,nL,6) rows of every environment are 16-byte aligned and sized)
};

template <typename T>
struct Params {
  // ---- model (device pointers)
  const T* cst;      // [nL*CREC]
  const T* csuc;     // [nL*12]  suc_H_i (R row-major 9, t 3)
  const T* pt_pos;   // [nc*3]
  const int* itab;   // packed int tables, see offsets below
  int itab_words;
  int nL, n, nc, depth;
  int floating, contact_model, enable_friction, flags;
  int o_parent, o_jtype, o_lvl_start, o_lvl_links, o_child_start, o_child_idx, o_pt_start, o_pt_idx,
      o_pt_body, o_pt_enabled;
  int o_anc, o_ldepth;  // rigid contacts only: ancestors of every link (root-first, [nL*depth]) and their count
  int o_rows8, n_rows8, o_rows16, n_rows16;  // packed level-walk rows (specialised step kernel), see b200sim_model_create
  int o_rows2_8, o_rows2_16;                 // same rows in the format of step2_kernel (sibling ranks instead of child ranges)
  int ws2_words;                             // step2_kernel: shared-memory words per environment (env2_ws_words)
  int* status;                               // optional per-environment status flags (b200sim_step_n_status), OR-ed into
  int fext_repr;                             // representation of `fext`: 0 inertial-fixed, 1 body-fixed, 2 mixed (api/common.py:160-222)
  // forward mode only: with mass_dir_period > 0 environment e differentiates w.r.t. the mass of link
  // mass_dir_first + e / mass_dir_period -- the tangent part of the model constants then holds the unit mass direction
  // of EVERY link and is kept for that link alone (many mass directions in one launch over replicas of the batch)
  long long mass_dir_period;
  int mass_dir_first;
  T dt, g, h_terrain, K, D, mu, pexp, qexp, tau_max, w_th, w_max;
  T reg;                // rigid contacts: Delassus regularisation
  T rx_tc, rx_zeta, rx_dmin, rx_dmax, rx_width, rx_mid, rx_pow;  // relaxed-rigid contacts (relaxed_rigid.py:30-82)
  // ---- batch
  long long B;
  const T *s, *sd, *q, *vlin, *omega, *p, *m, *tau, *fext;
  const T *Hin, *Vin;  // optional cached kinematics of the INPUT state: W_H_L (B,nL,4,4), W_v_WL (B,nL,6)
  T *s_o, *sd_o, *q_o, *vlin_o, *omega_o, *p_o, *m_o;
  T *W_H_B, *iXl, *W_H_L, *W_v;
  T *avd, *sdd_o;
  long long tau_step_stride;   // elements between the torque rows of consecutive steps (0: constant)
  long long fext_step_stride;  // same for the external link forces
  int nsteps;
  int mode;
  int envs_per_block;
  // ---- rigid contacts: work lists of the cascade (b200sim_rigid_kernels.cuh).  An item is
  // env | (impact_only << 31).  `work_*` is consumed, `over_*` is produced.
  const int* work_count;
  const int* work_list;
  int* over_count;
  int* over_list;
  int na_cap;  // active points the consumer's shared-memory workspace is sized for
  // split rigid level (b200sim_rigid_kernels.cuh): 0 = the rigid kernel solves its contact QP itself; 1 = it assembles the
  // QP of every work item into qp_buf and stops; 2 = it resumes with the solution rigid_qp_kernel left there.
  // Record of work item k at qp_buf + k * qp_stride: int na, rc, env, pad | S q[3 cap] | S x[3 cap] | S Q[packed, 3 cap]
  int qp_mode;
  unsigned char* qp_buf;
  long long qp_stride;
  unsigned long long* dbg;  // optional counters (b200sim_debug_counters): QP iterations, items per level, ...
};

// ------------------------------------------------------------------------------------
// small math
// ------------------------------------------------------------------------------------
template <typename T> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float eps() { return 1.1920928955078125e-07f; } };
template <> struct Lim<double> { static __device__ __forceinline__ double eps() { return 2.220446049250313e-16; } };
template <> struct Lim<DualD> { static __device__ __forceinline__ DualD eps() { return DualD(2.220446049250313e-16); } };

__device__ __forceinline__ void sincos_t(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_t(double x, double* s, double* c) { sincos(x, s, c); }
// float: MUFU-based sqrt / reciprocal (about 1 ulp, far inside the 1e-3 float32 tolerance) instead of the IEEE
// sequences (each ~15-20 dependent instructions on the latency-bound path); double stays IEEE
__device__ __forceinline__ float sqrt_t(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float pow_t(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double pow_t(double x, double y) { return pow(x, y); }
// forward-mode scalars drop their tangent part when `keep` is false (b200sim_dual.cuh); plain scalars are unchanged
__device__ __forceinline__ float keep_tangent(float x, bool) { return x; }
__device__ __forceinline__ double keep_tangent(double x, bool) { return x; }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_t(double x) { return fabs(x); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
// reciprocal of the joint-space inertia d_i > 0: MUFU.RCP (<= 1 ulp) for float, IEEE for double
__device__ __forceinline__ float rcp_t(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ double rcp_t(double x) { return 1.0 / x; }
__device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }

template <typename T>
__device__ __forceinline__ void cross3(const T* a, const T* b, T* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T>
__device__ __forceinline__ void cross3_add(const T* a, const T* b, T* o) {
  o[0] += a[1] * b[2] - a[2] * b[1];
  o[1] += a[2] * b[0] - a[0] * b[2];
  o[2] += a[0] * b[1] - a[1] * b[0];
}
template <typename T>
__device__ __forceinline__ T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// o = R v  (R row-major 3x3)
template <typename T>
__device__ __forceinline__ void mat3_vec(const T* R, const T* v, T* o) {
  o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
// o = R^T v
template <typename T>
__device__ __forceinline__ void mat3T_vec(const T* R, const T* v, T* o) {
  o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
// C = A B
template <typename T>
__device__ __forceinline__ void mat3_mul(const T* A, const T* B, T* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// sym6 (xx,xy,xz,yy,yz,zz) times vector
template <typename T>
__device__ __forceinline__ void sym3_vec(const T* S, const T* v, T* o) {
  o[0] = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
  o[1] = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
  o[2] = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
}
template <int N, typename T>
__device__ __forceinline__ void ldn(const T* src, T* dst) {
#pragma unroll
  for (int k = 0; k < N; ++k) dst[k] = src[k];
}
template <int N, typename T>
__device__ __forceinline__ void stn(T* dst, const T* src) {
#pragma unroll
  for (int k = 0; k < N; ++k) dst[k] = src[k];
}

// 16-byte vector stores to global memory (dst must be 16-byte aligned; N*sizeof(T) % 16 == 0)
template <int N>
__device__ __forceinline__ void stg_vec(float* dst, const float* src) {
  static_assert(N % 4 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(src[k], src[k + 1], src[k + 2], src[k + 3]);
}
template <int N>
__device__ __forceinline__ void stg_vec(double* dst, const double* src) {
  static_assert(N % 2 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(dst + k) = make_double2(src[k], src[k + 1]);
}
template <int N>
__device__ __forceinline__ void stg_vec(DualD* dst, const DualD* src) {
#pragma unroll
  for (int k = 0; k < N; ++k) *reinterpret_cast<double2*>(dst + k) = make_double2(src[k].v, src[k].d);
}
// 6-vector rows: (B,nL,6) rows are 24 B (float, 8-byte aligned) / 48 B (double, 16-byte aligned)
__device__ __forceinline__ void stg_vec6(float* dst, const float* src) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) *reinterpret_cast<float2*>(dst + k) = make_float2(src[k], src[k + 1]);
}
__device__ __forceinline__ void stg_vec6(double* dst, const double* src) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) *reinterpret_cast<double2*>(dst + k) = make_double2(src[k], src[k + 1]);
}

__device__ __forceinline__ void stg_vec6(DualD* dst, const DualD* src) { stg_vec<6>(dst, src); }

// 16-byte vector loads from global memory (src 16-byte aligned; 8-byte for float 6-rows)
template <int N>
__device__ __forceinline__ void ldg_vec(const float* src, float* dst) {
  static_assert(N % 4 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + k);
    dst[k] = v.x; dst[k + 1] = v.y; dst[k + 2] = v.z; dst[k + 3] = v.w;
  }
}
template <int N>
__device__ __forceinline__ void ldg_vec(const double* src, double* dst) {
  static_assert(N % 2 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 2) {
    const double2 v = *reinterpret_cast<const double2*>(src + k);
    dst[k] = v.x; dst[k + 1] = v.y;
  }
}
template <int N>
__device__ __forceinline__ void ldg_vec(const DualD* src, DualD* dst) {
#pragma unroll
  for (int k = 0; k < N; ++k) dst[k] = src[k];
}
__device__ __forceinline__ void ldg_vec6(const float* src, float* dst) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) {
    const float2 v = *reinterpret_cast<const float2*>(src + k);
    dst[k] = v.x; dst[k + 1] = v.y;
  }
}
__device__ __forceinline__ void ldg_vec6(const double* src, double* dst) { ldg_vec<6>(src, dst); }
__device__ __forceinline__ void ldg_vec6(const DualD* src, DualD* dst) { ldg_vec<6>(src, dst); }

// ---- async copies -------------------------------------------------------------------
// one element global -> shared without a register round trip (LDGSTS)
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
  __pipeline_memcpy_async(smem_dst, gsrc, sizeof(T));
}

// cooperative global -> shared copy of `words` elements (a multiple of 16 bytes, both
// sides 16-byte aligned) with 16-byte LDGSTS chunks
template <typename E>
__device__ __forceinline__ void stage_async(E* smem_dst, const E* gsrc, int words) {
  constexpr int per = 16 / sizeof(E);
  const int chunks = words / per;
  for (int k = threadIdx.x; k < chunks; k += blockDim.x)
    __pipeline_memcpy_async(smem_dst + (size_t)k * per, gsrc + (size_t)k * per, 16);
}
// TMA bulk store shared -> global (cp.async.bulk, SASS UBLKCP): bytes % 16 == 0, both
// addresses 16-byte aligned.  The generic-proxy writes to shared memory must be fenced
// before the async proxy reads them.
__device__ __forceinline__ void tma_store_bulk(void* gdst, const void* smem_src, unsigned bytes) {
  const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(smem_src));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMA bulk loads global -> shared, completion on an mbarrier (cp.async.bulk + complete_tx)
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "B200SIM_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra B200SIM_MBAR_DONE;\n"
      "bra B200SIM_MBAR_WAIT;\n"
      "B200SIM_MBAR_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_bulk(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc),
               "r"(bytes), "r"(a) : "memory");
}

// jaxlie SO3(wxyz).as_matrix() (reference call sites rbda/aba.py:79-86)
template <typename T>
__device__ __forceinline__ void quat_to_dcm(const T* q, T* R) {
  const T nsq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const T k = T(2) * rcp_t(nsq);
  const T xx = q[1] * q[1] * k, yy = q[2] * q[2] * k, zz = q[3] * q[3] * k;
  const T xy = q[1] * q[2] * k, xz = q[1] * q[3] * k, yz = q[2] * q[3] * k;
  const T wx = q[0] * q[1] * k, wy = q[0] * q[2] * k, wz = q[0] * q[3] * k;
  R[0] = T(1) - yy - zz; R[1] = xy - wz;        R[2] = xz + wy;
  R[3] = xy + wz;        R[4] = T(1) - xx - zz; R[5] = yz - wx;
  R[6] = xz - wy;        R[7] = yz + wx;        R[8] = T(1) - xx - yy;
}

// Solve M x = -b for symmetric positive definite 6x6 M (floating-base a0, rbda/aba.py:241).
template <typename T>
__device__ __forceinline__ void solve6_spd_neg(T M[6][6], const T* b, T* x) {
  // LDL^T, in place: L strictly lower in M, D on the diagonal.
  T dinv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    T dj = M[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) dj -= M[j][k] * M[j][k] * M[k][k];
    M[j][j] = dj;
    dinv[j] = T(1) / dj;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      T v = M[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k] * M[k][k];
      M[i][j] = v * dinv[j];
    }
  }
  T y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    T v = -b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= M[i][k] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    T v = y[i] * dinv[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= M[k][i] * x[k];
    x[i] = v;
  }
}

// External 6D force on a link, given in the representation `repr` of the data (api/model.py:2641-2646,
// api/common.py:160-222), added to the wrench (fe, ne) about the link origin in world axes (the frame F_i):
//   inertial-fixed: moment about the world origin          -> ne += n - p x f
//   mixed:          frame at the link origin, world axes    == F_i, taken as is
//   body-fixed:     frame at the link origin, link axes     -> rotate both parts by W_R_L
// (R, p): world pose of the link AT THE CURRENT STEP, so the forces of a fused rollout follow the links exactly like
// repeated `step` calls re-express them with every state's link transforms.
template <typename T>
__device__ __forceinline__ void add_external_wrench(const int repr, const T* fx, const T* R, const T* p, T* fe, T* ne) {
  const T f[3] = {fx[0], fx[1], fx[2]}, mo[3] = {fx[3], fx[4], fx[5]};
  if (repr == 1) {
    T a[3], b[3];
    mat3_vec(R, f, a);
    mat3_vec(R, mo, b);
    fe[0] += a[0]; fe[1] += a[1]; fe[2] += a[2];
    ne[0] += b[0]; ne[1] += b[1]; ne[2] += b[2];
  } else {
    fe[0] += f[0]; fe[1] += f[1]; fe[2] += f[2];
    ne[0] += mo[0]; ne[1] += mo[1]; ne[2] += mo[2];
    if (repr == 0) {
      T t[3];
      cross3(p, f, t);
      ne[0] -= t[0]; ne[1] -= t[1]; ne[2] -= t[2];
    }
  }
}

// words of T per environment: link records, point records, then 16 bytes that hold the
// mbarrier of the environment's bulk loads (env_mbar)
template <typename T>
__host__ __device__ inline size_t env_ws_words(int nL, int nc) {
  size_t w = (size_t)nL * REC + (size_t)nc * PTREC;
  w = (w + 3) & ~size_t(3);  // keep 16-byte alignment
  return w + 16 / sizeof(T);
}
template <typename T>
__device__ __forceinline__ unsigned long long* env_mbar(T* ws, int nL, int nc) {
  return reinterpret_cast<unsigned long long*>(ws + env_ws_words<T>(nL, nc) - 16 / sizeof(T));
}

// lam_H_i = lam_H_pre * J(s) * suc_H_i  (api/kin_dyn_parameters.py:396-451,
// math/joint_model.py:146-200, math/rotation.py:58-84) from the precomputed constant
// matrices: R = M0 + cos(s) M1 + sin(s) M2, t = TPRE + s RA.
template <typename T>
__device__ __forceinline__ void joint_rel_transform(const Params<T>& P, const int flags, const T* c, int jtype, int i, T s,
                                                    T* Rrel, T* trel) {
  T sn = T(0), cs = T(1);
  if (jtype == 1) sincos_t(s, &sn, &cs);
#pragma unroll
  for (int k = 0; k < 9; ++k) Rrel[k] = c[C_M0 + k] + cs * c[C_M1 + k] + sn * c[C_M2 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) trel[k] = c[C_TPRE + k] + s * c[C_RA + k];
  if (flags & F_SUC_NONID) {
    const T* su = P.csuc + (size_t)i * 12;
    T Rs[9], ts[3], tmp[9], t2[3];
    ldn<9>(su, Rs);
    ldn<3>(su + 9, ts);
    mat3_vec(Rrel, ts, t2);  // (Rpre RJ) t_suc
    trel[0] += t2[0]; trel[1] += t2[1]; trel[2] += t2[2];
    mat3_mul(Rrel, Rs, tmp);
    stn<9>(Rrel, tmp);
  }
}

// the 6x6 adjoint of H^-1 for H = (R, t): [[R^T, S(p')R^T],[0, R^T]], p' = -R^T t
template <typename T>
__device__ __forceinline__ void inverse_adjoint(T* X, const T* R, const T* t) {
  T pi[3];
  mat3T_vec(R, t, pi);
  pi[0] = -pi[0]; pi[1] = -pi[1]; pi[2] = -pi[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const T rt = R[3 * j + i];  // R^T(i,j)
      X[6 * i + j] = rt;
      X[6 * (i + 3) + (j + 3)] = rt;
      X[6 * (i + 3) + j] = T(0);
    }
  }
  // (S(p) M)(i,j) = (p x col_j(M))_i ; col_j(R^T) = row_j(R)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const T c0 = R[3 * j + 0], c1 = R[3 * j + 1], c2 = R[3 * j + 2];
    X[6 * 0 + 3 + j] = pi[1] * c2 - pi[2] * c1;
    X[6 * 1 + 3 + j] = pi[2] * c0 - pi[0] * c2;
    X[6 * 2 + 3 + j] = pi[0] * c1 - pi[1] * c0;
  }
}

template <typename T>
__device__ __forceinline__ void store_transform(T* out, const T* R, const T* p) {
  const T H[16] = {R[0], R[1], R[2], p[0], R[3], R[4], R[5], p[1], R[6], R[7], R[8], p[2], T(0), T(0), T(0), T(1)};
  stg_vec<16>(out, H);
}

// Base state of one environment, replicated in the registers of every lane of the group.
template <typename T>
struct BaseState {
  T qn[4];    // normalised quaternion (JaxSimModelData.base_orientation, api/data.py:267-286)
  T R[9];     // W_R_B
  T p[3];     // W_p_B
  T vlin[3];  // inertial-fixed base linear velocity
  T w[3];     // base angular velocity
};

// The world transform T = W_H_B * suc_H_i[0] * W_H_B^-1 that maps the ABA chain poses
// (rooted at W_H_B, rbda/aba.py:79-93) to the FK poses (rooted at W_H_B suc_H_i[0],
// rbda/forward_kinematics.py:66-67).  Identity for URDF floating-base models.
template <typename T>
struct FkMap {
  T R[9];
  T p[3];
};

template <typename T>
__device__ __forceinline__ void make_fk_map(const T* cst0, const BaseState<T>& b, FkMap<T>& m) {
  T Rg[9], tg[3], tmp[9];  // suc_H_i[0]: record 0 of the constants
  ldn<9>(cst0 + C_M0, Rg);
  ldn<3>(cst0 + C_TPRE, tg);
  mat3_mul(b.R, Rg, tmp);
  // R_T = R_B R_G R_B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) m.R[3 * i + j] = tmp[3 * i] * b.R[3 * j] + tmp[3 * i + 1] * b.R[3 * j + 1] + tmp[3 * i + 2] * b.R[3 * j + 2];
  // p_T = p_B + R_B t_G - R_T p_B
  T a[3], c[3];
  mat3_vec(b.R, tg, a);
  mat3_vec(m.R, b.p, c);
  m.p[0] = b.p[0] + a[0] - c[0];
  m.p[1] = b.p[1] + a[1] - c[1];
  m.p[2] = b.p[2] + a[2] - c[2];
}

// FK view of link i: pose (R,p) as the reference's forward_kinematics_model sees it and the
// link velocity in "mixed at the link origin" form (vlin = velocity of the origin, w).
template <typename T>
__device__ __forceinline__ void fk_view(const int flags, const bool floating, const BaseState<T>& b, const FkMap<T>& fm, const T* rec,
                                        T* R, T* p, T* vlin, T* w) {
  T Ra[9], pa[3], va[6];
  ldn<9>(rec + O_R, Ra);
  ldn<3>(rec + O_P, pa);
  ldn<6>(rec + O_V, va);
  if (!(flags & F_GENERIC_FK)) {
    stn<9>(R, Ra);
    stn<3>(p, pa);
    vlin[0] = va[0]; vlin[1] = va[1]; vlin[2] = va[2];
    w[0] = va[3]; w[1] = va[4]; w[2] = va[5];
    return;
  }
  mat3_mul(fm.R, Ra, R);
  mat3_vec(fm.R, pa, p);
  p[0] += fm.p[0]; p[1] += fm.p[1]; p[2] += fm.p[2];
  // inertial-fixed ABA-chain velocity of the link and of the chain root
  T Wv[3], d[3], dw[3], t[3];
  cross3(pa, va + 3, Wv);
  Wv[0] += va[0]; Wv[1] += va[1]; Wv[2] += va[2];
  const T f = floating ? T(1) : T(0);  // chain root velocity: W_v_WB if floating else 0
  d[0] = Wv[0] - f * b.vlin[0]; d[1] = Wv[1] - f * b.vlin[1]; d[2] = Wv[2] - f * b.vlin[2];
  dw[0] = va[3] - f * b.w[0]; dw[1] = va[4] - f * b.w[1]; dw[2] = va[5] - f * b.w[2];
  // X_T [d; dw] = [R_T d + p_T x (R_T dw); R_T dw]
  T rd[3], rw[3];
  mat3_vec(fm.R, d, rd);
  mat3_vec(fm.R, dw, rw);
  cross3(fm.p, rw, t);
  T Fv[3];
  Fv[0] = b.vlin[0] + rd[0] + t[0]; Fv[1] = b.vlin[1] + rd[1] + t[1]; Fv[2] = b.vlin[2] + rd[2] + t[2];
  w[0] = b.w[0] + rw[0]; w[1] = b.w[1] + rw[1]; w[2] = b.w[2] + rw[2];
  cross3(w, p, t);  // mixed at p: vlin = Fv + w x p
  vlin[0] = Fv[0] + t[0]; vlin[1] = Fv[1] + t[1]; vlin[2] = Fv[2] + t[2];
}

#ifndef B200SIM_KTHREADS
#define B200SIM_KTHREADS 256
#endif
template <int G>
struct LaunchBounds {
  static constexpr int kThreads = (G <= 8) ? B200SIM_KTHREADS : 512;  // 256 threads: the 255-register budget (288 rounds up to 384 threads = 168 registers)
};

// append an environment to the produced work list (rigid-contact cascade)
template <typename T>
__device__ __forceinline__ void over_push(const Params<T>& P, int env, int impact_only) {
  const int slot = atomicAdd(P.over_count, 1);
  P.over_list[slot] = env | (impact_only << 31);
}

// Phase timeline of one warp (diagnostic, b200sim_debug_counters): when the debug buffer is
// enabled, lane 0 of warp 0 of block 0 stores clock64() at the phase boundaries of its first
// environment into dbg[8 + k].  A uniform, predicted-not-taken branch otherwise.
#define B200SIM_PHASE_MARK(k)                                                                   \
  do {                                                                                           \
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0 && env0 == first) P.dbg[8 + (k)] = (unsigned long long)clock64(); \
  } while (0)

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
// SPEC = 1 is the instance for the common case -- MODE_STEP of a floating-base URDF model
// (identity successor transforms, FK poses == ABA chain poses) with soft contacts or no
// collidable points: the mode / flag / base-type tests fold at compile time, which removes
// the other modes' code from the instruction stream (the generic instance is ~6.7k SASS
// instructions walked once per step: instruction fetch is a measurable share of its time).
template <typename T, int G, int SPEC = 0>
__global__ void __launch_bounds__(LaunchBounds<G>::kThreads, 1) step_kernel(const Params<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm_cst = reinterpret_cast<T*>(smem_raw);
  const int nL = P.nL, n = P.n, nc = P.nc;
  T* sm_pt = sm_cst + (size_t)nL * CREC;
  const size_t pt_words = ((size_t)nc * 3 + 3) & ~size_t(3);
  int* sm_itab = reinterpret_cast<int*>(sm_pt + pt_words);
  const size_t itab_words = ((size_t)P.itab_words + 3) & ~size_t(3);
  T* ws_base = reinterpret_cast<T*>(sm_itab + itab_words);

  // ---- stage the model once per block: 16-byte cp.async chunks, all in flight at once
  // (the device blobs are padded to whole chunks by b200sim_model_create)
  // Programmatic dependent launch: let the NEXT launch on the stream be set up while this grid runs
  // (its blocks cannot become resident before ours exit -- one block fills an SM's shared memory --
  // but its launch latency disappears behind our execution).  No-ops for ordinary launches.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[8 + 0] = (unsigned long long)clock64();
  if (P.dbg && threadIdx.x == 0 && blockIdx.x < 512) {  // per-block start / end wall clock (ns)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.dbg[40 + 2 * blockIdx.x] = t;
  }
  stage_async(sm_cst, P.cst, nL * CREC);
  stage_async(sm_pt, P.pt_pos, (int)pt_words);
  stage_async(sm_itab, P.itab, (int)itab_words);
  __pipeline_commit();
  if (SPEC && (threadIdx.x & (G - 1)) == 0) {
    mbar_init(env_mbar(ws_base + (size_t)(threadIdx.x / G) * env_ws_words<T>(nL, nc), nL, nc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // The specialised instance reads nothing of the model before the kinematics of the first
  // environment: it waits for the staging there, behind the input loads it has issued meanwhile.
  if (!SPEC) __pipeline_wait_prior(0);
  __syncthreads();
  // everything above touched only the constant model; the state below may have been written by the
  // previous launch on the stream: wait for it to complete and flush (no-op without the PDL attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[8 + 1] = (unsigned long long)clock64();

  const int* parent = sm_itab + P.o_parent;
  const int* jtypes = sm_itab + P.o_jtype;
  const int* lvl_start = sm_itab + P.o_lvl_start;
  const int* lvl_links = sm_itab + P.o_lvl_links;
  const int* child_start = sm_itab + P.o_child_start;
  const int* child_idx = sm_itab + P.o_child_idx;
  const int* pt_start = sm_itab + P.o_pt_start;
  const int* pt_idx = sm_itab + P.o_pt_idx;
  const int* pt_body = sm_itab + P.o_pt_body;
  const int* pt_enabled = sm_itab + P.o_pt_enabled;
  // packed level-walk rows of the specialised instance (G = 8 or 16): one int per lane and row
  const int* rows = sm_itab + (G == 16 ? P.o_rows16 : P.o_rows8);
  const int nrows = SPEC ? (G == 16 ? P.n_rows16 : P.n_rows8) : 0;
  const int rlane = threadIdx.x & (G - 1);

  const int lane = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  const size_t wsw = env_ws_words<T>(nL, nc);
  T* ws = ws_base + (size_t)grp * wsw;
  T* ptws = ws + (size_t)nL * REC;
  const long long stride = (long long)gridDim.x * P.envs_per_block;
  const T dt = P.dt;
  const int mode = SPEC ? (int)MODE_STEP : P.mode;
  const int flags = SPEC ? (P.flags & (F_SQRT_P | F_SQRT_Q | F_TMA_STORE | F_BULK_IN)) : P.flags;
  const bool floating = SPEC ? true : (P.floating != 0);
  const bool with_contacts = (mode == MODE_STEP) || (mode == MODE_DYN);
  const bool soft = SPEC ? (nc > 0) : (with_contacts && (P.contact_model == 1) && nc > 0);
  const bool rigid = SPEC ? false : ((mode == MODE_STEP) && (P.contact_model >= 2) && nc > 0);  // rigid (2) or relaxed-rigid (3)
  const bool tma = (flags & F_TMA_STORE) != 0;

  // the number of loop trips is uniform across the block so that __syncwarp() is safe
  const long long first = (long long)blockIdx.x * P.envs_per_block;
  unsigned in_parity = 0;  // phase of the environment's mbarrier (one bulk-load transaction per trip)
  for (long long env0 = first; env0 < P.B; env0 += stride) {
    long long env = env0 + grp;
    bool active = env < P.B;
    // idle groups shadow DISTINCT valid environments with their stores masked (many groups fetching the
    // same rows serialise in L2: measured +3 us on the one block that carries the ragged tail)
    if (!active) env = env % P.B;

    // =========================================================== prefetch (one burst)
    // joint state, first-step torque reference and contact state go global -> shared with
    // cp.async; the 13 base scalars go to registers.  Nothing is consumed before the wait.
    if (tma) {
      // the previous environment's staging areas are reused below; in the compact final phase only
      // lane 0 of the group owns the bulk store, so the others must not run ahead of its wait
      tma_store_wait_read();
      __syncwarp();
    }
    const bool use_cached = (mode == MODE_STEP) && P.Hin && P.Vin && !(flags & F_GENERIC_FK);
    // bulk_in: the environment's (nL,4,4) and (nL,6) blocks are contiguous in HBM, so they arrive
    // with two cp.async.bulk (one lane, completion on the environment's mbarrier) into the head of
    // the still empty workspace -- [nL x 16 | nL x 6] words -- instead of 6 LDGSTS per link; the joint
    // state then goes to registers (its record slots overlap that staging area).
    const bool bulk_in = SPEC && use_cached && (flags & F_BULK_IN) && (16 * nL <= 54 * G);
    T s_r[4], sd_r[4], tr_r[4];
    if (bulk_in) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses of the last trip before the async writes
      __syncwarp();
      if (lane == 0) {
        unsigned long long* bar = env_mbar(ws, nL, nc);
        const unsigned bH = (unsigned)(nL * 16 * sizeof(T)), bV = (unsigned)(nL * 6 * sizeof(T));
        mbar_expect_tx(bar, bH + bV);
        tma_load_bulk(ws, P.Hin + env * nL * 16, bH, bar);
        tma_load_bulk(ws + (size_t)nL * 16, P.Vin + env * nL * 6, bV, bar);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = lane + t * G;
        s_r[t] = T(0); sd_r[t] = T(0); tr_r[t] = T(0);
        if (i >= 1 && i < nL) {
          s_r[t] = P.s[env * n + (i - 1)];
          sd_r[t] = P.sd[env * n + (i - 1)];
          if (P.tau) tr_r[t] = P.tau[env * n + (i - 1)];
        }
      }
    } else {
      if (use_cached) {
        // cached kinematics of the input state: rows [R | p] of W_H_L and the 6D velocity land in
        // the (still unused) IA / c slots of the record, 16 bytes per cp.async
        constexpr int per = 16 / sizeof(T);        // elements per 16-byte chunk
        for (int i = lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * REC;
          const T* H = P.Hin + (env * nL + i) * 16;
          const T* V = P.Vin + (env * nL + i) * 6;
#pragma unroll
          for (int k = 0; k < 12; k += per) __pipeline_memcpy_async(ri + O_X + k, H + k, 16);
          if (sizeof(T) == 4) {
#pragma unroll
            for (int k = 0; k < 6; k += 2) __pipeline_memcpy_async(ri + O_C + k, V + k, 8);
          } else {
#pragma unroll
            for (int k = 0; k < 6; k += per) __pipeline_memcpy_async(ri + O_C + k, V + k, 16);
          }
        }
      }
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        cp_async_elem(ri + O_S, P.s + env * n + (i - 1));
        cp_async_elem(ri + O_SD, P.sd + env * n + (i - 1));
        if (P.tau) cp_async_elem(ri + O_TREF, P.tau + env * n + (i - 1));
        else ri[O_TREF] = T(0);
      }
    }
    if (with_contacts) {
      for (int k = lane; k < nc; k += G) {
        T* pw = ptws + (size_t)k * PTREC + PT_M;
        if (P.m) {
          const T* src = P.m + (env * nc + k) * 3;
          cp_async_elem(pw, src); cp_async_elem(pw + 1, src + 1); cp_async_elem(pw + 2, src + 2);
        } else {
          pw[0] = T(0); pw[1] = T(0); pw[2] = T(0);
        }
      }
    }
    __pipeline_commit();
    B200SIM_PHASE_MARK(2);

    // =========================================================== phase 0: base
    BaseState<T> b;
    {
      const T* q = P.q + env * 4;
      T qr[4] = {q[0], q[1], q[2], q[3]};
      ldn<3>(P.p + env * 3, b.p);
      ldn<3>(P.vlin + env * 3, b.vlin);
      ldn<3>(P.omega + env * 3, b.w);
      const T nrm = sqrt_t(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      T den;
      if (mode == MODE_FK) den = (nrm == T(0)) ? T(1) : nrm;            // data.replace (api/data.py:441-447)
      else den = nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0));       // base_orientation (api/data.py:283-285)
      const T inv = rcp_t(den);
#pragma unroll
      for (int k = 0; k < 4; ++k) b.qn[k] = qr[k] * inv;
      quat_to_dcm(b.qn, b.R);
    }
    FkMap<T> fm;
    if (flags & F_GENERIC_FK) make_fk_map(sm_cst, b, fm);

    auto write_base_record = [&](const BaseState<T>& bs) {
      if (lane == 0) {
        T* r0 = ws;
        stn<9>(r0 + O_R, bs.R);
        stn<3>(r0 + O_P, bs.p);
        T v0[6];
        if (floating) {
          T t[3];
          cross3(bs.w, bs.p, t);  // velocity of the base origin: v_lin + w x p
          v0[0] = bs.vlin[0] + t[0]; v0[1] = bs.vlin[1] + t[1]; v0[2] = bs.vlin[2] + t[2];
          v0[3] = bs.w[0]; v0[4] = bs.w[1]; v0[5] = bs.w[2];
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) v0[k] = T(0);  // rbda/aba.py:109-121: v[0] stays zero
        }
        stn<6>(r0 + O_V, v0);
      }
    };

    // FK + velocity chain over the tree levels (rbda/forward_kinematics.py:80-113 and
    // pass 1 of rbda/aba.py:131-171, in F_i coordinates).  `for_aba` also stores the world
    // joint axis and the offset to the parent (not needed when only caches follow, and
    // their slots then hold the joint-adjoint staging).
    auto fk_chain = [&](const bool for_aba) {
      auto fk_link = [&](const int i, const int par, const int jt) {
          const T* rp = ws + (size_t)par * REC;
          T* ri = ws + (size_t)i * REC;
          T Rp[9], pp[3], vp[6], Rrel[9], trel[3];
          ldn<9>(rp + O_R, Rp);
          ldn<3>(rp + O_P, pp);
          ldn<6>(rp + O_V, vp);
          ldn<9>(ri + O_R, Rrel);
          ldn<3>(ri + O_P, trel);
          T R[9], r[3], pw[3];
          mat3_mul(Rp, Rrel, R);
          mat3_vec(Rp, trel, r);
          pw[0] = pp[0] + r[0]; pw[1] = pp[1] + r[1]; pw[2] = pp[2] + r[2];
          T ax[3], aw[3];
          ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
          mat3_vec(R, ax, aw);
          const T sdi = ri[O_SD];
          T v[6];
          cross3(vp + 3, r, v);
          v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
          v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
          if (jt == 1) { v[3] += sdi * aw[0]; v[4] += sdi * aw[1]; v[5] += sdi * aw[2]; }
          else if (jt == 2) { v[0] += sdi * aw[0]; v[1] += sdi * aw[1]; v[2] += sdi * aw[2]; }
          stn<9>(ri + O_R, R);
          stn<3>(ri + O_P, pw);
          stn<6>(ri + O_V, v);
          if (for_aba) {
            stn<3>(ri + O_RR, r);
            stn<3>(ri + O_AX, aw);
          }
      };
      if (SPEC) {
        // packed rows: the next row's entry is fetched while the current one is processed
        int e = nrows > 0 ? rows[rlane] : 0xFF;
        for (int r = 0; r < nrows; ++r) {
          __syncwarp();
          const int en = (r + 1 < nrows) ? rows[(r + 1) * G + rlane] : 0xFF;
          if ((e & 0xFF) != 0xFF) fk_link(e & 0xFF, (e >> 8) & 0xFF, (e >> 27) & 3);
          e = en;
        }
      } else {
        for (int l = 1; l <= P.depth; ++l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            fk_link(i, parent[i], jtypes[i]);
          }
        }
      }
      __syncwarp();
    };

    auto write_fk_caches = [&](const BaseState<T>& bs, const FkMap<T>& f) {
      for (int i = lane; i < nL; i += G) {
        T R[9], p[3], vl[3], w[3];
        fk_view(flags, floating, bs, f, ws + (size_t)i * REC, R, p, vl, w);
        if (P.W_H_L) store_transform(P.W_H_L + (env * nL + i) * 16, R, p);
        if (P.W_v) {
          T t[3];
          cross3(p, w, t);  // inertial-fixed linear part: vlin + p x w
          const T o[6] = {vl[0] + t[0], vl[1] + t[1], vl[2] + t[2], w[0], w[1], w[2]};
          stg_vec6(P.W_v + (env * nL + i) * 6, o);
        }
      }
    };

    // joint adjoint output of link i (R,t = lam_H_i): TMA from the record's staging area
    // or 128-bit stores from registers
    auto emit_joint_adjoint = [&](T* ri, long long i, const T* R, const T* t) {
      T X[36];
      inverse_adjoint(X, R, t);
      T* dst = P.iXl + (env * nL + i) * 36;
      if (tma) {
        stn<36>(ri + O_X, X);
        tma_store_bulk(dst, ri + O_X, 36 * sizeof(T));
      } else {
        stg_vec<36>(dst, X);
      }
    };

    B200SIM_PHASE_MARK(3);
    __pipeline_wait_prior(0);
    if (SPEC && env0 == first) __syncthreads();  // the model blob staged by all threads of the block
    B200SIM_PHASE_MARK(4);
    if (!use_cached) write_base_record(b);

    if (bulk_in) {
      // ========================================================= phases 1-2 from the caches (bulk)
      mbar_wait(env_mbar(ws, nL, nc), in_parity);
      in_parity ^= 1u;
      // The records (60 words per link) overwrite the staging area (22 words per link) from the
      // top: the trips run over DESCENDING link indices, each one reads its links' staged rows,
      // synchronises, then writes their records -- which only clobbers staging words of links that
      // earlier trips have consumed (16 nL <= 54 G, checked by bulk_in).
#pragma unroll
      for (int t = 3; t >= 0; --t) {
        if (t * G < nL) {
          const int i = lane + t * G;
          T H[12], V[6];
          if (i < nL) {
            ldn<12>(ws + (size_t)i * 16, H);
            ldn<6>(ws + (size_t)nL * 16 + (size_t)i * 6, V);
          }
          __syncwarp();
          if (i < nL) {
            T* ri = ws + (size_t)i * REC;
            const T R[9] = {H[0], H[1], H[2], H[4], H[5], H[6], H[8], H[9], H[10]};
            const T p[3] = {H[3], H[7], H[11]};
            T v[6], tt[3];
            cross3(V + 3, p, tt);  // velocity of the link origin: W_v_lin + w x p
            v[0] = V[0] + tt[0]; v[1] = V[1] + tt[1]; v[2] = V[2] + tt[2];
            v[3] = V[3]; v[4] = V[4]; v[5] = V[5];
            stn<9>(ri + O_R, R);
            stn<3>(ri + O_P, p);
            stn<6>(ri + O_V, v);
            if (i > 0) {
              T ax[3], aw[3];
              ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
              mat3_vec(R, ax, aw);
              stn<3>(ri + O_AX, aw);
              ri[O_S] = s_r[t];
              ri[O_SD] = sd_r[t];
              ri[O_TREF] = tr_r[t];
            }
          }
        }
      }
      __syncwarp();
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* rp = ws + (size_t)parent[i] * REC;
        const T r[3] = {ri[O_P] - rp[O_P], ri[O_P + 1] - rp[O_P + 1], ri[O_P + 2] - rp[O_P + 2]};
        stn<3>(ri + O_RR, r);
      }
      __syncwarp();
    } else if (use_cached) {
      // ========================================================= phases 1-2 from the caches
      // The input data carries the link transforms / velocities of its own state (they are
      // what the reference's contact code reads, api/contact.py:39-43): 88 B per link from
      // L2/HBM instead of a sincos, two 3x3 products and a sequential tree walk.
      for (int i = lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        T H[12], V[6];
        ldn<12>(ri + O_X, H);   // staged by the input burst
        ldn<6>(ri + O_C, V);
        const T R[9] = {H[0], H[1], H[2], H[4], H[5], H[6], H[8], H[9], H[10]};
        const T p[3] = {H[3], H[7], H[11]};
        T v[6], t[3];
        cross3(V + 3, p, t);  // velocity of the link origin: W_v_lin + w x p
        v[0] = V[0] + t[0]; v[1] = V[1] + t[1]; v[2] = V[2] + t[2];
        v[3] = V[3]; v[4] = V[4]; v[5] = V[5];
        stn<9>(ri + O_R, R);
        stn<3>(ri + O_P, p);
        stn<6>(ri + O_V, v);
        if (i > 0) {
          T ax[3], aw[3];
          ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
          mat3_vec(R, ax, aw);
          stn<3>(ri + O_AX, aw);
        }
      }
      __syncwarp();
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* rp = ws + (size_t)parent[i] * REC;
        const T r[3] = {ri[O_P] - rp[O_P], ri[O_P + 1] - rp[O_P + 1], ri[O_P + 2] - rp[O_P + 2]};
        stn<3>(ri + O_RR, r);
      }
      __syncwarp();
    } else {
      // ========================================================= phase 1: joint transforms
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        T Rrel[9], trel[3];
        joint_rel_transform(P, flags, sm_cst + (size_t)i * CREC, jtypes[i], i, ri[O_S], Rrel, trel);
        stn<9>(ri + O_R, Rrel);
        stn<3>(ri + O_P, trel);
        if (mode == MODE_FK && P.iXl && active) emit_joint_adjoint(ri, i, Rrel, trel);
      }
      // ========================================================= phase 2: FK chain
      fk_chain(mode != MODE_FK);
    }

    if (mode == MODE_FK) {
      // JaxSimModelData.build / replace: caches of the given state
      if (active) {
        if (lane == 0) {
          if (P.q_o) stn<4>(P.q_o + env * 4, b.qn);
          if (P.W_H_B) store_transform(P.W_H_B + env * 16, b.R, b.p);
          if (P.iXl) {
            // index 0: Ad((W_H_B suc_H_i[0])^-1)  (api/kin_dyn_parameters.py:417-449)
            T R0[9], p0[3], t[3], X[36];
            mat3_mul(b.R, sm_cst + C_M0, R0);      // suc_H_i[0] lives in record 0 of the constants
            mat3_vec(b.R, sm_cst + C_TPRE, t);
            p0[0] = b.p[0] + t[0]; p0[1] = b.p[1] + t[1]; p0[2] = b.p[2] + t[2];
            inverse_adjoint(X, R0, p0);
            stg_vec<36>(P.iXl + env * nL * 36, X);
          }
        }
        write_fk_caches(b, fm);
      }
      __syncwarp();
      continue;
    }

    B200SIM_PHASE_MARK(5);
    T Wa[6];
    for (int step = 0; step < P.nsteps; ++step) {
      const bool last = (step == P.nsteps - 1);
      const T* fext_step = P.fext ? P.fext + (long long)step * P.fext_step_stride : nullptr;

      // ========================================================= contacts (point-parallel)
      if (rigid) {
        // RigidContacts, level 0 of the cascade: this kernel only finishes environments whose
        // collidable points are all above the ground at t and at t+dt (then the contact forces
        // and the impact are exactly zero, rbda/contacts/rigid.py:222-436).  The others go to
        // the work list of the warp-per-environment rigid kernel; their stores are masked.
        bool touch = false;
        for (int k = lane; k < nc; k += G) {
          const T* rb = ws + (size_t)pt_body[k] * REC;
          const T* Lp = sm_pt + 3 * k;
          const T z = rb[O_P + 2] + rb[O_R + 6] * Lp[0] + rb[O_R + 7] * Lp[1] + rb[O_R + 8] * Lp[2];
          touch = touch || (pt_enabled[k] && (P.h_terrain - z > T(0)));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, touch);
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
        if (bal & gmask) {
          if (active && lane == 0) over_push(P, (int)env, 0);
          active = false;
        }
        // every environment of this warp is left to the rigid kernel (a standing batch: all of them): nothing to finish
        // here (no block-level barrier inside the trip loop; the steps loop runs once for the rigid cascade)
        if (__all_sync(0xffffffffu, !active)) break;
      }
      bool env_touches = false;  // some point of this environment is in contact (group-uniform after the ballot)
      if (soft) {
        for (int k = lane; k < nc; k += G) {
          const int bi = pt_body[k];
          const T* rb = ws + (size_t)bi * REC;
          T R[9], p[3], vl[3], w[3];
          fk_view(flags, floating, b, fm, rb, R, p, vl, w);
          T Lp[3], d[3], pc[3], pd[3];
          ldn<3>(sm_pt + 3 * k, Lp);
          mat3_vec(R, Lp, d);
          pc[0] = p[0] + d[0]; pc[1] = p[1] + d[1]; pc[2] = p[2] + d[2];
          cross3(w, d, pd);
          pd[0] += vl[0]; pd[1] += vl[1]; pd[2] += vl[2];
          T* pw = ptws + (size_t)k * PTREC;
          T m[3];
          ldn<3>(pw + PT_M, m);
          T f[3] = {T(0), T(0), T(0)};
          T md[3] = {T(0), T(0), T(0)};
          if (pt_enabled[k]) {
            // compute_penetration_data, FlatTerrain: n = z (rbda/contacts/common.py:25-63)
            const T delta = max_t(T(0), P.h_terrain - pc[2]);
            const T KoD = P.K * rcp_t(P.D);
            if (delta <= T(0)) {
              // no contact: zero force, the tangential deformation relaxes (soft.py:318-330).  A branch, not a
              // select: a warp whose points are all above the ground skips the Hunt/Crossley arithmetic.
              md[0] = -KoD * m[0]; md[1] = -KoD * m[1]; md[2] = -KoD * m[2];
            } else {
              const T ddot = -pd[2];
              const T eps = Lim<T>::eps();
              const T dp = (flags & F_SQRT_P) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.pexp);
              const T dq = (flags & F_SQRT_Q) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.qexp);
              const T Kdp = P.K * dp, Ddq = P.D * dq;
              const T fn = max_t(T(0), Kdp * delta + Ddq * ddot);
              // tangential quantities (n = z): v_t = (vx,vy,0), m_t = (mx,my,0), m_n = (0,0,mz)
              T ft0 = -(Kdp * m[0] + Ddq * pd[0]);
              T ft1 = -(Kdp * m[1] + Ddq * pd[1]);
              const T mufn = P.mu * fn;
              const bool sticking = ft0 * ft0 + ft1 * ft1 <= mufn * mufn;
              if (sticking) {
                md[0] = pd[0]; md[1] = pd[1]; md[2] = -KoD * m[2];
              } else {
                const T nrm = sqrt_t(ft0 * ft0 + ft1 * ft1);
                const T idn = rcp_t(nrm + eps * (nrm == T(0) ? T(1) : T(0)));
                const T sc = min_t(mufn, nrm) * idn;
                ft0 *= sc; ft1 *= sc;
                const T iD = rcp_t(Ddq);
                md[0] = -(ft0 + Kdp * m[0]) * iD; md[1] = -(ft1 + Kdp * m[1]) * iD; md[2] = T(0);
              }
              f[0] = ft0; f[1] = ft1; f[2] = fn;
              env_touches = true;
            }
          }
          // lever arm w.r.t. the origin of the ABA-chain frame of the body
          T lev[3];
          if (flags & F_GENERIC_FK) {
            lev[0] = pc[0] - rb[O_P]; lev[1] = pc[1] - rb[O_P + 1]; lev[2] = pc[2] - rb[O_P + 2];
          } else {
            lev[0] = d[0]; lev[1] = d[1]; lev[2] = d[2];
          }
          stn<3>(pw + PT_F, f);
          stn<3>(pw + PT_LEV, lev);
          if (mode == MODE_DYN) {
            // system_dynamics: the contact-state derivative itself (api/ode.py:174-225)
            if (active && P.m_o) {
              T* mo = P.m_o + (env * nc + k) * 3;
              mo[0] = md[0]; mo[1] = md[1]; mo[2] = md[2];
            }
          } else {
            // m+ = m + dt * m_dot (api/integrators.py:67-71); stays on chip between steps
            m[0] += dt * md[0]; m[1] += dt * md[1]; m[2] += dt * md[2];
            stn<3>(pw + PT_M, m);
            if (last && active && P.m_o) {
              T* mo = P.m_o + (env * nc + k) * 3;
              mo[0] = m[0]; mo[1] = m[1]; mo[2] = m[2];
            }
          }
        }
      } else if (mode == MODE_STEP && last && P.m_o && P.m && active) {
        for (int k = lane; k < nc * 3; k += G) P.m_o[env * nc * 3 + k] = P.m[env * nc * 3 + k];
      } else if (mode == MODE_DYN && P.m_o && active) {
        for (int k = lane; k < nc * 3; k += G) P.m_o[env * nc * 3 + k] = T(0);
      }
      {
        // airborne environments skip the per-link accumulation of the (all zero) contact wrenches below
        const unsigned bal = __ballot_sync(0xffffffffu, env_touches);
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
        env_touches = (bal & gmask) != 0u;
      }
      __syncwarp();
      B200SIM_PHASE_MARK(6);

      // ========================================================= phase 3: link-parallel
      for (int i = lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* c = sm_cst + (size_t)i * CREC;
        T R[9], p[3], v[6];
        ldn<9>(ri + O_R, R);
        ldn<3>(ri + O_P, p);
        ldn<6>(ri + O_V, v);
        // total external wrench on the link in F_i: contacts + user forces
        T fe[3] = {T(0), T(0), T(0)}, ne[3] = {T(0), T(0), T(0)};
        if (soft && env_touches) {
          const int e = pt_start[i + 1];
          for (int kk = pt_start[i]; kk < e; ++kk) {
            const T* pw = ptws + (size_t)pt_idx[kk] * PTREC;
            T f[3], lev[3];
            ldn<3>(pw + PT_F, f);
            ldn<3>(pw + PT_LEV, lev);
            fe[0] += f[0]; fe[1] += f[1]; fe[2] += f[2];
            cross3_add(lev, f, ne);
          }
        }
        if (fext_step) {
          const T* fx = fext_step + (env * nL + i) * 6;
          add_external_wrench(P.fext_repr, fx, R, p, fe, ne);
        }
        // link inertia in world axes about the link origin
        T mass = c[C_MASS];
        T com[3], cw[3], Dl[6];
        ldn<3>(c + C_COM, com);
        ldn<6>(c + C_DL, Dl);
        if (P.mass_dir_period > 0) {
          const bool keep = (long long)i == (long long)P.mass_dir_first + env / P.mass_dir_period;
          mass = keep_tangent(mass, keep);
#pragma unroll
          for (int k = 0; k < 6; ++k) Dl[k] = keep_tangent(Dl[k], keep);
        }
        mat3_vec(R, com, cw);
        T Dw[6];
        {
          const T Df[9] = {Dl[0], Dl[1], Dl[2], Dl[1], Dl[3], Dl[4], Dl[2], Dl[4], Dl[5]};
          T Tm[9];
          mat3_mul(R, Df, Tm);
          Dw[0] = Tm[0] * R[0] + Tm[1] * R[1] + Tm[2] * R[2];
          Dw[1] = Tm[0] * R[3] + Tm[1] * R[4] + Tm[2] * R[5];
          Dw[2] = Tm[0] * R[6] + Tm[1] * R[7] + Tm[2] * R[8];
          Dw[3] = Tm[3] * R[3] + Tm[4] * R[4] + Tm[5] * R[5];
          Dw[4] = Tm[3] * R[6] + Tm[4] * R[7] + Tm[5] * R[8];
          Dw[5] = Tm[6] * R[6] + Tm[7] * R[7] + Tm[8] * R[8];
        }
        // I v = [m (v + w x c); m c x v + D w]
        T fI[3], nI[3], t[3];
        cross3(v + 3, cw, t);
        fI[0] = mass * (v[0] + t[0]); fI[1] = mass * (v[1] + t[1]); fI[2] = mass * (v[2] + t[2]);
        sym3_vec(Dw, v + 3, nI);
        cross3(cw, v, t);
        nI[0] += mass * t[0]; nI[1] += mass * t[1]; nI[2] += mass * t[2];
        // pA = v x* (I v) - f_ext = [w x fI ; v x fI + w x nI] - [fe; ne]
        T pA[6];
        cross3(v + 3, fI, pA);
        cross3(v, fI, pA + 3);
        cross3_add(v + 3, nI, pA + 3);
        pA[0] -= fe[0]; pA[1] -= fe[1]; pA[2] -= fe[2];
        pA[3] -= ne[0]; pA[4] -= ne[1]; pA[5] -= ne[2];
        if (i == 0 && !floating) {
#pragma unroll
          for (int k = 0; k < 6; ++k) pA[k] = T(0);
        }
        // c_i = v x vJ (rbda/aba.py:143-144) and the resultant joint torque
        if (i > 0) {
          const int jt = jtypes[i];
          const T sdi = ri[O_SD];
          T aw[3];
          ldn<3>(ri + O_AX, aw);
          T cc[6];
          const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
          if (jt == 1) {
            cross3(v, vJ, cc);  // v_lin x vJ_ang
            cross3(v + 3, vJ, cc + 3);
          } else {
            cross3(v + 3, vJ, cc);  // w x vJ_lin
            cc[3] = cc[4] = cc[5] = T(0);
          }
          stn<6>(ri + O_C, cc);
          const T tref = ri[O_TREF];
          T tau;
          if (mode != MODE_STEP) {
            tau = tref;
          } else {
            // api/actuation_model.py:7-126
            const T si = ri[O_S];
            const T lower = min_t(si - c[C_SMIN], T(0));
            const T upper = max_t(si - c[C_SMAX], T(0));
            T tlim = -c[C_KS] * (lower + upper);
            tlim = tlim - tlim * c[C_KD] * sdi;
            T tfr = T(0);
            if (P.enable_friction) {
              const T sg = (sdi > T(0)) ? T(1) : ((sdi < T(0)) ? T(-1) : T(0));
              tfr = -(c[C_KC] * sg + c[C_KV] * sdi);
            }
            const T tt = tref + tfr + tlim;
            const T av = abs_t(sdi);
            T lim;
            if (av <= P.w_th) lim = P.tau_max;
            else if (av <= P.w_max) lim = P.tau_max * (T(1) - (av - P.w_th) / (P.w_max - P.w_th));
            else lim = T(0);
            tau = min_t(max_t(tt, -lim), lim);
          }
          ri[O_TAU] = tau;
          // torque reference of the NEXT step: its latency hides behind the ABA passes
          if (!last && P.tau && P.tau_step_stride)
            cp_async_elem(ri + O_TREF, P.tau + (long long)(step + 1) * P.tau_step_stride + env * n + (i - 1));
        }
        // articulated inertia init (overwrites R/p): A = m 1, B = -m S(c_w), D = D_w
        T IA[21];
        IA[0] = mass; IA[1] = T(0); IA[2] = T(0); IA[3] = mass; IA[4] = T(0); IA[5] = mass;
        IA[6] = T(0);            IA[7] = mass * cw[2];   IA[8] = -mass * cw[1];
        IA[9] = -mass * cw[2];   IA[10] = T(0);          IA[11] = mass * cw[0];
        IA[12] = mass * cw[1];   IA[13] = -mass * cw[0]; IA[14] = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) IA[15 + k] = Dw[k];
        if (i == 0 && !floating) {
#pragma unroll
          for (int k = 0; k < 21; ++k) IA[k] = T(0);
        }
        stn<21>(ri + O_IA, IA);
        stn<6>(ri + O_PA, pA);
      }
      __pipeline_commit();
      B200SIM_PHASE_MARK(7);

      // ========================================================= phase 4: ABA pass 2
      // children: a contiguous index range from the packed row (SPEC) or the child list (generic)
      auto pass2_link = [&](const int i, const int par, const int jt, const int cbase, const int ncld) {
          T* ri = ws + (size_t)i * REC;
          T A[6], Bm[9], D[6], pA[6];
          ldn<6>(ri + O_IA, A);
          ldn<9>(ri + O_IB, Bm);
          ldn<6>(ri + O_ID, D);
          ldn<6>(ri + O_PA, pA);
          {
            for (int cc = 0; cc < ncld; ++cc) {
              const T* rc = ws + (size_t)(SPEC ? (cbase + cc) : child_idx[cbase + cc]) * REC;
#pragma unroll
              for (int k = 0; k < 6; ++k) A[k] += rc[O_IA + k];
#pragma unroll
              for (int k = 0; k < 9; ++k) Bm[k] += rc[O_IB + k];
#pragma unroll
              for (int k = 0; k < 6; ++k) D[k] += rc[O_ID + k];
#pragma unroll
              for (int k = 0; k < 6; ++k) pA[k] += rc[O_PA + k];
            }
          }
          T aw[3], cI[6], r[3];
          ldn<3>(ri + O_AX, aw);
          ldn<6>(ri + O_C, cI);
          ldn<3>(ri + O_RR, r);
          T Ul[3], Ua[3], d, u;
          const T tau = ri[O_TAU];
          if (jt == 1) {
            mat3_vec(Bm, aw, Ul);  // B a
            sym3_vec(D, aw, Ua);   // D a
            d = dot3(aw, Ua);
            u = tau - dot3(aw, pA + 3);
          } else {
            sym3_vec(A, aw, Ul);    // A a
            mat3T_vec(Bm, aw, Ua);  // B^T a
            d = dot3(aw, Ul);
            u = tau - dot3(aw, pA);
          }
          const T dinv = rcp_t(d);
          stn<3>(ri + O_U, Ul);
          stn<3>(ri + O_U + 3, Ua);
          ri[O_DINV] = dinv;
          ri[O_UU] = u;
          if (par != 0 || floating) {
            // Ma = IA - U U^T / d
            const T Uls[3] = {Ul[0] * dinv, Ul[1] * dinv, Ul[2] * dinv};
            const T Uas[3] = {Ua[0] * dinv, Ua[1] * dinv, Ua[2] * dinv};
            A[0] -= Uls[0] * Ul[0]; A[1] -= Uls[0] * Ul[1]; A[2] -= Uls[0] * Ul[2];
            A[3] -= Uls[1] * Ul[1]; A[4] -= Uls[1] * Ul[2]; A[5] -= Uls[2] * Ul[2];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int bb = 0; bb < 3; ++bb) Bm[3 * a + bb] -= Uls[a] * Ua[bb];
            D[0] -= Uas[0] * Ua[0]; D[1] -= Uas[0] * Ua[1]; D[2] -= Uas[0] * Ua[2];
            D[3] -= Uas[1] * Ua[1]; D[4] -= Uas[1] * Ua[2]; D[5] -= Uas[2] * Ua[2];
            // pa = pA + Ma c + U u/d
            const T ud = u * dinv;
            T pa[6], t3[3];
            sym3_vec(A, cI, t3);
            pa[0] = pA[0] + t3[0] + Ul[0] * ud; pa[1] = pA[1] + t3[1] + Ul[1] * ud; pa[2] = pA[2] + t3[2] + Ul[2] * ud;
            mat3_vec(Bm, cI + 3, t3);
            pa[0] += t3[0]; pa[1] += t3[1]; pa[2] += t3[2];
            mat3T_vec(Bm, cI, t3);
            pa[3] = pA[3] + t3[0] + Ua[0] * ud; pa[4] = pA[4] + t3[1] + Ua[1] * ud; pa[5] = pA[5] + t3[2] + Ua[2] * ud;
            sym3_vec(D, cI + 3, t3);
            pa[3] += t3[0]; pa[4] += t3[1]; pa[5] += t3[2];
            // shift to the parent's origin: X = [[1, -S(r)],[0, 1]]
            //   B'' = B' - A' S(r)        (row_i(A' S) = row_i(A') x r)
            //   D'' = D' + S(r) B' + (S(r) B'')^T
            const T Af[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
            T B2[9];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              T rowx[3];
              cross3(Af + 3 * a, r, rowx);
              B2[3 * a] = Bm[3 * a] - rowx[0]; B2[3 * a + 1] = Bm[3 * a + 1] - rowx[1]; B2[3 * a + 2] = Bm[3 * a + 2] - rowx[2];
            }
            // SB1(i,j) = (r x col_j(B'))_i ; SB2(i,j) = (r x col_j(B''))_i
            T SB1[9], SB2[9];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const T c1[3] = {Bm[j], Bm[3 + j], Bm[6 + j]};
              const T c2[3] = {B2[j], B2[3 + j], B2[6 + j]};
              T o1[3], o2[3];
              cross3(r, c1, o1);
              cross3(r, c2, o2);
              SB1[j] = o1[0]; SB1[3 + j] = o1[1]; SB1[6 + j] = o1[2];
              SB2[j] = o2[0]; SB2[3 + j] = o2[1]; SB2[6 + j] = o2[2];
            }
            T D2[6];
            D2[0] = D[0] + SB1[0] + SB2[0];
            D2[1] = D[1] + SB1[1] + SB2[3];
            D2[2] = D[2] + SB1[2] + SB2[6];
            D2[3] = D[3] + SB1[4] + SB2[4];
            D2[4] = D[4] + SB1[5] + SB2[7];
            D2[5] = D[5] + SB1[8] + SB2[8];
            cross3_add(r, pa, pa + 3);
            stn<6>(ri + O_IA, A);
            stn<9>(ri + O_IB, B2);
            stn<6>(ri + O_ID, D2);
            stn<6>(ri + O_PA, pa);
          }
      };
      if (SPEC) {
        int e = nrows > 0 ? rows[(nrows - 1) * G + rlane] : 0xFF;
        for (int r = nrows - 1; r >= 0; --r) {
          __syncwarp();
          const int en = r > 0 ? rows[(r - 1) * G + rlane] : 0xFF;
          if ((e & 0xFF) != 0xFF) pass2_link(e & 0xFF, (e >> 8) & 0xFF, (e >> 27) & 3, (e >> 16) & 0xFF, (e >> 24) & 7);
          e = en;
        }
      } else {
        for (int l = P.depth; l >= 1; --l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            pass2_link(i, parent[i], jtypes[i], child_start[i], child_start[i + 1] - child_start[i]);
          }
        }
      }
      __syncwarp();
      B200SIM_PHASE_MARK(8);

      // ========================================================= phase 5: base acceleration
      if (lane == 0) {
        T a0[6];
        if (floating) {
          T* r0 = ws;
          T A[6], Bm[9], D[6], pA[6];
          ldn<6>(r0 + O_IA, A);
          ldn<9>(r0 + O_IB, Bm);
          ldn<6>(r0 + O_ID, D);
          ldn<6>(r0 + O_PA, pA);
          const int ce = child_start[1];
          for (int cc = child_start[0]; cc < ce; ++cc) {
            const T* rc = ws + (size_t)child_idx[cc] * REC;
#pragma unroll
            for (int k = 0; k < 6; ++k) A[k] += rc[O_IA + k];
#pragma unroll
            for (int k = 0; k < 9; ++k) Bm[k] += rc[O_IB + k];
#pragma unroll
            for (int k = 0; k < 6; ++k) D[k] += rc[O_ID + k];
#pragma unroll
            for (int k = 0; k < 6; ++k) pA[k] += rc[O_PA + k];
          }
          T M[6][6];
          M[0][0] = A[0]; M[0][1] = A[1]; M[0][2] = A[2]; M[1][1] = A[3]; M[1][2] = A[4]; M[2][2] = A[5];
          M[1][0] = A[1]; M[2][0] = A[2]; M[2][1] = A[4];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) { M[a][3 + bb] = Bm[3 * a + bb]; M[3 + bb][a] = Bm[3 * a + bb]; }
          M[3][3] = D[0]; M[3][4] = D[1]; M[3][5] = D[2]; M[4][4] = D[3]; M[4][5] = D[4]; M[5][5] = D[5];
          M[4][3] = D[1]; M[5][3] = D[2]; M[5][4] = D[4];
          solve6_spd_neg(M, pA, a0);
        } else {
          a0[0] = T(0); a0[1] = T(0); a0[2] = -P.g; a0[3] = T(0); a0[4] = T(0); a0[5] = T(0);
        }
        stn<6>(ws + O_V, a0);
      }

      B200SIM_PHASE_MARK(9);
      // ========================================================= phase 6: ABA pass 3
      auto pass3_link = [&](const int i, const int par, const int jt) {
          T* ri = ws + (size_t)i * REC;
          const T* rp = ws + (size_t)par * REC;
          T ap[6], r[3], cI[6], U[6], aw[3];
          ldn<6>(rp + O_V, ap);
          ldn<3>(ri + O_RR, r);
          ldn<6>(ri + O_C, cI);
          ldn<6>(ri + O_U, U);
          ldn<3>(ri + O_AX, aw);
          T a[6];
          cross3(ap + 3, r, a);
          a[0] += ap[0] + cI[0]; a[1] += ap[1] + cI[1]; a[2] += ap[2] + cI[2];
          a[3] = ap[3] + cI[3]; a[4] = ap[4] + cI[4]; a[5] = ap[5] + cI[5];
          const T sdd = (ri[O_UU] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * ri[O_DINV];
          if (jt == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
          else if (jt == 2) { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
          stn<6>(ri + O_V, a);
          ri[O_SDD] = sdd;
      };
      if (SPEC) {
        int e = nrows > 0 ? rows[rlane] : 0xFF;
        for (int r = 0; r < nrows; ++r) {
          __syncwarp();
          const int en = (r + 1 < nrows) ? rows[(r + 1) * G + rlane] : 0xFF;
          if ((e & 0xFF) != 0xFF) pass3_link(e & 0xFF, (e >> 8) & 0xFF, (e >> 27) & 3);
          e = en;
        }
      } else {
        for (int l = 1; l <= P.depth; ++l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            pass3_link(i, parent[i], jtypes[i]);
          }
        }
      }
      __syncwarp();

      B200SIM_PHASE_MARK(10);
      // base acceleration in inertial-fixed representation + gravity (rbda/aba.py:284-288)
      if (floating) {
        T a0[6];
        ldn<6>(ws + O_V, a0);
        cross3(b.p, a0 + 3, Wa);
        Wa[0] += a0[0]; Wa[1] += a0[1]; Wa[2] += a0[2] + P.g;
        Wa[3] = a0[3]; Wa[4] = a0[4]; Wa[5] = a0[5];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) Wa[k] = T(0);
      }
      if (mode != MODE_STEP) break;

      // ========================================================= phase 7: semi-implicit Euler
      // (api/integrators.py:14-88), base part replicated in every lane
      {
        BaseState<T> nb;
#pragma unroll
        for (int k = 0; k < 3; ++k) { nb.vlin[k] = b.vlin[k] + dt * Wa[k]; nb.w[k] = b.w[k] + dt * Wa[3 + k]; }
        T pd[3];
        cross3(nb.w, b.p, pd);
        pd[0] += nb.vlin[0]; pd[1] += nb.vlin[1]; pd[2] += nb.vlin[2];
        // Quaternion.derivative (math/quaternion.py:68-132), inertial-fixed omega, K = 0.1
        const T nw = sqrt_t(dot3(nb.w, nb.w));
        const T nq = sqrt_t(b.qn[0] * b.qn[0] + b.qn[1] * b.qn[1] + b.qn[2] * b.qn[2] + b.qn[3] * b.qn[3]);
        const T v0 = T(0.1) * nw * (T(1) - nq);
        const T qw = b.qn[0], qx = b.qn[1], qy = b.qn[2], qz = b.qn[3];
        const T wx = nb.w[0], wy = nb.w[1], wz = nb.w[2];
        T qd[4];
        qd[0] = T(0.5) * (qw * v0 - qx * wx - qy * wy - qz * wz);
        qd[1] = T(0.5) * (qx * v0 + qw * wx + qz * wy - qy * wz);
        qd[2] = T(0.5) * (qy * v0 - qz * wx + qw * wy + qx * wz);
        qd[3] = T(0.5) * (qz * v0 + qy * wx - qx * wy + qw * wz);
        T qn2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) qn2[k] = b.qn[k] + dt * qd[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) nb.p[k] = b.p[k] + dt * pd[k];
        // normalise (integrators.py:61-63) and again in data.replace (api/data.py:441-447)
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const T nn = sqrt_t(qn2[0] * qn2[0] + qn2[1] * qn2[1] + qn2[2] * qn2[2] + qn2[3] * qn2[3]);
          const T inv = rcp_t((nn == T(0)) ? T(1) : nn);
#pragma unroll
          for (int k = 0; k < 4; ++k) qn2[k] *= inv;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) nb.qn[k] = qn2[k];
        quat_to_dcm(nb.qn, nb.R);
        b = nb;
      }
      B200SIM_PHASE_MARK(11);
      const bool want_caches = last && (P.W_H_L || P.W_v);
      const bool compact = SPEC && last && tma && (P.iXl != nullptr) && (nL <= 4 * G + 1);
      if (last && active && lane == 0) {
        stn<4>(P.q_o + env * 4, b.qn);
        stn<3>(P.p_o + env * 3, b.p);
        stn<3>(P.vlin_o + env * 3, b.vlin);
        stn<3>(P.omega_o + env * 3, b.w);
        if (P.W_H_B) store_transform(P.W_H_B + env * 16, b.R, b.p);
        if (P.iXl && !compact) {
          T R0[9], p0[3], t[3], X[36];
          mat3_mul(b.R, sm_cst + C_M0, R0);
          mat3_vec(b.R, sm_cst + C_TPRE, t);
          p0[0] = b.p[0] + t[0]; p0[1] = b.p[1] + t[1]; p0[2] = b.p[2] + t[2];
          inverse_adjoint(X, R0, p0);
          stg_vec<36>(P.iXl + env * nL * 36, X);
        }
      }
      if (compact) {
        // ======================================================= compact final phase (see FS above)
        B200SIM_PHASE_MARK(12);
        __pipeline_wait_prior(0);
        T snr[4], sdr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int i = 1 + lane + t * G;
          snr[t] = T(0); sdr[t] = T(0);
          if (i < nL) {
            const T* ri = ws + (size_t)i * REC;
            const T sdn = ri[O_SD] + dt * ri[O_SDD];
            const T sn = ri[O_S] + dt * sdn;
            snr[t] = sn; sdr[t] = sdn;
            if (active) {
              P.sd_o[env * n + (i - 1)] = sdn;
              P.s_o[env * n + (i - 1)] = sn;
            }
          }
        }
        __syncwarp();  // every lane is done with the records (joint state, base acceleration)
        T* FX = ws + (((size_t)nL * FS + 3) & ~size_t(3));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int i = 1 + lane + t * G;
          if (i < nL) {
            T Rrel[9], trel[3], X[36];
            joint_rel_transform(P, flags, sm_cst + (size_t)i * CREC, jtypes[i], i, snr[t], Rrel, trel);
            T* fi = ws + (size_t)i * FS;
            stn<9>(fi + F_R, Rrel);
            stn<3>(fi + F_P, trel);
            fi[F_SD] = sdr[t];
            inverse_adjoint(X, Rrel, trel);
            stn<36>(FX + (size_t)i * 36, X);
          }
        }
        if (lane == 0) {
          // link 0: Ad((W_H_B suc_H_i[0])^-1) (api/kin_dyn_parameters.py:417-449) and the chain root
          T R0[9], p0[3], t[3], X[36];
          mat3_mul(b.R, sm_cst + C_M0, R0);
          mat3_vec(b.R, sm_cst + C_TPRE, t);
          p0[0] = b.p[0] + t[0]; p0[1] = b.p[1] + t[1]; p0[2] = b.p[2] + t[2];
          inverse_adjoint(X, R0, p0);
          stn<36>(FX, X);
          T v0[6];
          cross3(b.w, b.p, t);  // velocity of the base origin: v_lin + w x p
          v0[0] = b.vlin[0] + t[0]; v0[1] = b.vlin[1] + t[1]; v0[2] = b.vlin[2] + t[2];
          v0[3] = b.w[0]; v0[4] = b.w[1]; v0[5] = b.w[2];
          stn<9>(ws + F_R, b.R);
          stn<3>(ws + F_P, b.p);
          stn<6>(ws + F_V, v0);
        }
        // generic-proxy writes of every lane -> visible to the async proxy, then one bulk store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && active) tma_store_bulk(P.iXl + env * nL * 36, FX, (unsigned)(nL * 36 * sizeof(T)));
        B200SIM_PHASE_MARK(13);
        if (want_caches) {
          // FK + velocity chain over the tree levels on the compact records
          int e = nrows > 0 ? rows[rlane] : 0xFF;
          for (int r = 0; r < nrows; ++r) {
            __syncwarp();
            const int en = (r + 1 < nrows) ? rows[(r + 1) * G + rlane] : 0xFF;
            if ((e & 0xFF) != 0xFF) {
              const int i = e & 0xFF;
              const T* rp = ws + (size_t)((e >> 8) & 0xFF) * FS;
              T* ri = ws + (size_t)i * FS;
              T Rp[9], pp[3], vp[6], Rrel[9], trel[3];
              ldn<9>(rp + F_R, Rp);
              ldn<3>(rp + F_P, pp);
              ldn<6>(rp + F_V, vp);
              ldn<9>(ri + F_R, Rrel);
              ldn<3>(ri + F_P, trel);
              T R[9], r[3], pw[3];
              mat3_mul(Rp, Rrel, R);
              mat3_vec(Rp, trel, r);
              pw[0] = pp[0] + r[0]; pw[1] = pp[1] + r[1]; pw[2] = pp[2] + r[2];
              T ax[3], aw[3];
              ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
              mat3_vec(R, ax, aw);
              const T sdi = ri[F_SD];
              T v[6];
              cross3(vp + 3, r, v);
              v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
              v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
              const int jt = (e >> 27) & 3;
              if (jt == 1) { v[3] += sdi * aw[0]; v[4] += sdi * aw[1]; v[5] += sdi * aw[2]; }
              else if (jt == 2) { v[0] += sdi * aw[0]; v[1] += sdi * aw[1]; v[2] += sdi * aw[2]; }
              stn<9>(ri + F_R, R);
              stn<3>(ri + F_P, pw);
              stn<6>(ri + F_V, v);
            }
            e = en;
          }
          __syncwarp();
          B200SIM_PHASE_MARK(14);
          if (active) {
            for (int i = lane; i < nL; i += G) {
              const T* ri = ws + (size_t)i * FS;
              T R[9], p[3], v[6];
              ldn<9>(ri + F_R, R);
              ldn<3>(ri + F_P, p);
              ldn<6>(ri + F_V, v);
              if (P.W_H_L) store_transform(P.W_H_L + (env * nL + i) * 16, R, p);
              if (P.W_v) {
                T t[3];
                cross3(p, v + 3, t);  // inertial-fixed linear part: vlin + p x w
                const T o[6] = {v[0] + t[0], v[1] + t[1], v[2] + t[2], v[3], v[4], v[5]};
                stg_vec6(P.W_v + (env * nL + i) * 6, o);
              }
            }
          }
          B200SIM_PHASE_MARK(15);
        }
        __syncwarp();
      } else {
        // joints: new velocity/position (+ joint transforms of the new state when kinematics
        // are needed again: next step, or cache outputs)
        B200SIM_PHASE_MARK(12);
        const bool need_fk = !last || want_caches || (P.iXl != nullptr);
        __pipeline_wait_prior(0);  // next step's torque references have landed
        for (int i = 1 + lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * REC;
          const T sdn = ri[O_SD] + dt * ri[O_SDD];
          const T sn = ri[O_S] + dt * sdn;
          ri[O_SD] = sdn;
          ri[O_S] = sn;
          if (last && active) {
            P.sd_o[env * n + (i - 1)] = sdn;
            P.s_o[env * n + (i - 1)] = sn;
          }
          if (need_fk) {
            T Rrel[9], trel[3];
            joint_rel_transform(P, flags, sm_cst + (size_t)i * CREC, jtypes[i], i, sn, Rrel, trel);
            stn<9>(ri + O_R, Rrel);
            stn<3>(ri + O_P, trel);
            if (last && active && P.iXl) emit_joint_adjoint(ri, i, Rrel, trel);
          }
        }
        B200SIM_PHASE_MARK(13);
        // ========================================================= phase 8: FK of the new state
        if (!last) {
          // the next fused step starts like a fresh call: base_orientation normalises the
          // stored quaternion once more (api/data.py:283-285), bit-identical to repeated steps
          const T nrm = sqrt_t(b.qn[0] * b.qn[0] + b.qn[1] * b.qn[1] + b.qn[2] * b.qn[2] + b.qn[3] * b.qn[3]);
          const T inv = rcp_t(nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));
  #pragma unroll
          for (int k = 0; k < 4; ++k) b.qn[k] *= inv;
          quat_to_dcm(b.qn, b.R);
        }
        if (flags & F_GENERIC_FK) make_fk_map(sm_cst, b, fm);
        __syncwarp();  // every lane has read the base acceleration out of record 0
        if (!last || want_caches || rigid) {
          write_base_record(b);
          fk_chain(!last);
          B200SIM_PHASE_MARK(14);
          if (last && active && want_caches) write_fk_caches(b, fm);
          B200SIM_PHASE_MARK(15);
        }
      }
      if (rigid && P.contact_model == 2) {
        // a point below the ground at t+dt: the impact (rigid.py:385-436) is left to the rigid
        // kernel, which starts from the pre-impact result this kernel has just stored
        // (the relaxed-rigid model has no impact step, relaxed_rigid.py:262-281)
        bool touch = false;
        for (int k = lane; k < nc; k += G) {
          const T* rb = ws + (size_t)pt_body[k] * REC;
          const T* Lp = sm_pt + 3 * k;
          const T z = rb[O_P + 2] + rb[O_R + 6] * Lp[0] + rb[O_R + 7] * Lp[1] + rb[O_R + 8] * Lp[2];
          touch = touch || (pt_enabled[k] && (P.h_terrain - z > T(0)));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, touch);
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
        if ((bal & gmask) && active && lane == 0) over_push(P, (int)env, 1);
      }
      __syncwarp();
    }  // steps

    if (mode != MODE_STEP) {
      if (active) {
        if (lane == 0) stn<6>(P.avd + env * 6, Wa);
        if (lane == 0 && mode == MODE_DYN) {
          // system_position_dynamics (api/ode.py:134-171): Baumgarte K = 1.0
          T pd[3];
          cross3(b.w, b.p, pd);
          pd[0] += b.vlin[0]; pd[1] += b.vlin[1]; pd[2] += b.vlin[2];
          const T nw = sqrt_t(dot3(b.w, b.w));
          const T nq = sqrt_t(b.qn[0] * b.qn[0] + b.qn[1] * b.qn[1] + b.qn[2] * b.qn[2] + b.qn[3] * b.qn[3]);
          const T v0 = nw * (T(1) - nq);
          const T qw = b.qn[0], qx = b.qn[1], qy = b.qn[2], qz = b.qn[3];
          const T wx = b.w[0], wy = b.w[1], wz = b.w[2];
          const T qd[4] = {T(0.5) * (qw * v0 - qx * wx - qy * wy - qz * wz), T(0.5) * (qx * v0 + qw * wx + qz * wy - qy * wz),
                           T(0.5) * (qy * v0 - qz * wx + qw * wy + qx * wz), T(0.5) * (qz * v0 + qy * wx - qx * wy + qw * wz)};
          if (P.p_o) stn<3>(P.p_o + env * 3, pd);
          if (P.q_o) stn<4>(P.q_o + env * 4, qd);
        }
        for (int i = 1 + lane; i < nL; i += G) P.sdd_o[env * n + (i - 1)] = ws[(size_t)i * REC + O_SDD];
      }
      __syncwarp();
    }
  }
  if (tma) tma_store_wait_all();
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[8 + 16] = (unsigned long long)clock64();
  if (P.dbg && blockIdx.x < 512) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.dbg[40 + 2 * blockIdx.x + 1] = t;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      P.dbg[40 + 1024 + blockIdx.x] = smid;
    }
  }
}

}  // namespace b200sim
