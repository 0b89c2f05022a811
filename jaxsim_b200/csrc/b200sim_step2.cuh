// b200sim_step2.cuh -- second-generation fused step kernel (the hot path of every BASELINE
// soft-contact configuration): `jaxsim.api.model.step` of a floating-base URDF model with
// SoftContacts (or no collidable points) and SemiImplicitEuler, api/model.py:2601-2681.
//
// Same mathematics as step_kernel<T,G,1> (b200sim_kernels.cuh: G lanes per environment, ABA in
// world-aligned link-origin frames, level walks over packed rows).  What changed is dictated by
// what bounds the step on B200: at large batch the LSU data pipe of the SM (shared-memory +
// global wavefronts, 1 per cycle) runs at 76-83 % (profiles/r02_step2_lsu_bound.md), at batch
// 4096 it is the latency of one environment through one warp.  So this kernel minimises
// WAVEFRONTS per environment:
//
//  * Two lane mappings inside a warp.  Link-/point-parallel phases use G consecutive lanes per
//    environment (all lanes busy).  The level walks (ABA pass 2 / 3, FK) only have 3-5 links per
//    tree level: they use the TRANSPOSED mapping lane = slot * (32/G) + environment, so the busy
//    lanes of the 32/G environments of a warp are packed into the low quarter-warps and a 128-bit
//    shared-memory access costs 2 wavefronts instead of 4.  The two mappings exchange data through
//    the environments' shared-memory records only, separated by __syncwarp().
//  * 44-word link record instead of 60: in pass 2 the child ADDS its shifted articulated inertia
//    into the parent's record (ordered sub-rounds for siblings of one row: deterministic, race
//    free) instead of the parent gathering its children, so U, 1/d, u reuse the link's own dead
//    I^A slots; the spatial acceleration reuses the slot of c_i.  Every group of fields that is
//    read together starts on a 16-byte boundary: whole record rows move with LDS.128 / STS.128.
//  * Link poses are kept as 3x4 rows [R | p] like the (B,nL,4,4) leaves, so the cached input rows
//    are copied, not re-packed, and the cache outputs of the last step are staged in shared
//    memory in EXACTLY the layout of the output leaves -- W_H_L (nL,4,4), W_v (nL,6) and the joint
//    adjoints (nL,6,6) -- and leave with three cp.async.bulk (TMA) per environment: no global
//    store wavefronts (the per-link 128-bit stores of the 6x6 adjoints alone cost 200 L1 tag
//    requests per environment).  Inputs arrive with two cp.async.bulk per environment.
//  * One code path for "kinematics of a joint state" (joint transforms + FK walk) with run-time
//    strides: it serves the un-cached start, the steps of a fused rollout (records) and the cache
//    outputs of the last step (output-layout staging).
#pragma once

#include "b200sim_kernels.cuh"

namespace b200sim {

constexpr int R2 = 44;  // words per link record (44 = 4*11: 128-bit rows of consecutive links hit distinct banks)
// kinematics view of the record
constexpr int K_H = 0;    // 12 rows [R | p] of the world pose (relative pose before the FK walk)
constexpr int K_V = 12;   // 6  velocity of the link origin, world axes (lin, ang); word 0 carries sd before the walk
// ABA view: articulated inertia [[A,B],[B^T,D]] + bias force, children add into it
constexpr int K_IA = 0;   // 6 (sym)
constexpr int K_IB = 6;   // 9
constexpr int K_ID = 15;  // 6 (sym)
constexpr int K_PA = 21;  // 6
constexpr int K_TAU = 27; // 1  resultant joint torque (rides in the last row of the bias force)
// after the link's own pass-2 row its I^A is dead: U, 1/d, u live there until the next FK
constexpr int K_U = 0;    // 6
constexpr int K_DINV = 6;
constexpr int K_UU = 7;
constexpr int K_C = 28;   // 6  c_i, overwritten by the spatial acceleration a_i in pass 3
constexpr int K_AX = 34;  // 3  joint axis, world axes
constexpr int K_RR = 37;  // 3  p_i - p_parent, world axes
constexpr int K_S = 40, K_SD = 41, K_SDD = 42, K_TREF = 43;

constexpr int S2_NT = 4;  // link trips of the unrolled phases: nL <= 4 G

// shared-memory words of T per environment: the records + point records of the dynamics, or the
// output staging of the last step ([nL x 16 | nL x 6 | nL x 36]), whichever is larger (the mbarriers of the bulk
// loads live in a static shared array); padded so that consecutive environments of a warp sit 2 or 6 (G = 8) / 4 (G = 16) 16-byte rows apart
// modulo 8: the transposed walks then read 128-bit rows of neighbouring links without bank conflicts
__host__ __device__ inline size_t env2_ws_words(size_t ts, int nL, int nc, int G) {
  size_t w = (size_t)nL * R2 + (size_t)nc * PTREC;
  const size_t f = (size_t)nL * 16 + (((size_t)nL * 6 + 3) & ~size_t(3)) + (size_t)nL * 36;
  if (f > w) w = f;
  w = (w + 3) & ~size_t(3);
  const size_t per_row = 16 / ts;
  while (G == 16 ? ((w / per_row) % 8 != 4) : ((w / per_row) % 4 != 2)) w += per_row;
  return w;
}

// whole 16-byte rows of shared memory (p 16-byte aligned; N * sizeof(T) a multiple of 16)
template <int N>
__device__ __forceinline__ void ldv(const float* p, float* d) {
  static_assert(N % 4 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p + k);
    d[k] = v.x; d[k + 1] = v.y; d[k + 2] = v.z; d[k + 3] = v.w;
  }
}
template <int N>
__device__ __forceinline__ void ldv(const double* p, double* d) {
  static_assert(N % 2 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 2) {
    const double2 v = *reinterpret_cast<const double2*>(p + k);
    d[k] = v.x; d[k + 1] = v.y;
  }
}
template <int N>
__device__ __forceinline__ void stv(float* p, const float* d) {
  static_assert(N % 4 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 4) *reinterpret_cast<float4*>(p + k) = make_float4(d[k], d[k + 1], d[k + 2], d[k + 3]);
}
template <int N>
__device__ __forceinline__ void stv(double* p, const double* d) {
  static_assert(N % 2 == 0, "N");
#pragma unroll
  for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(p + k) = make_double2(d[k], d[k + 1]);
}
// six consecutive words at an 8-byte (float) / 16-byte (double) boundary
__device__ __forceinline__ void ld6(const float* p, float* d) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) {
    const float2 v = *reinterpret_cast<const float2*>(p + k);
    d[k] = v.x; d[k + 1] = v.y;
  }
}
__device__ __forceinline__ void ld6(const double* p, double* d) { ldv<6>(p, d); }
__device__ __forceinline__ void st6(float* p, const float* d) {
#pragma unroll
  for (int k = 0; k < 6; k += 2) *reinterpret_cast<float2*>(p + k) = make_float2(d[k], d[k + 1]);
}
__device__ __forceinline__ void st6(double* p, const double* d) { stv<6>(p, d); }

// rows [R | p] (3x4) <-> R (row-major 3x3), p
template <typename T>
__device__ __forceinline__ void split_pose(const T* H, T* R, T* p) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    R[3 * r] = H[4 * r]; R[3 * r + 1] = H[4 * r + 1]; R[3 * r + 2] = H[4 * r + 2];
    p[r] = H[4 * r + 3];
  }
}
template <typename T>
__device__ __forceinline__ void st_pose(T* dst, const T* R, const T* p) {
  const T H[12] = {R[0], R[1], R[2], p[0], R[3], R[4], R[5], p[1], R[6], R[7], R[8], p[2]};
  stv<12>(dst, H);
}

#define B200SIM_MARK2(k)                                                                                       \
  do {                                                                                                         \
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0 && env0 == first) P.dbg[8 + (k)] = (unsigned long long)clock64(); \
  } while (0)

// packed rows of the level walks (b200sim_model_create):
//   bits 0-7 link (0xFF: none) | 8-15 parent | 16-19 rank among the siblings of this row | 20-23 sub-rounds of the row |
//   27-28 joint type
template <typename T, int G, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) step2_kernel(const Params<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long sm_mbar[THREADS / G];  // one mbarrier per environment slot (bulk input loads)
  constexpr int W4 = 32 / G;  // environments per warp
  T* sm_cst = reinterpret_cast<T*>(smem_raw);
  const int nL = P.nL, n = P.n, nc = P.nc;
  T* sm_pt = sm_cst + (size_t)nL * CREC;
  const size_t pt_words = ((size_t)nc * 3 + 3) & ~size_t(3);
  int* sm_itab = reinterpret_cast<int*>(sm_pt + pt_words);
  const size_t itab_words = ((size_t)P.itab_words + 3) & ~size_t(3);
  T* ws_base = reinterpret_cast<T*>(sm_itab + itab_words);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[8 + 0] = (unsigned long long)clock64();
  if (P.dbg && threadIdx.x == 0 && blockIdx.x < 512) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.dbg[40 + 2 * blockIdx.x] = t;
  }
  stage_async(sm_cst, P.cst, nL * CREC);
  stage_async(sm_pt, P.pt_pos, (int)pt_words);
  stage_async(sm_itab, P.itab, (int)itab_words);
  __pipeline_commit();

  const int lane32 = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // link-parallel mapping: G consecutive lanes per environment
  const int lane = lane32 & (G - 1);
  const int grp = warp * W4 + lane32 / G;
  // walk mapping (transposed): lane32 = slot * W4 + environment
  const int slot = lane32 / W4;
  const int wgrp = warp * W4 + (lane32 & (W4 - 1));
  const size_t wsw = (size_t)P.ws2_words;
  T* ws = ws_base + (size_t)grp * wsw;
  T* wk = ws_base + (size_t)wgrp * wsw;
  T* ptws = ws + (size_t)nL * R2;
  unsigned long long* mbar = sm_mbar + grp;
  if (lane == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // everything above touched only the constant model; the state may have been written by the previous
  // launch on the stream (programmatic dependent launch): wait for it to complete and flush
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0) P.dbg[8 + 1] = (unsigned long long)clock64();

  const int* parent = sm_itab + P.o_parent;
  const int* jtypes = sm_itab + P.o_jtype;
  const int* pt_start = sm_itab + P.o_pt_start;
  const int* pt_idx = sm_itab + P.o_pt_idx;
  const int* pt_body = sm_itab + P.o_pt_body;
  const int* pt_enabled = sm_itab + P.o_pt_enabled;
  const int* rows = sm_itab + (G == 16 ? P.o_rows2_16 : P.o_rows2_8);
  const int nrows = (G == 16 ? P.n_rows16 : P.n_rows8);

  const long long stride = (long long)gridDim.x * P.envs_per_block;
  const T dt = P.dt;
  const int flags = P.flags;
  const bool use_cached = P.Hin && P.Vin;
  const bool bulk = use_cached && (flags & F_BULK_IN);
  const int VW = (nL * 6 + 3) & ~3;  // input staging: [V: nL x 6, padded][H: nL x 16]
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane32 / G * G));
  // output staging of the last step, in the layout of the output leaves
  const int o_FV = nL * 16, o_FX = nL * 16 + VW;

  const long long first = (long long)blockIdx.x * P.envs_per_block;
  unsigned in_parity = 0;
  for (long long env0 = first; env0 < P.B; env0 += stride) {
    long long env = env0 + grp;
    const bool active = env < P.B;
    if (!active) env = env % P.B;  // idle groups shadow distinct valid environments, stores masked

    // =========================================================== inputs (one burst)
    if (lane == 0) tma_store_wait_read();  // the bulk stores of the previous environment have read its staging
    if (bulk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses of the last trip before the async writes
      __syncwarp();
      if (lane == 0) {
        const unsigned bH = (unsigned)(nL * 16 * sizeof(T)), bV = (unsigned)(nL * 6 * sizeof(T));
        mbar_expect_tx(mbar, bH + bV);
        tma_load_bulk(ws, P.Vin + env * nL * 6, bV, mbar);
        tma_load_bulk(ws + VW, P.Hin + env * nL * 16, bH, mbar);
      }
    } else if (use_cached) {
      __syncwarp();
      constexpr int per = 16 / sizeof(T);
#pragma unroll
      for (int t = 0; t < S2_NT; ++t) {
        const int i = lane + t * G;
        if (i < nL) {
          const T* H = P.Hin + (env * nL + i) * 16;
          const T* V = P.Vin + (env * nL + i) * 6;
#pragma unroll
          for (int k = 0; k < 12; k += per) __pipeline_memcpy_async(ws + VW + (size_t)i * 16 + k, H + k, 16);
          if (sizeof(T) == 4) {
#pragma unroll
            for (int k = 0; k < 6; k += 2) __pipeline_memcpy_async(ws + (size_t)i * 6 + k, V + k, 8);
          } else {
#pragma unroll
            for (int k = 0; k < 6; k += per) __pipeline_memcpy_async(ws + (size_t)i * 6 + k, V + k, 16);
          }
        }
      }
    } else {
      __syncwarp();
    }
    T s_r[S2_NT], sd_r[S2_NT], tr_r[S2_NT];
#pragma unroll
    for (int t = 0; t < S2_NT; ++t) {
      const int i = lane + t * G;
      s_r[t] = T(0); sd_r[t] = T(0); tr_r[t] = T(0);
      if (i >= 1 && i < nL) {
        s_r[t] = P.s[env * n + (i - 1)];
        sd_r[t] = P.sd[env * n + (i - 1)];
        if (P.tau) tr_r[t] = P.tau[env * n + (i - 1)];
      }
    }
    for (int k = lane; k < nc; k += G) {
      T* pw = ptws + (size_t)k * PTREC + PT_M;
      if (P.m) {
        const T* src = P.m + (env * nc + k) * 3;
        cp_async_elem(pw, src); cp_async_elem(pw + 1, src + 1); cp_async_elem(pw + 2, src + 2);
      } else {
        pw[0] = T(0); pw[1] = T(0); pw[2] = T(0);
      }
    }
    __pipeline_commit();
    B200SIM_MARK2(2);

    // =========================================================== base state (replicated per lane)
    BaseState<T> b;
    {
      const T* q = P.q + env * 4;
      T qr[4] = {q[0], q[1], q[2], q[3]};
      ldn<3>(P.p + env * 3, b.p);
      ldn<3>(P.vlin + env * 3, b.vlin);
      ldn<3>(P.omega + env * 3, b.w);
      const T nrm = sqrt_t(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const T den = nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0));  // base_orientation (api/data.py:283-285)
      const T inv = rcp_t(den);
#pragma unroll
      for (int k = 0; k < 4; ++k) b.qn[k] = qr[k] * inv;
      quat_to_dcm(b.qn, b.R);
    }
    B200SIM_MARK2(3);
    __pipeline_wait_prior(0);
    if (env0 == first) __syncthreads();  // the model blob staged by all threads of the block
    B200SIM_MARK2(4);

    if (use_cached) {
      // ========================================================= kinematics of the input state from its caches
      // (what the reference's contact code reads, api/contact.py:39-43).  The records (44 words per link)
      // overwrite the staging area (22 words per link) from the top: trips over DESCENDING link indices, each
      // reads its links' staged rows (and the parent's position), synchronises, then writes their records.
      if (bulk) {
        mbar_wait(mbar, in_parity);
        in_parity ^= 1u;
      } else {
        __syncwarp();  // rows staged by other lanes' cp.async
      }
#pragma unroll
      for (int t = S2_NT - 1; t >= 0; --t) {
        if (t * G < nL) {
          const int i = lane + t * G;
          T H[12], V[6], pp[3];
          if (i < nL) {
            ldv<12>(ws + VW + (size_t)i * 16, H);
            ld6(ws + (size_t)i * 6, V);
            if (i > 0) {
              const T* Hp = ws + VW + (size_t)parent[i] * 16;
              pp[0] = Hp[3]; pp[1] = Hp[7]; pp[2] = Hp[11];
            }
          }
          __syncwarp();
          if (i < nL) {
            T* ri = ws + (size_t)i * R2;
            const T p[3] = {H[3], H[7], H[11]};
            T tt[3];
            cross3(V + 3, p, tt);  // velocity of the link origin: W_v_lin + w x p
            stv<12>(ri + K_H, H);
            const T v8[8] = {V[0] + tt[0], V[1] + tt[1], V[2] + tt[2], V[3], V[4], V[5], T(0), T(0)};
            stv<8>(ri + K_V, v8);
            if (i > 0) {
              T ax[3], aw[3];
              ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
              aw[0] = H[0] * ax[0] + H[1] * ax[1] + H[2] * ax[2];
              aw[1] = H[4] * ax[0] + H[5] * ax[1] + H[6] * ax[2];
              aw[2] = H[8] * ax[0] + H[9] * ax[1] + H[10] * ax[2];
              const T Q[12] = {T(0), T(0), aw[0], aw[1], aw[2], p[0] - pp[0], p[1] - pp[1], p[2] - pp[2],
                               s_r[t], sd_r[t], T(0), tr_r[t]};
              stv<12>(ri + 32, Q);  // (c tail), axis, r, s, sd, (sdd), torque reference
            }
          }
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int t = 0; t < S2_NT; ++t) {
        const int i = lane + t * G;
        if (i >= 1 && i < nL) {
          const T Q[4] = {s_r[t], sd_r[t], T(0), tr_r[t]};
          stv<4>(ws + (size_t)i * R2 + K_S, Q);
        }
      }
    }
    B200SIM_MARK2(5);

    // step -1 (only without input caches) is the kinematics of the input state: it runs the tail of the
    // step body (joint transforms + FK walk) on the unchanged joint state
    for (int step = use_cached ? 0 : -1; step < P.nsteps; ++step) {
      const bool last = (step == P.nsteps - 1);
      T a0[6];
      if (step >= 0) {
        const T* fext_step = P.fext ? P.fext + (long long)step * P.fext_step_stride : nullptr;
        // ======================================================= contacts (point-parallel)
        // collidable_points_pos_vel + compute_penetration_data (FlatTerrain) + Hunt/Crossley
        // (rbda/collidable_points.py:9-65, rbda/contacts/common.py:25-63, rbda/contacts/soft.py:195-444)
        bool env_touches = false;
        for (int k = lane; k < nc; k += G) {
          const T* rb = ws + (size_t)pt_body[k] * R2;
          T Kb[20];
          ldv<20>(rb, Kb);
          const T* vl = Kb + K_V;
          const T* w = Kb + K_V + 3;
          T Lp[3], d[3], pd[3];
          ldn<3>(sm_pt + 3 * k, Lp);
          d[0] = Kb[0] * Lp[0] + Kb[1] * Lp[1] + Kb[2] * Lp[2];
          d[1] = Kb[4] * Lp[0] + Kb[5] * Lp[1] + Kb[6] * Lp[2];
          d[2] = Kb[8] * Lp[0] + Kb[9] * Lp[1] + Kb[10] * Lp[2];
          const T pcz = Kb[11] + d[2];
          cross3(w, d, pd);
          pd[0] += vl[0]; pd[1] += vl[1]; pd[2] += vl[2];
          T* pw = ptws + (size_t)k * PTREC;
          T m[3];
          ldn<3>(pw + PT_M, m);
          T f[3] = {T(0), T(0), T(0)};
          T md[3] = {T(0), T(0), T(0)};
          if (pt_enabled[k]) {
            const T delta = max_t(T(0), P.h_terrain - pcz);
            const T KoD = P.K * rcp_t(P.D);
            if (delta <= T(0)) {
              // no contact: zero force, the tangential deformation relaxes (soft.py:318-330)
              md[0] = -KoD * m[0]; md[1] = -KoD * m[1]; md[2] = -KoD * m[2];
            } else {
              const T ddot = -pd[2];
              const T eps = Lim<T>::eps();
              const T dp = (flags & F_SQRT_P) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.pexp);
              const T dq = (flags & F_SQRT_Q) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.qexp);
              const T Kdp = P.K * dp, Ddq = P.D * dq;
              const T fn = max_t(T(0), Kdp * delta + Ddq * ddot);
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
          stn<3>(pw + PT_F, f);
          stn<3>(pw + PT_LEV, d);  // lever arm w.r.t. the origin of the body link
          // m+ = m + dt * m_dot (api/integrators.py:67-71); stays on chip between fused steps
          m[0] += dt * md[0]; m[1] += dt * md[1]; m[2] += dt * md[2];
          stn<3>(pw + PT_M, m);
          if (last && active && P.m_o) {
            T* mo = P.m_o + (env * nc + k) * 3;
            mo[0] = m[0]; mo[1] = m[1]; mo[2] = m[2];
          }
        }
        // airborne environments skip the per-link accumulation of the (all zero) contact wrenches
        env_touches = (__ballot_sync(0xffffffffu, env_touches) & gmask) != 0u;
        __syncwarp();
        B200SIM_MARK2(6);

        // ======================================================= link-parallel: inertias, bias forces, actuation
        for (int i = lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * R2;
          const T* c = sm_cst + (size_t)i * CREC;
          T Kk[20];
          ldv<20>(ri, Kk);
          T R[9], p[3];
          split_pose(Kk, R, p);
          const T* v = Kk + K_V;
          T fe[3] = {T(0), T(0), T(0)}, ne[3] = {T(0), T(0), T(0)};
          if (env_touches) {
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
          if (fext_step) add_external_wrench(P.fext_repr, fext_step + (env * nL + i) * 6, R, p, fe, ne);
          // link inertia in world axes about the link origin
          T Cc[16];
          ldv<16>(c + C_MASS, Cc);  // mass, com, D_link, limit spring / damper, limits, friction
          const T mass = Cc[0];
          const T* com = Cc + (C_COM - C_MASS);
          const T* Dl = Cc + (C_DL - C_MASS);
          T cw[3];
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
          // pA = v x* (I v) - f_ext
          T pA[6];
          cross3(v + 3, fI, pA);
          cross3(v, fI, pA + 3);
          cross3_add(v + 3, nI, pA + 3);
          pA[0] -= fe[0]; pA[1] -= fe[1]; pA[2] -= fe[2];
          pA[3] -= ne[0]; pA[4] -= ne[1]; pA[5] -= ne[2];
          T tau_i = T(0);
          if (i > 0) {
            // c_i = v x vJ (rbda/aba.py:143-144) and the resultant joint torque (api/actuation_model.py:7-126)
            const int jt = jtypes[i];
            T Q[12];
            ldv<12>(ri + 32, Q);  // (c tail), axis, r, s, sd, (sdd), torque reference
            const T* aw = Q + (K_AX - 32);
            const T si = Q[K_S - 32], sdi = Q[K_SD - 32], tref = Q[K_TREF - 32];
            T cc[6];
            const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
            if (jt == 1) {
              cross3(v, vJ, cc);
              cross3(v + 3, vJ, cc + 3);
            } else {
              cross3(v + 3, vJ, cc);
              cc[3] = cc[4] = cc[5] = T(0);
            }
            {
              const T c8[8] = {cc[0], cc[1], cc[2], cc[3], cc[4], cc[5], aw[0], aw[1]};
              stv<8>(ri + K_C, c8);
            }
            const T* lc = Cc + (C_KS - C_MASS);  // KS KD SMIN SMAX KC KV
            const T lower = min_t(si - lc[2], T(0));
            const T upper = max_t(si - lc[3], T(0));
            T tlim = -lc[0] * (lower + upper);
            tlim = tlim - tlim * lc[1] * sdi;
            T tfr = T(0);
            if (P.enable_friction) {
              const T sg = (sdi > T(0)) ? T(1) : ((sdi < T(0)) ? T(-1) : T(0));
              tfr = -(lc[4] * sg + lc[5] * sdi);
            }
            const T tt = tref + tfr + tlim;
            const T av = abs_t(sdi);
            T lim;
            if (av <= P.w_th) lim = P.tau_max;
            else if (av <= P.w_max) lim = P.tau_max * (T(1) - (av - P.w_th) / (P.w_max - P.w_th));
            else lim = T(0);
            tau_i = min_t(max_t(tt, -lim), lim);
            // torque reference of the NEXT fused step: its latency hides behind the ABA passes
            if (!last && P.tau && P.tau_step_stride)
              cp_async_elem(ri + K_TREF, P.tau + (long long)(step + 1) * P.tau_step_stride + env * n + (i - 1));
          }
          // articulated inertia init (overwrites the pose and the velocity): A = m 1, B = -m S(c_w), D = D_w
          T IA[28];
          IA[0] = mass; IA[1] = T(0); IA[2] = T(0); IA[3] = mass; IA[4] = T(0); IA[5] = mass;
          IA[6] = T(0);            IA[7] = mass * cw[2];   IA[8] = -mass * cw[1];
          IA[9] = -mass * cw[2];   IA[10] = T(0);          IA[11] = mass * cw[0];
          IA[12] = mass * cw[1];   IA[13] = -mass * cw[0]; IA[14] = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) IA[15 + k] = Dw[k];
#pragma unroll
          for (int k = 0; k < 6; ++k) IA[21 + k] = pA[k];
          IA[27] = tau_i;
          stv<28>(ri, IA);
        }
        __pipeline_commit();
        B200SIM_MARK2(7);

        // ======================================================= ABA pass 2 (rbda/aba.py:184-234), leaves first
        // transposed lane mapping from here to the end of pass 3
        {
          int e = nrows > 0 ? rows[(nrows - 1) * G + slot] : 0xFF;
          for (int r = nrows - 1; r >= 0; --r) {
            __syncwarp();
            const int en = r > 0 ? rows[(r - 1) * G + slot] : 0xFF;
            const bool valid = (e & 0xFF) != 0xFF;
            const int nsub = (e >> 20) & 15;
            const int rank = (e >> 16) & 15;
            T X[27];  // contribution to the parent: A(6) B(9) D(6) pA(6), about the parent's origin
            T* rp = wk + (size_t)((e >> 8) & 0xFF) * R2;
            if (valid) {
              T* ri = wk + (size_t)(e & 0xFF) * R2;
              const int jt = (e >> 27) & 3;
              T W[40];
              ldv<40>(ri, W);
              T* A = W + K_IA;
              T* Bm = W + K_IB;
              T* D = W + K_ID;
              const T* pA = W + K_PA;
              const T* aw = W + K_AX;
              const T* cI = W + K_C;
              const T* rr = W + K_RR;
              const T tau = W[K_TAU];
              T Ul[3], Ua[3], d, u;
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
              {
                const T U8[8] = {Ul[0], Ul[1], Ul[2], Ua[0], Ua[1], Ua[2], dinv, u};
                stv<8>(ri + K_U, U8);
              }
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
              T* pa = X + 21;
              T t3[3];
              sym3_vec(A, cI, t3);
              pa[0] = pA[0] + t3[0] + Ul[0] * ud; pa[1] = pA[1] + t3[1] + Ul[1] * ud; pa[2] = pA[2] + t3[2] + Ul[2] * ud;
              mat3_vec(Bm, cI + 3, t3);
              pa[0] += t3[0]; pa[1] += t3[1]; pa[2] += t3[2];
              mat3T_vec(Bm, cI, t3);
              pa[3] = pA[3] + t3[0] + Ua[0] * ud; pa[4] = pA[4] + t3[1] + Ua[1] * ud; pa[5] = pA[5] + t3[2] + Ua[2] * ud;
              sym3_vec(D, cI + 3, t3);
              pa[3] += t3[0]; pa[4] += t3[1]; pa[5] += t3[2];
              // shift to the parent's origin, X = [[1, -S(r)],[0, 1]]:
              //   B'' = B' - A' S(r)   (row_i(A' S) = row_i(A') x r);   D'' = D' + S(r) B' + (S(r) B'')^T
              const T Af[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
              T* B2 = X + 6;
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                T rowx[3];
                cross3(Af + 3 * a, rr, rowx);
                B2[3 * a] = Bm[3 * a] - rowx[0]; B2[3 * a + 1] = Bm[3 * a + 1] - rowx[1]; B2[3 * a + 2] = Bm[3 * a + 2] - rowx[2];
              }
              T SB1[9], SB2[9];
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const T c1[3] = {Bm[j], Bm[3 + j], Bm[6 + j]};
                const T c2[3] = {B2[j], B2[3 + j], B2[6 + j]};
                T o1[3], o2[3];
                cross3(rr, c1, o1);
                cross3(rr, c2, o2);
                SB1[j] = o1[0]; SB1[3 + j] = o1[1]; SB1[6 + j] = o1[2];
                SB2[j] = o2[0]; SB2[3 + j] = o2[1]; SB2[6 + j] = o2[2];
              }
              X[15] = D[0] + SB1[0] + SB2[0];
              X[16] = D[1] + SB1[1] + SB2[3];
              X[17] = D[2] + SB1[2] + SB2[6];
              X[18] = D[3] + SB1[4] + SB2[4];
              X[19] = D[4] + SB1[5] + SB2[7];
              X[20] = D[5] + SB1[8] + SB2[8];
              cross3_add(rr, pa, pa + 3);
#pragma unroll
              for (int k = 0; k < 6; ++k) X[k] = A[k];
            }
            // children of one parent add in the order of their rank: deterministic and race free
            for (int k = 0; k < nsub; ++k) {
              if (k) __syncwarp();
              if (valid && rank == k) {
                T Y[28];
                ldv<28>(rp, Y);  // word 27 is the parent's joint torque: rewritten unchanged
#pragma unroll
                for (int q = 0; q < 27; ++q) Y[q] += X[q];
                stv<28>(rp, Y);
              }
            }
            e = en;
          }
          __syncwarp();
        }
        B200SIM_MARK2(8);

        // ======================================================= base acceleration (rbda/aba.py:240-242)
        if (slot == 0) {
          T W[28];
          ldv<28>(wk, W);
          const T* A = W + K_IA;
          const T* Bm = W + K_IB;
          const T* D = W + K_ID;
          const T* pA = W + K_PA;
          T M[6][6];
          M[0][0] = A[0]; M[0][1] = A[1]; M[0][2] = A[2]; M[1][1] = A[3]; M[1][2] = A[4]; M[2][2] = A[5];
          M[1][0] = A[1]; M[2][0] = A[2]; M[2][1] = A[4];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) { M[a][3 + bb] = Bm[3 * a + bb]; M[3 + bb][a] = Bm[3 * a + bb]; }
          M[3][3] = D[0]; M[3][4] = D[1]; M[3][5] = D[2]; M[4][4] = D[3]; M[4][5] = D[4]; M[5][5] = D[5];
          M[4][3] = D[1]; M[5][3] = D[2]; M[5][4] = D[4];
          T x0[6];
          solve6_spd_neg(M, pA, x0);
          const T a8[8] = {x0[0], x0[1], x0[2], x0[3], x0[4], x0[5], T(0), T(0)};
          stv<8>(wk + K_C, a8);
        }
        B200SIM_MARK2(9);

        // ======================================================= ABA pass 3 (rbda/aba.py:244-282), root first
        {
          int e = nrows > 0 ? rows[slot] : 0xFF;
          for (int r = 0; r < nrows; ++r) {
            __syncwarp();
            const int en = (r + 1 < nrows) ? rows[(r + 1) * G + slot] : 0xFF;
            if ((e & 0xFF) != 0xFF) {
              T* ri = wk + (size_t)(e & 0xFF) * R2;
              const T* rp = wk + (size_t)((e >> 8) & 0xFF) * R2;
              const int jt = (e >> 27) & 3;
              T ap[8], U[8], W[12];
              ldv<8>(rp + K_C, ap);  // parent's acceleration (+ two words of its axis)
              ldv<8>(ri + K_U, U);   // U, 1/d, u
              ldv<12>(ri + K_C, W);  // c, axis, r
              const T* cI = W;
              const T* aw = W + (K_AX - K_C);
              const T* rr = W + (K_RR - K_C);
              T a[8];
              cross3(ap + 3, rr, a);
              a[0] += ap[0] + cI[0]; a[1] += ap[1] + cI[1]; a[2] += ap[2] + cI[2];
              a[3] = ap[3] + cI[3]; a[4] = ap[4] + cI[4]; a[5] = ap[5] + cI[5];
              const T sdd = (U[7] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * U[6];
              if (jt == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
              else { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
              a[6] = aw[0]; a[7] = aw[1];
              stv<8>(ri + K_C, a);
              ri[K_SDD] = sdd;
            }
            e = en;
          }
          __pipeline_wait_prior(0);  // next step's torque references have landed
          __syncwarp();
        }
        B200SIM_MARK2(10);
        // back to the link-parallel mapping: the base acceleration of this lane's environment
        {
          T a8[8];
          ldv<8>(ws + K_C, a8);
#pragma unroll
          for (int k = 0; k < 6; ++k) a0[k] = a8[k];
        }

        // ======================================================= semi-implicit Euler, base (api/integrators.py:14-88)
        {
          // base acceleration in inertial-fixed representation + gravity (rbda/aba.py:284-288)
          T Wa[6];
          cross3(b.p, a0 + 3, Wa);
          Wa[0] += a0[0]; Wa[1] += a0[1]; Wa[2] += a0[2] + P.g;
          Wa[3] = a0[3]; Wa[4] = a0[4]; Wa[5] = a0[5];
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
        if (last && active && lane == 0) {
          stn<4>(P.q_o + env * 4, b.qn);
          stn<3>(P.p_o + env * 3, b.p);
          stn<3>(P.vlin_o + env * 3, b.vlin);
          stn<3>(P.omega_o + env * 3, b.w);
          if (P.W_H_B) store_transform(P.W_H_B + env * 16, b.R, b.p);
        }
        if (!last) {
          // the next fused step starts like a fresh call: base_orientation normalises the stored
          // quaternion once more (api/data.py:283-285), bit-identical to repeated steps
          const T nrm = sqrt_t(b.qn[0] * b.qn[0] + b.qn[1] * b.qn[1] + b.qn[2] * b.qn[2] + b.qn[3] * b.qn[3]);
          const T inv = rcp_t(nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));
#pragma unroll
          for (int k = 0; k < 4; ++k) b.qn[k] *= inv;
          quat_to_dcm(b.qn, b.R);
        }
        B200SIM_MARK2(11);
      }

      // ========================================================= joints: Euler update + joint transforms
      // (api/kin_dyn_parameters.py:396-451) of the state the FK walk below needs: the next fused step's (in the
      // records) or the new state's for the cache outputs (in the output staging, which aliases the records:
      // every lane first reads the joint state of its links, then all write)
      const bool want_caches = last && (P.W_H_L || P.W_v);
      const bool need_kin = !last || want_caches || (P.iXl != nullptr);
      const bool stage_out = last && need_kin;
      T snr[S2_NT], sdr[S2_NT];
#pragma unroll
      for (int t = 0; t < S2_NT; ++t) {
        const int i = lane + t * G;
        snr[t] = T(0); sdr[t] = T(0);
        if (i >= 1 && i < nL) {
          T* ri = ws + (size_t)i * R2;
          T Q[4];
          ldv<4>(ri + K_S, Q);  // s, sd, sdd, tref
          if (step >= 0) {
            Q[1] += dt * Q[2];
            Q[0] += dt * Q[1];
            if (!last) stv<4>(ri + K_S, Q);
            if (last && active) {
              P.sd_o[env * n + (i - 1)] = Q[1];
              P.s_o[env * n + (i - 1)] = Q[0];
            }
          }
          snr[t] = Q[0]; sdr[t] = Q[1];
        }
      }
      B200SIM_MARK2(12);
      if (need_kin) {
        // pose rows / velocity of link i: records, or the output staging of the last step
        const int ks = stage_out ? 16 : R2, vs = stage_out ? 6 : R2;
        T* kb = ws;
        T* vb = stage_out ? ws + o_FV : ws + K_V;
        if (stage_out) __syncwarp();  // every lane has read its joint state out of the records
        for (int t = 0; t * G < nL; ++t) {
          const int i = lane + t * G;
          const T sn = t == 0 ? snr[0] : (t == 1 ? snr[1] : (t == 2 ? snr[2] : snr[3]));
          const T sdn = t == 0 ? sdr[0] : (t == 1 ? sdr[1] : (t == 2 ? sdr[2] : sdr[3]));
          if (i >= 1 && i < nL) {
            T Rrel[9], trel[3];
            joint_rel_transform(P, 0, sm_cst + (size_t)i * CREC, jtypes[i], i, sn, Rrel, trel);
            st_pose(kb + (size_t)i * ks, Rrel, trel);
            vb[(size_t)i * vs] = sdn;
            if (stage_out) {
              const T row3[4] = {T(0), T(0), T(0), T(1)};
              stv<4>(kb + (size_t)i * 16 + 12, row3);
              if (P.iXl) {
                T X[36];
                inverse_adjoint(X, Rrel, trel);
                stv<36>(ws + o_FX + (size_t)i * 36, X);
              }
            }
          }
        }
        if (lane == 0) {
          // chain root: the base link (suc_H_i[0] = I for the models this kernel serves)
          T v0[6], t[3];
          cross3(b.w, b.p, t);  // velocity of the base origin: v_lin + w x p
          v0[0] = b.vlin[0] + t[0]; v0[1] = b.vlin[1] + t[1]; v0[2] = b.vlin[2] + t[2];
          v0[3] = b.w[0]; v0[4] = b.w[1]; v0[5] = b.w[2];
          st_pose(kb, b.R, b.p);
          st6(vb, v0);
          if (stage_out) {
            const T row3[4] = {T(0), T(0), T(0), T(1)};
            stv<4>(kb + 12, row3);
            if (P.iXl) {
              // index 0: Ad((W_H_B suc_H_i[0])^-1)  (api/kin_dyn_parameters.py:417-449)
              T X[36];
              inverse_adjoint(X, b.R, b.p);
              stv<36>(ws + o_FX, X);
            }
          }
        }
        if (stage_out && P.iXl) {
          // the (nL,6,6) joint adjoints of the environment leave with ONE bulk copy, overlapping the FK walk
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && active) tma_store_bulk(P.iXl + env * nL * 36, ws + o_FX, (unsigned)(nL * 36 * sizeof(T)));
        }
        B200SIM_MARK2(13);
        if (!last || want_caches) {
          // ===================================================== FK + velocity walk over the tree levels
          // (rbda/forward_kinematics.py:80-113, pass 1 of rbda/aba.py:131-171 in F_i coordinates), transposed mapping
          T* kbw = wk;
          T* vbw = stage_out ? wk + o_FV : wk + K_V;
          int e = nrows > 0 ? rows[slot] : 0xFF;
          for (int r = 0; r < nrows; ++r) {
            __syncwarp();
            const int en = (r + 1 < nrows) ? rows[(r + 1) * G + slot] : 0xFF;
            if ((e & 0xFF) != 0xFF) {
              const int i = e & 0xFF, par = (e >> 8) & 0xFF;
              T Hp[12], vp[6], Ho[12];
              ldv<12>(kbw + (size_t)par * ks, Hp);
              ld6(vbw + (size_t)par * vs, vp);
              ldv<12>(kbw + (size_t)i * ks, Ho);
              const T sdi = vbw[(size_t)i * vs];
              T Rp[9], pp[3], Rrel[9], trel[3];
              split_pose(Hp, Rp, pp);
              split_pose(Ho, Rrel, trel);
              T R[9], rr[3], pw[3];
              mat3_mul(Rp, Rrel, R);
              mat3_vec(Rp, trel, rr);
              pw[0] = pp[0] + rr[0]; pw[1] = pp[1] + rr[1]; pw[2] = pp[2] + rr[2];
              T ax[3], aw[3];
              ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
              mat3_vec(R, ax, aw);
              T v[6];
              cross3(vp + 3, rr, v);
              v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
              v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
              if (((e >> 27) & 3) == 1) { v[3] += sdi * aw[0]; v[4] += sdi * aw[1]; v[5] += sdi * aw[2]; }
              else { v[0] += sdi * aw[0]; v[1] += sdi * aw[1]; v[2] += sdi * aw[2]; }
              st_pose(kbw + (size_t)i * ks, R, pw);
              st6(vbw + (size_t)i * vs, v);
              if (!last) {
                const T Q[8] = {T(0), T(0), aw[0], aw[1], aw[2], rr[0], rr[1], rr[2]};
                stv<8>(wk + (size_t)i * R2 + 32, Q);  // (dead acceleration tail), axis, r
              }
            }
            e = en;
          }
          __syncwarp();
          B200SIM_MARK2(14);
          if (stage_out) {
            // link velocities in inertial-fixed representation, in place, then W_H_L and W_v leave with one
            // bulk copy each
            if (P.W_v) {
              for (int i = lane; i < nL; i += G) {
                T v[6];
                ld6(vb + (size_t)i * 6, v);
                const T* Hk = kb + (size_t)i * 16;
                const T p[3] = {Hk[3], Hk[7], Hk[11]};
                T t[3];
                cross3(p, v + 3, t);  // inertial-fixed linear part: vlin + p x w
                v[0] += t[0]; v[1] += t[1]; v[2] += t[2];
                st6(vb + (size_t)i * 6, v);
              }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && active) {
              if (P.W_H_L) tma_store_bulk(P.W_H_L + env * nL * 16, kb, (unsigned)(nL * 16 * sizeof(T)));
              if (P.W_v) tma_store_bulk(P.W_v + env * nL * 6, vb, (unsigned)(nL * 6 * sizeof(T)));
            }
          }
          B200SIM_MARK2(15);
        }
      }
    }  // steps
  }
  tma_store_wait_all();
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
