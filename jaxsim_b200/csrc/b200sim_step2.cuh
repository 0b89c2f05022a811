// b200sim_step2.cuh -- second-generation fused step kernel (the hot path of every BASELINE
// soft-contact configuration): `jaxsim.api.model.step` of a floating-base URDF model with
// SoftContacts (or no collidable points) and SemiImplicitEuler, api/model.py:2601-2681.
//
// Same mathematics and lane mapping as step_kernel<T,G,1> (b200sim_kernels.cuh: G lanes per
// environment, ABA in world-aligned link-origin frames, level walks over packed rows); what
// changes is everything that bounded that kernel on B200 (profiles/r01_step_kernel_v8_warm.md:
// 7 resident warps per SM, 28 environments in flight, 16 % of the stall samples waiting for
// instructions, 37 % in the output phase):
//
//  * 44-word link record instead of 60.  Pass 2 of the ABA no longer lets the parent GATHER its
//    children's shifted articulated inertias (which forced every link to keep I^A, p^A next to
//    U, 1/d, u): the child ADDS its contribution into the parent's record (ordered sub-rounds
//    for siblings of the same row: deterministic, race free), so U, 1/d, u reuse the link's own
//    dead I^A slots; the spatial acceleration reuses the slot of c_i; the velocity lives in the
//    I^A area while kinematics are needed.  4.8 KB per environment instead of 6.4 KB: 44-46
//    environments (11 warps) per SM instead of 28 (7 warps).
//  * <= 168 registers (384-thread launch bound), no 4x unrolled output phase: the SASS shrinks
//    from 4.8k to ~3k instructions (instruction fetch was 16 % of the stall samples).
//  * the cached kinematics of the input state ALWAYS arrive with two cp.async.bulk (TMA) per
//    environment when the rows are 16-byte granular, joint state / torques go to registers.
//  * one code path for "kinematics of a joint state" (joint transforms + FK level walk): it
//    serves the un-cached start, the steps of a fused rollout and the cache outputs of the last
//    step; the joint adjoints leave as 128-bit stores straight from registers (no staging area,
//    which is what capped the old workspace at 56 words per link in its final phase).
//  * the 6x6 floating-base solve runs redundantly in every lane of the group (same latency, no
//    shared-memory round trip of the base acceleration).
#pragma once

#include "b200sim_kernels.cuh"

namespace b200sim {

constexpr int R2 = 44;  // words per link record (44 = 4*11: 128-bit rows of consecutive links hit distinct banks)
// Every group of fields that is read together starts on a 128-bit boundary, so the level walks move
// whole record rows with LDS.128 / STS.128.
// kinematics view of the record
constexpr int K_R = 0;    // 9  world rotation (relative rotation before the FK walk)
constexpr int K_P = 9;    // 3  world position (relative translation before the FK walk)
constexpr int K_V = 12;   // 6  velocity of the link origin, world axes (lin, ang)
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
constexpr int K_S = 34, K_SD = 35;
constexpr int K_AX = 36;  // 3  joint axis, world axes
constexpr int K_SDD = 39;
constexpr int K_RR = 40;  // 3  p_i - p_parent, world axes
constexpr int K_TREF = 43;

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

constexpr int S2_NT = 4;  // link trips of the unrolled input phase: nL <= 4 G

template <typename T>
__host__ __device__ inline size_t env2_ws_words(int nL, int nc) {
  size_t w = (size_t)nL * R2 + (size_t)nc * PTREC;
  w = (w + 3) & ~size_t(3);
  return w + 16 / sizeof(T);  // + the mbarrier of the environment's bulk loads
}
template <typename T>
__device__ __forceinline__ unsigned long long* env2_mbar(T* ws, int nL, int nc) {
  return reinterpret_cast<unsigned long long*>(ws + env2_ws_words<T>(nL, nc) - 16 / sizeof(T));
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

  const int lane = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  const size_t wsw = env2_ws_words<T>(nL, nc);
  T* ws = ws_base + (size_t)grp * wsw;
  T* ptws = ws + (size_t)nL * R2;
  if (lane == 0) {
    mbar_init(env2_mbar(ws, nL, nc), 1);
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
  const int VW = (nL * 6 + 3) & ~3;  // staging: [V: nL x 6, padded][H: nL x 16]
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));

  const long long first = (long long)blockIdx.x * P.envs_per_block;
  unsigned in_parity = 0;
  for (long long env0 = first; env0 < P.B; env0 += stride) {
    long long env = env0 + grp;
    const bool active = env < P.B;
    if (!active) env = env % P.B;  // idle groups shadow distinct valid environments, stores masked

    // =========================================================== inputs (one burst)
    if (bulk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic accesses of the last trip before the async writes
      __syncwarp();
      if (lane == 0) {
        unsigned long long* bar = env2_mbar(ws, nL, nc);
        const unsigned bH = (unsigned)(nL * 16 * sizeof(T)), bV = (unsigned)(nL * 6 * sizeof(T));
        mbar_expect_tx(bar, bH + bV);
        tma_load_bulk(ws, P.Vin + env * nL * 6, bV, bar);
        tma_load_bulk(ws + VW, P.Hin + env * nL * 16, bH, bar);
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
        mbar_wait(env2_mbar(ws, nL, nc), in_parity);
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
            const T R[9] = {H[0], H[1], H[2], H[4], H[5], H[6], H[8], H[9], H[10]};
            const T p[3] = {H[3], H[7], H[11]};
            T v[6], tt[3];
            cross3(V + 3, p, tt);  // velocity of the link origin: W_v_lin + w x p
            v[0] = V[0] + tt[0]; v[1] = V[1] + tt[1]; v[2] = V[2] + tt[2];
            v[3] = V[3]; v[4] = V[4]; v[5] = V[5];
            const T K[20] = {R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8], p[0], p[1], p[2],
                             v[0], v[1], v[2], v[3], v[4], v[5], T(0), T(0)};
            stv<20>(ri, K);
            if (i > 0) {
              T ax[3], aw[3];
              ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
              mat3_vec(R, ax, aw);
              const T Q[8] = {aw[0], aw[1], aw[2], T(0), p[0] - pp[0], p[1] - pp[1], p[2] - pp[2], tr_r[t]};
              stv<8>(ri + K_AX, Q);  // axis, (sdd), r, torque reference
              ri[K_S] = s_r[t];
              ri[K_SD] = sd_r[t];
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
          T* ri = ws + (size_t)i * R2;
          ri[K_S] = s_r[t];
          ri[K_SD] = sd_r[t];
          ri[K_TREF] = tr_r[t];
        }
      }
      __syncwarp();
    }
    B200SIM_MARK2(5);

    // step -1 (only without input caches) is the kinematics of the input state: it runs the tail of the
    // step body (joint transforms + FK walk) on the unchanged joint state
    for (int step = use_cached ? 0 : -1; step < P.nsteps; ++step) {
      const bool last = (step == P.nsteps - 1);
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
          const T* R = Kb + K_R;
          const T* vl = Kb + K_V;
          const T* w = Kb + K_V + 3;
          const T pz = Kb[K_P + 2];
          T Lp[3], d[3], pd[3];
          ldn<3>(sm_pt + 3 * k, Lp);
          mat3_vec(R, Lp, d);
          const T pcz = pz + d[2];
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
          const T* R = Kk + K_R;
          const T* p = Kk + K_P;
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
          if (fext_step) {
            const T* fx = fext_step + (env * nL + i) * 6;
            T f[3] = {fx[0], fx[1], fx[2]};
            fe[0] += f[0]; fe[1] += f[1]; fe[2] += f[2];
            ne[0] += fx[3]; ne[1] += fx[4]; ne[2] += fx[5];
            T t[3];
            cross3(p, f, t);  // moment about the link origin = moment about W origin - p x f
            ne[0] -= t[0]; ne[1] -= t[1]; ne[2] -= t[2];
          }
          // link inertia in world axes about the link origin
          const T mass = c[C_MASS];
          T com[3], cw[3], Dl[6];
          ldn<3>(c + C_COM, com);
          ldn<6>(c + C_DL, Dl);
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
            const T sdi = ri[K_SD];
            T aw[3];
            ldn<3>(ri + K_AX, aw);
            T cc[6];
            const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
            if (jt == 1) {
              cross3(v, vJ, cc);
              cross3(v + 3, vJ, cc + 3);
            } else {
              cross3(v + 3, vJ, cc);
              cc[3] = cc[4] = cc[5] = T(0);
            }
            stn<6>(ri + K_C, cc);
            const T tref = ri[K_TREF];
            const T si = ri[K_S];
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
            tau_i = min_t(max_t(tt, -lim), lim);
            // torque reference of the NEXT fused step: its latency hides behind the ABA passes
            if (!last && P.tau && P.tau_step_stride)
              cp_async_elem(ri + K_TREF, P.tau + (long long)(step + 1) * P.tau_step_stride + env * n + (i - 1));
          }
          // articulated inertia init (overwrites R, p, v): A = m 1, B = -m S(c_w), D = D_w
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
        {
          int e = nrows > 0 ? rows[(nrows - 1) * G + lane] : 0xFF;
          for (int r = nrows - 1; r >= 0; --r) {
            __syncwarp();
            const int en = r > 0 ? rows[(r - 1) * G + lane] : 0xFF;
            const bool valid = (e & 0xFF) != 0xFF;
            const int nsub = (e >> 20) & 15;
            const int rank = (e >> 16) & 15;
            T X[27];  // contribution to the parent: A(6) B(9) D(6) pA(6), about the parent's origin
            T* rp = ws + (size_t)((e >> 8) & 0xFF) * R2;
            if (valid) {
              T* ri = ws + (size_t)(e & 0xFF) * R2;
              const int jt = (e >> 27) & 3;
              T W[44];
              ldv<44>(ri, W);
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

        // ======================================================= base acceleration (rbda/aba.py:240-242), every lane
        T a0[6];
        {
          T W[28];
          ldv<28>(ws, W);
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
          solve6_spd_neg(M, pA, a0);
          if (lane == 0) {
            const T a8[8] = {a0[0], a0[1], a0[2], a0[3], a0[4], a0[5], T(0), T(0)};
            stv<8>(ws + K_C, a8);
          }
        }
        B200SIM_MARK2(9);

        // ======================================================= ABA pass 3 (rbda/aba.py:244-282), root first
        {
          int e = nrows > 0 ? rows[lane] : 0xFF;
          for (int r = 0; r < nrows; ++r) {
            __syncwarp();
            const int en = (r + 1 < nrows) ? rows[(r + 1) * G + lane] : 0xFF;
            if ((e & 0xFF) != 0xFF) {
              T* ri = ws + (size_t)(e & 0xFF) * R2;
              const T* rp = ws + (size_t)((e >> 8) & 0xFF) * R2;
              const int jt = (e >> 27) & 3;
              T ap[8], U[8], W[16];
              ldv<8>(rp + K_C, ap);  // parent's acceleration (+ its s, sd)
              ldv<8>(ri + K_U, U);   // U, 1/d, u
              ldv<16>(ri + K_C, W);  // c, s, sd, axis, (sdd), r, tref
              const T* cI = W;
              T* aw = W + (K_AX - K_C);
              const T* rr = W + (K_RR - K_C);
              T a[6];
              cross3(ap + 3, rr, a);
              a[0] += ap[0] + cI[0]; a[1] += ap[1] + cI[1]; a[2] += ap[2] + cI[2];
              a[3] = ap[3] + cI[3]; a[4] = ap[4] + cI[4]; a[5] = ap[5] + cI[5];
              const T sdd = (U[7] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * U[6];
              if (jt == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
              else { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
#pragma unroll
              for (int k = 0; k < 6; ++k) W[k] = a[k];
              aw[3] = sdd;           // K_SDD follows the axis
              stv<12>(ri + K_C, W);  // a, s, sd, axis, sdd
            }
            e = en;
          }
          __syncwarp();
        }
        B200SIM_MARK2(10);

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
        __pipeline_wait_prior(0);  // next step's torque references have landed
      }

      // ========================================================= joints: Euler update + joint transforms
      // (api/kin_dyn_parameters.py:396-451) of the state the FK walk below needs: the next fused step's, or the
      // new state's for the cache outputs
      const bool want_caches = last && (P.W_H_L || P.W_v);
      const bool need_kin = !last || want_caches || (P.iXl != nullptr);
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * R2;
        T sdn = ri[K_SD], sn = ri[K_S];
        if (step >= 0) {
          sdn += dt * ri[K_SDD];
          sn += dt * sdn;
          ri[K_SD] = sdn;
          ri[K_S] = sn;
          if (last && active) {
            P.sd_o[env * n + (i - 1)] = sdn;
            P.s_o[env * n + (i - 1)] = sn;
          }
        }
        if (need_kin) {
          T Rrel[9], trel[3];
          joint_rel_transform(P, 0, sm_cst + (size_t)i * CREC, jtypes[i], i, sn, Rrel, trel);
          {
            const T K[12] = {Rrel[0], Rrel[1], Rrel[2], Rrel[3], Rrel[4], Rrel[5], Rrel[6], Rrel[7], Rrel[8], trel[0], trel[1], trel[2]};
            stv<12>(ri, K);
          }
          if (last && active && P.iXl) {
            T X[36];
            inverse_adjoint(X, Rrel, trel);
            stg_vec<36>(P.iXl + (env * nL + i) * 36, X);
          }
        }
      }
      B200SIM_MARK2(12);
      if (need_kin) {
        if (lane == 0) {
          // chain root: the base link (suc_H_i[0] = I for the models this kernel serves)
          T v0[6], t[3];
          cross3(b.w, b.p, t);  // velocity of the base origin: v_lin + w x p
          v0[0] = b.vlin[0] + t[0]; v0[1] = b.vlin[1] + t[1]; v0[2] = b.vlin[2] + t[2];
          v0[3] = b.w[0]; v0[4] = b.w[1]; v0[5] = b.w[2];
          {
            const T K[20] = {b.R[0], b.R[1], b.R[2], b.R[3], b.R[4], b.R[5], b.R[6], b.R[7], b.R[8], b.p[0], b.p[1], b.p[2],
                             v0[0], v0[1], v0[2], v0[3], v0[4], v0[5], T(0), T(0)};
            stv<20>(ws, K);
          }
          if (last && active && P.iXl) {
            // index 0: Ad((W_H_B suc_H_i[0])^-1)  (api/kin_dyn_parameters.py:417-449)
            T X[36];
            inverse_adjoint(X, b.R, b.p);
            stg_vec<36>(P.iXl + env * nL * 36, X);
          }
        }
        // ======================================================= FK + velocity walk over the tree levels
        // (rbda/forward_kinematics.py:80-113, pass 1 of rbda/aba.py:131-171 in F_i coordinates)
        int e = nrows > 0 ? rows[lane] : 0xFF;
        for (int r = 0; r < nrows; ++r) {
          __syncwarp();
          const int en = (r + 1 < nrows) ? rows[(r + 1) * G + lane] : 0xFF;
          if ((e & 0xFF) != 0xFF) {
            const int i = e & 0xFF;
            const T* rp = ws + (size_t)((e >> 8) & 0xFF) * R2;
            T* ri = ws + (size_t)i * R2;
            T Kp[20], Ko[12];
            ldv<20>(rp, Kp);
            ldv<12>(ri, Ko);
            const T* Rp = Kp + K_R;
            const T* pp = Kp + K_P;
            const T* vp = Kp + K_V;
            const T* Rrel = Ko + K_R;
            const T* trel = Ko + K_P;
            T R[9], rr[3], pw[3];
            mat3_mul(Rp, Rrel, R);
            mat3_vec(Rp, trel, rr);
            pw[0] = pp[0] + rr[0]; pw[1] = pp[1] + rr[1]; pw[2] = pp[2] + rr[2];
            T ax[3], aw[3];
            ldn<3>(sm_cst + (size_t)i * CREC + C_AXIS, ax);
            mat3_vec(R, ax, aw);
            const T sdi = ri[K_SD];
            T v[6];
            cross3(vp + 3, rr, v);
            v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
            v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
            if (((e >> 27) & 3) == 1) { v[3] += sdi * aw[0]; v[4] += sdi * aw[1]; v[5] += sdi * aw[2]; }
            else { v[0] += sdi * aw[0]; v[1] += sdi * aw[1]; v[2] += sdi * aw[2]; }
            {
              const T K[20] = {R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8], pw[0], pw[1], pw[2],
                               v[0], v[1], v[2], v[3], v[4], v[5], T(0), T(0)};
              stv<20>(ri, K);
            }
            stn<3>(ri + K_RR, rr);
            stn<3>(ri + K_AX, aw);
          }
          e = en;
        }
        __syncwarp();
        B200SIM_MARK2(13);
        if (want_caches && active) {
          for (int i = lane; i < nL; i += G) {
            const T* ri = ws + (size_t)i * R2;
            T Kk[20];
            ldv<20>(ri, Kk);
            const T* R = Kk + K_R;
            const T* p = Kk + K_P;
            const T* v = Kk + K_V;
            if (P.W_H_L) store_transform(P.W_H_L + (env * nL + i) * 16, R, p);
            if (P.W_v) {
              T t[3];
              cross3(p, v + 3, t);  // inertial-fixed linear part: vlin + p x w
              const T o[6] = {v[0] + t[0], v[1] + t[1], v[2] + t[2], v[3], v[4], v[5]};
              stg_vec6(P.W_v + (env * nL + i) * 6, o);
            }
          }
        }
        B200SIM_MARK2(14);
      }
    }  // steps
  }
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
