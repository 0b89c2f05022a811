// b200sim_ws_kernel.cuh -- warp-specialised variant of the fused step kernel.
//
// Why: the single-role kernel (b200sim_kernels.cuh) is latency bound -- at batch 4096 every
// environment is resident, each warp walks an ~8900-instruction dependent stream and only
// ~7 warps per SM exist to hide its stalls (profiles/r01_step_kernel_v3.md: issue slots 32 %
// busy).  Here TWO warps cooperate on the same four environments with different ROLES, so
// the stream per warp is about half as long and twice as many warps are resident for the
// same shared memory:
//
//   role A ("rotations / matrices"):  sincos + joint rotations, the rotation chain of FK,
//       link inertias in world axes, the MATRIX half of ABA pass 2 (U, d, Ma, shift), the
//       LDL^T factorisation of the floating-base inertia, the joint-adjoint cache output;
//   role B ("translations / vectors"): joint translations + actuation model, the position /
//       velocity chain of FK, collidable points + Hunt/Crossley, bias forces, the VECTOR
//       half of pass 2 (u, pa), base substitution, pass 3, the integrator, the link
//       transform / velocity cache outputs.
//
// The two warps of a pair meet at named barriers (bar.sync id, 64); inside a role the
// usual __syncwarp() between tree levels applies.  Both roles execute the SAME barrier
// sequence by construction: barriers only appear in the role-independent skeleton below.
//
// Scope: MODE_STEP (any nsteps), soft contacts or none, floating-base URDF-style models
// (suc_H_i = I, FK chain == ABA chain).  Everything else runs on the single-role kernel;
// b200sim.cu picks.  Same arithmetic as the single-role kernel, operation for operation,
// except that the gyroscopic term uses R (D_l (R^T w)) instead of (R D_l R^T) w.
#pragma once

#include "b200sim_kernels.cuh"

namespace b200sim {
namespace ws {

constexpr int REC = 68;  // words per link (68 = 4*17)
// U (6) and w = IA c (6) overlay R, p once the kinematics of the step have been consumed
constexpr int O_R = 0;      // 9  A
constexpr int O_P = 9;      // 3  B
constexpr int O_U = 0;      // 6  A (pass 2 ->) B (pass 3)
constexpr int O_W = 6;      // 6  A -> B
constexpr int O_IA = 12;    // 21 A   (A 12..17, B 18..26, D 27..32); link 0: LDL^T factor for B
constexpr int O_PA = 33;    // 6  B
constexpr int O_C = 39;     // 6  B (chain) -> A (w), B
constexpr int O_DINV = 45;  // 1  A -> B
constexpr int O_TAU = 46;   // 1  B: resultant torque, then u_i
constexpr int O_AX = 47;    // 3  B (chain) -> A
constexpr int O_RR = 50;    // 3  B (chain) -> A
constexpr int O_V = 53;     // 6  B: velocity, later acceleration
constexpr int O_SD = 59;
constexpr int O_S = 60;
constexpr int O_SDD = 61;
constexpr int O_TREF = 62;
constexpr int O_X = 12;     // 36 staging of the joint adjoint (output phase), words 12..47

__host__ __device__ inline size_t env_ws_words(int nL, int nc) {
  size_t w = (size_t)nL * REC + (size_t)nc * PTREC;
  return (w + 3) & ~size_t(3);
}

__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(512, 1) step_kernel_ws(const Params<T> P) {
  constexpr int G = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm_cst = reinterpret_cast<T*>(smem_raw);
  const int nL = P.nL, n = P.n, nc = P.nc;
  T* sm_pt = sm_cst + (size_t)nL * CREC;
  const size_t pt_words = ((size_t)nc * 3 + 3) & ~size_t(3);
  int* sm_itab = reinterpret_cast<int*>(sm_pt + pt_words);
  const size_t itab_words = ((size_t)P.itab_words + 3) & ~size_t(3);
  T* ws_base = reinterpret_cast<T*>(sm_itab + itab_words);

  stage_async(sm_cst, P.cst, nL * CREC);
  stage_async(sm_pt, P.pt_pos, (int)pt_words);
  stage_async(sm_itab, P.itab, (int)itab_words);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();

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

  const int warp = threadIdx.x >> 5;
  const int role = warp & 1;   // 0: A, 1: B
  const int pair = warp >> 1;
  const int lane = threadIdx.x & (G - 1);
  const int grp = (threadIdx.x & 31) >> 3;  // environment within the pair
  const bool isA = role == 0;
  const size_t wsw = env_ws_words(nL, nc);
  T* ws = ws_base + (size_t)(pair * 4 + grp) * wsw;
  T* ptws = ws + (size_t)nL * REC;
  const long long stride = (long long)gridDim.x * P.envs_per_block;
  const T dt = P.dt;
  const bool soft = (P.contact_model == 1) && nc > 0;
  const bool tma = (P.flags & F_TMA_STORE) != 0;

  for (long long env0 = (long long)blockIdx.x * P.envs_per_block; env0 < P.B; env0 += stride) {
    long long env = env0 + pair * 4 + grp;
    const bool active = env < P.B;
    if (!active) env = P.B - 1;

    // ================================================================ S0: input burst
    if (isA) {
      if (tma) tma_store_wait_read();
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        cp_async_elem(ri + O_S, P.s + env * n + (i - 1));
        cp_async_elem(ri + O_SD, P.sd + env * n + (i - 1));
        if (P.tau) cp_async_elem(ri + O_TREF, P.tau + env * n + (i - 1));
        else ri[O_TREF] = T(0);
      }
    } else {
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
    BaseState<T> b;  // replicated in every lane of both roles
    {
      const T* q = P.q + env * 4;
      T qr[4] = {q[0], q[1], q[2], q[3]};
      ldn<3>(P.p + env * 3, b.p);
      ldn<3>(P.vlin + env * 3, b.vlin);
      ldn<3>(P.omega + env * 3, b.w);
      const T nrm = sqrt_t(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const T inv = T(1) / (nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));
#pragma unroll
      for (int k = 0; k < 4; ++k) b.qn[k] = qr[k] * inv;
      quat_to_dcm(b.qn, b.R);
    }
    __pipeline_wait_prior(0);
    pair_sync(pair);

    // ---- role pieces used more than once -------------------------------------------------
    // S1 / S6: relative joint transforms of the current joint positions
    auto joints_A = [&]() {  // rotations
      if (lane == 0) stn<9>(ws + O_R, b.R);
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* c = sm_cst + (size_t)i * CREC;
        T sn = T(0), cs = T(1);
        if (jtypes[i] == 1) sincos_t(ri[O_S], &sn, &cs);
        T Rrel[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Rrel[k] = c[C_M0 + k] + cs * c[C_M1 + k] + sn * c[C_M2 + k];
        stn<9>(ri + O_R, Rrel);
      }
    };
    auto joints_B = [&]() {  // translations + base record
      if (lane == 0) {
        stn<3>(ws + O_P, b.p);
        T v0[6];
        if (P.floating) {
          T t[3];
          cross3(b.w, b.p, t);
          v0[0] = b.vlin[0] + t[0]; v0[1] = b.vlin[1] + t[1]; v0[2] = b.vlin[2] + t[2];
          v0[3] = b.w[0]; v0[4] = b.w[1]; v0[5] = b.w[2];
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) v0[k] = T(0);
        }
        stn<6>(ws + O_V, v0);
      }
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* c = sm_cst + (size_t)i * CREC;
        const T s = ri[O_S];
        T trel[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) trel[k] = c[C_TPRE + k] + s * c[C_RA + k];
        stn<3>(ri + O_P, trel);
      }
    };
    // one tree level of the kinematic chain: A rotations, B positions / velocities
    auto chain_level = [&](const int l, const bool for_aba) {
      const int e = lvl_start[l + 1];
      for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
        const int i = lvl_links[idx];
        const T* rp = ws + (size_t)parent[i] * REC;
        T* ri = ws + (size_t)i * REC;
        T Rp[9];
        ldn<9>(rp + O_R, Rp);
        if (isA) {
          T Rrel[9], R[9];
          ldn<9>(ri + O_R, Rrel);
          mat3_mul(Rp, Rrel, R);
          stn<9>(ri + O_R, R);
        } else {
          T pp[3], vp[6], trel[3], pax[3];
          ldn<3>(rp + O_P, pp);
          ldn<6>(rp + O_V, vp);
          ldn<3>(ri + O_P, trel);
          ldn<3>(sm_cst + (size_t)i * CREC + C_PAX, pax);
          T r[3], pw[3], aw[3];
          mat3_vec(Rp, trel, r);
          pw[0] = pp[0] + r[0]; pw[1] = pp[1] + r[1]; pw[2] = pp[2] + r[2];
          mat3_vec(Rp, pax, aw);
          const T sdi = ri[O_SD];
          T v[6];
          cross3(vp + 3, r, v);
          v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
          v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
          const int jt = jtypes[i];
          const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
          T cc[6];
          if (jt == 1) {
            cross3(v, vJ, cc);  // uses the parent-shifted linear velocity: v_lin x vJ_ang
            v[3] += vJ[0]; v[4] += vJ[1]; v[5] += vJ[2];
            cross3(v + 3, vJ, cc + 3);  // (w + vJ) x vJ = w x vJ
          } else {
            v[0] += vJ[0]; v[1] += vJ[1]; v[2] += vJ[2];
            cross3(v + 3, vJ, cc);  // w x vJ_lin
            cc[3] = cc[4] = cc[5] = T(0);
          }
          stn<3>(ri + O_P, pw);
          stn<6>(ri + O_V, v);
          if (for_aba) {
            stn<3>(ri + O_RR, r);
            stn<3>(ri + O_AX, aw);
            stn<6>(ri + O_C, cc);
          }
        }
      }
    };

    // ================================================================ S1: joint transforms
    if (isA) {
      joints_A();
    } else {
      joints_B();
      // actuation model (api/actuation_model.py:7-126): resultant torque of step 0
      for (int i = 1 + lane; i < nL; i += G) {
        T* ri = ws + (size_t)i * REC;
        const T* c = sm_cst + (size_t)i * CREC;
        const T si = ri[O_S], sdi = ri[O_SD], tref = ri[O_TREF];
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
        ri[O_TAU] = min_t(max_t(tt, -lim), lim);
      }
    }
    pair_sync(pair);
    // ================================================================ S2: kinematic chain
    for (int l = 1; l <= P.depth; ++l) {
      chain_level(l, true);
      pair_sync(pair);
    }

    for (int step = 0; step < P.nsteps; ++step) {
      const bool last = (step == P.nsteps - 1);
      const T* fext_step = P.fext ? P.fext + (long long)step * P.fext_step_stride : nullptr;

      // ============================================================== S3
      if (isA) {
        // link inertias in world axes about the link origins -> articulated inertia init
        for (int i = lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * REC;
          const T* c = sm_cst + (size_t)i * CREC;
          T R[9];
          ldn<9>(ri + O_R, R);
          const T mass = c[C_MASS];
          T com[3], cw[3], Dl[6], Dw[6];
          ldn<3>(c + C_COM, com);
          ldn<6>(c + C_DL, Dl);
          mat3_vec(R, com, cw);
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
          T IA[21];
          IA[0] = mass; IA[1] = T(0); IA[2] = T(0); IA[3] = mass; IA[4] = T(0); IA[5] = mass;
          IA[6] = T(0);            IA[7] = mass * cw[2];   IA[8] = -mass * cw[1];
          IA[9] = -mass * cw[2];   IA[10] = T(0);          IA[11] = mass * cw[0];
          IA[12] = mass * cw[1];   IA[13] = -mass * cw[0]; IA[14] = T(0);
#pragma unroll
          for (int k = 0; k < 6; ++k) IA[15 + k] = Dw[k];
          if (i == 0 && !P.floating) {
#pragma unroll
            for (int k = 0; k < 21; ++k) IA[k] = T(0);
          }
          stn<21>(ri + O_IA, IA);
        }
      } else {
        // collidable points + Hunt/Crossley (same code as the single-role kernel)
        if (soft) {
          for (int k = lane; k < nc; k += G) {
            const int bi = pt_body[k];
            const T* rb = ws + (size_t)bi * REC;
            T R[9], p[3], va[6];
            ldn<9>(rb + O_R, R);
            ldn<3>(rb + O_P, p);
            ldn<6>(rb + O_V, va);
            T Lp[3], d[3], pc[3], pd[3];
            ldn<3>(sm_pt + 3 * k, Lp);
            mat3_vec(R, Lp, d);
            pc[0] = p[0] + d[0]; pc[1] = p[1] + d[1]; pc[2] = p[2] + d[2];
            cross3(va + 3, d, pd);
            pd[0] += va[0]; pd[1] += va[1]; pd[2] += va[2];
            T* pw = ptws + (size_t)k * PTREC;
            T m[3];
            ldn<3>(pw + PT_M, m);
            T f[3] = {T(0), T(0), T(0)};
            T md[3] = {T(0), T(0), T(0)};
            if (pt_enabled[k]) {
              const T delta = max_t(T(0), P.h_terrain - pc[2]);
              const T ddot = (delta > T(0)) ? -pd[2] : T(0);
              const T eps = Lim<T>::eps();
              const T dp = (P.flags & F_SQRT_P) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.pexp);
              const T dq = (P.flags & F_SQRT_Q) ? sqrt_t(delta + eps) : pow_t(delta + eps, P.qexp);
              const T Kdp = P.K * dp, Ddq = P.D * dq;
              const T fn = max_t(T(0), Kdp * delta + Ddq * ddot);
              T ft0 = -(Kdp * m[0] + Ddq * pd[0]);
              T ft1 = -(Kdp * m[1] + Ddq * pd[1]);
              const T mufn = P.mu * fn;
              const bool nocontact = delta <= T(0);
              const bool sticking = nocontact || (ft0 * ft0 + ft1 * ft1 <= mufn * mufn);
              const T nrm = sqrt_t(ft0 * ft0 + ft1 * ft1);
              const T idn = T(1) / (nrm + eps * (nrm == T(0) ? T(1) : T(0)));
              if (!sticking) {
                const T sc = min_t(mufn, nrm) * idn;
                ft0 *= sc; ft1 *= sc;
              }
              if (nocontact) { ft0 = T(0); ft1 = T(0); }
              const T KoD = P.K / P.D;
              if (nocontact) {
                md[0] = -KoD * m[0]; md[1] = -KoD * m[1]; md[2] = -KoD * m[2];
              } else if (sticking) {
                md[0] = pd[0]; md[1] = pd[1]; md[2] = -KoD * m[2];
              } else {
                const T iD = T(1) / Ddq;
                md[0] = -(ft0 + Kdp * m[0]) * iD; md[1] = -(ft1 + Kdp * m[1]) * iD; md[2] = T(0);
              }
              f[0] = ft0; f[1] = ft1; f[2] = fn;
            }
            stn<3>(pw + PT_F, f);
            stn<3>(pw + PT_LEV, d);
            m[0] += dt * md[0]; m[1] += dt * md[1]; m[2] += dt * md[2];
            stn<3>(pw + PT_M, m);
            if (last && active && P.m_o) {
              T* mo = P.m_o + (env * nc + k) * 3;
              mo[0] = m[0]; mo[1] = m[1]; mo[2] = m[2];
            }
          }
        } else if (last && P.m_o && P.m && active) {
          for (int k = lane; k < nc * 3; k += G) P.m_o[env * nc * 3 + k] = P.m[env * nc * 3 + k];
        }
        __syncwarp();
        // bias forces pA = v x* (I v) - f_ext (rbda/aba.py:160) with I v through R, D_l
        for (int i = lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * REC;
          const T* c = sm_cst + (size_t)i * CREC;
          T R[9], p[3], v[6];
          ldn<9>(ri + O_R, R);
          ldn<3>(ri + O_P, p);
          ldn<6>(ri + O_V, v);
          T fe[3] = {T(0), T(0), T(0)}, ne[3] = {T(0), T(0), T(0)};
          if (soft) {
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
            cross3(p, f, t);
            ne[0] -= t[0]; ne[1] -= t[1]; ne[2] -= t[2];
          }
          const T mass = c[C_MASS];
          T com[3], cw[3], Dl[6];
          ldn<3>(c + C_COM, com);
          ldn<6>(c + C_DL, Dl);
          mat3_vec(R, com, cw);
          T fI[3], nI[3], t[3], wl[3], Dwl[3];
          cross3(v + 3, cw, t);
          fI[0] = mass * (v[0] + t[0]); fI[1] = mass * (v[1] + t[1]); fI[2] = mass * (v[2] + t[2]);
          mat3T_vec(R, v + 3, wl);   // angular velocity in link axes
          sym3_vec(Dl, wl, Dwl);
          mat3_vec(R, Dwl, nI);      // D_w w = R D_l R^T w
          cross3(cw, v, t);
          nI[0] += mass * t[0]; nI[1] += mass * t[1]; nI[2] += mass * t[2];
          T pA[6];
          cross3(v + 3, fI, pA);
          cross3(v, fI, pA + 3);
          cross3_add(v + 3, nI, pA + 3);
          pA[0] -= fe[0]; pA[1] -= fe[1]; pA[2] -= fe[2];
          pA[3] -= ne[0]; pA[4] -= ne[1]; pA[5] -= ne[2];
          if (i == 0 && !P.floating) {
#pragma unroll
            for (int k = 0; k < 6; ++k) pA[k] = T(0);
          }
          stn<6>(ri + O_PA, pA);
        }
        // joint force references of the NEXT fused step (hidden behind A's sweep)
        if (!last && P.tau && P.tau_step_stride) {
          for (int i = 1 + lane; i < nL; i += G)
            cp_async_elem(ws + (size_t)i * REC + O_TREF, P.tau + (long long)(step + 1) * P.tau_step_stride + env * n + (i - 1));
        }
        __pipeline_commit();
      }
      pair_sync(pair);  // kinematics consumed: R, p may now be overwritten by U, w

      // ============================================================== S4: A matrix sweep
      if (isA) {
        for (int l = P.depth; l >= 1; --l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            T* ri = ws + (size_t)i * REC;
            T A[6], Bm[9], D[6];
            ldn<6>(ri + O_IA, A);
            ldn<9>(ri + O_IA + 6, Bm);
            ldn<6>(ri + O_IA + 15, D);
            const int ce = child_start[i + 1];
            for (int cc = child_start[i]; cc < ce; ++cc) {
              const T* rc = ws + (size_t)child_idx[cc] * REC + O_IA;
#pragma unroll
              for (int k = 0; k < 6; ++k) A[k] += rc[k];
#pragma unroll
              for (int k = 0; k < 9; ++k) Bm[k] += rc[6 + k];
#pragma unroll
              for (int k = 0; k < 6; ++k) D[k] += rc[15 + k];
            }
            T aw[3], cI[6], r[3];
            ldn<3>(ri + O_AX, aw);
            ldn<6>(ri + O_C, cI);
            ldn<3>(ri + O_RR, r);
            const int jt = jtypes[i];
            T Ul[3], Ua[3], d;
            if (jt == 1) {
              mat3_vec(Bm, aw, Ul);
              sym3_vec(D, aw, Ua);
              d = dot3(aw, Ua);
            } else {
              sym3_vec(A, aw, Ul);
              mat3T_vec(Bm, aw, Ua);
              d = dot3(aw, Ul);
            }
            const T dinv = rcp_t(d);
            // w = IA c (the vector sweep needs Ma c = w - U (U.c)/d)
            T w6[6], t3[3];
            sym3_vec(A, cI, w6);
            mat3_vec(Bm, cI + 3, t3);
            w6[0] += t3[0]; w6[1] += t3[1]; w6[2] += t3[2];
            mat3T_vec(Bm, cI, w6 + 3);
            sym3_vec(D, cI + 3, t3);
            w6[3] += t3[0]; w6[4] += t3[1]; w6[5] += t3[2];
            stn<3>(ri + O_U, Ul);
            stn<3>(ri + O_U + 3, Ua);
            stn<6>(ri + O_W, w6);
            ri[O_DINV] = dinv;
            const int par = parent[i];
            if (par != 0 || P.floating) {
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
              const T Af[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
              T B2[9];
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                T rowx[3];
                cross3(Af + 3 * a, r, rowx);
                B2[3 * a] = Bm[3 * a] - rowx[0]; B2[3 * a + 1] = Bm[3 * a + 1] - rowx[1]; B2[3 * a + 2] = Bm[3 * a + 2] - rowx[2];
              }
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
              stn<6>(ri + O_IA, A);
              stn<9>(ri + O_IA + 6, B2);
              stn<6>(ri + O_IA + 15, D2);
            }
          }
        }
        __syncwarp();
        // LDL^T of the floating-base articulated inertia; B substitutes (rbda/aba.py:241)
        if (lane == 0 && P.floating) {
          T* r0 = ws;
          T A[6], Bm[9], D[6];
          ldn<6>(r0 + O_IA, A);
          ldn<9>(r0 + O_IA + 6, Bm);
          ldn<6>(r0 + O_IA + 15, D);
          const int ce = child_start[1];
          for (int cc = child_start[0]; cc < ce; ++cc) {
            const T* rc = ws + (size_t)child_idx[cc] * REC + O_IA;
#pragma unroll
            for (int k = 0; k < 6; ++k) A[k] += rc[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) Bm[k] += rc[6 + k];
#pragma unroll
            for (int k = 0; k < 6; ++k) D[k] += rc[15 + k];
          }
          T M[6][6];
          M[0][0] = A[0]; M[1][1] = A[3]; M[2][2] = A[5];
          M[1][0] = A[1]; M[2][0] = A[2]; M[2][1] = A[4];
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) M[3 + bb][a] = Bm[3 * a + bb];
          M[3][3] = D[0]; M[4][4] = D[3]; M[5][5] = D[5];
          M[4][3] = D[1]; M[5][3] = D[2]; M[5][4] = D[4];
          // in-place LDL^T on the lower triangle (same recurrences as solve6_spd_neg)
          T dinv6[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            T dj = M[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) dj -= M[j][k] * M[j][k] * M[k][k];
            M[j][j] = dj;
            dinv6[j] = T(1) / dj;
#pragma unroll
            for (int i2 = j + 1; i2 < 6; ++i2) {
              T v = M[i2][j];
#pragma unroll
              for (int k = 0; k < j; ++k) v -= M[i2][k] * M[j][k] * M[k][k];
              M[i2][j] = v * dinv6[j];
            }
          }
          T F[21];
          int o = 0;
#pragma unroll
          for (int i2 = 1; i2 < 6; ++i2)
#pragma unroll
            for (int k = 0; k < i2; ++k) F[o++] = M[i2][k];
#pragma unroll
          for (int k = 0; k < 6; ++k) F[15 + k] = dinv6[k];
          stn<21>(r0 + O_IA, F);
        }
      }
      pair_sync(pair);

      // ============================================================== S5: B vector sweep, pass 3, integrator
      T Wa[6];
      BaseState<T> nb = b;
      if (!isA) {
        for (int l = P.depth; l >= 1; --l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            T* ri = ws + (size_t)i * REC;
            T pA[6];
            ldn<6>(ri + O_PA, pA);
            const int ce = child_start[i + 1];
            for (int cc = child_start[i]; cc < ce; ++cc) {
              const T* rc = ws + (size_t)child_idx[cc] * REC + O_PA;
#pragma unroll
              for (int k = 0; k < 6; ++k) pA[k] += rc[k];
            }
            T aw[3], cI[6], r[3], U[6], w6[6];
            ldn<3>(ri + O_AX, aw);
            ldn<6>(ri + O_C, cI);
            ldn<3>(ri + O_RR, r);
            ldn<6>(ri + O_U, U);
            ldn<6>(ri + O_W, w6);
            const T dinv = ri[O_DINV];
            const int jt = jtypes[i];
            const T u = ri[O_TAU] - ((jt == 1) ? dot3(aw, pA + 3) : dot3(aw, pA));
            ri[O_TAU] = u;
            const int par = parent[i];
            if (par != 0 || P.floating) {
              // pa = pA + Ma c + U u/d, Ma c = IA c - U (U.c)/d
              const T Uc = U[0] * cI[0] + U[1] * cI[1] + U[2] * cI[2] + U[3] * cI[3] + U[4] * cI[4] + U[5] * cI[5];
              const T k1 = (u - Uc) * dinv;
              T pa[6];
#pragma unroll
              for (int k = 0; k < 6; ++k) pa[k] = pA[k] + w6[k] + U[k] * k1;
              cross3_add(r, pa, pa + 3);
              stn<6>(ri + O_PA, pa);
            }
          }
        }
        __syncwarp();
        // base acceleration (substitution with A's factor), replicated result via record 0
        if (lane == 0) {
          T a0[6];
          if (P.floating) {
            T* r0 = ws;
            T pA[6], F[21];
            ldn<6>(r0 + O_PA, pA);
            ldn<21>(r0 + O_IA, F);
            const int ce = child_start[1];
            for (int cc = child_start[0]; cc < ce; ++cc) {
              const T* rc = ws + (size_t)child_idx[cc] * REC + O_PA;
#pragma unroll
              for (int k = 0; k < 6; ++k) pA[k] += rc[k];
            }
            // L y = -pA ; x = L^-T (D^-1 y).  L(i,k) = F[i(i-1)/2 + k]
            T y[6];
#pragma unroll
            for (int i2 = 0; i2 < 6; ++i2) {
              T v = -pA[i2];
#pragma unroll
              for (int k = 0; k < i2; ++k) v -= F[i2 * (i2 - 1) / 2 + k] * y[k];
              y[i2] = v;
            }
#pragma unroll
            for (int i2 = 5; i2 >= 0; --i2) {
              T v = y[i2] * F[15 + i2];
#pragma unroll
              for (int k = i2 + 1; k < 6; ++k) v -= F[k * (k - 1) / 2 + i2] * a0[k];
              a0[i2] = v;
            }
          } else {
            a0[0] = T(0); a0[1] = T(0); a0[2] = -P.g; a0[3] = T(0); a0[4] = T(0); a0[5] = T(0);
          }
          stn<6>(ws + O_V, a0);
        }
        // pass 3 (rbda/aba.py:251-277)
        for (int l = 1; l <= P.depth; ++l) {
          __syncwarp();
          const int e = lvl_start[l + 1];
          for (int idx = lvl_start[l] + lane; idx < e; idx += G) {
            const int i = lvl_links[idx];
            T* ri = ws + (size_t)i * REC;
            const T* rp = ws + (size_t)parent[i] * REC;
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
            const T sdd = (ri[O_TAU] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * ri[O_DINV];
            const int jt = jtypes[i];
            if (jt == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
            else if (jt == 2) { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
            stn<6>(ri + O_V, a);
            ri[O_SDD] = sdd;
          }
        }
        __syncwarp();
        if (P.floating) {
          T a0[6];
          ldn<6>(ws + O_V, a0);
          cross3(b.p, a0 + 3, Wa);
          Wa[0] += a0[0]; Wa[1] += a0[1]; Wa[2] += a0[2] + P.g;
          Wa[3] = a0[3]; Wa[4] = a0[4]; Wa[5] = a0[5];
        } else {
#pragma unroll
          for (int k = 0; k < 6; ++k) Wa[k] = T(0);
        }
        // semi-implicit Euler, base (api/integrators.py:14-88)
        {
#pragma unroll
          for (int k = 0; k < 3; ++k) { nb.vlin[k] = b.vlin[k] + dt * Wa[k]; nb.w[k] = b.w[k] + dt * Wa[3 + k]; }
          T pd[3];
          cross3(nb.w, b.p, pd);
          pd[0] += nb.vlin[0]; pd[1] += nb.vlin[1]; pd[2] += nb.vlin[2];
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
#pragma unroll
          for (int rep = 0; rep < 2; ++rep) {
            const T nn = sqrt_t(qn2[0] * qn2[0] + qn2[1] * qn2[1] + qn2[2] * qn2[2] + qn2[3] * qn2[3]);
            const T inv = T(1) / ((nn == T(0)) ? T(1) : nn);
#pragma unroll
            for (int k = 0; k < 4; ++k) qn2[k] *= inv;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) nb.qn[k] = qn2[k];
        }
        // the new base state travels to role A through record 0: the 13 words from O_C on
        // belong to joint quantities, which link 0 does not have
        if (lane == 0) {
          T* x = ws + O_C;  // O_C(6) O_DINV O_TAU O_AX(3) O_RR(2 of 3): 13 words, unused by link 0
          stn<4>(x, nb.qn);
          stn<3>(x + 4, nb.p);
          stn<3>(x + 7, nb.vlin);
          stn<3>(x + 10, nb.w);
        }
        if (last && active && lane == 0) {
          stn<4>(P.q_o + env * 4, nb.qn);
          stn<3>(P.p_o + env * 3, nb.p);
          stn<3>(P.vlin_o + env * 3, nb.vlin);
          stn<3>(P.omega_o + env * 3, nb.w);
        }
        __pipeline_wait_prior(0);  // next step's torque references
        // joints: integrate; for a following step also its resultant torque
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
          if (!last) {
            const T* c = sm_cst + (size_t)i * CREC;
            const T tref = ri[O_TREF];
            const T lower = min_t(sn - c[C_SMIN], T(0));
            const T upper = max_t(sn - c[C_SMAX], T(0));
            T tlim = -c[C_KS] * (lower + upper);
            tlim = tlim - tlim * c[C_KD] * sdn;
            T tfr = T(0);
            if (P.enable_friction) {
              const T sg = (sdn > T(0)) ? T(1) : ((sdn < T(0)) ? T(-1) : T(0));
              tfr = -(c[C_KC] * sg + c[C_KV] * sdn);
            }
            const T tt = tref + tfr + tlim;
            const T av = abs_t(sdn);
            T lim;
            if (av <= P.w_th) lim = P.tau_max;
            else if (av <= P.w_max) lim = P.tau_max * (T(1) - (av - P.w_th) / (P.w_max - P.w_th));
            else lim = T(0);
            ri[O_TAU] = min_t(max_t(tt, -lim), lim);
          }
        }
      }
      pair_sync(pair);

      // both roles adopt the new base state (B computed it; A reads it from record 0)
      {
        const T* x = ws + O_C;
        ldn<4>(x, nb.qn);
        ldn<3>(x + 4, nb.p);
        ldn<3>(x + 7, nb.vlin);
        ldn<3>(x + 10, nb.w);
        b = nb;
        if (!last) {
          // a following fused step starts like a fresh call (api/data.py:283-285)
          const T nrm = sqrt_t(b.qn[0] * b.qn[0] + b.qn[1] * b.qn[1] + b.qn[2] * b.qn[2] + b.qn[3] * b.qn[3]);
          const T inv = T(1) / (nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));
#pragma unroll
          for (int k = 0; k < 4; ++k) b.qn[k] *= inv;
        }
        quat_to_dcm(b.qn, b.R);
      }
      const bool want_caches = last && (P.W_H_L || P.W_v);
      const bool need_fk = !last || want_caches || (P.iXl != nullptr);

      // ============================================================== S6: joints of the new state
      if (need_fk) {
        if (isA) joints_A();
        else joints_B();
      }
      pair_sync(pair);
      if (last && P.iXl && isA && active) {
        // joint adjoints (A): needs R_rel (own) and t_rel (B, previous barrier)
        if (lane == 0) {
          T X[36];
          T R0[9], p0[3], t[3];
          mat3_mul(b.R, sm_cst + C_M0, R0);
          mat3_vec(b.R, sm_cst + C_TPRE, t);
          p0[0] = b.p[0] + t[0]; p0[1] = b.p[1] + t[1]; p0[2] = b.p[2] + t[2];
          inverse_adjoint(X, R0, p0);
          stg_vec<36>(P.iXl + env * nL * 36, X);
          if (P.W_H_B) store_transform(P.W_H_B + env * 16, b.R, b.p);
        }
        for (int i = 1 + lane; i < nL; i += G) {
          T* ri = ws + (size_t)i * REC;
          T Rrel[9], trel[3], X[36];
          ldn<9>(ri + O_R, Rrel);
          ldn<3>(ri + O_P, trel);
          inverse_adjoint(X, Rrel, trel);
          T* dst = P.iXl + (env * nL + i) * 36;
          if (tma) {
            stn<36>(ri + O_X, X);
            tma_store_bulk(dst, ri + O_X, 36 * sizeof(T));
          } else {
            stg_vec<36>(dst, X);
          }
        }
      } else if (last && P.W_H_B && !P.iXl && isA && active && lane == 0) {
        store_transform(P.W_H_B + env * 16, b.R, b.p);
      }
      // ============================================================== S7: chain of the new state
      if (!last || want_caches) {
        for (int l = 1; l <= P.depth; ++l) {
          chain_level(l, !last);
          pair_sync(pair);
        }
        if (last && active && !isA) {
          for (int i = lane; i < nL; i += G) {
            const T* ri = ws + (size_t)i * REC;
            T R[9], p[3], v[6];
            ldn<9>(ri + O_R, R);
            ldn<3>(ri + O_P, p);
            ldn<6>(ri + O_V, v);
            if (P.W_H_L) store_transform(P.W_H_L + (env * nL + i) * 16, R, p);
            if (P.W_v) {
              T t[3];
              cross3(p, v + 3, t);
              const T o[6] = {v[0] + t[0], v[1] + t[1], v[2] + t[2], v[3], v[4], v[5]};
              stg_vec6(P.W_v + (env * nL + i) * 6, o);
            }
          }
        }
      }
      pair_sync(pair);
    }  // steps
  }
  if (tma) tma_store_wait_all();
}

}  // namespace ws
}  // namespace b200sim
