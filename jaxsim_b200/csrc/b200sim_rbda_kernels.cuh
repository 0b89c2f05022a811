// b200sim_rbda_kernels.cuh -- RNEA and CRBA on the same tree-walk skeleton as the step
// kernel: same shared-memory records, same world-aligned
// link-origin frames F_i = (p_i, world axes), same level-synchronous recursions.
//
//   rnea_kernel  == vmapped rbda.rnea  (src/jaxsim/rbda/rnea.py:12-238)
//   crba_kernel  == vmapped rbda.crba  (src/jaxsim/rbda/crba.py:10-170)
#pragma once

#include "b200sim_kernels.cuh"

namespace b200sim {

// Shared-memory context of one thread block (model staged once) + this group's workspace.
template <typename T>
struct BlockCtx {
  T* cst;
  T* pt;
  int* itab;
  const int *parent, *jtypes, *lvl_start, *lvl_links, *child_start, *child_idx;
  T* ws;
};

template <typename T, int G>
__device__ __forceinline__ BlockCtx<T> stage_model(const Params<T>& P, unsigned char* smem_raw) {
  BlockCtx<T> c;
  c.cst = reinterpret_cast<T*>(smem_raw);
  c.pt = c.cst + (size_t)P.nL * CREC;
  const size_t pt_words = ((size_t)P.nc * 3 + 3) & ~size_t(3);
  c.itab = reinterpret_cast<int*>(c.pt + pt_words);
  const size_t itab_words = ((size_t)P.itab_words + 3) & ~size_t(3);
  T* ws_base = reinterpret_cast<T*>(c.itab + itab_words);
  stage_async(c.cst, P.cst, P.nL * CREC);
  stage_async(c.itab, P.itab, (int)itab_words);
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();
  c.parent = c.itab + P.o_parent;
  c.jtypes = c.itab + P.o_jtype;
  c.lvl_start = c.itab + P.o_lvl_start;
  c.lvl_links = c.itab + P.o_lvl_links;
  c.child_start = c.itab + P.o_child_start;
  c.child_idx = c.itab + P.o_child_idx;
  c.ws = ws_base + (size_t)(threadIdx.x / G) * env_ws_words<T>(P.nL, P.nc);
  return c;
}

// level-synchronous FK + velocity chain (same as the step kernel's, always storing r, a_w)
template <typename T, int G>
__device__ __forceinline__ void fk_chain_full(const Params<T>& P, const BlockCtx<T>& c, int lane) {
  for (int l = 1; l <= P.depth; ++l) {
    __syncwarp();
    const int e = c.lvl_start[l + 1];
    for (int idx = c.lvl_start[l] + lane; idx < e; idx += G) {
      const int i = c.lvl_links[idx];
      const T* rp = c.ws + (size_t)c.parent[i] * REC;
      T* ri = c.ws + (size_t)i * REC;
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
      ldn<3>(c.cst + (size_t)i * CREC + C_AXIS, ax);
      mat3_vec(R, ax, aw);
      const T sdi = ri[O_SD];
      T v[6];
      cross3(vp + 3, r, v);
      v[0] += vp[0]; v[1] += vp[1]; v[2] += vp[2];
      v[3] = vp[3]; v[4] = vp[4]; v[5] = vp[5];
      const int jt = c.jtypes[i];
      if (jt == 1) { v[3] += sdi * aw[0]; v[4] += sdi * aw[1]; v[5] += sdi * aw[2]; }
      else if (jt == 2) { v[0] += sdi * aw[0]; v[1] += sdi * aw[1]; v[2] += sdi * aw[2]; }
      stn<9>(ri + O_R, R);
      stn<3>(ri + O_P, pw);
      stn<6>(ri + O_V, v);
      stn<3>(ri + O_RR, r);
      stn<3>(ri + O_AX, aw);
    }
  }
  __syncwarp();
}

// rigid-body inertia of link i in F_i: mass, c_w = R com, D_w = R D_link R^T
template <typename T>
__device__ __forceinline__ void link_inertia_world(const T* c, const T* R, T& mass, T* cw, T* Dw) {
  mass = c[C_MASS];
  T com[3], Dl[6];
  ldn<3>(c + C_COM, com);
  ldn<6>(c + C_DL, Dl);
  mat3_vec(R, com, cw);
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

// extra per-call pointers of the RBDA entry points
template <typename T>
struct RbdaArgs {
  const T* avd;   // (B,6) inertial-fixed base acceleration (RNEA in)
  const T* sdd;   // (B,n) joint accelerations (RNEA in)
  T* W_f0;        // (B,6) base wrench (RNEA out)
  T* tau_o;       // (B,n) joint forces (RNEA out)
  T* M;           // (B,6+n,6+n) mass matrix (CRBA out, pre-zeroed)
};

// ======================================================================================
// RNEA
// ======================================================================================
template <typename T, int G>
__global__ void __launch_bounds__(LaunchBounds<G>::kThreads, 1) rnea_kernel(const Params<T> P, const RbdaArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BlockCtx<T> c = stage_model<T, G>(P, smem_raw);
  const int nL = P.nL, n = P.n;
  const int lane = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  T* ws = c.ws;
  const long long stride = (long long)gridDim.x * P.envs_per_block;
  for (long long env0 = (long long)blockIdx.x * P.envs_per_block; env0 < P.B; env0 += stride) {
    long long env = env0 + grp;
    const bool active = env < P.B;
    if (!active) env = P.B - 1;
    for (int i = 1 + lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      cp_async_elem(ri + O_S, P.s + env * n + (i - 1));
      cp_async_elem(ri + O_SD, P.sd + env * n + (i - 1));
      if (A.sdd) cp_async_elem(ri + O_SDD, A.sdd + env * n + (i - 1));
      else ri[O_SDD] = T(0);
    }
    __pipeline_commit();
    // base
    T qn[4], Rb[9], pb[3], vl[3], w[3];
    {
      const T* q = P.q + env * 4;
      T qr[4] = {q[0], q[1], q[2], q[3]};
      ldn<3>(P.p + env * 3, pb);
      ldn<3>(P.vlin + env * 3, vl);
      ldn<3>(P.omega + env * 3, w);
      const T nrm = sqrt_t(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const T inv = T(1) / (nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));
#pragma unroll
      for (int k = 0; k < 4; ++k) qn[k] = qr[k] * inv;
      quat_to_dcm(qn, Rb);
    }
    T a0[6];  // base spatial acceleration in F_0 (gravity folded in, rbda/rnea.py:108-123)
    if (P.floating) {
      T av[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      if (A.avd) ldn<6>(A.avd + env * 6, av);
      T t[3];
      cross3(pb, av + 3, t);  // W -> F_0: lin - p x ang
      a0[0] = av[0] - t[0]; a0[1] = av[1] - t[1]; a0[2] = av[2] - P.g - t[2];
      a0[3] = av[3]; a0[4] = av[4]; a0[5] = av[5];
    } else {
      a0[0] = T(0); a0[1] = T(0); a0[2] = -P.g; a0[3] = T(0); a0[4] = T(0); a0[5] = T(0);
    }
    __pipeline_wait_prior(0);
    if (lane == 0) {
      stn<9>(ws + O_R, Rb);
      stn<3>(ws + O_P, pb);
      T v0[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      if (P.floating) {
        T t[3];
        cross3(w, pb, t);
        v0[0] = vl[0] + t[0]; v0[1] = vl[1] + t[1]; v0[2] = vl[2] + t[2];
        v0[3] = w[0]; v0[4] = w[1]; v0[5] = w[2];
      }
      stn<6>(ws + O_V, v0);
      stn<6>(ws + O_C, a0);
    }
    for (int i = 1 + lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      T Rrel[9], trel[3];
      joint_rel_transform(P, P.flags, c.cst + (size_t)i * CREC, c.jtypes[i], i, ri[O_S], Rrel, trel);
      stn<9>(ri + O_R, Rrel);
      stn<3>(ri + O_P, trel);
    }
    fk_chain_full<T, G>(P, c, lane);
    // acceleration chain: a_i = X a_parent + S sdd + v_i x vJ   (rnea.py:152-153)
    for (int l = 1; l <= P.depth; ++l) {
      __syncwarp();
      const int e = c.lvl_start[l + 1];
      for (int idx = c.lvl_start[l] + lane; idx < e; idx += G) {
        const int i = c.lvl_links[idx];
        T* ri = ws + (size_t)i * REC;
        const T* rp = ws + (size_t)c.parent[i] * REC;
        T ap[6], r[3], aw[3], v[6];
        ldn<6>(rp + O_C, ap);
        ldn<3>(ri + O_RR, r);
        ldn<3>(ri + O_AX, aw);
        ldn<6>(ri + O_V, v);
        const T sdi = ri[O_SD], sddi = ri[O_SDD];
        T a[6];
        cross3(ap + 3, r, a);
        a[0] += ap[0]; a[1] += ap[1]; a[2] += ap[2];
        a[3] = ap[3]; a[4] = ap[4]; a[5] = ap[5];
        const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
        T cc[3];
        if (c.jtypes[i] == 1) {
          a[3] += sddi * aw[0]; a[4] += sddi * aw[1]; a[5] += sddi * aw[2];
          cross3(v, vJ, cc);
          a[0] += cc[0]; a[1] += cc[1]; a[2] += cc[2];
          cross3(v + 3, vJ, cc);
          a[3] += cc[0]; a[4] += cc[1]; a[5] += cc[2];
        } else {
          a[0] += sddi * aw[0]; a[1] += sddi * aw[1]; a[2] += sddi * aw[2];
          cross3(v + 3, vJ, cc);
          a[0] += cc[0]; a[1] += cc[1]; a[2] += cc[2];
        }
        stn<6>(ri + O_C, a);
      }
    }
    __syncwarp();
    // link forces f_i = I a + v x* I v - f_ext   (rnea.py:162-168), stored in O_PA
    for (int i = lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      T R[9], p[3], v[6], a[6];
      ldn<9>(ri + O_R, R);
      ldn<3>(ri + O_P, p);
      ldn<6>(ri + O_V, v);
      ldn<6>(ri + O_C, a);
      T mass, cw[3], Dw[6];
      link_inertia_world(c.cst + (size_t)i * CREC, R, mass, cw, Dw);
      T f[6], t[3], fI[3], nI[3];
      // I a
      cross3(a + 3, cw, t);
      f[0] = mass * (a[0] + t[0]); f[1] = mass * (a[1] + t[1]); f[2] = mass * (a[2] + t[2]);
      sym3_vec(Dw, a + 3, f + 3);
      cross3(cw, a, t);
      f[3] += mass * t[0]; f[4] += mass * t[1]; f[5] += mass * t[2];
      // v x* (I v)
      cross3(v + 3, cw, t);
      fI[0] = mass * (v[0] + t[0]); fI[1] = mass * (v[1] + t[1]); fI[2] = mass * (v[2] + t[2]);
      sym3_vec(Dw, v + 3, nI);
      cross3(cw, v, t);
      nI[0] += mass * t[0]; nI[1] += mass * t[1]; nI[2] += mass * t[2];
      cross3_add(v + 3, fI, f);
      cross3_add(v, fI, f + 3);
      cross3_add(v + 3, nI, f + 3);
      if (P.fext) {
        const T* fx = P.fext + (env * nL + i) * 6;
        const T fe[3] = {fx[0], fx[1], fx[2]};
        f[0] -= fe[0]; f[1] -= fe[1]; f[2] -= fe[2];
        cross3(p, fe, t);
        f[3] -= fx[3] - t[0]; f[4] -= fx[4] - t[1]; f[5] -= fx[5] - t[2];
      }
      if (i == 0 && !P.floating) {
#pragma unroll
        for (int k = 0; k < 6; ++k) f[k] = T(0);
      }
      stn<6>(ri + O_PA, f);
    }
    // backward pass: tau_i = S^T f_i, f_parent += X^T f_i   (rnea.py:193-229)
    for (int l = P.depth; l >= 1; --l) {
      __syncwarp();
      const int e = c.lvl_start[l + 1];
      for (int idx = c.lvl_start[l] + lane; idx < e; idx += G) {
        const int i = c.lvl_links[idx];
        T* ri = ws + (size_t)i * REC;
        T f[6], aw[3], r[3];
        ldn<6>(ri + O_PA, f);
        const int ce = c.child_start[i + 1];
        for (int cc = c.child_start[i]; cc < ce; ++cc) {
          const T* rc = ws + (size_t)c.child_idx[cc] * REC;
#pragma unroll
          for (int k = 0; k < 6; ++k) f[k] += rc[O_PA + k];
        }
        ldn<3>(ri + O_AX, aw);
        ldn<3>(ri + O_RR, r);
        const T tau = (c.jtypes[i] == 1) ? dot3(aw, f + 3) : dot3(aw, f);
        if (active) A.tau_o[env * n + (i - 1)] = tau;
        cross3_add(r, f, f + 3);  // moment about the parent's origin
        stn<6>(ri + O_PA, f);
      }
    }
    __syncwarp();
    if (lane == 0 && active) {
      T f[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      if (P.floating) {
        ldn<6>(ws + O_PA, f);
        const int ce = c.child_start[1];
        for (int cc = c.child_start[0]; cc < ce; ++cc) {
          const T* rc = ws + (size_t)c.child_idx[cc] * REC;
#pragma unroll
          for (int k = 0; k < 6; ++k) f[k] += rc[O_PA + k];
        }
        cross3_add(pb, f, f + 3);  // W_f0 = B_X_W^T f0: moment about the world origin
      }
      stn<6>(A.W_f0 + env * 6, f);
    }
    __syncwarp();
  }
}

// ======================================================================================
// CRBA (body-fixed representation: the base frame is the identity, crba.py:33-35)
// ======================================================================================
template <typename T, int G>
__global__ void __launch_bounds__(LaunchBounds<G>::kThreads, 1) crba_kernel(const Params<T> P, const RbdaArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BlockCtx<T> c = stage_model<T, G>(P, smem_raw);
  const int nL = P.nL, n = P.n;
  const int lane = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  T* ws = c.ws;
  const int N = 6 + n;
  const long long stride = (long long)gridDim.x * P.envs_per_block;
  for (long long env0 = (long long)blockIdx.x * P.envs_per_block; env0 < P.B; env0 += stride) {
    long long env = env0 + grp;
    const bool active = env < P.B;
    if (!active) env = P.B - 1;
    T* Mo = A.M + env * (long long)N * N;
    for (int i = 1 + lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      cp_async_elem(ri + O_S, P.s + env * n + (i - 1));
      ri[O_SD] = T(0);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    if (lane == 0) {
      const T I3[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
      const T z6[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      stn<9>(ws + O_R, I3);
      stn<3>(ws + O_P, z6);
      stn<6>(ws + O_V, z6);
    }
    for (int i = 1 + lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      T Rrel[9], trel[3];
      joint_rel_transform(P, P.flags, c.cst + (size_t)i * CREC, c.jtypes[i], i, ri[O_S], Rrel, trel);
      stn<9>(ri + O_R, Rrel);
      stn<3>(ri + O_P, trel);
    }
    fk_chain_full<T, G>(P, c, lane);
    // composite inertia init = link inertia in F_i (overwrites R, p)
    for (int i = lane; i < nL; i += G) {
      T* ri = ws + (size_t)i * REC;
      T R[9];
      ldn<9>(ri + O_R, R);
      T mass, cw[3], Dw[6];
      link_inertia_world(c.cst + (size_t)i * CREC, R, mass, cw, Dw);
      T IA[21];
      IA[0] = mass; IA[1] = T(0); IA[2] = T(0); IA[3] = mass; IA[4] = T(0); IA[5] = mass;
      IA[6] = T(0);            IA[7] = mass * cw[2];   IA[8] = -mass * cw[1];
      IA[9] = -mass * cw[2];   IA[10] = T(0);          IA[11] = mass * cw[0];
      IA[12] = mass * cw[1];   IA[13] = -mass * cw[0]; IA[14] = T(0);
#pragma unroll
      for (int k = 0; k < 6; ++k) IA[15 + k] = Dw[k];
      stn<21>(ri + O_IA, IA);
    }
    // backward pass: Mc_parent += X^T Mc_i X ; F_i = Mc_i S_i ; M_ii = S_i^T F_i (crba.py:84-95)
    for (int l = P.depth; l >= 1; --l) {
      __syncwarp();
      const int e = c.lvl_start[l + 1];
      for (int idx = c.lvl_start[l] + lane; idx < e; idx += G) {
        const int i = c.lvl_links[idx];
        T* ri = ws + (size_t)i * REC;
        T Am[6], Bm[9], D[6];
        ldn<6>(ri + O_IA, Am);
        ldn<9>(ri + O_IB, Bm);
        ldn<6>(ri + O_ID, D);
        const int ce = c.child_start[i + 1];
        for (int cc = c.child_start[i]; cc < ce; ++cc) {
          const T* rc = ws + (size_t)c.child_idx[cc] * REC;
#pragma unroll
          for (int k = 0; k < 6; ++k) Am[k] += rc[O_IA + k];
#pragma unroll
          for (int k = 0; k < 9; ++k) Bm[k] += rc[O_IB + k];
#pragma unroll
          for (int k = 0; k < 6; ++k) D[k] += rc[O_ID + k];
        }
        T aw[3], r[3], F[6];
        ldn<3>(ri + O_AX, aw);
        ldn<3>(ri + O_RR, r);
        T mii;
        if (c.jtypes[i] == 1) {
          mat3_vec(Bm, aw, F);
          sym3_vec(D, aw, F + 3);
          mii = dot3(aw, F + 3);
        } else {
          sym3_vec(Am, aw, F);
          mat3T_vec(Bm, aw, F + 3);
          mii = dot3(aw, F);
        }
        stn<6>(ri + O_U, F);
        if (active) Mo[(long long)(5 + i) * N + (5 + i)] = mii;
        // shift the composite inertia to the parent's origin (same algebra as ABA pass 2)
        const T Af[9] = {Am[0], Am[1], Am[2], Am[1], Am[3], Am[4], Am[2], Am[4], Am[5]};
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
        stn<6>(ri + O_IA, Am);
        stn<9>(ri + O_IB, B2);
        stn<6>(ri + O_ID, D2);
      }
    }
    __syncwarp();
    // locked 6x6 inertia of the whole tree in the base frame (crba.py:168)
    if (lane == 0 && active) {
      T Am[6], Bm[9], D[6];
      ldn<6>(ws + O_IA, Am);
      ldn<9>(ws + O_IB, Bm);
      ldn<6>(ws + O_ID, D);
      const int ce = c.child_start[1];
      for (int cc = c.child_start[0]; cc < ce; ++cc) {
        const T* rc = ws + (size_t)c.child_idx[cc] * REC;
#pragma unroll
        for (int k = 0; k < 6; ++k) Am[k] += rc[O_IA + k];
#pragma unroll
        for (int k = 0; k < 9; ++k) Bm[k] += rc[O_IB + k];
#pragma unroll
        for (int k = 0; k < 6; ++k) D[k] += rc[O_ID + k];
      }
      const T Af[9] = {Am[0], Am[1], Am[2], Am[1], Am[3], Am[4], Am[2], Am[4], Am[5]};
      const T Df[9] = {D[0], D[1], D[2], D[1], D[3], D[4], D[2], D[4], D[5]};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          Mo[(long long)a * N + b] = Af[3 * a + b];
          Mo[(long long)a * N + 3 + b] = Bm[3 * a + b];
          Mo[(long long)(3 + b) * N + a] = Bm[3 * a + b];
          Mo[(long long)(3 + a) * N + 3 + b] = Df[3 * a + b];
        }
    }
    // off-diagonal blocks: walk each joint's force up to the root (crba.py:97-162)
    for (int i = 1 + lane; i < nL; i += G) {
      T F[6];
      ldn<6>(ws + (size_t)i * REC + O_U, F);
      int j = i;
      while (true) {
        const T* rj = ws + (size_t)j * REC;
        T r[3];
        ldn<3>(rj + O_RR, r);
        cross3_add(r, F, F + 3);  // X_j^T F: moment about the parent's origin
        j = c.parent[j];
        if (j == 0) break;
        T aw[3];
        ldn<3>(ws + (size_t)j * REC + O_AX, aw);
        const T mij = (c.jtypes[j] == 1) ? dot3(aw, F + 3) : dot3(aw, F);
        if (active) {
          Mo[(long long)(5 + i) * N + (5 + j)] = mij;
          Mo[(long long)(5 + j) * N + (5 + i)] = mij;
        }
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          Mo[(long long)k * N + (5 + i)] = F[k];
          Mo[(long long)(5 + i) * N + k] = F[k];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace b200sim
