// jaxsim_b200 -- the rigid-contact variant of the simulation step (BASELINE config 3).
//
// Reference path (src/jaxsim): api/model.py:2601-2681 (step) with
// rbda/contacts/rigid.py:222-436 (compute_contact_forces, update_velocity_after_impact),
// api/contact.py:214-511 (contact Jacobians and their derivative), rbda/mass_inverse.py,
// rbda/jacobian.py, api/ode.py:16-131, api/integrators.py:14-88.
//
// The reference assembles dense objects per environment -- M^-1 (6+n)^2, the stacked contact
// Jacobians J, J_dot (nc,6,6+n), the Delassus matrix J M^-1 J^T (3nc)^2, a KKT matrix
// (6+n+3nc)^2 for the impact -- and hands a QP with 3nc unknowns / 6nc inequalities to qpax.
// None of M, M^-1, J, J_dot is ever formed here.  One warp owns one environment and keeps its
// whole state in shared memory:
//
//   * kinematics, ABA (free acceleration) exactly as in the soft-contact kernel: world-aligned
//     link-origin frames F_i, level-synchronous tree walks;
//   * J_dot nu + J nu_dot_free of a contact point is the classical acceleration of that point,
//     a_lin + alpha x rho + omega x pdot (plus gravity), read off the ABA's pass-3 link
//     accelerations -- with the reference's quirk that `pdot` (and the penetration rate) come
//     from the CACHED link velocities of the input data, which are the pre-impact velocities of
//     the previous step (rigid.py:429-434 does not refresh the caches, api/contact.py:470-477
//     reads them);
//   * the Delassus matrix comes column by column from articulated-body impulse responses
//     (unit force at an active point -> up the tree with the pass-2 quantities U_i, 1/d_i ->
//     6x6 base solve -> down to every link that carries active points), one column per lane,
//     O(active points x depth) instead of O((6+n)^3); only the ACTIVE points enter (the
//     reference's QP pins the inactive ones to zero through rows 4-5 of G, rigid.py:478-492);
//   * the QP (friction pyramids) is solved by a warp-cooperative primal-dual interior-point
//     method (Mehrotra predictor-corrector) on a packed lower-triangular Hessian in shared
//     memory, iterated to the resolution of the arithmetic (the reference stops qpax at 1e-3);
//   * the contact forces act through one more articulated-body response (no second full ABA);
//   * the impact (rigid.py:163-220: lstsq on the KKT system) is the M-orthogonal projection of
//     nu onto {J_active nu = 0}: Delassus of the new configuration, a rank-revealing Cholesky
//     (dependent constraints of coplanar points are skipped, which is what the minimum-norm
//     lstsq solution amounts to for nu), and an impulse response.
//
// Supported: floating-base models whose base link frame is the ABA chain root (no
// F_GENERIC_FK), enabled collidable points forming a prefix of the point list, nsteps == 1.
#pragma once

#include "b200sim_kernels.cuh"

namespace b200sim {

// per collidable point (words of T): lever arm rho in F_body (3), bias term (3), force (3)
constexpr int RPT = 9;
constexpr int RP_LEV = 0, RP_B = 3, RP_F = 6;
constexpr int RIGID_MAX_WARPS = 8;

// Packed lower triangle, every row starting on a 16-byte boundary (rows of odd length carry one pad element): the
// row-times-row products of the factorisation then read two elements per shared-memory access (the contact-QP kernel
// is bound by shared-memory wavefronts in that loop, profiles/r02_rigid_qp_kernel.md).
//   prow(r) = start of row r = sum of the padded lengths (i + 2) & ~1 of the rows above; npk(N) = words of an order-N matrix
__host__ __device__ __forceinline__ int prow(int r) { const int m = r >> 1; return 2 * (m + 1) * (m + (r & 1)); }
__host__ __device__ __forceinline__ int pidx(int r, int c) { return prow(r) + c; }  // r >= c
__host__ __device__ __forceinline__ size_t npk(size_t N) { return (size_t)prow((int)N); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

struct RigidLayout {  // byte offsets inside one environment's (= one warp's) workspace
  size_t links, pts, ainv, ucol, Qp, Hp, vecN, vecM, ints, total;
};

__host__ __device__ inline size_t rl_align(size_t x) { return (x + 15) & ~size_t(15); }

// `cap` = number of simultaneously active points the solver arrays are sized for (<= nc)
// `qp_mode` (Params::qp_mode): the launches of a split level never run the interior-point method, so they do not carry
// its vectors (mode 1, assemble only: nor the second matrix) -- one or two more resident warps per SM
template <typename T, typename S>
__host__ __device__ inline RigidLayout rigid_layout(int nL, int nc, int depth, int cap, int qp_mode = 0) {
  RigidLayout L;
  size_t o = 0;
  const size_t N = 3 * (size_t)cap, M = qp_mode ? 0 : 5 * (size_t)cap, NP = npk(N);
  const size_t dd = depth > 0 ? depth : 1;
  L.links = o; o = rl_align(o + sizeof(T) * (size_t)nL * REC);
  L.pts = o;   o = rl_align(o + sizeof(T) * (size_t)nc * RPT);
  L.ainv = o;  o = rl_align(o + sizeof(T) * 36);
  L.ucol = o;  o = rl_align(o + sizeof(T) * 32 * dd);
  L.Qp = o;    o = rl_align(o + sizeof(S) * NP);
  L.Hp = o;    o = rl_align(o + sizeof(S) * (qp_mode == 1 ? 0 : NP));
  L.vecN = o;  o = rl_align(o + sizeof(S) * 7 * N);
  L.vecM = o;  o = rl_align(o + sizeof(S) * 9 * M);
  L.ints = o;  o = rl_align(o + sizeof(int) * (5 * (size_t)nc + (size_t)nL + 4));
  L.total = o;
  return L;
}

// Split rigid level: record of one work item in global memory, written by the assembling launch (qp_mode 1), solved
// by rigid_qp_kernel, consumed by the resuming launch (qp_mode 2):  int na, rc, env, pad | S q[3 cap] | S x[3 cap] |
// S Q[packed lower triangle of order 3 cap; an item with na active points uses the leading 3 na (3 na + 1) / 2 entries]
template <typename S>
__host__ __device__ inline size_t qp_record_problem_bytes(int cap) {
  const size_t N = 3 * (size_t)cap;
  return rl_align(16 + sizeof(S) * (2 * N + npk(N)));
}
// ... followed by what the resuming launch would otherwise recompute (16 of its 74 us per item): the link records, point
// records and base inverse inertia of the assembling launch [words of T: nL REC + nc RPT + 36, as laid out in the
// workspace] and its active-point / contact-link lists [2 nc + nL ints]; header word 3 carries the number of contact links
// Measured (ErgoCub-like, 16 384 standing environments): float64 data 6.0 -> 5.1 ms per step; float32 data 3.44 -> 3.48 ms
// (the float32 recomputation is as cheap as pulling 13 KB per item back from HBM), so only float64 records carry the state.
template <typename T>
struct QpSaveState { static constexpr bool value = sizeof(T) == 8; };
template <typename T, typename S>
__host__ __device__ inline size_t qp_record_bytes(int nL, int nc, int cap) {
  if (!QpSaveState<T>::value) return qp_record_problem_bytes<S>(cap);
  const size_t state = rl_align(sizeof(T) * (size_t)nL * REC) + rl_align(sizeof(T) * (size_t)nc * RPT) + rl_align(sizeof(T) * 36);
  return qp_record_problem_bytes<S>(cap) + state + rl_align(sizeof(int) * (2 * (size_t)nc + (size_t)nL));
}

template <typename S> struct QpTol;
template <> struct QpTol<float> {
  static __device__ __forceinline__ float tol() { return 1e-5f; }
  static __device__ __forceinline__ float pivot_floor() { return 1e-7f; }
  static constexpr int max_iter = 40;
};
template <> struct QpTol<double> {
  static __device__ __forceinline__ double tol() { return 1e-12; }
  static __device__ __forceinline__ double pivot_floor() { return 1e-15; }
  static constexpr int max_iter = 60;
};

// rank decisions of the impact solve depend on the precision the Delassus matrix was COMPUTED in
template <typename T> struct RankTol;
template <> struct RankTol<float> { static __device__ __forceinline__ double tol() { return 1e-4; } };
template <> struct RankTol<double> { static __device__ __forceinline__ double tol() { return 1e-10; } };

constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) { return rsqrt(x); }

// phase timeline of one warp (diagnostic, like B200SIM_PHASE_MARK of the step kernel): dbg[8 + 16 + k], k = 1..15
#define B200SIM_RIGID_MARK(k)                                                                                   \
  do {                                                                                                           \
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0 && it0 == 0) P.dbg[8 + 16 + (k)] = (unsigned long long)clock64(); \
  } while (0)

template <typename S>
__device__ __forceinline__ S warp_sum(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <typename S>
__device__ __forceinline__ S warp_max(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max_t(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// maximum of FINITE operands as compare + select (3 instructions; fmax on doubles is an 8-instruction sequence, and the
// solver takes ~60 maxima per lane and iteration).  Non-finite iterates are caught through the sums (see qp_pyramids).
template <typename S>
__device__ __forceinline__ S qmax(S a, S b) { return a > b ? a : b; }
template <typename S>
__device__ __forceinline__ S warp_qmax(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = qmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
template <typename S>
__device__ __forceinline__ S warp_min(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min_t(v, __shfl_xor_sync(FULL, v, o));
  return v;
}


// y = A x for a packed lower-triangular symmetric A (rows lane-strided)
template <typename S>
__device__ __forceinline__ void sym_matvec(const S* Ap, const S* x, S* y, int N, int lane) {
  for (int r = lane; r < N; r += 32) {
    S acc = S(0);
    const S* row = Ap + prow(r);
    for (int c = 0; c <= r; ++c) acc += row[c] * x[c];
    for (int c = r + 1; c < N; ++c) acc += Ap[pidx(c, r)] * x[c];
    y[r] = acc;
  }
}

// In-place Cholesky of a packed lower-triangular matrix, warp-cooperative, LEFT-looking: at
// step k every lane owns rows r > k and forms L[r,k] = (H[r,k] - <L[r,0:k], L[k,0:k]>) / L[k,k]
// -- equal work per lane (a k-term dot product over a contiguous packed row against a broadcast
// row), one barrier per step; the diagonal H[r,r] is kept as the running Schur complement.
// `dg` holds the diagonal before elimination on entry (pivots are floored at pivot_floor * dg:
// the interior-point Hessian is positive definite up to rounding) and 1 / L_kk on exit, so
// that the triangular solves multiply instead of divide.
template <typename S>
__device__ __forceinline__ void chol_packed(S* __restrict__ Hp, S* __restrict__ dg, int N, int lane) {
  for (int k = 0; k < N; ++k) {
    S d = max_t(Hp[pidx(k, k)], QpTol<S>::pivot_floor() * dg[k]);
    if (!(d > S(0))) d = S(1);
    const S ipiv = rsqrt_t(d);
    const S* rowk = Hp + prow(k);
    for (int r = k + 1 + lane; r < N; r += 32) {
      S* row = Hp + prow(r);
      // four independent partial sums, two elements of either row per shared-memory access
      S a0 = row[k], a1 = S(0), a2 = S(0), a3 = S(0);
      int c = 0;
      for (; c + 3 < k; c += 4) {
        const auto p0 = ld2(row + c), q0 = ld2(rowk + c), p1 = ld2(row + c + 2), q1 = ld2(rowk + c + 2);
        a0 -= p0.x * q0.x; a1 -= p0.y * q0.y;
        a2 -= p1.x * q1.x; a3 -= p1.y * q1.y;
      }
      // the 0-3 leftover terms, predicated (no loop)
      if (c + 1 < k) {
        const auto p0 = ld2(row + c), q0 = ld2(rowk + c);
        a0 -= p0.x * q0.x; a1 -= p0.y * q0.y;
        c += 2;
      }
      if (c < k) a2 -= row[c] * rowk[c];
      const S v = ((a0 + a1) + (a2 + a3)) * ipiv;
      row[k] = v;
      row[r] -= v * v;
    }
    __syncwarp();  // every lane has read dg[k] (the pivot floor) before lane 0 replaces it (racecheck: write after read)
    if (lane == 0) dg[k] = ipiv;
    __syncwarp();
  }
}

// solve L L^T x = y in place; invd = 1 / diag(L).  The vector lives in registers (lane owns
// rows lane, lane+32, lane+64): the pivot element travels by shuffle, no shared-memory round
// trip or barrier inside the substitution loops.  N <= 96.
// NS = number of 32-row slots a lane may own (N <= 32 NS): the per-step work of the two substitution loops -- the
// critical path of an interior-point iteration next to the factorisation -- shrinks with it.
template <typename S, int NS = 3>
__device__ __forceinline__ void chol_solve(const S* __restrict__ Lp, const S* __restrict__ invd, S* y, int N, int lane) {
  S yr[NS];
  const S* rowr[NS];  // start of this lane's packed rows
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    yr[j] = (lane + 32 * j < N) ? y[lane + 32 * j] : S(0);
    rowr[j] = Lp + prow(lane + 32 * j);
  }
  for (int k = 0; k < N; ++k) {
    const int slot = k >> 5;
    S src = yr[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) src = (slot == j) ? yr[j] : src;
    const S yk = __shfl_sync(FULL, src, k & 31) * invd[k];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int r = lane + 32 * j;
      if (r > k && r < N) yr[j] -= rowr[j][k] * yk;
      else if (r == k) yr[j] = yk;
    }
  }
  const S* row = Lp + prow(N);
  for (int k = N - 1; k >= 0; --k) {
    const int slot = k >> 5;
    S src = yr[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) src = (slot == j) ? yr[j] : src;
    const S xk = __shfl_sync(FULL, src, k & 31) * invd[k];
    row -= (k + 2) & ~1;  // == Lp + prow(k)
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int r = lane + 32 * j;
      if (r < k) yr[j] -= row[r] * xk;
      else if (r == k) yr[j] = xk;
    }
  }
#pragma unroll
  for (int j = 0; j < NS; ++j)
    if (lane + 32 * j < N) y[lane + 32 * j] = yr[j];
  __syncwarp();
}

// Solve A x = y for a symmetric positive SEMI-definite, consistent system (packed lower A):
// Cholesky with diagonal pivoting, stopped when the largest remaining diagonal falls below
// rtol * (largest initial diagonal); the unknowns outside the selected independent set are zero.
// L is built in permuted position space (packed), d = running diagonal, tmp = N scratch.
template <typename S>
__device__ __noinline__ void psd_solve_pivoted(const S* Ap, S* Lp, S* d, S* y, S* tmp, int* perm, int N, S rtol,
                                               int lane) {
  for (int i = lane; i < N; i += 32) { perm[i] = i; d[i] = Ap[pidx(i, i)]; }
  __syncwarp();
  int rank = 0;
  S dref = S(0);
  for (int k = 0; k < N; ++k) {
    S best = S(-1);
    int bp = k;
    for (int pos = k + lane; pos < N; pos += 32) {
      const S v = d[pos];
      if (v > best) { best = v; bp = pos; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const S ob = __shfl_xor_sync(FULL, best, o);
      const int op = __shfl_xor_sync(FULL, bp, o);
      if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
    }
    if (k == 0) dref = best;
    if (!(best > rtol * dref)) break;
    rank = k + 1;
    if (bp != k) {
      if (lane == 0) {
        const int tp = perm[k]; perm[k] = perm[bp]; perm[bp] = tp;
        const S td = d[k]; d[k] = d[bp]; d[bp] = td;
      }
      for (int c = lane; c < k; c += 32) {
        const S a = Lp[pidx(k, c)];
        Lp[pidx(k, c)] = Lp[pidx(bp, c)];
        Lp[pidx(bp, c)] = a;
      }
    }
    __syncwarp();
    const S ipiv = S(1) / sqrt_t(best);
    const int pk = perm[k];
    const S* rowk = Lp + pidx(k, 0);
    for (int pos = k + 1 + lane; pos < N; pos += 32) {
      const int pi = perm[pos];
      S v = (pi >= pk) ? Ap[pidx(pi, pk)] : Ap[pidx(pk, pi)];
      const S* row = Lp + prow(pos);
      {  // as in chol_packed: independent partial sums, two elements of either row per shared-memory access
        S a1 = S(0), a2 = S(0), a3 = S(0);
        int c = 0;
        for (; c + 3 < k; c += 4) {
          const auto p0 = ld2(row + c), q0 = ld2(rowk + c), p1 = ld2(row + c + 2), q1 = ld2(rowk + c + 2);
          v -= p0.x * q0.x; a1 -= p0.y * q0.y;
          a2 -= p1.x * q1.x; a3 -= p1.y * q1.y;
        }
        if (c + 1 < k) {
          const auto p0 = ld2(row + c), q0 = ld2(rowk + c);
          v -= p0.x * q0.x; a1 -= p0.y * q0.y;
          c += 2;
        }
        if (c < k) a2 -= row[c] * rowk[c];
        v = (v + a1) + (a2 + a3);
      }
      v *= ipiv;
      Lp[pidx(pos, k)] = v;
      d[pos] -= v * v;
    }
    if (lane == 0) d[k] = ipiv;  // position k is settled: its slot now holds 1 / L_kk
    __syncwarp();
  }
  for (int i = lane; i < N; i += 32) tmp[i] = y[perm[i]];
  __syncwarp();
  if (N <= 32) chol_solve<S, 1>(Lp, d, tmp, rank, lane);
  else if (N <= 64) chol_solve<S, 2>(Lp, d, tmp, rank, lane);
  else chol_solve<S, 3>(Lp, d, tmp, rank, lane);
  for (int i = lane; i < N; i += 32) y[perm[i]] = (i < rank) ? tmp[i] : S(0);
  __syncwarp();
}

// friction-pyramid rows of one point (rigid.py:478-489 rows 0-4): G v and G^T w
template <typename S>
__device__ __forceinline__ void pyr_G(S mu, const S* v, S* o) {
  o[0] = v[0] - mu * v[2]; o[1] = v[1] - mu * v[2]; o[2] = -v[0] - mu * v[2]; o[3] = -v[1] - mu * v[2]; o[4] = -v[2];
}
template <typename S>
__device__ __forceinline__ void pyr_GT(S mu, const S* w, S* o) {
  o[0] = w[0] - w[2]; o[1] = w[1] - w[3]; o[2] = -mu * (w[0] + w[1] + w[2] + w[3]) - w[4];
}

// min 1/2 x'Qx + q'x  s.t. pyramid constraints per point; N = 3 na.  Result in x.
// Primal-dual interior point, Mehrotra predictor-corrector.  Divisions are the expensive
// operation (float64 especially): 1/s and 1/z are formed once per iteration and every
// ratio below multiplies by them; the Cholesky factor carries 1/L_kk.
// Q is read once per iteration, in order (H <- Q), so it may live in global memory (rigid_qp_kernel); NS as in chol_solve.
template <typename S, int NS = 3>
__device__ __noinline__ int qp_pyramids(const S* Qp, S* Hp, S* vN, S* vM, int na, S mu_f, int lane, const S tol) {
  const int N = 3 * na, M = 5 * na;
  S* x = vN;
  S* q = vN + N;
  S* rd = vN + 2 * N;
  S* dxa = vN + 3 * N;
  S* dx = vN + 4 * N;
  S* dg = vN + 5 * N;
  S* xb = vN + 6 * N;  // best iterate so far (the answer if the iteration stalls at the resolution of S)
  S* s = vM;
  S* z = vM + M;
  S* rp = vM + 2 * M;
  S* dsa = vM + 3 * M;
  S* dza = vM + 4 * M;
  S* ds = vM + 5 * M;
  S* dz = vM + 6 * M;
  S* is = vM + 7 * M;  // 1 / s
  S* iz = vM + 8 * M;  // 1 / z
  for (int i = lane; i < N; i += 32) { x[i] = S(0); xb[i] = S(0); }
  for (int j = lane; j < M; j += 32) { s[j] = S(1); z[j] = S(1); }
  S best = S(1e30);
  bool converged = false;
  // Normalisation.  The constraints are a cone (G x <= 0), so x = (alpha sigma) y solves the problem iff y solves it
  // for Q' = alpha^2 Q, q' = alpha q / sigma.  With alpha^2 = 1 / mean(diag Q) and sigma = 0.3 max|alpha q| the solution,
  // its slacks and its multipliers are O(1) -- where the start s = z = 1 and the "1 +" floors of the merit assume them to
  // be.  In physical units (forces of 1e2-1e3 N, accelerations of 1e1-1e2 m/s^2) the interior-point iteration spent a
  // quarter of its steps growing s: 12.4 -> 8.6 iterations on the standing ErgoCub-like contact problems at equal accuracy
  // of Q x (scripts/experiments/qp_polish_proto.py; the tolerances were tightened by 10 to pay for the floors that no longer bind).
  S qm = S(0), dsum = S(0);
  for (int i = lane; i < N; i += 32) { qm = qmax(qm, abs_t(q[i])); dsum += Qp[pidx(i, i)]; }
  qm = warp_qmax(qm);
  dsum = warp_sum(dsum);
  const S alpha2 = (dsum > S(0)) ? S(N) / dsum : S(1);
  const S alpha = sqrt_t(alpha2);
  const S sigma = (qm > S(0)) ? S(0.3) * alpha * qm : S(1);
  {
    const S qs = alpha / sigma;
    for (int i = lane; i < N; i += 32) q[i] *= qs;
    qm *= qs;
  }
  __syncwarp();
  const int NP = prow(N);
  const S inv_M = S(1) / S(M);
  int it = 0, it_prog = 0;
  for (; it < QpTol<S>::max_iter; ++it) {
    // H <- Q (the factor of the previous iteration is dead); residuals
    for (int e = lane; e < NP; e += 32) Hp[e] = alpha2 * Qp[e];
    __syncwarp();
    sym_matvec(Hp, x, rd, N, lane);
    __syncwarp();
    S xQx = S(0), qx = S(0), xm = S(0), Qxm = S(0);
    for (int i = lane; i < N; i += 32) {
      xQx += x[i] * rd[i]; qx += q[i] * x[i];
      xm = qmax(xm, abs_t(x[i])); Qxm = qmax(Qxm, abs_t(rd[i]));
    }
    xQx = warp_sum(xQx); qx = warp_sum(qx); xm = warp_qmax(xm); Qxm = warp_qmax(Qxm);
    S rdn = S(0), rpn = S(0), sz = S(0);
    for (int a = lane; a < na; a += 32) {
      S gz[3], gx[5];
      pyr_GT(mu_f, z + 5 * a, gz);
      pyr_G(mu_f, x + 3 * a, gx);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const S v = rd[3 * a + d] + q[3 * a + d] + gz[d];
        rd[3 * a + d] = v;
        rdn = qmax(rdn, abs_t(v));
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const S sj = s[5 * a + j], zj = z[5 * a + j];
        const S v = gx[j] + sj;
        rp[5 * a + j] = v;
        rpn = qmax(rpn, abs_t(v));
        sz += sj * zj;
        is[5 * a + j] = S(1) / sj;
        iz[5 * a + j] = S(1) / zj;
      }
    }
    rdn = warp_qmax(rdn); rpn = warp_qmax(rpn); sz = warp_sum(sz);
    const S mu = sz * inv_M;
    // A non-finite iterate (the Hessian Q + G' diag(z/s) G loses definiteness to rounding once z/s spans the whole
    // exponent range, a few iterations past the resolution of S) must never become the answer.  The maxima above may
    // drop a NaN (compare + select), the SUMS cannot: x'Qx, q'x and s'z carry any NaN / Inf of x, s, z.
    {
      const S chk = abs_t(xQx) + abs_t(qx) + abs_t(sz);
      if (!(chk < S(sizeof(S) == 8 ? 1e290 : 1e30f))) break;  // keep the best finite iterate found so far
    }
    const S m_d = rdn / (S(1) + qm + Qxm), m_p = rpn / (S(1) + xm), m_g = mu / (S(1) + abs_t(S(0.5) * xQx + qx));
    const S merit = qmax(m_d, qmax(m_p, m_g));
    if (merit < best) {
      if (merit < S(0.5) * best) it_prog = it;
      best = merit;
      for (int i = lane; i < N; i += 32) xb[i] = x[i];
    }
    // stalled above the tolerance (a degenerate problem at the resolution of S): ten iterations without halving the
    // merit.  A healthy iteration gains a factor 3-10 per step; the stragglers used to run into max_iter, and ONE of
    // them holds its whole launch (60 iterations = 1.7 ms).
    if (it - it_prog >= 10) break;
    // converged, or the gap is far below the tolerance while a residual stalls at rounding level
    if (!(merit > tol)) { converged = true; break; }
    if (!(m_g > S(1e-3) * tol) || !(mu > S(0))) break;
    // H = Q + G' diag(z/s) G
    for (int a = lane; a < na; a += 32) {
      S w[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) w[j] = z[5 * a + j] * is[5 * a + j];
      const int r0 = 3 * a;
      Hp[pidx(r0, r0)] += w[0] + w[2];
      Hp[pidx(r0 + 1, r0 + 1)] += w[1] + w[3];
      Hp[pidx(r0 + 2, r0)] += -mu_f * (w[0] - w[2]);
      Hp[pidx(r0 + 2, r0 + 1)] += -mu_f * (w[1] - w[3]);
      Hp[pidx(r0 + 2, r0 + 2)] += mu_f * mu_f * (w[0] + w[1] + w[2] + w[3]) + w[4];
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) dg[i] = Hp[pidx(i, i)];
    __syncwarp();
    chol_packed(Hp, dg, N, lane);
    // predictor: r_c = s z
    for (int a = lane; a < na; a += 32) {
      S t[5], g[3];
#pragma unroll
      for (int j = 0; j < 5; ++j) t[j] = z[5 * a + j] * (rp[5 * a + j] - s[5 * a + j]) * is[5 * a + j];
      pyr_GT(mu_f, t, g);
#pragma unroll
      for (int d = 0; d < 3; ++d) dxa[3 * a + d] = -(rd[3 * a + d] + g[d]);
    }
    __syncwarp();
    chol_solve<S, NS>(Hp, dg, dxa, N, lane);
    // largest step keeping s, z positive: alpha = 1 / max_j(-ds_j / s_j, -dz_j / z_j)
    S rmax = S(1);
    for (int a = lane; a < na; a += 32) {
      S g[5];
      pyr_G(mu_f, dxa + 3 * a, g);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const S sj = s[5 * a + j], zj = z[5 * a + j];
        const S dsj = -rp[5 * a + j] - g[j];
        const S dzj = -(sj * zj + zj * dsj) * is[5 * a + j];
        dsa[5 * a + j] = dsj; dza[5 * a + j] = dzj;
        rmax = qmax(rmax, qmax(-dsj * is[5 * a + j], -dzj * iz[5 * a + j]));
      }
    }
    rmax = warp_qmax(rmax);
    const S amax = S(1) / rmax;  // rmax >= 1
    __syncwarp();  // ds_a / dz_a were written point-wise, read element-wise below
    S mua = S(0);
    for (int j = lane; j < M; j += 32) mua += (s[j] + amax * dsa[j]) * (z[j] + amax * dza[j]);
    mua = warp_sum(mua) * inv_M;
    S sg = mua / mu;
    sg = sg * sg * sg;
    // corrector: r_c = s z + ds_a dz_a - sigma mu
    for (int a = lane; a < na; a += 32) {
      S t[5], g[3];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const S sj = s[5 * a + j], zj = z[5 * a + j];
        const S rc = sj * zj + dsa[5 * a + j] * dza[5 * a + j] - sg * mu;
        t[j] = (zj * rp[5 * a + j] - rc) * is[5 * a + j];
      }
      pyr_GT(mu_f, t, g);
#pragma unroll
      for (int d = 0; d < 3; ++d) dx[3 * a + d] = -(rd[3 * a + d] + g[d]);
    }
    __syncwarp();
    chol_solve<S, NS>(Hp, dg, dx, N, lane);
    S rm2 = S(0);
    for (int a = lane; a < na; a += 32) {
      S g[5];
      pyr_G(mu_f, dx + 3 * a, g);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const S sj = s[5 * a + j], zj = z[5 * a + j];
        const S rc = sj * zj + dsa[5 * a + j] * dza[5 * a + j] - sg * mu;
        const S dsj = -rp[5 * a + j] - g[j];
        const S dzj = -(rc + zj * dsj) * is[5 * a + j];
        ds[5 * a + j] = dsj; dz[5 * a + j] = dzj;
        rm2 = qmax(rm2, qmax(-dsj * is[5 * a + j], -dzj * iz[5 * a + j]));
      }
    }
    rm2 = warp_qmax(rm2);
    // alpha = min(1, 0.99 / rm2).  (A fraction to the boundary that grows with the closing gap, max(0.99, 1 - mu), saves
    // 0.3 iterations on average but doubles the longest ones -- 14 -> 20 -- and the launch waits for those: not adopted.)
    const S al = (rm2 > S(0.99)) ? S(0.99) / rm2 : S(1);
    __syncwarp();
    for (int i = lane; i < N; i += 32) x[i] += al * dx[i];
    for (int j = lane; j < M; j += 32) { s[j] += al * ds[j]; z[j] += al * dz[j]; }
    __syncwarp();
  }
  __syncwarp();
  {
    const S xs = alpha * sigma;  // back to physical units
    for (int i = lane; i < N; i += 32) x[i] = xb[i] * xs;
  }
  __syncwarp();
  // bit 16: the iteration ended without meeting the tolerance (iteration limit, stall at the resolution of S, or a
  // non-finite iterate): the best iterate is returned.  The reference ignores qpax's flag (rigid.py:359-362).
  return it | ((converged || best <= S(100) * tol) ? 0 : 0x10000);
}

// the instance whose substitution loops fit the problem; both cascades (monolithic / split) go through here, so that they
// run the same code on the same numbers (tests/test_gpu_rigid.py: bit-identical results)
template <typename S>
__device__ __forceinline__ int qp_solve(const S* Qp, S* Hp, S* vN, S* vM, int na, S mu_f, int lane, const S tol) {
  const int N = 3 * na;
  if (N <= 32) return qp_pyramids<S, 1>(Qp, Hp, vN, vM, na, mu_f, lane, tol);
  if (N <= 64) return qp_pyramids<S, 2>(Qp, Hp, vN, vM, na, mu_f, lane, tol);
  return qp_pyramids<S, 3>(Qp, Hp, vN, vM, na, mu_f, lane, tol);
}

// ------------------------------------------------------------------------------------
// the kernel: one warp per environment
// ------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(32 * RIGID_MAX_WARPS, 1) rigid_step_kernel(const Params<T> P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm_cst = reinterpret_cast<T*>(smem_raw);
  const int nL = P.nL, n = P.n, nc = P.nc;
  T* sm_pt = sm_cst + (size_t)nL * CREC;
  const size_t pt_words = ((size_t)nc * 3 + 3) & ~size_t(3);
  int* sm_itab = reinterpret_cast<int*>(sm_pt + pt_words);
  const size_t itab_words = ((size_t)P.itab_words + 3) & ~size_t(3);
  unsigned char* ws_base = reinterpret_cast<unsigned char*>(sm_itab + itab_words);

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
  const int* anc = sm_itab + P.o_anc;        // [nL * depth] ancestors root-first, incl. the link itself
  const int* ldepth = sm_itab + P.o_ldepth;  // [nL]
  const int depth = P.depth;

  const int lane = threadIdx.x & 31;
  const int wrp = threadIdx.x >> 5;
  const int cap = P.na_cap;
  // RelaxedRigidContacts never runs the interior-point method either: the layout of a resuming launch (no solver vectors)
  const RigidLayout L = rigid_layout<T, S>(nL, nc, depth, cap, P.contact_model == 3 ? 2 : P.qp_mode);
  unsigned char* wb = ws_base + (size_t)wrp * L.total;
  T* ws = reinterpret_cast<T*>(wb + L.links);
  T* pts = reinterpret_cast<T*>(wb + L.pts);
  T* ainv = reinterpret_cast<T*>(wb + L.ainv);
  T* ucol = reinterpret_cast<T*>(wb + L.ucol);
  S* Qp = reinterpret_cast<S*>(wb + L.Qp);
  S* Hp = reinterpret_cast<S*>(wb + L.Hp);
  S* vN = reinterpret_cast<S*>(wb + L.vecN);
  S* vM = reinterpret_cast<S*>(wb + L.vecM);
  int* aidx = reinterpret_cast<int*>(wb + L.ints);  // [nc] compact index of an active point or -1
  int* alist = aidx + nc;                           // [nc] point index of compact index
  int* clist = alist + nc;                          // [nL] links that carry active points
  int* perm = clist + nL;                           // [3 nc] pivot order of the impact solve
  const T dt = P.dt;
  const bool relaxed = P.contact_model == 3;  // RelaxedRigidContacts: same assembly, a linear solve, no impact
  const long long stride = (long long)gridDim.x * P.envs_per_block;

  // work items: every environment of the batch, or the list produced by the previous level of
  // the cascade (env | impact_only << 31)
  const long long total = P.work_list ? (long long)(*P.work_count) : P.B;
  for (long long it0 = (long long)blockIdx.x * P.envs_per_block; it0 < total; it0 += stride) {
    const long long item_idx = it0 + wrp;
    if (item_idx >= total) continue;  // whole warp: no block-level barrier inside the loop
    long long env = item_idx;
    bool impact_only = false;
    if (P.work_list) {
      const int item = P.work_list[item_idx];
      env = item & 0x7fffffff;
      impact_only = item < 0;
    }
    int* qp_hdr = P.qp_mode ? reinterpret_cast<int*>(P.qp_buf + item_idx * P.qp_stride) : nullptr;
    if (P.qp_mode == 1) {
      if (lane == 0) { qp_hdr[0] = 0; qp_hdr[1] = 0; qp_hdr[2] = (int)env; }  // nothing to solve unless set below
      if (impact_only) continue;
      __syncwarp();
    }
    if (P.dbg && lane == 0 && P.qp_mode != 1) atomicAdd(P.dbg + (impact_only ? 5 : 4), 1ull);

    // ============================================================== input state
    // (impact_only: the pre-impact result of this step, stored by the previous level)
    BaseState<T> b;
    if (!impact_only) {
      for (int i = 1 + lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        ri[O_S] = P.s[env * n + (i - 1)];
        ri[O_SD] = P.sd[env * n + (i - 1)];
        ri[O_TREF] = P.tau ? P.tau[env * n + (i - 1)] : T(0);
      }
      const T* q = P.q + env * 4;
      const T qr[4] = {q[0], q[1], q[2], q[3]};
      ldn<3>(P.p + env * 3, b.p);
      ldn<3>(P.vlin + env * 3, b.vlin);
      ldn<3>(P.omega + env * 3, b.w);
      const T nrm = sqrt_t(qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]);
      const T inv = T(1) / (nrm + Lim<T>::eps() * (nrm == T(0) ? T(1) : T(0)));  // api/data.py:283-285
#pragma unroll
      for (int k = 0; k < 4; ++k) b.qn[k] = qr[k] * inv;
      quat_to_dcm(b.qn, b.R);
    } else {
      for (int i = 1 + lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        ri[O_S] = P.s_o[env * n + (i - 1)];
        ri[O_SD] = P.sd_o[env * n + (i - 1)];
      }
      ldn<4>(P.q_o + env * 4, b.qn);
      ldn<3>(P.p_o + env * 3, b.p);
      ldn<3>(P.vlin_o + env * 3, b.vlin);
      ldn<3>(P.omega_o + env * 3, b.w);
      quat_to_dcm(b.qn, b.R);
    }

    auto write_base_record = [&](const BaseState<T>& bs) {
      if (lane == 0) {
        T* r0 = ws;
        stn<9>(r0 + O_R, bs.R);
        stn<3>(r0 + O_P, bs.p);
        T v0[6], t[3];
        cross3(bs.w, bs.p, t);
        v0[0] = bs.vlin[0] + t[0]; v0[1] = bs.vlin[1] + t[1]; v0[2] = bs.vlin[2] + t[2];
        v0[3] = bs.w[0]; v0[4] = bs.w[1]; v0[5] = bs.w[2];
        stn<6>(r0 + O_V, v0);
      }
    };

    // joint transforms (relative) into the records, then the FK / velocity chain
    auto kinematics = [&](const BaseState<T>& bs, const bool emit_adjoints) {
      write_base_record(bs);
      for (int i = 1 + lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        T Rrel[9], trel[3];
        joint_rel_transform(P, P.flags, sm_cst + (size_t)i * CREC, jtypes[i], i, ri[O_S], Rrel, trel);
        stn<9>(ri + O_R, Rrel);
        stn<3>(ri + O_P, trel);
        if (emit_adjoints && P.iXl) {
          T X[36];
          inverse_adjoint(X, Rrel, trel);
          stg_vec<36>(P.iXl + (env * nL + i) * 36, X);
        }
      }
      for (int l = 1; l <= depth; ++l) {
        __syncwarp();
        const int e = lvl_start[l + 1];
        for (int idx = lvl_start[l] + lane; idx < e; idx += 32) {
          const int i = lvl_links[idx];
          const T* rp = ws + (size_t)parent[i] * REC;
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
          const int jt = jtypes[i];
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
    };

    // collidable points of the current kinematics: lever arms, active set, and either the
    // bias of the contact-acceleration equation (phase A) or the point velocity (impact)
    //   returns the number of active points; fills aidx / alist / clist (ncl links)
    int ncl = 0;
    auto contact_points = [&](const bool impact) -> int {
      for (int k = lane; k < nc; k += 32) {
        const int bi = pt_body[k];
        const T* rb = ws + (size_t)bi * REC;
        T R[9], p[3], v[6];
        ldn<9>(rb + O_R, R);
        ldn<3>(rb + O_P, p);
        ldn<6>(rb + O_V, v);
        T Lp[3], d[3], pc[3], pd[3];
        ldn<3>(sm_pt + 3 * k, Lp);
        mat3_vec(R, Lp, d);
        pc[0] = p[0] + d[0]; pc[1] = p[1] + d[1]; pc[2] = p[2] + d[2];
        cross3(v + 3, d, pd);
        pd[0] += v[0]; pd[1] += v[1]; pd[2] += v[2];
        const T delta = max_t(T(0), P.h_terrain - pc[2]);  // rbda/contacts/common.py:25-63, FlatTerrain
        const bool act = pt_enabled[k] && (delta > T(0));
        T* pw = pts + (size_t)k * RPT;
        stn<3>(pw + RP_LEV, d);
        T bv[3];
        if (impact) {
          bv[0] = pd[0]; bv[1] = pd[1]; bv[2] = pd[2];
        } else {
          // pdot as the reference's contact code sees it: from the cached link velocities
          T ps[3] = {pd[0], pd[1], pd[2]};
          if (P.Vin) {
            T V[6];
            ldg_vec6(P.Vin + (env * nL + bi) * 6, V);
            cross3(V + 3, pc, ps);
            ps[0] += V[0]; ps[1] += V[1]; ps[2] += V[2];
          }
          const T ddot = -ps[2];
          cross3(v + 3, ps, bv);                       // omega x pdot  (api/contact.py:470-477 term)
          if (!relaxed) {
            bv[2] -= P.K * delta + P.D * ddot;         // Baumgarte, n = z (rigid.py:527-539)
          } else {
            // RelaxedRigidContacts._regularizers (relaxed_rigid.py:533-653), FlatTerrain: the position in
            // the constraint frame is (0, 0, -delta); impedance per COMPONENT as the reference evaluates it
            // (imp_x = |pos| / width is a 3-vector); K, D derive from the time constant (:585-591)
            const T xmax = P.rx_dmax, xmin = P.rx_dmin;
            const T Kc = T(1) / ((xmax * P.rx_tc * P.rx_zeta) * (xmax * P.rx_tc * P.rx_zeta));
            const T Dc = T(2) / (xmax * P.rx_tc);
            const T mu2 = P.mu * P.mu;
            const T imass = T(1) / sm_cst[(size_t)bi * CREC + C_MASS];
            const T posv[3] = {T(0), T(0), -delta};
            T rr3[3];
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) {
              const T ix = abs_t(posv[c3]) / P.rx_width;
              T iy;
              if (ix < P.rx_mid) iy = pow_t(ix, P.rx_pow) / pow_t(P.rx_mid, P.rx_pow - T(1));
              else iy = T(1) - pow_t(max_t(T(1) - ix, T(0)), P.rx_pow) / pow_t(T(1) - P.rx_mid, P.rx_pow - T(1));
              T xi = xmin + iy * (xmax - xmin);
              xi = min_t(max_t(xi, xmin), xmax);
              if (ix > T(1)) xi = xmax;
              const T aref = -(Dc * ps[c3] + Kc * xi * posv[c3]);
              bv[c3] -= aref;                          // b = J nu_dot_free + J_dot nu - a_ref (:386-391)
              rr3[c3] = (T(2) * mu2 * (T(1) - xi) / (xi + T(1e-12))) * (T(1) + mu2) * imass;
            }
            stn<3>(pw + RP_F, rr3);                    // diagonal regulariser, consumed right after delassus()
          }
        }
        stn<3>(pw + RP_B, bv);
        aidx[k] = act ? 1 : 0;
      }
      __syncwarp();
      int base = 0;
      for (int k0 = 0; k0 < nc; k0 += 32) {
        const int k = k0 + lane;
        const bool a = (k < nc) && (aidx[k] != 0);
        const unsigned m = __ballot_sync(FULL, a);
        if (k < nc) {
          const int ci = a ? base + __popc(m & ((1u << lane) - 1u)) : -1;
          aidx[k] = ci;
          if (a) alist[ci] = k;
        }
        base += __popc(m);
      }
      __syncwarp();
      // links that carry active points (ordered by link index)
      int nl = 0;
      for (int i0 = 0; i0 < nL; i0 += 32) {
        const int i = i0 + lane;
        bool has = false;
        if (i < nL) {
          const int e = pt_start[i + 1];
          for (int kk = pt_start[i]; kk < e; ++kk) has = has || (aidx[pt_idx[kk]] >= 0);
        }
        const unsigned m = __ballot_sync(FULL, has);
        if (has) clist[nl + __popc(m & ((1u << lane) - 1u))] = i;
        nl += __popc(m);
      }
      ncl = nl;
      __syncwarp();
      return base;
    };

    // per-link initialisation of the ABA: inertia about the link origin in world axes, and
    // (bias == true) velocity-product bias force, c_i, resultant joint torque
    auto link_init = [&](const bool bias) {
      for (int i = lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        const T* c = sm_cst + (size_t)i * CREC;
        T R[9], p[3], v[6];
        ldn<9>(ri + O_R, R);
        ldn<3>(ri + O_P, p);
        ldn<6>(ri + O_V, v);
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
        T pA[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        T cc[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        T tau = T(0);
        if (bias) {
          T fe[3] = {T(0), T(0), T(0)}, ne[3] = {T(0), T(0), T(0)};
          if (P.fext) add_external_wrench(P.fext_repr, P.fext + (env * nL + i) * 6, R, p, fe, ne);
          T fI[3], nI[3], t[3];
          cross3(v + 3, cw, t);
          fI[0] = mass * (v[0] + t[0]); fI[1] = mass * (v[1] + t[1]); fI[2] = mass * (v[2] + t[2]);
          sym3_vec(Dw, v + 3, nI);
          cross3(cw, v, t);
          nI[0] += mass * t[0]; nI[1] += mass * t[1]; nI[2] += mass * t[2];
          cross3(v + 3, fI, pA);
          cross3(v, fI, pA + 3);
          cross3_add(v + 3, nI, pA + 3);
          pA[0] -= fe[0]; pA[1] -= fe[1]; pA[2] -= fe[2];
          pA[3] -= ne[0]; pA[4] -= ne[1]; pA[5] -= ne[2];
          if (i > 0) {
            const int jt = jtypes[i];
            const T sdi = ri[O_SD];
            T aw[3];
            ldn<3>(ri + O_AX, aw);
            const T vJ[3] = {sdi * aw[0], sdi * aw[1], sdi * aw[2]};
            if (jt == 1) {
              cross3(v, vJ, cc);
              cross3(v + 3, vJ, cc + 3);
            } else {
              cross3(v + 3, vJ, cc);
            }
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
            const T tt = ri[O_TREF] + tfr + tlim;
            const T av = abs_t(sdi);
            T lim;
            if (av <= P.w_th) lim = P.tau_max;
            else if (av <= P.w_max) lim = P.tau_max * (T(1) - (av - P.w_th) / (P.w_max - P.w_th));
            else lim = T(0);
            tau = min_t(max_t(tt, -lim), lim);
          }
        }
        T IA[21];
        IA[0] = mass; IA[1] = T(0); IA[2] = T(0); IA[3] = mass; IA[4] = T(0); IA[5] = mass;
        IA[6] = T(0);            IA[7] = mass * cw[2];   IA[8] = -mass * cw[1];
        IA[9] = -mass * cw[2];   IA[10] = T(0);          IA[11] = mass * cw[0];
        IA[12] = mass * cw[1];   IA[13] = -mass * cw[0]; IA[14] = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) IA[15 + k] = Dw[k];
        stn<21>(ri + O_IA, IA);
        stn<6>(ri + O_PA, pA);
        stn<6>(ri + O_C, cc);
        ri[O_TAU] = tau;
      }
    };

    // ABA pass 2 (rbda/aba.py:173-231) + inverse of the base articulated inertia
    auto pass2 = [&]() {
      for (int l = depth; l >= 1; --l) {
        __syncwarp();
        const int e = lvl_start[l + 1];
        for (int idx = lvl_start[l] + lane; idx < e; idx += 32) {
          const int i = lvl_links[idx];
          T* ri = ws + (size_t)i * REC;
          T A[6], Bm[9], D[6], pA[6];
          ldn<6>(ri + O_IA, A);
          ldn<9>(ri + O_IB, Bm);
          ldn<6>(ri + O_ID, D);
          ldn<6>(ri + O_PA, pA);
          const int ce = child_start[i + 1];
          for (int cc = child_start[i]; cc < ce; ++cc) {
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
          T aw[3], cI[6], r[3];
          ldn<3>(ri + O_AX, aw);
          ldn<6>(ri + O_C, cI);
          ldn<3>(ri + O_RR, r);
          const int jt = jtypes[i];
          T Ul[3], Ua[3], d, u;
          const T tau = ri[O_TAU];
          if (jt == 1) {
            mat3_vec(Bm, aw, Ul);
            sym3_vec(D, aw, Ua);
            d = dot3(aw, Ua);
            u = tau - dot3(aw, pA + 3);
          } else {
            sym3_vec(A, aw, Ul);
            mat3T_vec(Bm, aw, Ua);
            d = dot3(aw, Ul);
            u = tau - dot3(aw, pA);
          }
          const T dinv = T(1) / d;
          stn<3>(ri + O_U, Ul);
          stn<3>(ri + O_U + 3, Ua);
          ri[O_DINV] = dinv;
          ri[O_UU] = u;
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
          cross3_add(r, pa, pa + 3);
          stn<6>(ri + O_IA, A);
          stn<9>(ri + O_IB, B2);
          stn<6>(ri + O_ID, D2);
          stn<6>(ri + O_PA, pa);
        }
      }
      __syncwarp();
      // base: total articulated inertia, its inverse (one column per lane), total bias
      if (lane < 6) {
        T* r0 = ws;
        T A[6], Bm[9], D[6];
        ldn<6>(r0 + O_IA, A);
        ldn<9>(r0 + O_IB, Bm);
        ldn<6>(r0 + O_ID, D);
        const int ce = child_start[1];
        for (int cc = child_start[0]; cc < ce; ++cc) {
          const T* rc = ws + (size_t)child_idx[cc] * REC;
#pragma unroll
          for (int k = 0; k < 6; ++k) A[k] += rc[O_IA + k];
#pragma unroll
          for (int k = 0; k < 9; ++k) Bm[k] += rc[O_IB + k];
#pragma unroll
          for (int k = 0; k < 6; ++k) D[k] += rc[O_ID + k];
        }
        T Mx[6][6];
        Mx[0][0] = A[0]; Mx[0][1] = A[1]; Mx[0][2] = A[2]; Mx[1][1] = A[3]; Mx[1][2] = A[4]; Mx[2][2] = A[5];
        Mx[1][0] = A[1]; Mx[2][0] = A[2]; Mx[2][1] = A[4];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int bb = 0; bb < 3; ++bb) { Mx[a][3 + bb] = Bm[3 * a + bb]; Mx[3 + bb][a] = Bm[3 * a + bb]; }
        Mx[3][3] = D[0]; Mx[3][4] = D[1]; Mx[3][5] = D[2]; Mx[4][4] = D[3]; Mx[4][5] = D[4]; Mx[5][5] = D[5];
        Mx[4][3] = D[1]; Mx[5][3] = D[2]; Mx[5][4] = D[4];
        T e[6], x[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) e[k] = (k == lane) ? T(-1) : T(0);
        solve6_spd_neg(Mx, e, x);  // x = M^-1 e_lane
#pragma unroll
        for (int k = 0; k < 6; ++k) ainv[6 * k + lane] = x[k];
      }
      __syncwarp();
    };

    // total bias force at the base (own + children), all lanes
    auto base_bias = [&](T* pA0) {
      ldn<6>(ws + O_PA, pA0);
      const int ce = child_start[1];
      for (int cc = child_start[0]; cc < ce; ++cc) {
        const T* rc = ws + (size_t)child_idx[cc] * REC;
#pragma unroll
        for (int k = 0; k < 6; ++k) pA0[k] += rc[O_PA + k];
      }
    };
    auto ainv_neg_mul = [&](const T* f, T* a) {  // a = -Ainv f
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T acc = T(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += ainv[6 * r + c] * f[c];
        a[r] = -acc;
      }
    };

    // articulated-body response of the whole tree to point forces (RP_F of the active
    // points): delta acceleration of every link into O_C, of every joint into O_TAU.
    // Uses U_i, 1/d_i, r_i, axes and Ainv of the last pass2().
    auto tree_response = [&]() {
      for (int i = lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        T w[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        const int e = pt_start[i + 1];
        for (int kk = pt_start[i]; kk < e; ++kk) {
          const int k = pt_idx[kk];
          if (aidx[k] < 0) continue;
          const T* pw = pts + (size_t)k * RPT;
          T f[3], lev[3];
          ldn<3>(pw + RP_F, f);
          ldn<3>(pw + RP_LEV, lev);
          w[0] -= f[0]; w[1] -= f[1]; w[2] -= f[2];
          T t[3];
          cross3(lev, f, t);
          w[3] -= t[0]; w[4] -= t[1]; w[5] -= t[2];
        }
        stn<6>(ri + O_PA, w);
      }
      for (int l = depth; l >= 1; --l) {
        __syncwarp();
        const int e = lvl_start[l + 1];
        for (int idx = lvl_start[l] + lane; idx < e; idx += 32) {
          const int i = lvl_links[idx];
          T* ri = ws + (size_t)i * REC;
          T pA[6];
          ldn<6>(ri + O_PA, pA);
          const int ce = child_start[i + 1];
          for (int cc = child_start[i]; cc < ce; ++cc) {
            const T* rc = ws + (size_t)child_idx[cc] * REC;
#pragma unroll
            for (int k = 0; k < 6; ++k) pA[k] += rc[O_PA + k];
          }
          T aw[3], U[6], r[3];
          ldn<3>(ri + O_AX, aw);
          ldn<6>(ri + O_U, U);
          ldn<3>(ri + O_RR, r);
          const T u = (jtypes[i] == 1) ? -dot3(aw, pA + 3) : -dot3(aw, pA);
          ri[O_UU] = u;
          const T ud = u * ri[O_DINV];
#pragma unroll
          for (int k = 0; k < 6; ++k) pA[k] += U[k] * ud;
          cross3_add(r, pA, pA + 3);
          stn<6>(ri + O_PA, pA);
        }
      }
      __syncwarp();
      if (lane == 0) {
        T pA0[6], a0[6];
        base_bias(pA0);
        ainv_neg_mul(pA0, a0);
        stn<6>(ws + O_C, a0);
      }
      for (int l = 1; l <= depth; ++l) {
        __syncwarp();
        const int e = lvl_start[l + 1];
        for (int idx = lvl_start[l] + lane; idx < e; idx += 32) {
          const int i = lvl_links[idx];
          T* ri = ws + (size_t)i * REC;
          const T* rp = ws + (size_t)parent[i] * REC;
          T ap[6], r[3], U[6], aw[3];
          ldn<6>(rp + O_C, ap);
          ldn<3>(ri + O_RR, r);
          ldn<6>(ri + O_U, U);
          ldn<3>(ri + O_AX, aw);
          T a[6];
          cross3(ap + 3, r, a);
          a[0] += ap[0]; a[1] += ap[1]; a[2] += ap[2];
          a[3] = ap[3]; a[4] = ap[4]; a[5] = ap[5];
          const T sdd = (ri[O_UU] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * ri[O_DINV];
          if (jtypes[i] == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
          else { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
          stn<6>(ri + O_C, a);
          ri[O_TAU] = sdd;
        }
      }
      __syncwarp();
    };

    // Delassus matrix of the active points, packed lower triangle, one column per lane:
    // response to a unit force e_d at point a, evaluated at every active point
    auto delassus = [&](const int na, const S reg) {
      const int N = 3 * na;
      for (int col = lane; col < N; col += 32) {
        const int a = col / 3, d = col - 3 * a;
        const int k = alist[a];
        const int lk = pt_body[k];
        T lev[3];
        ldn<3>(pts + (size_t)k * RPT + RP_LEV, lev);
        T pa[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
        pa[d] = T(-1);
        {
          T e[3] = {T(0), T(0), T(0)}, t[3];
          e[d] = T(1);
          cross3(lev, e, t);
          pa[3] = -t[0]; pa[4] = -t[1]; pa[5] = -t[2];
        }
        const int dk = ldepth[lk];
        int i = lk;
        for (int t = dk - 1; t >= 0; --t) {
          const T* ri = ws + (size_t)i * REC;
          T aw[3], U[6], r[3];
          ldn<3>(ri + O_AX, aw);
          ldn<6>(ri + O_U, U);
          ldn<3>(ri + O_RR, r);
          const T u = (jtypes[i] == 1) ? -dot3(aw, pa + 3) : -dot3(aw, pa);
          ucol[t * 32 + lane] = u;
          const T ud = u * ri[O_DINV];
#pragma unroll
          for (int q = 0; q < 6; ++q) pa[q] += U[q] * ud;
          cross3_add(r, pa, pa + 3);
          i = parent[i];
        }
        T a0[6];
        ainv_neg_mul(pa, a0);
        for (int ci = 0; ci < ncl; ++ci) {
          const int j = clist[ci];
          const int dj = ldepth[j];
          T acc[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc[q] = a0[q];
          for (int t = 0; t < dj; ++t) {
            const int li = anc[j * depth + t];
            const T* ri = ws + (size_t)li * REC;
            T aw[3], U[6], r[3];
            ldn<3>(ri + O_AX, aw);
            ldn<6>(ri + O_U, U);
            ldn<3>(ri + O_RR, r);
            cross3_add(acc + 3, r, acc);
            const T ut = (t < dk && anc[lk * depth + t] == li) ? ucol[t * 32 + lane] : T(0);
            const T sdd = (ut - (U[0] * acc[0] + U[1] * acc[1] + U[2] * acc[2] + U[3] * acc[3] + U[4] * acc[4] + U[5] * acc[5])) * ri[O_DINV];
            if (jtypes[li] == 1) { acc[3] += sdd * aw[0]; acc[4] += sdd * aw[1]; acc[5] += sdd * aw[2]; }
            else { acc[0] += sdd * aw[0]; acc[1] += sdd * aw[1]; acc[2] += sdd * aw[2]; }
          }
          const int e = pt_start[j + 1];
          for (int kk = pt_start[j]; kk < e; ++kk) {
            const int k2 = pt_idx[kk];
            const int a2 = aidx[k2];
            if (a2 < 0) continue;
            T lev2[3], val[3];
            ldn<3>(pts + (size_t)k2 * RPT + RP_LEV, lev2);
            cross3(acc + 3, lev2, val);
            val[0] += acc[0]; val[1] += acc[1]; val[2] += acc[2];
#pragma unroll
            for (int d2 = 0; d2 < 3; ++d2) {
              const int row = 3 * a2 + d2;
              if (row >= col) Qp[pidx(row, col)] = S(val[d2]) + ((row == col) ? reg : S(0));
            }
          }
        }
      }
      __syncwarp();
    };

    if (!impact_only) {
    // ============================================================== phase A: state at t
    B200SIM_RIGID_MARK(1);
    // resuming launch of a split level: the assembling launch saved its workspace next to the contact problem
    const bool restored = QpSaveState<T>::value && (P.qp_mode == 2) && (qp_hdr[0] > 0);
    const size_t state_bytes = L.ucol - L.links;  // link records | point records | base inverse inertia
    unsigned char* saved = P.qp_mode ? reinterpret_cast<unsigned char*>(qp_hdr) + qp_record_problem_bytes<S>(cap) : nullptr;
    int na = 0;
    if (restored) {
      const int4* src = reinterpret_cast<const int4*>(saved);
      int4* dst = reinterpret_cast<int4*>(wb + L.links);
      const int n16 = (int)(state_bytes / 16);
      // eight loads in flight per lane: one warp pulls its 13 KB in ~4 memory round trips instead of 26
      for (int i0 = 0; i0 < n16; i0 += 256) {
        int4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + 32 * u + lane;
          if (i < n16) v[u] = __ldcs(src + i);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + 32 * u + lane;
          if (i < n16) dst[i] = v[u];
        }
      }
      const int* isrc = reinterpret_cast<const int*>(saved + state_bytes);
      for (int i = lane; i < 2 * nc + nL; i += 32) aidx[i] = isrc[i];  // aidx | alist | clist are contiguous
      na = qp_hdr[0];
      ncl = qp_hdr[3];
      __syncwarp();
    } else {
    kinematics(b, false);
    na = contact_points(false);
    B200SIM_RIGID_MARK(2);
    if (na > cap) {  // more active points than this level's workspace holds: next level, untouched
      if (lane == 0 && P.qp_mode != 2) over_push(P, (int)env, 0);  // (the assembling launch already pushed it)
      continue;
    }
    if (P.qp_mode == 1 && na == 0) continue;
    link_init(true);
    pass2();
    // free acceleration (ABA pass 3, rbda/aba.py:233-288)
    if (lane == 0) {
      T pA0[6], a0[6];
      base_bias(pA0);
      ainv_neg_mul(pA0, a0);
      stn<6>(ws + O_V, a0);
    }
    for (int l = 1; l <= depth; ++l) {
      __syncwarp();
      const int e = lvl_start[l + 1];
      for (int idx = lvl_start[l] + lane; idx < e; idx += 32) {
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
        const T sdd = (ri[O_UU] - (U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5])) * ri[O_DINV];
        if (jtypes[i] == 1) { a[3] += sdd * aw[0]; a[4] += sdd * aw[1]; a[5] += sdd * aw[2]; }
        else { a[0] += sdd * aw[0]; a[1] += sdd * aw[1]; a[2] += sdd * aw[2]; }
        stn<6>(ri + O_V, a);
        ri[O_SDD] = sdd;
      }
    }
    __syncwarp();
    }  // !restored

    B200SIM_RIGID_MARK(3);
    T a0t[6];  // base acceleration in F_0 (gravity-shifted), free + contact response
    ldn<6>(ws + O_V, a0t);
    if (na > 0) {
      // q = J_dot nu + J nu_dot_free - baumgarte  (rigid.py:316-343)
      S* q = vN + 3 * na;
      for (int a = lane; a < na; a += 32) {
        const int k = alist[a];
        const T* rb = ws + (size_t)pt_body[k] * REC;
        const T* pw = pts + (size_t)k * RPT;
        T acc[6], lev[3], bv[3], val[3];
        ldn<6>(rb + O_V, acc);
        ldn<3>(pw + RP_LEV, lev);
        ldn<3>(pw + RP_B, bv);
        cross3(acc + 3, lev, val);
        q[3 * a + 0] = S(val[0] + acc[0] + bv[0]);
        q[3 * a + 1] = S(val[1] + acc[1] + bv[1]);
        q[3 * a + 2] = S(val[2] + acc[2] + P.g + bv[2]);
      }
      if (!relaxed && P.qp_mode == 2) {
        // the contact forces were found by rigid_qp_kernel from the problem the assembling launch left in the record
        const S* xg = reinterpret_cast<const S*>(qp_hdr + 4) + 3 * cap;
        for (int i = lane; i < 3 * na; i += 32) vN[i] = xg[i];
        __syncwarp();
      } else if (!relaxed) {
        delassus(na, S(P.reg));
        B200SIM_RIGID_MARK(4);
        if (P.qp_mode == 1) {
          S* qg = reinterpret_cast<S*>(qp_hdr + 4);
          S* Qg = qg + 6 * cap;
          const int N = 3 * na, NP = prow(N);
          for (int i = lane; i < N; i += 32) qg[i] = q[i];
          for (int e = lane; e < NP; e += 32) Qg[e] = Qp[e];
          if (QpSaveState<T>::value) {
            const int4* src = reinterpret_cast<const int4*>(wb + L.links);
            int4* dst = reinterpret_cast<int4*>(saved);
            for (int i = lane; i < (int)(state_bytes / 16); i += 32) dst[i] = src[i];
            int* idst = reinterpret_cast<int*>(saved + state_bytes);
            for (int i = lane; i < 2 * nc + nL; i += 32) idst[i] = aidx[i];
          }
          if (lane == 0) { qp_hdr[0] = na; qp_hdr[3] = ncl; }
          __syncwarp();
          continue;
        }
#ifdef B200SIM_RIGID_DEBUG
        if (P.dbg && lane == 0 && env == 0) {  // dump of the contact problem of environment 0 (diagnostic builds only)
          double* D = reinterpret_cast<double*>(P.dbg + 1600);
          D[0] = (double)na;
          for (int k = 0; k < 6; ++k) D[1 + k] = (double)a0t[k];
          for (int k = 0; k < 3 * na; ++k) D[8 + k] = (double)q[k];
          for (int k = 0; k < 3 * na; ++k) D[8 + 96 + k] = (double)Qp[pidx(k, k)];
        }
        __syncwarp();
#endif
        // iterate to the resolution of the DATA: float32 states carry 6e-8 relative rounding, so a float64 solve of a
        // float32 problem stops at 1e-9 of the normalised problem (3-4 interior-point iterations earlier than the 1e-12 of
        // float64 data)
        const S qp_tol = (sizeof(S) == 8 && sizeof(T) == 4) ? S(1e-9) : QpTol<S>::tol();
        const int qp_rc = qp_solve<S>(Qp, Hp, vN, vM, na, S(P.mu), lane, qp_tol);
        const int qp_it = qp_rc & 0xFFFF;
        if (P.status && lane == 0 && (qp_rc & 0x10000)) atomicOr(P.status + env, 8);  // B200SIM_STATUS_QP_NOT_CONVERGED
#ifdef B200SIM_RIGID_DEBUG
        if (P.dbg && lane == 0 && env == 0) {
          double* D = reinterpret_cast<double*>(P.dbg + 1600);
          for (int k = 0; k < 3 * na; ++k) D[8 + 192 + k] = (double)vN[k];
          D[7] = (double)qp_it;
        }
        __syncwarp();
#endif
        if (P.dbg && lane == 0) {
          atomicAdd(P.dbg + 0, (unsigned long long)qp_it);
          atomicAdd(P.dbg + 1, 1ull);
          atomicMax(P.dbg + 2, (unsigned long long)qp_it);
          atomicAdd(P.dbg + 3, (unsigned long long)na);
        }
      } else {
        // RelaxedRigid (relaxed_rigid.py:380-505): the reference minimises |A x + b|^2, A = Delassus + diag(r),
        // with L-BFGS; A is positive definite on the active points, so the minimiser is x = -A^-1 b:
        // one pivoted Cholesky solve instead of the interior-point iteration of the rigid model.
        delassus(na, S(0));
        const int N = 3 * na;
        S* lam = vN;
        for (int a = lane; a < na; a += 32) {
          const T* pw = pts + (size_t)alist[a] * RPT;
#pragma unroll
          for (int c3 = 0; c3 < 3; ++c3) {
            Qp[pidx(3 * a + c3, 3 * a + c3)] += S(pw[RP_F + c3]);
            lam[3 * a + c3] = -q[3 * a + c3];
          }
        }
        __syncwarp();
        psd_solve_pivoted<S>(Qp, Hp, vN + 5 * N, lam, vN + N, perm, N, S(sizeof(S) == 8 ? 1e-14 : 1e-7), lane);
      }
      B200SIM_RIGID_MARK(5);
      for (int a = lane; a < na; a += 32) {
        T* pw = pts + (size_t)alist[a] * RPT;
        pw[RP_F] = T(vN[3 * a]); pw[RP_F + 1] = T(vN[3 * a + 1]); pw[RP_F + 2] = T(vN[3 * a + 2]);
      }
      __syncwarp();
      tree_response();
      for (int i = 1 + lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        ri[O_SDD] += ri[O_TAU];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) a0t[k] += ws[O_C + k];
#ifdef B200SIM_RIGID_DEBUG
      if (P.dbg && lane == 0 && env == 0) {
        double* D = reinterpret_cast<double*>(P.dbg + 1600);
        for (int k = 0; k < 6; ++k) D[8 + 288 + k] = (double)a0t[k];
        for (int i = 1; i < nL && i < 64; ++i) D[8 + 300 + i] = (double)ws[(size_t)i * REC + O_SDD];
      }
#endif
    }
    __syncwarp();

    B200SIM_RIGID_MARK(6);
    // ============================================================== semi-implicit Euler
    T Wa[6];
    cross3(b.p, a0t + 3, Wa);
    Wa[0] += a0t[0]; Wa[1] += a0t[1]; Wa[2] += a0t[2] + P.g;
    Wa[3] = a0t[3]; Wa[4] = a0t[4]; Wa[5] = a0t[5];
    {
      BaseState<T> nb;
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
      quat_to_dcm(nb.qn, nb.R);
      b = nb;
    }
    for (int i = 1 + lane; i < nL; i += 32) {
      T* ri = ws + (size_t)i * REC;
      const T sdn = ri[O_SD] + dt * ri[O_SDD];
      ri[O_SD] = sdn;
      ri[O_S] = ri[O_S] + dt * sdn;
    }
    __syncwarp();

    // ============================================================== phase B: state at t+dt
    // kinematics of the new state with the PRE-impact velocities: these are the caches the
    // reference returns (rigid.py:429-434 leaves them untouched)
    B200SIM_RIGID_MARK(7);
    kinematics(b, true);
    if (lane == 0) {
      stn<4>(P.q_o + env * 4, b.qn);
      stn<3>(P.p_o + env * 3, b.p);
      if (P.W_H_B) store_transform(P.W_H_B + env * 16, b.R, b.p);
      if (P.iXl) {
        T R0[9], p0[3], t[3], X[36];
        mat3_mul(b.R, sm_cst + C_M0, R0);
        mat3_vec(b.R, sm_cst + C_TPRE, t);
        p0[0] = b.p[0] + t[0]; p0[1] = b.p[1] + t[1]; p0[2] = b.p[2] + t[2];
        inverse_adjoint(X, R0, p0);
        stg_vec<36>(P.iXl + env * nL * 36, X);
      }
    }
    for (int i = lane; i < nL; i += 32) {
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
    for (int i = 1 + lane; i < nL; i += 32) P.s_o[env * n + (i - 1)] = ws[(size_t)i * REC + O_S];
    } else {
      kinematics(b, false);
    }

    // impact (rigid.py:385-436): project nu onto {velocity of the active points = 0}
    B200SIM_RIGID_MARK(8);
    const int na2 = relaxed ? 0 : contact_points(true);  // RelaxedRigid: no impact step (relaxed_rigid.py:262-281)
    B200SIM_RIGID_MARK(9);
    if (na2 > cap) {
      // next level applies the impact to the pre-impact result, which must then be complete
      if (!impact_only) {
        for (int i = 1 + lane; i < nL; i += 32) P.sd_o[env * n + (i - 1)] = ws[(size_t)i * REC + O_SD];
        if (lane == 0) {
          stn<3>(P.vlin_o + env * 3, b.vlin);
          stn<3>(P.omega_o + env * 3, b.w);
        }
        __threadfence();
      }
      if (lane == 0) over_push(P, (int)env, 1);
      continue;
    }
    T dv0[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    if (P.dbg && lane == 0 && na2 > 0) { atomicAdd(P.dbg + 6, 1ull); atomicAdd(P.dbg + 7, (unsigned long long)na2); }
    if (na2 > 0) {
      link_init(false);
      pass2();
      delassus(na2, S(0));
      B200SIM_RIGID_MARK(10);
      const int N2 = 3 * na2;
      S* lam = vN;
      S* dg = vN + 5 * N2;
      for (int a = lane; a < na2; a += 32) {
        const T* pw = pts + (size_t)alist[a] * RPT;
        lam[3 * a] = -S(pw[RP_B]); lam[3 * a + 1] = -S(pw[RP_B + 1]); lam[3 * a + 2] = -S(pw[RP_B + 2]);
      }
      __syncwarp();
      psd_solve_pivoted<S>(Qp, Hp, dg, lam, vN + N2, perm, N2, S(RankTol<T>::tol()), lane);
      B200SIM_RIGID_MARK(11);
      for (int a = lane; a < na2; a += 32) {
        T* pw = pts + (size_t)alist[a] * RPT;
        pw[RP_F] = T(lam[3 * a]); pw[RP_F + 1] = T(lam[3 * a + 1]); pw[RP_F + 2] = T(lam[3 * a + 2]);
      }
      __syncwarp();
      tree_response();
      ldn<6>(ws + O_C, dv0);
      for (int i = 1 + lane; i < nL; i += 32) {
        T* ri = ws + (size_t)i * REC;
        ri[O_SD] += ri[O_TAU];
      }
    }
    __syncwarp();
    for (int i = 1 + lane; i < nL; i += 32) P.sd_o[env * n + (i - 1)] = ws[(size_t)i * REC + O_SD];
    if (lane == 0 && na2 == 0) {
      stn<3>(P.vlin_o + env * 3, b.vlin);
      stn<3>(P.omega_o + env * 3, b.w);
    } else if (lane == 0) {
      // base: F_0 coordinates [pdot_B; omega] -> inertial-fixed linear part
      T t[3], w2[3], pdB[3];
      cross3(b.w, b.p, t);
      pdB[0] = b.vlin[0] + t[0] + dv0[0]; pdB[1] = b.vlin[1] + t[1] + dv0[1]; pdB[2] = b.vlin[2] + t[2] + dv0[2];
      w2[0] = b.w[0] + dv0[3]; w2[1] = b.w[1] + dv0[4]; w2[2] = b.w[2] + dv0[5];
      cross3(w2, b.p, t);
      const T vl[3] = {pdB[0] - t[0], pdB[1] - t[1], pdB[2] - t[2]};
      stn<3>(P.vlin_o + env * 3, vl);
      stn<3>(P.omega_o + env * 3, w2);
    }
    B200SIM_RIGID_MARK(12);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// the contact QPs of a split rigid level, one warp per work item
// ------------------------------------------------------------------------------------
// The interior-point iteration is a chain of dependent shared-memory round trips on ONE warp; inside the monolithic
// kernel an environment also holds its link records, so only 5-7 warps fit an SM and the schedulers idle 80 % of the
// time (profiles/r01_rigid_kernel_v1.md).  On its own the solver needs H and a few vectors in shared memory (Q stays
// in the item's record: it is read once per iteration, in order): 11.7 KB for 12 active points -- 19-20 resident warps.
struct QpLayout { size_t Hp, vecN, vecM, total; };
template <typename S>
__host__ __device__ inline QpLayout qp_layout(int cap) {
  QpLayout L;
  const size_t N = 3 * (size_t)cap, M = 5 * (size_t)cap, NP = npk(N);
  size_t o = 0;
  L.Hp = o;   o = rl_align(o + sizeof(S) * NP);
  L.vecN = o; o = rl_align(o + sizeof(S) * 7 * N);
  L.vecM = o; o = rl_align(o + sizeof(S) * 9 * M);
  L.total = o;
  return L;
}

constexpr int QP_WARPS = 2;  // warps per block (no block-level barrier: small blocks even out the iteration counts)

// MINB = resident blocks per SM the register allocation aims at (8: 128 registers, 9: 112, 10: 96)
template <typename S, int MINB>
__global__ void __launch_bounds__(32 * QP_WARPS, MINB) rigid_qp_kernel(const int* work_count, int* next_item, unsigned char* qp_buf,
                                                                  long long qp_stride, int cap, S mu_f, S tol, int* status,
                                                                  unsigned long long* dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const QpLayout L = qp_layout<S>(cap);
  unsigned char* wb = smem_raw + (size_t)wrp * L.total;
  S* Hp = reinterpret_cast<S*>(wb + L.Hp);
  S* vN = reinterpret_cast<S*>(wb + L.vecN);
  S* vM = reinterpret_cast<S*>(wb + L.vecM);
  const int total = *work_count;
  // iteration counts differ from item to item (5 ... 60): the warps draw items from a counter instead of striding
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(next_item, 1);
    item = __shfl_sync(FULL, item, 0);
    if (item >= total) break;
    int* hdr = reinterpret_cast<int*>(qp_buf + (long long)item * qp_stride);
    const int na = hdr[0];
    if (na <= 0) continue;  // whole warp: no contact at t, impact-only, or handed to the next level
    const int N = 3 * na;
    S* qg = reinterpret_cast<S*>(hdr + 4);
    S* xg = qg + 3 * cap;
    const S* Qg = qg + 6 * cap;
    for (int i = lane; i < N; i += 32) vN[N + i] = qg[i];  // qp_pyramids reads q at vN + N
    __syncwarp();
    const int rc = qp_solve<S>(Qg, Hp, vN, vM, na, mu_f, lane, tol);
    for (int i = lane; i < N; i += 32) xg[i] = vN[i];
    if (lane == 0) {
      hdr[1] = rc;
      if (status && (rc & 0x10000)) atomicOr(status + hdr[2], 8);  // B200SIM_STATUS_QP_NOT_CONVERGED
      if (dbg) {
        const int it = rc & 0xFFFF;
        atomicAdd(dbg + 0, (unsigned long long)it);
        atomicAdd(dbg + 1, 1ull);
        atomicMax(dbg + 2, (unsigned long long)it);
        atomicAdd(dbg + 3, (unsigned long long)na);
      }
    }
    __syncwarp();
  }
}

}  // namespace b200sim
