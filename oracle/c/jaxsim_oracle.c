/*
 * jaxsim_oracle.c -- plain-C restatement of the reference's `jaxsim.api.model.step` path.
 *
 * TEST INFRASTRUCTURE ONLY (same role and rules as oracle/jaxsim_oracle.py; nothing under
 * jaxsim_b200/ links or loads it).  It exists for two reasons: (1) parity at BASELINE's FULL
 * batch sizes in seconds instead of minutes, (2) a multi-threaded CPU baseline that is not
 * handicapped by NumPy dispatch overhead.  It is itself pinned against the NumPy oracle in
 * tests/test_c_oracle.py (both restate the same reference lines; citations below are
 * relative to /root/reference/src/jaxsim).
 *
 * Like the NumPy oracle it keeps the reference's formulation on purpose: dense 6x6
 * adjoints, body-fixed link frames, [linear; angular] ordering, one link after the other --
 * NOT the world-aligned/structured formulation of the CUDA kernels.
 *
 * Build: make -C oracle   (gcc -O2 -pthread -shared -fPIC)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200sim.h" /* only for struct B200SimModelDesc (the model image) */

#define MAXL 128

typedef double M6[6][6];
typedef double M3[3][3];

static void wedge(const double v[3], M3 S) {
  S[0][0] = 0; S[0][1] = -v[2]; S[0][2] = v[1];
  S[1][0] = v[2]; S[1][1] = 0; S[1][2] = -v[0];
  S[2][0] = -v[1]; S[2][1] = v[0]; S[2][2] = 0;
}
static void m3mul(const M3 A, const M3 B, M3 C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[i][k] * B[k][j]; C[i][j] = s; }
}
static void m3T(const M3 A, M3 B) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[i][j] = A[j][i]; }
static void m6mul(const M6 A, const M6 B, M6 C) {
  M6 T;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < 6; ++k) s += A[i][k] * B[k][j]; T[i][j] = s; }
  memcpy(C, T, sizeof(M6));
}
static void m6T(const M6 A, M6 B) { M6 T; for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) T[i][j] = A[j][i]; memcpy(B, T, sizeof(M6)); }
static void m6vec(const M6 A, const double x[6], double y[6]) {
  double t[6];
  for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < 6; ++k) s += A[i][k] * x[k]; t[i] = s; }
  memcpy(y, t, sizeof(t));
}

/* jaxlie SO3(wxyz).as_matrix() (rbda/aba.py:79-86) */
static void quat_to_dcm(const double q[4], M3 R) {
  const double nsq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], k = 2.0 / nsq;
  double q2[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) q2[i][j] = q[i] * q[j] * k;
  R[0][0] = 1 - q2[2][2] - q2[3][3]; R[0][1] = q2[1][2] - q2[3][0]; R[0][2] = q2[1][3] + q2[2][0];
  R[1][0] = q2[1][2] + q2[3][0]; R[1][1] = 1 - q2[1][1] - q2[3][3]; R[1][2] = q2[2][3] - q2[1][0];
  R[2][0] = q2[1][3] - q2[2][0]; R[2][1] = q2[2][3] + q2[1][0]; R[2][2] = 1 - q2[1][1] - q2[2][2];
}
/* Adjoint.from_rotation_and_translation (math/adjoint.py:66-107) */
static void adjoint(const M3 R, const double p[3], int inverse, M6 X) {
  M3 S, T, Rt;
  memset(X, 0, sizeof(M6));
  wedge(p, S);
  if (!inverse) {
    m3mul(S, R, T);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { X[i][j] = R[i][j]; X[i][3 + j] = T[i][j]; X[3 + i][3 + j] = R[i][j]; }
  } else {
    m3T(R, Rt);
    m3mul(Rt, S, T);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { X[i][j] = Rt[i][j]; X[i][3 + j] = -T[i][j]; X[3 + i][3 + j] = Rt[i][j]; }
  }
}
/* Adjoint.inverse (math/adjoint.py:135-160) */
static void adjoint_inverse(const M6 X, M6 Y) {
  M3 Rt, T, A, B;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Rt[i][j] = X[j][i]; T[i][j] = X[i][3 + j]; }
  m3mul(Rt, T, A);
  m3mul(A, Rt, B);
  M6 Z;
  memset(Z, 0, sizeof(M6));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Z[i][j] = Rt[i][j]; Z[i][3 + j] = -B[i][j]; Z[3 + i][3 + j] = Rt[i][j]; }
  memcpy(Y, Z, sizeof(M6));
}
/* Adjoint.to_transform (math/adjoint.py:109-133): H = [R, vee(X_12 R^T); 0 1] */
static void adjoint_to_transform(const M6 X, double H[16]) {
  M3 R, O, Rt, M;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { R[i][j] = X[i][j]; O[i][j] = X[i][3 + j]; }
  m3T(R, Rt);
  m3mul(O, Rt, M);
  const double p[3] = {0.5 * (M[2][1] - M[1][2]), 0.5 * (M[0][2] - M[2][0]), 0.5 * (M[1][0] - M[0][1])};
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) H[4 * i + j] = R[i][j]; H[4 * i + 3] = p[i]; }
  H[12] = 0; H[13] = 0; H[14] = 0; H[15] = 1;
}
/* Cross.vx / vx_star (math/cross.py:13-58) */
static void cross_vx(const double v[6], M6 X) {
  M3 Sw, Sv;
  memset(X, 0, sizeof(M6));
  wedge(v + 3, Sw);
  wedge(v, Sv);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { X[i][j] = Sw[i][j]; X[i][3 + j] = Sv[i][j]; X[3 + i][3 + j] = Sw[i][j]; }
}
static void cross_vx_star(const double v[6], M6 X) { M6 T; cross_vx(v, T); for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) X[i][j] = -T[j][i]; }

/* Rotation.from_axis_angle (math/rotation.py:58-84) */
static void rot_axis_angle(const double vec[3], M3 R) {
  const double th = sqrt(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]);
  const double s = sin(th), c = cos(th), c1 = 2 * sin(th / 2) * sin(th / 2), st = (th == 0) ? 1.0 : th;
  const double u[3] = {vec[0] / st, vec[1] / st, vec[2] / st};
  M3 S, Q;
  wedge(u, S);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Q[i][j] = c * (i == j) - s * S[i][j] + c1 * u[i] * u[j];
  m3T(Q, R);
}
/* 4x4 homogeneous helpers */
static void h_mul(const double A[16], const double B[16], double C[16]) {
  double T[16];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j]; T[4 * i + j] = s; }
  memcpy(C, T, sizeof(T));
}
static void h_to_adjoint(const double H[16], int inverse, M6 X) {
  M3 R; double p[3];
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i][j] = H[4 * i + j]; p[i] = H[4 * i + 3]; }
  adjoint(R, p, inverse, X);
}

/* KinDynParameters.joint_transforms (api/kin_dyn_parameters.py:396-451) */
static void joint_transforms(const B200SimModelDesc *d, const double *s, const double W_H_B[16], M6 *iXl) {
  const int nL = d->n_links;
  for (int i = 0; i < nL; ++i) {
    double J[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, T[16], H[16];
    const double *pre = d->lam_H_pre + 16 * (size_t)i;
    double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (i == 0) { memcpy(J, W_H_B, sizeof(J)); pre = eye; }
    else {
      const double *ax = d->joint_axis + 3 * (size_t)i;
      const double si = s[i - 1];
      if (d->joint_type[i] == 1) {
        const double v[3] = {si * ax[0], si * ax[1], si * ax[2]};
        M3 R; rot_axis_angle(v, R);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) J[4 * r + c] = R[r][c];
      } else if (d->joint_type[i] == 2) { J[3] = si * ax[0]; J[7] = si * ax[1]; J[11] = si * ax[2]; }
    }
    h_mul(pre, J, T);
    h_mul(T, d->suc_H_i + 16 * (size_t)i, H);
    h_to_adjoint(H, 1, iXl[i]);
  }
}

static void motion_subspace(const B200SimModelDesc *d, int i, double S[6]) {
  memset(S, 0, 6 * sizeof(double));
  const double *ax = d->joint_axis + 3 * (size_t)i;
  if (d->joint_type[i] == 1) { S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2]; }
  else if (d->joint_type[i] == 2) { S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2]; }
}

/* Inertia.to_sixd (math/inertia.py:14-41) */
static void link_inertia(const B200SimModelDesc *d, int i, M6 M) {
  const double m = d->link_mass[i], *c = d->link_com + 3 * (size_t)i, *I6 = d->link_inertia + 6 * (size_t)i;
  M3 C, Ct, CCt;
  const double I[3][3] = {{I6[0], I6[1], I6[2]}, {I6[1], I6[3], I6[4]}, {I6[2], I6[4], I6[5]}};
  wedge(c, C); m3T(C, Ct); m3mul(C, Ct, CCt);
  memset(M, 0, sizeof(M6));
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) {
    M[r][k] = m * (r == k); M[r][3 + k] = m * Ct[r][k]; M[3 + r][k] = m * C[r][k]; M[3 + r][3 + k] = I[r][k] + m * CCt[r][k];
  }
}

/* rbda.forward_kinematics_model (rbda/forward_kinematics.py:12-113) + data.replace caches */
static void forward_kinematics(const B200SimModelDesc *d, const double *p, const double *qn, const double *s,
                               const double *vl, const double *w, const double *sd, double *W_H_B, double *iXl_out,
                               double *W_H_L, double *W_v) {
  const int nL = d->n_links;
  M3 R; quat_to_dcm(qn, R);
  double H[16];
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) H[4 * i + j] = R[i][j]; H[4 * i + 3] = p[i]; }
  H[12] = 0; H[13] = 0; H[14] = 0; H[15] = 1;
  if (W_H_B) memcpy(W_H_B, H, sizeof(H));
  static _Thread_local M6 iXl[MAXL], WX[MAXL];
  joint_transforms(d, s, H, iXl);
  if (iXl_out) memcpy(iXl_out, iXl, (size_t)nL * sizeof(M6));
  adjoint_inverse(iXl[0], WX[0]);
  for (int k = 0; k < 3; ++k) { W_v[k] = vl[k]; W_v[3 + k] = w[k]; }
  for (int i = 1; i < nL; ++i) {
    M6 lXi; double S[6], vJ[6], t[6];
    const int par = d->parent[i];
    adjoint_inverse(iXl[i], lXi);
    m6mul(WX[par], lXi, WX[i]);
    motion_subspace(d, i, S);
    for (int k = 0; k < 6; ++k) vJ[k] = S[k] * sd[i - 1];
    m6vec(WX[i], vJ, t);
    for (int k = 0; k < 6; ++k) W_v[6 * i + k] = W_v[6 * par + k] + t[k];
  }
  for (int i = 0; i < nL; ++i) adjoint_to_transform(WX[i], W_H_L + 16 * (size_t)i);
}

/* compute_resultant_torques + tn_curve_fn (api/actuation_model.py:7-126) */
static void resultant_torques(const B200SimModelDesc *d, const double *s, const double *sd, const double *tref, double *tau) {
  for (int j = 0; j < d->n_dofs; ++j) {
    const double lower = fmin(s[j] - d->position_limits_min[j], 0.0), upper = fmax(s[j] - d->position_limits_max[j], 0.0);
    double tl = -d->position_limit_spring[j] * (lower + upper);
    tl = tl - tl * d->position_limit_damper[j] * sd[j];
    double tf = 0;
    if (d->enable_friction) {
      const double sg = (sd[j] > 0) - (sd[j] < 0);
      tf = -(d->friction_static[j] * sg + d->friction_viscous[j] * sd[j]);
    }
    const double tt = (tref ? tref[j] : 0.0) + tf + tl, av = fabs(sd[j]);
    double lim;
    if (av <= d->omega_th) lim = d->torque_max;
    else if (av <= d->omega_max) lim = d->torque_max * (1 - (av - d->omega_th) / (d->omega_max - d->omega_th));
    else lim = 0;
    tau[j] = fmin(fmax(tt, -lim), lim);
  }
}

/* hunt_crossley_contact_model (rbda/contacts/soft.py:195-339), FlatTerrain */
static void hunt_crossley(const B200SimModelDesc *d, const double pc[3], const double pd[3], const double m[3], double f[3], double md[3]) {
  const double eps = 2.220446049250313e-16, K = d->soft_K, D = d->soft_D, mu = d->soft_mu;
  const double delta = fmax(0.0, d->terrain_height - pc[2]);
  const double ddot = (delta > 0) ? -pd[2] : 0.0;
  const double dp = pow(delta + eps, d->soft_p), dq = pow(delta + eps, d->soft_q);
  const double fn = fmax(0.0, (K * dp) * delta + (D * dq) * ddot);
  const double vt[3] = {pd[0], pd[1], 0}, mn[3] = {0, 0, m[2]}, mt[3] = {m[0], m[1], 0};
  double ft[3];
  for (int k = 0; k < 3; ++k) ft[k] = -((K * dp) * mt[k] + (D * dq) * vt[k]);
  const double ft2 = ft[0] * ft[0] + ft[1] * ft[1] + ft[2] * ft[2];
  const int nocontact = delta <= 0, sticking = nocontact || (ft2 <= (mu * fn) * (mu * fn));
  const double nrm = sqrt(ft2), den = nrm + eps * (nrm == 0);
  if (!sticking) { const double sc = fmin(mu * fn, nrm); for (int k = 0; k < 3; ++k) ft[k] = sc * (ft[k] / den); }
  if (nocontact) ft[0] = ft[1] = ft[2] = 0;
  const int status = sticking + nocontact;
  for (int k = 0; k < 3; ++k) {
    const double no = -(K / D) * m[k], st = vt[k] - (K / D) * mn[k], sl = -(ft[k] + (K * dp) * mt[k]) / (D * dq);
    md[k] = status == 0 ? sl : (status == 1 ? st : no);
  }
  f[0] = ft[0]; f[1] = ft[1]; f[2] = fn + ft[2];
}

/* 6x6 solve with partial pivoting (jnp.linalg.solve, rbda/aba.py:241) */
static void solve6(M6 A, double b[6]) {
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
    if (piv != c) { for (int k = 0; k < 6; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; } double t = b[c]; b[c] = b[piv]; b[piv] = t; }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < 6; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 5; r >= 0; --r) { double s = b[r]; for (int k = r + 1; k < 6; ++k) s -= A[r][k] * b[k]; b[r] = s / A[r][r]; }
}

/* rbda.aba (rbda/aba.py:12-292) */
static void aba(const B200SimModelDesc *d, const double *p, const double *qn, const double *s, const double *vl,
                const double *w, const double *sd, const double *tau, const double *W_f, double *W_a, double *sdd) {
  const int nL = d->n_links;
  static _Thread_local M6 iXl[MAXL], MA[MAXL], iX0[MAXL], Mi[MAXL];
  static _Thread_local double v[MAXL][6], c[MAXL][6], pA[MAXL][6], U[MAXL][6], dd[MAXL], uu[MAXL], a[MAXL][6];
  M3 R; quat_to_dcm(qn, R);
  M6 W_X_B, B_X_W;
  adjoint(R, p, 0, W_X_B);
  adjoint(R, p, 1, B_X_W);
  double H[16];
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) H[4 * i + j] = R[i][j]; H[4 * i + 3] = p[i]; }
  H[12] = 0; H[13] = 0; H[14] = 0; H[15] = 1;
  joint_transforms(d, s, H, iXl);
  const double W_g[6] = {0, 0, d->gravity, 0, 0, 0}, W_v[6] = {vl[0], vl[1], vl[2], w[0], w[1], w[2]};
  memset(v, 0, sizeof(double) * 6 * nL); memset(c, 0, sizeof(double) * 6 * nL); memset(pA, 0, sizeof(double) * 6 * nL);
  memset(MA, 0, sizeof(M6) * nL); memset(iX0, 0, sizeof(M6) * nL);
  for (int k = 0; k < 6; ++k) iX0[0][k][k] = 1;
  for (int i = 0; i < nL; ++i) link_inertia(d, i, Mi[i]);
  if (d->floating_base) {
    M6 X, T; double t[6], t2[6];
    m6vec(B_X_W, W_v, v[0]);
    memcpy(MA[0], Mi[0], sizeof(M6));
    cross_vx_star(v[0], X); m6mul(X, MA[0], T); m6vec(T, v[0], t);
    m6T(W_X_B, X); m6vec(X, W_f, t2);
    for (int k = 0; k < 6; ++k) pA[0][k] = t[k] - t2[k];
  }
  for (int i = 1; i < nL; ++i) { /* pass 1 (:131-171) */
    const int par = d->parent[i];
    double S[6], vJ[6], t[6], t2[6];
    M6 X, T, Y;
    motion_subspace(d, i, S);
    for (int k = 0; k < 6; ++k) vJ[k] = S[k] * sd[i - 1];
    m6vec(iXl[i], v[par], t);
    for (int k = 0; k < 6; ++k) v[i][k] = t[k] + vJ[k];
    cross_vx(v[i], X); m6vec(X, vJ, c[i]);
    memcpy(MA[i], Mi[i], sizeof(M6));
    m6mul(iXl[i], iX0[par], iX0[i]);
    m6mul(iX0[i], B_X_W, T); adjoint_inverse(T, Y); m6T(Y, Y); /* i_Xf_W */
    cross_vx_star(v[i], X); m6mul(X, Mi[i], T); m6vec(T, v[i], t);
    m6vec(Y, W_f + 6 * (size_t)i, t2);
    for (int k = 0; k < 6; ++k) pA[i][k] = t[k] - t2[k];
  }
  for (int i = nL - 1; i >= 1; --i) { /* pass 2 (:184-234) */
    const int par = d->parent[i];
    double S[6], t[6];
    motion_subspace(d, i, S);
    m6vec(MA[i], S, U[i]);
    dd[i] = 0; uu[i] = tau[i - 1];
    for (int k = 0; k < 6; ++k) { dd[i] += S[k] * U[i][k]; uu[i] -= S[k] * pA[i][k]; }
    if (par != 0 || d->floating_base) {
      M6 Ma, Xt, T;
      double pa[6];
      for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) Ma[r][k] = MA[i][r][k] - (U[i][r] / dd[i]) * U[i][k];
      m6vec(Ma, c[i], t);
      for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + t[k] + U[i][k] * (uu[i] / dd[i]);
      m6T(iXl[i], Xt); m6mul(Xt, Ma, T); m6mul(T, iXl[i], T);
      for (int r = 0; r < 6; ++r) for (int k = 0; k < 6; ++k) MA[par][r][k] += T[r][k];
      m6vec(Xt, pa, t);
      for (int k = 0; k < 6; ++k) pA[par][k] += t[k];
    }
  }
  double a0[6];
  if (d->floating_base) {
    M6 N;
    for (int r = 0; r < 6; ++r) { for (int k = 0; k < 6; ++k) N[r][k] = -MA[0][r][k]; a0[r] = pA[0][r]; }
    solve6(N, a0);
  } else {
    m6vec(B_X_W, W_g, a0);
    for (int k = 0; k < 6; ++k) a0[k] = -a0[k];
  }
  memcpy(a[0], a0, sizeof(a0));
  for (int i = 1; i < nL; ++i) { /* pass 3 (:251-277) */
    const int par = d->parent[i];
    double S[6], ai[6];
    motion_subspace(d, i, S);
    m6vec(iXl[i], a[par], ai);
    double ua = 0;
    for (int k = 0; k < 6; ++k) { ai[k] += c[i][k]; }
    for (int k = 0; k < 6; ++k) ua += U[i][k] * ai[k];
    sdd[i - 1] = (uu[i] - ua) / dd[i];
    for (int k = 0; k < 6; ++k) a[i][k] = ai[k] + S[k] * sdd[i - 1];
  }
  if (d->floating_base) { m6vec(W_X_B, a[0], W_a); for (int k = 0; k < 6; ++k) W_a[k] += W_g[k]; }
  else memset(W_a, 0, 6 * sizeof(double));
}

/* One js.model.step (api/model.py:2601-2681) of ONE environment: SoftContacts / no contacts,
 * SemiImplicitEuler.  Arrays are the per-environment rows of the batched arrays. */
static void step_one(const B200SimModelDesc *d, const double *s, const double *sd, const double *q, const double *vl,
                     const double *w, const double *p, const double *m, const double *tref, const double *fext,
                     double *s_o, double *sd_o, double *q_o, double *vl_o, double *w_o, double *p_o, double *m_o,
                     double *W_H_B, double *iXl, double *W_H_L, double *W_v) {
  const int nL = d->n_links, n = d->n_dofs, nc = d->n_points;
  const double dt = d->time_step;
  static _Thread_local double tau[MAXL], W_f[MAXL * 6], Hl[MAXL * 16], Wv[MAXL * 6], sdd[MAXL];
  /* caches of the input state (data.replace semantics: normalise with where(norm==0,1,norm)) */
  double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double qr[4], qn[4];
  for (int k = 0; k < 4; ++k) qr[k] = q[k] / (nq == 0 ? 1.0 : nq);
  forward_kinematics(d, p, qr, s, vl, w, sd, NULL, NULL, Hl, Wv);
  resultant_torques(d, s, sd, tref, tau);
  for (int k = 0; k < 6 * nL; ++k) W_f[k] = fext ? fext[k] : 0.0;
  if (d->contact_model == B200SIM_CONTACT_SOFT) {
    for (int k = 0; k < nc; ++k) {
      const int b = d->point_body[k];
      const double *H = Hl + 16 * (size_t)b, *Lp = d->point_position + 3 * (size_t)k, *v = Wv + 6 * (size_t)b;
      double pc[3], pd[3], f[3] = {0, 0, 0}, md[3] = {0, 0, 0};
      for (int r = 0; r < 3; ++r) pc[r] = H[4 * r] * Lp[0] + H[4 * r + 1] * Lp[1] + H[4 * r + 2] * Lp[2] + H[4 * r + 3];
      /* [I, -S(p)] v  (rbda/collidable_points.py:49-53) */
      pd[0] = v[0] - (pc[1] * v[5] - pc[2] * v[4]);
      pd[1] = v[1] - (pc[2] * v[3] - pc[0] * v[5]);
      pd[2] = v[2] - (pc[0] * v[4] - pc[1] * v[3]);
      const double *mk = m ? m + 3 * (size_t)k : NULL;
      const double m0[3] = {mk ? mk[0] : 0, mk ? mk[1] : 0, mk ? mk[2] : 0};
      if (d->point_enabled[k]) hunt_crossley(d, pc, pd, m0, f, md);
      /* W_f = [f; p x f] summed on the parent link (soft.py:378-386, api/contact.py:557-603) */
      double *Wf = W_f + 6 * (size_t)b;
      Wf[0] += f[0]; Wf[1] += f[1]; Wf[2] += f[2];
      Wf[3] += pc[1] * f[2] - pc[2] * f[1]; Wf[4] += pc[2] * f[0] - pc[0] * f[2]; Wf[5] += pc[0] * f[1] - pc[1] * f[0];
      if (m_o) for (int r = 0; r < 3; ++r) m_o[3 * (size_t)k + r] = m0[r] + dt * md[r];
    }
  } else if (m_o) {
    for (int k = 0; k < 3 * nc; ++k) m_o[k] = m ? m[k] : 0.0;
  }
  /* base_orientation (api/data.py:267-286) */
  for (int k = 0; k < 4; ++k) qn[k] = q[k] / (nq + 2.220446049250313e-16 * (nq == 0));
  double W_a[6];
  aba(d, p, qn, s, vl, w, sd, tau, W_f, W_a, sdd);
  /* semi_implicit_euler_integration (api/integrators.py:14-88) */
  double vn[3], wn[3], pdot[3];
  for (int k = 0; k < 3; ++k) { vn[k] = vl[k] + dt * W_a[k]; wn[k] = w[k] + dt * W_a[3 + k]; }
  pdot[0] = vn[0] + (wn[1] * p[2] - wn[2] * p[1]);
  pdot[1] = vn[1] + (wn[2] * p[0] - wn[0] * p[2]);
  pdot[2] = vn[2] + (wn[0] * p[1] - wn[1] * p[0]);
  const double nw = sqrt(wn[0] * wn[0] + wn[1] * wn[1] + wn[2] * wn[2]);
  const double nqn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
  const double v0 = 0.1 * nw * (1 - nqn), qw = qn[0], qx = qn[1], qy = qn[2], qz = qn[3];
  const double qd[4] = {0.5 * (qw * v0 - qx * wn[0] - qy * wn[1] - qz * wn[2]), 0.5 * (qx * v0 + qw * wn[0] + qz * wn[1] - qy * wn[2]),
                        0.5 * (qy * v0 - qz * wn[0] + qw * wn[1] + qx * wn[2]), 0.5 * (qz * v0 + qy * wn[0] - qx * wn[1] + qw * wn[2])};
  double q2[4];
  for (int k = 0; k < 4; ++k) q2[k] = qn[k] + dt * qd[k];
  for (int rep = 0; rep < 2; ++rep) { /* integrators.py:61-63 then data.replace (api/data.py:441-447) */
    const double nn = sqrt(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2] + q2[3] * q2[3]);
    for (int k = 0; k < 4; ++k) q2[k] /= (nn == 0 ? 1.0 : nn);
  }
  for (int k = 0; k < 3; ++k) { p_o[k] = p[k] + dt * pdot[k]; vl_o[k] = vn[k]; w_o[k] = wn[k]; }
  for (int k = 0; k < 4; ++k) q_o[k] = q2[k];
  for (int j = 0; j < n; ++j) { sd_o[j] = sd[j] + dt * sdd[j]; s_o[j] = s[j] + dt * sd_o[j]; }
  if (W_H_L) forward_kinematics(d, p_o, q_o, s_o, vl_o, w_o, sd_o, W_H_B, iXl, W_H_L, W_v);
}

/* Batched entry point: the environments are split over `nthreads` POSIX threads (libgomp is not
 * available in this image).  Cache outputs may be NULL (all or none). */
typedef struct {
  const B200SimModelDesc *d;
  int64_t e0, e1;
  const double *s, *sd, *q, *vl, *w, *p, *m, *tref, *fext;
  double *s_o, *sd_o, *q_o, *vl_o, *w_o, *p_o, *m_o, *W_H_B, *iXl, *W_H_L, *W_v;
} Job;

static void *worker(void *arg) {
  const Job *j = (const Job *)arg;
  const B200SimModelDesc *d = j->d;
  const int nL = d->n_links, n = d->n_dofs, nc = d->n_points;
  const int caches = j->W_H_L != NULL;
  for (int64_t e = j->e0; e < j->e1; ++e) {
    step_one(d, j->s + e * n, j->sd + e * n, j->q + e * 4, j->vl + e * 3, j->w + e * 3, j->p + e * 3,
             j->m ? j->m + e * nc * 3 : NULL, j->tref ? j->tref + e * n : NULL, j->fext ? j->fext + e * nL * 6 : NULL,
             j->s_o + e * n, j->sd_o + e * n, j->q_o + e * 4, j->vl_o + e * 3, j->w_o + e * 3, j->p_o + e * 3,
             j->m_o ? j->m_o + e * nc * 3 : NULL, caches ? j->W_H_B + e * 16 : NULL, caches ? j->iXl + e * nL * 36 : NULL,
             caches ? j->W_H_L + e * nL * 16 : NULL, caches ? j->W_v + e * nL * 6 : NULL);
  }
  return NULL;
}

int oracle_step(const B200SimModelDesc *d, int64_t B, int nthreads, const double *s, const double *sd, const double *q,
                const double *vl, const double *w, const double *p, const double *m, const double *tref, const double *fext,
                double *s_o, double *sd_o, double *q_o, double *vl_o, double *w_o, double *p_o, double *m_o,
                double *W_H_B, double *iXl, double *W_H_L, double *W_v) {
  if (!d || d->n_links > MAXL || d->n_links < 1) return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > B) nthreads = B > 0 ? (int)B : 1;
  Job jobs[256];
  pthread_t th[256];
  const int64_t chunk = (B + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    Job j = {d, t * chunk, (t + 1) * chunk < B ? (t + 1) * chunk : B, s, sd, q, vl, w, p, m, tref, fext,
             s_o, sd_o, q_o, vl_o, w_o, p_o, m_o, W_H_B, iXl, W_H_L, W_v};
    jobs[t] = j;
  }
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
  worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}
