// b200sim_dual.cuh -- forward-mode AD scalar for the step kernel (BASELINE config 5).
//
// step_kernel<T, G> is generic in its scalar type.  Instantiating it with Dual<double>
// propagates ONE tangent direction through the whole step (joint transforms, FK, contacts,
// ABA, integrator, caches): a Jacobian-vector product d(step)/d(theta) . theta_dot for
// theta = (joint positions, link masses, ... any input leaf), exactly what
// `jax.jvp(js.model.step, ...)` returns in the reference
// (tests/test_automatic_differentiation.py:346-420 checks it against finite differences).
//
// Derivative conventions follow JAX so that gradients agree with the reference's AD:
//   max/min: the selected operand's tangent, ties split 1/2-1/2 (jnp.maximum);
//   comparisons / where / sign / clip: on the primal value, the taken branch's tangent;
//   norm at zero: zero tangent (safe_norm custom JVP, math/utils.py:23-40);
//   pow(x, p): p x^(p-1) (the epsilon inside the base keeps it finite, soft.py:246-252).
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace b200sim {

template <typename S>
struct alignas(2 * sizeof(S)) Dual {
  S v, d;
  __host__ __device__ Dual() {}
  __host__ __device__ constexpr Dual(S value) : v(value), d(S(0)) {}
  __host__ __device__ constexpr Dual(int value) : v(S(value)), d(S(0)) {}
  __host__ __device__ constexpr Dual(S value, S tangent) : v(value), d(tangent) {}
};

#define B200_HD __host__ __device__ __forceinline__

template <typename S> B200_HD Dual<S> operator+(Dual<S> a, Dual<S> b) { return Dual<S>(a.v + b.v, a.d + b.d); }
template <typename S> B200_HD Dual<S> operator-(Dual<S> a, Dual<S> b) { return Dual<S>(a.v - b.v, a.d - b.d); }
template <typename S> B200_HD Dual<S> operator-(Dual<S> a) { return Dual<S>(-a.v, -a.d); }
template <typename S> B200_HD Dual<S> operator*(Dual<S> a, Dual<S> b) { return Dual<S>(a.v * b.v, a.v * b.d + a.d * b.v); }
template <typename S> B200_HD Dual<S> operator/(Dual<S> a, Dual<S> b) {
  const S inv = S(1) / b.v;
  const S q = a.v * inv;
  return Dual<S>(q, (a.d - q * b.d) * inv);
}
template <typename S> B200_HD Dual<S>& operator+=(Dual<S>& a, Dual<S> b) { a.v += b.v; a.d += b.d; return a; }
template <typename S> B200_HD Dual<S>& operator-=(Dual<S>& a, Dual<S> b) { a.v -= b.v; a.d -= b.d; return a; }
template <typename S> B200_HD Dual<S>& operator*=(Dual<S>& a, Dual<S> b) { a = a * b; return a; }
template <typename S> B200_HD bool operator<(Dual<S> a, Dual<S> b) { return a.v < b.v; }
template <typename S> B200_HD bool operator>(Dual<S> a, Dual<S> b) { return a.v > b.v; }
template <typename S> B200_HD bool operator<=(Dual<S> a, Dual<S> b) { return a.v <= b.v; }
template <typename S> B200_HD bool operator>=(Dual<S> a, Dual<S> b) { return a.v >= b.v; }
template <typename S> B200_HD bool operator==(Dual<S> a, Dual<S> b) { return a.v == b.v; }
template <typename S> B200_HD bool operator!=(Dual<S> a, Dual<S> b) { return a.v != b.v; }

using DualD = Dual<double>;

__device__ __forceinline__ void sincos_t(DualD x, DualD* s, DualD* c) {
  double sv, cv;
  sincos(x.v, &sv, &cv);
  *s = DualD(sv, cv * x.d);
  *c = DualD(cv, -sv * x.d);
}
__device__ __forceinline__ DualD sqrt_t(DualD x) {
  const double r = sqrt(x.v);
  return DualD(r, x.v == 0.0 ? 0.0 : 0.5 * x.d / r);  // safe_norm: zero tangent at zero
}
__device__ __forceinline__ DualD pow_t(DualD x, DualD y) {
  const double r = pow(x.v, y.v);
  return DualD(r, y.v * pow(x.v, y.v - 1.0) * x.d + (y.d != 0.0 ? log(x.v) * r * y.d : 0.0));
}
__device__ __forceinline__ DualD abs_t(DualD x) { return x.v < 0.0 ? -x : (x.v > 0.0 ? x : DualD(0.0, 0.0)); }
__device__ __forceinline__ DualD max_t(DualD a, DualD b) {
  if (a.v > b.v) return a;
  if (a.v < b.v) return b;
  return DualD(a.v, 0.5 * (a.d + b.d));
}
__device__ __forceinline__ DualD min_t(DualD a, DualD b) {
  if (a.v < b.v) return a;
  if (a.v > b.v) return b;
  return DualD(a.v, 0.5 * (a.d + b.d));
}
__device__ __forceinline__ DualD keep_tangent(DualD x, bool keep) { return DualD(x.v, keep ? x.d : 0.0); }
__device__ __forceinline__ DualD rcp_t(DualD x) {
  const double r = 1.0 / x.v;
  return DualD(r, -r * r * x.d);
}

}  // namespace b200sim
