// common.cuh -- scalar/dual arithmetic and quaternion helpers shared by every kernel of libmpxb.
//
// Everything here is templated on the scalar type T in {float, Dual}.  Dual carries one forward-mode tangent,
// so a warp whose lanes seed different input directions produces one Jacobian column per lane -- this is how
// the linearisation kernels reproduce the reference's jax.jacobian / jax.grad calls (trajax linearize,
// optimizers.py:165-173) including the derivative conventions of where / clip / MJX's gradient-safe norm.
// The helpers are __host__ __device__ so that tests/cpu_harness can compile the same arithmetic with g++ and
// compare it with the oracle on a machine without a GPU; the product path only ever runs the __global__ kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace mpxb {

struct Dual {
  float v, d;
  HD Dual() : v(0.f), d(0.f) {}
  HD Dual(float v_) : v(v_), d(0.f) {}
  HD Dual(float v_, float d_) : v(v_), d(d_) {}
};

HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
HD Dual operator/(Dual a, Dual b) {
  float inv = 1.f / b.v;
  float q = a.v * inv;
  return Dual(q, (a.d - q * b.d) * inv);
}
HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
HD Dual operator+(Dual a, float b) { return Dual(a.v + b, a.d); }
HD Dual operator+(float a, Dual b) { return Dual(a + b.v, b.d); }
HD Dual operator-(Dual a, float b) { return Dual(a.v - b, a.d); }
HD Dual operator-(float a, Dual b) { return Dual(a - b.v, -b.d); }
HD Dual operator*(Dual a, float b) { return Dual(a.v * b, a.d * b); }
HD Dual operator*(float a, Dual b) { return Dual(a * b.v, a * b.d); }
HD Dual operator/(Dual a, float b) { float inv = 1.f / b; return Dual(a.v * inv, a.d * inv); }
HD Dual operator/(float a, Dual b) { return Dual(a) / b; }
HD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
HD Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
HD Dual& operator+=(Dual& a, float b) { a.v += b; return a; }

HD float val(float x) { return x; }
HD float val(Dual x) { return x.v; }
HD float tan_of(float) { return 0.f; }
HD float tan_of(Dual x) { return x.d; }

HD float xsqrt(float x) { return sqrtf(x); }
HD Dual xsqrt(Dual x) { float s = sqrtf(x.v); return Dual(s, 0.5f * x.d / s); }
HD float xsin(float x) { return sinf(x); }
HD Dual xsin(Dual x) { return Dual(sinf(x.v), cosf(x.v) * x.d); }
HD float xcos(float x) { return cosf(x); }
HD Dual xcos(Dual x) { return Dual(cosf(x.v), -sinf(x.v) * x.d); }
HD float xlog(float x) { return logf(x); }
HD Dual xlog(Dual x) { return Dual(logf(x.v), x.d / x.v); }
HD float xatan2(float y, float x) { return atan2f(y, x); }
HD Dual xatan2(Dual y, Dual x) {
  float den = x.v * x.v + y.v * y.v;
  return Dual(atan2f(y.v, x.v), (x.v * y.d - y.v * x.d) / den);
}
// jnp.clip: value clamped, tangent passes only strictly inside (lo, hi)
HD float xclip(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
HD Dual xclip(Dual x, float lo, float hi) {
  if (x.v < lo) return Dual(lo, 0.f);
  if (x.v > hi) return Dual(hi, 0.f);
  return x;
}

template <class T> HD void cross3(const T* a, const T* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}
template <class T, class S> HD void cross3m(const T* a, const S* b, T* c) {
  T c0 = a[1] * b[2] - a[2] * b[1];
  T c1 = a[2] * b[0] - a[0] * b[2];
  T c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}

// mjx math.norm: zero value and zero gradient when allclose(x, 0) (atol 1e-8)
template <class T, int K> HD T safe_norm(const T* x) {
  bool is_zero = true;
  for (int i = 0; i < K; ++i) is_zero = is_zero && (fabsf(val(x[i])) <= 1e-8f);
  if (is_zero) return T(0.f);
  T s = x[0] * x[0];
  for (int i = 1; i < K; ++i) s = s + x[i] * x[i];
  return xsqrt(s);
}
// mjx math.normalize_with_norm: x / (n + 1e-6 * (n == 0))
template <class T, int K> HD T normalize_with_norm(const T* x, T* out) {
  T n = safe_norm<T, K>(x);
  T den = (val(n) == 0.f) ? n + 1e-6f : n;
  for (int i = 0; i < K; ++i) out[i] = x[i] / den;
  return n;
}

template <class A, class B, class C> HD void quat_mul(const A* u, const B* v, C* o) {
  C o0 = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  C o1 = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  C o2 = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  C o3 = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
}

// mjx math.quat_to_mat (no normalisation: scales with |q|^2 like the reference)
template <class T> HD void quat_to_mat(const T* q, T* R) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}

// mjx math.rotate(vec, quat)
template <class T, class S> HD void quat_rotate(const T* q, const S* v, T* o) {
  T s = q[0];
  T uv = q[1] * v[0] + q[2] * v[1] + q[3] * v[2];
  T uu = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  T c[3];
  c[0] = q[2] * v[2] - q[3] * v[1];
  c[1] = q[3] * v[0] - q[1] * v[2];
  c[2] = q[1] * v[1] - q[2] * v[0];
  T k = s * s - uu;
  for (int i = 0; i < 3; ++i) o[i] = 2.f * (uv * q[1 + i]) + k * v[i] + 2.f * s * c[i];
}

// mjx math.quat_sub(u, v): rotation vector of v^-1 * u, angle wrapped to (-pi, pi]
template <class T> HD void quat_sub(const T* u, const float* v, T* out) {
  float vi[4] = {v[0], -v[1], -v[2], -v[3]};
  T q[4];
  quat_mul(vi, u, q);
  T axis[3];
  T s = normalize_with_norm<T, 3>(q + 1, axis);
  T angle = 2.f * xatan2(s, q[0]);
  if (val(angle) > 3.14159265358979323846f) angle = angle - 6.28318530717958647692f;
  for (int i = 0; i < 3; ++i) out[i] = axis[i] * angle;
}

// mjx math.quat_integrate(q, v, dt)
template <class T> HD void quat_integrate(const T* q, const T* v, float dt, T* out) {
  T axis[3];
  T n = normalize_with_norm<T, 3>(v, axis);
  T angle = dt * n;
  T s = xsin(angle * 0.5f), c = xcos(angle * 0.5f);
  T qr[4] = {c, axis[0] * s, axis[1] * s, axis[2] * s};
  T qq[4];
  quat_mul(q, qr, qq);
  normalize_with_norm<T, 4>(qq, out);
}

// reference objectives.py:7-13 `penalty` (relaxed log barrier), forward-mode exact
template <class T> HD T penalty(T c, float alpha, float sigma) {
  T ct = xclip(c, 1e-10f, 1e6f);
  T logb = -alpha * xlog(ct);
  T r;
  if (val(c) > sigma) {
    r = logb;
  } else {
    T z = (c - 2.f * sigma) / sigma;
    r = (alpha * 0.5f) * (z * z - 1.f) + logb;
  }
  return xclip(r, 0.f, 1e8f);
}
// jax.grad(jax.grad(penalty)) at a scalar: second derivative, zero wherever the outer clip is active
HD float penalty_dd(float c, float alpha, float sigma) {
  float p = val(penalty<float>(c, alpha, sigma));
  if (!(p > 0.f && p < 1e8f)) return 0.f;
  float inside = (c > 1e-10f && c < 1e6f) ? 1.f : 0.f;
  float ct = fminf(fmaxf(c, 1e-10f), 1e6f);
  float dd = alpha * inside / (ct * ct);
  if (!(c > sigma)) dd += alpha / (sigma * sigma);
  return dd;
}

// 1 / sqrt(d) of a Cholesky pivot.  Device: hardware reciprocal square root + one Newton step (about 1 ulp) -- IEEE
// sqrt followed by an IEEE divide is ~120 cycles longer per pivot and the pivots are the serial chain of every
// factorisation here; a pivot d <= 0 (or NaN) gives NaN, so a non-SPD matrix still shows up in-band like jax
// cho_factor.  Host (tests/cpu_harness): 1 / sqrtf.
HD float chol_pivot_inv(float d) {
#ifdef __CUDA_ARCH__
  float inv = rsqrtf(d);
  inv = inv * fmaf(-0.5f * d, inv * inv, 1.5f);
  return (d > 0.f) ? inv : __int_as_float(0x7fc00000);
#else
  return 1.f / sqrtf(d);
#endif
}

#ifdef __CUDACC__
// Cholesky of a K x K SPD matrix (K <= 32, lower triangle of Mat valid) by ONE warp with the rows in registers: lane i
// keeps row i, the update is right-looking with one shuffle per (column, row) pair, no barrier and no shared-memory
// traffic inside.  Every entry sees the same operations in the same order as the shared-memory column form
// (s = M_ik - sum_j L_ij L_kj, j ascending, fmaf(-a, b, s); L_ij = s / sqrt(pivot)), so the factor is bitwise the same.
// Writes the strictly-lower factor back into Mat and the reciprocal diagonal into dinv.
template <int K, int LD>
__device__ __forceinline__ void warp_cholesky(float* Mat, float* dinv, int lane) {
  static_assert(K <= 32, "one row per lane");
  float row[K];
  const int r = lane < K ? lane : K - 1;  // spare lanes mirror the last row; their results are discarded
#pragma unroll
  for (int c = 0; c < K; ++c) row[c] = (c <= r) ? Mat[r * LD + c] : 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float sjj = __shfl_sync(0xffffffffu, row[j], j);
    const float dj = chol_pivot_inv(sjj);
    const float lij = row[j] * dj;
    if (lane == j) dinv[j] = dj;
    if (lane > j && lane < K) Mat[lane * LD + j] = lij;
#pragma unroll
    for (int c = j + 1; c < K; ++c) {
      const float lcj = __shfl_sync(0xffffffffu, lij, c);
      row[c] = fmaf(-lij, lcj, row[c]);
    }
  }
}
#endif

}  // namespace mpxb
