// srbd.cuh -- single-rigid-body quadruped: dynamics, cost and residuals, templated on the scalar type.
// Follows reference mpx/utils/models.py:6-37 (quadruped_srbd_dynamics) and
// mpx/utils/objectives.py:15-75 (quadruped_srbd_obj, quadruped_srbd_hessian_gn).
#pragma once
#include "common.cuh"

namespace mpxb {

struct SrbdParams {
  float mass, dt;
  float inertia[9];
  float inertia_inv[9];
};

constexpr int SRBD_N = 13, SRBD_M = 12, SRBD_NREF = 17, SRBD_NPAR = 16, SRBD_NW = 24;

// x = [p3 | quat4 | dp3 | omega3], u = grf12, par = [contact4 | foot12]  (models.py:8-18)
template <class T> HD void srbd_dynamics(const SrbdParams& P, const T* x, const T* u, const float* par, T* xn) {
  const T* p = x;
  const T* quat = x + 3;
  const T* dp = x + 7;
  const T* om = x + 10;
  const float* contact = par;
  const float* legs = par + 4;
  const float inv_m = 1.f / P.mass;
  T fsum[3] = {T(0.f), T(0.f), T(0.f)};
  T wrench[3] = {T(0.f), T(0.f), T(0.f)};
  for (int f = 0; f < 4; ++f) {
    T arm[3], t[3];
    for (int i = 0; i < 3; ++i) {
      fsum[i] = fsum[i] + u[3 * f + i] * contact[f];
      arm[i] = legs[3 * f + i] - p[i];
    }
    cross3(arm, u + 3 * f, t);
    for (int i = 0; i < 3; ++i) wrench[i] = wrench[i] + t[i] * contact[f];
  }
  const float grav[3] = {0.f, 0.f, -9.81f};
  T dpn[3];
  for (int i = 0; i < 3; ++i) dpn[i] = dp[i] + (grav[i] + inv_m * fsum[i]) * P.dt;
  T R[9];
  quat_to_mat(quat, R);
  T Iw[3], tb[3], gyro[3], rhs[3], omn[3];
  for (int i = 0; i < 3; ++i) {
    Iw[i] = P.inertia[3 * i] * om[0] + P.inertia[3 * i + 1] * om[1] + P.inertia[3 * i + 2] * om[2];
    tb[i] = R[i] * wrench[0] + R[3 + i] * wrench[1] + R[6 + i] * wrench[2];  // R^T wrench
  }
  cross3(om, Iw, gyro);
  for (int i = 0; i < 3; ++i) rhs[i] = tb[i] - gyro[i];
  for (int i = 0; i < 3; ++i)
    omn[i] = om[i] + (P.inertia_inv[3 * i] * rhs[0] + P.inertia_inv[3 * i + 1] * rhs[1] + P.inertia_inv[3 * i + 2] * rhs[2]) * P.dt;
  for (int i = 0; i < 3; ++i) xn[i] = p[i] + dpn[i] * P.dt;
  quat_integrate(quat, omn, P.dt, xn + 3);
  for (int i = 0; i < 3; ++i) { xn[7 + i] = dpn[i]; xn[10 + i] = omn[i]; }
}

template <class T> HD T quad3(const float* W, int ldw, int off, const T* r) {
  T s(0.f);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) s = s + r[i] * (W[(off + i) * ldw + off + j] * r[j]);
  return s;
}

// objectives.py:15-37.  W is [24,24] row-major; ref = reference[t] (17 entries).
template <class T> HD T srbd_cost(const float* W, const float* ref, const T* x, const T* u, bool terminal) {
  T r[3];
  T cost(0.f);
  for (int i = 0; i < 3; ++i) r[i] = x[i] - ref[i];
  cost = cost + quad3(W, SRBD_NW, 0, r);
  quat_sub(x + 3, ref + 3, r);
  cost = cost + quad3(W, SRBD_NW, 3, r);
  for (int i = 0; i < 3; ++i) r[i] = x[7 + i] - ref[7 + i];
  cost = cost + quad3(W, SRBD_NW, 6, r);
  for (int i = 0; i < 3; ++i) r[i] = x[10 + i] - ref[10 + i];
  cost = cost + quad3(W, SRBD_NW, 9, r);
  if (!terminal) {
    for (int i = 0; i < 12; ++i)
      for (int j = 0; j < 12; ++j) cost = cost + u[i] * (W[(12 + i) * SRBD_NW + 12 + j] * u[j]);
    for (int f = 0; f < 4; ++f) {
      T cone = 0.7f * u[3 * f + 2] - xsqrt(u[3 * f + 1] * u[3 * f + 1] + u[3 * f] * u[3 * f] + 1e-2f);
      cost = cost + xclip(penalty(cone, 0.1f, 5.f), 1e-6f, 1e6f);
    }
  }
  return 0.5f * cost;
}

}  // namespace mpxb
