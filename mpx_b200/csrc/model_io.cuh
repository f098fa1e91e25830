// model_io.cuh -- argument blocks shared by the model kernels (SRBD and whole-body): the iterate and problem data
// of one mpc call (optimizers.py:609-616), the LQ data a linearise kernel writes, and the trial-point outputs of
// the filter line search (optimizers.py:518-540).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace mpxb {

constexpr int NALPHA = 11;  // alpha_values = exp2(-arange(11))  (optimizers.py:518)

struct ModelIO {
  int E, N;
  const float *reference, *parameter, *W;
  size_t w_stride;  // floats between consecutive envs' W (0 = shared)
  const float *x0, *X, *U, *V;
  const int32_t* done;
  // first environment of a sliced launch (the first-pass trial kernels and their reduction only: E is then the slice
  // length; every pointer keeps addressing the whole batch)
  int e0;
};

struct LinOut {
  float *A, *B, *Q, *R, *M, *q, *r, *c, *cost;
  // Row strides of A / B (dense: n / m).  Structured mode (compact != 0; lq3.cuh): A and B share one block
  // [A_t | B_t] of row stride ldA = ldB per knot (B points at its first column); only the leading 13 + 2 nj columns of
  // A are live, followed by padA zero columns, B by padB zero columns; Q / R / M / q / r are not written; instead every
  // knot gets one packed block PK[t] of `pks` floats:
  //   diag(Q_t) @oQd | q_t @oq | c_t @oc | 4x4 quaternion block of Q_t @oQq | 3x3 cone blocks @oQc | diag(R_t) @oRd | r_t @oR
  // (the layout of Lq3Dims).  c is still written densely as well (the line search reads it).
  int ldA, ldB;
  int padA, padB;
  int compact;
  float* PK;
  int pks, oQd, oq, oc, oQq, oQc, oRd, oR;
};

struct TrialOut {
  const float *dX, *dU;
  float *ls_theta, *ls_phi;
  // lazy filter line search: the launch covers trial points ia0 .. ia0+nia-1.  The warp of trial point 0 also
  // reduces theta_k = sum c^2 and phi_k = sum cost of the CURRENT iterate and records whether alpha = 1 is rejected
  // (need[e]); a launch with ia0 > 0 skips every environment whose need flag is clear.
  const float *c, *cost, *slope;
  float *theta_k, *phi_k;
  int32_t* need;
  int ia0, nia;
  // merit-function line search (optimizers.py:185-217, 295-354, 637-669), selected by MPXB_OPT_LINE_SEARCH: every
  // trial point also needs vc = sum((V + alpha dV) * c_new); trial point 0 adds sum(V c), sum(dV^2), sum(dV c)
  int merit;
  const float* dV;
  float *ls_vc, *vc_k, *dv2, *dvc;
  // cooperative whole-body trial kernel: per-knot partial sums [E, NALPHA, N+1], reduced by trial_reduce_kernel
  float *th_part, *ph_part, *vc_part;
  // second pass of the lazy search: how many environments rejected alpha = 1 decides ON THE DEVICE which of the two
  // whole-body trial kernels does the work (few -> the cooperative one, many -> one warp per trial point); both are
  // launched, the one whose condition fails exits at once.  sel: 0 run, 1 run iff *need_count <= thr, 2 iff > thr
  int32_t* need_count;   // [1 + E]: count, then the environments that rejected alpha = 1 in arrival order
  int sel, thr;
  // a single-point pass behind alpha = 1 (ia0 > 0, nia == 1) clears need[e] when its point is accepted, so that the
  // remaining points are only evaluated for the environments that reject it as well
  int update_need;
};

#if defined(__CUDACC__)
__device__ __forceinline__ bool trial_selected(const TrialOut& o) {
  if (o.sel == 0) return true;
  const int c = *o.need_count;
  return (o.sel == 1) ? (c <= o.thr) : (c > o.thr);
}
#endif

// whole_body_interface (mpc_utils.py:275-313) arguments: see wb_interface_kernel
struct WbiArgs {
  int E;
  const float *Kp, *Kd, *qpos, *qvel, *grf, *foot_ref, *foot_ref_dot, *contact;
  float *tau, *J;
};

#if defined(__CUDACC__)
// The 3-case filter acceptance test of parallel_filter_line_search (optimizers.py:541-566) with JAX's NaN semantics
__device__ __forceinline__ bool filter_accept(float theta_new, float phi_new, float theta_k, float phi_k, float slope,
                                              float alpha) {
  const float theta_max = 1e-2f, theta_min = 1e-6f, eta = 1e-4f, gamma_phi = 1e-6f, gamma_theta = 1e-6f;
  const bool cond1 = theta_new > theta_max;
  const bool case1 = cond1 && (theta_new < (1.f - gamma_theta) * theta_k);
  const bool cond2 = (fmaxf(theta_new, theta_k) < theta_min) && (slope < 0.f);
  // jnp.maximum propagates NaN; fmaxf does not -> test explicitly
  const bool cond2n = cond2 && !(theta_new != theta_new) && !(theta_k != theta_k);
  const bool case2 = cond2n && (phi_new < phi_k + eta * alpha * slope);
  const bool cond3 = !(cond1 || cond2n);
  const bool case3 = cond3 && ((phi_new < phi_k - gamma_phi * phi_k) || (theta_new < (1.f - gamma_theta) * theta_k));
  return case1 || case2 || case3;
}

// Merit line search: rho (merit_rho, optimizers.py:185-198), the current merit g + sum((V + rho/2 c) c) and its slope
// sum(q dX) + sum(r dU) + 2 sum(dV c) - rho sum(c c) (optimizers.py:201-217)
// mode: the MPXB_LS_* value (1: optimizers.py:201-217, 2 sum(dV c); 2: the legacy slope old_optimazers.py:154-170, sum(dV c))
__device__ __forceinline__ void merit_terms(float theta_k, float phi_k, float vc_k, float dv2, float dvc, float slope_lq,
                                            float* rho, float* merit_k, float* mslope, int mode = 1) {
  *rho = (theta_k > 1e-12f) ? 2.f * sqrtf(dv2 / theta_k) : 1e-2f;
  *merit_k = phi_k + vc_k + 0.5f * *rho * theta_k;
  *mslope = slope_lq + (mode == 2 ? 1.f : 2.f) * dvc - *rho * theta_k;
}
// parallel_line_search's step_acceptance (True = the Armijo test FAILS); a NaN merit is replaced by the current one
__device__ __forceinline__ bool merit_rejected(float theta_new, float phi_new, float vc_new, float rho, float merit_k,
                                               float mslope, float alpha) {
  float mn = phi_new + vc_new + 0.5f * rho * theta_new;
  if (mn != mn) mn = merit_k;
  return mn > merit_k + alpha * 1e-4f * mslope;
}

// Tail of the trial kernels (one warp per (env, trial point)): store theta / phi; trial point 0 adds theta_k, phi_k
// of the current iterate (fixed-order warp reduction: deterministic) and the lazy-evaluation flag.
__device__ __forceinline__ void trial_finish(const ModelIO& io, const TrialOut& o, int e, int ia, int lane, int n,
                                             float theta, float phi, float vc) {
  if (lane == 0) {
    o.ls_theta[e * NALPHA + ia] = theta;
    o.ls_phi[e * NALPHA + ia] = phi;
    if (o.merit) o.ls_vc[e * NALPHA + ia] = vc;
  }
  if (ia != 0) {
    if (o.update_need && o.need && lane == 0) {
      bool ok;
      if (o.merit) {
        float rho, mk, ms;
        merit_terms(o.theta_k[e], o.phi_k[e], o.vc_k[e], o.dv2[e], o.dvc[e], o.slope[e], &rho, &mk, &ms, o.merit);
        ok = !merit_rejected(theta, phi, vc, rho, mk, ms, exp2f(-float(ia)));
      } else {
        ok = filter_accept(theta, phi, o.theta_k[e], o.phi_k[e], o.slope[e], exp2f(-float(ia)));
      }
      if (ok) o.need[e] = 0;
    }
    return;
  }
  const int N = io.N;
  const size_t nx = size_t(N + 1) * n;
  const float* c = o.c + size_t(e) * nx;
  float th = 0.f, ph = 0.f;
  for (size_t i = lane; i < nx; i += 32) th = fmaf(c[i], c[i], th);
  for (int t = lane; t <= N; t += 32) ph += o.cost[size_t(e) * (N + 1) + t];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    th += __shfl_xor_sync(0xffffffffu, th, s);
    ph += __shfl_xor_sync(0xffffffffu, ph, s);
  }
  float vk = 0.f, d2 = 0.f, dc = 0.f;
  if (o.merit) {
    const float* V = io.V + size_t(e) * nx;
    const float* dV = o.dV + size_t(e) * nx;
    for (size_t i = lane; i < nx; i += 32) {
      vk = fmaf(V[i], c[i], vk);
      d2 = fmaf(dV[i], dV[i], d2);
      dc = fmaf(dV[i], c[i], dc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      vk += __shfl_xor_sync(0xffffffffu, vk, s);
      d2 += __shfl_xor_sync(0xffffffffu, d2, s);
      dc += __shfl_xor_sync(0xffffffffu, dc, s);
    }
  }
  if (lane == 0) {
    o.theta_k[e] = th;
    o.phi_k[e] = ph;
    int nd = 0;
    if (o.merit) {
      o.vc_k[e] = vk; o.dv2[e] = d2; o.dvc[e] = dc;
      float rho, mk, ms;
      merit_terms(th, ph, vk, d2, dc, o.slope[e], &rho, &mk, &ms, o.merit);
      nd = merit_rejected(theta, phi, vc, rho, mk, ms, 1.f) ? 1 : 0;
    } else {
      nd = filter_accept(theta, phi, th, ph, o.slope[e], 1.f) ? 0 : 1;
    }
    if (o.need) {
      o.need[e] = nd;
      // count + compact list of the environments that rejected alpha = 1 (need_count[1..]): the second pass of the
      // cooperative kernel walks the list instead of all E environments.  The list's order depends on the atomics, the
      // results do not: every (env, trial point, knot) item writes its own slot
      if (nd && o.need_count) o.need_count[1 + atomicAdd(o.need_count, 1)] = e;
    }
  }
}

// Sums the per-knot partials of the cooperative trial kernel (fixed order: lanes stride over knots, xor tree) and
// finishes like the warp-per-trial kernels.  One warp per (env, trial point of this launch).
template <int DUMMY>
__global__ void trial_reduce_kernel(ModelIO io, TrialOut o, int n) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= io.E * o.nia || !trial_selected(o)) return;
  const int le = warp / o.nia, e = le + io.e0, ia = o.ia0 + warp - le * o.nia;
  if (io.done && io.done[e]) return;
  if (o.ia0 > 0 && o.need && !o.need[e]) return;
  const int N = io.N;
  const size_t base = (size_t(e) * NALPHA + ia) * (N + 1);
  float theta = 0.f, phi = 0.f, vc = 0.f;
  for (int t = lane; t <= N; t += 32) {
    theta += o.th_part[base + t];
    phi += o.ph_part[base + t];
    if (o.merit) vc += o.vc_part[base + t];
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    theta += __shfl_xor_sync(0xffffffffu, theta, s);
    phi += __shfl_xor_sync(0xffffffffu, phi, s);
    vc += __shfl_xor_sync(0xffffffffu, vc, s);
  }
  trial_finish(io, o, e, ia, lane, n, theta, phi, vc);
}
#endif

}  // namespace mpxb
