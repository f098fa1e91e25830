// sqp.cuh -- workspace layout and the model-independent kernels of one SQP iteration:
// filter line-search decision + iterate update + convergence bookkeeping (optimizers.py:468-578,
// mpc_wrapper.py:456-467) and the warm-start shift (mpc_wrapper.py:79-98).
#pragma once
#include "model_io.cuh"
#include <cuda_runtime.h>
#include "common.cuh"
#include "lq3.cuh"

namespace mpxb {


// Offsets (in floats) of each per-batch array inside the caller's workspace.
struct WsLayout {
  size_t A, B, Q, R, M, q, r, c, K, k, P, p, dX, dU, dV, cost, ls_theta, ls_phi, slope, g, Xt, Ut, Vt, last_cost,
      done, scan, theta_k, phi_k, need, ls_vc, vc_k, dv2, dvc, th_part, ph_part, vc_part, need_count,
      // structured path (lq3.cuh): [A | B] blocks, packed knot blocks, compact K / P, packed p | k
      ABc, PK, Kc, Pc, OKc, total;
};

inline size_t align64(size_t x) { return (x + 63) & ~size_t(63); }

inline WsLayout make_layout(int E, int N, int n, int m, size_t scan_floats, const Lq3Layout* s3 = nullptr) {
  WsLayout L;
  size_t o = 0;
  auto take = [&](size_t cnt) { size_t r = o; o = align64(o + cnt); return r; };
  const size_t e = size_t(E);
  L.A = take(e * N * n * n);
  L.B = take(e * N * n * m);
  L.Q = take(e * (N + 1) * n * n);
  L.R = take(e * N * m * m);
  L.M = take(e * N * n * m);
  L.q = take(e * (N + 1) * n);
  L.r = take(e * N * m);
  L.c = take(e * (N + 1) * n);
  L.K = take(e * N * m * n);
  L.k = take(e * N * m);
  L.P = take(e * (N + 1) * n * n);
  L.p = take(e * (N + 1) * n);
  L.dX = take(e * (N + 1) * n);
  L.dU = take(e * N * m);
  L.dV = take(e * (N + 1) * n);
  L.cost = take(e * (N + 1));
  L.ls_theta = take(e * NALPHA);
  L.ls_phi = take(e * NALPHA);
  L.slope = take(e);
  L.g = take(e);
  L.Xt = take(e * (N + 1) * n);  // ping-pong iterate for max_iter > 1
  L.Ut = take(e * N * m);
  L.Vt = take(e * (N + 1) * n);
  L.last_cost = take(e);
  L.done = take(e);  // int32
  L.scan = take(scan_floats);
  L.theta_k = take(e);
  L.phi_k = take(e);
  L.need = take(e);  // int32
  L.ls_vc = take(e * NALPHA);
  L.vc_k = take(e);
  L.dv2 = take(e);
  L.dvc = take(e);
  L.th_part = take(e * NALPHA * (N + 1));
  L.ph_part = take(e * NALPHA * (N + 1));
  L.vc_part = take(e * NALPHA * (N + 1));
  L.need_count = take(1 + e);  // int32: the count, then the list of environments that rejected alpha = 1
  L.ABc = L.PK = L.Kc = L.Pc = L.OKc = 0;
  if (s3) {
    L.ABc = take(e * N * s3->szAB);
    L.PK = take(e * (N + 1) * s3->PKS);
    L.Kc = take(e * N * s3->szK);
    L.Pc = take(e * (N + 1) * s3->szP);
    L.OKc = take(e * (N + 1) * s3->OKS);
  }
  L.total = o;
  return L;
}

struct SelectArgs {
  int E, N, n, m;
  const float *X_in, *U_in, *V_in, *dX, *dU, *dV;
  const float *theta_k, *phi_k, *ls_theta, *ls_phi, *slope;  // theta_k / phi_k: written by the trial kernel
  float *X, *U, *V;
  float* g_out;
  float* last_cost;      // per env, offline stopping rule state (may be null)
  int32_t* done;         // per env (may be null)
  int32_t *alpha_index, *iters, *status;  // may be null
  int check_convergence;  // 1 when max_iter > 1
  int first_iter;
  int merit;  // MPXB_LS_*: 0 filter, 1 merit-function line search (parallel_line_search), 2 sequential Armijo (legacy)
  const float *ls_vc, *vc_k, *dv2, *dvc;
  int stop_rule;  // MPXB_STOP_*: 0 runOffline's rule after the step, 1 the legacy continuation criterion before it
};

// One CTA per env: theta_k, phi_k, filter tests for the 11 trial points, chosen alpha, iterate update.
__global__ void select_update_kernel(SelectArgs a) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const int N = a.N, n = a.n, m = a.m;
  const size_t nx = size_t(N + 1) * n, nu = size_t(N) * m;
  const float* Xi = a.X_in + e * nx;
  const float* Ui = a.U_in + e * nu;
  const float* Vi = a.V_in + e * nx;
  float* Xo = a.X + e * nx;
  float* Uo = a.U + e * nu;
  float* Vo = a.V + e * nx;
  if (a.done && a.done[e]) {  // converged earlier: carry the iterate through the ping-pong buffers
    for (size_t i = tid; i < nx; i += blockDim.x) { Xo[i] = Xi[i]; Vo[i] = Vi[i]; }
    for (size_t i = tid; i < nu; i += blockDim.x) Uo[i] = Ui[i];
    return;
  }
  __shared__ int s_idx, s_stop;
  __shared__ float s_alpha, s_part[32];
  const bool legacy = a.stop_rule == 1 && a.check_convergence && a.done;
  float d2 = 0.f;
  if (legacy) {
    // |d|^2 = sum dX^2 + sum dU^2 (+ sum dV^2, which the trial kernel has reduced) for the continuation criterion
    const float* dXe = a.dX + e * nx;
    const float* dUe = a.dU + e * nu;
    for (size_t i = tid; i < nx; i += blockDim.x) d2 = fmaf(dXe[i], dXe[i], d2);
    for (size_t i = tid; i < nu; i += blockDim.x) d2 = fmaf(dUe[i], dUe[i], d2);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, k);
    if ((tid & 31) == 0) s_part[tid >> 5] = d2;
  }
  if (tid == 0) s_stop = 0;
  __syncthreads();
  if (tid == 0) {
    // theta_k = sum c^2, phi_k = sum_t cost_t of the current iterate (reduced by the trial kernel)
    const float theta_k = a.theta_k[e], phi_k = a.phi_k[e];
    const float slope = a.slope[e];
    // first accepted trial point (largest alpha), else index -1 of the 11 trials (optimizers.py:568-572).  Ascending
    // with early exit: when the line search ran lazily the entries behind the first accepted one were never evaluated.
    int best = NALPHA - 1;
    bool exhausted = false;
    if (a.merit) {
      // parallel_line_search (optimizers.py:295-354): first trial point that passes the Armijo test on the merit
      // function; index 0 when none does (argmin of an all-True acceptance vector).  MPXB_LS_ARMIJO: the sequential
      // line_search (old_optimazers.py:173-251) on the same points -- none accepted = the last point, no_errors false
      float rho, mk, ms;
      merit_terms(theta_k, phi_k, a.vc_k[e], a.dv2[e], a.dvc[e], slope, &rho, &mk, &ms, a.merit);
      if (legacy) {
        // continuation_criterion (old_optimazers.py:435-447) on the direction just computed, before the step
        float dd = a.dv2[e];
        for (int w = 0; w < int(blockDim.x >> 5); ++w) dd += s_part[w];
        const bool slope_ok = fabsf(ms) > 1e-4f, delta_ok = dd > 0.f, c_ok = theta_k > 1e-4f;
        if (!((slope_ok && delta_ok) || c_ok)) {
          s_stop = 1;
          a.done[e] = 1;
          if (a.g_out) a.g_out[e] = phi_k;
          if (a.iters && a.first_iter) a.iters[e] = 0;
          if (a.status && a.first_iter) a.status[e] = 0;
        }
      }
      best = (a.merit == 2) ? NALPHA - 1 : 0;
      exhausted = (a.merit == 2);
      for (int i = 0; i < NALPHA && !s_stop; ++i) {
        if (!merit_rejected(a.ls_theta[e * NALPHA + i], a.ls_phi[e * NALPHA + i], a.ls_vc[e * NALPHA + i], rho, mk, ms,
                            exp2f(-float(i)))) {
          best = i;
          // alpha_min = 2^-10: a step accepted AT alpha_min still ends the loop (no_errors = alpha > alpha_min)
          exhausted = (a.merit == 2) && (i == NALPHA - 1);
          break;
        }
      }
    } else {
      for (int i = 0; i < NALPHA; ++i) {
        if (filter_accept(a.ls_theta[e * NALPHA + i], a.ls_phi[e * NALPHA + i], theta_k, phi_k, slope, exp2f(-float(i)))) {
          best = i;
          break;
        }
      }
    }
    s_idx = exhausted ? -(best + 1) : best;   // negative: MPXB_LS_ARMIJO ran out of step sizes (status bit 1)
    s_alpha = exp2f(-float(best));
    if (!s_stop) {
      if (a.g_out) a.g_out[e] = phi_k;
      if (a.alpha_index) a.alpha_index[e] = best;
      if (a.iters) a.iters[e] = (a.first_iter ? 0 : a.iters[e]) + 1;
      if (exhausted && a.done) a.done[e] = 1;
    }
    if (a.check_convergence && a.done && !legacy) {
      // mpc_wrapper.py:456-467: l2_cost = g*g at the NEW point; stop when last_cost - l2_cost < 1e-3 and sum c^2 < 1e-5
      const float g_new = a.ls_phi[e * NALPHA + best];
      const float c2_new = a.ls_theta[e * NALPHA + best];
      const float l2 = g_new * g_new;
      const float last = a.first_iter ? 1e10f : a.last_cost[e];
      if ((last - l2 < 1e-3f) && (c2_new < 1e-5f)) a.done[e] = 1;
      a.last_cost[e] = l2;
    }
  }
  __syncthreads();
  if (s_stop) {  // the legacy criterion ended the loop: the iterate goes through unchanged
    for (size_t i = tid; i < nx; i += blockDim.x) { Xo[i] = Xi[i]; Vo[i] = Vi[i]; }
    for (size_t i = tid; i < nu; i += blockDim.x) Uo[i] = Ui[i];
    return;
  }
  const float alpha = s_alpha;
  const float* dX = a.dX + e * nx;
  const float* dU = a.dU + e * nu;
  const float* dV = a.dV + e * nx;
  int bad = 0;
  for (size_t i = tid; i < nx; i += blockDim.x) {
    const float xv = fmaf(alpha, dX[i], Xi[i]);
    const float vv = fmaf(alpha, dV[i], Vi[i]);
    Xo[i] = xv;
    Vo[i] = vv;
    bad |= !(isfinite(xv) && isfinite(vv));
  }
  for (size_t i = tid; i < nu; i += blockDim.x) {
    const float uv = fmaf(alpha, dU[i], Ui[i]);
    Uo[i] = uv;
    bad |= !isfinite(uv);
  }
  bad = __syncthreads_or(bad);
  if (tid == 0 && a.status) a.status[e] = (bad ? 1 : 0) | (s_idx < 0 ? 2 : 0);
}

// update_and_extract_helper (mpc_wrapper.py:79-98 / 261-282) and the SRBD shift (mpc_wrapper_srbd.py:135-139)
__global__ void shift_extract_kernel(int E, int N, int n, int m, int nj, int shift, int wb, const float* X,
                                     const float* U, const float* V, const float* x0, const float* u_ref, float* X0,
                                     float* U0, float* V0, float* tau, float* q, float* dq) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const size_t nx = size_t(N + 1) * n, nu = size_t(N) * m;
  const float* Xe = X + e * nx;
  const float* Ue = U + e * nu;
  const float* Ve = V + e * nx;
  float* X0e = X0 + e * nx;
  float* U0e = U0 + e * nu;
  float* V0e = V0 + e * nx;
  const bool nan = wb && (Ue[0] != Ue[0]);  // jnp.isnan(U[0,0]); the SRBD wrapper has no guard
  if (wb && tau) {
    // safe: tau = U[0,:nj], q = X[0,7:7+nj], dq = X[1,13+nj:13+2nj]; unsafe: same rows of the PREVIOUS plan's row 1
    for (int i = tid; i < nj; i += blockDim.x) {
      tau[e * nj + i] = nan ? U0e[m + i] : Ue[i];
      if (q) q[e * nj + i] = nan ? X0e[n + 7 + i] : Xe[7 + i];
      if (dq) dq[e * nj + i] = nan ? X0e[n + 13 + nj + i] : Xe[n + 13 + nj + i];
    }
  }
  __syncthreads();
  // (32-bit indices, no division by the runtime row length: the flat size_t form spent most of the kernel in 64-bit
  // divide sequences)
  const int nxi = (N + 1) * n, nui = N * m, nt = int(blockDim.x);
  if (nan) {
    for (int t = 0; t <= N; ++t)
      for (int j = tid; j < n; j += nt) { X0e[t * n + j] = x0[size_t(e) * n + j]; V0e[t * n + j] = 0.f; }
    for (int t = 0; t < N; ++t)
      for (int j = tid; j < m; j += nt) U0e[t * m + j] = u_ref[j];
    return;
  }
  // rows t <= N - shift are one contiguous block moved down by `shift` rows; the rows behind repeat the last knot
  const int sx = shift * n, su = shift * m;
  for (int i = tid; i < nxi - sx; i += nt) { X0e[i] = Xe[i + sx]; V0e[i] = Ve[i + sx]; }
  for (int t = max(N + 1 - shift, 0); t <= N; ++t)
    for (int j = tid; j < n; j += nt) { X0e[t * n + j] = Xe[N * n + j]; V0e[t * n + j] = Ve[N * n + j]; }
  for (int i = tid; i < nui - su; i += nt) U0e[i] = Ue[i + su];
  for (int t = max(N - shift, 0); t < N; ++t)
    for (int j = tid; j < m; j += nt) U0e[t * m + j] = Ue[(N - 1) * m + j];
}

// Debug / parity views of the structured path: rebuild the dense arrays the reference's LQ solver works with
// (A, B, Q, R, M, q, r, K, k, P, p) from the compact ones.  One CTA per (environment, knot).
struct ExpandArgs {
  int E, N;
  Lq3Layout s;
  const float *ABc, *PK, *Kc, *Pc, *OKc;
  float *A, *B, *Q, *R, *M, *q, *r, *K, *k, *P, *p;
};
__global__ void lq3_expand_kernel(ExpandArgs a) {
  const int e = blockIdx.x / (a.N + 1), t = blockIdx.x - e * (a.N + 1);
  const int tid = threadIdx.x, NT = blockDim.x;
  const Lq3Layout& s = a.s;
  const int n = s.n, m = s.m, nh = s.nh, N = a.N;
  const float* pk = a.PK + (size_t(e) * (N + 1) + t) * s.PKS;
  const float* ok = a.OKc + (size_t(e) * (N + 1) + t) * s.OKS;
  const float* Pc = a.Pc + (size_t(e) * (N + 1) + t) * s.szP;
  // Q_t structure: diag + quaternion block + cone blocks
  auto qent = [&](int i, int j) -> float {
    float v = (i == j) ? pk[s.oQd + i] : 0.f;
    if (i >= 3 && i < 7 && j >= 3 && j < 7) v = pk[s.oQq + (i - 3) * 4 + (j - 3)];
    if (s.ncone > 0 && i >= s.cone0 && j >= s.cone0 && (i - s.cone0) / 3 == (j - s.cone0) / 3) {
      const int blk = (i - s.cone0) / 3;
      v += pk[s.oQc + 9 * blk + 3 * (i - s.cone0 - 3 * blk) + (j - s.cone0 - 3 * blk)];
    }
    return v;
  };
  float* Q = a.Q + (size_t(e) * (N + 1) + t) * n * n;
  float* P = a.P + (size_t(e) * (N + 1) + t) * n * n;
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    const float qv = qent(i, j);
    Q[idx] = qv;
    P[idx] = (i < nh && j < nh) ? Pc[i * s.LH + j] : ((i >= nh && j >= nh) ? qv : 0.f);
  }
  for (int i = tid; i < n; i += NT) {
    a.q[(size_t(e) * (N + 1) + t) * n + i] = pk[s.oq + i];
    a.p[(size_t(e) * (N + 1) + t) * n + i] = (i < nh) ? ok[i] : pk[s.oq + i];
  }
  if (t == N) return;
  const float* Ac = a.ABc + (size_t(e) * N + t) * s.szAB;
  const float* Bc = Ac + s.LH;
  const float* Kc = a.Kc + (size_t(e) * N + t) * s.szK;
  float* A = a.A + (size_t(e) * N + t) * n * n;
  float* B = a.B + (size_t(e) * N + t) * n * m;
  float* M = a.M + (size_t(e) * N + t) * n * m;
  float* R = a.R + (size_t(e) * N + t) * m * m;
  float* K = a.K + (size_t(e) * N + t) * m * n;
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    A[idx] = (j < nh) ? Ac[i * s.LAB + j] : 0.f;
  }
  for (int idx = tid; idx < n * m; idx += NT) {
    const int i = idx / m, j = idx - i * m;
    B[idx] = Bc[i * s.LAB + j];
    M[idx] = 0.f;
  }
  for (int idx = tid; idx < m * m; idx += NT) {
    const int i = idx / m, j = idx - i * m;
    R[idx] = (i == j) ? pk[s.oRd + i] : 0.f;
  }
  for (int idx = tid; idx < m * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    K[idx] = (j < nh) ? Kc[i * s.LH + j] : 0.f;
  }
  for (int i = tid; i < m; i += NT) {
    a.r[(size_t(e) * N + t) * m + i] = pk[s.oR + i];
    a.k[(size_t(e) * N + t) * m + i] = ok[s.LH + i];
  }
}

}  // namespace mpxb
