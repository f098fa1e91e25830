// srbd_kernels.cuh -- SRBD model kernels: fused evaluate + linearise + quadratise (one warp per knot per
// environment, lanes = forward-mode directions) and the 11-alpha trial evaluation of the filter line search.
// Reference: optimizers.py:580-602 (model_evaluator), 150-173 (linearize / hessian_approx / lagrangian),
// 518-540 (trial points); models.py:6-37; objectives.py:15-75.
#pragma once
#include "model_io.cuh"
#include "srbd.cuh"
#include "sqp.cuh"

namespace mpxb {

// grid: ceil(E*(N+1) / (blockDim/32)) ; one warp per (env, knot)
__global__ void srbd_linearize_kernel(SrbdParams P, ModelIO io, LinOut o) {
  constexpr int n = SRBD_N, m = SRBD_M;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * (blockDim.x >> 5) + wib;
  const int N = io.N;
  if (warp >= io.E * (N + 1)) return;
  const int e = warp / (N + 1), t = warp - e * (N + 1);
  if (io.done && io.done[e]) return;
  __shared__ float sG[8][12];  // d quat_sub / d quat (3x4) per warp
  const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
  const float* Ut = io.U + (size_t(e) * N + t) * m;
  const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
  const float* ref = io.reference + (size_t(e) * (N + 1) + t) * SRBD_NREF;
  const float* par = io.parameter + (size_t(e) * (N + 1) + t) * SRBD_NPAR;
  const float* W = io.W + size_t(e) * io.w_stride;
  const bool term = (t == N);

  Dual x[n], u[m];
#pragma unroll
  for (int i = 0; i < n; ++i) x[i] = Dual(Xt[i], lane == i ? 1.f : 0.f);
#pragma unroll
  for (int i = 0; i < m; ++i) u[i] = Dual(term ? 0.f : Ut[i], lane == n + i ? 1.f : 0.f);

  float atv = 0.f;  // column `lane` of [A B] dotted with V_{t+1}
  if (!term) {
    Dual xn[n];
    srbd_dynamics(P, x, u, par, xn);
    const float* Vn = Vt + n;
    const float* Xn = Xt + n;
    float* A = o.A + (size_t(e) * N + t) * n * n;
    float* B = o.B + (size_t(e) * N + t) * n * m;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      atv = fmaf(xn[i].d, Vn[i], atv);
      if (lane < n) A[i * n + lane] = xn[i].d;
      else if (lane < n + m) B[i * m + lane - n] = xn[i].d;
      if (lane == i) o.c[(size_t(e) * (N + 1) + t + 1) * n + i] = xn[i].v - Xn[i];
    }
  }
  if (t == 0 && lane < n) o.c[size_t(e) * (N + 1) * n + lane] = io.x0[size_t(e) * n + lane] - Xt[lane];

  const Dual cst = srbd_cost(W, ref, x, u, term);
  if (lane == 0) o.cost[size_t(e) * (N + 1) + t] = cst.v;
  if (lane < n) o.q[(size_t(e) * (N + 1) + t) * n + lane] = cst.d + atv - Vt[lane];
  else if (lane < n + m && !term) o.r[(size_t(e) * N + t) * m + lane - n] = cst.d + atv;

  // G = d quat_sub(quat, quat_ref) / d quat : lanes 3..6 hold one column each
  {
    Dual rs[3];
    quat_sub(x + 3, ref + 3, rs);
    if (lane >= 3 && lane < 7) {
      sG[wib][0 * 4 + lane - 3] = rs[0].d;
      sG[wib][1 * 4 + lane - 3] = rs[1].d;
      sG[wib][2 * 4 + lane - 3] = rs[2].d;
    }
  }
  __syncwarp();
  const float* G = sG[wib];
  // residual Jacobian wrt x is the 12 x 13 matrix [I3 0 0 0; 0 G 0 0; 0 0 I3 0; 0 0 0 I3] (objectives.py:41-64)
  // WJ(a, j) = sum_b W[a][b] Jx[b][j]
  auto WJ = [&](int a, int j) -> float {
    if (j < 3) return W[a * SRBD_NW + j];
    if (j < 7) return W[a * SRBD_NW + 3] * G[j - 3] + W[a * SRBD_NW + 4] * G[4 + j - 3] + W[a * SRBD_NW + 5] * G[8 + j - 3];
    return W[a * SRBD_NW + j - 1];
  };
  float* Q = o.Q + (size_t(e) * (N + 1) + t) * n * n;
  for (int idx = lane; idx < n * n; idx += 32) {
    const int i = idx / n, j = idx - i * n;
    float s;
    if (i < 3) s = WJ(i, j);
    else if (i < 7) s = G[i - 3] * WJ(3, j) + G[4 + i - 3] * WJ(4, j) + G[8 + i - 3] * WJ(5, j);
    else s = WJ(i - 1, j);
    Q[idx] = s;
  }
  if (!term) {
    float* M = o.M + (size_t(e) * N + t) * n * m;
    for (int idx = lane; idx < n * m; idx += 32) {
      const int i = idx / m, j = idx - i * m;
      float s;
      if (i < 3) s = W[i * SRBD_NW + 12 + j];
      else if (i < 7)
        s = G[i - 3] * W[3 * SRBD_NW + 12 + j] + G[4 + i - 3] * W[4 * SRBD_NW + 12 + j] + G[8 + i - 3] * W[5 * SRBD_NW + 12 + j];
      else s = W[(i - 1) * SRBD_NW + 12 + j];
      M[idx] = s;
    }
    // R = W_uu + Jf^T diag(clip(penalty'', 1e-6, 1e6) * contact) Jf   (objectives.py:66-75)
    float* R = o.R + (size_t(e) * N + t) * m * m;
    const float* contact = ref + 13;
    for (int idx = lane; idx < m * m; idx += 32) {
      const int i = idx / m, j = idx - i * m;
      float s = W[(12 + i) * SRBD_NW + 12 + j];
      const int f = i / 3;
      if (f == j / 3) {
        const float fx = Ut[3 * f], fy = Ut[3 * f + 1], fz = Ut[3 * f + 2];
        const float sq = sqrtf(fy * fy + fx * fx + 1e-2f);
        const float cone = 0.7f * fz - sq;
        const float hp = fminf(fmaxf(penalty_dd(cone, 0.1f, 5.f), 1e-6f), 1e6f) * contact[f];
        const float jf[3] = {-fx / sq, -fy / sq, 0.7f};
        s = fmaf(hp * jf[i - 3 * f], jf[j - 3 * f], s);
      }
      R[idx] = s;
    }
  }
}

// one warp per (env, alpha): lanes stride over knots; deterministic shuffle reduction
__global__ void srbd_trial_kernel(SrbdParams P, ModelIO io, TrialOut o) {
  constexpr int n = SRBD_N, m = SRBD_M;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int N = io.N;
  if (warp >= io.E * o.nia || !trial_selected(o)) return;
  const int e = warp / o.nia, ia = o.ia0 + warp - e * o.nia;
  if (io.done && io.done[e]) return;
  if (o.ia0 > 0 && o.need && !o.need[e]) return;  // alpha = 1 was accepted: the smaller steps are never looked at
  const float alpha = exp2f(-float(ia));
  const float* W = io.W + size_t(e) * io.w_stride;
  float theta = 0.f, phi = 0.f, vc = 0.f;
  for (int t = lane; t <= N; t += 32) {
    const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
    const float* dXt = o.dX + (size_t(e) * (N + 1) + t) * n;
    const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
    const float* dVt = o.merit ? o.dV + (size_t(e) * (N + 1) + t) * n : Vt;
    const float* ref = io.reference + (size_t(e) * (N + 1) + t) * SRBD_NREF;
    const float* par = io.parameter + (size_t(e) * (N + 1) + t) * SRBD_NPAR;
    float x[n], u[m];
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = fmaf(alpha, dXt[i], Xt[i]);
    if (t < N) {
      const float* Ut = io.U + (size_t(e) * N + t) * m;
      const float* dUt = o.dU + (size_t(e) * N + t) * m;
#pragma unroll
      for (int i = 0; i < m; ++i) u[i] = fmaf(alpha, dUt[i], Ut[i]);
    } else {
#pragma unroll
      for (int i = 0; i < m; ++i) u[i] = 0.f;
    }
    phi += srbd_cost(W, ref, x, u, t == N);
    if (t < N) {
      float xn[n];
      srbd_dynamics(P, x, u, par, xn);
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float r = xn[i] - fmaf(alpha, dXt[n + i], Xt[n + i]);
        theta = fmaf(r, r, theta);
        if (o.merit) vc = fmaf(r, fmaf(alpha, dVt[n + i], Vt[n + i]), vc);
      }
    }
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < n; ++i) {
        const float r = io.x0[size_t(e) * n + i] - x[i];
        theta = fmaf(r, r, theta);
        if (o.merit) vc = fmaf(r, fmaf(alpha, dVt[i], Vt[i]), vc);
      }
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    theta += __shfl_xor_sync(0xffffffffu, theta, s);
    phi += __shfl_xor_sync(0xffffffffu, phi, s);
    vc += __shfl_xor_sync(0xffffffffu, vc, s);
  }
  trial_finish(io, o, e, ia, lane, n, theta, phi, vc);
}

}  // namespace mpxb
