// lq_small.cuh -- the sequential LQ sub-problem (reference primal_tvlqr.py:11-52 lqr_step, 103-144 tvlqr, 478-498
// rollout; dual_tvlqr.py:6-19 dual_lqr) for SMALL systems (n, m <= 16: the SRBD model, n = 13, m = 12), built for the
// latency of a single instance: the WHOLE horizon of one environment -- every A_t, B_t, Q_t, R_t, M_t, q_t, r_t, c_t and
// every result K_t, k_t, P_t, p_t, dX, dU -- lives in the shared memory of one CTA (127 KB for SRBD at N = 25), so
// the time loop touches no global memory at all: one coalesced load phase with every request in flight at once, 4 block
// barriers per knot, one coalesced store phase.  The general kernel (lq.cuh) streams each knot through cp.async and
// spends 8.9 us per knot on a 13 x 13 problem; this one spends 2.9 us (E = 1 SRBD: LQ 223 -> 72 us, r02w), 40 % of it
// at the barrier behind the per-column factorisation + solve, the serial chain of the knot.
//
// Per knot (the reference recursion, dense general form):
//   S1   PA = P+ A,  PB = P+ B,  w = p+ + P+ c
//   S2   X = A' PA,  H = B' PA + M',  G = R + B' PB,  h = r + B' w,  g = q + A' w
//   S34  every right-hand-side column thread forms sym(G), factors it in registers (Cholesky, lower; NaN on a non-SPD
//        G like jax cho_factor) and solves its own column of [K k] = -G^-1 [H h]
//   S5   P = sym(Q + X + K' H),  p = g + K' h
// then the rollout on one warp (two lanes per row, warp barriers only) and dV = P dX + p, slope = sum q.dX + sum r.dU
// over all knots at once.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"
#include "lq.cuh"

namespace mpxb {

template <int n, int m> __host__ __device__ inline size_t lq_small_floats(int N) {
  return size_t(N) * (n * n + 2 * n * m + m * m + m + m * n + m + m) + size_t(N + 1) * (2 * n * n + 5 * n) +
         (2 * n * n + 2 * n * m + m * m + 2 * n + m) + 64;
}
template <int n, int m> inline size_t lq_small_bytes(int N) { return lq_small_floats<n, m>(N) * sizeof(float); }

template <int n, int m, int NT>
__global__ void __launch_bounds__(NT, 1) lq_small_kernel(LqArgs a) {
  static_assert(n <= 16 && m <= 16 && NT >= 64, "small systems: one warp rolls out, column threads factor in registers");
  extern __shared__ __align__(16) float sm_small[];
  const int e = blockIdx.x, tid = threadIdx.x, N = a.N;
  if (a.done && a.done[e]) return;
  constexpr int nn = n * n, nm = n * m, mm = m * m;
  float* sA = sm_small;              // [N][n][n]
  float* sB = sA + N * nn;           // [N][n][m]
  float* sQ = sB + N * nm;           // [N+1][n][n]
  float* sR = sQ + (N + 1) * nn;     // [N][m][m]
  float* sM = sR + N * mm;           // [N][n][m]
  float* sq = sM + N * nm;           // [N+1][n]
  float* sr = sq + (N + 1) * n;      // [N][m]
  float* sc = sr + N * m;            // [N+1][n]
  float* sK = sc + (N + 1) * n;      // [N][m][n]
  float* sk = sK + N * nm;           // [N][m]
  float* sP = sk + N * m;            // [N+1][n][n]
  float* sp = sP + (N + 1) * nn;     // [N+1][n]
  float* sdx = sp + (N + 1) * n;     // [N+1][n]
  float* sdu = sdx + (N + 1) * n;    // [N][m]
  float* sdv = sdu + N * m;          // [N+1][n]
  float* wPA = sdv + (N + 1) * n;    // [n][n]
  float* wPB = wPA + nn;             // [n][m]
  float* wX = wPB + nm;              // [n][n]
  float* wH = wX + nn;               // [m][n]
  float* wG = wH + nm;               // [m][m]
  float* ww = wG + mm;               // [n]
  float* wg = ww + n;                // [n]
  float* wh = wg + n;                // [m]
  float* sred = wh + m;              // [NT / 32]
  // rollout scratch [N][n n + n] (closed-loop maps) over Q | R | M, which are dead once the backward pass is done
  static_assert(n * n + m * m + n * m >= n * n + n, "closed-loop maps fit over Q | R | M");
  auto wScr = [&](int t, int q) -> float& { return sQ[t * (nn + n) + q]; };

  // ---- load phase: the whole horizon, coalesced, every request in flight
  {
    const size_t eN = size_t(e) * N, eN1 = size_t(e) * (N + 1);
    // 4-byte asynchronous copies (LDGSTS): no register staging, every request of the thread in flight before the wait
    auto ld = [&](float* dst, const float* src, int cnt) {
      for (int i = tid; i < cnt; i += NT) cp_async4(dst + i, src + i);
    };
    ld(sA, a.A + eN * nn, N * nn);
    ld(sB, a.B + eN * nm, N * nm);
    ld(sQ, a.Q + eN1 * nn, (N + 1) * nn);
    ld(sR, a.R + eN * mm, N * mm);
    ld(sM, a.M + eN * nm, N * nm);
    ld(sq, a.q + eN1 * n, (N + 1) * n);
    ld(sr, a.r + eN * m, N * m);
    ld(sc, a.c + eN1 * n, (N + 1) * n);
    cp_async_commit();
    cp_async_wait_all();
  }
  __syncthreads();
  // terminal value function: P_N = Q_N, p_N = q_N (primal_tvlqr.py:136)
  for (int i = tid; i < nn; i += NT) sP[N * nn + i] = sQ[N * nn + i];
  for (int i = tid; i < n; i += NT) sp[N * n + i] = sq[N * n + i];
  __syncthreads();

  for (int t = N - 1; t >= 0; --t) {
    const float *A = sA + t * nn, *B = sB + t * nm, *Pn = sP + (t + 1) * nn, *pn = sp + (t + 1) * n, *cb = sc + (t + 1) * n;
    // ---- S1
    for (int it = tid; it < nn + nm + n; it += NT) {
      if (it < nn) {
        const int i = it / n, j = it - i * n;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(Pn[i * n + k], A[k * n + j], s);
        wPA[it] = s;
      } else if (it < nn + nm) {
        const int q = it - nn, i = q / m, j = q - i * m;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(Pn[i * n + k], B[k * m + j], s);
        wPB[q] = s;
      } else {
        const int i = it - nn - nm;
        float s = pn[i];
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(Pn[i * n + k], cb[k], s);
        ww[i] = s;
      }
    }
    __syncthreads();
    // ---- S2
    for (int it = tid; it < nn + nm + mm + m + n; it += NT) {
      if (it < nn) {
        const int i = it / n, j = it - i * n;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(A[k * n + i], wPA[k * n + j], s);
        wX[it] = s;
      } else if (it < nn + nm) {
        const int q = it - nn, i = q / n, j = q - i * n;   // H [m][n]
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(B[k * m + i], wPA[k * n + j], s);
        wH[q] = s + sM[t * nm + j * m + i];
      } else if (it < nn + nm + mm) {
        const int q = it - nn - nm, i = q / m, j = q - i * m;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(B[k * m + i], wPB[k * m + j], s);
        wG[q] = sR[t * mm + q] + s;
      } else if (it < nn + nm + mm + m) {
        const int i = it - nn - nm - mm;
        float s = sr[t * m + i];
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(B[k * m + i], ww[k], s);
        wh[i] = s;
      } else {
        const int i = it - nn - nm - mm - m;
        float s = sq[t * n + i];
#pragma unroll
        for (int k = 0; k < n; ++k) s = fmaf(A[k * n + i], ww[k], s);
        wg[i] = s;
      }
    }
    __syncthreads();
    // ---- S34: [K k] = -sym(G)^-1 [H h], one thread per right-hand-side column, factor in registers
    for (int col = tid; col <= n; col += NT) {
      float L[m * (m + 1) / 2];
#pragma unroll
      for (int i = 0; i < m; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) L[i * (i + 1) / 2 + j] = 0.5f * (wG[i * m + j] + wG[j * m + i]);
      float gi[m];
#pragma unroll
      for (int j = 0; j < m; ++j) {
        float d = L[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k2 = 0; k2 < j; ++k2) d = fmaf(-L[j * (j + 1) / 2 + k2], L[j * (j + 1) / 2 + k2], d);
        const float inv = chol_pivot_inv(d);
        gi[j] = inv;
#pragma unroll
        for (int i = j + 1; i < m; ++i) {
          float s = L[i * (i + 1) / 2 + j];
#pragma unroll
          for (int k2 = 0; k2 < j; ++k2) s = fmaf(-L[i * (i + 1) / 2 + k2], L[j * (j + 1) / 2 + k2], s);
          L[i * (i + 1) / 2 + j] = s * inv;
        }
      }
      float y[m];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        float s = -((col < n) ? wH[i * n + col] : wh[i]);
#pragma unroll
        for (int k2 = 0; k2 < i; ++k2) s = fmaf(-L[i * (i + 1) / 2 + k2], y[k2], s);
        y[i] = s * gi[i];
      }
#pragma unroll
      for (int i = m - 1; i >= 0; --i) {
        float s = y[i];
#pragma unroll
        for (int k2 = i + 1; k2 < m; ++k2) s = fmaf(-L[k2 * (k2 + 1) / 2 + i], y[k2], s);
        y[i] = s * gi[i];
      }
      if (col < n) {
#pragma unroll
        for (int i = 0; i < m; ++i) sK[t * nm + i * n + col] = y[i];
      } else {
#pragma unroll
        for (int i = 0; i < m; ++i) sk[t * m + i] = y[i];
      }
    }
    __syncthreads();
    // ---- S5: P_t = sym(Q_t + X + K' H) on the unordered pairs, p_t = g + K' h
    {
      const float *K = sK + t * nm, *Q = sQ + t * nn;
      float* Pt = sP + t * nn;
      for (int it = tid; it < nn + n; it += NT) {
        if (it < nn) {
          const int i = it / n, j = it - i * n;
          if (i > j) continue;
          float zij = wX[i * n + j], zji = wX[j * n + i];
#pragma unroll
          for (int k = 0; k < m; ++k) {
            zij = fmaf(K[k * n + i], wH[k * n + j], zij);
            zji = fmaf(K[k * n + j], wH[k * n + i], zji);
          }
          const float s = 0.5f * ((Q[i * n + j] + zij) + (Q[j * n + i] + zji));
          Pt[i * n + j] = s;
          Pt[j * n + i] = s;
        } else {
          const int i = it - nn;
          float s = wg[i];
#pragma unroll
          for (int k = 0; k < m; ++k) s = fmaf(K[k * n + i], wh[k], s);
          sp[t * n + i] = s;
        }
      }
    }
    __syncthreads();
  }

  // ---- forward rollout (primal_tvlqr.py:484-498).  The serial chain over the horizon is one 13-term mat-vec per
  // knot: the closed-loop maps  Acl_t = A_t + B_t K_t,  ccl_t = B_t k_t + c_{t+1}  are formed for all knots at once
  // (over A_t / c_{t+1}, which nothing reads afterwards), warp 0 walks  dX_{t+1} = Acl_t dX_t + ccl_t  with two lanes
  // per row and warp barriers only, then  dU_t = K_t dX_t + k_t  for all knots at once.
  for (int it = tid; it < N * (nn + n); it += NT) {
    const int t = it / (nn + n), q = it - t * (nn + n);
    if (q < nn) {
      const int i = q / n, j = q - i * n;
      float s = sA[t * nn + q];
#pragma unroll
      for (int k = 0; k < m; ++k) s = fmaf(sB[t * nm + i * m + k], sK[t * nm + k * n + j], s);
      wScr(t, q) = s;
    } else {
      const int i = q - nn;
      float s = sc[(t + 1) * n + i];
#pragma unroll
      for (int k = 0; k < m; ++k) s = fmaf(sB[t * nm + i * m + k], sk[t * m + k], s);
      wScr(t, q) = s;
    }
  }
  __syncthreads();
  if (tid < 32) {
    const int row = tid >> 1, part = tid & 1;
    if (tid < n) sdx[tid] = sc[tid];
    __syncwarp();
    for (int t = 0; t < N; ++t) {
      const float* dx = sdx + t * n;
      const float* Acl = &wScr(t, 0);
      float s = 0.f;
      if (row < n)
        for (int j = part; j < n; j += 2) s = fmaf(Acl[row * n + j], dx[j], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (row < n && part == 0) sdx[(t + 1) * n + row] = s + Acl[nn + row];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int it = tid; it < N * m; it += NT) {
    const int t = it / m, i = it - t * m;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < n; ++j) s = fmaf(sK[t * nm + i * n + j], sdx[t * n + j], s);
    sdu[it] = s + sk[it];
  }
  __syncthreads();
  // ---- dual (dual_tvlqr.py:19): dV_t = P_t dX_t + p_t ; slope = sum q.dX + sum r.dU (fixed-order reduction)
  float slope_acc = 0.f;
  for (int it = tid; it < (N + 1) * n; it += NT) {
    const int t = it / n, i = it - t * n;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < n; ++j) s = fmaf(sP[t * nn + i * n + j], sdx[t * n + j], s);
    sdv[it] = s + sp[it];
    slope_acc = fmaf(sq[it], sdx[it], slope_acc);
  }
  for (int it = tid; it < N * m; it += NT) slope_acc = fmaf(sr[it], sdu[it], slope_acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) slope_acc += __shfl_xor_sync(0xffffffffu, slope_acc, o);
  if ((tid & 31) == 0) sred[tid >> 5] = slope_acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NT / 32; ++w) s += sred[w];
    a.slope[e] = s;
  }
  // ---- store phase
  {
    const size_t eN = size_t(e) * N, eN1 = size_t(e) * (N + 1);
    auto stg = [&](float* dst, const float* src, int cnt) {
      for (int i = tid; i < cnt; i += NT) dst[i] = src[i];
    };
    stg(a.dX + eN1 * n, sdx, (N + 1) * n);
    stg(a.dU + eN * m, sdu, N * m);
    stg(a.dV + eN1 * n, sdv, (N + 1) * n);
    stg(a.K + eN * nm, sK, N * nm);
    stg(a.k + eN * m, sk, N * m);
    stg(a.P + eN1 * nn, sP, (N + 1) * nn);
    stg(a.p + eN1 * n, sp, (N + 1) * n);
  }
}

}  // namespace mpxb
