// lq.cuh -- the LQ sub-problem of the SQP step: sequential Riccati backward pass, forward rollout and dual
// update, one CTA per environment with the value function resident in shared memory.
//
// Follows reference primal_tvlqr.py:11-52 (lqr_step), 103-144 (tvlqr), 478-498 (rollout) and
// dual_tvlqr.py:6-19 (dual_lqr).  All matrices are FP32; products are register-tiled FFMA loops over
// shared-memory operands padded to a multiple of 4 so the inner loops carry no bounds checks.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace mpxb {

struct LqArgs {
  int E, N;
  // linearisation (inputs), per env: A[N,n,n] B[N,n,m] Q[N+1,n,n] R[N,m,m] M[N,n,m] q[N+1,n] r[N,m] c[N+1,n]
  const float *A, *B, *Q, *R, *M, *q, *r, *c;
  // outputs: K[N,m,n] k[N,m] P[N+1,n,n] p[N+1,n] dX[N+1,n] dU[N,m] dV[N+1,n] slope[E]
  float *K, *k, *P, *p, *dX, *dU, *dV, *slope;
  const int32_t* done;  // per env: skip when != 0 (device-side SQP loop), may be null
};

constexpr int lq_pad(int n) { return (n + 3) & ~3; }

template <int n, int m> struct LqSmemLayout {
  static constexpr int LN = lq_pad(n), LM = lq_pad(m);
  // three n x n buffers (P / X, A / Y, T1), B, PB (n x m), H (m x n), G (m x m), K (m x (n+1)), vectors
  static constexpr int oP = 0;
  static constexpr int oA = oP + LN * LN;
  static constexpr int oT = oA + LN * LN;
  static constexpr int oB = oT + LN * LN;    // [LN][LM]
  static constexpr int oPB = oB + LN * LM;   // [LN][LM]
  static constexpr int oH = oPB + LN * LM;   // [LM][LN]
  static constexpr int oG = oH + LM * LN;    // [LM][LM]
  static constexpr int oK = oG + LM * LM;    // [LM][LN]
  static constexpr int oVec = oK + LM * LN;  // p, w, c, q, h, kk, r : 4*LN + 3*LM
  static constexpr int total = oVec + 4 * LN + 3 * LM + 8;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// C[MT*4, NT*4] = op(A) * op(B); op(A)(i,k) = TA ? A[k*lda+i] : A[i*lda+k]; op(B)(k,j) = TB ? B[j*ldb+k] : B[k*ldb+j].
// All dimensions are multiples of 4 (zero padded).  Each thread owns 4x4 tiles tid, tid+nthreads, ...
template <bool TA, bool TB>
__device__ __forceinline__ void gemm_tiles(float* __restrict__ C, int ldc, const float* __restrict__ A, int lda,
                                           const float* __restrict__ B, int ldb, int MT, int NT, int K, int tid,
                                           int nthreads) {
  for (int tile = tid; tile < MT * NT; tile += nthreads) {
    const int i0 = (tile / NT) * 4, j0 = (tile % NT) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < K; ++kk) {
      float av[4], bv[4];
      if (TA) {
        const float4 t = *reinterpret_cast<const float4*>(A + kk * lda + i0);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
      } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = A[(i0 + a) * lda + kk];
      }
      if (TB) {
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = B[(j0 + b) * ldb + kk];
      } else {
        const float4 t = *reinterpret_cast<const float4*>(B + kk * ldb + j0);
        bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      *reinterpret_cast<float4*>(C + (i0 + a) * ldc + j0) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  }
}

// load a [rows, cols] row-major global matrix into a zero-padded [prow, ld] shared tile
__device__ __forceinline__ void load_padded(float* __restrict__ dst, int ld, int prow, const float* __restrict__ src,
                                            int rows, int cols, int tid, int nthreads) {
  for (int idx = tid; idx < prow * ld; idx += nthreads) {
    const int i = idx / ld, j = idx - i * ld;
    dst[idx] = (i < rows && j < cols) ? src[i * cols + j] : 0.f;
  }
}

// Forward rollout + dual update + slope, shared by the sequential and the scan backward passes
// (primal_tvlqr.py:484-498 / 523-549: the prefix composition of the affine maps evaluated left to right;
// dual_tvlqr.py:19).  dx, dxn: [LN] shared scratch, du: [LM] shared scratch.
template <int n, int m, int NT>
__device__ __forceinline__ void lq_rollout_dual(const LqArgs& a, int e, int tid, float* dx, float* dxn, float* du) {
  const int N = a.N;
  const float* gA = a.A + size_t(e) * N * n * n;
  const float* gB = a.B + size_t(e) * N * n * m;
  const float* gq = a.q + size_t(e) * (N + 1) * n;
  const float* gr = a.r + size_t(e) * N * m;
  const float* gc = a.c + size_t(e) * (N + 1) * n;
  const float* gK = a.K + size_t(e) * N * m * n;
  const float* gk = a.k + size_t(e) * N * m;
  const float* gP = a.P + size_t(e) * (N + 1) * n * n;
  const float* gp = a.p + size_t(e) * (N + 1) * n;
  float* gdX = a.dX + size_t(e) * (N + 1) * n;
  float* gdU = a.dU + size_t(e) * N * m;
  float* gdV = a.dV + size_t(e) * (N + 1) * n;
  const int warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;
  for (int i = tid; i < n; i += NT) dx[i] = gc[i];
  float slope_acc = 0.f;  // meaningful in warp 0 lanes
  __syncthreads();
  for (int t = 0; t <= N; ++t) {
    // dV_t = P_t dx + p_t ; (t < N) du = K_t dx + k_t
    for (int row = warp; row < n + (t < N ? m : 0); row += nwarp) {
      const float* src = (row < n) ? gP + size_t(t) * n * n + size_t(row) * n : gK + size_t(t) * m * n + size_t(row - n) * n;
      float s = 0.f;
      for (int j = lane; j < n; j += 32) s = fmaf(src[j], dx[j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        if (row < n) gdV[size_t(t) * n + row] = s + gp[size_t(t) * n + row];
        else du[row - n] = s + gk[size_t(t) * m + row - n];
      }
    }
    for (int i = tid; i < n; i += NT) gdX[size_t(t) * n + i] = dx[i];
    __syncthreads();
    if (warp == 0) {
      for (int i = lane; i < n; i += 32) slope_acc = fmaf(gq[size_t(t) * n + i], dx[i], slope_acc);
      if (t < N)
        for (int i = lane; i < m; i += 32) slope_acc = fmaf(gr[size_t(t) * m + i], du[i], slope_acc);
    }
    if (t == N) break;
    for (int i = tid; i < m; i += NT) gdU[size_t(t) * m + i] = du[i];
    // dx_next = A_t dx + B_t du + c_{t+1}
    for (int row = warp; row < n; row += nwarp) {
      const float* ar = gA + size_t(t) * n * n + size_t(row) * n;
      const float* br = gB + size_t(t) * n * m + size_t(row) * m;
      float s = 0.f;
      for (int j = lane; j < n; j += 32) s = fmaf(ar[j], dx[j], s);
      for (int j = lane; j < m; j += 32) s = fmaf(br[j], du[j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) dxn[row] = s + gc[size_t(t + 1) * n + row];
    }
    __syncthreads();
    for (int i = tid; i < n; i += NT) dx[i] = dxn[i];
    __syncthreads();
  }
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) slope_acc += __shfl_xor_sync(0xffffffffu, slope_acc, o);
    if (lane == 0) a.slope[e] = slope_acc;
  }
}

template <int n, int m, int NT>
__global__ void __launch_bounds__(NT) lq_seq_kernel(LqArgs a) {
  using L = LqSmemLayout<n, m>;
  constexpr int LN = L::LN, LM = L::LM;
  extern __shared__ __align__(16) float smem[];
  float* sP = smem + L::oP;
  float* sA = smem + L::oA;
  float* sT = smem + L::oT;
  float* sB = smem + L::oB;
  float* sPB = smem + L::oPB;
  float* sH = smem + L::oH;
  float* sG = smem + L::oG;
  float* sK = smem + L::oK;
  float* sp = smem + L::oVec;  // p   [LN]
  float* sw = sp + LN;         // w = p + P c
  float* sc = sw + LN;         // c_t
  float* sq = sc + LN;         // scratch n-vector
  float* sh = sq + LN;         // h [LM]
  float* skk = sh + LM;        // k [LM]
  float* sr = skk + LM;        // r_t [LM]
  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int N = a.N;
  const float* gA = a.A + size_t(e) * N * n * n;
  const float* gB = a.B + size_t(e) * N * n * m;
  const float* gQ = a.Q + size_t(e) * (N + 1) * n * n;
  const float* gR = a.R + size_t(e) * N * m * m;
  const float* gM = a.M + size_t(e) * N * n * m;
  const float* gq = a.q + size_t(e) * (N + 1) * n;
  const float* gr = a.r + size_t(e) * N * m;
  const float* gc = a.c + size_t(e) * (N + 1) * n;
  float* gK = a.K + size_t(e) * N * m * n;
  float* gk = a.k + size_t(e) * N * m;
  float* gP = a.P + size_t(e) * (N + 1) * n * n;
  float* gp = a.p + size_t(e) * (N + 1) * n;

  // terminal value function: P_N = Q_N, p_N = q_N   (primal_tvlqr.py:136)
  load_padded(sP, LN, LN, gQ + size_t(N) * n * n, n, n, tid, NT);
  for (int i = tid; i < LN; i += NT) sp[i] = (i < n) ? gq[size_t(N) * n + i] : 0.f;
  for (int idx = tid; idx < n * n; idx += NT) gP[size_t(N) * n * n + idx] = gQ[size_t(N) * n * n + idx];
  for (int i = tid; i < n; i += NT) gp[size_t(N) * n + i] = gq[size_t(N) * n + i];
  __syncthreads();

  for (int t = N - 1; t >= 0; --t) {
    load_padded(sA, LN, LN, gA + size_t(t) * n * n, n, n, tid, NT);
    load_padded(sB, LM, LN, gB + size_t(t) * n * m, n, m, tid, NT);
    for (int i = tid; i < LN; i += NT) sc[i] = (i < n) ? gc[size_t(t + 1) * n + i] : 0.f;
    for (int i = tid; i < LM; i += NT) sr[i] = (i < m) ? gr[size_t(t) * m + i] : 0.f;
    __syncthreads();
    // w = p + P c
    for (int i = tid; i < LN; i += NT) {
      float s = sp[i];
      for (int j = 0; j < n; ++j) s = fmaf(sP[i * LN + j], sc[j], s);
      sw[i] = s;
    }
    // T1 = P A,  PB = P B
    gemm_tiles<false, false>(sT, LN, sP, LN, sA, LN, LN / 4, LN / 4, n, tid, NT);
    gemm_tiles<false, false>(sPB, LM, sP, LN, sB, LM, LN / 4, LM / 4, n, tid, NT);
    __syncthreads();
    // X = A^T T1 -> sP (old P is dead);  H = B^T T1 ;  G = B^T PB ;  h = B^T w + r
    gemm_tiles<true, false>(sP, LN, sA, LN, sT, LN, LN / 4, LN / 4, n, tid, NT);
    gemm_tiles<true, false>(sH, LN, sB, LM, sT, LN, LM / 4, LN / 4, n, tid, NT);
    gemm_tiles<true, false>(sG, LM, sB, LM, sPB, LM, LM / 4, LM / 4, n, tid, NT);
    for (int i = tid; i < LM; i += NT) {
      float s = sr[i];
      for (int j = 0; j < n; ++j) s = fmaf(sB[j * LM + i], sw[j], s);
      sh[i] = s;
    }
    // q-scratch = q_t + A^T w
    for (int i = tid; i < LN; i += NT) {
      float s = (i < n) ? gq[size_t(t) * n + i] : 0.f;
      for (int j = 0; j < n; ++j) s = fmaf(sA[j * LN + i], sw[j], s);
      sq[i] = s;
    }
    __syncthreads();
    // H += M^T ; G = sym(R + G)
    for (int idx = tid; idx < m * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      sH[i * LN + j] += gM[size_t(t) * n * m + j * m + i];
    }
    for (int idx = tid; idx < m * m; idx += NT) {
      const int i = idx / m, j = idx - i * m;
      if (i <= j) {
        const float gij = gR[size_t(t) * m * m + i * m + j] + sG[i * LM + j];
        const float gji = gR[size_t(t) * m * m + j * m + i] + sG[j * LM + i];
        const float s = 0.5f * (gij + gji);
        // both halves are read only by this thread before being written
        sG[i * LM + j] = s;
        sG[j * LM + i] = s;
      }
    }
    __syncthreads();
    // Cholesky G = L L^T in place (lower), one warp; NaN on a non-SPD G propagates like jax cho_factor
    if (tid < 32) {
      for (int j = 0; j < m; ++j) {
        float d = sG[j * LM + j];
        for (int k2 = 0; k2 < j; ++k2) d = fmaf(-sG[j * LM + k2], sG[j * LM + k2], d);
        d = sqrtf(d);
        const float inv = 1.f / d;
        __syncwarp();
        if (tid == 0) sG[j * LM + j] = d;
        for (int i = j + 1 + tid; i < m; i += 32) {
          float s = sG[i * LM + j];
          for (int k2 = 0; k2 < j; ++k2) s = fmaf(-sG[i * LM + k2], sG[j * LM + k2], s);
          sG[i * LM + j] = s * inv;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // [K k] = -G^-1 [H h]: one thread per right-hand-side column
    for (int col = tid; col <= n; col += NT) {
      float y[m];
#pragma unroll 1
      for (int i = 0; i < m; ++i) {
        float s = -((col < n) ? sH[i * LN + col] : sh[i]);
        for (int k2 = 0; k2 < i; ++k2) s = fmaf(-sG[i * LM + k2], y[k2], s);
        y[i] = s / sG[i * LM + i];
      }
#pragma unroll 1
      for (int i = m - 1; i >= 0; --i) {
        float s = y[i];
        for (int k2 = i + 1; k2 < m; ++k2) s = fmaf(-sG[k2 * LM + i], y[k2], s);
        y[i] = s / sG[i * LM + i];
      }
      if (col < n) {
        for (int i = 0; i < m; ++i) sK[i * LN + col] = y[i];
      } else {
        for (int i = 0; i < m; ++i) skk[i] = y[i];
      }
    }
    // zero the padding of K (rows m..LM-1 and cols n..LN-1) once per step so K^T H tiles stay clean
    for (int idx = tid; idx < LM * LN; idx += NT) {
      const int i = idx / LN, j = idx - i * LN;
      if (i >= m || j >= n) sK[idx] = 0.f;
    }
    __syncthreads();
    // Y = K^T H -> sA (A is dead); then Y += Q + sym(X)
    gemm_tiles<true, false>(sA, LN, sK, LN, sH, LN, LN / 4, LN / 4, m, tid, NT);
    // p_new = q + A^T w + K^T h
    for (int i = tid; i < LN; i += NT) {
      float s = sq[i];
      for (int j = 0; j < m; ++j) s = fmaf(sK[j * LN + i], sh[j], s);
      sp[i] = (i < n) ? s : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      sA[i * LN + j] += gQ[size_t(t) * n * n + idx] + 0.5f * (sP[i * LN + j] + sP[j * LN + i]);
    }
    __syncthreads();
    // P = sym(Y)
    for (int idx = tid; idx < LN * LN; idx += NT) {
      const int i = idx / LN, j = idx - i * LN;
      sP[idx] = (i < n && j < n) ? 0.5f * (sA[i * LN + j] + sA[j * LN + i]) : 0.f;
    }
    // store gains
    for (int idx = tid; idx < m * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      gK[size_t(t) * m * n + idx] = sK[i * LN + j];
    }
    for (int i = tid; i < m; i += NT) gk[size_t(t) * m + i] = skk[i];
    for (int i = tid; i < n; i += NT) gp[size_t(t) * n + i] = sp[i];
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      gP[size_t(t) * n * n + idx] = sP[i * LN + j];
    }
  }
  __syncthreads();

  // ---- forward rollout + dual (primal_tvlqr.py:484-498, dual_tvlqr.py:19) ----
  lq_rollout_dual<n, m, NT>(a, e, tid, sw, sq, sh);
}

}  // namespace mpxb
