// lq_pscan.cuh -- the LQ backward pass as a PARALLEL associative scan over the horizon (reference tvlqr_gpu,
// primal_tvlqr.py:337-475: "O(log T) parallel time complexity"), for the small-batch / long-horizon regime where one
// CTA per environment leaves the GPU empty (single-instance controllers, the Talos horizon sweep).
//
// Same elements and the same associative product as lq_scan.cuh (primal_tvlqr.py:364-448); what differs is the
// schedule: a Hillis-Steele suffix scan, one CTA per (environment, knot), ceil(log2(N+1)) launches:
//     level d = 1, 2, 4, ...:   e'_t = e_t (x) e_{t+d}   if t + d <= N,   e_t otherwise,
// where e_t covers the knots [t, t + d) before the level and [t, t + 2d) after it.  An element whose range reaches
// the terminal knot has A^ = c^ = C^ = 0 (the terminal element is (0, 0, 0, q_N, Q_N)), so for it only (p^, P^) are
// computed and stored -- those are exactly the value-function terms the scan is run for.
// Combine of an earlier element l with a later one r (primal_tvlqr.py:364-394), with W = I + P^_r C^_l, Z' = W^-1
// (Gauss-Jordan with partial pivoting: the reference inverts with LU) and (I + C^_l P^_r)^-1 = Z'^T because P^ and
// C^ are symmetric:
//     P^ = A^_l' (Z' P^_r) A^_l + P^_l             p^ = A^_l' Z' (p^_r + P^_r c^_l) + p^_l
//     Y = A^_r Z'^T ;  A^ = Y A^_l ;  C^ = Y C^_l A^_r' + C^_r ;  c^ = Y (c^_l - C^_l p^_r) + c^_r
// Launches: elements (1) + levels + gains (1, one CTA per knot: getKs, primal_tvlqr.py:455-473) + rollout / dual (1).
#pragma once
#include "lq_scan.cuh"

namespace mpxb {

struct PscanBuf { float *A, *C, *P, *c, *p; };  // [cnt] x (n*n | n*n | n*n | n | n), cnt = E * (N + 1)

__host__ __device__ inline size_t pscan_buf_floats(size_t cnt, int n) { return cnt * (size_t(3) * n * n + 2 * n); }
__host__ __device__ inline PscanBuf pscan_buf(float* ws, int which, size_t cnt, int n) {
  float* b = ws + size_t(which) * pscan_buf_floats(cnt, n);
  const size_t nn = size_t(n) * n;
  return {b, b + cnt * nn, b + 2 * cnt * nn, b + 3 * cnt * nn, b + 3 * cnt * nn + cnt * n};
}
// environments up to which MPXB_LQ_SCAN runs the parallel schedule (above it every SM already has its own
// environments and the one-CTA-per-environment kernels win)
constexpr int PSCAN_MAX_ENVS = 64;
inline bool pscan_applies(int E) { return E <= PSCAN_MAX_ENVS; }
inline size_t lq_pscan_floats(int E, int N, int n) { return 2 * pscan_buf_floats(size_t(E) * (N + 1), n); }

template <int n, int m> struct PscanInitSmem {
  static constexpr int LN = lq_pad(n), LM = lq_pad(m);
  static constexpr int oW = 0;                 // product scratch   [LN][LN]
  static constexpr int oB = oW + LN * LN;      // B                 [LN][LM]
  static constexpr int oBR = oB + LN * LM;     // B R^-1            [LN][LM]
  static constexpr int oM = oBR + LN * LM;     // M                 [LN][LM]
  static constexpr int oMR = oM + LN * LM;     // M R^-1            [LN][LM]
  static constexpr int oR = oMR + LN * LM;     // R -> chol         [LM][LM]
  static constexpr int oRi = oR + LM * LM;     // R^-1              [LM][LM]
  static constexpr int oVec = oRi + LM * LM;   // r [LM]
  static constexpr int total = oVec + LM + 8;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// scan elements (primal_tvlqr.py:396-448): one CTA per (knot, environment)
template <int n, int m, int NT>
__global__ void __launch_bounds__(NT) pscan_init_kernel(LqArgs a, PscanBuf o) {
  using L = PscanInitSmem<n, m>;
  constexpr int LN = L::LN, LM = L::LM;
  extern __shared__ __align__(16) float smem[];
  float* sW = smem + L::oW;
  float* sB = smem + L::oB;
  float* sBR = smem + L::oBR;
  float* sM = smem + L::oM;
  float* sMR = smem + L::oMR;
  float* sR = smem + L::oR;
  float* sRi = smem + L::oRi;
  float* sr = smem + L::oVec;
  const int t = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int N = a.N;
  const size_t el = size_t(e) * (N + 1) + t;
  const float* gQ = a.Q + el * n * n;
  const float* gq = a.q + el * n;
  float* oP = o.P + el * n * n;
  float* op = o.p + el * n;
  if (t == N) {  // terminal element (0, 0, 0, q_N, Q_N): A^, c^, C^ are implied zeros and never read
    for (int i = tid; i < n * n; i += NT) oP[i] = gQ[i];
    for (int i = tid; i < n; i += NT) op[i] = gq[i];
    return;
  }
  const size_t kt = size_t(e) * N + t;
  const float* gA = a.A + kt * n * n;
  const float* gc = a.c + el * n + n;  // c[t + 1]: the offset of knot t
  load_padded(sB, LM, LN, a.B + kt * n * m, n, m, tid, NT);
  load_padded(sM, LM, LN, a.M + kt * n * m, n, m, tid, NT);
  load_padded(sR, LM, LM, a.R + kt * m * m, m, m, tid, NT);
  for (int i = tid; i < LM; i += NT) sr[i] = (i < m) ? a.r[kt * m + i] : 0.f;
  __syncthreads();
  // R^-1 by Cholesky (primal_tvlqr.py:396-401)
  if (tid < 32) {
    for (int j = 0; j < m; ++j) {
      float d = sR[j * LM + j];
      for (int k2 = 0; k2 < j; ++k2) d = fmaf(-sR[j * LM + k2], sR[j * LM + k2], d);
      d = sqrtf(d);
      const float inv = 1.f / d;
      __syncwarp();
      if (tid == 0) sR[j * LM + j] = d;
      for (int i = j + 1 + tid; i < m; i += 32) {
        float s = sR[i * LM + j];
        for (int k2 = 0; k2 < j; ++k2) s = fmaf(-sR[i * LM + k2], sR[j * LM + k2], s);
        sR[i * LM + j] = s * inv;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int col = tid; col < LM; col += NT) {  // R^-1 column by column: L L' x = e_col
    float y[m > 0 ? m : 1];
    if (col < m) {
#pragma unroll 1
      for (int i = 0; i < m; ++i) {
        float s = (i == col) ? 1.f : 0.f;
        for (int k2 = 0; k2 < i; ++k2) s = fmaf(-sR[i * LM + k2], y[k2], s);
        y[i] = s / sR[i * LM + i];
      }
#pragma unroll 1
      for (int i = m - 1; i >= 0; --i) {
        float s = y[i];
        for (int k2 = i + 1; k2 < m; ++k2) s = fmaf(-sR[k2 * LM + i], y[k2], s);
        y[i] = s / sR[i * LM + i];
      }
    }
    for (int i = 0; i < LM; ++i) sRi[i * LM + col] = (col < m && i < m) ? y[i] : 0.f;
  }
  __syncthreads();
  gemm_tiles<false, false>(sBR, LM, sB, LM, sRi, LM, LN / 4, LM / 4, m, tid, NT);
  gemm_tiles<false, false>(sMR, LM, sM, LM, sRi, LM, LN / 4, LM / 4, m, tid, NT);
  __syncthreads();
  // c^ = c - BR r ; p^ = q - MR r
  for (int i = tid; i < n; i += NT) {
    float s1 = gc[i], s2 = gq[i];
    for (int j = 0; j < m; ++j) { s1 = fmaf(-sBR[i * LM + j], sr[j], s1); s2 = fmaf(-sMR[i * LM + j], sr[j], s2); }
    o.c[el * n + i] = s1;
    op[i] = s2;
  }
  // A^ = A - BR M'
  gemm_tiles<false, true>(sW, LN, sBR, LM, sM, LM, LN / 4, LN / 4, m, tid, NT);
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    o.A[el * n * n + idx] = gA[idx] - sW[i * LN + j];
  }
  __syncthreads();
  // C^ = BR B'
  gemm_tiles<false, true>(sW, LN, sBR, LM, sB, LM, LN / 4, LN / 4, m, tid, NT);
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    o.C[el * n * n + idx] = sW[i * LN + j];
  }
  __syncthreads();
  // P^ = Q - MR M'
  gemm_tiles<false, true>(sW, LN, sMR, LM, sM, LM, LN / 4, LN / 4, m, tid, NT);
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    oP[idx] = gQ[idx] - sW[i * LN + j];
  }
}

template <int n> struct PscanCombineSmem {
  static constexpr int LN = lq_pad(n);
  static constexpr int oS = 0;                   // five [LN][LN] buffers
  static constexpr int oVec = 5 * LN * LN;       // v / y, u, p_r, c_l, GJ multipliers, scratch  [LN] each ; pivot
  static constexpr int total = oVec + 6 * LN + 8;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// one level of the suffix scan: dst_t = src_t (x) src_{t+d}
template <int n, int NT>
__global__ void __launch_bounds__(NT) pscan_combine_kernel(int N, int d, const int32_t* done, PscanBuf src, PscanBuf dst) {
  using L = PscanCombineSmem<n>;
  constexpr int LN = L::LN, TN = LN / 4;
  extern __shared__ __align__(16) float smem[];
  float* S0 = smem;
  float* S1 = S0 + LN * LN;
  float* S2 = S1 + LN * LN;
  float* S3 = S2 + LN * LN;
  float* S4 = S3 + LN * LN;
  float* sv = smem + L::oVec;  // p_r + P_r c_l  ->  y = Z' (...)
  float* su = sv + LN;         // c_l - C_l p_r
  float* spr = su + LN;        // p_r
  float* scl = spr + LN;       // c_l
  float* sf = scl + LN;        // GJ multipliers
  int* s_piv = reinterpret_cast<int*>(sf + LN + LN);
  const int t = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  if (done && done[e]) return;
  const size_t nn = size_t(n) * n;
  const size_t l = size_t(e) * (N + 1) + t;
  const int rt = t + d;
  const bool l_final = t + d > N;  // src_t already covers the terminal knot (then it has no partner either)
  if (rt > N) {                    // no partner: the element is carried over
    for (int i = tid; i < n * n; i += NT) dst.P[l * nn + i] = src.P[l * nn + i];
    for (int i = tid; i < n; i += NT) dst.p[l * n + i] = src.p[l * n + i];
    if (!l_final) {  // (unreachable: kept so the carry-over rule reads like the scan's definition)
      for (int i = tid; i < n * n; i += NT) { dst.A[l * nn + i] = src.A[l * nn + i]; dst.C[l * nn + i] = src.C[l * nn + i]; }
      for (int i = tid; i < n; i += NT) dst.c[l * n + i] = src.c[l * n + i];
    }
    return;
  }
  const size_t r = l + d;
  const bool r_final = rt + d > N;  // the partner reaches the terminal knot: A^_r = c^_r = C^_r = 0
  // ---- 1: P_r -> S0, C_l -> S1 ; v = p_r + P_r c_l ; u = c_l - C_l p_r ; W = I + P_r C_l -> S2
  load_padded(S0, LN, LN, src.P + r * nn, n, n, tid, NT);
  load_padded(S1, LN, LN, src.C + l * nn, n, n, tid, NT);
  for (int i = tid; i < LN; i += NT) {
    spr[i] = (i < n) ? src.p[r * n + i] : 0.f;
    scl[i] = (i < n) ? src.c[l * n + i] : 0.f;
  }
  __syncthreads();
  gemm_tiles<false, false>(S2, LN, S0, LN, S1, LN, TN, TN, n, tid, NT);
  for (int idx = tid; idx < 2 * LN; idx += NT) {
    const int i = idx >> 1;
    if (idx & 1) {
      float s = scl[i];
      for (int j = 0; j < n; ++j) s = fmaf(-S1[i * LN + j], spr[j], s);
      su[i] = s;
    } else {
      float s = spr[i];
      for (int j = 0; j < n; ++j) s = fmaf(S0[i * LN + j], scl[j], s);
      sv[i] = s;
    }
  }
  for (int idx = tid; idx < LN * LN; idx += NT) {
    const int i = idx / LN, j = idx - i * LN;
    S3[idx] = (i == j) ? 1.f : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < LN; i += NT) S2[i * LN + i] += 1.f;
  __syncthreads();
  // ---- 2: Z' = W^-1 -> S3 ; y = Z' v -> sv
  gauss_jordan<NT>(S2, LN, S3, LN, n, sv, n, sf, s_piv, tid);
  // ---- 3: X = Z' P_r -> S2
  gemm_tiles<false, false>(S2, LN, S3, LN, S0, LN, TN, TN, n, tid, NT);
  __syncthreads();
  // ---- 4: A_l -> S0 ; T = X A_l -> S4
  load_padded(S0, LN, LN, src.A + l * nn, n, n, tid, NT);
  __syncthreads();
  gemm_tiles<false, false>(S4, LN, S2, LN, S0, LN, TN, TN, n, tid, NT);
  __syncthreads();
  // ---- 5: P = A_l' T + P_l -> S2 -> global ; p = A_l' y + p_l
  gemm_tiles<true, false>(S2, LN, S0, LN, S4, LN, TN, TN, n, tid, NT);
  for (int i = tid; i < n; i += NT) {
    float s = src.p[l * n + i];
    for (int j = 0; j < n; ++j) s = fmaf(S0[j * LN + i], sv[j], s);
    dst.p[l * n + i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    dst.P[l * nn + idx] = S2[i * LN + j] + src.P[l * nn + idx];
  }
  if (r_final) return;  // the new element reaches the terminal knot as well
  __syncthreads();
  // ---- 6: A_r -> S2 ; Y = A_r Z'^T -> S4
  load_padded(S2, LN, LN, src.A + r * nn, n, n, tid, NT);
  __syncthreads();
  gemm_tiles<false, true>(S4, LN, S2, LN, S3, LN, TN, TN, n, tid, NT);
  __syncthreads();
  // ---- 7: A = Y A_l -> S3 -> global ; c = Y u + c_r
  gemm_tiles<false, false>(S3, LN, S4, LN, S0, LN, TN, TN, n, tid, NT);
  for (int i = tid; i < n; i += NT) {
    float s = src.c[r * n + i];
    for (int j = 0; j < n; ++j) s = fmaf(S4[i * LN + j], su[j], s);
    dst.c[l * n + i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    dst.A[l * nn + idx] = S3[i * LN + j];
  }
  // ---- 8: D = Y C_l -> S0 (A_l is dead)
  gemm_tiles<false, false>(S0, LN, S4, LN, S1, LN, TN, TN, n, tid, NT);
  __syncthreads();
  // ---- 9: C = D A_r' + C_r -> S3 -> global
  gemm_tiles<false, true>(S3, LN, S0, LN, S2, LN, TN, TN, n, tid, NT);
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    dst.C[l * nn + idx] = S3[i * LN + j] + src.C[r * nn + idx];
  }
}

template <int n, int m> struct PscanGainSmem {
  static constexpr int LN = lq_pad(n), LM = lq_pad(m);
  static constexpr int oP = 0;                  // P_{t+1}          [LN][LN]
  static constexpr int oA = oP + LN * LN;       // A_t              [LN][LN]
  static constexpr int oH = oA + LN * LN;       // H -> K           [LM][LN]
  static constexpr int oBtP = oH + LM * LN;     // B' P+            [LM][LN]
  static constexpr int oB = oBtP + LM * LN;     // B                [LN][LM]
  static constexpr int oM = oB + LN * LM;       // M                [LN][LM]
  static constexpr int oG = oM + LN * LM;       // G                [LM][LM]
  static constexpr int oR = oG + LM * LM;       // R                [LM][LM]
  static constexpr int oVec = oR + LM * LM;     // p+, c [LN] ; r, h, GJ multipliers [LM] ; pivot
  static constexpr int total = oVec + 2 * LN + 3 * LM + 8;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// value function out of the scan + gains (getKs, primal_tvlqr.py:455-473): one CTA per (knot, environment)
template <int n, int m, int NT>
__global__ void __launch_bounds__(NT) pscan_gains_kernel(LqArgs a, PscanBuf fin) {
  using L = PscanGainSmem<n, m>;
  constexpr int LN = L::LN, LM = L::LM;
  extern __shared__ __align__(16) float smem[];
  float* sP = smem + L::oP;
  float* sA = smem + L::oA;
  float* sH = smem + L::oH;
  float* sBtP = smem + L::oBtP;
  float* sB = smem + L::oB;
  float* sM = smem + L::oM;
  float* sG = smem + L::oG;
  float* sR = smem + L::oR;
  float* sp = smem + L::oVec;
  float* sc = sp + LN;
  float* sr = sc + LN;
  float* sh = sr + LM;
  float* sfm = sh + LM;
  int* s_piv = reinterpret_cast<int*>(sfm + LM);
  const int t = blockIdx.x, e = blockIdx.y, tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int N = a.N;
  const size_t nn = size_t(n) * n;
  const size_t el = size_t(e) * (N + 1) + t;
  for (int i = tid; i < n * n; i += NT) a.P[el * nn + i] = fin.P[el * nn + i];
  for (int i = tid; i < n; i += NT) a.p[el * n + i] = fin.p[el * n + i];
  if (t == N) return;
  const size_t kt = size_t(e) * N + t;
  load_padded(sP, LN, LN, fin.P + (el + 1) * nn, n, n, tid, NT);
  load_padded(sA, LN, LN, a.A + kt * nn, n, n, tid, NT);
  load_padded(sB, LM, LN, a.B + kt * n * m, n, m, tid, NT);
  load_padded(sM, LM, LN, a.M + kt * n * m, n, m, tid, NT);
  load_padded(sR, LM, LM, a.R + kt * m * m, m, m, tid, NT);
  for (int i = tid; i < LN; i += NT) {
    sp[i] = (i < n) ? fin.p[(el + 1) * n + i] : 0.f;
    sc[i] = (i < n) ? a.c[(el + 1) * n + i] : 0.f;
  }
  for (int i = tid; i < LM; i += NT) sr[i] = (i < m) ? a.r[kt * m + i] : 0.f;
  __syncthreads();
  // BtP = B' P+ ; G = R + BtP B ; H = BtP A + M' ; h = B' p+ + BtP c + r
  gemm_tiles<true, false>(sBtP, LN, sB, LM, sP, LN, LM / 4, LN / 4, n, tid, NT);
  __syncthreads();
  gemm_tiles<false, false>(sG, LM, sBtP, LN, sB, LM, LM / 4, LM / 4, n, tid, NT);
  gemm_tiles<false, false>(sH, LN, sBtP, LN, sA, LN, LM / 4, LN / 4, n, tid, NT);
  for (int i = tid; i < LM; i += NT) {
    float s = sr[i];
    for (int j = 0; j < n; ++j) s = fmaf(sB[j * LM + i], sp[j], s);
    for (int j = 0; j < n; ++j) s = fmaf(sBtP[i * LN + j], sc[j], s);
    sh[i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < LM * LM; idx += NT) {
    const int i = idx / LM, j = idx - i * LM;
    sG[idx] = (i < m && j < m) ? sG[idx] + sR[idx] : (i == j ? 1.f : 0.f);
  }
  for (int idx = tid; idx < LM * LN; idx += NT) {
    const int i = idx / LN, j = idx - i * LN;
    sH[idx] = (i < m && j < n) ? -(sH[idx] + sM[j * LM + i]) : 0.f;  // -H
  }
  for (int i = tid; i < LM; i += NT) sh[i] = (i < m) ? -sh[i] : 0.f;
  __syncthreads();
  gauss_jordan<NT>(sG, LM, sH, LN, n, sh, m, sfm, s_piv, tid);  // sH rows 0..m-1 = K, sh = k
  for (int idx = tid; idx < m * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    a.K[kt * m * n + idx] = sH[i * LN + j];
  }
  for (int i = tid; i < m; i += NT) a.k[kt * m + i] = sh[i];
}

// forward rollout + dual update + slope (primal_tvlqr.py:484-498 evaluated left to right, dual_tvlqr.py:19)
template <int n, int m, int NT>
__global__ void __launch_bounds__(NT) pscan_rollout_kernel(LqArgs a, int arena) {
  extern __shared__ __align__(16) float smem[];
  const int e = blockIdx.x, tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int nzc = (a.nzc > 0 && a.nzc < n) ? a.nzc : n;
  const int LZ = (nzc + 3) & ~3;
  const int acols = (LZ < n) ? LZ : n;
  lq_rollout_stream<n, m, NT>(a, e, tid, smem, acols, arena);
}

template <int n, int m> cudaError_t launch_lq_pscan(const LqArgs& a, float* ws, cudaStream_t st) {
  constexpr int NT = 256, NTC = 512;
  auto k_init = pscan_init_kernel<n, m, NT>;
  auto k_comb = pscan_combine_kernel<n, NTC>;
  auto k_gain = pscan_gains_kernel<n, m, NT>;
  auto k_roll = pscan_rollout_kernel<n, m, NT>;
  const size_t roll_bytes = lq_rt_bytes<n, m>(a.nzc);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t err = cudaFuncSetAttribute(k_init, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanInitSmem<n, m>::bytes));
    if (err == cudaSuccess) err = cudaFuncSetAttribute(k_comb, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanCombineSmem<n>::bytes));
    if (err == cudaSuccess) err = cudaFuncSetAttribute(k_gain, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanGainSmem<n, m>::bytes));
    if (err == cudaSuccess) err = cudaFuncSetAttribute(k_roll, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lq_rt_bytes<n, m>(0)));
    if (err != cudaSuccess) return err;
    configured[dev] = true;
  }
  const size_t cnt = size_t(a.E) * (a.N + 1);
  PscanBuf buf[2] = {pscan_buf(ws, 0, cnt, n), pscan_buf(ws, 1, cnt, n)};
  const dim3 grid(a.N + 1, a.E);
  k_init<<<grid, NT, PscanInitSmem<n, m>::bytes, st>>>(a, buf[0]);
  int cur = 0;
  for (int d = 1; d <= a.N; d <<= 1) {
    k_comb<<<grid, NTC, PscanCombineSmem<n>::bytes, st>>>(a.N, d, a.done, buf[cur], buf[cur ^ 1]);
    cur ^= 1;
  }
  k_gain<<<grid, NT, PscanGainSmem<n, m>::bytes, st>>>(a, buf[cur]);
  k_roll<<<a.E, NT, roll_bytes, st>>>(a, int(roll_bytes / sizeof(float)));
  return cudaGetLastError();
}

inline int dispatch_lq_pscan(int n, int m, const LqArgs& a, float* ws, cudaStream_t st) {
  cudaError_t err;
  if (!ws) return MPXB_EINVAL;
#define MPXB_LQP_CASE(NN, MM) if (n == NN && m == MM) { err = launch_lq_pscan<NN, MM>(a, ws, st); return err == cudaSuccess ? MPXB_OK : MPXB_ECUDA; }
  MPXB_LQP_CASE(13, 12)
  MPXB_LQP_CASE(61, 12)
  MPXB_LQP_CASE(75, 19)
  MPXB_LQP_CASE(81, 46)
  MPXB_LQP_CASE(2, 1)
  MPXB_LQP_CASE(4, 1)
  MPXB_LQP_CASE(8, 2)
#undef MPXB_LQP_CASE
  return MPXB_EUNSUPPORTED;
}

}  // namespace mpxb
