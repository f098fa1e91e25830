// lq_scan.cuh -- the associative-scan form of the LQ backward pass (reference tvlqr_gpu, primal_tvlqr.py:337-475)
// with the LU-based gain solve (getKs, primal_tvlqr.py:455-473), followed by the rollout / dual update.
//
// Scan elements (primal_tvlqr.py:396-448):  e_t = (A^, c^, C^, p^, P^) = (A - B R^-1 M', c - B R^-1 r, B R^-1 B',
// q - M R^-1 r, Q - M R^-1 M'),  e_N = (0, 0, 0, q_N, Q_N).  Combine of an earlier element l with a later one r
// (primal_tvlqr.py:364-394):
//     Z' = (I + P^_r C^_l)^-1 ;  p^ = A^_l' Z' (p^_r + P^_r c^_l) + p^_l ;  P^ = A^_l' Z' P^_r A^_l + P^_l
//     (and A^ = A^_r Z A^_l, c^ = A^_r Z (c^_l - C^_l p^_r) + c^_r, C^ = A^_r Z C^_l A^_r' + C^_r).
// The suffix scan is evaluated here in the right-to-left schedule, one CTA per environment: because the terminal
// element has A^ = c^ = C^ = 0, every suffix product has too, so each combine only updates (p^, P^) -- one n x n
// solve by Gauss-Jordan with partial pivoting (the reference inverts I + P C with LU, np.linalg.inv) and two GEMMs.
// It is the same associative product as lax.associative_scan's tree schedule and differs from it in rounding only
// (the float64 oracle agrees to 1e-8: tests/test_oracle_lq.py::test_scan_orders_and_sequential_agree).
// Gains: [K k] = -(R + B'P+B)^-1 [B'P+A + M', B'p+ + B'P+c + r] by LU with partial pivoting, no symmetrisation.
#pragma once
#include "lq.cuh"

namespace mpxb {

template <int n, int m> struct LqScanSmem {
  static constexpr int LN = lq_pad(n), LM = lq_pad(m);
  static constexpr int oP = 0;               // P+ -> X = Z' P+          [LN][LN]
  static constexpr int oC = oP + LN * LN;    // C^ -> T = X A^           [LN][LN]
  static constexpr int oW = oC + LN * LN;    // W = I + P+ C^            [LN][LN]
  static constexpr int oA = oW + LN * LN;    // A -> A^                  [LN][LN]
  static constexpr int oB = oA + LN * LN;    // B                        [LN][LM]
  static constexpr int oBR = oB + LN * LM;   // B R^-1                   [LN][LM]
  static constexpr int oM = oBR + LN * LM;   // M                        [LN][LM]
  static constexpr int oMR = oM + LN * LM;   // M R^-1                   [LN][LM]
  static constexpr int oBtP = oMR + LN * LM; // B' P+  /  H              [LM][LN]
  static constexpr int oR = oBtP + LM * LN;  // R -> chol                [LM][LM]
  static constexpr int oRi = oR + LM * LM;   // R^-1                     [LM][LM]
  static constexpr int oG = oRi + LM * LM;   // G                        [LM][LM]
  static constexpr int oVec = oG + LM * LM;  // p, v, c, c^, q^, x, f(LN each) ; r, h, (LM each) ; misc
  static constexpr int total = oVec + 8 * LN + 4 * LM + 16;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// Gauss-Jordan elimination with partial pivoting on the k x k system W (ld ldw), applied to `nrhs` right-hand-side
// columns stored row-major in Rhs (ld ldr) and to one extra vector v.  On return Rhs = W^-1 Rhs, v = W^-1 v.
// fcol: [k] shared scratch.  All threads of the CTA must call.
template <int NT>
__device__ __forceinline__ void gauss_jordan(float* W, int ldw, float* Rhs, int ldr, int nrhs, float* v, int k,
                                             float* fcol, int* s_piv, int tid) {
  for (int c = 0; c < k; ++c) {
    if (tid < 32) {
      float best = -1.f;
      int bi = c;
      for (int i = c + tid; i < k; i += 32) {
        const float a = fabsf(W[i * ldw + c]);
        if (a > best || !(a == a)) { best = (a == a) ? a : 3.4e38f; bi = i; }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) *s_piv = bi;
    }
    __syncthreads();
    const int pv = *s_piv;
    if (pv != c) {
      for (int j = tid; j < k + nrhs + 1; j += NT) {
        float* a;
        float* b;
        if (j < k) { a = W + c * ldw + j; b = W + pv * ldw + j; }
        else if (j < k + nrhs) { a = Rhs + c * ldr + (j - k); b = Rhs + pv * ldr + (j - k); }
        else { a = v + c; b = v + pv; }
        const float t = *a; *a = *b; *b = t;
      }
      __syncthreads();
    }
    const float inv = 1.f / W[c * ldw + c];
    for (int i = tid; i < k; i += NT) fcol[i] = (i == c) ? 0.f : W[i * ldw + c] * inv;
    __syncthreads();
    // eliminate column c from every other row; columns <= c of W are already final and never read again
    const int wc = k - c - 1, tot = wc + nrhs + 1;
    for (int idx = tid; idx < k * tot; idx += NT) {
      const int i = idx / tot, j = idx - i * tot;
      const float f = fcol[i];
      if (f != 0.f || !(f == f)) {
        if (j < wc) W[i * ldw + c + 1 + j] = fmaf(-f, W[c * ldw + c + 1 + j], W[i * ldw + c + 1 + j]);
        else if (j < wc + nrhs) Rhs[i * ldr + (j - wc)] = fmaf(-f, Rhs[c * ldr + (j - wc)], Rhs[i * ldr + (j - wc)]);
        else v[i] = fmaf(-f, v[c], v[i]);
      }
    }
    __syncthreads();
    for (int j = tid; j < tot; j += NT) {
      if (j < wc) W[c * ldw + c + 1 + j] *= inv;
      else if (j < wc + nrhs) Rhs[c * ldr + (j - wc)] *= inv;
      else v[c] *= inv;
    }
    __syncthreads();
  }
}

template <int n, int m, int NT>
__global__ void __launch_bounds__(NT) lq_scan_kernel(LqArgs a) {
  using L = LqScanSmem<n, m>;
  constexpr int LN = L::LN, LM = L::LM;
  extern __shared__ __align__(16) float smem[];
  float* sP = smem + L::oP;
  float* sC = smem + L::oC;
  float* sW = smem + L::oW;
  float* sA = smem + L::oA;
  float* sB = smem + L::oB;
  float* sBR = smem + L::oBR;
  float* sM = smem + L::oM;
  float* sMR = smem + L::oMR;
  float* sBtP = smem + L::oBtP;
  float* sR = smem + L::oR;
  float* sRi = smem + L::oRi;
  float* sG = smem + L::oG;
  float* sp = smem + L::oVec;  // p+            [LN]
  float* sv = sp + LN;         // p+ + P+ c^    [LN]
  float* sc = sv + LN;         // c_t           [LN]
  float* sch = sc + LN;        // c^            [LN]
  float* sqh = sch + LN;       // q^            [LN]
  float* sx = sqh + LN;        // scratch       [LN]
  float* sf = sx + LN;         // GJ multipliers[LN]
  float* sy = sf + LN;         // scratch       [LN]
  float* sr = sy + LN;         // r_t           [LM]
  float* sh = sr + LM;         // h / k         [LM]
  float* st = sh + LM;         // R^-1 r        [LM]
  float* sfm = st + LM;        // GJ multipliers[LM]
  int* s_piv = reinterpret_cast<int*>(sfm + LM);
  const int e = blockIdx.x, tid = threadIdx.x;
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

  // terminal element: (0, 0, 0, q_N, Q_N)
  load_padded(sP, LN, LN, gQ + size_t(N) * n * n, n, n, tid, NT);
  for (int i = tid; i < LN; i += NT) sp[i] = (i < n) ? gq[size_t(N) * n + i] : 0.f;
  for (int idx = tid; idx < n * n; idx += NT) gP[size_t(N) * n * n + idx] = gQ[size_t(N) * n * n + idx];
  for (int i = tid; i < n; i += NT) gp[size_t(N) * n + i] = gq[size_t(N) * n + i];
  __syncthreads();

  for (int t = N - 1; t >= 0; --t) {
    load_padded(sA, LN, LN, gA + size_t(t) * n * n, n, n, tid, NT);
    load_padded(sB, LM, LN, gB + size_t(t) * n * m, n, m, tid, NT);
    load_padded(sM, LM, LN, gM + size_t(t) * n * m, n, m, tid, NT);
    load_padded(sR, LM, LM, gR + size_t(t) * m * m, m, m, tid, NT);
    for (int i = tid; i < LN; i += NT) sc[i] = (i < n) ? gc[size_t(t + 1) * n + i] : 0.f;
    for (int i = tid; i < LM; i += NT) sr[i] = (i < m) ? gr[size_t(t) * m + i] : 0.f;
    __syncthreads();

    // ---- gains from (P+, p+):  BtP = B' P+ ; G = R + BtP B ; H = BtP A + M' ; h = B' p+ + BtP c + r
    gemm_tiles<true, false>(sBtP, LN, sB, LM, sP, LN, LM / 4, LN / 4, n, tid, NT);
    __syncthreads();
    gemm_tiles<false, false>(sG, LM, sBtP, LN, sB, LM, LM / 4, LM / 4, n, tid, NT);
    gemm_tiles<false, false>(sW, LN, sBtP, LN, sA, LN, LM / 4, LN / 4, n, tid, NT);  // H (before M') in sW rows 0..LM-1
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
      sW[idx] = (i < m && j < n) ? -(sW[idx] + sM[j * LM + i]) : 0.f;   // -H
    }
    for (int i = tid; i < LM; i += NT) sh[i] = (i < m) ? -sh[i] : 0.f;
    __syncthreads();
    gauss_jordan<NT>(sG, LM, sW, LN, n, sh, m, sfm, s_piv, tid);       // sW rows 0..m-1 = K, sh = k
    for (int idx = tid; idx < m * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      gK[size_t(t) * m * n + idx] = sW[i * LN + j];
    }
    for (int i = tid; i < m; i += NT) gk[size_t(t) * m + i] = sh[i];
    __syncthreads();

    // ---- scan element of knot t:  R^-1 by Cholesky (primal_tvlqr.py:396-401)
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
    for (int col = tid; col < LM; col += NT) {   // R^-1 column by column: L L' x = e_col
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
    // C^ = BR B' ; A^ = A - BR M' ; P^_l = Q - MR M' (kept in global Q, added at the end) ; c^ = c - BR r ; q^ = q - MR r
    gemm_tiles<false, true>(sC, LN, sBR, LM, sB, LM, LN / 4, LN / 4, m, tid, NT);
    gemm_tiles<false, true>(sW, LN, sBR, LM, sM, LM, LN / 4, LN / 4, m, tid, NT);   // BR M'
    for (int i = tid; i < LN; i += NT) {
      float s1 = sc[i], s2 = (i < n) ? gq[size_t(t) * n + i] : 0.f;
      for (int j = 0; j < m; ++j) { s1 = fmaf(-sBR[i * LM + j], sr[j], s1); s2 = fmaf(-sMR[i * LM + j], sr[j], s2); }
      sch[i] = s1;
      sqh[i] = s2;
    }
    __syncthreads();
    for (int idx = tid; idx < LN * LN; idx += NT) sA[idx] -= sW[idx];
    __syncthreads();
    // W = I + P+ C^ ; v = p+ + P+ c^
    gemm_tiles<false, false>(sW, LN, sP, LN, sC, LN, LN / 4, LN / 4, n, tid, NT);
    for (int i = tid; i < LN; i += NT) {
      float s = sp[i];
      for (int j = 0; j < n; ++j) s = fmaf(sP[i * LN + j], sch[j], s);
      sv[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < LN; i += NT) sW[i * LN + i] += 1.f;
    __syncthreads();
    // [X | x] = W^-1 [P+ | v]   (Z' P+ and Z' (p+ + P+ c^))
    gauss_jordan<NT>(sW, LN, sP, LN, n, sv, n, sf, s_piv, tid);
    // T = X A^ -> sC ; then P = A^' T + P^_l, p = A^' x + q^
    gemm_tiles<false, false>(sC, LN, sP, LN, sA, LN, LN / 4, LN / 4, n, tid, NT);
    __syncthreads();
    gemm_tiles<true, false>(sP, LN, sA, LN, sC, LN, LN / 4, LN / 4, n, tid, NT);
    gemm_tiles<false, true>(sW, LN, sMR, LM, sM, LM, LN / 4, LN / 4, m, tid, NT);   // MR M'
    for (int i = tid; i < LN; i += NT) {
      float s = sqh[i];
      for (int j = 0; j < n; ++j) s = fmaf(sA[j * LN + i], sv[j], s);
      sx[i] = (i < n) ? s : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < LN * LN; idx += NT) {
      const int i = idx / LN, j = idx - i * LN;
      const float v = (i < n && j < n) ? sP[idx] + gQ[size_t(t) * n * n + i * n + j] - sW[idx] : 0.f;
      sP[idx] = v;
      if (i < n && j < n) gP[size_t(t) * n * n + i * n + j] = v;
    }
    for (int i = tid; i < LN; i += NT) {
      sp[i] = sx[i];
      if (i < n) gp[size_t(t) * n + i] = sx[i];
    }
    __syncthreads();
  }
  __syncthreads();
  lq_rollout_dual<n, m, NT>(a, e, tid, sv, sx, sh);
}

template <int n, int m> cudaError_t configure_lq_scan() {
  return cudaFuncSetAttribute(lq_scan_kernel<n, m, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(LqScanSmem<n, m>::bytes));
}
template <int n, int m> cudaError_t launch_lq_scan(const LqArgs& a, cudaStream_t st) {
  constexpr int NT = 256;
  auto kern = lq_scan_kernel<n, m, NT>;
  kern<<<a.E, NT, LqScanSmem<n, m>::bytes, st>>>(a);
  return cudaGetLastError();
}

inline size_t lq_scan_floats(int, int, int, int) { return 0; }  // everything lives in shared memory

inline int dispatch_lq_scan(int n, int m, const LqArgs& a, float*, cudaStream_t st) {
  cudaError_t err;
#define MPXB_LQS_CASE(NN, MM) if (n == NN && m == MM) { err = launch_lq_scan<NN, MM>(a, st); return err == cudaSuccess ? MPXB_OK : MPXB_ECUDA; }
  MPXB_LQS_CASE(13, 12)
  MPXB_LQS_CASE(61, 12)
  MPXB_LQS_CASE(75, 19)
  MPXB_LQS_CASE(81, 46)
  MPXB_LQS_CASE(2, 1)
  MPXB_LQS_CASE(4, 1)
  MPXB_LQS_CASE(8, 2)
#undef MPXB_LQS_CASE
  return MPXB_EUNSUPPORTED;
}

}  // namespace mpxb
