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

// Gauss-Jordan with partial pivoting for the scan kernels: same elimination as gauss_jordan() in lq_scan.cuh (row
// exchange with the largest |pivot|, every other row reduced against the scaled pivot row), laid out for throughput:
// the scaled pivot row is staged in `prow` once, and each thread then updates 4 x 4 register patches of [W | Rhs]
// with float4 accesses (no integer division, no per-element index arithmetic), three barriers per pivot plus one
// when rows are exchanged.  ldw, ldr multiples of 4 and >= the padded row count; W, Rhs, prow 16-byte aligned;
// prow holds ldw + ldr floats.  Columns of W left of the pivot are dead (never read again) and may be overwritten.
template <int NT>
__device__ __forceinline__ void gauss_jordan_tiled(float* W, int ldw, float* Rhs, int ldr, float* v, int k, float* fcol,
                                                   float* prow, int* s_piv, int tid) {
  const int RT = (k + 3) >> 2;       // row tiles
  const int CR = ldr >> 2;           // column chunks of Rhs
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
    const int c0 = (c + 1) & ~3;       // first live column chunk of W (aligned down: may include dead columns)
    if (pv != c) {
      for (int j = c + tid; j < ldw + ldr + 1; j += NT) {
        float *x, *y;
        if (j < ldw) { x = W + c * ldw + j; y = W + pv * ldw + j; }
        else if (j < ldw + ldr) { x = Rhs + c * ldr + (j - ldw); y = Rhs + pv * ldr + (j - ldw); }
        else { x = v + c; y = v + pv; }
        const float t = *x; *x = *y; *y = t;
      }
      __syncthreads();
    }
    {  // multipliers and the scaled pivot row
      const float inv = 1.f / W[c * ldw + c];
      for (int j = tid; j < ldw + ldr + 1 + 4 * RT; j += NT) {
        if (j < ldw) prow[j] = W[c * ldw + j] * inv;
        else if (j < ldw + ldr) prow[j] = Rhs[c * ldr + (j - ldw)] * inv;
        else if (j == ldw + ldr) fcol[4 * RT] = v[c] * inv;
        else { const int i = j - ldw - ldr - 1; fcol[i] = (i < k && i != c) ? W[i * ldw + c] : 0.f; }
      }
    }
    __syncthreads();
    const int CW = (ldw - c0) >> 2;    // live column chunks of W
    const int CT = CW + CR;
    for (int patch = tid; patch < RT * CT; patch += NT) {
      const int rt = patch / CT, ct = patch - rt * CT;
      float* base;
      int ld;
      const float* pr;
      if (ct < CW) { base = W + c0 + 4 * ct; ld = ldw; pr = prow + c0 + 4 * ct; }
      else { base = Rhs + 4 * (ct - CW); ld = ldr; pr = prow + ldw + 4 * (ct - CW); }
      const float4 pvv = *reinterpret_cast<const float4*>(pr);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * rt + a;
        float4* cell = reinterpret_cast<float4*>(base + i * ld);
        if (i == c) {
          *cell = pvv;
        } else {
          const float f = fcol[i];
          if (f != 0.f || !(f == f)) {
            float4 x = *cell;
            x.x = fmaf(-f, pvv.x, x.x); x.y = fmaf(-f, pvv.y, x.y); x.z = fmaf(-f, pvv.z, x.z); x.w = fmaf(-f, pvv.w, x.w);
            *cell = x;
          }
        }
      }
    }
    if (tid < k) {
      const float vc = fcol[4 * RT];
      v[tid] = (tid == c) ? vc : fmaf(-fcol[tid], vc, v[tid]);
    }
    __syncthreads();
  }
}

// Gauss-Jordan with partial pivoting, OUT OF PLACE: pivot step c reads [W | Rhs | v] from buffer c & 1 and writes the
// whole updated system to the other buffer, so nothing is staged and nothing is exchanged in place -- the row swap is
// a renaming of the source rows -- and a pivot costs ONE barrier (gauss_jordan_tiled: three or four, each a full
// shared-memory round trip, which is what an 81-pivot elimination is made of).  Every warp finds the pivot itself
// (three loads and a shuffle butterfly on (|value| bits, ~row) packed in 64 bits; ties to the smallest row, NaN wins
// and propagates), every thread scales its own chunk of the pivot row.  Same arithmetic as gauss_jordan_tiled.
// Buffers: W[2] [.][ldw], Rhs[2] [.][ldr], v[2]; ldw, ldr multiples of 4, rows padded to a multiple of 4 (padded rows
// carry zeros in the pivot columns and are copied through); the result is in buffer k & 1.
template <int NT>
__device__ __forceinline__ void gauss_jordan_pp(float* W0, float* W1, int ldw, float* R0, float* R1, int ldr, float* v0,
                                                float* v1, int k, int tid) {
  const int RT = (k + 3) >> 2;
  const int CR = ldr >> 2;
  const int lane = tid & 31;
  for (int c = 0; c < k; ++c) {
    const float* Wo = (c & 1) ? W1 : W0;
    float* Wn = (c & 1) ? W0 : W1;
    const float* Ro = (c & 1) ? R1 : R0;
    float* Rn = (c & 1) ? R0 : R1;
    const float* vo = (c & 1) ? v1 : v0;
    float* vn = (c & 1) ? v0 : v1;
    unsigned long long best = 0ull;
    for (int i = c + lane; i < k; i += 32) {
      const unsigned long long pk = (static_cast<unsigned long long>(__float_as_uint(fabsf(Wo[i * ldw + c]))) << 32) |
                                    static_cast<unsigned>(0x7fffffff - i);
      best = pk > best ? pk : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      best = ob > best ? ob : best;
    }
    const int pv = 0x7fffffff - static_cast<int>(static_cast<unsigned>(best & 0xffffffffull));
    const float inv = 1.f / Wo[pv * ldw + c];
    const int c0 = (c + 1) & ~3;       // first live column chunk of W (aligned down: may include dead columns)
    const int CW = (ldw - c0) >> 2;
    const int CT = CW + CR;
    for (int patch = tid; patch < RT * CT; patch += NT) {
      const int rt = patch / CT, ct = patch - rt * CT;
      const float* ob;
      float* nb;
      int ld;
      if (ct < CW) { ob = Wo + c0 + 4 * ct; nb = Wn + c0 + 4 * ct; ld = ldw; }
      else { ob = Ro + 4 * (ct - CW); nb = Rn + 4 * (ct - CW); ld = ldr; }
      float4 pr = *reinterpret_cast<const float4*>(ob + pv * ld);
      pr.x *= inv; pr.y *= inv; pr.z *= inv; pr.w *= inv;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * rt + a;
        float4 x;
        if (i == c) {
          x = pr;
        } else {
          const int src = (i == pv) ? c : i;     // the exchanged row
          const float f = Wo[src * ldw + c];
          x = *reinterpret_cast<const float4*>(ob + src * ld);
          x.x = fmaf(-f, pr.x, x.x); x.y = fmaf(-f, pr.y, x.y); x.z = fmaf(-f, pr.z, x.z); x.w = fmaf(-f, pr.w, x.w);
        }
        *reinterpret_cast<float4*>(nb + i * ld) = x;
      }
    }
    if (tid < 4 * RT) {
      const float vc = vo[pv] * inv;
      const int src = (tid == pv) ? c : tid;
      vn[tid] = (tid == c) ? vc : fmaf(-Wo[src * ldw + c], vc, vo[src]);
    }
    __syncthreads();
  }
}

// Gauss-Jordan with partial pivoting, REGISTER RESIDENT: every thread keeps its 4 x 4 patches of [W | Rhs | v] in
// registers for the whole elimination (two patches per thread at the shipped sizes); per pivot the owners of column c
// publish it (fcol), every warp finds the pivot among the rows not used yet, the owners of the pivot row publish it
// (prow), and each thread updates its own registers: ~40 instructions and two barriers per pivot where the
// shared-memory forms (gauss_jordan_tiled / _pp) spend ~250 instructions per warp on addressing and staging.  Rows are
// never exchanged: the pivot row stays where it is, is marked used, and the rows are written back permuted at the end
// (same pivots, same arithmetic as the exchange form).  fcol / prow are double-buffered by the parity of the pivot.
// W [LDW rows][LDW], R [LDW rows][LDR] (in: right-hand sides, out: W^-1 R), v [LDW] (in / out); k <= LDW pivots;
// rows >= k must carry zeros in the pivot columns.  scratch: 2 * (LDW + LDW + LDR + 4) floats + 2 * LDW ints.
template <int NT, int LDW, int LDR>
__device__ __forceinline__ void gauss_jordan_reg(float* W, float* R, float* v, int k, float* scratch, int tid) {
  constexpr int RT = LDW / 4, CW = LDW / 4, CR = LDR / 4, CT = CW + CR + 1, NP = RT * CT, PPT = (NP + NT - 1) / NT;
  constexpr int PROW = LDW + LDR + 4;
  float* fcol = scratch;                    // [2][LDW]
  float* prow = scratch + 2 * LDW;          // [2][PROW]
  int* colof = reinterpret_cast<int*>(prow + 2 * PROW);   // [LDW]: physical row -> the pivot column it served
  const int lane = tid & 31;
  float4 reg[PPT][4];
  int prt[PPT], pct[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * NT;
    prt[q] = p < NP ? p / CT : -1;
    pct[q] = p < NP ? p - (p / CT) * CT : -1;
    if (prt[q] >= 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = 4 * prt[q] + a;
        if (pct[q] < CW) reg[q][a] = *reinterpret_cast<const float4*>(W + i * LDW + 4 * pct[q]);
        else if (pct[q] < CW + CR) reg[q][a] = *reinterpret_cast<const float4*>(R + i * LDR + 4 * (pct[q] - CW));
        else reg[q][a] = make_float4(v[i], 0.f, 0.f, 0.f);
      }
    }
  }
  for (int i = tid; i < LDW; i += NT) colof[i] = i;
  unsigned usedmask = 0u;   // bit u: row lane + 32 u has served as a pivot (kept identically by every warp)
  static_assert(LDW <= 32 * 32, "used-row mask");
  for (int c = 0; c < k; ++c) {
    float* fc = fcol + (c & 1) * LDW;
    float* pw = prow + (c & 1) * PROW;
    // column c, from its owners
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (pct[q] == (c >> 2)) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 x = reg[q][a];
          const int w = c & 3;
          fc[4 * prt[q] + a] = w == 0 ? x.x : (w == 1 ? x.y : (w == 2 ? x.z : x.w));
        }
      }
    __syncthreads();
    // pivot = arg max |column c| over the rows not used yet, ties to the smallest row, NaN wins (it must propagate):
    // every warp on its own -- each lane reduces its rows (lane, lane + 32, ...; their used flags live in a register
    // mask), then two hardware warp reductions (max of the value bits, min of the rows that attain it)
    unsigned bbits = 0u;
    int brow = 0x7fffffff;
#pragma unroll
    for (int u = 0; u < (LDW + 31) / 32; ++u) {
      const int i = lane + 32 * u;
      if (i < k && !((usedmask >> u) & 1u)) {
        const unsigned b = __float_as_uint(fabsf(fc[i])) + 1u;   // + 1: a candidate always beats "none" (0); order kept
        if (b > bbits) { bbits = b; brow = i; }
      }
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, bbits);
    const int pv = static_cast<int>(__reduce_min_sync(0xffffffffu, bbits == wmax ? static_cast<unsigned>(brow) : 0x7fffffffu));
    if ((pv & 31) == lane) usedmask |= 1u << (pv >> 5);
    const float inv = 1.f / fc[pv];
    // the pivot row, from its owners
#pragma unroll
    for (int q = 0; q < PPT; ++q)
      if (prt[q] == (pv >> 2)) {
        const int a = pv & 3;
        const float4 x = a == 0 ? reg[q][0] : (a == 1 ? reg[q][1] : (a == 2 ? reg[q][2] : reg[q][3]));
        *reinterpret_cast<float4*>(pw + 4 * pct[q]) = x;
      }
    if (tid == 0) colof[pv] = c;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (prt[q] < 0) continue;
      if (pct[q] < CW && 4 * pct[q] + 3 < c) continue;   // columns of W left of the pivot are dead
      float4 pr = *reinterpret_cast<const float4*>(pw + 4 * pct[q]);
      pr.x *= inv; pr.y *= inv; pr.z *= inv; pr.w *= inv;
      const float4 f4 = *reinterpret_cast<const float4*>(fc + 4 * prt[q]);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float f = a == 0 ? f4.x : (a == 1 ? f4.y : (a == 2 ? f4.z : f4.w));
        const bool piv = 4 * prt[q] + a == pv;
        float4 x = reg[q][a];
        x.x = piv ? pr.x : fmaf(-f, pr.x, x.x); x.y = piv ? pr.y : fmaf(-f, pr.y, x.y);
        x.z = piv ? pr.z : fmaf(-f, pr.z, x.z); x.w = piv ? pr.w : fmaf(-f, pr.w, x.w);
        reg[q][a] = x;
      }
    }
  }
  __syncthreads();
  // write back W^-1 R and W^-1 v, rows in pivot order
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    if (prt[q] < 0 || pct[q] < CW) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = colof[4 * prt[q] + a];
      if (pct[q] < CW + CR) *reinterpret_cast<float4*>(R + row * LDR + 4 * (pct[q] - CW)) = reg[q][a];
      else v[row] = reg[q][a].x;
    }
  }
  __syncthreads();
}

// C[MT*4, NT*4] = A * B with A [.][lda] and B [.][ldb] row-major, K a multiple of 4 (zero padded): 4 x 4 register tiles,
// both operands read as float4 (A along k -- a broadcast for the threads that share a tile row --, B along j), i.e.
// 8 LDS.128 per 64 FMA where gemm_tiles<false, false> issues 16 scalar + 4 vector loads.
__device__ __forceinline__ void gemm_nn4(float* __restrict__ C, int ldc, const float* __restrict__ A, int lda,
                                         const float* __restrict__ B, int ldb, int MT, int NT, int K, int tid, int nthreads) {
  for (int tile = tid; tile < MT * NT; tile += nthreads) {
    const int i0 = (tile / NT) * 4, j0 = (tile % NT) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 2
    for (int kk = 0; kk < K; kk += 4) {
      float av[4][4], bv[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float4 t = *reinterpret_cast<const float4*>(A + (i0 + a) * lda + kk);
        av[a][0] = t.x; av[a][1] = t.y; av[a][2] = t.z; av[a][3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(B + (kk + q) * ldb + j0);
        bv[q][0] = t.x; bv[q][1] = t.y; bv[q][2] = t.z; bv[q][3] = t.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a][q], bv[q][b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      *reinterpret_cast<float4*>(C + (i0 + a) * ldc + j0) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
  }
}

// load a [n, n] row-major global matrix TRANSPOSED into a zero-padded [LN][LN] shared tile
__device__ __forceinline__ void load_padded_t(float* __restrict__ dst, int ld, const float* __restrict__ src, int n, int tid,
                                              int nthreads) {
  for (int idx = tid; idx < ld * ld; idx += nthreads) {
    const int j = idx / ld, i = idx - j * ld;        // consecutive threads read consecutive global addresses src[j][i]
    if (j < n && i < n) dst[i * ld + j] = src[j * n + i];
    else dst[idx] = 0.f;                             // (i, j) -> idx is a bijection on the padding as well
  }
}

template <int n, int m> struct PscanInitSmem {
  static constexpr int LN = lq_pad(n), LM = lq_pad(m);
  static constexpr int oW = 0;                 // product scratch   [LN][LN]
  static constexpr int oB = oW + LN * LN;      // B                 [LN][LM]
  static constexpr int oBR = oB + LN * LM;     // B R^-1            [LN][LM]
  static constexpr int oM = oBR + LN * LM;     // M                 [LN][LM]
  static constexpr int oMR = oM + LN * LM;     // M R^-1            [LN][LM]
  static constexpr int oR = oMR + LN * LM;     // R                 [2][LM][LM]
  static constexpr int oRi = oR + 2 * LM * LM; // I -> R^-1         [2][LM][LM]  (R, R^-1: ping-pong pairs of the elimination)
  static constexpr int oVec = oRi + 2 * LM * LM;  // r [LM] ; dummy right-hand-side vector [LM] ; elimination scratch [8 LM + 8]
  static constexpr int total = oVec + 10 * LM + 16;
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
  float* sdum = sr + LM;
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
  for (int i = tid; i < LM; i += NT) { sr[i] = (i < m) ? a.r[kt * m + i] : 0.f; sdum[i] = 0.f; sdum[LM + i] = 0.f; }
  for (int idx = tid; idx < LM * LM; idx += NT) sRi[idx] = (idx / LM == idx % LM && idx / LM < m) ? 1.f : 0.f;
  __syncthreads();
  // R^-1 (primal_tvlqr.py:396-401 factorises R by Cholesky and solves against I; here the same inverse comes out of
  // the Gauss-Jordan elimination, m pivots of a few hundred cycles instead of a one-warp factorisation)
  gauss_jordan_reg<NT, LM, LM>(sR, sRi, sdum, m, sdum + LM, tid);
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
  static constexpr int oS = 0;                   // six [LN][LN] buffers
  static constexpr int oVec = 6 * LN * LN;       // v / y, u, p_r, c_l [LN] each ; elimination scratch [8 LN + 8]
  static constexpr int total = oVec + 12 * LN + 16;
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
  float* S5 = S4 + LN * LN;
  float* sv = smem + L::oVec;  // p_r + P_r c_l  ->  y = Z' (...)
  float* su = sv + LN;         // c_l - C_l p_r
  float* spr = su + LN;        // p_r
  float* scl = spr + LN;       // c_l
  float* sscr = scl + LN;      // elimination scratch
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
  gemm_nn4(S2, LN, S0, LN, S1, LN, TN, TN, LN, tid, NT);
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
  // ---- 2: Z' = W^-1 -> S3 ; y = Z' v -> sv   (register-resident elimination)
  gauss_jordan_reg<NT, LN, LN>(S2, S3, sv, n, sscr, tid);
  // ---- 3: X = Z' P_r -> S2
  gemm_nn4(S2, LN, S3, LN, S0, LN, TN, TN, LN, tid, NT);
  __syncthreads();
  // ---- 4: A_l -> S0 ; T = X A_l -> S4
  load_padded(S0, LN, LN, src.A + l * nn, n, n, tid, NT);
  __syncthreads();
  gemm_nn4(S4, LN, S2, LN, S0, LN, TN, TN, LN, tid, NT);
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
  // ---- 6: A_r' -> S2 ; Y' = Z' A_r' -> S4   (Y = A_r Z'^T kept transposed: every product below reads rows)
  load_padded_t(S2, LN, src.A + r * nn, n, tid, NT);
  __syncthreads();
  gemm_nn4(S4, LN, S3, LN, S2, LN, TN, TN, LN, tid, NT);
  __syncthreads();
  // ---- 7: A = Y A_l -> S5 -> global ; c = Y u + c_r
  gemm_tiles<true, false>(S5, LN, S4, LN, S0, LN, TN, TN, n, tid, NT);
  for (int i = tid; i < n; i += NT) {
    float s = src.c[r * n + i];
    for (int j = 0; j < n; ++j) s = fmaf(S4[j * LN + i], su[j], s);
    dst.c[l * n + i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    dst.A[l * nn + idx] = S5[i * LN + j];
  }
  // ---- 8: D = Y C_l -> S0 (A_l is dead)
  gemm_tiles<true, false>(S0, LN, S4, LN, S1, LN, TN, TN, n, tid, NT);
  __syncthreads();
  // ---- 9: C = D A_r' + C_r -> S3 -> global
  gemm_nn4(S3, LN, S0, LN, S2, LN, TN, TN, LN, tid, NT);
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
  static constexpr int oH = oA + LN * LN;       // H -> K           [2][LM][LN]  (G, H, h: ping-pong pairs of the elimination)
  static constexpr int oBtP = oH + 2 * LM * LN; // B' P+            [LM][LN]
  static constexpr int oB = oBtP + LM * LN;     // B                [LN][LM]
  static constexpr int oM = oB + LN * LM;       // M                [LN][LM]
  static constexpr int oG = oM + LN * LM;       // G                [2][LM][LM]
  static constexpr int oR = oG + 2 * LM * LM;   // R                [LM][LM]
  static constexpr int oVec = oR + LM * LM;     // p+, c [LN] ; r, h [LM] ; elimination scratch [6 LM + 2 LN + 8]
  static constexpr int total = oVec + 4 * LN + 8 * LM + 16;
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
  gauss_jordan_reg<NT, LM, LN>(sG, sH, sh, m, sh + LM, tid);               // sH rows 0..m-1 = K, sh = k
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

// > 48 KB of dynamic shared memory is a per-device function attribute: set at model creation (mpxb_model_create) and by
// the stand-alone mpxb_tvlqr, never inside a launch sequence that may be under stream capture
template <int n, int m> cudaError_t configure_lq_pscan() {
  constexpr int NT = 256, NTC = 512;
  cudaError_t err = cudaFuncSetAttribute(pscan_init_kernel<n, m, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanInitSmem<n, m>::bytes));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(pscan_combine_kernel<n, NTC>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanCombineSmem<n>::bytes));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(pscan_gains_kernel<n, m, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PscanGainSmem<n, m>::bytes));
  if (err == cudaSuccess) err = cudaFuncSetAttribute(pscan_rollout_kernel<n, m, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lq_rt_bytes<n, m>(0)));
  return err;
}

template <int n, int m> cudaError_t launch_lq_pscan(const LqArgs& a, float* ws, cudaStream_t st) {
  // threads per CTA: 256 / 512 also for the small sizes (two-warp CTAs measured slower for SRBD, tick 0.26 vs 0.23 ms:
  // the kernels are bound by load and barrier latency, which more warps overlap)
  constexpr int NT = 256, NTC = 512;
  static_assert(NT >= n || n > 256, "rollout / elimination index rows by thread");
  static_assert(NTC >= n && NT >= m, "gauss_jordan_tiled updates the right-hand-side vector one row per thread");
  auto k_init = pscan_init_kernel<n, m, NT>;
  auto k_comb = pscan_combine_kernel<n, NTC>;
  auto k_gain = pscan_gains_kernel<n, m, NT>;
  auto k_roll = pscan_rollout_kernel<n, m, NT>;
  const size_t roll_bytes = lq_rt_bytes<n, m>(a.nzc);
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
