// lq3.cuh -- the LQ sub-problem for STRUCTURED whole-body problems: sequential Riccati backward pass, forward rollout and
// dual update (reference primal_tvlqr.py:11-52 lqr_step, 103-144 tvlqr, 478-498 rollout; dual_tvlqr.py:6-19 dual_lqr)
// in the form the structure of the shipped whole-body problems admits, with Blackwell data movement.
//
// Structure (holds for every shipped whole-body config; the host enables this path only after checking it,
// MPXB_OPT_COST_STRUCTURE):
//   * the state is x = [head | tail], head = [p, quat, q, v, w, dq] (nh = 13 + 2 nj), tail = [feet, GRFs]; the tail
//     never enters the dynamics (models.py:61,95,109-115), so columns >= nh of A_t are zero;
//   * W is diagonal (config_*.py: block_diag of diag(...)), so the Gauss-Newton blocks (objectives.py:127-202,
//     252-323, 378-450) are  Q_t = diag + a 4x4 quaternion block (+ 3x3 friction-cone blocks on the GRF tail of the
//     quadrupeds),  R_t = diag,  M_t = 0.
// Then  P_t = blockdiag(P_hh dense nh x nh,  Q_t,tail)  for every t,  p_t,tail = q_t,tail,  K_t has zero tail columns,
// and lqr_step reduces to
//   w = p+ + P+ c,  T1 = P+ Az (head rows: dense P_hh A_hh; tail rows: sparse Q_tail A_th),  PB = P+ B,
//   X = Az' T1,  H = B' T1,  G = sym(R + B' PB),  h = B' w + r,  [K k] = -G^-1 [H h],
//   P_hh = sym(Q_hh + X + K' H),  p_h = q_h + Az' w + K' h
// which is the reference recursion with the structural zeros never multiplied: the same numbers in the same order of
// summation, at 0.4x the FLOPs, a third of the shared memory and without Q / M / the tail of P ever touching HBM.
//
// Data movement: every per-knot block lives in HBM in exactly the layout the kernel wants in shared memory (rows padded
// to 16 bytes), so one elected thread moves it with ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA path:
// UBLKCP in SASS) that completes on an mbarrier (transaction count); results leave shared memory the same way
// (cp.async.bulk.global.shared::cta + bulk groups).  No thread computes a global address inside the knot loop.
//
// Compute: 64 threads per CTA (one CTA per environment), 8x8 / 8x4 / 4x8 register tiles on k-major shared-memory
// operands read as float4: 4 FMA per shared-memory float loaded, which is the balance point of the SM
// (128 FMA lanes against 32 shared-memory words per clock) -- the 4x4 tiles of lq.cuh are shared-memory bound by 2x.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace mpxb {

// Dimensions and packed layouts shared by the linearise kernels (writers), this kernel and the expand kernel.
template <int NH_, int NTL_, int M_, int NCONE_> struct Lq3Dims {
  static constexpr int NH = NH_, NTL = NTL_, M = M_, NCONE = NCONE_, n = NH_ + NTL_;
  static constexpr int LH = (NH_ + 7) & ~7;    // padded head: row stride of A, P_hh, K, H
  static constexpr int LMB = (M_ + 3) & ~3;    // padded control dim: row stride of B, PB, G
  static constexpr int LN = (n + 3) & ~3;
  // packed knot block PK[t] (floats): diag(Q_t) | q_t | c_t | quaternion block of Q_t (4x4) | cone blocks (NCONE x 3x3)
  // | diag(R_t) | r_t
  static constexpr int oQd = 0, oq = LN, oc = 2 * LN, oQq = 3 * LN, oQc = oQq + 16;
  static constexpr int oRd = oQc + ((9 * NCONE_ + 3) & ~3), oR = oRd + LMB, PKS = oR + LMB;
  static constexpr int OKS = LH + LMB;         // packed result block OK[t]: p_t head | k_t
  static constexpr int cone0 = n - 3 * NCONE_;  // first state index of the cone blocks (the GRF tail of a quadruped)
  static constexpr int szA = n * LH, szB = n * LMB, szP = LH * LH, szK = LMB * LH;  // floats per knot
};

// Runtime mirror of Lq3Dims (host side: workspace layout, views)
struct Lq3Layout {
  int nh, ntl, m, ncone, n, LH, LMB, LN, oQd, oq, oc, oQq, oQc, oRd, oR, PKS, OKS, cone0, szA, szB, szP, szK;
};
inline Lq3Layout lq3_layout(int nh, int ntl, int m, int ncone) {
  Lq3Layout L;
  L.nh = nh; L.ntl = ntl; L.m = m; L.ncone = ncone; L.n = nh + ntl;
  L.LH = (nh + 7) & ~7; L.LMB = (m + 3) & ~3; L.LN = (L.n + 3) & ~3;
  L.oQd = 0; L.oq = L.LN; L.oc = 2 * L.LN; L.oQq = 3 * L.LN; L.oQc = L.oQq + 16;
  L.oRd = L.oQc + ((9 * ncone + 3) & ~3); L.oR = L.oRd + L.LMB; L.PKS = L.oR + L.LMB;
  L.OKS = L.LH + L.LMB; L.cone0 = L.n - 3 * ncone;
  L.szA = L.n * L.LH; L.szB = L.n * L.LMB; L.szP = L.LH * L.LH; L.szK = L.LMB * L.LH;
  return L;
}

struct Lq3Args {
  int E, N;
  const float *A;    // [E][N][n][LH]     leading nh columns live, the rest zero
  const float *B;    // [E][N][n][LMB]
  const float *PK;   // [E][N+1][PKS]
  float *Kc;         // [E][N][LMB][LH]   rows >= m and columns >= nh are zero
  float *Pc;         // [E][N+1][LH][LH]  head block of P_t
  float *OK;         // [E][N+1][OKS]     p_t head | k_t
  float *dX, *dU, *dV, *slope;   // dense [E][N+1][n], [E][N][m], [E][N+1][n], [E]
  const int32_t* done;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy, tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ register tiles
// acc[TI][TJ] += A[0:K, i0:i0+TI]^T B[0:K, j0:j0+TJ]   (both operands k-major in shared memory, float4 loads)
template <int TI, int TJ>
__device__ __forceinline__ void rt_atb(float (&acc)[TI][TJ], const float* __restrict__ A, int lda, int i0,
                                       const float* __restrict__ B, int ldb, int j0, int K) {
  const float* pa = A + i0;
  const float* pb = B + j0;
#pragma unroll 2
  for (int kk = 0; kk < K; ++kk) {
    float av[TI], bv[TJ];
#pragma unroll
    for (int x = 0; x < TI; x += 4) {
      const float4 t = *reinterpret_cast<const float4*>(pa + x);
      av[x] = t.x; av[x + 1] = t.y; av[x + 2] = t.z; av[x + 3] = t.w;
    }
#pragma unroll
    for (int y = 0; y < TJ; y += 4) {
      const float4 t = *reinterpret_cast<const float4*>(pb + y);
      bv[y] = t.x; bv[y + 1] = t.y; bv[y + 2] = t.z; bv[y + 3] = t.w;
    }
#pragma unroll
    for (int x = 0; x < TI; ++x)
#pragma unroll
      for (int y = 0; y < TJ; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    pa += lda;
    pb += ldb;
  }
}
template <int TI, int TJ> __device__ __forceinline__ void rt_zero(float (&acc)[TI][TJ]) {
#pragma unroll
  for (int x = 0; x < TI; ++x)
#pragma unroll
    for (int y = 0; y < TJ; ++y) acc[x][y] = 0.f;
}
template <int TI, int TJ>
__device__ __forceinline__ void rt_store(float* __restrict__ C, int ldc, int i0, int j0, const float (&acc)[TI][TJ]) {
#pragma unroll
  for (int x = 0; x < TI; ++x)
#pragma unroll
    for (int y = 0; y < TJ; y += 4)
      *reinterpret_cast<float4*>(C + (i0 + x) * ldc + j0 + y) = make_float4(acc[x][y], acc[x][y + 1], acc[x][y + 2], acc[x][y + 3]);
}
// rows >= row_end are not stored (the head tiles of T1 / PB overlap the first tail rows, which are written elsewhere)
template <int TI, int TJ>
__device__ __forceinline__ void rt_store_rows(float* __restrict__ C, int ldc, int i0, int j0, const float (&acc)[TI][TJ], int row_end) {
#pragma unroll
  for (int x = 0; x < TI; ++x)
    if (i0 + x < row_end) {
#pragma unroll
      for (int y = 0; y < TJ; y += 4)
        *reinterpret_cast<float4*>(C + (i0 + x) * ldc + j0 + y) = make_float4(acc[x][y], acc[x][y + 1], acc[x][y + 2], acc[x][y + 3]);
    }
}
template <int TI, int TJ>
__device__ __forceinline__ void rt_load(float (&acc)[TI][TJ], const float* __restrict__ C, int ldc, int i0, int j0) {
#pragma unroll
  for (int x = 0; x < TI; ++x)
#pragma unroll
    for (int y = 0; y < TJ; y += 4) {
      const float4 t = *reinterpret_cast<const float4*>(C + (i0 + x) * ldc + j0 + y);
      acc[x][y] = t.x; acc[x][y + 1] = t.y; acc[x][y + 2] = t.z; acc[x][y + 3] = t.w;
    }
}

// Shared-memory arena (floats).  The rollout reuses it as two stages of {K, A, B, P, PK, OK}.
template <class D> struct Lq3Smem {
  static constexpr int LH = D::LH, LMB = D::LMB, LN = D::LN, n = D::n;
  static constexpr int TROWS = (n + 3) & ~3;            // rows of T1 / PB (k extent of the second products)
  static constexpr int oP0 = 0;                         // P ping  [LH][LH]
  static constexpr int oP1 = oP0 + LH * LH;             // P pong
  static constexpr int oA = oP1 + LH * LH;              // A_t  [n][LH]
  static constexpr int oT = oA + D::szA;                // T1   [TROWS][LH]
  static constexpr int oB = oT + TROWS * LH;            // B_t  [n][LMB]
  static constexpr int oPB = oB + D::szB;               // PB   [TROWS][LMB]
  static constexpr int oH = oPB + TROWS * LMB;          // H    [LMB][LH]
  static constexpr int oG = oH + LMB * LH;              // G    [LMB][LMB]
  static constexpr int oK = oG + LMB * LMB;             // K    [LMB][LH]
  static constexpr int oPK0 = oK + LMB * LH;            // PK ping / pong
  static constexpr int oPK1 = oPK0 + D::PKS;
  static constexpr int oOK = oPK1 + D::PKS;             // OK staging: p head | k
  static constexpr int oVec = oOK + D::OKS;             // p [LN] | w [LN] | qh [LH] | h [LMB] | gi [LMB]
  static constexpr int backward = oVec + 2 * LN + LH + 2 * LMB;
  static constexpr int stage = D::szK + D::szA + D::szB + D::szP + D::PKS + D::OKS;
  static constexpr int rollout = 2 * stage + 2 * LN + LMB + 8;
  static constexpr int total = (backward > rollout ? backward : rollout) + 8;   // + mbarriers (4 x 8 bytes)
  static constexpr size_t bytes = size_t(total) * sizeof(float);
};

// resident CTAs per SM the shared-memory footprint allows (capped at 6)
template <class D> __host__ __device__ constexpr int lq3_min_blocks() {
  constexpr int b = int(232448 / (Lq3Smem<D>::bytes + 1024));
  return b < 1 ? 1 : (b > 6 ? 6 : b);
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NH, int NTL, int M, int NCONE, int NT>
__global__ void __launch_bounds__(NT, lq3_min_blocks<Lq3Dims<NH, NTL, M, NCONE>>()) lq3_kernel(Lq3Args a) {
  using D = Lq3Dims<NH, NTL, M, NCONE>;
  using S = Lq3Smem<D>;
  constexpr int LH = D::LH, LMB = D::LMB, LN = D::LN, n = D::n, m = M, TROWS = S::TROWS;
  constexpr int TH = LH / 8, TM4 = LMB / 4;
  extern __shared__ __align__(128) float smem3[];
  float* const smem = smem3;
  const int e = blockIdx.x, tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int N = a.N;
  float* sPb[2] = {smem + S::oP0, smem + S::oP1};
  float* sA = smem + S::oA;
  float* sT = smem + S::oT;
  float* sB = smem + S::oB;
  float* sPB = smem + S::oPB;
  float* sH = smem + S::oH;
  float* sG = smem + S::oG;
  float* sK = smem + S::oK;
  float* sPKb[2] = {smem + S::oPK0, smem + S::oPK1};
  float* sOK = smem + S::oOK;
  float* sp = smem + S::oVec;   // p_{t+1} (full n)
  float* sw = sp + LN;          // w = p + P c
  float* sqh = sw + LN;         // q_h + Az' w   [LH]
  float* sh = sqh + LH;         // h [LMB]
  float* sgi = sh + LMB;        // 1 / diag(chol G) [LMB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::total - 8);   // [0] backward loads, [1..2] rollout stages

  const float* gA = a.A + size_t(e) * N * D::szA;
  const float* gB = a.B + size_t(e) * N * D::szB;
  const float* gPK = a.PK + size_t(e) * (N + 1) * D::PKS;
  float* gK = a.Kc + size_t(e) * N * D::szK;
  float* gP = a.Pc + size_t(e) * (N + 1) * D::szP;
  float* gOK = a.OK + size_t(e) * (N + 1) * D::OKS;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
  }
  __syncthreads();
  // first loads: A_{N-1}, B_{N-1}, PK[N] (terminal) and PK[N-1]
  if (tid == 0) {
    mbar_expect_tx(&bars[0], uint32_t((D::szA + D::szB + 2 * D::PKS) * sizeof(float)));
    bulk_g2s(sA, gA + size_t(N - 1) * D::szA, D::szA * sizeof(float), &bars[0]);
    bulk_g2s(sB, gB + size_t(N - 1) * D::szB, D::szB * sizeof(float), &bars[0]);
    bulk_g2s(sPKb[0], gPK + size_t(N) * D::PKS, D::PKS * sizeof(float), &bars[0]);
    bulk_g2s(sPKb[1], gPK + size_t(N - 1) * D::PKS, D::PKS * sizeof(float), &bars[0]);
  }
  // zero what no copy ever writes: the pad rows of T1 / PB (k extent TROWS > n), H / K / G pad rows, vectors
  for (int i = tid; i < (TROWS - n) * LH; i += NT) sT[n * LH + i] = 0.f;
  for (int i = tid; i < (TROWS - n) * LMB; i += NT) sPB[n * LMB + i] = 0.f;
  for (int i = tid; i < LMB * LH; i += NT) { sH[i] = 0.f; sK[i] = 0.f; }
  for (int i = tid; i < LMB * LMB; i += NT) sG[i] = 0.f;
  for (int i = tid; i < 2 * LN + LH + 2 * LMB; i += NT) sp[i] = 0.f;
  for (int i = tid; i < D::OKS; i += NT) sOK[i] = 0.f;
  uint32_t ph0 = 0;
  mbar_wait(&bars[0], ph0);
  ph0 ^= 1;
  // terminal value function: P_N = Q_N, p_N = q_N (primal_tvlqr.py:136); head block expanded to dense
  int cur = 0;
  {
    const float* pkN = sPKb[0];
    float* P = sPb[0];
    for (int i = tid; i < LH * LH; i += NT) {
      const int r = i / LH, c = i - r * LH;
      float v = 0.f;
      if (r == c && r < NH) v = pkN[D::oQd + r];
      if (r >= 3 && r < 7 && c >= 3 && c < 7) v = pkN[D::oQq + (r - 3) * 4 + (c - 3)];
      P[i] = v;
    }
    for (int i = tid; i < n; i += NT) sp[i] = pkN[D::oq + i];
    for (int i = tid; i < LH; i += NT) sOK[i] = (i < NH) ? pkN[D::oq + i] : 0.f;
  }
  __syncthreads();
  fence_async_smem();
  if (tid == 0) {
    bulk_s2g(gP + size_t(N) * D::szP, sPb[0], D::szP * sizeof(float));
    bulk_s2g(gOK + size_t(N) * D::OKS, sOK, D::OKS * sizeof(float));
    bulk_commit();
  }
  int pkn = 0;  // buffer that holds PK[t+1]; PK[t] is in the other one

  for (int t = N - 1; t >= 0; --t) {
    const float* P = sPb[cur];        // P_{t+1} head (symmetric)
    float* X = sPb[cur ^ 1];          // X, XY, then P_t
    const float* pk1 = sPKb[pkn];      // PK[t+1]: c_{t+1}, tail of Q_{t+1} (= tail of P_{t+1})
    const float* pk0 = sPKb[pkn ^ 1];  // PK[t]:   Q_t, q_t, R_t, r_t
    // ---- S1: T1 = P+ Az, PB = P+ B (head rows by tiles, tail rows sparse), w = p+ + P+ c
    for (int tile = tid; tile < TH * TH + TH * TM4; tile += NT) {
      if (tile < TH * TH) {
        const int i0 = (tile / TH) * 8, j0 = (tile % TH) * 8;
        float acc[8][8];
        rt_zero(acc);
        rt_atb<8, 8>(acc, P, LH, i0, sA, LH, j0, NH);
        rt_store_rows<8, 8>(sT, LH, i0, j0, acc, NH);
      } else {
        const int tt = tile - TH * TH;
        const int i0 = (tt / TM4) * 8, j0 = (tt % TM4) * 4;
        float acc[8][4];
        rt_zero(acc);
        rt_atb<8, 4>(acc, P, LH, i0, sB, LMB, j0, NH);
        rt_store_rows<8, 4>(sPB, LMB, i0, j0, acc, NH);
      }
    }
    // tail rows r = NH .. n-1 of T1 / PB: P_tt = diag(Qd) (+ 3x3 cone blocks), rows of A / B below the head
    for (int idx = tid; idx < NTL * (LH + LMB) + n; idx += NT) {
      if (idx < NTL * (LH + LMB)) {
        const int rr = idx / (LH + LMB), cc = idx - rr * (LH + LMB);
        const int r = NH + rr;
        const bool isA = cc < LH;
        const float* src = isA ? sA + cc : sB + (cc - LH);
        const int ld = isA ? LH : LMB;
        float v = pk1[D::oQd + r] * src[r * ld];
        if (NCONE > 0 && r >= D::cone0) {
          const int blk = (r - D::cone0) / 3, k0 = D::cone0 + 3 * blk, ri = r - k0;
          const float* cb = pk1 + D::oQc + 9 * blk + 3 * ri;
          v = fmaf(cb[0], src[k0 * ld], fmaf(cb[1], src[(k0 + 1) * ld], fmaf(cb[2], src[(k0 + 2) * ld], v)));
        }
        if (isA) sT[r * LH + cc] = v; else sPB[r * LMB + (cc - LH)] = v;
      } else {
        const int i = idx - NTL * (LH + LMB);
        const float* cv = pk1 + D::oc;
        float s = sp[i];
        if (i < NH) {
          for (int j = 0; j < NH; ++j) s = fmaf(P[j * LH + i], cv[j], s);
        } else {
          s = fmaf(pk1[D::oQd + i], cv[i], s);
          if (NCONE > 0 && i >= D::cone0) {
            const int blk = (i - D::cone0) / 3, k0 = D::cone0 + 3 * blk, ri = i - k0;
            const float* cb = pk1 + D::oQc + 9 * blk + 3 * ri;
            s = fmaf(cb[0], cv[k0], fmaf(cb[1], cv[k0 + 1], fmaf(cb[2], cv[k0 + 2], s)));
          }
        }
        sw[i] = s;
      }
    }
    __syncthreads();
    // ---- S2: X = Az' T1, H = B' T1, G = B' PB, h = B' w + r, qh = q_h + Az' w
    for (int tile = tid; tile < TH * TH + TM4 * TH + TM4 * TM4 + LMB + LH; tile += NT) {
      if (tile < TH * TH) {
        const int i0 = (tile / TH) * 8, j0 = (tile % TH) * 8;
        float acc[8][8];
        rt_zero(acc);
        rt_atb<8, 8>(acc, sA, LH, i0, sT, LH, j0, n);
        rt_store<8, 8>(X, LH, i0, j0, acc);
      } else if (tile < TH * TH + TM4 * TH) {
        const int tt = tile - TH * TH;
        const int i0 = (tt / TH) * 4, j0 = (tt % TH) * 8;
        float acc[4][8];
        rt_zero(acc);
        rt_atb<4, 8>(acc, sB, LMB, i0, sT, LH, j0, n);
        rt_store<4, 8>(sH, LH, i0, j0, acc);
      } else if (tile < TH * TH + TM4 * TH + TM4 * TM4) {
        const int tt = tile - TH * TH - TM4 * TH;
        const int i0 = (tt / TM4) * 4, j0 = (tt % TM4) * 4;
        float acc[4][4];
        rt_zero(acc);
        rt_atb<4, 4>(acc, sB, LMB, i0, sPB, LMB, j0, n);
        rt_store<4, 4>(sG, LMB, i0, j0, acc);
      } else if (tile < TH * TH + TM4 * TH + TM4 * TM4 + LMB) {
        const int i = tile - TH * TH - TM4 * TH - TM4 * TM4;
        float s = (i < m) ? pk0[D::oR + i] : 0.f;
        for (int j = 0; j < n; ++j) s = fmaf(sB[j * LMB + i], sw[j], s);
        sh[i] = (i < m) ? s : 0.f;
      } else {
        const int i = tile - TH * TH - TM4 * TH - TM4 * TM4 - LMB;
        float s = (i < NH) ? pk0[D::oq + i] : 0.f;
        for (int j = 0; j < n; ++j) s = fmaf(sA[j * LH + i], sw[j], s);
        sqh[i] = (i < NH) ? s : 0.f;
      }
    }
    __syncthreads();
    // A_t, B_t and PK[t+1] are dead: fetch knot t-1 while S3..S7 run.  The bulk stores of the previous knot have long
    // finished reading K / OK / P (they were issued a whole S1 + S2 ago); wait for them before those buffers are rewritten.
    if (tid == 0) {
      bulk_wait_read0();
      if (t > 0) {
        mbar_expect_tx(&bars[0], uint32_t((D::szA + D::szB + D::PKS) * sizeof(float)));
        bulk_g2s(sA, gA + size_t(t - 1) * D::szA, D::szA * sizeof(float), &bars[0]);
        bulk_g2s(sB, gB + size_t(t - 1) * D::szB, D::szB * sizeof(float), &bars[0]);
        bulk_g2s(sPKb[pkn], gPK + size_t(t - 1) * D::PKS, D::PKS * sizeof(float), &bars[0]);
      }
    }
    // ---- S3: G = sym(R + G), R = diag (M = 0: H needs no correction)
    for (int idx = tid; idx < m * m; idx += NT) {
      const int i = idx / m, j = idx - i * m;
      if (i <= j) {
        float gij = sG[i * LMB + j], gji = sG[j * LMB + i];
        if (i == j) { const float rd = pk0[D::oRd + i]; gij += rd; gji += rd; }
        const float s = 0.5f * (gij + gji);
        sG[i * LMB + j] = s;
        sG[j * LMB + i] = s;
      }
    }
    __syncthreads();
    // ---- S4: Cholesky G = L L' in place (lower), one warp; NaN on a non-SPD G propagates like jax cho_factor
    if (tid < 32) {
      for (int j = 0; j < m; ++j) {
        float d = sG[j * LMB + j];
        for (int k2 = 0; k2 < j; ++k2) d = fmaf(-sG[j * LMB + k2], sG[j * LMB + k2], d);
        const float inv = 1.f / sqrtf(d);
        __syncwarp();
        if (tid == 0) sgi[j] = inv;
        for (int i = j + 1 + tid; i < m; i += 32) {
          float s = sG[i * LMB + j];
          for (int k2 = 0; k2 < j; ++k2) s = fmaf(-sG[i * LMB + k2], sG[j * LMB + k2], s);
          sG[i * LMB + j] = s * inv;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- S5: [K k] = -G^-1 [H h]: one thread per right-hand-side column (head columns + k)
    for (int col = tid; col <= NH; col += NT) {
      float y[m];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        float s = -((col < NH) ? sH[i * LH + col] : sh[i]);
#pragma unroll
        for (int k2 = 0; k2 < i; ++k2) s = fmaf(-sG[i * LMB + k2], y[k2], s);
        y[i] = s * sgi[i];
      }
#pragma unroll
      for (int i = m - 1; i >= 0; --i) {
        float s = y[i];
#pragma unroll
        for (int k2 = i + 1; k2 < m; ++k2) s = fmaf(-sG[k2 * LMB + i], y[k2], s);
        y[i] = s * sgi[i];
      }
      if (col < NH) {
#pragma unroll
        for (int i = 0; i < m; ++i) sK[i * LH + col] = y[i];
      } else {
#pragma unroll
        for (int i = 0; i < m; ++i) sOK[LH + i] = y[i];
      }
    }
    __syncthreads();
    // ---- S6: XY = X + K' H, p_h = qh + K' h
    for (int tile = tid; tile < TH * TH + LH; tile += NT) {
      if (tile < TH * TH) {
        const int i0 = (tile / TH) * 8, j0 = (tile % TH) * 8;
        float acc[8][8];
        rt_load<8, 8>(acc, X, LH, i0, j0);
        rt_atb<8, 8>(acc, sK, LH, i0, sH, LH, j0, m);
        rt_store<8, 8>(X, LH, i0, j0, acc);
      } else {
        const int i = tile - TH * TH;
        float s = sqh[i];
        for (int j = 0; j < m; ++j) s = fmaf(sK[j * LH + i], sh[j], s);
        s = (i < NH) ? s : 0.f;
        sOK[i] = s;
      }
    }
    __syncthreads();
    // ---- S7: P_hh = sym(Q_hh + XY) over unordered pairs; p_{t} (full) for the next step
    for (int idx = tid; idx < (LH / 2 + 1) * LH; idx += NT) {
      // unordered pairs {i, (i + d) mod LH}: every pair once, consecutive lanes walk consecutive i
      const int d = idx / LH, i = idx - d * LH;
      int j = i + d;
      if (j >= LH) j -= LH;
      if ((2 * d == LH && i >= LH / 2) || i >= NH || j >= NH) continue;
      float qij = 0.f, qji = 0.f;
      if (i == j) { qij = pk0[D::oQd + i]; qji = qij; }
      if (i >= 3 && i < 7 && j >= 3 && j < 7) { qij = pk0[D::oQq + (i - 3) * 4 + (j - 3)]; qji = pk0[D::oQq + (j - 3) * 4 + (i - 3)]; }
      const float s = 0.5f * ((qij + X[i * LH + j]) + (qji + X[j * LH + i]));
      X[i * LH + j] = s;
      X[j * LH + i] = s;
    }
    for (int i = tid; i < n; i += NT) sp[i] = (i < NH) ? sOK[i] : pk0[D::oq + i];
    __syncthreads();
    fence_async_smem();
    if (tid == 0) {
      bulk_s2g(gP + size_t(t) * D::szP, X, D::szP * sizeof(float));
      bulk_s2g(gK + size_t(t) * D::szK, sK, D::szK * sizeof(float));
      bulk_s2g(gOK + size_t(t) * D::OKS, sOK, D::OKS * sizeof(float));
      bulk_commit();
    }
    cur ^= 1;
    pkn ^= 1;
    if (t > 0) {
      mbar_wait(&bars[0], ph0);
      ph0 ^= 1;
    }
  }

  // ------------------------------------------------------------------------------------------ rollout + dual
  // dx_0 = c_0, du_t = K_t dx_t + k_t, dx_{t+1} = A_t dx_t + B_t du_t + c_{t+1} (primal_tvlqr.py:478-498);
  // dV_t = P_t dx_t + p_t (dual_tvlqr.py:19) with the block structure of P_t; slope = sum q.dX + sum r.dU.
  if (tid == 0) bulk_wait0();   // the results written above are read back below
  __syncthreads();
  float* stg[2] = {smem, smem + S::stage};
  float* sdx = smem + 2 * S::stage;   // [2][LN]
  float* sdu = sdx + 2 * LN;          // [LMB]
  float* sred = sdu + LMB;            // [NT / 32]
  constexpr int oSK = 0, oSA = oSK + D::szK, oSB = oSA + D::szA, oSP = oSB + D::szB, oSPK = oSP + D::szP, oSOK = oSPK + D::PKS;
  auto fetch = [&](int t, int b) {   // knot t = 0..N (t == N: only P, PK, OK)
    float* st = stg[b];
    uint32_t bytes = uint32_t((D::szP + D::PKS + D::OKS) * sizeof(float));
    if (t < N) bytes += uint32_t((D::szK + D::szA + D::szB) * sizeof(float));
    mbar_expect_tx(&bars[1 + b], bytes);
    if (t < N) {
      bulk_g2s(st + oSK, gK + size_t(t) * D::szK, D::szK * sizeof(float), &bars[1 + b]);
      bulk_g2s(st + oSA, gA + size_t(t) * D::szA, D::szA * sizeof(float), &bars[1 + b]);
      bulk_g2s(st + oSB, gB + size_t(t) * D::szB, D::szB * sizeof(float), &bars[1 + b]);
    }
    bulk_g2s(st + oSP, gP + size_t(t) * D::szP, D::szP * sizeof(float), &bars[1 + b]);
    bulk_g2s(st + oSPK, gPK + size_t(t) * D::PKS, D::PKS * sizeof(float), &bars[1 + b]);
    bulk_g2s(st + oSOK, gOK + size_t(t) * D::OKS, D::OKS * sizeof(float), &bars[1 + b]);
  };
  if (tid == 0) { fetch(0, 0); fetch(1, 1); }
  float* gdX = a.dX + size_t(e) * (N + 1) * n;
  float* gdU = a.dU + size_t(e) * N * m;
  float* gdV = a.dV + size_t(e) * (N + 1) * n;
  uint32_t phs[2] = {0, 0};
  float slope_acc = 0.f;
  constexpr int PD = (NT / m >= 8) ? 8 : ((NT / m >= 4) ? 4 : ((NT / m >= 2) ? 2 : 1));   // lanes per row of K
  constexpr int PX = (NT / n >= 8) ? 8 : ((NT / n >= 4) ? 4 : ((NT / n >= 2) ? 2 : 1));   // lanes per row of A / P
  for (int t = 0; t <= N; ++t) {
    const int b = t & 1;
    const float* st = stg[b];
    mbar_wait(&bars[1 + b], phs[b]);
    phs[b] ^= 1;
    float* dx = sdx + b * LN;
    float* dxn = sdx + (b ^ 1) * LN;
    const float* pk = st + oSPK;
    if (t == 0) {
      for (int i = tid; i < n; i += NT) dx[i] = pk[D::oc + i];
      __syncthreads();
    }
    // du = K dx_h + k
    if (t < N) {
      const float* K = st + oSK;
      for (int row0 = 0; row0 < m; row0 += NT / PD) {
        const int row = row0 + tid / PD, part = tid % PD;
        float s = 0.f;
        if (row < m)
          for (int j = part; j < NH; j += PD) s = fmaf(K[row * LH + j], dx[j], s);
#pragma unroll
        for (int o = PD / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < m && part == 0) {
          s += st[oSOK + LH + row];
          sdu[row] = s;
          gdU[size_t(t) * m + row] = s;
          slope_acc = fmaf(pk[D::oR + row], s, slope_acc);
        }
      }
    }
    // dV_t = P_t dx + p_t: dense head block, sparse tail (Q_t tail); dX_t out; slope
    {
      const float* P = st + oSP;
      for (int row0 = 0; row0 < n; row0 += NT / PX) {
        const int row = row0 + tid / PX, part = tid % PX;
        float s = 0.f;
        if (row < NH) {
          for (int j = part; j < NH; j += PX) s = fmaf(P[row * LH + j], dx[j], s);
        }
#pragma unroll
        for (int o = PX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < n && part == 0) {
          if (row < NH) {
            s += st[oSOK + row];
          } else {
            s = fmaf(pk[D::oQd + row], dx[row], pk[D::oq + row]);
            if (NCONE > 0 && row >= D::cone0) {
              const int blk = (row - D::cone0) / 3, k0 = D::cone0 + 3 * blk, ri = row - k0;
              const float* cb = pk + D::oQc + 9 * blk + 3 * ri;
              s = fmaf(cb[0], dx[k0], fmaf(cb[1], dx[k0 + 1], fmaf(cb[2], dx[k0 + 2], s)));
            }
          }
          gdV[size_t(t) * n + row] = s;
          gdX[size_t(t) * n + row] = dx[row];
          slope_acc = fmaf(pk[D::oq + row], dx[row], slope_acc);
        }
      }
    }
    if (t == N) break;
    __syncthreads();   // du complete
    // dx+ = A dx_h + B du + c_{t+1}: c_{t+1} sits in the OTHER stage's PK block (knot t+1, already requested)
    mbar_wait(&bars[1 + (b ^ 1)], phs[b ^ 1]);   // peek: do not flip the phase, the next iteration waits again
    {
      const float* Am = st + oSA;
      const float* Bm = st + oSB;
      const float* cn = stg[b ^ 1] + oSPK + D::oc;
      for (int row0 = 0; row0 < n; row0 += NT / PX) {
        const int row = row0 + tid / PX, part = tid % PX;
        float s = 0.f;
        if (row < n) {
          for (int j = part; j < NH; j += PX) s = fmaf(Am[row * LH + j], dx[j], s);
          for (int j = part; j < m; j += PX) s = fmaf(Bm[row * LMB + j], sdu[j], s);
        }
#pragma unroll
        for (int o = PX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < n && part == 0) dxn[row] = s + cn[row];
      }
    }
    __syncthreads();   // dx+ complete, every thread is done with stage b
    if (tid == 0 && t + 2 <= N) fetch(t + 2, b);
  }
  // slope: fixed-order block reduction (deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) slope_acc += __shfl_xor_sync(0xffffffffu, slope_acc, o);
  if ((tid & 31) == 0) sred[tid >> 5] = slope_acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w2 = 0; w2 < NT / 32; ++w2) s += sred[w2];
    a.slope[e] = s;
  }
}

}  // namespace mpxb
