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
  int nzc;              // leading columns of A that can be non-zero (0 or >= n: all); columns >= nzc must hold zeros
};

__host__ __device__ constexpr int lq_pad(int n) { return (n + 3) & ~3; }

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
  static constexpr int oVec = oK + LM * LN;  // p, w, c, q, h, kk, r, 1/diag(chol G) : 4*LN + 4*LM
  static constexpr int oMR = oVec + 4 * LN + 4 * LM + 8;  // M_t [n][m] and R_t [m][m] prefetched one knot ahead
  static constexpr int total = oMR + n * m + m * m;
  static constexpr size_t bytes = size_t(total) * sizeof(float);
  // resident CTAs per SM that shared memory allows (227 KB usable, 1 KB reserved per CTA), capped at 4
  static constexpr int min_blocks = (232448 / (bytes + 1024)) < 1 ? 1 : ((232448 / (bytes + 1024)) > 4 ? 4 : int(232448 / (bytes + 1024)));
};

// Layout of lq_seq_kernel's arena for a given number of live columns of A (LZ = pad4(nzc)): A and T1 = P A only
// need LZ columns, Q_t is staged unpadded ([n][n], odd stride) over T1.  For the whole-body quadrupeds (n = 61,
// nzc = 37) this is 56 KB instead of 67 KB, i.e. four resident CTAs per SM instead of three.
struct LqRtLayout {
  int oP, oA, oT, oB, oPB, oH, oG, oK, oVec, oR, total;
};
template <int n, int m> __host__ __device__ inline LqRtLayout lq_rt_layout(int LZ) {
  constexpr int LN = lq_pad(n), LM = lq_pad(m);
  LqRtLayout o;
  o.oP = 0;
  o.oA = LN * LN;                 // [LN][LZ]
  o.oT = o.oA + LN * LZ;          // T1 [LN][LZ], then Q_t [n][n]
  const int szT = (LN * LZ > lq_pad(n * n + 3)) ? LN * LZ : lq_pad(n * n + 3);  // + 3: cp_async_flat slack
  o.oB = o.oT + szT;              // [LN][LM]
  o.oPB = o.oB + LN * LM;         // [LN][LM]
  o.oH = o.oPB + LN * LM;         // [LM][LN]
  o.oG = o.oH + LM * LN;          // [LM][LM]
  o.oK = o.oG + LM * LM;          // [LM][LN]
  o.oVec = o.oK + LM * LN;        // p, w, c, q [LN]; h, kk, r, 1/diag(chol G) [LM]
  o.oR = o.oVec + 5 * LN + 4 * LM + 8;  // R_t [m][m], prefetched one knot ahead
  o.total = o.oR + lq_pad(m * m);
  return o;
}
// resident CTAs per SM (capped at 4) the register allocation is tuned for: the occupancy shared memory allows with
// the number of live columns of A the shipped whole-body models have (13 + 2 nj of n = 13 + 2 nj + 6 nc [+ 3 nc])
template <int n, int m> __host__ __device__ constexpr int lq_rt_min_blocks() {
  constexpr int LN = lq_pad(n), LM = lq_pad(m);
  constexpr int nzc = (n == 61 && m == 12) ? 37 : (n == 75 && m == 19) ? 51 : (n == 81 && m == 46) ? 57 : n;
  constexpr int LZ = lq_pad(nzc);
  constexpr int szT = (LN * LZ > lq_pad(n * n + 3)) ? LN * LZ : lq_pad(n * n + 3);
  constexpr int total = LN * LN + LN * LZ + szT + 2 * LN * LM + 2 * LM * LN + LM * LM + 5 * LN + 4 * LM + 8 + lq_pad(m * m);
  constexpr int b = 232448 / (total * 4 + 1024);
  return b < 1 ? 1 : (b > 4 ? 4 : b);
}
template <int n, int m> inline size_t lq_rt_bytes(int nzc) {
  const int z = (nzc > 0 && nzc < n) ? nzc : n;
  return size_t(lq_rt_layout<n, m>((z + 3) & ~3).total) * sizeof(float);
}

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

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
// Asynchronous copy of `count` contiguous floats whose global address is only 4-byte aligned: the destination is
// shifted by the source's offset inside its 16-byte line (0..3 floats; `dst_base` is 16-byte aligned and has 3 floats
// of slack), so everything but a head / tail of at most 3 floats moves as 16-byte cp.async.  Returns the shifted
// destination.
__device__ __forceinline__ float* cp_async_flat(float* dst_base, const float* src, int count, int tid, int nthreads) {
  const int sh = int((reinterpret_cast<unsigned long long>(src) >> 2) & 3ull);
  float* dst = dst_base + sh;
  int head = (4 - sh) & 3;
  if (head > count) head = count;
  if (tid < head) cp_async4(dst + tid, src + tid);
  const int nvec = (count - head) >> 2;
  for (int v = tid; v < nvec; v += nthreads) cp_async16(dst + head + 4 * v, src + head + 4 * v);
  const int done = head + 4 * nvec;
  if (tid < count - done) cp_async4(dst + done + tid, src + done + tid);
  return dst;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// asynchronous copy of the leading `cols` columns of a [rows, gcols] global matrix into a [., ld] shared tile
__device__ __forceinline__ void cp_async_matrix(float* dst, int ld, const float* src, int rows, int gcols, int cols, int tid,
                                                int nthreads) {
  // column j = tid mod CW with CW the power of two >= cols: no integer division, coalesced row segments
  int cw = 1, sh = 0;
  while (cw < cols && cw < nthreads) { cw <<= 1; ++sh; }
  const int j = tid & (cw - 1), step = nthreads >> sh;
  if (step == 0) {  // more columns than threads (not reached with the shipped sizes)
    for (int idx = tid; idx < rows * cols; idx += nthreads) {
      const int i = idx / cols, jj = idx - i * cols;
      cp_async4(dst + i * ld + jj, src + i * gcols + jj);
    }
    return;
  }
  if (j < cols)
    for (int i = tid >> sh; i < rows; i += step) cp_async4(dst + i * ld + j, src + i * gcols + j);
}

__host__ __device__ constexpr int lq_parts(int rows, int nthreads) {  // lanes that share one row: largest power of two with rows * p <= nthreads
  int p = 32;
  while (p > 1 && rows * p > nthreads) p >>= 1;
  return p;
}

// Forward rollout + dual update + slope for the sequential backward pass (primal_tvlqr.py:478-498: dx_0 = c_0,
// du_t = K_t dx_t + k_t, dx_{t+1} = A_t dx_t + B_t du_t + c_{t+1}; dual_tvlqr.py:19: dV_t = P_t dx_t + p_t).
// The only serial chain is dx; its per-knot operands (K, A, B, q, r, c, k) are streamed through a cp.async double
// buffer so a knot costs two shared-memory mat-vecs and three barriers.  dV needs no chain: it is evaluated for all
// knots at once afterwards, one warp per row of P.  `smem` is the whole (now dead) backward-pass arena.
template <int n, int m, int NT>
__device__ __forceinline__ void lq_rollout_stream(const LqArgs& a, int e, int tid, float* smem, int acols, int arena) {
  constexpr int LN = lq_pad(n), LM = lq_pad(m);
  constexpr int PD = lq_parts(m, NT), PX = lq_parts(n, NT);
  // one stage: K | A (acols wide) | B | q | r | c | k | [P | p].  P_t, p_t ride inside both stages when that fits the
  // arena (pmode 2), else in ONE extra buffer refilled as soon as dV_t is done (pmode 1), else dV is a separate pass
  const int SB0 = 2 * m * n + n * acols + 2 * n + 2 * m;
  const int extras = 2 * LN + LM + NT / 32;
  const int pmode = (2 * (SB0 + n * n + n) + extras <= arena) ? 2 : ((2 * SB0 + n * n + n + extras <= arena) ? 1 : 0);
  const bool withP = pmode != 0;
  const int SB = SB0 + (pmode == 2 ? n * n + n : 0);
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
  float* stage = smem;
  float* sPone = smem + 2 * SB;                               // pmode 1: the single P | p buffer
  float* sdx = sPone + (pmode == 1 ? n * n + n : 0);          // [2][LN]
  float* sdu = sdx + 2 * LN;   // [LM]
  float* sred = sdu + LM;      // [NT / 32]
  auto prefetch = [&](int t, int buf) {
    float* st = stage + buf * SB;
    float* sK = st;
    float* sA = sK + m * n;
    float* sB = sA + n * acols;
    float* sq = sB + n * m;
    float* sr = sq + n;
    float* sc = sr + m;
    float* sk = sc + n;
    float* sPm = sk + m;
    if (pmode == 2) {
      for (int i = tid; i < n * n; i += NT) cp_async4(sPm + i, gP + size_t(t) * n * n + i);
      for (int i = tid; i < n; i += NT) cp_async4(sPm + n * n + i, gp + size_t(t) * n + i);
    }
    for (int i = tid; i < m * n; i += NT) cp_async4(sK + i, gK + size_t(t) * m * n + i);
    cp_async_matrix(sA, acols, gA + size_t(t) * n * n, n, n, acols, tid, NT);
    for (int i = tid; i < n * m; i += NT) cp_async4(sB + i, gB + size_t(t) * n * m + i);
    for (int i = tid; i < n; i += NT) { cp_async4(sq + i, gq + size_t(t) * n + i); cp_async4(sc + i, gc + size_t(t + 1) * n + i); }
    for (int i = tid; i < m; i += NT) { cp_async4(sr + i, gr + size_t(t) * m + i); cp_async4(sk + i, gk + size_t(t) * m + i); }
    cp_async_commit();
  };
  auto prefetch_p = [&](int t) {  // pmode 1
    for (int i = tid; i < n * n; i += NT) cp_async4(sPone + i, gP + size_t(t) * n * n + i);
    for (int i = tid; i < n; i += NT) cp_async4(sPone + n * n + i, gp + size_t(t) * n + i);
    cp_async_commit();
  };
  __syncthreads();  // the backward pass is done with the arena
  if (pmode == 1) prefetch_p(0);
  prefetch(0, 0);
  for (int i = tid; i < n; i += NT) sdx[i] = gc[i];
  float slope_acc = 0.f;
  for (int t = 0; t < N; ++t) {
    const float* st = stage + (t & 1) * SB;
    const float* sK = st;
    const float* sA = sK + m * n;
    const float* sB = sA + n * acols;
    const float* sq = sB + n * m;
    const float* sr = sq + n;
    const float* sc = sr + m;
    const float* sk = sc + n;
    const float* sPm = (pmode == 1) ? sPone : sk + m;
    const float* dx = sdx + (t & 1) * LN;
    float* dxn = sdx + ((t + 1) & 1) * LN;
    if (t + 1 < N) {
      prefetch(t + 1, (t + 1) & 1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      cp_async_wait_all();
    }
    __syncthreads();
    {  // du = K dx + k : PD lanes per row
      const int row = tid / PD, part = tid - row * PD;
      float s = 0.f;
      if (row < m)
        for (int j = part; j < n; j += PD) s = fmaf(sK[row * n + j], dx[j], s);
#pragma unroll
      for (int o = PD / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (row < m && part == 0) {
        s += sk[row];
        sdu[row] = s;
        gdU[size_t(t) * m + row] = s;
        slope_acc = fmaf(sr[row], s, slope_acc);
      }
    }
    if (tid < n) {
      gdX[size_t(t) * n + tid] = dx[tid];
      slope_acc = fmaf(sq[tid], dx[tid], slope_acc);
    }
    if (withP) {  // dV_t = P_t dx_t + p_t, off the dx chain
      const int row = tid / PX, part = tid - row * PX;
      float s = 0.f;
      if (row < n)
        for (int j = part; j < n; j += PX) s = fmaf(sPm[row * n + j], dx[j], s);
#pragma unroll
      for (int o = PX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (row < n && part == 0) gdV[size_t(t) * n + row] = s + sPm[n * n + row];
    }
    __syncthreads();
    if (pmode == 1 && t + 1 < N) prefetch_p(t + 1);  // every worker is done with P_t
    {  // dx+ = A dx + B du + c : PX lanes per row
      const int row = tid / PX, part = tid - row * PX;
      float s = 0.f;
      if (row < n) {
        for (int j = part; j < acols; j += PX) s = fmaf(sA[row * acols + j], dx[j], s);
        for (int j = part; j < m; j += PX) s = fmaf(sB[row * m + j], sdu[j], s);
      }
#pragma unroll
      for (int o = PX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (row < n && part == 0) dxn[row] = s + sc[row];
    }
    __syncthreads();
  }
  {
    const float* dx = sdx + (N & 1) * LN;
    if (tid < n) {
      gdX[size_t(N) * n + tid] = dx[tid];
      slope_acc = fmaf(gq[size_t(N) * n + tid], dx[tid], slope_acc);
    }
  }
  // slope = sum q.dX + sum r.dU : fixed-order block reduction (deterministic)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) slope_acc += __shfl_xor_sync(0xffffffffu, slope_acc, o);
  if ((tid & 31) == 0) sred[tid >> 5] = slope_acc;
  __syncthreads();  // also publishes gdX to the whole block
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NT / 32; ++w) s += sred[w];
    a.slope[e] = s;
  }
  // dV_t = P_t dx_t + p_t for the knots not covered above (all of them when P did not fit the stages): one warp per
  // row, four rows in flight
  const int warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;
  const int rows = (N + 1) * n;
  for (int base = (withP ? N * n : 0) + warp * 4; base < rows; base += nwarp * 4) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = base + u;
      s[u] = 0.f;
      if (r < rows) {
        const int t = r / n;
        const float* prow = gP + size_t(r) * n;
        const float* dx = gdX + size_t(t) * n;
        for (int j = lane; j < n; j += 32) s[u] = fmaf(prow[j], dx[j], s[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
      const int r = base + u;
      if (lane == 0 && r < rows) gdV[r] = s[u] + gp[r];
    }
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

// acc[4][4] += A[0:K, i0:i0+4]^T B[0:K, j0:j0+4]  (both operands k-major in shared memory, one LDS.128 each per k)
__device__ __forceinline__ void tile_atb(float (&acc)[4][4], const float* __restrict__ A, int lda, int i0,
                                         const float* __restrict__ B, int ldb, int j0, int K) {
#pragma unroll 4
  for (int kk = 0; kk < K; ++kk) {
    const float4 ta = *reinterpret_cast<const float4*>(A + kk * lda + i0);
    const float4 tb = *reinterpret_cast<const float4*>(B + kk * ldb + j0);
    const float av[4] = {ta.x, ta.y, ta.z, ta.w}, bv[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
  }
}
__device__ __forceinline__ void tile_zero(float (&acc)[4][4]) {
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
}
__device__ __forceinline__ void tile_store(float* __restrict__ C, int ldc, int i0, int j0, const float (&acc)[4][4]) {
#pragma unroll
  for (int x = 0; x < 4; ++x)
    *reinterpret_cast<float4*>(C + (i0 + x) * ldc + j0) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
}
// Sequential Riccati backward pass + rollout, one CTA per environment.
//
// Per knot (lqr_step, primal_tvlqr.py:11-52), with Az = the leading nzc columns of A_t (columns >= nzc are zero:
// the foot / GRF tail of the whole-body state never enters the dynamics; nzc = n for a generic problem):
//   S1  T1 = P Az, PB = P B, w = p + P c                      (P is bitwise symmetric -> P^T form, LDS.128 operands)
//   S2  X = Az^T T1 (nzc x nzc block), H = B^T T1, G = B^T PB, h = B^T w + r, q' = q + A^T w
//       -> A_{t-1}, B_{t-1} and Q_t are prefetched with cp.async while S3..S6 run
//   S3  H += M^T, G = sym(R + G)      S4  Cholesky of G      S5  [K k] = -G^-1 [H h]
//   S6  XY = X + K^T H, p = q' + K^T h
//   S7  P = sym(Q + XY)   (pair-owned, conflict-free skewed walk; the reference's inner sym(A^T P A) only reorders
//       roundings of the same symmetric sum)
template <int n, int m, int NT>
__global__ void __launch_bounds__(NT, lq_rt_min_blocks<n, m>()) lq_seq_kernel(LqArgs a) {
  constexpr int LN = lq_pad(n), LM = lq_pad(m), TN = LN / 4, TM = LM / 4;
  extern __shared__ __align__(16) float smem[];
  const int nzc = (a.nzc > 0 && a.nzc < n) ? a.nzc : n;
  const int LZ = (nzc + 3) & ~3, TZ = LZ / 4;
  const LqRtLayout L = lq_rt_layout<n, m>(LZ);
  float* sP = smem + L.oP;    // P_{t+1}; X / XY of the step; P_t
  float* sA = smem + L.oA;    // A_t (leading LZ columns)
  float* sT = smem + L.oT;    // T1 = P A; then Q_t
  float* sB = smem + L.oB;    // [LN][LM]
  float* sPB = smem + L.oPB;  // [LN][LM]
  float* sH = smem + L.oH;    // [LM][LN]
  float* sG = smem + L.oG;    // [LM][LM]
  float* sK = smem + L.oK;    // [LM][LN]
  float* sp = smem + L.oVec;  // p   [LN]
  float* sw = sp + LN;         // w = p + P c
  float* sc = sw + LN;         // c_{t+1}
  float* sq = sc + LN;         // q_t + A^T w
  float* sh = sq + LN;         // h [LM]
  float* skk = sh + LM;        // k [LM]
  float* sr = skk + LM;        // r_t [LM]
  float* sgi = sr + LM;        // reciprocal Cholesky diagonal of G [LM]
  float* sqin = sgi + LM + 8;  // q_t [LN], prefetched one knot ahead (like c_{t+1} in sc and r_t in sr)
  float* sR = smem + L.oR;     // R_t [m][m]
  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  if (a.done && a.done[e]) return;
  const int N = a.N;
  const int acols = (LZ < n) ? LZ : n;  // columns of A fetched (those in [nzc, LZ) are stored zeros)
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

  // zero every buffer once: the padding of the cp.async targets is never written again
  for (int i = tid; i < L.total; i += NT) smem[i] = 0.f;
  __syncthreads();
  // terminal value function: P_N = Q_N, p_N = q_N   (primal_tvlqr.py:136); first A, B
  // (stored transposed: the products below read P^T, and Q_N is the one P that need not be symmetric)
  for (int idx = tid; idx < n * n; idx += NT) {
    const int i = idx / n, j = idx - i * n;
    cp_async4(sP + j * LN + i, gQ + size_t(N) * n * n + idx);
  }
  cp_async_matrix(sA, LZ, gA + size_t(N - 1) * n * n, n, n, acols, tid, NT);
  cp_async_matrix(sB, LM, gB + size_t(N - 1) * n * m, n, m, m, tid, NT);
  for (int i = tid; i < m * m; i += NT) cp_async4(sR + i, gR + size_t(N - 1) * m * m + i);
  for (int i = tid; i < n * m; i += NT) cp_async4(sK + i, gM + size_t(N - 1) * n * m + i);  // M_t rides in K's buffer until S5
  for (int i = tid; i < n; i += NT) { cp_async4(sc + i, gc + size_t(N) * n + i); cp_async4(sqin + i, gq + size_t(N - 1) * n + i); }
  for (int i = tid; i < m; i += NT) cp_async4(sr + i, gr + size_t(N - 1) * m + i);
  cp_async_commit();
  for (int i = tid; i < n; i += NT) { const float v = gq[size_t(N) * n + i]; sp[i] = v; gp[size_t(N) * n + i] = v; }
  for (int idx = tid; idx < n * n; idx += NT) gP[size_t(N) * n * n + idx] = gQ[size_t(N) * n * n + idx];

  for (int t = N - 1; t >= 0; --t) {
    cp_async_wait_all();
    __syncthreads();
    // ---- S1
    for (int tile = tid; tile < TN * TZ + TN * TM + LN; tile += NT) {
      float acc[4][4];
      if (tile < TN * TZ) {
        const int i0 = (tile / TZ) * 4, j0 = (tile % TZ) * 4;
        tile_zero(acc);
        tile_atb(acc, sP, LN, i0, sA, LZ, j0, n);
        tile_store(sT, LZ, i0, j0, acc);
      } else if (tile < TN * TZ + TN * TM) {
        const int tt = tile - TN * TZ;
        const int i0 = (tt / TM) * 4, j0 = (tt % TM) * 4;
        tile_zero(acc);
        tile_atb(acc, sP, LN, i0, sB, LM, j0, n);
        tile_store(sPB, LM, i0, j0, acc);
      } else {
        const int i = tile - TN * TZ - TN * TM;
        float s = sp[i];
        for (int j = 0; j < n; ++j) s = fmaf(sP[j * LN + i], sc[j], s);
        sw[i] = (i < n) ? s : 0.f;
      }
    }
    __syncthreads();
    // ---- S2
    for (int tile = tid; tile < TZ * TZ + TM * TZ + TM * TM + LM + LN; tile += NT) {
      float acc[4][4];
      if (tile < TZ * TZ) {
        const int i0 = (tile / TZ) * 4, j0 = (tile % TZ) * 4;
        tile_zero(acc);
        tile_atb(acc, sA, LZ, i0, sT, LZ, j0, n);
        tile_store(sP, LN, i0, j0, acc);  // P_{t+1} is dead
      } else if (tile < TZ * TZ + TM * TZ) {
        const int tt = tile - TZ * TZ;
        const int i0 = (tt / TZ) * 4, j0 = (tt % TZ) * 4;
        tile_zero(acc);
        tile_atb(acc, sB, LM, i0, sT, LZ, j0, n);
        tile_store(sH, LN, i0, j0, acc);
      } else if (tile < TZ * TZ + TM * TZ + TM * TM) {
        const int tt = tile - TZ * TZ - TM * TZ;
        const int i0 = (tt / TM) * 4, j0 = (tt % TM) * 4;
        tile_zero(acc);
        tile_atb(acc, sB, LM, i0, sPB, LM, j0, n);
        tile_store(sG, LM, i0, j0, acc);
      } else if (tile < TZ * TZ + TM * TZ + TM * TM + LM) {
        const int i = tile - TZ * TZ - TM * TZ - TM * TM;
        float s = sr[i];
        for (int j = 0; j < n; ++j) s = fmaf(sB[j * LM + i], sw[j], s);
        sh[i] = (i < m) ? s : 0.f;
      } else {
        const int i = tile - TZ * TZ - TM * TZ - TM * TM - LM;
        float s = (i < n) ? sqin[i] : 0.f;
        if (i < LZ)
          for (int j = 0; j < n; ++j) s = fmaf(sA[j * LZ + i], sw[j], s);
        sq[i] = s;
      }
    }
    __syncthreads();
    // prefetch: Q_t over T1, A_{t-1} / B_{t-1} over A_t / B_t
    const float* sQ = cp_async_flat(sT, gQ + size_t(t) * n * n, n * n, tid, NT);  // unpadded [n][n], 16-byte copies
    if (t > 0) {
      cp_async_matrix(sA, LZ, gA + size_t(t - 1) * n * n, n, n, acols, tid, NT);
      cp_async_matrix(sB, LM, gB + size_t(t - 1) * n * m, n, m, m, tid, NT);
      for (int i = tid; i < n; i += NT) { cp_async4(sc + i, gc + size_t(t) * n + i); cp_async4(sqin + i, gq + size_t(t - 1) * n + i); }
      for (int i = tid; i < m; i += NT) cp_async4(sr + i, gr + size_t(t - 1) * m + i);
    }
    cp_async_commit();
    // ---- S3: H (columns >= LZ hold no B^T P A part) += M^T ; G = sym(R + G)   (M_t, R_t arrived with A_t, B_t)
    for (int idx = tid; idx < m * n; idx += NT) {
      const int j = idx / m, i = idx - j * m;  // M is [n][m]
      const float mv = sK[idx];  // M_t, staged in K's buffer (K_t is written in S5)
      sH[i * LN + j] = (j < LZ ? sH[i * LN + j] : 0.f) + mv;
    }
    for (int idx = tid; idx < m * m; idx += NT) {
      const int i = idx / m, j = idx - i * m;
      if (i <= j) {
        const float gij = sR[i * m + j] + sG[i * LM + j];
        const float gji = sR[j * m + i] + sG[j * LM + i];
        const float s = 0.5f * (gij + gji);
        sG[i * LM + j] = s;  // both halves are read only by this thread before being written
        sG[j * LM + i] = s;
      }
    }
    __syncthreads();
    if (t > 0) {  // R_{t-1}: its buffer was last read in S3
      for (int i = tid; i < m * m; i += NT) cp_async4(sR + i, gR + size_t(t - 1) * m * m + i);
      cp_async_commit();
    }
    // ---- S4: Cholesky G = L L^T in place (lower), one warp; NaN on a non-SPD G propagates like jax cho_factor
    if (tid < 32) {
      for (int j = 0; j < m; ++j) {
        float d = sG[j * LM + j];
        for (int k2 = 0; k2 < j; ++k2) d = fmaf(-sG[j * LM + k2], sG[j * LM + k2], d);
        const float inv = chol_pivot_inv(d);
        __syncwarp();
        if (tid == 0) sgi[j] = inv;
        for (int i = j + 1 + tid; i < m; i += 32) {
          float s = sG[i * LM + j];
          for (int k2 = 0; k2 < j; ++k2) s = fmaf(-sG[i * LM + k2], sG[j * LM + k2], s);
          sG[i * LM + j] = s * inv;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- S5: [K k] = -G^-1 [H h]: one thread per right-hand-side column
    for (int col = tid; col <= n; col += NT) {
      float y[m];
#pragma unroll
      for (int i = 0; i < m; ++i) {
        float s = -((col < n) ? sH[i * LN + col] : sh[i]);
#pragma unroll
        for (int k2 = 0; k2 < i; ++k2) s = fmaf(-sG[i * LM + k2], y[k2], s);
        y[i] = s * sgi[i];
      }
#pragma unroll
      for (int i = m - 1; i >= 0; --i) {
        float s = y[i];
#pragma unroll
        for (int k2 = i + 1; k2 < m; ++k2) s = fmaf(-sG[k2 * LM + i], y[k2], s);
        y[i] = s * sgi[i];
      }
      if (col < n) {
#pragma unroll
        for (int i = 0; i < m; ++i) { sK[i * LN + col] = y[i]; gK[size_t(t) * m * n + i * n + col] = y[i]; }
      } else {
#pragma unroll
        for (int i = 0; i < m; ++i) { skk[i] = y[i]; gk[size_t(t) * m + i] = y[i]; }
      }
    }
    __syncthreads();
    // ---- S6: XY = X + K^T H (X lives in the leading LZ x LZ block of sP), p = q' + K^T h
    for (int tile = tid; tile < TN * TN + LN; tile += NT) {
      if (tile < TN * TN) {
        const int ti = tile / TN, tj = tile % TN;
        const int i0 = ti * 4, j0 = tj * 4;
        float acc[4][4];
        if (ti < TZ && tj < TZ) {
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float4 v = *reinterpret_cast<const float4*>(sP + (i0 + x) * LN + j0);
            acc[x][0] = v.x; acc[x][1] = v.y; acc[x][2] = v.z; acc[x][3] = v.w;
          }
        } else {
          tile_zero(acc);
        }
        tile_atb(acc, sK, LN, i0, sH, LN, j0, m);
        tile_store(sP, LN, i0, j0, acc);
      } else {
        const int i = tile - TN * TN;
        float s = sq[i];
        for (int j = 0; j < m; ++j) s = fmaf(sK[j * LN + i], sh[j], s);
        s = (i < n) ? s : 0.f;
        sp[i] = s;
        if (i < n) gp[size_t(t) * n + i] = s;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (t > 0) {  // K_t is dead: stage M_{t-1} in its buffer
      for (int i = tid; i < n * m; i += NT) cp_async4(sK + i, gM + size_t(t - 1) * n * m + i);
      cp_async_commit();
    }
    // ---- S7: P = sym(Q + XY): unordered pairs {i, (i + d) mod LN}; lanes walk i, so both accesses hit 32 banks
    for (int idx = tid; idx < (LN / 2 + 1) * LN; idx += NT) {
      const int d = idx / LN, i = idx - d * LN;
      int j = i + d;
      if (j >= LN) j -= LN;
      if ((2 * d == LN && i >= LN / 2) || i >= n || j >= n) continue;
      const float zij = sQ[i * n + j] + sP[i * LN + j];
      const float zji = sQ[j * n + i] + sP[j * LN + i];
      const float s = 0.5f * (zij + zji);
      sP[i * LN + j] = s;
      sP[j * LN + i] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += NT) {
      const int i = idx / n, j = idx - i * n;
      gP[size_t(t) * n * n + idx] = sP[i * LN + j];
    }
    // the next S1 reads sP (after its barrier); sT (Q) is dead, T1 overwrites it after that barrier as well
  }
  __syncthreads();

  // ---- forward rollout + dual (primal_tvlqr.py:484-498, dual_tvlqr.py:19) ----
  lq_rollout_stream<n, m, NT>(a, e, tid, smem, acols, L.total);
}

}  // namespace mpxb
