// wb_kernels.cuh -- whole-body model kernels: fused evaluate + linearise + quadratise and the 11-alpha trial
// evaluation of the filter line search.  Reference: optimizers.py:580-602 (model_evaluator), 150-173
// (linearize / hessian_approx / lagrangian), 518-540 (trial points); models.py:39-217; objectives.py:77-450.
//
// Linearisation: one CTA per (environment, knot); thread d carries forward-mode direction d of [qpos | qvel | u]
// through the whole dynamics (the foot / GRF tail of the state does not enter f, so those columns of A are
// zero: models.py:61,95,109-115).  A second pass carries the cost gradient; the Gauss-Newton blocks follow from
// the residual map (identity except the 3x4 quaternion-error Jacobian) and the penalty curvatures.
#pragma once
#include "../../include/mpxb.h"
#include "model_io.cuh"
#include "wb.cuh"
#include "wb_tree.cuh"
#include <stdlib.h>

namespace mpxb {

// residual row(s) of state index j (objectives.py residual(): identity except the quaternion block)
struct WbResMap {
  int cnt, a[3];
  float c[3];
};
__device__ __forceinline__ WbResMap wb_resmap_x(const WbModel& md, int j, const float* G, const float* contact) {
  WbResMap r;
  const int nj = md.nj, nf = 3 * md.nc;
  r.cnt = 1; r.c[0] = 1.f;
  if (j < 3) {
    r.a[0] = j;
  } else if (j < 7) {
    r.cnt = 3;
    for (int k = 0; k < 3; ++k) { r.a[k] = 3 + k; r.c[k] = G[k * 4 + j - 3]; }
  } else if (j < 13 + 2 * nj) {
    r.a[0] = j - 1;
  } else if (j < 13 + 2 * nj + nf) {
    r.a[0] = j - 1;
    if (md.kind == MPXB_KIND_QUAD_WB && !md.swing_tracking) r.c[0] = contact[(j - 13 - 2 * nj) / 3];
  } else {
    r.a[0] = j - 1 + nj;  // GRF carried in the state (quadruped / H1)
  }
  return r;
}

// Shared-memory scratch of wb_knot_cost_blocks
template <int NV, int NC> struct WbCostSmem {
  static constexpr int NJ = NV - 6, NF = 3 * NC, NX = 13 + 2 * NJ + 2 * NF, NU = NJ + NF, NW = 12 + 3 * NJ + 2 * NF;
  float r[NW], y[NW], part[NW], pen[NX + NU], dpen[NX + NU], rc[NX];
  int ra[NX];
};

// d penalty / dc with the reference's clip / where derivative conventions (forward mode on the scalar)
__device__ __forceinline__ Dual wb_pen(float c, float sgn, float alpha, float sigma) {
  const Dual p = penalty(Dual(c, sgn), alpha, sigma);
  return p;
}

// P5 of the linearise kernels: cost value + gradient, q, r and the Gauss-Newton blocks Q, R, M of one knot, in three
// parts so that the kernel can overlap them with the dynamics derivatives:
//   wb_cost_prepare  residuals, quaternion-error Jacobian, residual map, penalty values / slopes   (needs x, u only)
//   wb_gn_blocks     Q, M, R                                                                       (needs prepare)
//   wb_cost_finish   y = 0.5 (W_bb + W_bb') r_b, cost, gradient, q_t, r_t                           (needs [A|B]'V)
// Each part is written for worker `tid` of `NT` (any subset of the CTA's threads).
//
// The reference takes jax.grad of the cost (optimizers.py:111-120,165-169 through trajax linearize/quadratize); the
// cost is 0.5 * sum_blocks r_b' W_bb r_b + penalties (objectives.py:77-125, 204-250, 325-376), every residual is
// an identity map of one entry of (x, u) except the 3x4 quaternion-error block, so the gradient is assembled in
// closed form: y = 0.5 (W_bb + W_bb') r_b, grad = (dr/dx)' y + penalty slopes.
template <int NV, int NC>
__device__ __forceinline__ void wb_cost_prepare(const WbModel& md, const ModelIO& io, int e, int t, const float* sx,
                                                const float* su, float* sG, WbCostSmem<NV, NC>& cs, const float* ref,
                                                int tid, int NT) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  const int N = io.N, n = md.n, m = md.m, nj = NJ, nc = NC, nf = NF;
  const bool term = (t == N);
  const float* W = io.W + size_t(e) * io.w_stride;
  const int nw = md.nw;
  const float* contact = ref + 13 + nj + nf;
  const int ua0 = 12 + 2 * nj + nf;  // residual row of u[0]
  const int g0 = 13 + 2 * nj + nf, v0 = 13 + nj;
  const bool quad = (md.kind == MPXB_KIND_QUAD_WB);
  const bool masked = quad && !md.swing_tracking;  // contact_map (objectives.py:110-113)
  (void)N; (void)n; (void)m; (void)nc; (void)nf; (void)term; (void)W; (void)nw; (void)contact; (void)ua0; (void)g0; (void)v0;
  (void)quad; (void)masked;
  const int nact = term ? 12 + 2 * nj : nw;  // the terminal cost drops the foot / torque / GRF terms
  // ---- residuals, the quaternion-error Jacobian, the residual map of every state entry
  for (int a = tid; a < nw; a += NT) {
    float r = 0.f;
    if (a < 3) r = sx[a] - ref[a];
    else if (a < 6) r = 0.f;  // filled below
    else if (a < 12 + nj) r = sx[a + 1] - ref[a + 1];
    else if (a < 12 + 2 * nj) r = sx[a + 1];
    else if (a < ua0) {
      r = sx[a + 1] - ref[a + 1 - nj];
      if (masked) r *= contact[(a - 12 - 2 * nj) / 3];
    } else if (a < ua0 + nj) r = su[a - ua0];
    else {
      const int i = a - ua0 - nj;
      r = (md.grf_in_u ? su[nj + i] : sx[g0 + i]) - ref[13 + nj + nf + nc + i];
    }
    if (a < 3 || a >= 6) cs.r[a] = (a < nact) ? r : 0.f;
  }
  if (tid < 4) {
    Dual q[4], rs[3];
    for (int i = 0; i < 4; ++i) q[i] = Dual(sx[3 + i], i == tid ? 1.f : 0.f);
    quat_sub(q, ref + 3, rs);
    for (int k = 0; k < 3; ++k) {
      sG[k * 4 + tid] = rs[k].d;
      if (tid == 0) cs.r[3 + k] = rs[k].v;
    }
  }
  for (int j = tid; j < n; j += NT) {
    int a = j - 1;
    float c = 1.f;
    if (j < 3) a = j;
    else if (j < 7) a = -1;
    else if (j >= 13 + 2 * nj && j < g0) { if (masked) c = contact[(j - 13 - 2 * nj) / 3]; }
    else if (j >= g0) a = j - 1 + nj;
    cs.ra[j] = a;
    cs.rc[j] = c;
  }
  // ---- penalties (quadruped only): value and slope per owning variable
  for (int d = tid; d < n + m; d += NT) {
    float pv = 0.f, pd = 0.f;
    if (quad && !term) {
      if (d >= v0 && d < v0 + nj) {  // joint speed +-10
        const Dual p1 = wb_pen(10.01f - sx[d], -1.f, 1.f, 1.f), p2 = wb_pen(sx[d] + 10.01f, 1.f, 1.f, 1.f);
        pv = p1.v + p2.v; pd = p1.d + p2.d;
      } else if (d >= n && d < n + nj) {  // torque +-44
        const float tq = su[d - n];
        const Dual p1 = wb_pen(44.01f - tq, -1.f, 1.f, 1.f), p2 = wb_pen(tq + 44.01f, 1.f, 1.f, 1.f);
        pv = p1.v + p2.v; pd = p1.d + p2.d;
      } else if (d >= g0 && d < g0 + nf) {  // friction cone (mu 0.5, eps 1e-2), masked by contact
        const int c = (d - g0) / 3, k = d - g0 - 3 * c;
        const float gx = sx[g0 + 3 * c], gy = sx[g0 + 3 * c + 1], gz = sx[g0 + 3 * c + 2];
        const float sq = sqrtf(gy * gy + gx * gx + 1e-2f);
        const Dual p = wb_pen(0.5f * gz - sq, 1.f, 0.1f, 5.f);
        const float jf = (k == 0) ? -gx / sq : (k == 1 ? -gy / sq : 0.5f);
        pd = p.d * jf * contact[c];
        pv = (k == 2) ? p.v * contact[c] : 0.f;
      }
    }
    cs.pen[d] = pv;
    cs.dpen[d] = pd;
  }
}

// cost value only (trial evaluations): cs.part[a] = r_a (W_bb r_b)_a for the active residual rows
template <int NV, int NC>
__device__ __forceinline__ void wb_cost_rowsums(const WbModel& md, const ModelIO& io, int e, int t, WbCostSmem<NV, NC>& cs,
                                                int tid, int NT) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  const int nj = NJ, nf = NF, nw = md.nw;
  const float* W = io.W + size_t(e) * io.w_stride;
  const int ua0 = 12 + 2 * nj + nf;
  const int bb[10] = {0, 3, 6, 6 + nj, 9 + nj, 12 + nj, 12 + 2 * nj, ua0, ua0 + nj, ua0 + nj + nf};
  const int nact = (t == io.N) ? 12 + 2 * nj : nw;
  for (int a = tid; a < nw; a += NT) {
    float s1 = 0.f;
    if (a < nact) {
      int blk = 0;
#pragma unroll
      for (int k = 1; k < 9; ++k) blk += (a >= bb[k]) ? 1 : 0;
      for (int b = bb[blk]; b < bb[blk + 1]; ++b) s1 = fmaf(W[a * nw + b], cs.r[b], s1);
    }
    cs.part[a] = cs.r[a] * s1;
  }
}

template <int NV, int NC, class Sync>
__device__ __forceinline__ void wb_cost_finish(const WbModel& md, const ModelIO& io, const LinOut& o, int e, int t,
                                               float* sgrad, const float* satv, const float* sG,
                                               WbCostSmem<NV, NC>& cs, const float* ref, int tid, int NT, const float* Wov,
                                               Sync sync) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  const int N = io.N, n = md.n, m = md.m, nj = NJ, nc = NC, nf = NF;
  const bool term = (t == N);
  const float* W = Wov ? Wov : io.W + size_t(e) * io.w_stride;
  const int nw = md.nw;
  const float* contact = ref + 13 + nj + nf;
  const int ua0 = 12 + 2 * nj + nf;  // residual row of u[0]
  const int g0 = 13 + 2 * nj + nf, v0 = 13 + nj;
  const bool quad = (md.kind == MPXB_KIND_QUAD_WB);
  const bool masked = quad && !md.swing_tracking;  // contact_map (objectives.py:110-113)
  (void)N; (void)n; (void)m; (void)nc; (void)nf; (void)term; (void)W; (void)nw; (void)contact; (void)ua0; (void)g0; (void)v0;
  (void)quad; (void)masked;
  const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
  // block boundaries of the residual vector
  const int bb[10] = {0, 3, 6, 6 + nj, 9 + nj, 12 + nj, 12 + 2 * nj, ua0, ua0 + nj, ua0 + nj + nf};
  const int nact = term ? 12 + 2 * nj : nw;
  // ---- y = 0.5 (W_bb + W_bb') r_b ; cost parts r_a (W_bb r_b)_a
  for (int a = tid; a < nw; a += NT) {
    float s1 = 0.f, s2 = 0.f;
    if (a < nact) {
      int blk = 0;
#pragma unroll
      for (int k = 1; k < 9; ++k) blk += (a >= bb[k]) ? 1 : 0;
      const int lo = bb[blk], hi = bb[blk + 1];
      if (o.compact) {  // structured mode: W is diagonal
        s1 = W[a * nw + a] * cs.r[a];
        s2 = s1;
      } else {
        for (int b = lo; b < hi; ++b) {
          const float rb = cs.r[b];
          s1 = fmaf(W[a * nw + b], rb, s1);
          s2 = fmaf(W[b * nw + a], rb, s2);
        }
      }
    }
    cs.y[a] = 0.5f * (s1 + s2);
    cs.part[a] = cs.r[a] * s1;
  }
  sync();
  if (tid < 32) {
    float s = 0.f;
    for (int a = tid; a < nw; a += 32) s += cs.part[a];
    for (int d = tid; d < n + m; d += 32) s += cs.pen[d];
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (tid == 0) o.cost[size_t(e) * (N + 1) + t] = 0.5f * s;
  }
  // q_t = grad_x cost + A^T V_{t+1} - V_t ; r_t = grad_u cost + B^T V_{t+1}   (optimizers.py:111-120,165-169)
  for (int j = tid; j < n; j += NT) {
    float gj;
    if (j >= 3 && j < 7) gj = sG[j - 3] * cs.y[3] + sG[4 + j - 3] * cs.y[4] + sG[8 + j - 3] * cs.y[5];
    else gj = cs.rc[j] * cs.y[cs.ra[j]];
    gj = fmaf(0.5f, cs.dpen[j], gj);
    sgrad[j] = gj;
    float* qdst = o.compact ? o.PK + (size_t(e) * (N + 1) + t) * o.pks + o.oq : o.q + (size_t(e) * (N + 1) + t) * n;
    qdst[j] = gj + satv[j] - Vt[j];
  }
  if (!term)
    for (int i = tid; i < m; i += NT) {
      const float gi = fmaf(0.5f, cs.dpen[n + i], cs.y[ua0 + i]);
      sgrad[n + i] = gi;
      float* rdst = o.compact ? o.PK + (size_t(e) * (N + 1) + t) * o.pks + o.oR : o.r + (size_t(e) * N + t) * m;
      rdst[i] = gi + satv[n + i];
    }

}
template <int NV, int NC>
__device__ __forceinline__ void wb_cost_finish(const WbModel& md, const ModelIO& io, const LinOut& o, int e, int t,
                                               float* sgrad, const float* satv, const float* sG,
                                               WbCostSmem<NV, NC>& cs, const float* ref, int tid, int NT, const float* Wov = nullptr) {
  wb_cost_finish<NV, NC>(md, io, o, e, t, sgrad, satv, sG, cs, ref, tid, NT, Wov, [] { __syncthreads(); });
}

// Gauss-Newton blocks for a DIAGONAL weight matrix W (every shipped config; the host checks it before selecting the
// structured path): Q_t = diag + 4x4 quaternion block (+ 3x3 friction-cone blocks on the GRF tail of the quadrupeds),
// R_t = diag, M_t = 0 -- written into the packed knot block instead of three dense matrices.  Same arithmetic as
// wb_gn_blocks on the entries that can be non-zero.
template <int NV, int NC>
__device__ __forceinline__ void wb_gn_compact(const WbModel& md, const ModelIO& io, const LinOut& o, int e, int t,
                                              const float* sx, const float* su, const float* sG,
                                              const WbCostSmem<NV, NC>& cs, const float* ref, int tid, int NT) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  const int N = io.N, n = md.n, m = md.m, nj = NJ, nf = NF;
  const float* W = io.W + size_t(e) * io.w_stride;
  const int nw = md.nw;
  const float* contact = ref + 13 + nj + nf;
  const int ua0 = 12 + 2 * nj + nf;
  const int g0 = 13 + 2 * nj + nf, v0 = 13 + nj;
  const bool quad = (md.kind == MPXB_KIND_QUAD_WB);
  float* pk = o.PK + (size_t(e) * (N + 1) + t) * o.pks;
  for (int j = tid; j < n; j += NT) {
    const int aj = cs.ra[j];
    float v = 0.f;
    if (aj >= 0) {
      const float cj = cs.rc[j];
      v = cj * cj * W[aj * nw + aj];
      if (quad && j >= v0 && j < v0 + nj) {
        const float dq = sx[j];
        v += fminf(fmaxf(penalty_dd(10.01f - dq, 0.1f, 5.f), -1e6f), 1e6f) +
             fminf(fmaxf(penalty_dd(dq + 10.01f, 0.1f, 5.f), -1e6f), 1e6f);
      }
    }
    pk[o.oQd + j] = v;
  }
  for (int idx = tid; idx < 16; idx += NT) {
    const int i = idx >> 2, k = idx & 3;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) s = fmaf(sG[r * 4 + i], sG[r * 4 + k] * W[(3 + r) * nw + 3 + r], s);
    pk[o.oQq + idx] = s;
  }
  if (quad)
    for (int idx = tid; idx < 9 * NC; idx += NT) {
      const int c = idx / 9, a = (idx - 9 * c) / 3, b = idx - 9 * c - 3 * a;
      const float gx = sx[g0 + 3 * c], gy = sx[g0 + 3 * c + 1], gz = sx[g0 + 3 * c + 2];
      const float sq = sqrtf(gy * gy + gx * gx + 1e-1f);
      const float hcone = fminf(fmaxf(penalty_dd(0.5f * gz - sq, 0.1f, 5.f), -1e6f), 1e6f) * contact[c];
      const float jf3[3] = {-gx / sq, -gy / sq, 0.5f};
      pk[o.oQc + idx] = hcone * jf3[a] * jf3[b];
    }
  if (t < N)
    for (int i = tid; i < m; i += NT) {
      float s = W[(ua0 + i) * nw + ua0 + i];
      if (quad && i < nj) {
        const float tq = su[i];
        s += fminf(fmaxf(penalty_dd(44.01f - tq, 1.f, 1.f), -1e6f), 1e6f) +
             fminf(fmaxf(penalty_dd(tq + 44.01f, 1.f, 1.f), -1e6f), 1e6f);
      }
      pk[o.oRd + i] = s;
    }
}

template <int NV, int NC>
__device__ __forceinline__ void wb_gn_blocks(const WbModel& md, const ModelIO& io, const LinOut& o, int e, int t,
                                             const float* sx, const float* su, const float* sG,
                                             const WbCostSmem<NV, NC>& cs, const float* ref, int tid, int NT, const float* Wov = nullptr) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  const int N = io.N, n = md.n, m = md.m, nj = NJ, nc = NC, nf = NF;
  const bool term = (t == N);
  const float* W = Wov ? Wov : io.W + size_t(e) * io.w_stride;
  const int nw = md.nw;
  const float* contact = ref + 13 + nj + nf;
  const int ua0 = 12 + 2 * nj + nf;  // residual row of u[0]
  const int g0 = 13 + 2 * nj + nf, v0 = 13 + nj;
  const bool quad = (md.kind == MPXB_KIND_QUAD_WB);
  const bool masked = quad && !md.swing_tracking;  // contact_map (objectives.py:110-113)
  (void)N; (void)n; (void)m; (void)nc; (void)nf; (void)term; (void)W; (void)nw; (void)contact; (void)ua0; (void)g0; (void)v0;
  (void)quad; (void)masked;
  // ---- Gauss-Newton blocks Q = Jx^T W Jx (+ penalty curvature), R, M   (objectives.py:127-202, 252-323, 378-450)
  float* Q = o.Q + (size_t(e) * (N + 1) + t) * n * n;
  // columns of the identity-mapped state entries: Q[i][j] = c_i c_j W[a_i][a_j], rows walked block by block
  for (int j = tid; j < n; j += NT) {
    const int aj = cs.ra[j];
    if (aj < 0) continue;  // quaternion columns: below
    const float cj = cs.rc[j];
    const float* Wc = W + aj;
    float* Qc = Q + j;
    float curv_speed = 0.f;
    if (quad && j >= v0 && j < v0 + nj) {
      const float dq = sx[j];
      curv_speed = fminf(fmaxf(penalty_dd(10.01f - dq, 0.1f, 5.f), -1e6f), 1e6f) +
                   fminf(fmaxf(penalty_dd(dq + 10.01f, 0.1f, 5.f), -1e6f), 1e6f);
    }
    float hcone = 0.f, jfj = 0.f, jf3[3] = {0.f, 0.f, 0.f};
    const int cj_foot = (quad && j >= g0) ? (j - g0) / 3 : -1;
    if (cj_foot >= 0) {
      // friction cone curvature: eps = 1e-1 here (objectives.py:167), clip +-1e6, masked by contact
      const float gx = sx[g0 + 3 * cj_foot], gy = sx[g0 + 3 * cj_foot + 1], gz = sx[g0 + 3 * cj_foot + 2];
      const float sq = sqrtf(gy * gy + gx * gx + 1e-1f);
      hcone = fminf(fmaxf(penalty_dd(0.5f * gz - sq, 0.1f, 5.f), -1e6f), 1e6f) * contact[cj_foot];
      jf3[0] = -gx / sq; jf3[1] = -gy / sq; jf3[2] = 0.5f;
      jfj = jf3[j - g0 - 3 * cj_foot];
    }
    for (int i = 0; i < 3; ++i) Qc[i * n] = cj * Wc[i * nw];
    {
      const float w3 = cj * Wc[3 * nw], w4 = cj * Wc[4 * nw], w5 = cj * Wc[5 * nw];
      for (int i = 3; i < 7; ++i) Qc[i * n] = sG[i - 3] * w3 + sG[4 + i - 3] * w4 + sG[8 + i - 3] * w5;
    }
#pragma unroll 8
    for (int i = 7; i < 13 + 2 * nj; ++i) {
      float s = cj * Wc[(i - 1) * nw];
      if (i == j) s += curv_speed;
      Qc[i * n] = s;
    }
#pragma unroll 4
    for (int i = 13 + 2 * nj; i < g0; ++i) Qc[i * n] = cs.rc[i] * cj * Wc[(i - 1) * nw];
#pragma unroll 4
    for (int i = g0; i < n; ++i) {
      float s = cj * Wc[(i - 1 + nj) * nw];
      if (cj_foot >= 0 && (i - g0) / 3 == cj_foot) s = fmaf(hcone * jf3[i - g0 - 3 * cj_foot], jfj, s);
      Qc[i * n] = s;
    }
  }
  // the four quaternion columns: Q[i][3+k] = sum_b (dr_i)' W[.][3..5] G[.][k]
  for (int idx = tid; idx < 4 * n; idx += NT) {
    const int i = idx >> 2, k = idx & 3;
    const float g3 = sG[k], g4 = sG[4 + k], g5 = sG[8 + k];
    const int ai = cs.ra[i];
    float s;
    if (ai >= 0) {
      s = cs.rc[i] * (g3 * W[ai * nw + 3] + g4 * W[ai * nw + 4] + g5 * W[ai * nw + 5]);
    } else {
      s = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        s = fmaf(sG[r * 4 + i - 3], g3 * W[(3 + r) * nw + 3] + g4 * W[(3 + r) * nw + 4] + g5 * W[(3 + r) * nw + 5], s);
    }
    Q[i * n + 3 + k] = s;
  }
  if (!term) {
    float* M = o.M + (size_t(e) * N + t) * n * m;
    for (int idx = tid; idx < n * m; idx += NT) {
      const int i = idx / m, j = idx - i * m;
      const int ai = cs.ra[i];
      float s;
      if (ai >= 0) s = cs.rc[i] * W[ai * nw + ua0 + j];
      else s = sG[i - 3] * W[3 * nw + ua0 + j] + sG[4 + i - 3] * W[4 * nw + ua0 + j] + sG[8 + i - 3] * W[5 * nw + ua0 + j];
      M[idx] = s;
    }
    float* R = o.R + (size_t(e) * N + t) * m * m;
    for (int idx = tid; idx < m * m; idx += NT) {
      const int i = idx / m, j = idx - i * m;
      float s = W[(ua0 + i) * nw + ua0 + j];
      if (quad && i == j && i < nj) {
        const float tq = su[i];
        s += fminf(fmaxf(penalty_dd(44.01f - tq, 1.f, 1.f), -1e6f), 1e6f) +
             fminf(fmaxf(penalty_dd(tq + 44.01f, 1.f, 1.f), -1e6f), 1e6f);
      }
      R[idx] = s;
    }
  }
}

// Forward-mode path stack of the linearise kernel: the primal half is identical in every lane and lives once per
// warp, the tangent half is private to the lane (direction) and lane-interleaved in shared memory.
struct WbStackSplit {
  float* prim;
  float* tan;
  int stride;
  __device__ __forceinline__ Dual ld(int i) const { return Dual(prim[i], tan[i * stride]); }
  __device__ __forceinline__ void st(int i, Dual v) const { prim[i] = v.v; tan[i * stride] = v.d; }
};

template <int NV, int NC, int MAXD, int NT> struct WbLinSmem {
  static constexpr int NJ = NV - 6, NF = 3 * NC, NX = 13 + 2 * NJ + 2 * NF, NU = NJ + NF, NDS = 2 * NV + 1;
  WbKnot<NV, NC> k;
  float sx[NX], su[NU], svn[NX], sgrad[NX + NU], satv[NX + NU], sG[12];
  WbCostSmem<NV, NC> cs;
  float prim[NT / 32][MAXD * WB_SLOT];
  float tan[MAXD * WB_SLOT * NDS];
};

// One CTA per (environment, knot).  Phases (DESIGN.md "linearise kernel"):
//   P1  thread 0: primal path-stack pass  -> qM, qfrc_bias, J, feet in shared memory
//   P2  all threads: cooperative contact solve (Cholesky of qM, M^-1 J, J'M^-1J, GRFs, x+)
//   P3  thread d < nq+nv: forward-mode path-stack pass of state direction d, tangents folded by WbTangentSink
//   P4  thread d < nq+nv+m: analytic contact-solve derivative -> column d of [A | B], A'V / B'V
//   P5  all threads: cost gradient (forward mode), q, r, Gauss-Newton Q, R, M
template <int NV, int NC, int MAXD, int NT>
__global__ void __launch_bounds__(NT) wb_linearize_kernel(const WbModel* __restrict__ mdp, ModelIO io, LinOut o) {
  using S = WbLinSmem<NV, NC, MAXD, NT>;
  constexpr int NJ = S::NJ, NF = S::NF, NX = S::NX, NU = S::NU, NDS = S::NDS, NQ = NV + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const WbModel& md = *mdp;
  const int tid = threadIdx.x;
  const int N = io.N;
  const int e = blockIdx.x / (N + 1), t = blockIdx.x - e * (N + 1);
  if (io.done && io.done[e]) return;
  const int n = md.n, m = md.m, nj = NJ, nc = NC;
  const bool term = (t == N);
  const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
  const float* Ut = io.U + (size_t(e) * N + t) * m;
  const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
  const float* ref = io.reference + (size_t(e) * (N + 1) + t) * md.nref;
  const float* par = io.parameter + (size_t(e) * (N + 1) + t) * md.npar;
  const float* W = io.W + size_t(e) * io.w_stride;
  float* sx = sm.sx;
  float* su = sm.su;
  float* svn = sm.svn;
  float* sgrad = sm.sgrad;
  float* satv = sm.satv;
  float* sG = sm.sG;
  for (int i = tid; i < n; i += NT) {
    sx[i] = Xt[i];
    svn[i] = term ? 0.f : Vt[n + i];
    satv[i] = 0.f;
  }
  for (int i = tid; i < m; i += NT) { su[i] = term ? 0.f : Ut[i]; satv[n + i] = 0.f; }
  if (!term) {
    for (int i = tid; i < NV * NV; i += NT) sm.k.M[i] = 0.f;
    for (int i = tid; i < NV * NF; i += NT) sm.k.J[i] = 0.f;
  }
  __syncthreads();

  if (!term) {
    // ---- P1: primal tree pass
    if (tid == 0) {
      WbStackPlain<float> stk{sm.prim[0]};
      WbStoreSink<float> sink{sm.k.M, NV, sm.k.D, sm.k.J, NF, sm.k.feet};
      WbInPlain<float> in{sx, sx + NQ};
      wb_dfs<float>(md, in, stk, sink);
    }
    __syncthreads();
    // ---- P2: contact solve, next state, constraint residual
    wb_s3_primal<NV, NC>(md, sm.k, sx, su, par, tid, NT, [] { __syncthreads(); });
    {
      float* c = o.c + (size_t(e) * (N + 1) + t + 1) * n;
      const float* Xn = Xt + n;
      for (int i = tid; i < n; i += NT) c[i] = sm.k.xn[i] - Xn[i];
    }
    // ---- P3 + P4: one column of [A | B] per thread
    float* A = o.A + (size_t(e) * N + t) * n * n;
    float* B = o.B + (size_t(e) * N + t) * n * m;
    if (tid < NDS + m) {
      float t1[NV], t2[NV], ee[NV], hh[NF], dfeet[NF];
      const int is_u = tid >= NDS;
      if (!is_u) {
#pragma unroll
        for (int i = 0; i < NV; ++i) { t1[i] = 0.f; t2[i] = 0.f; ee[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < NF; ++i) { hh[i] = 0.f; dfeet[i] = 0.f; }
        WbStackSplit stk{sm.prim[tid >> 5], sm.tan + tid, NDS};
        WbTangentSink sink{t1, t2, ee, hh, dfeet, sm.k.w, sm.k.z, sm.k.a, sm.k.y, sm.k.grf, sm.k.wvz};
        WbInSeed seed{sx, NQ, tid};
        wb_dfs<Dual>(md, seed, stk, sink);
      }
      float atv = 0.f;
      float* dst = is_u ? B + (tid - NDS) : A + tid;
      const int ld = is_u ? m : n;
      wb_s3_tangent<NV, NC>(md, sm.k, par, is_u, is_u ? tid - NDS : tid, t1, t2, ee, hh, dfeet,
                            [&](int i, float v) { dst[i * ld] = v; atv = fmaf(v, svn[i], atv); });
      satv[is_u ? n + tid - NDS : tid] = atv;
    }
    // the foot / GRF tail of the state does not enter the dynamics
    const int ntail = n - NDS;
    for (int idx = tid; idx < n * ntail; idx += NT) {
      const int i = idx / ntail, j = NDS + idx - i * ntail;
      A[i * n + j] = 0.f;
    }
  }
  if (t == 0)
    for (int i = tid; i < n; i += NT) o.c[size_t(e) * (N + 1) * n + i] = io.x0[size_t(e) * n + i] - sx[i];

  wb_cost_prepare<NV, NC>(md, io, e, t, sx, su, sG, sm.cs, ref, tid, NT);
  __syncthreads();
  wb_cost_finish<NV, NC>(md, io, o, e, t, sgrad, satv, sG, sm.cs, ref, tid, NT);
  wb_gn_blocks<NV, NC>(md, io, o, e, t, sx, su, sG, sm.cs, ref, tid, NT);
}

// ---------------------------------------------------------------------------------------------------------------
// Linearise kernel v3: per-body tree quantities in shared memory + analytic derivatives (wb_tree.cuh).
//   S1  leaf-path workers: kinematics, joint motion vectors, spatial velocities / bias accelerations
//   S2  body workers: spatial inertia, momentum, bias force, foot offsets         S3  composites (component workers)
//   S4  dof workers: U_i, qfrc_bias ; pair workers: qM, foot Jacobians
//   P2  cooperative contact solve (as v2)                                          S5  acceleration problems w, a, z
//   S6  (direction, body) force perturbations + subtree sums
//   S7  thread d: folded tangents of direction d (analytic) -> column d of [A | B] (wb_s3_tangent), A'V / B'V
//   P5  cost gradient, q, r, Gauss-Newton blocks
// resident knots per SM the register allocation is tuned for.  Quadrupeds: 8 knots of 64 threads at 128 registers (16
// warps per SM) since the per-knot shared memory is 26.6 KB (narrow folded-tangent block over the tree's scratch
// block); at 36 KB per knot 6 knots at 168 registers were the best shared memory allowed (Go2 x1024: 1.76 -> 1.4 ms).
// H1 x4096: 4 CTAs of 128 threads at 128 registers -- 16 warps per SM,
// no spills -- 11.5 ms vs 12.3 ms for 4 x 96 threads at 168 and 14.7 ms for 3 x 128 at 168, r02y)
#ifndef MPXB_LIN3_MINB_H1
#define MPXB_LIN3_MINB_H1 4
#endif
#ifndef MPXB_LIN3_MINB
#define MPXB_LIN3_MINB (NT <= 64 ? 8 : (NT <= 96 ? 4 : (NV <= 25 ? MPXB_LIN3_MINB_H1 : 2)))
#endif
// knots per CTA: the kernel is ~16 k SASS instructions of straight-line code that every CTA executes once, and
// instruction fetch is 15 % of its stall samples; KPB > 1 puts KPB knots -- NT threads, one shared-memory struct and one
// named barrier each, otherwise independent -- into one CTA, so that a fetched line serves KPB x NT / 32 warps
#ifndef MPXB_LIN3_KPB
#ifndef MPXB_LIN3_KPB64
#define MPXB_LIN3_KPB64 2
#endif
#define MPXB_LIN3_KPB(NT) ((NT) <= 64 ? MPXB_LIN3_KPB64 : 1)
#endif
template <int NB, int NV, int NC, int NI, int NT> struct WbLin3Smem {
  static constexpr int NJ = NV - 6, NF = 3 * NC, NX = 13 + 2 * NJ + 2 * NF, NU = NJ + NF, NDS = 2 * NV + 1;
  static constexpr int NW = 12 + 3 * NJ + 2 * NF;
  WbKnot<NV, NC> k;
  WbCostSmem<NV, NC> cs;
  float sref[13 + NJ + 7 * NC];
  float sx[NX], su[NU], svn[NX], sgrad[NX + NU], satv[NX + NU], sG[12];
  // tr | fold | dw are dead once the columns of [A | B] are out: the weight matrix W is staged over them for the
  // Gauss-Newton / cost stage (its reads would otherwise go to L2: 216 KB of shared memory leave ~40 KB of L1)
  alignas(16) WbTree<NB, NV, NC, NI> tr;
  // folded tangents of every x-direction (wb_tree_fold_coop): t1 for all nq + nv directions ([row][direction]); the
  // rows t2 | e | h | dfeet exist for the nq position directions only (+ one all-zero column).  That narrow block lives
  // over the tree's scratch block when it fits (the quadrupeds: 36 -> 27 KB per knot, 8 knots per SM instead of 6)
  static constexpr int FW = 2 * NV + 1, FWN = NV + 2, SZN = 2 * NV * FWN, SZH = 2 * NF * FWN;
  // first fit over the scratch block: t2 | e (quadrupeds, H1), then h | dfeet behind them (quadrupeds)
  static constexpr int SCR = WbTree<NB, NV, NC, NI>::SCRATCH;
  static constexpr bool N_OVER_TREE = SZN <= SCR, H_OVER_TREE = N_OVER_TREE && SZN + SZH <= SCR;
  float fold[NV * FW];
  float foldn_own[N_OVER_TREE ? 4 : SZN], foldh_own[H_OVER_TREE ? 4 : SZH];
  WbDirW<NV> dw[NV - 2];   // direction frames of the 4 + nj position directions that move a subtree
  __device__ float* foldn() { return N_OVER_TREE ? tr.scratch() : foldn_own; }
  __device__ float* foldh() { return H_OVER_TREE ? tr.scratch() + SZN : foldh_own; }
};

__device__ __forceinline__ void wb_cp_async16(float* dst, const float* src) {
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}

template <int NB, int NV, int NC, int NI, int NT, int KPB>
__global__ void __launch_bounds__(NT * KPB, (MPXB_LIN3_MINB + KPB - 1) / KPB)
    wb_linearize3_kernel(const WbModel* __restrict__ mdp, ModelIO io, LinOut o) {
  using S = WbLin3Smem<NB, NV, NC, NI, NT>;
  constexpr int NDS = S::NDS, NQ = NV + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int slot = KPB > 1 ? int(threadIdx.x) / NT : 0;
  const int tid = int(threadIdx.x) - slot * NT;
  S& sm = reinterpret_cast<S*>(smem_raw)[slot];
  const WbModel& md = *mdp;
  const int N = io.N;
  const int item = int(blockIdx.x) * KPB + slot;
  if (item >= io.E * (N + 1)) return;   // the knots of a CTA share nothing, not even a barrier
  const int e = item / (N + 1), t = item - e * (N + 1);
  if (io.done && io.done[e]) return;
  // barrier over the NT workers of this knot
  auto sync = [slot] {
    if constexpr (KPB == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(NT) : "memory");
  };
  const int n = md.n, m = md.m;
  const bool term = (t == N);
  const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
  const float* Ut = io.U + (size_t(e) * N + t) * m;
  const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
  const float* par = io.parameter + (size_t(e) * (N + 1) + t) * md.npar;
  float* sx = sm.sx;
  float* su = sm.su;
  float* svn = sm.svn;
  float* satv = sm.satv;
  for (int i = tid; i < n; i += NT) {
    sx[i] = Xt[i];
    svn[i] = term ? 0.f : Vt[n + i];
    satv[i] = 0.f;
  }
  for (int i = tid; i < m; i += NT) { su[i] = term ? 0.f : Ut[i]; satv[n + i] = 0.f; }
  {
    const float* ref = io.reference + (size_t(e) * (N + 1) + t) * md.nref;
    for (int i = tid; i < md.nref; i += NT) sm.sref[i] = ref[i];
  }
  sync();
  // structured mode: the narrow tree stages S1-S4 (4 to ~20 work items) run on warp 0 alone, with warp-level barriers,
  // while the other warps do the cost preparation and the Gauss-Newton diagonals, which need x, u and the reference only
  const bool overlap = o.compact && !term;
  if (!overlap) {
    wb_cost_prepare<NV, NC>(md, io, e, t, sx, su, sm.sG, sm.cs, sm.sref, tid, NT);
    sync();
  }

  if (!term) {
    if (overlap) {
      if (tid < 32) {
        wb_tree_fk(md, sm.tr, sx, sx + NQ, tid, 32);
        __syncwarp();
        wb_tree_bodies(md, sm.tr, sx, sm.k.feet, tid, 32);
        __syncwarp();
        wb_tree_composite(md, sm.tr, tid, 32);
        __syncwarp();
        wb_tree_dofs(md, sm.tr, sm.k.D, tid, 32);
      } else {
        wb_cost_prepare<NV, NC>(md, io, e, t, sx, su, sm.sG, sm.cs, sm.sref, tid - 32, NT - 32);
        if constexpr (NT == 64) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(KPB + 1 + slot), "n"(NT - 32) : "memory");
        wb_gn_compact<NV, NC>(md, io, o, e, t, sx, su, sm.sG, sm.cs, sm.sref, tid - 32, NT - 32);
      }
      sync();
    } else {
      wb_tree_fk(md, sm.tr, sx, sx + NQ, tid, NT);
      sync();
      wb_tree_bodies(md, sm.tr, sx, sm.k.feet, tid, NT);
      sync();
      wb_tree_composite(md, sm.tr, tid, NT);
      sync();
      wb_tree_dofs(md, sm.tr, sm.k.D, tid, NT);
      sync();
    }
    wb_tree_mass_jac(md, sm.tr, sm.k.M, sm.k.J, tid, NT);
    sync();
    // explicit qM^-1 for the quadrupeds and H1 (no spills, shorter dependent chains); Talos keeps the triangular solves
    if constexpr (NV <= MPXB_INV_MAXNV) wb_s3_primal_inv<NV, NC>(md, sm.k, sx, su, par, tid, NT, sync);
    else wb_s3_primal<NV, NC>(md, sm.k, sx, su, par, tid, NT, sync);
    {
      float* c = o.c + (size_t(e) * (N + 1) + t + 1) * n;
      float* cpk = o.compact ? o.PK + (size_t(e) * (N + 1) + t + 1) * o.pks + o.oc : nullptr;
      const float* Xn = Xt + n;
      for (int i = tid; i < n; i += NT) {
        const float v = sm.k.xn[i] - Xn[i];
        c[i] = v;
        if (cpk) cpk[i] = v;
      }
    }
    wb_tree_acc(md, sm.tr, sm.k.w, sm.k.a, sm.k.z, tid, NT, sync);
    sync();
    wb_tree_items(md, sm.tr, tid, NT, sync);
    sync();
    float* A = o.A + (size_t(e) * N + t) * n * o.ldA;
    float* B = o.B + (size_t(e) * N + t) * n * o.ldB;
    // S7: folded tangents of all directions as cooperative work items, then thread d -> column d of [A | B].  The
    // Gauss-Newton blocks follow without a barrier, so the warps that hold few columns start on them early.
    float* const foldn = sm.foldn();
    float* const foldh = sm.foldh();
    wb_tree_fold_coop(md, sm.tr, sm.k.J, sm.k.w, sm.k.a, sm.k.z, sm.k.y, sm.k.grf, sm.k.wvz, sm.dw, sm.fold, S::FW, foldn, foldh,
                      S::FWN, tid, NT, sync, reinterpret_cast<const WbFoldTable*>(mdp + 1));
    sync();
    if (tid < NDS + m) {
      constexpr int FW = S::FW, NF3 = 3 * NC;
      float t1[NV];
      const int is_u = tid >= NDS;
      const int cc = is_u ? 0 : tid;
      if (!is_u) {
#pragma unroll
        for (int i = 0; i < NV; ++i) t1[i] = sm.fold[i * FW + cc];
      }
      // a velocity direction has t2 = t1 and e = h = dfeet = 0 (the zero column NQ of the narrow block)
      constexpr int FWN = S::FWN;
      const bool posdir = !is_u && cc < NQ;
      const int cn = posdir ? cc : NQ;
      const WbStrided t2 = posdir ? WbStrided{foldn + cc, FWN} : WbStrided{sm.fold + cc, FW};
      const WbStrided ee{foldn + NV * FWN + cn, FWN}, hh{foldh + cn, FWN}, dfeet{foldh + NF3 * FWN + cn, FWN};
      float atv = 0.f;
      float* dst = is_u ? B + (tid - NDS) : A + tid;
      const int ld = is_u ? o.ldB : o.ldA;
      auto sink = [&](int i, float v) { dst[i * ld] = v; atv = fmaf(v, svn[i], atv); };
      if constexpr (NV <= MPXB_INV_MAXNV) wb_s3_tangent_inv<NV, NC>(md, sm.k, par, is_u, is_u ? tid - NDS : tid, t1, t2, ee, hh, dfeet, sink);
      else wb_s3_tangent<NV, NC>(md, sm.k, par, is_u, is_u ? tid - NDS : tid, t1, t2, ee, hh, dfeet, sink);
      satv[is_u ? n + tid - NDS : tid] = atv;
    }
    // the foot / GRF tail of the state does not enter the dynamics (structured mode: only the pad columns exist)
    const int ntail = o.compact ? o.padA : n - NDS;
    for (int idx = tid; idx < n * ntail; idx += NT) {
      const int i = idx / ntail, j = NDS + idx - i * ntail;
      A[i * o.ldA + j] = 0.f;
    }
    const int nbpad = o.compact ? o.padB : 0;
    for (int idx = tid; idx < n * nbpad; idx += NT) {
      const int i = idx / nbpad, j = m + idx - i * nbpad;
      B[i * o.ldB + j] = 0.f;
    }
  }
  if (t == 0)
    for (int i = tid; i < n; i += NT) {
      const float v = io.x0[size_t(e) * n + i] - sx[i];
      o.c[size_t(e) * (N + 1) * n + i] = v;
      if (o.compact) o.PK[size_t(e) * (N + 1) * o.pks + o.oc + i] = v;
    }
  if (o.compact) {
    // diagonal W: no staging, the Gauss-Newton blocks are a few hundred numbers
    if (!overlap) wb_gn_compact<NV, NC>(md, io, o, e, t, sx, su, sm.sG, sm.cs, sm.sref, tid, NT);
    wb_cost_finish<NV, NC>(md, io, o, e, t, sm.sgrad, satv, sm.sG, sm.cs, sm.sref, tid, NT, nullptr, sync);
    return;
  }
  // stage W over the dead tree / fold arrays (16-byte cp.async when it is 16-byte aligned and fits)
  const float* Wg = io.W + size_t(e) * io.w_stride;
  const float* Wuse = nullptr;
  const int nww = md.nw * md.nw;
  sync();  // every worker is done with tr / fold / dw
  if ((reinterpret_cast<unsigned long long>(Wg) & 15ull) == 0 && (nww & 3) == 0 &&
      size_t(nww) * sizeof(float) <= sizeof(sm.tr) + sizeof(sm.fold) + sizeof(sm.dw)) {
    float* sW = reinterpret_cast<float*>(&sm.tr);
    for (int v = tid; v < nww / 4; v += NT) wb_cp_async16(sW + 4 * v, Wg + 4 * v);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    sync();
    Wuse = sW;
  }
  wb_gn_blocks<NV, NC>(md, io, o, e, t, sx, su, sm.sG, sm.cs, sm.sref, tid, NT, Wuse);
  wb_cost_finish<NV, NC>(md, io, o, e, t, sm.sgrad, satv, sm.sG, sm.cs, sm.sref, tid, NT, Wuse, sync);
}

// one warp per (env, alpha): lanes stride over knots; deterministic shuffle reduction
template <int NB, int NV, int NC>
__global__ void wb_trial_kernel(const WbModel* __restrict__ mdp, ModelIO io, TrialOut o) {
  constexpr int NJ = NV - 6, NF = 3 * NC;
  constexpr int NX = 13 + 2 * NJ + 2 * NF, NU = NJ + NF;
  __shared__ WbModel md;
  {
    const int* src = reinterpret_cast<const int*>(mdp);
    int* dst = reinterpret_cast<int*>(&md);
    for (int i = threadIdx.x; i < int(sizeof(WbModel) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int N = io.N;
  if (warp >= io.E * o.nia || !trial_selected(o)) return;
  const int e = warp / o.nia, ia = o.ia0 + warp - e * o.nia;
  if (io.done && io.done[e]) return;
  if (o.ia0 > 0 && o.need && !o.need[e]) return;  // alpha = 1 was accepted: the smaller steps are never looked at
  const int n = md.n, m = md.m;
  const float alpha = exp2f(-float(ia));
  const float* W = io.W + size_t(e) * io.w_stride;
  float theta = 0.f, phi = 0.f, vc = 0.f;
  for (int t = lane; t <= N; t += 32) {
    const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
    const float* dXt = o.dX + (size_t(e) * (N + 1) + t) * n;
    const float* Vt = io.V + (size_t(e) * (N + 1) + t) * n;
    const float* dVt = o.merit ? o.dV + (size_t(e) * (N + 1) + t) * n : Vt;
    const float* ref = io.reference + (size_t(e) * (N + 1) + t) * md.nref;
    const float* par = io.parameter + (size_t(e) * (N + 1) + t) * md.npar;
    float x[NX], u[NU];
    for (int i = 0; i < n; ++i) x[i] = fmaf(alpha, dXt[i], Xt[i]);
    if (t < N) {
      const float* Ut = io.U + (size_t(e) * N + t) * m;
      const float* dUt = o.dU + (size_t(e) * N + t) * m;
      for (int i = 0; i < m; ++i) u[i] = fmaf(alpha, dUt[i], Ut[i]);
    } else {
      for (int i = 0; i < m; ++i) u[i] = 0.f;
    }
    phi += wb_cost<float, NJ, NC>(md, W, ref, x, u, t == N);
    if (t < N) {
      float xn[NX];
      wb_dynamics<float, NB, NV, NC>(md, x, u, par, xn);
      for (int i = 0; i < n; ++i) {
        const float r = xn[i] - fmaf(alpha, dXt[n + i], Xt[n + i]);
        theta = fmaf(r, r, theta);
        if (o.merit) vc = fmaf(r, fmaf(alpha, dVt[n + i], Vt[n + i]), vc);
      }
    }
    if (t == 0) {
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

// ---------------------------------------------------------------------------------------------------------------
// whole_body_interface (mpc_utils.py:275-313): SRBD ground-reaction forces -> joint torques at the whole-body rate.
// One CTA per environment: qM, qfrc_bias, foot positions and foot Jacobians from the tree stages of the linearise
// kernel, then  csa = Kp (foot_ref - feet) + Kd (foot_ref_dot - J' qvel),
//   tau_fb = D[6:] + (M pinv(J') csa)[6:]   with pinv(J') = J (J'J)^-1 by Cholesky.  J [nv, 3 nc] includes the six
//                                            floating-base rows, which carry every foot force, so it keeps full column
//                                            rank where a 3x3 leg Jacobian does not (straight knee: lambda_min(J'J) =
//                                            0.04 for Aliengo, tests/test_gpu_wbi.py) and the reference's SVD
//                                            pseudo-inverse (mpc_utils.py:304) equals this left inverse; a J that did
//                                            lose rank (two feet at one point) would give NaN torques here, in-band,
//   tau_mpc = -(J grf)[6:],   tau = contact ? tau_mpc : tau_fb   per leg.
template <int NB, int NV, int NC> struct WbiSmem {
  static constexpr int NF = 3 * NC;
  WbTree<NB, NV, NC, 1> tr;
  float M[NV * NV], D[NV], J[NV * NF], feet[NF], G[NF * NF], gi[NF], csa[NF], x[NF], y[NV], qpos[NV + 1], qvel[NV];
};
template <int NB, int NV, int NC, int NT>
__global__ void __launch_bounds__(NT) wb_interface_kernel(const WbModel* __restrict__ mdp, WbiArgs a) {
  using S = WbiSmem<NB, NV, NC>;
  constexpr int NF = S::NF, NJ = NV - 6;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const WbModel& md = *mdp;
  const int e = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < NV + 1; i += NT) sm.qpos[i] = a.qpos[size_t(e) * (NV + 1) + i];
  for (int i = tid; i < NV; i += NT) sm.qvel[i] = a.qvel[size_t(e) * NV + i];
  __syncthreads();
  wb_tree_fk(md, sm.tr, sm.qpos, sm.qvel, tid, NT);
  __syncthreads();
  wb_tree_bodies(md, sm.tr, sm.qpos, sm.feet, tid, NT);
  __syncthreads();
  wb_tree_composite(md, sm.tr, tid, NT);
  __syncthreads();
  wb_tree_dofs(md, sm.tr, sm.D, tid, NT);
  __syncthreads();
  wb_tree_mass_jac(md, sm.tr, sm.M, sm.J, tid, NT);
  __syncthreads();
  const float* Kp = a.Kp;
  const float* Kd = a.Kd;
  // cartesian-space action and G = J'J
  for (int c = tid; c < NF; c += NT) {
    float s = 0.f;
    for (int k = 0; k < NF; ++k) {
      float ld = 0.f;
      for (int i = 0; i < NV; ++i) ld = fmaf(sm.J[i * NF + k], sm.qvel[i], ld);
      s = fmaf(Kp[c * NF + k], a.foot_ref[size_t(e) * NF + k] - sm.feet[k], s);
      s = fmaf(Kd[c * NF + k], a.foot_ref_dot[size_t(e) * NF + k] - ld, s);
    }
    sm.csa[c] = s;
  }
  for (int idx = tid; idx < NF * NF; idx += NT) {
    const int r = idx / NF, c = idx - r * NF;
    float s = 0.f;
    for (int i = 0; i < NV; ++i) s = fmaf(sm.J[i * NF + r], sm.J[i * NF + c], s);
    sm.G[idx] = s;
  }
  if (a.J)
    for (int idx = tid; idx < NV * NF; idx += NT) a.J[size_t(e) * NV * NF + idx] = sm.J[idx];
  __syncthreads();
  if (tid == 0) {  // 12 x 12 Cholesky solve: x = (J'J)^-1 csa
    for (int j = 0; j < NF; ++j) {
      float d = sm.G[j * NF + j];
      for (int p = 0; p < j; ++p) d = fmaf(-sm.G[j * NF + p], sm.G[j * NF + p], d);
      const float inv = chol_pivot_inv(d);
      sm.gi[j] = inv;
      for (int i = j + 1; i < NF; ++i) {
        float v = sm.G[i * NF + j];
        for (int p = 0; p < j; ++p) v = fmaf(-sm.G[i * NF + p], sm.G[j * NF + p], v);
        sm.G[i * NF + j] = v * inv;
      }
    }
    for (int i = 0; i < NF; ++i) {
      float v = sm.csa[i];
      for (int p = 0; p < i; ++p) v = fmaf(-sm.G[i * NF + p], sm.x[p], v);
      sm.x[i] = v * sm.gi[i];
    }
    for (int i = NF - 1; i >= 0; --i) {
      float v = sm.x[i];
      for (int p = i + 1; p < NF; ++p) v = fmaf(-sm.G[p * NF + i], sm.x[p], v);
      sm.x[i] = v * sm.gi[i];
    }
  }
  __syncthreads();
  for (int i = tid; i < NV; i += NT) {
    float s = 0.f;
    for (int c = 0; c < NF; ++c) s = fmaf(sm.J[i * NF + c], sm.x[c], s);
    sm.y[i] = s;
  }
  __syncthreads();
  for (int j = tid; j < NJ; j += NT) {
    float my = 0.f, jg = 0.f;
    for (int i = 0; i < NV; ++i) my = fmaf(sm.M[(6 + j) * NV + i], sm.y[i], my);
    for (int c = 0; c < NF; ++c) jg = fmaf(sm.J[(6 + j) * NF + c], a.grf[size_t(e) * NF + c], jg);
    const float mask = a.contact[size_t(e) * NC + j / 3];  // leg-major joints, three per leg (mpc_utils.py:308)
    a.tau[size_t(e) * NJ + j] = -jg * mask + (1.f - mask) * (sm.D[6 + j] + my);
  }
}

template <int NB, int NV, int NC>
inline int wb_class_interface(const WbModel& h, const WbModel* dev, const WbiArgs& a, cudaStream_t st) {
  constexpr int NT = 64;
  using S = WbiSmem<NB, NV, NC>;
  if (!h.tree_ok || h.nbody > NB || h.nv != NV || h.nc != NC || h.nj != 3 * NC) return MPXB_EUNSUPPORTED;
  auto kern = wb_interface_kernel<NB, NV, NC, NT>;
  kern<<<a.E, NT, sizeof(S), st>>>(dev, a);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

template <int NV, int NC, int MAXD, int NT>
inline int wb_launch_linearize_t(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st) {
  using S = WbLinSmem<NV, NC, MAXD, NT>;
  if (h.nv != NV || h.nc != NC || h.maxdepth > MAXD || 2 * NV + 1 + h.m > NT) return MPXB_EUNSUPPORTED;
  auto kern = wb_linearize_kernel<NV, NC, MAXD, NT>;
  kern<<<io.E * (io.N + 1), NT, sizeof(S), st>>>(dev, io, lo);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

template <int NB, int NV, int NC, int NI, int NT>
inline int wb_launch_linearize3_t(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st) {
  using S = WbLin3Smem<NB, NV, NC, NI, NT>;
  constexpr int KPB = MPXB_LIN3_KPB(NT);
  static_assert(sizeof(S) % 16 == 0, "the per-knot structs are laid out back to back");
  auto kern = wb_linearize3_kernel<NB, NV, NC, NI, NT, KPB>;
  const int items = io.E * (io.N + 1);
  kern<<<(items + KPB - 1) / KPB, NT * KPB, KPB * sizeof(S), st>>>(dev, io, lo);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

// One robot class = one translation unit (wb_inst_quad.cu / wb_inst_h1.cu / wb_inst_talos.cu) so the classes compile
// in parallel.  v3 (analytic derivatives) whenever the model's tables fit; force_v2 selects the forward-mode kernel.
template <int NB, int NV, int NC, int MAXD, int NI, int NT>
inline int wb_class_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st,
                              int force_v2) {
  if (h.tree_ok && !force_v2 && h.nbody <= NB && h.nitems <= NI) return wb_launch_linearize3_t<NB, NV, NC, NI, NT>(h, dev, io, lo, st);
  if (lo.compact) return MPXB_EUNSUPPORTED;  // the structured layout is written by the v3 kernel only
  return wb_launch_linearize_t<NV, NC, MAXD, NT>(h, dev, io, lo, st);
}
// ---------------------------------------------------------------------------------------------------------------
// Cooperative trial evaluation: one CTA per (env, trial point, knot).  The tree stages and the contact solve of the
// linearise kernel evaluate x+ = f(x, u) at (X, U) + alpha (dX, dU) with 64 workers instead of one thread with
// kilobytes of local memory; the closed-form residual machinery of the cost stage gives the cost value.  Per-knot
// partials of theta = sum c_new^2, phi = sum cost and (merit mode) sum (V + alpha dV) c_new go to the workspace and
// are summed by trial_reduce_kernel.
template <int NB, int NV, int NC> struct WbTrialSmem {
  static constexpr int NJ = NV - 6, NF = 3 * NC, NX = 13 + 2 * NJ + 2 * NF, NU = NJ + NF;
  WbKnot<NV, NC, (NV > MPXB_INV_MAXNV)> k;   // the quadruped / H1 trial solve forms no inverse: no L^-1 buffer
  WbTree<NB, NV, NC, 1> tr;
  WbCostSmem<NV, NC> cs;
  float sx[NX], su[NU], sxn[NX], svn[NX], sref[13 + NJ + 7 * NC], sG[12], red[3][2];
};

// resident CTAs per SM the register allocation aims at: what shared memory allows, at most 16 one-warp items resp.
// 8 two-warp items (128 registers per thread either way)
template <int NB, int NV, int NC, int NT> __host__ __device__ constexpr int wb_trial_min_blocks() {
  constexpr int cap = NT == 32 ? 16 : 8;
  constexpr int fit = int(232448 / (sizeof(WbTrialSmem<NB, NV, NC>) + 1024));
  return fit < 1 ? 1 : (fit > cap ? cap : fit);
}

template <int NB, int NV, int NC, int NT>
__global__ void __launch_bounds__(NT, wb_trial_min_blocks<NB, NV, NC, NT>()) wb_trial3_kernel(const WbModel* __restrict__ mdp, ModelIO io, TrialOut o) {
  using S = WbTrialSmem<NB, NV, NC>;
  constexpr int NQ = NV + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S& sm = *reinterpret_cast<S*>(smem_raw);
  const WbModel& md = *mdp;
  const int tid = threadIdx.x;
  const int N = io.N;
  const int per_env = o.nia * (N + 1);
  if (!trial_selected(o)) return;
  // grid-stride over the (env, trial point, knot) items: the first pass launches one CTA per item, the second pass a
  // fixed grid whose CTAs skip the environments that accepted alpha = 1
  // a later pass (ia0 > 0) walks the compact list of environments that rejected alpha = 1 (trial_finish)
  const bool listed = o.ia0 > 0 && o.need && o.need_count;
  const int nenv = listed ? *o.need_count : io.E;
  // NT == 32: one warp per item, every barrier is a warp barrier
  auto sync = [] {
    if constexpr (NT == 32) __syncwarp();
    else __syncthreads();
  };
  for (int item = blockIdx.x; item < nenv * per_env; item += gridDim.x) {
  const int le = item / per_env, rem = item - le * per_env;
  const int e = listed ? o.need_count[1 + le] : le + io.e0;
  const int ia = o.ia0 + rem / (N + 1), t = rem % (N + 1);
  if (io.done && io.done[e]) continue;
  if (o.ia0 > 0 && o.need && !o.need[e]) continue;  // accepted in an earlier pass: the smaller steps are never looked at
  const int n = md.n, m = md.m;
  const bool term = (t == N);
  const float alpha = exp2f(-float(ia));
  const float* Xt = io.X + (size_t(e) * (N + 1) + t) * n;
  const float* dXt = o.dX + (size_t(e) * (N + 1) + t) * n;
  const float* par = io.parameter + (size_t(e) * (N + 1) + t) * md.npar;
  for (int i = tid; i < n; i += NT) {
    sm.sx[i] = fmaf(alpha, dXt[i], Xt[i]);
    if (!term) {
      sm.sxn[i] = fmaf(alpha, dXt[n + i], Xt[n + i]);
      if (o.merit) {
        const size_t vi = (size_t(e) * (N + 1) + t + 1) * n + i;
        sm.svn[i] = fmaf(alpha, o.dV[vi], io.V[vi]);
      }
    }
  }
  for (int i = tid; i < m; i += NT)
    sm.su[i] = term ? 0.f : fmaf(alpha, o.dU[(size_t(e) * N + t) * m + i], io.U[(size_t(e) * N + t) * m + i]);
  {
    const float* ref = io.reference + (size_t(e) * (N + 1) + t) * md.nref;
    for (int i = tid; i < md.nref; i += NT) sm.sref[i] = ref[i];
  }
  sync();
  // ---- cost value: residuals + penalties, then 0.5 (sum_a r_a (W_bb r_b)_a + sum pen).  It needs x, u and the
  // reference only, so it runs on warp 1 while warp 0 walks the narrow tree stages S1-S4 (as in the linearise kernel)
  float theta = 0.f, vc = 0.f;
  if (term || NT == 32) {
    wb_cost_prepare<NV, NC>(md, io, e, t, sm.sx, sm.su, sm.sG, sm.cs, sm.sref, tid, NT);
    sync();
    wb_cost_rowsums<NV, NC>(md, io, e, t, sm.cs, tid, NT);
  }
  if (!term) {
    static_assert(NT == 64 || NT == 32, "two warps: tree | cost, or one warp for everything");
    if (NT == 32 || tid < 32) {
      wb_tree_fk(md, sm.tr, sm.sx, sm.sx + NQ, tid, 32);
      __syncwarp();
      wb_tree_bodies(md, sm.tr, sm.sx, sm.k.feet, tid, 32);
      __syncwarp();
      wb_tree_composite(md, sm.tr, tid, 32);
      __syncwarp();
      wb_tree_dofs(md, sm.tr, sm.k.D, tid, 32);
    } else {
      wb_cost_prepare<NV, NC>(md, io, e, t, sm.sx, sm.su, sm.sG, sm.cs, sm.sref, tid - 32, 32);
      __syncwarp();
      wb_cost_rowsums<NV, NC>(md, io, e, t, sm.cs, tid - 32, 32);
    }
    sync();
    wb_tree_mass_jac(md, sm.tr, sm.k.M, sm.k.J, tid, NT);
    sync();
    // robots with a contact solve: forward substitutions only (wb_s3_primal_fwd); Talos (GRFs are controls): one solve
    // (the launcher sends a model of these classes without a contact solve to the warp form)
    if constexpr (NV <= MPXB_INV_MAXNV) {
      wb_s3_primal_fwd<NV, NC>(md, sm.k, sm.sx, sm.su, par, tid, NT, sync);
    } else {
      wb_s3_primal<NV, NC, false>(md, sm.k, sm.sx, sm.su, par, tid, NT, sync);
    }
    for (int i = tid; i < n; i += NT) {
      const float r = sm.k.xn[i] - sm.sxn[i];
      theta = fmaf(r, r, theta);
      if (o.merit) vc = fmaf(r, sm.svn[i], vc);
    }
  }
  if (t == 0)
    for (int i = tid; i < n; i += NT) {
      const float r = io.x0[size_t(e) * n + i] - sm.sx[i];
      theta = fmaf(r, r, theta);
      if (o.merit) vc = fmaf(r, fmaf(alpha, o.dV[size_t(e) * (N + 1) * n + i], io.V[size_t(e) * (N + 1) * n + i]), vc);
    }
  float phi = 0.f;
  for (int a = tid; a < md.nw; a += NT) phi += sm.cs.part[a];
  for (int d = tid; d < n + m; d += NT) phi += sm.cs.pen[d];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    theta += __shfl_xor_sync(0xffffffffu, theta, s);
    phi += __shfl_xor_sync(0xffffffffu, phi, s);
    vc += __shfl_xor_sync(0xffffffffu, vc, s);
  }
  if constexpr (NT == 32) {
    if (tid == 0) {
      const size_t idx = (size_t(e) * NALPHA + ia) * (N + 1) + t;
      o.th_part[idx] = theta;
      o.ph_part[idx] = 0.5f * phi;
      if (o.merit) o.vc_part[idx] = vc;
    }
  } else {
    if ((tid & 31) == 0 && (tid >> 5) < 2) { sm.red[0][tid >> 5] = theta; sm.red[1][tid >> 5] = phi; sm.red[2][tid >> 5] = vc; }
    __syncthreads();
    if (tid == 0) {
      const size_t idx = (size_t(e) * NALPHA + ia) * (N + 1) + t;
      o.th_part[idx] = sm.red[0][0] + sm.red[0][1];
      o.ph_part[idx] = 0.5f * (sm.red[1][0] + sm.red[1][1]);
      if (o.merit) o.vc_part[idx] = sm.red[2][0] + sm.red[2][1];
    }
  }
  sync();  // the next item reuses the shared arrays
  }
}

// Per-device function attributes (> 48 KB dynamic shared memory) of every kernel of a robot class: called from
// mpxb_model_create, so that no launch sequence -- possibly under stream capture -- has to touch them.
template <int NB, int NV, int NC, int MAXD, int NI, int NT>
inline int wb_class_configure(bool with_interface) {
  auto set = [](auto kern, size_t bytes) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)) == cudaSuccess; };
  bool ok = set(wb_linearize_kernel<NV, NC, MAXD, NT>, sizeof(WbLinSmem<NV, NC, MAXD, NT>));
  ok = ok && set(wb_linearize3_kernel<NB, NV, NC, NI, NT, MPXB_LIN3_KPB(NT)>, MPXB_LIN3_KPB(NT) * sizeof(WbLin3Smem<NB, NV, NC, NI, NT>));
  ok = ok && set(wb_trial3_kernel<NB, NV, NC, 64>, sizeof(WbTrialSmem<NB, NV, NC>));
  ok = ok && set(wb_trial3_kernel<NB, NV, NC, 32>, sizeof(WbTrialSmem<NB, NV, NC>));
  if (with_interface) {
    if constexpr (NV == 6 + 3 * NC) ok = ok && set(wb_interface_kernel<NB, NV, NC, 64>, sizeof(WbiSmem<NB, NV, NC>));
  }
  return ok ? MPXB_OK : MPXB_ECUDA;
}

template <int NB, int NV, int NC>
inline int wb_class_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to_in, cudaStream_t st) {
  static const int legacy = [] { const char* v = getenv("MPXB_WB_TRIAL"); return (v && v[0] == '1') ? 1 : ((v && v[0] == '2') ? 2 : 0); }();
  // the cooperative kernel wins for the quadrupeds and H1 at any batch size (H1 x4096: 6.0 vs 6.5 ms, r02) and for
  // every robot when the batch is too small to fill the GPU with one warp per trial point (single-instance latency)
  // MPXB_WB_TRIAL=1: warp form only; =2: cooperative form for every class and batch size (measurement switches)
  const bool coop_ok = legacy != 1 && h.tree_ok && h.nbody <= NB && (NV > MPXB_INV_MAXNV || h.contact_solve) &&
                       (legacy == 2 || NV <= MPXB_INV_MAXNV || io.E * (io.N + 1) <= 148 * 8);
  TrialOut to = to_in;
  // first pass / eager evaluation: cooperative (if available) resp. warp-per-trial; second pass: both, selected on the
  // device by the number of environments that still need the smaller steps
  const bool second = to.ia0 > 0 && to.need_count != nullptr;
  const bool run_warp = !coop_ok || second || to.nia == NALPHA;
  const bool run_coop = coop_ok && to.nia != NALPHA;
  // second pass: the warp form (one warp per trial point, serial over the knots) pays off only when the rejecting
  // environments alone fill the GPU; up to a grid's worth of cooperative items (148 x 32 CTAs) the cooperative form
  // runs whatever the count -- a single instance that rejects alpha = 1 took 0.4 ms extra in the warp form
  {
    const int coop_envs = (148 * 32) / ((NALPHA - 1) * (io.N + 1));
    to.thr = io.E / 8 > coop_envs ? io.E / 8 : coop_envs;
  }
  if (run_warp) {
    to.sel = (second && coop_ok) ? 2 : 0;
    const int wpb = 2;
    const int blocks = (io.E * to.nia + wpb - 1) / wpb;
    wb_trial_kernel<NB, NV, NC><<<blocks, wpb * 32, 0, st>>>(dev, io, to);
    if (cudaGetLastError() != cudaSuccess) return MPXB_ECUDA;
  }
  if (run_coop) {
    to.sel = second ? 1 : 0;
    // one warp per item for the quadrupeds and H1 (no CTA barrier; 16 resp. 12 items resident per SM once the item's
    // shared memory carries neither L^-1 nor the acceleration problems): Go2 x1024 0.71 -> 0.41 ms, H1 x4096 5.24 ->
    // 4.87 ms (r02y).  MPXB_TRIAL_NT=32 / 64 forces a form (measurement switch)
    static const int nt32 = [] { const char* v = getenv("MPXB_TRIAL_NT"); return (v && v[0] == '3') ? 1 : ((v && v[0] == '6') ? 0 : (NV <= MPXB_INV_MAXNV ? 1 : 0)); }();
    using S = WbTrialSmem<NB, NV, NC>;
    const int items = io.E * to.nia * (io.N + 1);
    const int grid = (second && items > 148 * 32) ? 148 * 32 : items;
    if (nt32) wb_trial3_kernel<NB, NV, NC, 32><<<grid, 32, sizeof(S), st>>>(dev, io, to);
    else wb_trial3_kernel<NB, NV, NC, 64><<<grid, 64, sizeof(S), st>>>(dev, io, to);
    const int wpb = 4;
    trial_reduce_kernel<0><<<(io.E * to.nia + wpb - 1) / wpb, wpb * 32, 0, st>>>(io, to, h.n);
  }
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

}  // namespace mpxb
