// refgen.cuh -- gait timer + swing / foothold reference generator on the device: the step immediately before
// the solve in every wrapper tick.
//
// Follows reference mpx/utils/mpc_utils.py:7-14 (timer_run), 15-36 (terrain_orientation), 38-144
// (reference_generator) and 146-270 (reference_generator_srbd).  One thread per (environment, leg): the knots are
// a sequential recurrence on the leg's gait phase, lift-off point and previous foot reference; the leg-independent
// columns of the row (base pose / velocity references, joint reference) are written by the env's threads striding
// over the knots.  A block holds whole environments, so liftoff may be updated in place.
#pragma once
#include <cuda_runtime.h>
#include "../../include/mpxb.h"
#include "common.cuh"

namespace mpxb {

struct RefGenArgs {
  int E, N, nj, nc, srbd, n, use_terrain, advance_first;
  float dt, mpc_dt, duty_factor, step_freq, step_height, mass, clearence_speed;
  float foot0[3 * MPXB_MAX_CONTACT];
  float q0[MPXB_MAX_DOF];
  const float* x;        // [E, n]
  const float* foot;     // [E, *] rows of foot_stride floats, 3 nc used
  int foot_stride;
  const float* input;    // [E, 7]
  const float* contact;  // [E, nc] measured contact, or null (= 0)
  const float* timer_in; // [E, nc]
  float* timer_out;      // [E, nc] gait phase advanced by mpc_dt (may alias timer_in)
  float* liftoff;        // [E, 3 nc] in / out
  float* reference;      // [E, N+1, nref]
  float* parameter;      // [E, N+1, npar]
  float* foot_ref_dot;   // [E, N+1, 3 nc] or null
};

__device__ __forceinline__ float rg_advance(float t, float dt, float step_freq) {
  t = t + dt * step_freq;
  return (t > 1.f) ? t - 1.f : t;
}

__device__ __forceinline__ float rg_safe_norm3(const float* v) {
  const bool zero = fabsf(v[0]) <= 1e-8f && fabsf(v[1]) <= 1e-8f && fabsf(v[2]) <= 1e-8f;
  return zero ? 0.f : sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

// mpc_utils.py:15-36: quaternion (w,x,y,z) of the frame [front-back | left-right | normal] spanned by the lift-off
// points; matrix -> quaternion by the largest-pivot rule of Rotation.from_matrix
__device__ inline void rg_terrain_orientation(const float* lo, float cy, float sy, float* quat) {
  float vfb[3], vlr[3] = {-sy, cy, 0.f}, nrm[3];
  for (int i = 0; i < 3; ++i) vfb[i] = (lo[i] + lo[3 + i] - lo[6 + i] - lo[9 + i]) / 2.f;
  nrm[0] = vfb[1] * vlr[2] - vfb[2] * vlr[1];
  nrm[1] = vfb[2] * vlr[0] - vfb[0] * vlr[2];
  nrm[2] = vfb[0] * vlr[1] - vfb[1] * vlr[0];
  const float n0 = rg_safe_norm3(vfb), n1 = rg_safe_norm3(vlr), n2 = rg_safe_norm3(nrm);
  float R[3][3];
  for (int i = 0; i < 3; ++i) { R[i][0] = vfb[i] / n0; R[i][1] = vlr[i] / n1; R[i][2] = nrm[i] / n2; }
  const float tr = R[0][0] + R[1][1] + R[2][2];
  const float dec[4] = {R[0][0], R[1][1], R[2][2], tr};
  int choice = 0;
  for (int i = 1; i < 4; ++i) if (dec[i] > dec[choice]) choice = i;
  float q[4];  // x, y, z, w
  if (choice != 3) {
    const int i = choice, j = (i + 1) % 3, k = (i + 2) % 3;
    q[i] = 1.f - tr + 2.f * R[i][i];
    q[j] = R[j][i] + R[i][j];
    q[k] = R[k][i] + R[i][k];
    q[3] = R[k][j] - R[j][k];
  } else {
    q[0] = R[2][1] - R[1][2];
    q[1] = R[0][2] - R[2][0];
    q[2] = R[1][0] - R[0][1];
    q[3] = 1.f + tr;
  }
  const float nq = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  quat[0] = q[3] / nq; quat[1] = q[0] / nq; quat[2] = q[1] / nq; quat[3] = q[2] / nq;
}

template <int EPB>
__global__ void refgen_kernel(RefGenArgs a) {
  const int nc = a.nc, N = a.N, nj = a.nj;
  const int le = threadIdx.x / nc, leg = threadIdx.x - le * nc;
  const int e = blockIdx.x * EPB + le;
  const bool active = (le < EPB) && (e < a.E);
  const int nref = a.srbd ? 13 + nc : 13 + nj + 7 * nc;
  const int npar = a.srbd ? 4 * nc : nc;
  float timers[MPXB_MAX_CONTACT];
  float lo[3 * MPXB_MAX_CONTACT];
  float timer_next = 0.f;
  if (active) {
    for (int c = 0; c < nc; ++c) {
      const float t = a.timer_in[size_t(e) * nc + c];
      const float adv = rg_advance(t, a.mpc_dt, a.step_freq);
      timers[c] = a.advance_first ? adv : t;
      if (c == leg) timer_next = adv;
    }
    for (int i = 0; i < 3 * nc; ++i) lo[i] = a.liftoff[size_t(e) * 3 * nc + i];
  }
  __syncthreads();  // every read of timer_in / liftoff precedes the in-place writes
  if (!active) return;
  if (a.timer_out) a.timer_out[size_t(e) * nc + leg] = timer_next;

  const float* x = a.x + size_t(e) * a.n;
  const float* in = a.input + size_t(e) * 7;
  const float* foot = a.foot + size_t(e) * a.foot_stride;
  const float quat[4] = {x[3], x[4], x[5], x[6]};
  const float* dp = a.srbd ? x + 7 : x + 7 + nj;
  const float yaw = atan2f(2.f * (quat[0] * quat[3] + quat[1] * quat[2]), 1.f - 2.f * (quat[2] * quat[2] + quat[3] * quat[3]));
  const float cy = cosf(yaw), sy = sinf(yaw);
  float lzsum = 0.f;
  for (int c = 0; c < nc; ++c) lzsum += lo[3 * c + 2];
  const float height = in[6] + lzsum / float(nc);
  const float p[3] = {x[0], x[1], height};
  float quat0[4] = {1.f, 0.f, 0.f, 0.f};
  if (a.use_terrain) rg_terrain_orientation(lo, cy, sy, quat0);
  const float pitch = asinf(2.f * (quat0[0] * quat0[2] - quat0[3] * quat0[1]));
  const float cp = cosf(pitch), sp = sinf(pitch);
  // ref_lin_vel = Ryaw Rpitch input[:3]
  const float v1[3] = {cp * in[0] + sp * in[2], in[1], -sp * in[0] + cp * in[2]};
  const float rlv[3] = {cy * v1[0] - sy * v1[1], sy * v1[0] + cy * v1[1], v1[2]};
  const float nrm = sqrtf(rlv[0] * rlv[0] + rlv[1] * rlv[1] + rlv[2] * rlv[2]) + 1e-6f;
  const float v0x = -rlv[0] / nrm * a.clearence_speed, v0y = -rlv[1] / nrm * a.clearence_speed;
  // foothold of this leg (mpc_utils.py:92-96 / 203-207)
  const float f0[3] = {a.foot0[3 * leg], a.foot0[3 * leg + 1], a.foot0[3 * leg + 2]};
  const float f0px = p[0] + cy * f0[0] - sy * f0[1], f0py = p[1] + sy * f0[0] + cy * f0[1];
  const float kf = sqrtf(in[6] / 9.81f);
  const float fhx = 0.5f * rlv[0] * a.duty_factor / a.step_freq + kf * (dp[0] - rlv[0]) + f0px;
  const float fhy = 0.5f * rlv[1] * a.duty_factor / a.step_freq + kf * (dp[1] - rlv[1]) + f0py;
  // early contact of this leg (mpc_utils.py:77-81)
  const float ct = rg_advance(timers[leg], a.dt, a.step_freq);
  const bool des = ct < a.duty_factor;
  const float meas = a.contact ? a.contact[size_t(e) * nc + leg] : 0.f;
  const bool early = (!des) && (meas == 1.f) && (ct > 0.5f + 0.5f * a.duty_factor);

  float* ref = a.reference + size_t(e) * (N + 1) * nref;
  float* par = a.parameter + size_t(e) * (N + 1) * npar;
  // leg-independent columns
  for (int t = leg; t <= N; t += nc) {
    float* r = ref + size_t(t) * nref;
    r[0] = float(t) * a.dt * rlv[0] + p[0];
    r[1] = float(t) * a.dt * rlv[1] + p[1];
    r[2] = height;
    for (int i = 0; i < 4; ++i) r[3 + i] = quat0[i];
    int o = 7;
    if (!a.srbd) { for (int i = 0; i < nj; ++i) r[7 + i] = a.q0[i]; o = 7 + nj; }
    for (int i = 0; i < 3; ++i) { r[o + i] = rlv[i]; r[o + 3 + i] = in[3 + i]; }
  }
  float lx = lo[3 * leg], ly = lo[3 * leg + 1], lz = lo[3 * leg + 2];
  float prev[3] = {foot[3 * leg], foot[3 * leg + 1], foot[3 * leg + 2]};
  bool prev_contact = false;
  const int o_foot = 13 + nj, o_con = a.srbd ? 13 : 13 + nj + 3 * nc, o_grf = 13 + nj + 4 * nc;
  for (int t = 0; t <= N; ++t) {
    float csum = 0.f;
    for (int c = 0; c < nc; ++c) {
      timers[c] = rg_advance(timers[c], a.dt, a.step_freq);
      csum += (timers[c] < a.duty_factor) ? 1.f : 0.f;
    }
    const float tm = timers[leg];
    const bool nc_ = tm < a.duty_factor;
    if (!nc_ && prev_contact) { lx = prev[0]; ly = prev[1]; lz = prev[2]; }
    const float val = (tm - a.duty_factor) / (1.f - a.duty_factor);
    const float val2 = val * val, val3 = val2 * val, val4 = val2 * val2;
    float s[3], sd[3];
    {
      const float a2 = 3.f * (fhx - lx) - 2.f * v0x, a3 = v0x - 2.f * (fhx - lx);
      s[0] = lx + v0x * val + a2 * val2 + a3 * val3;
      sd[0] = 2.f * a2 * val + 3.f * a3 * val2 + v0x;
    }
    {
      const float a2 = 3.f * (fhy - ly) - 2.f * v0y, a3 = v0y - 2.f * (fhy - ly);
      s[1] = ly + v0y * val + a2 * val2 + a3 * val3;
      sd[1] = 2.f * a2 * val + 3.f * a3 * val2 + v0y;
    }
    {
      const float s0 = 0.7f, sh = lz + a.step_height;
      const float ca = 16.f * sh - 8.f * lz - 8.f * lz - 2.f * s0;
      const float cb = 5.f * s0 + 14.f * lz + 18.f * lz - 32.f * sh;
      const float cc = 16.f * sh - 5.f * lz - 11.f * lz - 4.f * s0;
      s[2] = ca * val4 + cb * val3 + cc * val2 + s0 * val + lz;
      sd[2] = 4.f * ca * val3 + 3.f * cb * val2 + 2.f * cc * val + s0;
    }
    const bool hold = nc_ || early;
    float cur[3];
    for (int i = 0; i < 3; ++i) cur[i] = hold ? prev[i] : s[i];
    const float ncf = nc_ ? 1.f : 0.f;
    if (a.srbd) {
      ref[size_t(t) * nref + o_con + leg] = ncf;
      par[size_t(t) * npar + leg] = ncf;
      for (int i = 0; i < 3; ++i) par[size_t(t) * npar + nc + 3 * leg + i] = cur[i];
    } else {
      float* r = ref + size_t(t) * nref;
      for (int i = 0; i < 3; ++i) r[o_foot + 3 * leg + i] = cur[i];
      r[o_con + leg] = ncf;
      r[o_grf + 3 * leg] = 0.f;
      r[o_grf + 3 * leg + 1] = 0.f;
      r[o_grf + 3 * leg + 2] = ncf * a.mass * 9.81f / (csum + 1e-5f);
      par[size_t(t) * npar + leg] = ncf;
    }
    if (a.foot_ref_dot)
      for (int i = 0; i < 3; ++i) a.foot_ref_dot[(size_t(e) * (N + 1) + t) * 3 * nc + 3 * leg + i] = hold ? 0.f : sd[i];
    for (int i = 0; i < 3; ++i) prev[i] = cur[i];
    prev_contact = nc_;
  }
  float* lout = a.liftoff + size_t(e) * 3 * nc;
  lout[3 * leg] = lx; lout[3 * leg + 1] = ly; lout[3 * leg + 2] = lz;
}

inline int launch_refgen(const RefGenArgs& a, cudaStream_t st) {
  constexpr int EPB = 16;
  const int blocks = (a.E + EPB - 1) / EPB;
  refgen_kernel<EPB><<<blocks, EPB * a.nc, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

}  // namespace mpxb
