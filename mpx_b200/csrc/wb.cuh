// wb.cuh -- whole-body (MuJoCo-style kinematic tree) legged-robot model: discrete dynamics, cost and residual
// maps, templated on the scalar type (float or Dual) exactly like srbd.cuh.
//
// Follows reference mpx/utils/models.py:39-117 (quadruped_wb_dynamics), 119-169 (h1_wb_dynamics), 170-217
// (talos_wb_dynamics) and mpx/utils/objectives.py:77-450.  The MJX calls those functions make
// (mjx.fwd_position / fwd_velocity / jac; mujoco-mjx==3.3.0 is a third-party dependency absent from the
// reference tree) are restated from the published algorithm in the same COM-centred spatial-vector convention:
//   kinematics -> com_pos (subtree COM, cinert, cdof) -> crb (dense qM + armature) -> cholesky
//   com_vel -> rne (qfrc_bias) ; support.jac (foot Jacobians about the subtree COM).
// Spatial motion vectors are [angular, linear]; cinert is (xx,yy,zz,xy,xz,yz, m*off[3], m).
// Only free + hinge joints, at most one joint per body (true for aliengo, go2, h1, talos).
#pragma once
#include <string.h>

#include "../../include/mpxb.h"
#include "common.cuh"

namespace mpxb {

constexpr int WB_MAXB = 32, WB_MAXV = 32, WB_MAXC = 8;
constexpr int WB_JNT_FREE = 0, WB_JNT_HINGE = 3;

struct WbModel {
  int kind, nbody, nv, nj, nc, n, m, nref, npar, nw, swing_tracking, grf_in_u, grf_as_state, contact_solve;
  float dt, alpha, reg, gravity[3];
  int parent[WB_MAXB], dofadr[WB_MAXB], jtype[WB_MAXB];
  float body_pos[WB_MAXB][3], body_quat[WB_MAXB][4], ipos[WB_MAXB][3], iquat[WB_MAXB][4], mass[WB_MAXB],
      inertia[WB_MAXB][3];
  float jnt_pos[WB_MAXB][3], jnt_axis[WB_MAXB][3], qpos0[WB_MAXB];
  float armature[WB_MAXV];
  int dof_body[WB_MAXV], dof_parent[WB_MAXV];
  int contact_body[WB_MAXC];
  float contact_pos[WB_MAXC][3];
  uint32_t contact_dofmask[WB_MAXC];  // dofs whose body is the contact body or one of its ancestors
};

// mpxb_model_desc (host, include/mpxb.h) -> WbModel: derived dof tables and contact ancestor masks
inline int wb_build_model(const mpxb_model_desc& d, WbModel* o) {
  if (d.nbody < 2 || d.nbody > WB_MAXB || d.nv < 6 || d.nv > WB_MAXV || d.n_contact < 1 || d.n_contact > WB_MAXC)
    return MPXB_EUNSUPPORTED;
  WbModel& w = *o;
  memset(&w, 0, sizeof(w));
  w.kind = d.kind;
  w.nbody = d.nbody; w.nv = d.nv; w.nj = d.n_joints; w.nc = d.n_contact;
  w.n = d.n; w.m = d.m; w.nref = d.nref; w.npar = d.npar; w.nw = d.nw;
  w.swing_tracking = d.swing_tracking;
  w.grf_in_u = (d.kind == MPXB_KIND_TALOS_WB);
  w.grf_as_state = !w.grf_in_u;
  w.contact_solve = !w.grf_in_u;
  w.dt = d.dt; w.alpha = d.baumgarte_alpha; w.reg = d.contact_reg;
  if (w.nv != w.nj + 6) return MPXB_EINVAL;
  const int nf = 3 * w.nc;
  if (w.n != 13 + 2 * w.nj + nf + (w.grf_as_state ? nf : 0)) return MPXB_EINVAL;
  if (w.m != w.nj + (w.grf_in_u ? nf : 0)) return MPXB_EINVAL;
  if (w.nref != 13 + w.nj + 7 * w.nc || w.npar < w.nc || w.nw != 12 + 3 * w.nj + 6 * w.nc) return MPXB_EINVAL;
  for (int i = 0; i < 3; ++i) w.gravity[i] = d.gravity[i];
  int lastdof[WB_MAXB];
  for (int b = 0; b < w.nbody; ++b) {
    w.parent[b] = d.body_parentid[b];
    w.dofadr[b] = d.body_dofadr[b];
    w.jtype[b] = d.body_jnttype[b];
    if (b > 0 && (w.parent[b] < 0 || w.parent[b] >= b)) return MPXB_EINVAL;
    for (int i = 0; i < 3; ++i) {
      w.body_pos[b][i] = d.body_pos[b][i]; w.ipos[b][i] = d.body_ipos[b][i]; w.inertia[b][i] = d.body_inertia[b][i];
      w.jnt_pos[b][i] = d.jnt_pos[b][i]; w.jnt_axis[b][i] = d.jnt_axis[b][i];
    }
    for (int i = 0; i < 4; ++i) { w.body_quat[b][i] = d.body_quat[b][i]; w.iquat[b][i] = d.body_iquat[b][i]; }
    w.mass[b] = d.body_mass[b];
    w.qpos0[b] = d.jnt_qpos0[b];
    const int pl = (b == 0) ? -1 : lastdof[w.parent[b]];
    if (w.jtype[b] == WB_JNT_FREE) {
      if (b != 1 || w.dofadr[b] != 0) return MPXB_EUNSUPPORTED;  // one floating base, first in the tree
      for (int k = 0; k < 6; ++k) { w.dof_body[k] = b; w.dof_parent[k] = (k == 0) ? pl : k - 1; }
      lastdof[b] = 5;
    } else if (w.jtype[b] == WB_JNT_HINGE) {
      const int dd = w.dofadr[b];
      if (dd < 6 || dd >= w.nv) return MPXB_EINVAL;
      w.dof_body[dd] = b;
      w.dof_parent[dd] = pl;
      lastdof[b] = dd;
    } else {
      w.jtype[b] = -1;
      lastdof[b] = pl;
    }
  }
  for (int i = 0; i < w.nv; ++i) w.armature[i] = d.dof_armature[i];
  for (int c = 0; c < w.nc; ++c) {
    w.contact_body[c] = d.contact_body[c];
    if (w.contact_body[c] <= 0 || w.contact_body[c] >= w.nbody) return MPXB_EINVAL;
    for (int i = 0; i < 3; ++i) w.contact_pos[c][i] = d.contact_pos[c][i];
    uint32_t mask = 0;
    for (int b = w.contact_body[c]; b > 0; b = w.parent[b])
      for (int k = 0; k < w.nv; ++k)
        if (w.dof_body[k] == b) mask |= (1u << k);
    w.contact_dofmask[c] = mask;
  }
  return MPXB_OK;
}

// ---------------------------------------------------------------- spatial algebra (mjx math)
template <class TI, class TV, class TO> HD void inert_mul(const TI* I, const TV* v, TO* o) {
  o[0] = I[0] * v[0] + I[3] * v[1] + I[4] * v[2] + (I[7] * v[5] - I[8] * v[4]);
  o[1] = I[3] * v[0] + I[1] * v[1] + I[5] * v[2] + (I[8] * v[3] - I[6] * v[5]);
  o[2] = I[4] * v[0] + I[5] * v[1] + I[2] * v[2] + (I[6] * v[4] - I[7] * v[3]);
  o[3] = I[9] * v[3] - (I[7] * v[2] - I[8] * v[1]);
  o[4] = I[9] * v[4] - (I[8] * v[0] - I[6] * v[2]);
  o[5] = I[9] * v[5] - (I[6] * v[1] - I[7] * v[0]);
}
template <class TU, class TV, class TO> HD void motion_cross(const TU* u, const TV* v, TO* o) {
  TO a0 = u[1] * v[2] - u[2] * v[1], a1 = u[2] * v[0] - u[0] * v[2], a2 = u[0] * v[1] - u[1] * v[0];
  TO l0 = (u[4] * v[2] - u[5] * v[1]) + (u[1] * v[5] - u[2] * v[4]);
  TO l1 = (u[5] * v[0] - u[3] * v[2]) + (u[2] * v[3] - u[0] * v[5]);
  TO l2 = (u[3] * v[1] - u[4] * v[0]) + (u[0] * v[4] - u[1] * v[3]);
  o[0] = a0; o[1] = a1; o[2] = a2; o[3] = l0; o[4] = l1; o[5] = l2;
}
template <class TV, class TF, class TO> HD void motion_cross_force(const TV* v, const TF* f, TO* o) {
  TO a0 = (v[1] * f[2] - v[2] * f[1]) + (v[4] * f[5] - v[5] * f[4]);
  TO a1 = (v[2] * f[0] - v[0] * f[2]) + (v[5] * f[3] - v[3] * f[5]);
  TO a2 = (v[0] * f[1] - v[1] * f[0]) + (v[3] * f[4] - v[4] * f[3]);
  TO l0 = v[1] * f[5] - v[2] * f[4], l1 = v[2] * f[3] - v[0] * f[5], l2 = v[0] * f[4] - v[1] * f[3];
  o[0] = a0; o[1] = a1; o[2] = a2; o[3] = l0; o[4] = l1; o[5] = l2;
}

// In-place Cholesky A = L L^T of the leading k x k block (row-major, leading dimension ld); only the lower
// triangle is read and written.  A non-SPD input yields NaN like jax.scipy.linalg.cho_factor.
template <class T> HD void chol_lower(T* A, int k, int ld) {
  for (int j = 0; j < k; ++j) {
    T d = A[j * ld + j];
    for (int p = 0; p < j; ++p) d = d - A[j * ld + p] * A[j * ld + p];
    d = xsqrt(d);
    A[j * ld + j] = d;
    T inv = 1.f / d;
    for (int i = j + 1; i < k; ++i) {
      T s = A[i * ld + j];
      for (int p = 0; p < j; ++p) s = s - A[i * ld + p] * A[j * ld + p];
      A[i * ld + j] = s * inv;
    }
  }
}
// x <- (L L^T)^-1 x for a strided vector x[i * inc]
template <class T, class TX> HD void chol_solve(const T* L, int k, int ld, TX* x, int inc) {
  for (int i = 0; i < k; ++i) {
    TX s = x[i * inc];
    for (int p = 0; p < i; ++p) s = s - L[i * ld + p] * x[p * inc];
    x[i * inc] = s / L[i * ld + i];
  }
  for (int i = k - 1; i >= 0; --i) {
    TX s = x[i * inc];
    for (int p = i + 1; p < k; ++p) s = s - L[p * ld + i] * x[p * inc];
    x[i * inc] = s / L[i * ld + i];
  }
}

// ---------------------------------------------------------------- geometry: kinematics + com_pos + feet
// Outputs: cinert[nbody][10], cdof[nv][6] (both about the subtree COM of the floating base), com[3],
// feet[3*nc] (geom_xpos of the contact spheres).  Scratch: xpos[NB][3], xquat[NB][4].
template <class T, int NB, int NV, int NC>
HD void wb_geometry(const WbModel& md, const T* qpos, T (*xpos)[3], T (*xquat)[4], T (*cinert)[10], T (*cdof)[6],
                    T* com, T* feet) {
  const int nb = md.nbody;
  xpos[0][0] = xpos[0][1] = xpos[0][2] = T(0.f);
  xquat[0][0] = T(1.f); xquat[0][1] = xquat[0][2] = xquat[0][3] = T(0.f);
  // mjx smooth.kinematics; joint axis -> cdof[.][0:3], joint anchor -> cdof[.][3:6] until the COM is known
  for (int b = 1; b < nb; ++b) {
    const int p = md.parent[b];
    T pos[3], quat[4], r[3];
    quat_rotate(xquat[p], md.body_pos[b], r);
    for (int i = 0; i < 3; ++i) pos[i] = xpos[p][i] + r[i];
    quat_mul(xquat[p], md.body_quat[b], quat);
    const int jt = md.jtype[b];
    if (jt == WB_JNT_FREE) {
      const int d = md.dofadr[b];
      for (int i = 0; i < 3; ++i) pos[i] = qpos[i];
      normalize_with_norm<T, 4>(qpos + 3, quat);
      T R[9];
      quat_to_mat(quat, R);
      for (int k = 0; k < 3; ++k) {
        for (int i = 0; i < 3; ++i) {
          cdof[d + k][i] = T(0.f);
          cdof[d + k][3 + i] = T(i == k ? 1.f : 0.f);
          cdof[d + 3 + k][i] = R[3 * i + k];  // column k of xmat
          cdof[d + 3 + k][3 + i] = pos[i];    // anchor (shifted to the COM below)
        }
      }
    } else if (jt == WB_JNT_HINGE) {
      const int d = md.dofadr[b];
      T anchor[3], axis[3];
      quat_rotate(quat, md.jnt_pos[b], r);
      for (int i = 0; i < 3; ++i) anchor[i] = r[i] + pos[i];
      quat_rotate(quat, md.jnt_axis[b], axis);
      const T half = (qpos[d + 1] - md.qpos0[b]) * 0.5f;  // qposadr = dofadr + 1 behind one free joint
      const T s = xsin(half), c = xcos(half);
      T qloc[4] = {c, md.jnt_axis[b][0] * s, md.jnt_axis[b][1] * s, md.jnt_axis[b][2] * s};
      T q2[4];
      quat_mul(quat, qloc, q2);
      for (int i = 0; i < 4; ++i) quat[i] = q2[i];
      quat_rotate(quat, md.jnt_pos[b], r);
      for (int i = 0; i < 3; ++i) {
        pos[i] = anchor[i] - r[i];
        cdof[d][i] = axis[i];
        cdof[d][3 + i] = anchor[i];
      }
    }
    for (int i = 0; i < 3; ++i) xpos[b][i] = pos[i];
    for (int i = 0; i < 4; ++i) xquat[b][i] = quat[i];
  }
  // feet = geom_xpos of the contact spheres
  for (int c = 0; c < md.nc; ++c) {
    const int b = md.contact_body[c];
    T r[3];
    quat_rotate(xquat[b], md.contact_pos[c], r);
    for (int i = 0; i < 3; ++i) feet[3 * c + i] = xpos[b][i] + r[i];
  }
  // subtree COM of the root (every non-world body hangs off the single floating base)
  T msum[3] = {T(0.f), T(0.f), T(0.f)};
  float mtot = 0.f;
  for (int b = 1; b < nb; ++b) {
    T r[3];
    quat_rotate(xquat[b], md.ipos[b], r);
    for (int i = 0; i < 3; ++i) msum[i] = msum[i] + (xpos[b][i] + r[i]) * md.mass[b];
    mtot += md.mass[b];
  }
  for (int i = 0; i < 3; ++i) com[i] = msum[i] / mtot;
  // cinert about the COM
  for (int b = 1; b < nb; ++b) {
    T r[3], off[3], qi[4], R[9];
    quat_rotate(xquat[b], md.ipos[b], r);
    for (int i = 0; i < 3; ++i) off[i] = xpos[b][i] + r[i] - com[i];
    quat_mul(xquat[b], md.iquat[b], qi);
    quat_to_mat(qi, R);
    const float mb = md.mass[b];
    const float* I = md.inertia[b];
    const T o2 = off[0] * off[0] + off[1] * off[1] + off[2] * off[2];
    const int ia[6] = {0, 1, 2, 0, 0, 1}, ib[6] = {0, 1, 2, 1, 2, 2};
    for (int k = 0; k < 6; ++k) {
      const int a = ia[k], c = ib[k];
      T v = R[3 * a] * (R[3 * c] * I[0]) + R[3 * a + 1] * (R[3 * c + 1] * I[1]) + R[3 * a + 2] * (R[3 * c + 2] * I[2]);
      v = v - off[a] * off[c] * mb;
      if (a == c) v = v + o2 * mb;
      cinert[b][k] = v;
    }
    for (int i = 0; i < 3; ++i) cinert[b][6 + i] = off[i] * mb;
    cinert[b][9] = T(mb);
  }
  for (int k = 0; k < 10; ++k) cinert[0][k] = T(0.f);
  // cdof linear part: axis x (com - anchor)
  for (int d = 0; d < md.nv; ++d) {
    const int b = md.dof_body[d];
    if (md.jtype[b] == WB_JNT_FREE && d - md.dofadr[b] < 3) continue;
    T o[3], l[3];
    for (int i = 0; i < 3; ++i) o[i] = com[i] - cdof[d][3 + i];
    cross3(cdof[d], o, l);
    for (int i = 0; i < 3; ++i) cdof[d][3 + i] = l[i];
  }
}

// mjx smooth.crb + dense make_m: M[nv][ld] (full symmetric) ; crb is scratch [NB][10]
template <class T, int NB, int NV>
HD void wb_mass_matrix(const WbModel& md, const T (*cinert)[10], const T (*cdof)[6], T (*crb)[10], T* M, int ld) {
  const int nb = md.nbody, nv = md.nv;
  for (int b = 0; b < nb; ++b)
    for (int k = 0; k < 10; ++k) crb[b][k] = cinert[b][k];
  for (int b = nb - 1; b > 0; --b) {
    const int p = md.parent[b];
    if (p > 0)
      for (int k = 0; k < 10; ++k) crb[p][k] = crb[p][k] + crb[b][k];
  }
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) M[i * ld + j] = T(0.f);
  for (int i = 0; i < nv; ++i) {
    T f[6];
    inert_mul(crb[md.dof_body[i]], cdof[i], f);
    int j = i;
    while (j >= 0) {
      T s = f[0] * cdof[j][0];
      for (int k = 1; k < 6; ++k) s = s + f[k] * cdof[j][k];
      M[i * ld + j] = s;
      M[j * ld + i] = s;
      j = md.dof_parent[j];
    }
    M[i * ld + i] = M[i * ld + i] + md.armature[i];
  }
}

// mjx smooth.com_vel + rne: qfrc_bias[nv].  Scratch: cvel, cacc, cfrc [NB][6].
template <class TG, class TV, class T, int NB, int NV>
HD void wb_bias(const WbModel& md, const TG (*cinert)[10], const TG (*cdof)[6], const TV* qvel, T (*cvel)[6],
                T (*cacc)[6], T (*cfrc)[6], T* bias) {
  const int nb = md.nbody, nv = md.nv;
  for (int k = 0; k < 6; ++k) { cvel[0][k] = T(0.f); cfrc[0][k] = T(0.f); }
  for (int k = 0; k < 3; ++k) { cacc[0][k] = T(0.f); cacc[0][3 + k] = T(-md.gravity[k]); }
  for (int b = 1; b < nb; ++b) {
    const int p = md.parent[b];
    T cv[6], ca[6];
    for (int k = 0; k < 6; ++k) { cv[k] = cvel[p][k]; ca[k] = cacc[p][k]; }
    const int jt = md.jtype[b], d = md.dofadr[b];
    if (jt == WB_JNT_FREE) {
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 6; ++k) cv[k] = cv[k] + cdof[d + a][k] * qvel[d + a];
      T cd[3][6];
      for (int a = 0; a < 3; ++a) motion_cross(cv, cdof[d + 3 + a], cd[a]);
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 6; ++k) {
          cv[k] = cv[k] + cdof[d + 3 + a][k] * qvel[d + 3 + a];
          ca[k] = ca[k] + cd[a][k] * qvel[d + 3 + a];
        }
    } else if (jt == WB_JNT_HINGE) {
      T cd[6];
      motion_cross(cv, cdof[d], cd);
      for (int k = 0; k < 6; ++k) {
        ca[k] = ca[k] + cd[k] * qvel[d];
        cv[k] = cv[k] + cdof[d][k] * qvel[d];
      }
    }
    for (int k = 0; k < 6; ++k) { cvel[b][k] = cv[k]; cacc[b][k] = ca[k]; }
    T f1[6], f2[6], f3[6];
    inert_mul(cinert[b], ca, f1);
    inert_mul(cinert[b], cv, f2);
    motion_cross_force(cv, f2, f3);
    for (int k = 0; k < 6; ++k) cfrc[b][k] = f1[k] + f3[k];
  }
  for (int b = nb - 1; b > 0; --b) {
    const int p = md.parent[b];
    if (p > 0)
      for (int k = 0; k < 6; ++k) cfrc[p][k] = cfrc[p][k] + cfrc[b][k];
  }
  for (int i = 0; i < nv; ++i) {
    const int b = md.dof_body[i];
    T s = cdof[i][0] * cfrc[b][0];
    for (int k = 1; k < 6; ++k) s = s + cdof[i][k] * cfrc[b][k];
    bias[i] = s;
  }
}

// mjx support.jac translational part for every contact point: J[nv][ld], columns 3c..3c+2
template <class T, int NV, int NC>
HD void wb_contact_jacobian(const WbModel& md, const T (*cdof)[6], const T* com, const T* feet, T* J, int ld) {
  for (int c = 0; c < md.nc; ++c) {
    T off[3];
    for (int i = 0; i < 3; ++i) off[i] = feet[3 * c + i] - com[i];
    const uint32_t mask = md.contact_dofmask[c];
    for (int d = 0; d < md.nv; ++d) {
      if ((mask >> d) & 1u) {
        T cr[3];
        cross3(cdof[d], off, cr);
        for (int i = 0; i < 3; ++i) J[d * ld + 3 * c + i] = cdof[d][3 + i] + cr[i];
      } else {
        for (int i = 0; i < 3; ++i) J[d * ld + 3 * c + i] = T(0.f);
      }
    }
  }
}

// ---------------------------------------------------------------- discrete dynamics x+ = f(x, u)
// x = [p3 | quat4 | q nj | v3 | omega3 | dq nj | foot 3nc | grf 3nc (quadruped / H1)], u = [tau nj | grf 3nc (Talos)]
template <class T, int NB, int NV, int NC>
HD void wb_dynamics(const WbModel& md, const T* x, const T* u, const float* contact, T* xn) {
  constexpr int NF = 3 * NC;
  const int nv = md.nv, nj = md.nj, nc = md.nc, nf = 3 * nc;
  const T* qpos = x;
  const T* qvel = x + 7 + nj;
  T cinert[NB][10], cdof[NV][6], com[3], feet[NF];
  T M[NV * NV], bias[NV], J[NV * NF];
  {
    T xpos[NB][3], xquat[NB][4];
    wb_geometry<T, NB, NV, NC>(md, qpos, xpos, xquat, cinert, cdof, com, feet);
  }
  {
    T crb[NB][10];
    wb_mass_matrix<T, NB, NV>(md, cinert, cdof, crb, M, NV);
  }
  {
    T cvel[NB][6], cacc[NB][6], cfrc[NB][6];
    wb_bias<T, T, T, NB, NV>(md, cinert, cdof, qvel, cvel, cacc, cfrc, bias);
  }
  wb_contact_jacobian<T, NV, NC>(md, cdof, com, feet, J, NF);
  chol_lower(M, nv, NV);  // qLD (mjx factor_m, dense path)
  T grf[NF], rhsv[NV];
  for (int i = 0; i < nv; ++i) rhsv[i] = (i < 6 ? T(0.f) : u[i - 6]) - bias[i];  // tau - D
  if (md.contact_solve) {
    // models.py:93-107 / 146-160: grf = chol(J^T M^-1 J (+ reg I))^-1 (-J^T M^-1 (tau - D) - 2 alpha J^T qvel)
    T MiJ[NV * NF], JMJ[NF * NF], w[NV];
    for (int i = 0; i < nv; ++i) {
      w[i] = rhsv[i];
      for (int c = 0; c < nf; ++c) MiJ[i * NF + c] = J[i * NF + c];
    }
    chol_solve(M, nv, NV, w, 1);
    for (int c = 0; c < nf; ++c) chol_solve(M, nv, NV, MiJ + c, NF);
    for (int a = 0; a < nf; ++a) {
      for (int b = 0; b <= a; ++b) {
        T s = J[a] * MiJ[b];
        for (int i = 1; i < nv; ++i) s = s + J[i * NF + a] * MiJ[i * NF + b];
        if (a == b) s = s + md.reg;
        JMJ[a * NF + b] = s;
      }
      T s = J[a] * (w[0] + (2.f * md.alpha) * qvel[0]);
      for (int i = 1; i < nv; ++i) s = s + J[i * NF + a] * (w[i] + (2.f * md.alpha) * qvel[i]);
      grf[a] = -s;
    }
    chol_lower(JMJ, nf, NF);
    chol_solve(JMJ, nf, NF, grf, 1);
    for (int c = 0; c < nc; ++c)
      for (int i = 0; i < 3; ++i) grf[3 * c + i] = grf[3 * c + i] * contact[c];
  } else {
    // models.py:206-209: the GRFs are controls
    for (int c = 0; c < nc; ++c)
      for (int i = 0; i < 3; ++i) grf[3 * c + i] = u[nj + 3 * c + i] * contact[c];
  }
  // v+ = qvel + M^-1 (tau - D + J grf) dt
  for (int i = 0; i < nv; ++i) {
    T s = rhsv[i];
    for (int c = 0; c < nf; ++c) s = s + J[i * NF + c] * grf[c];
    rhsv[i] = s;
  }
  chol_solve(M, nv, NV, rhsv, 1);
  const float dt = md.dt;
  T* vn = xn + 7 + nj;
  for (int i = 0; i < nv; ++i) vn[i] = qvel[i] + rhsv[i] * dt;
  for (int i = 0; i < 3; ++i) xn[i] = x[i] + vn[i] * dt;
  quat_integrate(x + 3, vn + 3, dt, xn + 3);
  for (int i = 0; i < nj; ++i) xn[7 + i] = x[7 + i] + vn[6 + i] * dt;
  T* tail = xn + 7 + nj + nv;
  for (int i = 0; i < nf; ++i) tail[i] = feet[i];
  if (md.grf_as_state)
    for (int i = 0; i < nf; ++i) tail[nf + i] = grf[i];
}

// ---------------------------------------------------------------- cost (objectives.py:77-125, 204-250, 325-376)
template <class T> HD T wb_qform(const float* W, int ldw, int off, int k, const T* r) {
  T s(0.f);
  for (int i = 0; i < k; ++i) {
    T row(0.f);
    for (int j = 0; j < k; ++j) row = row + W[(off + i) * ldw + off + j] * r[j];
    s = s + r[i] * row;
  }
  return s;
}

// reference row layout: [p3 | quat4 | q nj | dp3 | omega3 | foot 3nc | contact nc | grf 3nc]
template <class T, int NJ, int NC>
HD T wb_cost(const WbModel& md, const float* W, const float* ref, const T* x, const T* u, bool terminal) {
  constexpr int NF = 3 * NC;
  const int nj = md.nj, nc = md.nc, nf = 3 * nc, nw = md.nw;
  const T* dq = x + 13 + nj;
  const T* p_leg = x + 13 + 2 * nj;
  const T* grf = md.grf_in_u ? u + nj : x + 13 + 2 * nj + nf;
  const float* contact = ref + 13 + nj + nf;
  const float* grf_ref = ref + 13 + nj + nf + nc;
  T r[(NJ > NF ? NJ : NF)];
  T cost(0.f);
  for (int i = 0; i < 3; ++i) r[i] = x[i] - ref[i];
  cost = cost + wb_qform(W, nw, 0, 3, r);
  quat_sub(x + 3, ref + 3, r);
  cost = cost + wb_qform(W, nw, 3, 3, r);
  for (int i = 0; i < nj; ++i) r[i] = x[7 + i] - ref[7 + i];
  cost = cost + wb_qform(W, nw, 6, nj, r);
  for (int i = 0; i < 3; ++i) r[i] = x[7 + nj + i] - ref[7 + nj + i];
  cost = cost + wb_qform(W, nw, 6 + nj, 3, r);
  for (int i = 0; i < 3; ++i) r[i] = x[10 + nj + i] - ref[10 + nj + i];
  cost = cost + wb_qform(W, nw, 9 + nj, 3, r);
  cost = cost + wb_qform(W, nw, 12 + nj, nj, dq);
  if (!terminal) {
    const bool masked = (md.kind == 1) && !md.swing_tracking;  // contact_map (objectives.py:110-113)
    for (int i = 0; i < nf; ++i) {
      r[i] = p_leg[i] - ref[13 + nj + i];
      if (masked) r[i] = r[i] * contact[i / 3];
    }
    cost = cost + wb_qform(W, nw, 12 + 2 * nj, nf, r);
    cost = cost + wb_qform(W, nw, 12 + 2 * nj + nf, nj, u);
    for (int i = 0; i < nf; ++i) r[i] = grf[i] - grf_ref[i];
    cost = cost + wb_qform(W, nw, 12 + 3 * nj + nf, nf, r);
    if (md.kind == 1) {
      // objectives.py:97-108,120: friction cone (mu 0.5, eps 1e-2, masked by contact), torque +-44, joint speed +-10
      for (int c = 0; c < nc; ++c) {
        T cone = 0.5f * grf[3 * c + 2] - xsqrt(grf[3 * c + 1] * grf[3 * c + 1] + grf[3 * c] * grf[3 * c] + 1e-2f);
        cost = cost + penalty(cone, 0.1f, 5.f) * contact[c];
      }
      for (int i = 0; i < nj; ++i) {
        cost = cost + penalty(44.01f - u[i], 1.f, 1.f) + penalty(u[i] + 44.01f, 1.f, 1.f);
        cost = cost + penalty(10.01f - dq[i], 1.f, 1.f) + penalty(dq[i] + 10.01f, 1.f, 1.f);
      }
    }
  }
  return 0.5f * cost;
}

}  // namespace mpxb
