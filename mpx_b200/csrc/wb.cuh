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
constexpr int WB_MAXLEAF = 12, WB_MAXPATH = 12, WB_MAXDIR = 4 + 2 * WB_MAXV, WB_MAXITEMS = 1024;
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
  int depth[WB_MAXB];                 // base = 0 (bodies are numbered depth-first, parents before children)
  int maxdepth;
  // tables of the analytic linearisation (wb_tree.cuh).  Bodies are numbered depth-first, so the subtree of body b
  // is the index range [b, b + sub_size[b]).  Directions: 0..3 raw base-quaternion components, 4..4+nj-1 hinge
  // positions (dof D+2), then the nv velocities; items enumerate (direction, body in the direction's subtree).
  int tree_ok;                        // 0: tables exceed the static bounds, use the forward-mode path
  int sub_size[WB_MAXB];
  uint32_t dof_anc[WB_MAXV];          // ancestor dofs of dof i, including i
  int nleaf, leaf_len[WB_MAXLEAF], leaf_path[WB_MAXLEAF][WB_MAXPATH];
  int ndir, nitems;
  uint16_t dir_off[WB_MAXDIR];
  uint8_t dir_b0[WB_MAXDIR], dir_nb[WB_MAXDIR];
  uint8_t item_dir[WB_MAXITEMS];
  // FK shortcuts, exact: bit 0 = body_quat is the identity, bit 1 = jnt_pos is zero (true for most MJCF bodies; the
  // products they save are multiplications by exact ones and zeros)
  uint8_t body_flags[WB_MAXB];
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
    w.body_flags[b] = uint8_t(((w.body_quat[b][0] == 1.f && w.body_quat[b][1] == 0.f && w.body_quat[b][2] == 0.f && w.body_quat[b][3] == 0.f) ? 1 : 0) |
                              ((w.jnt_pos[b][0] == 0.f && w.jnt_pos[b][1] == 0.f && w.jnt_pos[b][2] == 0.f) ? 2 : 0));
    w.depth[b] = (b <= 1) ? 0 : w.depth[w.parent[b]] + 1;
    if (b > 1 && w.parent[b] < 1) return MPXB_EUNSUPPORTED;  // a single tree hanging off the floating base
    if (w.depth[b] + 1 > w.maxdepth) w.maxdepth = w.depth[b] + 1;
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
  // ---- analytic-linearisation tables
  w.tree_ok = 1;
  for (int b = w.nbody - 1; b >= 1; --b) {
    w.sub_size[b] += 1;
    if (b > 1) w.sub_size[w.parent[b]] += w.sub_size[b];
  }
  for (int b = 2; b < w.nbody; ++b) {  // depth-first numbering <=> every subtree is a contiguous index range
    const int p = w.parent[b];
    if (!(b > p && b < p + w.sub_size[p])) w.tree_ok = 0;
  }
  for (int i = 0; i < w.nv; ++i) {
    uint32_t mk = 0;
    for (int j = i; j >= 0; j = w.dof_parent[j]) mk |= (1u << j);
    w.dof_anc[i] = mk;
  }
  w.nleaf = 0;
  for (int b = 1; b < w.nbody; ++b) {
    bool leaf = true;
    for (int c = b + 1; c < w.nbody; ++c) leaf = leaf && (w.parent[c] != b);
    if (!leaf) continue;
    if (w.nleaf >= WB_MAXLEAF || w.depth[b] + 1 > WB_MAXPATH) { w.tree_ok = 0; break; }
    const int len = w.depth[b] + 1;
    w.leaf_len[w.nleaf] = len;
    int a = b;
    for (int l = len - 1; l >= 0; --l) { w.leaf_path[w.nleaf][l] = a; a = w.parent[a]; }
    ++w.nleaf;
  }
  w.ndir = 4 + w.nj + w.nv;
  w.nitems = 0;
  for (int D = 0; D < w.ndir && w.tree_ok; ++D) {
    int b0 = 1;
    if (D >= 4 && D < 4 + w.nj) b0 = w.dof_body[D + 2];
    else if (D >= 4 + w.nj) b0 = w.dof_body[D - 4 - w.nj];
    w.dir_b0[D] = uint8_t(b0);
    w.dir_nb[D] = uint8_t(w.sub_size[b0]);
    w.dir_off[D] = uint16_t(w.nitems);
    if (w.nitems + w.sub_size[b0] > WB_MAXITEMS) { w.tree_ok = 0; break; }
    for (int k = 0; k < w.sub_size[b0]; ++k) w.item_dir[w.nitems++] = uint8_t(D);
  }
  return MPXB_OK;
}

// Work-item table of the folded-tangent stage (wb_tree_fold_coop, F2): the (x-direction, dof) items sorted by the branch
// they take -- dof moves with the direction | dof is an ancestor | velocity direction | structurally zero -- each
// segment padded to a multiple of 32 with WB_FOLD_NONE, so every warp runs one branch (the straight (col, dof)
// enumeration mixes all four in a warp: 14 of 32 lanes active).  Lives in device memory right behind the WbModel.
constexpr int WB_MAXFOLD = (WB_MAXV + 1 + WB_MAXV) * WB_MAXV + 4 * 32;
constexpr uint16_t WB_FOLD_NONE = 0xFFFFu;
struct WbFoldTable {
  int end_mv, end_anc, end_vel, end_zero;  // segment ends (multiples of 32); item = col << 5 | dof
  uint16_t item[WB_MAXFOLD];
};
inline void wb_build_fold_table(const WbModel& w, WbFoldTable* ft) {
  memset(ft, 0xFF, sizeof(*ft));
  const int nq = w.nv + 1, ncol = nq + w.nv;
  int n = 0;
  auto type_of = [&](int col, int i) {
    if (col < 3) return 3;
    const int D = col - 3, b0 = w.dir_b0[D], nbd = w.dir_nb[D], bi = w.dof_body[i];
    if (col >= nq) return ((bi >= b0 && bi < b0 + nbd) || ((w.dof_anc[col - nq] >> i) & 1u)) ? 2 : 3;
    const bool isq = D < 4;
    if (isq ? (i >= 3) : (bi >= b0 && bi < b0 + nbd)) return 0;
    return (((isq ? 7u : w.dof_anc[D + 2]) >> i) & 1u) ? 1 : 3;
  };
  int ends[4];
  for (int ty = 0; ty < 4; ++ty) {
    for (int col = 0; col < ncol; ++col)
      for (int i = 0; i < w.nv; ++i)
        if (type_of(col, i) == ty) ft->item[n++] = uint16_t(col << 5 | i);
    while (n & 31) ft->item[n++] = WB_FOLD_NONE;
    ends[ty] = n;
  }
  ft->end_mv = ends[0]; ft->end_anc = ends[1]; ft->end_vel = ends[2]; ft->end_zero = ends[3];
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

// ====================================================================================================
// Single-pass depth-first evaluation of the tree quantities with a path stack (v2 of the linearisation).
//
// The bodies are visited in index order (depth-first pre-order); slot d of the stack holds the current ancestor
// at depth d: pose, joint motion axis, spatial velocity / bias acceleration, and the composite inertia / force
// accumulated from its finished descendants.  Mass-matrix entries, bias forces, foot positions and foot
// Jacobians are handed to a `sink` the moment they are final, so a forward-mode (Dual) pass can fold their
// tangents into a handful of matrix-vector products without ever storing dM/dJ per direction.
// Reference point of all spatial quantities: the base position qpos[0:3] (mjx uses the subtree COM; qM,
// qfrc_bias, the contact Jacobians and geom_xpos do not depend on that choice).
namespace mpxb {

constexpr int WB_SLOT = 41;
enum { WS_POS = 0, WS_QUAT = 3, WS_CDOF = 7, WS_CVEL = 13, WS_CACC = 19, WS_CRB = 25, WS_CFRC = 35 };

template <class T> struct WbStackPlain {  // host / single-thread storage
  T* p;
  HD T ld(int i) const { return p[i]; }
  HD void st(int i, T v) const { p[i] = v; }
};

// cdof of dof j, which lives on the body in stack slot sd
template <class T, class Stk>
HD void wb_slot_cdof(const WbModel& md, const Stk& stk, int sd, int sb, int j, const T* Rb, T* cd) {
  if (md.jtype[sb] == WB_JNT_FREE) {
    const int k = j - md.dofadr[sb];
    for (int i = 0; i < 3; ++i) {
      cd[i] = (k < 3) ? T(0.f) : Rb[3 * i + (k - 3)];
      cd[3 + i] = T((k < 3 && i == k) ? 1.f : 0.f);
    }
  } else {
    for (int i = 0; i < 6; ++i) cd[i] = stk.ld(sd * WB_SLOT + WS_CDOF + i);
  }
}

template <class T, class Stk, class Sink>
HD void wb_dfs_finish(const WbModel& md, const Stk& stk, const int* slot_body, int dd, const T* Rb, Sink& sink) {
  const int b = slot_body[dd];
  T crb[10], cfrc[6];
  for (int i = 0; i < 10; ++i) crb[i] = stk.ld(dd * WB_SLOT + WS_CRB + i);
  for (int i = 0; i < 6; ++i) cfrc[i] = stk.ld(dd * WB_SLOT + WS_CFRC + i);
  const int jt = md.jtype[b];
  const int nd = (jt == WB_JNT_FREE) ? 6 : (jt == WB_JNT_HINGE ? 1 : 0);
  for (int a = nd - 1; a >= 0; --a) {
    const int i = md.dofadr[b] + a;
    T cdi[6], f[6];
    wb_slot_cdof(md, stk, dd, b, i, Rb, cdi);
    inert_mul(crb, cdi, f);
    T bs = cdi[0] * cfrc[0];
    for (int k = 1; k < 6; ++k) bs = bs + cdi[k] * cfrc[k];
    sink.bias(i, bs);
    for (int sd = dd; sd >= 0; --sd) {
      const int sb = slot_body[sd];
      const int jts = md.jtype[sb];
      if (jts != WB_JNT_FREE && jts != WB_JNT_HINGE) continue;
      const int j0 = md.dofadr[sb];
      const int jhi = (jts == WB_JNT_FREE) ? ((sd == dd) ? i : j0 + 5) : j0;
      for (int j = jhi; j >= j0; --j) {
        T cdj[6];
        if (j == i) { for (int k = 0; k < 6; ++k) cdj[k] = cdi[k]; }
        else wb_slot_cdof(md, stk, sd, sb, j, Rb, cdj);
        T mij = f[0] * cdj[0];
        for (int k = 1; k < 6; ++k) mij = mij + f[k] * cdj[k];
        if (j == i) mij = mij + md.armature[i];
        sink.mass(i, j, mij);
      }
    }
  }
  if (dd > 0) {
    const int pd = dd - 1;
    for (int i = 0; i < 10; ++i) stk.st(pd * WB_SLOT + WS_CRB + i, stk.ld(pd * WB_SLOT + WS_CRB + i) + crb[i]);
    for (int i = 0; i < 6; ++i) stk.st(pd * WB_SLOT + WS_CFRC + i, stk.ld(pd * WB_SLOT + WS_CFRC + i) + cfrc[i]);
  }
}

// TV: scalar type of the velocity-dependent quantities (== T, or float with T = float)
template <class T> struct WbInPlain {  // (qpos, qvel) from plain arrays
  const T* qp; const T* qv;
  HD T q(int i) const { return qp[i]; }
  HD T v(int i) const { return qv[i]; }
};
struct WbInSeed {  // forward-mode seed on entry `dir` of x = [qpos | qvel | ...]
  const float* x; int nq, dir;
  HD Dual q(int i) const { return Dual(x[i], i == dir ? 1.f : 0.f); }
  HD Dual v(int i) const { return Dual(x[nq + i], nq + i == dir ? 1.f : 0.f); }
};

template <class T, class In, class Stk, class Sink>
HD void wb_dfs(const WbModel& md, const In& in, const Stk& stk, Sink& sink) {
  int slot_body[WB_MAXB];
  T Rb[9];
  const T ref[3] = {in.q(0), in.q(1), in.q(2)};
  int cur = -1;
  const int nb = md.nbody;
  for (int b = 1; b < nb; ++b) {
    const int d = md.depth[b];
    for (int dd = cur; dd >= d; --dd) wb_dfs_finish(md, stk, slot_body, dd, Rb, sink);
    cur = d;
    slot_body[d] = b;
    const int so = d * WB_SLOT;
    T pos[3], quat[4], cdof[6], cv[6], ca[6];
    const int jt = md.jtype[b];
    if (d == 0) {
      for (int i = 0; i < 3; ++i) pos[i] = ref[i];
      {
        const T q4[4] = {in.q(3), in.q(4), in.q(5), in.q(6)};
        normalize_with_norm<T, 4>(q4, quat);
      }
      quat_to_mat(quat, Rb);
      // mjx com_vel for the free joint: translation first, angular cdof_dot against the translated velocity
      for (int i = 0; i < 3; ++i) { cv[i] = T(0.f); cv[3 + i] = in.v(i); ca[i] = T(0.f); ca[3 + i] = T(-md.gravity[i]); }
      T cd[3][6];
      for (int a = 0; a < 3; ++a) {
        T cdr[6] = {Rb[a], Rb[3 + a], Rb[6 + a], T(0.f), T(0.f), T(0.f)};
        motion_cross(cv, cdr, cd[a]);
      }
      for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 6; ++k) {
          if (k < 3) cv[k] = cv[k] + Rb[3 * k + a] * in.v(3 + a);
          ca[k] = ca[k] + cd[a][k] * in.v(3 + a);
        }
      for (int i = 0; i < 6; ++i) cdof[i] = T(0.f);
    } else {
      T ppos[3], pquat[4], r[3];
      for (int i = 0; i < 3; ++i) ppos[i] = stk.ld(so - WB_SLOT + WS_POS + i);
      for (int i = 0; i < 4; ++i) pquat[i] = stk.ld(so - WB_SLOT + WS_QUAT + i);
      for (int i = 0; i < 6; ++i) { cv[i] = stk.ld(so - WB_SLOT + WS_CVEL + i); ca[i] = stk.ld(so - WB_SLOT + WS_CACC + i); }
      quat_rotate(pquat, md.body_pos[b], r);
      for (int i = 0; i < 3; ++i) pos[i] = ppos[i] + r[i];
      quat_mul(pquat, md.body_quat[b], quat);
      if (jt == WB_JNT_HINGE) {
        const int dj = md.dofadr[b];
        T anchor[3], axis[3], o[3], l[3];
        quat_rotate(quat, md.jnt_pos[b], r);
        for (int i = 0; i < 3; ++i) anchor[i] = r[i] + pos[i];
        quat_rotate(quat, md.jnt_axis[b], axis);
        const T half = (in.q(dj + 1) - md.qpos0[b]) * 0.5f;
        const T sn = xsin(half), cs = xcos(half);
        T qloc[4] = {cs, md.jnt_axis[b][0] * sn, md.jnt_axis[b][1] * sn, md.jnt_axis[b][2] * sn};
        T q2[4];
        quat_mul(quat, qloc, q2);
        for (int i = 0; i < 4; ++i) quat[i] = q2[i];
        quat_rotate(quat, md.jnt_pos[b], r);
        for (int i = 0; i < 3; ++i) { pos[i] = anchor[i] - r[i]; o[i] = ref[i] - anchor[i]; }
        cross3(axis, o, l);
        for (int i = 0; i < 3; ++i) { cdof[i] = axis[i]; cdof[3 + i] = l[i]; }
        T cd[6];
        motion_cross(cv, cdof, cd);
        const T qd = in.v(dj);
        for (int k = 0; k < 6; ++k) { ca[k] = ca[k] + cd[k] * qd; cv[k] = cv[k] + cdof[k] * qd; }
      } else {
        for (int i = 0; i < 6; ++i) cdof[i] = T(0.f);
      }
    }
    for (int i = 0; i < 3; ++i) stk.st(so + WS_POS + i, pos[i]);
    for (int i = 0; i < 4; ++i) stk.st(so + WS_QUAT + i, quat[i]);
    for (int i = 0; i < 6; ++i) { stk.st(so + WS_CDOF + i, cdof[i]); stk.st(so + WS_CVEL + i, cv[i]); stk.st(so + WS_CACC + i, ca[i]); }
    // spatial inertia about the reference point, local force
    {
      T r[3], off[3], qi[4], R[9], ci[10];
      quat_rotate(quat, md.ipos[b], r);
      for (int i = 0; i < 3; ++i) off[i] = pos[i] + r[i] - ref[i];
      quat_mul(quat, md.iquat[b], qi);
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
        ci[k] = v;
      }
      for (int i = 0; i < 3; ++i) ci[6 + i] = off[i] * mb;
      ci[9] = T(mb);
      T f1[6], f2[6], f3[6];
      inert_mul(ci, ca, f1);
      inert_mul(ci, cv, f2);
      motion_cross_force(cv, f2, f3);
      for (int i = 0; i < 10; ++i) stk.st(so + WS_CRB + i, ci[i]);
      for (int i = 0; i < 6; ++i) stk.st(so + WS_CFRC + i, f1[i] + f3[i]);
    }
    // contact points carried by this body: position and translational Jacobian over the path
    for (int c = 0; c < md.nc; ++c) {
      if (md.contact_body[c] != b) continue;
      T r[3], foot[3], off[3];
      quat_rotate(quat, md.contact_pos[c], r);
      for (int i = 0; i < 3; ++i) { foot[i] = pos[i] + r[i]; off[i] = foot[i] - ref[i]; }
      sink.foot(c, foot);
      for (int sd = d; sd >= 0; --sd) {
        const int sb = slot_body[sd];
        const int jts = md.jtype[sb];
        if (jts != WB_JNT_FREE && jts != WB_JNT_HINGE) continue;
        const int j0 = md.dofadr[sb], j1 = j0 + (jts == WB_JNT_FREE ? 5 : 0);
        for (int j = j1; j >= j0; --j) {
          T cdj[6], cr[3], jc[3];
          if (sd == d && jts == WB_JNT_HINGE) { for (int k = 0; k < 6; ++k) cdj[k] = cdof[k]; }
          else wb_slot_cdof(md, stk, sd, sb, j, Rb, cdj);
          cross3(cdj, off, cr);
          for (int i = 0; i < 3; ++i) jc[i] = cdj[3 + i] + cr[i];
          sink.jac(j, c, jc);
        }
      }
    }
  }
  for (int dd = cur; dd >= 0; --dd) wb_dfs_finish(md, stk, slot_body, dd, Rb, sink);
}

// Sink that stores the primal terms: M [nv][ldm] (full symmetric; entries off the tree paths must be pre-zeroed),
// bias [nv], J [nv][ldj] (pre-zeroed), feet [3 nc].
template <class T> struct WbStoreSink {
  T* Mp; int ldm; T* biasp; T* Jp; int ldj; T* feetp;
  HD void mass(int i, int j, T v) { Mp[i * ldm + j] = v; Mp[j * ldm + i] = v; }
  HD void bias(int i, T v) { biasp[i] = v; }
  HD void foot(int c, const T* p) { for (int k = 0; k < 3; ++k) feetp[3 * c + k] = p[k]; }
  HD void jac(int j, int c, const T* v) { for (int k = 0; k < 3; ++k) Jp[j * ldj + 3 * c + k] = v[k]; }
};

// Forward-mode sink: folds the tangents of the mass matrix / bias force / foot Jacobian entries of ONE direction
// into the matrix-vector products the analytic contact-solve derivative needs (see wb_s3_tangent):
//   t1 = -dD - dM w          t2 = -dD - dM a + dJ grf          e = dJ y - dM z          h = dJ' (w + 2 alpha qvel + z)
struct WbTangentSink {
  float *t1, *t2, *e, *h, *dfeet;                     // per-direction accumulators (pre-zeroed)
  const float *w, *z, *a, *y, *grf, *wvz;             // primal vectors of the knot
  HD void mass(int i, int j, Dual v) {
    const float d = v.d;
    t1[i] -= d * w[j]; t2[i] -= d * a[j]; e[i] -= d * z[j];
    if (i != j) { t1[j] -= d * w[i]; t2[j] -= d * a[i]; e[j] -= d * z[i]; }
  }
  HD void bias(int i, Dual v) { t1[i] -= v.d; t2[i] -= v.d; }
  HD void foot(int c, const Dual* p) { for (int k = 0; k < 3; ++k) dfeet[3 * c + k] = p[k].d; }
  HD void jac(int j, int c, const Dual* v) {
    for (int k = 0; k < 3; ++k) {
      const float d = v[k].d;
      const int col = 3 * c + k;
      e[j] += d * y[col]; t2[j] += d * grf[col]; h[col] += d * wvz[j];
    }
  }
};

// x <- (L L')^-1 x with the strictly-lower part of L in Lm (ld) and the RECIPROCAL of its diagonal in idL; K is the
// exact size so the loops unroll and x stays in registers
template <int K> HD void wb_tri_solve(const float* Lm, int ld, const float* idL, float* x) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float s = x[i];
#pragma unroll
    for (int p = 0; p < i; ++p) s = fmaf(-Lm[i * ld + p], x[p], s);
    x[i] = s * idL[i];
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    float s = x[i];
#pragma unroll
    for (int p = i + 1; p < K; ++p) s = fmaf(-Lm[p * ld + i], x[p], s);
    x[i] = s * idL[i];
  }
}

#ifdef __CUDACC__
// one-warp register Cholesky: common.cuh (warp_cholesky), shared with the LQ kernel
template <int K, int LD> __device__ __forceinline__ void wb_warp_cholesky(float* Mat, float* dinv, int lane) {
  warp_cholesky<K, LD>(Mat, dinv, lane);
}
#endif

// Robot classes up to this many dofs form qM^-1 explicitly (wb_s3_primal_inv / wb_s3_tangent_inv): the quadrupeds and
// H1 (no spills at 168 registers; H1 x4096 linearise 21.0 -> 12.7 ms).  Talos (28 dofs, GRFs are controls) keeps the
// triangular solves: with the explicit inverse two of its scan-mode golden checks move out of tolerance and E = 1 gains
// nothing (measured, r02).
#ifndef MPXB_INV_MAXNV
#define MPXB_INV_MAXNV 25
#endif

// Per-knot primal workspace of the analytic contact solve (shared memory in the kernel)
template <int NV, int NC, bool INV = true> struct WbKnot {
  static constexpr int NF = 3 * NC, NX = 13 + 2 * (NV - 6) + 2 * NF;
  float M[NV * NV];      // qM -> strictly-lower Cholesky factor in place (diagonal in dL)
  float J[NV * NF], MiJ[NV * NF], JMJ[NF * NF];
  // dL, dLc: reciprocal Cholesky diagonals.  tmd sits in front of the read-only D on purpose: the compiler reads tmd with
  // 16-byte loads that run two floats past its end, and with w behind it racecheck flags that (unused) over-read against
  // the workers writing w in the same phase.
  float dL[NV], dLc[NF], tmd[NV], D[NV], w[NV], a[NV], z[NV], wvz[NV];
  float y[NF], grf[NF], feet[NF];
  float xn[NX], Gq[16], Gw[12];
  // explicit-inverse variant (NV <= MPXB_INV_MAXNV): L^-1 and the right-hand side of the GRF solve
  // (INV = false: trial items, which never form the inverse -- wb_s3_primal_fwd)
  float Li[(INV && NV <= MPXB_INV_MAXNV) ? NV * NV : NV], yr[NF];
};

// Cooperative primal contact solve (models.py:93-115 / 146-167 / 206-215): `lane` of `nl` workers, `sync()` between
// the dependent steps.  Host: lane 0 of 1 with a no-op sync.
// TANGENT = false skips what only the derivative needs (z, wvz, the quat_integrate Jacobians): trial evaluations.
template <int NV, int NC, bool TANGENT = true, class Sync>
HD void wb_s3_primal(const WbModel& md, WbKnot<NV, NC>& k, const float* x, const float* u, const float* contact, int lane,
                     int nl, Sync sync) {
  constexpr int NF = 3 * NC;
  constexpr int nv = NV, nj = NV - 6, nf = NF;  // every robot class is instantiated for its exact sizes (wb_launch.h)
  const float* qvel = x + 7 + nj;
  for (int i = lane; i < nv; i += nl) k.tmd[i] = (i < 6 ? 0.f : u[i - 6]) - k.D[i];
  for (int idx = lane; idx < nv * NF; idx += nl) k.MiJ[idx] = k.J[idx];
  sync();
  // Cholesky of qM, one column per step; every worker recomputes the pivot so a single barrier per column suffices
#ifdef __CUDA_ARCH__
  if (nl >= 32) {  // device: warp 0 factors with the rows in registers (wb_warp_cholesky), the other warps wait
    if (lane < 32) wb_warp_cholesky<NV, NV>(k.M, k.dL, lane);
    sync();
  } else
#endif
  for (int j = 0; j < nv; ++j) {
    float sjj = k.M[j * NV + j];
    for (int p = 0; p < j; ++p) sjj = fmaf(-k.M[j * NV + p], k.M[j * NV + p], sjj);
    const float dj = chol_pivot_inv(sjj);
    for (int i = j + 1 + lane; i < nv; i += nl) {
      float s = k.M[i * NV + j];
      for (int p = 0; p < j; ++p) s = fmaf(-k.M[i * NV + p], k.M[j * NV + p], s);
      k.M[i * NV + j] = s * dj;
    }
    if (lane == 0) k.dL[j] = dj;
    sync();
  }
  if (md.contact_solve) {
    for (int c = lane; c <= nf; c += nl) {
      if (c < nf) {
        float col[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) col[i] = k.MiJ[i * NF + c];
        wb_tri_solve<NV>(k.M, NV, k.dL, col);
#pragma unroll
        for (int i = 0; i < NV; ++i) k.MiJ[i * NF + c] = col[i];
      } else {
        float col[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) col[i] = k.tmd[i];
        wb_tri_solve<NV>(k.M, NV, k.dL, col);
#pragma unroll
        for (int i = 0; i < NV; ++i) k.w[i] = col[i];
      }
    }
    sync();
    for (int idx = lane; idx < nf * nf + nf; idx += nl) {
      if (idx < nf * nf) {
        const int a = idx / nf, b = idx - a * nf;
        if (b <= a) {
          float s = 0.f;
          for (int i = 0; i < nv; ++i) s = fmaf(k.J[i * NF + a], k.MiJ[i * NF + b], s);
          if (a == b) s += md.reg;
          k.JMJ[a * NF + b] = s;
        }
      } else {
        const int a = idx - nf * nf;
        float s = 0.f;
        for (int i = 0; i < nv; ++i) s = fmaf(k.J[i * NF + a], k.w[i] + 2.f * md.alpha * qvel[i], s);
        k.y[a] = -s;
      }
    }
    sync();
#ifdef __CUDA_ARCH__
    if (nl >= 32 && NF <= 32) {
      if constexpr (NF <= 32) {
        if (lane < 32) wb_warp_cholesky<NF, NF>(k.JMJ, k.dLc, lane);
      }
      sync();
    } else
#endif
    for (int j = 0; j < nf; ++j) {
      float sjj = k.JMJ[j * NF + j];
      for (int p = 0; p < j; ++p) sjj = fmaf(-k.JMJ[j * NF + p], k.JMJ[j * NF + p], sjj);
      const float dj = chol_pivot_inv(sjj);
      for (int i = j + 1 + lane; i < nf; i += nl) {
        float s = k.JMJ[i * NF + j];
        for (int p = 0; p < j; ++p) s = fmaf(-k.JMJ[i * NF + p], k.JMJ[j * NF + p], s);
        k.JMJ[i * NF + j] = s * dj;
      }
      if (lane == 0) k.dLc[j] = dj;
      sync();
    }
    if (lane == 0) {
      float yy[NF];
#pragma unroll
      for (int c = 0; c < NF; ++c) yy[c] = k.y[c];
      wb_tri_solve<NF>(k.JMJ, NF, k.dLc, yy);
#pragma unroll
      for (int c = 0; c < NF; ++c) k.y[c] = yy[c];
    }
    sync();
    for (int c = lane; c < nf; c += nl) k.grf[c] = k.y[c] * contact[c / 3];
    sync();
    for (int i = lane; i < nv; i += nl) {
      float sz = 0.f, sa = k.w[i];
      for (int c = 0; c < nf; ++c) { sz = fmaf(k.MiJ[i * NF + c], k.y[c], sz); sa = fmaf(k.MiJ[i * NF + c], k.grf[c], sa); }
      k.z[i] = sz;
      k.a[i] = sa;
      k.wvz[i] = k.w[i] + 2.f * md.alpha * qvel[i] + sz;
    }
  } else {
    for (int c = lane; c < nf; c += nl) { k.grf[c] = u[nj + c] * contact[c / 3]; k.y[c] = 0.f; }
    for (int i = lane; i < nv; i += nl) { k.w[i] = 0.f; k.z[i] = 0.f; k.wvz[i] = 0.f; }
    sync();
    if (lane == 0) {
      float col[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = k.tmd[i];
        for (int c = 0; c < nf; ++c) s = fmaf(k.J[i * NF + c], k.grf[c], s);
        col[i] = s;
      }
      wb_tri_solve<NV>(k.M, NV, k.dL, col);
#pragma unroll
      for (int i = 0; i < NV; ++i) k.a[i] = col[i];
    }
  }
  sync();
  // semi-implicit Euler step and the Jacobians of quat_integrate wrt (quat, omega+)
  const float dt = md.dt;
  float* vn = k.xn + 7 + nj;
  for (int i = lane; i < nv; i += nl) vn[i] = qvel[i] + k.a[i] * dt;
  sync();
  for (int i = lane; i < 3; i += nl) k.xn[i] = x[i] + vn[i] * dt;
  for (int i = lane; i < nj; i += nl) k.xn[7 + i] = x[7 + i] + vn[6 + i] * dt;
  for (int i = lane; i < nf; i += nl) {
    k.xn[7 + nj + nv + i] = k.feet[i];
    if (md.grf_as_state) k.xn[7 + nj + nv + nf + i] = k.grf[i];
  }
  for (int s = (TANGENT ? lane : 7 + lane); s < 8; s += nl) {
    if (s < 7) {
      Dual q[4], om[3], qo[4];
      for (int i = 0; i < 4; ++i) q[i] = Dual(x[3 + i], s == i ? 1.f : 0.f);
      for (int i = 0; i < 3; ++i) om[i] = Dual(vn[3 + i], s == 4 + i ? 1.f : 0.f);
      quat_integrate(q, om, dt, qo);
      for (int r = 0; r < 4; ++r) {
        if (s < 4) k.Gq[r * 4 + s] = qo[r].d;
        else k.Gw[r * 3 + s - 4] = qo[r].d;
      }
    } else {
      float qo[4];
      quat_integrate(x + 3, vn + 3, dt, qo);
      for (int r = 0; r < 4; ++r) k.xn[3 + r] = qo[r];
    }
  }
  sync();
}

// ---------------------------------------------------------------------------------------------------------------
// Explicit-inverse variant of the contact solve (v3 linearise kernel, cooperative trial kernel).  Same quantities as
// wb_s3_primal, but after the two Cholesky factorisations the inverses qM^-1 = L^-T L^-1 and (J'M^-1J)^-1 are formed
// once per knot (k.M and k.JMJ then hold the full symmetric inverses), so that every later solve -- M^-1 J, w, the
// GRFs, and the two qM solves + one contact solve PER COLUMN of [A | B] in wb_s3_tangent_inv -- is a mat-vec with
// independent accumulators instead of a 300-deep dependent FMA chain.  The chains were what bounded those stages
// (12 warps per SM cannot hide them).
template <int NV, int NC, bool TANGENT = true, class Sync>
HD void wb_s3_primal_inv(const WbModel& md, WbKnot<NV, NC>& k, const float* x, const float* u, const float* contact,
                         int lane, int nl, Sync sync) {
  constexpr int NF = 3 * NC;
  constexpr int nv = NV, nj = NV - 6, nf = NF;  // every robot class is instantiated for its exact sizes (wb_launch.h)
  const float* qvel = x + 7 + nj;
  for (int i = lane; i < nv; i += nl) k.tmd[i] = (i < 6 ? 0.f : u[i - 6]) - k.D[i];
  sync();
  // Cholesky of qM, one column per step; every worker recomputes the pivot so a single barrier per column suffices
#ifdef __CUDA_ARCH__
  if (nl >= 32) {  // device: warp 0 factors with the rows in registers (wb_warp_cholesky), the other warps wait
    if (lane < 32) wb_warp_cholesky<NV, NV>(k.M, k.dL, lane);
    sync();
  } else
#endif
  for (int j = 0; j < nv; ++j) {
    float sjj = k.M[j * NV + j];
    for (int p = 0; p < j; ++p) sjj = fmaf(-k.M[j * NV + p], k.M[j * NV + p], sjj);
    const float dj = chol_pivot_inv(sjj);
    for (int i = j + 1 + lane; i < nv; i += nl) {
      float s = k.M[i * NV + j];
      for (int p = 0; p < j; ++p) s = fmaf(-k.M[i * NV + p], k.M[j * NV + p], s);
      k.M[i * NV + j] = s * dj;
    }
    if (lane == 0) k.dL[j] = dj;
    sync();
  }
  // L^-1, one column per worker
  for (int j = lane; j < nv; j += nl) {
    k.Li[j * NV + j] = k.dL[j];
    for (int i = j + 1; i < nv; ++i) {
      float s = 0.f;
      for (int p = j; p < i; ++p) s = fmaf(k.M[i * NV + p], k.Li[p * NV + j], s);
      k.Li[i * NV + j] = -s * k.dL[i];
    }
  }
  sync();
  // qM^-1 = L^-T L^-1 (full symmetric) over the factor
  for (int idx = lane; idx < nv * nv; idx += nl) {
    const int i = idx / nv, j = idx - i * nv;
    float s = 0.f;
    for (int q = (i > j ? i : j); q < nv; ++q) s = fmaf(k.Li[q * NV + i], k.Li[q * NV + j], s);
    k.M[i * NV + j] = s;
  }
  sync();
  if (md.contact_solve) {
    for (int idx = lane; idx < nv * nf + nv; idx += nl) {
      if (idx < nv * nf) {
        const int i = idx / nf, c = idx - i * nf;
        float s = 0.f;
        for (int q = 0; q < nv; ++q) s = fmaf(k.M[i * NV + q], k.J[q * NF + c], s);
        k.MiJ[i * NF + c] = s;
      } else {
        const int i = idx - nv * nf;
        float s = 0.f;
        for (int q = 0; q < nv; ++q) s = fmaf(k.M[i * NV + q], k.tmd[q], s);
        k.w[i] = s;
      }
    }
    sync();
    for (int idx = lane; idx < nf * nf + nf; idx += nl) {
      if (idx < nf * nf) {
        const int a = idx / nf, b = idx - a * nf;
        if (b <= a) {
          float s = 0.f;
          for (int i = 0; i < nv; ++i) s = fmaf(k.J[i * NF + a], k.MiJ[i * NF + b], s);
          if (a == b) s += md.reg;
          k.JMJ[a * NF + b] = s;
        }
      } else {
        const int a = idx - nf * nf;
        float s = 0.f;
        for (int i = 0; i < nv; ++i) s = fmaf(k.J[i * NF + a], k.w[i] + 2.f * md.alpha * qvel[i], s);
        k.yr[a] = -s;
      }
    }
    sync();
#ifdef __CUDA_ARCH__
    if (nl >= 32 && NF <= 32) {
      if constexpr (NF <= 32) {
        if (lane < 32) wb_warp_cholesky<NF, NF>(k.JMJ, k.dLc, lane);
      }
      sync();
    } else
#endif
    for (int j = 0; j < nf; ++j) {
      float sjj = k.JMJ[j * NF + j];
      for (int p = 0; p < j; ++p) sjj = fmaf(-k.JMJ[j * NF + p], k.JMJ[j * NF + p], sjj);
      const float dj = chol_pivot_inv(sjj);
      for (int i = j + 1 + lane; i < nf; i += nl) {
        float s = k.JMJ[i * NF + j];
        for (int p = 0; p < j; ++p) s = fmaf(-k.JMJ[i * NF + p], k.JMJ[j * NF + p], s);
        k.JMJ[i * NF + j] = s * dj;
      }
      if (lane == 0) k.dLc[j] = dj;
      sync();
    }
    // the contact system stays a Cholesky solve: for H1's regularised line contacts (models.py:155-157) an explicit
    // inverse of J'M^-1J loses three digits
    if (lane == 0) {
      float yy[NF];
#pragma unroll
      for (int c = 0; c < NF; ++c) yy[c] = k.yr[c];
      wb_tri_solve<NF>(k.JMJ, NF, k.dLc, yy);
#pragma unroll
      for (int c = 0; c < NF; ++c) { k.y[c] = yy[c]; k.grf[c] = yy[c] * contact[c / 3]; }
    }
    sync();
    for (int i = lane; i < nv; i += nl) {
      float sz = 0.f, sa = k.w[i];
      for (int c = 0; c < nf; ++c) { sz = fmaf(k.MiJ[i * NF + c], k.y[c], sz); sa = fmaf(k.MiJ[i * NF + c], k.grf[c], sa); }
      k.z[i] = sz;
      k.a[i] = sa;
      k.wvz[i] = k.w[i] + 2.f * md.alpha * qvel[i] + sz;
    }
  } else {
    for (int c = lane; c < nf; c += nl) { k.grf[c] = u[nj + c] * contact[c / 3]; k.y[c] = 0.f; }
    for (int i = lane; i < nv; i += nl) { k.w[i] = 0.f; k.z[i] = 0.f; k.wvz[i] = 0.f; }
    sync();
    for (int i = lane; i < nv; i += nl) {  // tmd <- tau - D + J grf, then a = qM^-1 tmd
      float s = k.tmd[i];
      for (int c = 0; c < nf; ++c) s = fmaf(k.J[i * NF + c], k.grf[c], s);
      k.Li[i] = s;  // L^-1 is dead: scratch
    }
    sync();
    for (int i = lane; i < nv; i += nl) {
      float s = 0.f;
      for (int q = 0; q < nv; ++q) s = fmaf(k.M[i * NV + q], k.Li[q], s);
      k.a[i] = s;
    }
  }
  sync();
  // semi-implicit Euler step and the Jacobians of quat_integrate wrt (quat, omega+)
  const float dt = md.dt;
  float* vn = k.xn + 7 + nj;
  for (int i = lane; i < nv; i += nl) vn[i] = qvel[i] + k.a[i] * dt;
  sync();
  for (int i = lane; i < 3; i += nl) k.xn[i] = x[i] + vn[i] * dt;
  for (int i = lane; i < nj; i += nl) k.xn[7 + i] = x[7 + i] + vn[6 + i] * dt;
  for (int i = lane; i < nf; i += nl) {
    k.xn[7 + nj + nv + i] = k.feet[i];
    if (md.grf_as_state) k.xn[7 + nj + nv + nf + i] = k.grf[i];
  }
  for (int s = (TANGENT ? lane : 7 + lane); s < 8; s += nl) {
    if (s < 7) {
      Dual q[4], om[3], qo[4];
      for (int i = 0; i < 4; ++i) q[i] = Dual(x[3 + i], s == i ? 1.f : 0.f);
      for (int i = 0; i < 3; ++i) om[i] = Dual(vn[3 + i], s == 4 + i ? 1.f : 0.f);
      quat_integrate(q, om, dt, qo);
      for (int r = 0; r < 4; ++r) {
        if (s < 4) k.Gq[r * 4 + s] = qo[r].d;
        else k.Gw[r * 3 + s - 4] = qo[r].d;
      }
    } else {
      float qo[4];
      quat_integrate(x + 3, vn + 3, dt, qo);
      for (int r = 0; r < 4; ++r) k.xn[3 + r] = qo[r];
    }
  }
  sync();
}

// Primal contact solve for TRIAL evaluations (device only): nothing but x+ = f(x, u) is wanted, so neither qM^-1 nor
// M^-1 J is ever formed.  With qM = L L' and the forward substitutions  Y = L^-1 J,  u = L^-1 (tau - D):
//   J' M^-1 J = Y' Y,   J' M^-1 (tau - D) = Y' u,   qacc = L^-T (u + Y grf)
// i.e. one forward substitution per column (NF + 1 workers, registers), two small products and ONE backward
// substitution, done by a warp with the vector spread over its lanes -- about a third of the operations of the
// explicit-inverse route the linearise kernel takes (there every later derivative column re-uses qM^-1).
// Same quantities as wb_s3_primal / wb_s3_primal_inv up to rounding.
#ifdef __CUDACC__
template <int NV, int NC, class Knot, class Sync>
__device__ __forceinline__ void wb_s3_primal_fwd(const WbModel& md, Knot& k, const float* x, const float* u,
                                                 const float* contact, int lane, int nl, Sync sync) {
  constexpr int NF = 3 * NC;
  constexpr int nv = NV, nj = NV - 6, nf = NF;
  static_assert(NV <= 32 && NF <= 32, "one warp: a lane per row");
  const float* qvel = x + 7 + nj;
  for (int i = lane; i < nv; i += nl) k.tmd[i] = (i < 6 ? 0.f : u[i - 6]) - k.D[i];
  sync();
  if (lane < 32) wb_warp_cholesky<NV, NV>(k.M, k.dL, lane);
  sync();
  // Y = L^-1 J (columns 0..nf-1, into MiJ), u = L^-1 tmd (column nf, into w)
  for (int c = lane; c <= nf; c += nl) {
    float y[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float sacc = (c < nf) ? k.J[i * NF + c] : k.tmd[i];
#pragma unroll
      for (int p = 0; p < i; ++p) sacc = fmaf(-k.M[i * NV + p], y[p], sacc);
      y[i] = sacc * k.dL[i];
    }
    if (c < nf) {
#pragma unroll
      for (int i = 0; i < NV; ++i) k.MiJ[i * NF + c] = y[i];
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) k.w[i] = y[i];
    }
  }
  sync();
  for (int idx = lane; idx < nf * nf + nf; idx += nl) {
    if (idx < nf * nf) {
      const int a = idx / nf, b = idx - a * nf;
      if (b <= a) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < nv; ++i) sacc = fmaf(k.MiJ[i * NF + a], k.MiJ[i * NF + b], sacc);
        if (a == b) sacc += md.reg;
        k.JMJ[a * NF + b] = sacc;
      }
    } else {
      const int a = idx - nf * nf;
      float sacc = 0.f, sv = 0.f;
#pragma unroll
      for (int i = 0; i < nv; ++i) {
        sacc = fmaf(k.MiJ[i * NF + a], k.w[i], sacc);
        sv = fmaf(k.J[i * NF + a], qvel[i], sv);
      }
      k.yr[a] = -(sacc + 2.f * md.alpha * sv);
    }
  }
  sync();
  if (lane < 32) wb_warp_cholesky<NF, NF>(k.JMJ, k.dLc, lane);
  sync();
  if (lane == 0) {
    float yy[NF];
#pragma unroll
    for (int c = 0; c < NF; ++c) yy[c] = k.yr[c];
    wb_tri_solve<NF>(k.JMJ, NF, k.dLc, yy);
#pragma unroll
    for (int c = 0; c < NF; ++c) { k.y[c] = yy[c]; k.grf[c] = yy[c] * contact[c / 3]; }
  }
  sync();
  // qacc = L^-T (u + Y grf): backward substitution across the lanes of warp 0 (lane i holds entry i)
  if (lane < 32) {
    float zi = 0.f;
    if (lane < nv) {
      zi = k.w[lane];
#pragma unroll
      for (int c = 0; c < nf; ++c) zi = fmaf(k.MiJ[lane * NF + c], k.grf[c], zi);
    }
    float ai = 0.f;
#pragma unroll
    for (int i = NV - 1; i >= 0; --i) {
      const float xi = __shfl_sync(0xffffffffu, zi, i) * k.dL[i];
      if (lane == i) ai = xi;
      if (lane < i) zi = fmaf(-k.M[i * NV + lane], xi, zi);
    }
    if (lane < nv) k.a[lane] = ai;
  }
  sync();
  // semi-implicit Euler step (as the tail of wb_s3_primal_inv, TANGENT = false)
  const float dt = md.dt;
  float* vn = k.xn + 7 + nj;
  for (int i = lane; i < nv; i += nl) vn[i] = qvel[i] + k.a[i] * dt;
  sync();
  for (int i = lane; i < 3; i += nl) k.xn[i] = x[i] + vn[i] * dt;
  for (int i = lane; i < nj; i += nl) k.xn[7 + i] = x[7 + i] + vn[6 + i] * dt;
  for (int i = lane; i < nf; i += nl) {
    k.xn[7 + nj + nv + i] = k.feet[i];
    if (md.grf_as_state) k.xn[7 + nj + nv + nf + i] = k.grf[i];
  }
  if (lane == 0) {
    float qo[4];
    quat_integrate(x + 3, vn + 3, dt, qo);
    for (int r = 0; r < 4; ++r) k.xn[3 + r] = qo[r];
  }
  sync();
}
#endif

// wb_s3_tangent on what wb_s3_primal_inv leaves: k.M = qM^-1 (explicit), k.JMJ / k.dLc = Cholesky factor of J'M^-1J
template <int NV, int NC, class V1, class V2, class Out>
HD void wb_s3_tangent_inv(const WbModel& md, const WbKnot<NV, NC>& k, const float* contact, int is_u, int dir, V1 t1, V2 t2,
                          V2 e, V2 h, V2 dfeet, Out out) {
  constexpr int NF = 3 * NC, NJ = NV - 6, NQ = NV + 1;
  float dgrf[NF], xa[NV];
  const int kv = (!is_u && dir >= NQ) ? dir - NQ : -1;  // velocity direction index
  if (md.contact_solve) {
    float xw[NV], rhs[NF];
    if (is_u) {
#pragma unroll
      for (int i = 0; i < NV; ++i) xw[i] = k.M[i * NV + 6 + dir];  // qM^-1 e_{6+dir}
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < NV; ++q) s = fmaf(k.M[i * NV + q], t1[q], s);
        xw[i] = s;  // dw
      }
    }
#pragma unroll
    for (int c = 0; c < NF; ++c) {
      float s = is_u ? 0.f : -h[c];
#pragma unroll
      for (int i = 0; i < NV; ++i) s = fmaf(-k.J[i * NF + c], xw[i], s);
      if (!is_u) {
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fmaf(-k.MiJ[i * NF + c], e[i], s);
        if (kv >= 0) s = fmaf(-2.f * md.alpha, k.J[kv * NF + c], s);
      }
      rhs[c] = s;
    }
    wb_tri_solve<NF>(k.JMJ, NF, k.dLc, rhs);
#pragma unroll
    for (int c = 0; c < NF; ++c) dgrf[c] = rhs[c] * contact[c / 3];
  } else {
#pragma unroll
    for (int c = 0; c < NF; ++c) dgrf[c] = (is_u && dir == NJ + c) ? contact[c / 3] : 0.f;
  }
  {
    float xr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = is_u ? ((i == 6 + dir && dir < NJ) ? 1.f : 0.f) : t2[i];
#pragma unroll
      for (int c = 0; c < NF; ++c) s = fmaf(k.J[i * NF + c], dgrf[c], s);
      xr[i] = s;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NV; ++q) s = fmaf(k.M[i * NV + q], xr[q], s);
      xa[i] = s;  // da
    }
  }
  const float dt = md.dt;
#pragma unroll
  for (int i = 0; i < NV; ++i) xa[i] = ((i == kv) ? 1.f : 0.f) + dt * xa[i];  // d v+
#pragma unroll
  for (int i = 0; i < 3; ++i) out(i, ((!is_u && dir == i) ? 1.f : 0.f) + dt * xa[i]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float s = (!is_u && dir >= 3 && dir < 7) ? k.Gq[r * 4 + dir - 3] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) s = fmaf(k.Gw[r * 3 + c], xa[3 + c], s);
    out(3 + r, s);
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i) out(7 + i, ((!is_u && dir == 7 + i) ? 1.f : 0.f) + dt * xa[6 + i]);
#pragma unroll
  for (int i = 0; i < NV; ++i) out(NQ + i, xa[i]);
#pragma unroll
  for (int c = 0; c < NF; ++c) out(NQ + NV + c, is_u ? 0.f : dfeet[c]);
  if (md.grf_as_state) {
#pragma unroll
    for (int c = 0; c < NF; ++c) out(NQ + NV + NF + c, dgrf[c]);
  }
}

// One column of [A | B] from the folded tangents of one direction (see WbTangentSink), handed entry by entry to
// out(i, value), i = 0..n-1 in increasing order.
//   is_u == 0: state direction `dir` in [0, 7+nj+nv) ; is_u == 1: control direction `dir` in [0, m).
// t1, t2, e, h, dfeet are the direction's accumulators (ignored for a control direction).  Requires nv == NV, nc == NC.
// element i of a per-thread vector interleaved across the threads of a CTA in shared memory ([i][thread])
struct WbStrided {
  float* p;
  int stride;
  HD float& operator[](int i) const { return p[i * stride]; }
};

template <int NV, int NC, class V1, class V2, class Out>
HD void wb_s3_tangent(const WbModel& md, const WbKnot<NV, NC>& k, const float* contact, int is_u, int dir, V1 t1, V2 t2,
                      V2 e, V2 h, V2 dfeet, Out out) {
  constexpr int NF = 3 * NC, NJ = NV - 6, NQ = NV + 1;
  float dgrf[NF], xa[NV];
  const int kv = (!is_u && dir >= NQ) ? dir - NQ : -1;  // velocity direction index
  if (md.contact_solve) {
    float xw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) xw[i] = is_u ? ((i == 6 + dir) ? 1.f : 0.f) : t1[i];
    wb_tri_solve<NV>(k.M, NV, k.dL, xw);  // dw
#pragma unroll
    for (int c = 0; c < NF; ++c) {
      float s = is_u ? 0.f : -h[c];
#pragma unroll
      for (int i = 0; i < NV; ++i) s = fmaf(-k.J[i * NF + c], xw[i], s);
      if (!is_u) {
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fmaf(-k.MiJ[i * NF + c], e[i], s);
        if (kv >= 0) s = fmaf(-2.f * md.alpha, k.J[kv * NF + c], s);
      }
      dgrf[c] = s;
    }
    wb_tri_solve<NF>(k.JMJ, NF, k.dLc, dgrf);
#pragma unroll
    for (int c = 0; c < NF; ++c) dgrf[c] *= contact[c / 3];
  } else {
#pragma unroll
    for (int c = 0; c < NF; ++c) dgrf[c] = (is_u && dir == NJ + c) ? contact[c / 3] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = is_u ? ((i == 6 + dir && dir < NJ) ? 1.f : 0.f) : t2[i];
#pragma unroll
    for (int c = 0; c < NF; ++c) s = fmaf(k.J[i * NF + c], dgrf[c], s);
    xa[i] = s;
  }
  wb_tri_solve<NV>(k.M, NV, k.dL, xa);  // da
  const float dt = md.dt;
#pragma unroll
  for (int i = 0; i < NV; ++i) xa[i] = ((i == kv) ? 1.f : 0.f) + dt * xa[i];  // d v+
#pragma unroll
  for (int i = 0; i < 3; ++i) out(i, ((!is_u && dir == i) ? 1.f : 0.f) + dt * xa[i]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float s = (!is_u && dir >= 3 && dir < 7) ? k.Gq[r * 4 + dir - 3] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) s = fmaf(k.Gw[r * 3 + c], xa[3 + c], s);
    out(3 + r, s);
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i) out(7 + i, ((!is_u && dir == 7 + i) ? 1.f : 0.f) + dt * xa[6 + i]);
#pragma unroll
  for (int i = 0; i < NV; ++i) out(NQ + i, xa[i]);
#pragma unroll
  for (int c = 0; c < NF; ++c) out(NQ + NV + c, is_u ? 0.f : dfeet[c]);
  if (md.grf_as_state) {
#pragma unroll
    for (int c = 0; c < NF; ++c) out(NQ + NV + NF + c, dgrf[c]);
  }
}

}  // namespace mpxb
