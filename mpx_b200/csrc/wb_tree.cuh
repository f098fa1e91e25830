// wb_tree.cuh -- whole-body linearisation, v3: all tree quantities of one knot held per body in shared memory and
// the derivatives of qM, qfrc_bias and the foot Jacobians taken ANALYTICALLY instead of by forward-mode passes.
//
// What is differentiated is still the reference's dynamics (models.py:39-217 with MJX 3.3.0 fwd_position /
// fwd_velocity / support.jac, restated in wb.cuh); only the way the Jacobians [A | B] are obtained changes.
// jax.jacobian (trajax linearize, optimizers.py:165-166) is replaced by the closed-form spatial-algebra identities
// for a kinematic tree: perturbing joint d moves the subtree of d rigidly with the motion vector S_d, so for every
// body b in sub(d), about a fixed world reference point,
//     dS_j = S_d x S_j (j below d)          dI_b = S_d x* I_b - I_b S_d x
//     dv_b = S_d x (v_b - v_d)              dc_b = S_d x (c_b - c_d) - (S_d x v_d) x (v_b - v_d)
// and the raw base quaternion components act as four virtual joints S = (omega_k, 0) placed between the
// translational and the rotational dofs of the free joint.  With U_i = Ic_i S_i (the CRBA force columns) this gives
// the contractions the contact-solve derivative (wb_s3_tangent) needs in O(nv) per direction:
//     d(M u)_i  = U_i . mu                          if dof i moves with d      mu = -S_d x alpha_d(u)
//               = S_i . (S_d x* F_d(u) + Ic_d mu)   if i is an ancestor of d
//     d(D)_i    = S_i . G'_i   /   S_i . (S_d x* Fb_d + G'_d)      G' = subtree sums of
//                 g'_b = I_b (mu_c + kappa x (v_b - v_d)) + kappa x* (I_b v_b) + v_b x* (I_b kappa),  kappa = -S_d x v_d
//     dD/dqvel_d: g''_b = I_b (S_d x (v_b - r_d)) + S_d x* (I_b v_b) + v_b x* (I_b S_d)
//     dJ_ic     = omega_d x J_ic  (i moves with d)   /   omega_i x S_d(r_c)  (i ancestor of d)
// Everything is __host__ __device__ and written for `lane` of `nl` cooperating workers with `sync()` between the
// dependent stages, so tests/cpu_harness runs the same code with one worker.
#pragma once
#include "wb.cuh"

namespace mpxb {

template <int NB, int NV, int NC, int NI> struct WbTree {
  // ---- scratch block: nothing from here to `F` is read once the direction frames exist (wb_tree_fold_coop, F1), so
  // the linearise kernel lets the narrow half of the folded tangents live over it (scratch() / SCRATCH floats)
  float pos[NB][3], quat[NB][4];
  float v[NB][6], c[NB][6], I[NB][10], Ic[NB][10], h[NB][6], Fb[NB][6];
  // acceleration problems of the linearise kernel only: trial / interface instantiations (NI == 1) carry no items
  // and drop them (3 KB of an H1 trial item's 22 KB)
  float al[NI > 1 ? 3 : 1][NI > 1 ? NB : 1][6], F[NI > 1 ? 3 : 1][NI > 1 ? NB : 1][6];
  // ---- live until the columns of [A | B] are out
  float S[NV][6], U[NV][6], rv[NV][6];
  float rc[NC][3];  // foot position relative to the reference point (base position)
  float wq[4][3];   // world angular rate of a unit perturbation of each raw base-quaternion component
  float vt[6], cg[6];
  float g[NI][6];
  static constexpr int SCRATCH = NB * (3 + 4 + 6 + 6 + 10 + 10 + 6 + 6) + 2 * (NI > 1 ? 3 : 1) * (NI > 1 ? NB : 1) * 6;
  HD float* scratch() { return &pos[0][0]; }
};

template <class A, class B> HD float dot6(const A* a, const B* b) {
  float s = a[0] * b[0];
#pragma unroll
  for (int k = 1; k < 6; ++k) s = fmaf(a[k], b[k], s);
  return s;
}

// ---- stage 1: forward kinematics / velocities, one worker per leaf path (shared ancestors are recomputed and
// written with identical values)
template <int NB, int NV, int NC, int NI>
HD void wb_tree_fk(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, const float* qpos, const float* qvel, int lane, int nl) {
  for (int lf = lane; lf < md.nleaf; lf += nl) {
    float pos[3], quat[4], cv[6], ca[6], vt0[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float ref[3] = {qpos[0], qpos[1], qpos[2]};
    const int len = md.leaf_len[lf];
    for (int lvl = 0; lvl < len; ++lvl) {
      const int b = md.leaf_path[lf][lvl];
      const int jt = md.jtype[b];
      if (lvl == 0) {
        float Rb[9];
        for (int i = 0; i < 3; ++i) pos[i] = ref[i];
        normalize_with_norm<float, 4>(qpos + 3, quat);
        quat_to_mat(quat, Rb);
        for (int i = 0; i < 3; ++i) { cv[i] = 0.f; cv[3 + i] = qvel[i]; ca[i] = 0.f; ca[3 + i] = -md.gravity[i]; }
        for (int i = 0; i < 6; ++i) { tr.vt[i] = cv[i]; tr.cg[i] = ca[i]; vt0[i] = cv[i]; }
        float cd[3][6];
        for (int a = 0; a < 3; ++a) {
          const float cdr[6] = {Rb[a], Rb[3 + a], Rb[6 + a], 0.f, 0.f, 0.f};
          motion_cross(cv, cdr, cd[a]);
          for (int i = 0; i < 3; ++i) {
            tr.S[a][i] = 0.f; tr.S[a][3 + i] = (i == a) ? 1.f : 0.f;
            tr.S[3 + a][i] = Rb[3 * i + a]; tr.S[3 + a][3 + i] = 0.f;
          }
        }
        for (int a = 0; a < 3; ++a)
          for (int k = 0; k < 6; ++k) {
            if (k < 3) cv[k] = cv[k] + Rb[3 * k + a] * qvel[3 + a];
            ca[k] = ca[k] + cd[a][k] * qvel[3 + a];
          }
        // world angular rate of d quat_raw[k]: 2 vec(d qhat * conj(qhat)) -- derivative data, like rv below: the
        // trial / interface instantiations (NI == 1) skip both
        if constexpr (NI > 1)
        for (int k = 0; k < 4; ++k) {
          Dual q4[4], qn[4];
          for (int i = 0; i < 4; ++i) q4[i] = Dual(qpos[3 + i], i == k ? 1.f : 0.f);
          normalize_with_norm<Dual, 4>(q4, qn);
          const float dq[4] = {qn[0].d, qn[1].d, qn[2].d, qn[3].d};
          const float qc[4] = {quat[0], -quat[1], -quat[2], -quat[3]};
          float o[4];
          quat_mul(dq, qc, o);
          for (int i = 0; i < 3; ++i) tr.wq[k][i] = 2.f * o[1 + i];
        }
      } else {
        float r[3], q2[4];
        const unsigned fl = md.body_flags[b];   // bit 0: body_quat = identity, bit 1: jnt_pos = 0 (exact shortcuts)
        quat_rotate(quat, md.body_pos[b], r);
        for (int i = 0; i < 3; ++i) pos[i] = pos[i] + r[i];
        if (!(fl & 1u)) {
          quat_mul(quat, md.body_quat[b], q2);
          for (int i = 0; i < 4; ++i) quat[i] = q2[i];
        }
        if (jt == WB_JNT_HINGE) {
          const int dj = md.dofadr[b];
          float anchor[3], axis[3], o[3], l[3], cdof[6], cd[6];
          if (fl & 2u) {
            for (int i = 0; i < 3; ++i) anchor[i] = pos[i];
          } else {
            quat_rotate(quat, md.jnt_pos[b], r);
            for (int i = 0; i < 3; ++i) anchor[i] = r[i] + pos[i];
          }
          quat_rotate(quat, md.jnt_axis[b], axis);
          const float half = (qpos[dj + 1] - md.qpos0[b]) * 0.5f;
          const float sn = sinf(half), cs = cosf(half);
          const float qloc[4] = {cs, md.jnt_axis[b][0] * sn, md.jnt_axis[b][1] * sn, md.jnt_axis[b][2] * sn};
          quat_mul(quat, qloc, q2);
          for (int i = 0; i < 4; ++i) quat[i] = q2[i];
          if (fl & 2u) {
            for (int i = 0; i < 3; ++i) o[i] = ref[i] - anchor[i];   // pos stays at the anchor
          } else {
            quat_rotate(quat, md.jnt_pos[b], r);
            for (int i = 0; i < 3; ++i) { pos[i] = anchor[i] - r[i]; o[i] = ref[i] - anchor[i]; }
          }
          cross3(axis, o, l);
          for (int i = 0; i < 3; ++i) { cdof[i] = axis[i]; cdof[3 + i] = l[i]; }
          // rv = v_body + v_parent, the reference velocity of d qfrc_bias / d qvel
          motion_cross(cv, cdof, cd);
          const float qd = qvel[dj];
          for (int k = 0; k < 6; ++k) {
            tr.S[dj][k] = cdof[k];
            if constexpr (NI > 1) tr.rv[dj][k] = cv[k] + (cv[k] + cdof[k] * qd);
            ca[k] = ca[k] + cd[k] * qd;
            cv[k] = cv[k] + cdof[k] * qd;
          }
        }
      }
      for (int i = 0; i < 3; ++i) tr.pos[b][i] = pos[i];
      for (int i = 0; i < 4; ++i) tr.quat[b][i] = quat[i];
      for (int i = 0; i < 6; ++i) { tr.v[b][i] = cv[i]; tr.c[b][i] = ca[i]; }
      if (NI > 1 && lvl == 0)
        for (int k = 0; k < 6; ++k) {
          for (int a = 0; a < 3; ++a) tr.rv[a][k] = 0.f;
          for (int a = 3; a < 6; ++a) tr.rv[a][k] = cv[k] + vt0[k];  // own copy: the other leaf workers write tr.vt too
        }
    }
  }
}

// ---- stage 2: per body spatial inertia about the reference point, momentum, bias force; foot offsets
template <int NB, int NV, int NC, int NI>
HD void wb_tree_bodies(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, const float* qpos, float* feet, int lane, int nl) {
  const int nb = md.nbody;
  for (int b = 1 + lane; b < nb + NC; b += nl) {
    if (b >= nb) {
      const int c = b - nb, bc = md.contact_body[c];
      float r[3];
      quat_rotate(tr.quat[bc], md.contact_pos[c], r);
      for (int i = 0; i < 3; ++i) {
        feet[3 * c + i] = tr.pos[bc][i] + r[i];
        tr.rc[c][i] = feet[3 * c + i] - qpos[i];
      }
      continue;
    }
    float r[3], off[3], qi[4], R[9], ci[10];
    quat_rotate(tr.quat[b], md.ipos[b], r);
    for (int i = 0; i < 3; ++i) off[i] = tr.pos[b][i] + r[i] - qpos[i];
    quat_mul(tr.quat[b], md.iquat[b], qi);
    quat_to_mat(qi, R);
    const float mb = md.mass[b];
    const float* I = md.inertia[b];
    const float o2 = off[0] * off[0] + off[1] * off[1] + off[2] * off[2];
    const int ia[6] = {0, 1, 2, 0, 0, 1}, ib[6] = {0, 1, 2, 1, 2, 2};
    for (int k = 0; k < 6; ++k) {
      const int a = ia[k], c = ib[k];
      float v = R[3 * a] * (R[3 * c] * I[0]) + R[3 * a + 1] * (R[3 * c + 1] * I[1]) + R[3 * a + 2] * (R[3 * c + 2] * I[2]);
      v = v - off[a] * off[c] * mb;
      if (a == c) v = v + o2 * mb;
      ci[k] = v;
    }
    for (int i = 0; i < 3; ++i) ci[6 + i] = off[i] * mb;
    ci[9] = mb;
    float f1[6], f2[6], f3[6];
    inert_mul(ci, tr.c[b], f1);
    inert_mul(ci, tr.v[b], f2);
    motion_cross_force(tr.v[b], f2, f3);
    for (int i = 0; i < 10; ++i) { tr.I[b][i] = ci[i]; tr.Ic[b][i] = ci[i]; }
    for (int i = 0; i < 6; ++i) { tr.h[b][i] = f2[i]; tr.Fb[b][i] = f1[i] + f3[i]; }
  }
}

// ---- stage 3: composite inertia / bias force, one worker per component, bodies leaf to root
template <int NB, int NV, int NC, int NI>
HD void wb_tree_composite(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, int lane, int nl) {
  for (int k = lane; k < 16; k += nl) {
    for (int b = md.nbody - 1; b > 1; --b) {
      const int p = md.parent[b];
      if (k < 10) tr.Ic[p][k] += tr.Ic[b][k];
      else tr.Fb[p][k - 10] += tr.Fb[b][k - 10];
    }
  }
}

// ---- stage 4: U_i, qfrc_bias, qM (full symmetric), foot Jacobians
template <int NB, int NV, int NC, int NI>
HD void wb_tree_dofs(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, float* D, int lane, int nl) {
  for (int i = lane; i < NV; i += nl) {
    const int b = md.dof_body[i];
    float f[6];
    inert_mul(tr.Ic[b], tr.S[i], f);
    for (int k = 0; k < 6; ++k) tr.U[i][k] = f[k];
    D[i] = dot6(tr.S[i], tr.Fb[b]);
  }
}
template <int NB, int NV, int NC, int NI>
HD void wb_tree_mass_jac(const WbModel& md, const WbTree<NB, NV, NC, NI>& tr, float* M, float* J, int lane, int nl) {
  constexpr int NF = 3 * NC;
  constexpr int nv = NV;  // classes are instantiated for their exact sizes (wb_launch.h)
  for (int idx = lane; idx < NV * NV; idx += nl) {
    const int i = idx / NV, j = idx - i * NV;
    float s = 0.f;
    if (i < nv && j < nv) {
      if ((md.dof_anc[i] >> j) & 1u) s = dot6(tr.U[i], tr.S[j]);
      else if ((md.dof_anc[j] >> i) & 1u) s = dot6(tr.U[j], tr.S[i]);
      if (i == j) s += md.armature[i];
    }
    M[idx] = s;
  }
  for (int idx = lane; idx < NV * NC; idx += nl) {
    const int i = idx / NC, c = idx - i * NC;
    float jc[3] = {0.f, 0.f, 0.f};
    if (i < nv && c < NC && ((md.contact_dofmask[c] >> i) & 1u)) {
      float cr[3];
      cross3(tr.S[i], tr.rc[c], cr);
      for (int k = 0; k < 3; ++k) jc[k] = tr.S[i][3 + k] + cr[k];
    }
    for (int k = 0; k < 3; ++k) J[i * NF + 3 * c + k] = jc[k];
  }
}

// ---- stage 5 (after the contact solve): the three pure-acceleration problems u in {w, a, z}
template <int NB, int NV, int NC, int NI, class Sync>
HD void wb_tree_acc(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, const float* w, const float* a, const float* z, int lane,
                    int nl, Sync sync) {
  const int nb = md.nbody;
  for (int idx = lane; idx < 18; idx += nl) {
    const int kk = idx / 6, k = idx - kk * 6;
    const float* u = (kk == 0) ? w : (kk == 1 ? a : z);
    for (int b = 1; b < nb; ++b) {
      float s = (b == 1) ? 0.f : tr.al[kk][md.parent[b]][k];
      const int jt = md.jtype[b], d = md.dofadr[b];
      if (jt == WB_JNT_FREE) {
        for (int j = 0; j < 6; ++j) s = fmaf(tr.S[d + j][k], u[d + j], s);
      } else if (jt == WB_JNT_HINGE) {
        s = fmaf(tr.S[d][k], u[d], s);
      }
      tr.al[kk][b][k] = s;
    }
  }
  sync();
  for (int idx = lane; idx < 3 * (nb - 1); idx += nl) {
    const int kk = idx / (nb - 1), b = 1 + idx - kk * (nb - 1);
    float f[6];
    inert_mul(tr.I[b], tr.al[kk][b], f);
    for (int k = 0; k < 6; ++k) tr.F[kk][b][k] = f[k];
  }
  sync();
  for (int idx = lane; idx < 18; idx += nl) {
    const int kk = idx / 6, k = idx - kk * 6;
    for (int b = nb - 1; b > 1; --b) tr.F[kk][md.parent[b]][k] += tr.F[kk][b][k];
  }
}

// motion vector, reference velocity / bias acceleration of position direction D (0..3 quaternion, 4.. hinge dof D+2)
template <int NB, int NV, int NC, int NI>
HD void wb_dir_frame(const WbModel& md, const WbTree<NB, NV, NC, NI>& tr, int D, float* Sd, const float*& vd, const float*& cd) {
  if (D < 4) {
    for (int i = 0; i < 3; ++i) { Sd[i] = tr.wq[D][i]; Sd[3 + i] = 0.f; }
    vd = tr.vt; cd = tr.cg;
  } else {
    const int dof = D + 2, b = md.dof_body[dof];
    for (int i = 0; i < 6; ++i) Sd[i] = tr.S[dof][i];
    vd = tr.v[b]; cd = tr.c[b];
  }
}

// ---- stage 6: per (direction, body-in-subtree) force perturbations g' / g'', then subtree sums inside each direction
template <int NB, int NV, int NC, int NI, class Sync>
HD void wb_tree_items(const WbModel& md, WbTree<NB, NV, NC, NI>& tr, int lane, int nl, Sync sync) {
  constexpr int npos = 4 + NV - 6;
  for (int it = lane; it < md.nitems; it += nl) {
    const int D = md.item_dir[it];
    const int b = md.dir_b0[D] + it - md.dir_off[D];
    float t[6], t2[6], t3[6];
    if (D < npos) {
      float Sd[6], kap[6], mu[6], dv[6], rho[6];
      const float *vd, *cd;
      wb_dir_frame(md, tr, D, Sd, vd, cd);
      motion_cross(Sd, vd, kap);
      motion_cross(Sd, cd, mu);
      for (int k = 0; k < 6; ++k) { kap[k] = -kap[k]; dv[k] = tr.v[b][k] - vd[k]; }
      motion_cross(kap, dv, rho);
      for (int k = 0; k < 6; ++k) rho[k] = rho[k] - mu[k];
      inert_mul(tr.I[b], rho, t);
      motion_cross_force(kap, tr.h[b], t2);
      float Ik[6];
      inert_mul(tr.I[b], kap, Ik);
      motion_cross_force(tr.v[b], Ik, t3);
    } else {
      const int dof = D - npos;
      float dv[6], dc[6], Is[6];
      for (int k = 0; k < 6; ++k) dv[k] = tr.v[b][k] - tr.rv[dof][k];
      motion_cross(tr.S[dof], dv, dc);
      inert_mul(tr.I[b], dc, t);
      motion_cross_force(tr.S[dof], tr.h[b], t2);
      inert_mul(tr.I[b], tr.S[dof], Is);
      motion_cross_force(tr.v[b], Is, t3);
    }
    for (int k = 0; k < 6; ++k) tr.g[it][k] = t[k] + t2[k] + t3[k];
  }
  sync();
  constexpr int ndir = npos + NV;
  for (int idx = lane; idx < ndir * 6; idx += nl) {
    const int D = idx / 6, k = idx - D * 6;
    const int b0 = md.dir_b0[D], off = md.dir_off[D];
    for (int lb = md.dir_nb[D] - 1; lb > 0; --lb) tr.g[off + md.parent[b0 + lb] - b0][k] += tr.g[off + lb][k];
  }
}

// ---- stage 7: folded tangents of one x-direction `col` (0..nq+nv-1) -- the vectors WbTangentSink accumulates in v2.
// y, grf, wvz, J are the primal vectors of the knot (WbKnot).  Outputs must have NV / 3 NC entries.
template <int NB, int NV, int NC, int NI, class V1, class V2>
HD void wb_tree_fold(const WbModel& md, const WbTree<NB, NV, NC, NI>& tr, const float* J, const float* w, const float* a,
                     const float* z, const float* y, const float* grf, const float* wvz, int col, V1 t1, V2 t2, V2 e, V2 h,
                     V2 dfeet) {
  constexpr int NF = 3 * NC;
#pragma unroll
  for (int i = 0; i < NV; ++i) { t1[i] = 0.f; t2[i] = 0.f; e[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < NF; ++i) { h[i] = 0.f; dfeet[i] = 0.f; }
  if (col < 3) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 3; ++i) dfeet[3 * c + i] = (i == col) ? 1.f : 0.f;
    return;
  }
  const int D = col - 3, npos = 4 + NV - 6;
  const int b0 = md.dir_b0[D], nbd = md.dir_nb[D], off = md.dir_off[D];
  if (D >= npos) {
    const int dof = D - npos;
    const uint32_t anc = md.dof_anc[dof];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int bi = md.dof_body[i];
      float s = 0.f;
      if (bi >= b0 && bi < b0 + nbd) s = dot6(tr.S[i], tr.g[off + bi - b0]);
      else if ((anc >> i) & 1u) s = dot6(tr.S[i], tr.g[off]);
      t1[i] = -s;
      t2[i] = -s;
    }
    return;
  }
  const bool isq = D < 4;
  float Sd[6];
  const float *vd, *cd;
  wb_dir_frame(md, tr, D, Sd, vd, cd);
  const uint32_t anc = isq ? 7u : md.dof_anc[D + 2];
  float X[4][6], mu[3][6];
  motion_cross_force(Sd, tr.Fb[b0], X[0]);
#pragma unroll
  for (int k = 0; k < 6; ++k) X[0][k] += tr.g[off][k];
#pragma unroll
  for (int kk = 0; kk < 3; ++kk) {
    const float* u = (kk == 0) ? w : (kk == 1 ? a : z);
    float ad[6], t[6];
    if (isq) { ad[0] = ad[1] = ad[2] = 0.f; ad[3] = u[0]; ad[4] = u[1]; ad[5] = u[2]; }
    else { for (int k = 0; k < 6; ++k) ad[k] = tr.al[kk][b0][k]; }
    motion_cross(Sd, ad, mu[kk]);
#pragma unroll
    for (int k = 0; k < 6; ++k) mu[kk][k] = -mu[kk][k];
    motion_cross_force(Sd, tr.F[kk][b0], X[1 + kk]);
    inert_mul(tr.Ic[b0], mu[kk], t);
#pragma unroll
    for (int k = 0; k < 6; ++k) X[1 + kk][k] += t[k];
  }
  uint32_t moves = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int bi = md.dof_body[i];
    const bool mv = isq ? (i >= 3) : (bi >= b0 && bi < b0 + nbd);
    if (mv) {
      moves |= 1u << i;
      const float dD = dot6(tr.S[i], tr.g[off + bi - b0]);
      t1[i] = -(dD + dot6(tr.U[i], mu[0]));
      t2[i] = -(dD + dot6(tr.U[i], mu[1]));
      e[i] = -dot6(tr.U[i], mu[2]);
    } else if ((anc >> i) & 1u) {
      const float dD = dot6(tr.S[i], X[0]);
      t1[i] = -(dD + dot6(tr.S[i], X[1]));
      t2[i] = -(dD + dot6(tr.S[i], X[2]));
      e[i] = -dot6(tr.S[i], X[3]);
    }
  }
  // foot Jacobian terms
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int bc = md.contact_body[c];
    if (bc < b0 || bc >= b0 + nbd) continue;
    float p[3], hc[3] = {0.f, 0.f, 0.f};
    cross3(Sd, tr.rc[c], p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] += Sd[3 + k]; dfeet[3 * c + k] = p[k]; }
    const uint32_t mask = md.contact_dofmask[c];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (!((mask >> i) & 1u)) continue;
      float dJ[3];
      if ((moves >> i) & 1u) cross3(Sd, J + i * NF + 3 * c, dJ);
      else cross3(tr.S[i], p, dJ);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        t2[i] = fmaf(dJ[k], grf[3 * c + k], t2[i]);
        e[i] = fmaf(dJ[k], y[3 * c + k], e[i]);
        hc[k] = fmaf(dJ[k], wvz[i], hc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) h[3 * c + k] = hc[k];
  }
}

// ---- stage 7, cooperative form: the same folded tangents for ALL x-directions of the knot at once, as independent
// work items instead of one long thread per direction (which kept 19 of 64 lanes busy).
//   F1  worker per position direction: its frame (S_d, the four X vectors, the three mu, moves / ancestor masks)
//   F2  items (direction, dof)  -> t1, t2, e   and items (direction, foot) -> h, dfeet
// Results go to `fold` ([row][col], rows t1 NV, ld = FW >= nq + nv, all x-directions) and to the NARROW blocks `foldn`
// (rows t2 NV | e NV) and `foldh` (rows h 3NC | dfeet 3NC), both [row][col] with ld = FWN >= nq + 1, position
// directions only: for a velocity direction t2 = t1 and e = h = dfeet = 0, which the readers take from t1 and from the
// all-zero column nq of the narrow blocks; from there the column threads of the tangent stage pick them up.  F1 is the
// last reader of the tree's scratch block, so the narrow blocks may live over it.
template <int NV> struct WbDirW {
  float Sd[6], X[4][6], mu[3][6];
  uint32_t moves, anc;
  int b0, nbd, off, pad;
};
template <int NB, int NV, int NC, int NI, class Sync>
HD void wb_tree_fold_coop(const WbModel& md, const WbTree<NB, NV, NC, NI>& tr, const float* J, const float* w,
                          const float* a, const float* z, const float* y, const float* grf, const float* wvz,
                          WbDirW<NV>* dw, float* fold, int FW, float* foldn, float* foldh, int FWN, int lane, int nl, Sync sync,
                          const WbFoldTable* ft = nullptr) {
  constexpr int NF = 3 * NC, NQ = NV + 1;
  constexpr int npos = 4 + NV - 6, ncol = NQ + NV;
  // F1
  for (int col = 3 + lane; col < NQ; col += nl) {
    const int D = col - 3;
    WbDirW<NV>& d = dw[col - 3];
    const int b0 = md.dir_b0[D], nbd = md.dir_nb[D], off = md.dir_off[D];
    const bool isq = D < 4;
    float Sd[6];
    const float *vd, *cd;
    wb_dir_frame(md, tr, D, Sd, vd, cd);
    float X0[6];
    motion_cross_force(Sd, tr.Fb[b0], X0);
    for (int k = 0; k < 6; ++k) { d.Sd[k] = Sd[k]; d.X[0][k] = X0[k] + tr.g[off][k]; }
    for (int kk = 0; kk < 3; ++kk) {
      const float* u = (kk == 0) ? w : (kk == 1 ? a : z);
      float ad[6], t[6], mu[6], Xk[6];
      if (isq) { ad[0] = ad[1] = ad[2] = 0.f; ad[3] = u[0]; ad[4] = u[1]; ad[5] = u[2]; }
      else { for (int k = 0; k < 6; ++k) ad[k] = tr.al[kk][b0][k]; }
      motion_cross(Sd, ad, mu);
      for (int k = 0; k < 6; ++k) mu[k] = -mu[k];
      motion_cross_force(Sd, tr.F[kk][b0], Xk);
      inert_mul(tr.Ic[b0], mu, t);
      for (int k = 0; k < 6; ++k) { d.mu[kk][k] = mu[k]; d.X[1 + kk][k] = Xk[k] + t[k]; }
    }
    uint32_t moves = 0;
    for (int i = 0; i < NV; ++i) {
      const int bi = md.dof_body[i];
      if (isq ? (i >= 3) : (bi >= b0 && bi < b0 + nbd)) moves |= 1u << i;
    }
    d.moves = moves;
    d.anc = isq ? 7u : md.dof_anc[D + 2];
    d.b0 = b0; d.nbd = nbd; d.off = off;
  }
  sync();
  // F2
  const int nitem1 = ncol * NV, nitem2 = ncol * NC;
  for (int r = lane; r < 2 * NV + 2 * NF; r += nl) (r < 2 * NV ? foldn + r * FWN : foldh + (r - 2 * NV) * FWN)[NQ] = 0.f;   // the zero columns
  if (ft) {
    // table-driven (col, dof) items: the segments are multiples of 32 items, so a warp runs one branch
    const int end_mv = ft->end_mv, end_anc = ft->end_anc, end_vel = ft->end_vel, end_zero = ft->end_zero;
    for (int it = lane; it < end_zero; it += nl) {
      const unsigned ent = ft->item[it];
      if (ent == WB_FOLD_NONE) continue;
      const int col = int(ent >> 5), i = int(ent & 31u);
      float t1 = 0.f, t2 = 0.f, e = 0.f;
      if (it < end_anc) {
        const WbDirW<NV>& d = dw[col - 3];
        const bool mv = it < end_mv;
        if (mv) {
          const float dD = dot6(tr.S[i], tr.g[d.off + md.dof_body[i] - d.b0]);
          t1 = -(dD + dot6(tr.U[i], d.mu[0]));
          t2 = -(dD + dot6(tr.U[i], d.mu[1]));
          e = -dot6(tr.U[i], d.mu[2]);
        } else {
          const float dD = dot6(tr.S[i], d.X[0]);
          t1 = -(dD + dot6(tr.S[i], d.X[1]));
          t2 = -(dD + dot6(tr.S[i], d.X[2]));
          e = -dot6(tr.S[i], d.X[3]);
        }
        for (int c = 0; c < NC; ++c) {  // foot Jacobian terms of this dof, feet in fixed order
          const int bc = md.contact_body[c];
          if (bc < d.b0 || bc >= d.b0 + d.nbd || !((md.contact_dofmask[c] >> i) & 1u)) continue;
          float dJ[3];
          if (mv) {
            cross3(d.Sd, J + i * NF + 3 * c, dJ);
          } else {
            float pc[3];
            cross3(d.Sd, tr.rc[c], pc);
            for (int k = 0; k < 3; ++k) pc[k] += d.Sd[3 + k];
            cross3(tr.S[i], pc, dJ);
          }
          for (int k = 0; k < 3; ++k) {
            t2 = fmaf(dJ[k], grf[3 * c + k], t2);
            e = fmaf(dJ[k], y[3 * c + k], e);
          }
        }
      } else if (it < end_vel) {
        const int D = col - 3, b0 = md.dir_b0[D], nbd = md.dir_nb[D], off = md.dir_off[D];
        const int bi = md.dof_body[i];
        const float s = dot6(tr.S[i], tr.g[(bi >= b0 && bi < b0 + nbd) ? off + bi - b0 : off]);
        t1 = -s; t2 = -s;
      }
      fold[i * FW + col] = t1;
      if (col < NQ) {
        foldn[i * FWN + col] = t2;
        foldn[(NV + i) * FWN + col] = e;
      }
    }
  }
  for (int it = (ft ? nitem1 + lane : lane); it < nitem1 + nitem2; it += nl) {
    if (it < nitem1) {
      const int col = it / NV, i = it - col * NV;
      float t1 = 0.f, t2 = 0.f, e = 0.f;
      if (col >= NQ) {  // velocity direction
        const int D = col - 3, b0 = md.dir_b0[D], nbd = md.dir_nb[D], off = md.dir_off[D];
        const int bi = md.dof_body[i];
        float s = 0.f;
        if (bi >= b0 && bi < b0 + nbd) s = dot6(tr.S[i], tr.g[off + bi - b0]);
        else if ((md.dof_anc[D - npos] >> i) & 1u) s = dot6(tr.S[i], tr.g[off]);
        t1 = -s; t2 = -s;
      } else if (col >= 3) {
        const WbDirW<NV>& d = dw[col - 3];
        const bool mv = (d.moves >> i) & 1u;
        if (mv) {
          const float dD = dot6(tr.S[i], tr.g[d.off + md.dof_body[i] - d.b0]);
          t1 = -(dD + dot6(tr.U[i], d.mu[0]));
          t2 = -(dD + dot6(tr.U[i], d.mu[1]));
          e = -dot6(tr.U[i], d.mu[2]);
        } else if ((d.anc >> i) & 1u) {
          const float dD = dot6(tr.S[i], d.X[0]);
          t1 = -(dD + dot6(tr.S[i], d.X[1]));
          t2 = -(dD + dot6(tr.S[i], d.X[2]));
          e = -dot6(tr.S[i], d.X[3]);
        }
        for (int c = 0; c < NC; ++c) {  // foot Jacobian terms of this dof, feet in fixed order
          const int bc = md.contact_body[c];
          if (bc < d.b0 || bc >= d.b0 + d.nbd || !((md.contact_dofmask[c] >> i) & 1u)) continue;
          float dJ[3];
          if (mv) {
            cross3(d.Sd, J + i * NF + 3 * c, dJ);
          } else {
            float pc[3];
            cross3(d.Sd, tr.rc[c], pc);
            for (int k = 0; k < 3; ++k) pc[k] += d.Sd[3 + k];
            cross3(tr.S[i], pc, dJ);
          }
          for (int k = 0; k < 3; ++k) {
            t2 = fmaf(dJ[k], grf[3 * c + k], t2);
            e = fmaf(dJ[k], y[3 * c + k], e);
          }
        }
      }
      fold[i * FW + col] = t1;
      if (col < NQ) {
        foldn[i * FWN + col] = t2;
        foldn[(NV + i) * FWN + col] = e;
      }
    } else {
      const int it2 = it - nitem1;
      const int col = it2 / NC, c = it2 - col * NC;
      float hc[3] = {0.f, 0.f, 0.f}, df[3] = {0.f, 0.f, 0.f};
      if (col < 3) {
        df[col] = 1.f;
      } else if (col < NQ) {
        const WbDirW<NV>& d = dw[col - 3];
        const int bc = md.contact_body[c];
        if (bc >= d.b0 && bc < d.b0 + d.nbd) {
          float pc[3];
          cross3(d.Sd, tr.rc[c], pc);
          for (int k = 0; k < 3; ++k) { pc[k] += d.Sd[3 + k]; df[k] = pc[k]; }
          const uint32_t mask = md.contact_dofmask[c];
          for (int i = 0; i < NV; ++i) {
            if (!((mask >> i) & 1u)) continue;
            float dJ[3];
            if ((d.moves >> i) & 1u) cross3(d.Sd, J + i * NF + 3 * c, dJ);
            else cross3(tr.S[i], pc, dJ);
            for (int k = 0; k < 3; ++k) hc[k] = fmaf(dJ[k], wvz[i], hc[k]);
          }
        }
      }
      if (col < NQ)
        for (int k = 0; k < 3; ++k) {
          foldh[(3 * c + k) * FWN + col] = hc[k];
          foldh[(NF + 3 * c + k) * FWN + col] = df[k];
        }
    }
  }
}

}  // namespace mpxb
