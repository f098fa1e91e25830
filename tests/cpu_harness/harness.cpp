// harness.cpp -- TEST INFRASTRUCTURE.  Compiles the kernels' __host__ __device__ arithmetic (common.cuh, srbd.cuh,
// wb.cuh) with g++ so that the `-m "not gpu"` tests can compare exactly the code the CUDA kernels run against the
// float64 oracle on a machine without a GPU.  Nothing in the product path loads this library.
#include <vector>

#include "../../mpx_b200/csrc/srbd.cuh"
#include "../../mpx_b200/csrc/wb.cuh"
#include "../../mpx_b200/csrc/wb_tree.cuh"

using namespace mpxb;

namespace {
template <int NB, int NV, int NC>
void dyn_t(const WbModel& md, const float* x, const float* u, const float* contact, float* xn, float* A, float* B) {
  const int n = md.n, m = md.m;
  wb_dynamics<float, NB, NV, NC>(md, x, u, contact, xn);
  if (!A) return;
  std::vector<Dual> xd(n), ud(m), xo(n);
  for (int d = 0; d < n + m; ++d) {
    for (int i = 0; i < n; ++i) xd[i] = Dual(x[i], i == d ? 1.f : 0.f);
    for (int i = 0; i < m; ++i) ud[i] = Dual(u[i], n + i == d ? 1.f : 0.f);
    wb_dynamics<Dual, NB, NV, NC>(md, xd.data(), ud.data(), contact, xo.data());
    for (int i = 0; i < n; ++i) {
      if (d < n) A[i * n + d] = xo[i].d;
      else B[i * m + d - n] = xo[i].d;
    }
  }
}
template <int NJ, int NC>
void cost_t(const WbModel& md, const float* W, const float* ref, const float* x, const float* u, int terminal,
            float* cost, float* grad) {
  const int n = md.n, m = md.m;
  *cost = wb_cost<float, NJ, NC>(md, W, ref, x, u, terminal != 0);
  std::vector<Dual> xd(n), ud(m);
  for (int d = 0; d < n + m; ++d) {
    for (int i = 0; i < n; ++i) xd[i] = Dual(x[i], i == d ? 1.f : 0.f);
    for (int i = 0; i < m; ++i) ud[i] = Dual(u[i], n + i == d ? 1.f : 0.f);
    grad[d] = wb_cost<Dual, NJ, NC>(md, W, ref, xd.data(), ud.data(), terminal != 0).d;
  }
}
// x+ and [A | B] through the v2 path of the linearise kernel: primal path-stack pass, cooperative contact solve
// (run by one worker here), one forward-mode path-stack pass per state direction folded by WbTangentSink, and the
// analytic contact-solve derivative wb_s3_tangent.
template <int NV, int NC>
void hybrid_t(const WbModel& md, const float* x, const float* u, const float* contact, float* xn, float* A, float* B) {
  constexpr int NF = 3 * NC;
  const int n = md.n, m = md.m, nj = md.nj, nv = md.nv, nq = 7 + nj, nf = 3 * md.nc;
  static WbKnot<NV, NC> k;
  static float stack[WB_MAXB * WB_SLOT];
  static Dual dstack[WB_MAXB * WB_SLOT];
  for (int i = 0; i < NV * NV; ++i) k.M[i] = 0.f;
  for (int i = 0; i < NV * NF; ++i) k.J[i] = 0.f;
  WbStackPlain<float> stk{stack};
  WbStoreSink<float> sink{k.M, NV, k.D, k.J, NF, k.feet};
  WbInPlain<float> in{x, x + nq};
  wb_dfs<float>(md, in, stk, sink);
  wb_s3_primal<NV, NC>(md, k, x, u, contact, 0, 1, [] {});
  for (int i = 0; i < n; ++i) xn[i] = k.xn[i];
  std::vector<float> col(n);
  for (int d = 0; d < nq + nv + m; ++d) {
    float t1[NV] = {0}, t2[NV] = {0}, e[NV] = {0}, h[NF] = {0}, dfeet[NF] = {0};
    const int is_u = d >= nq + nv;
    if (!is_u) {
      WbStackPlain<Dual> ds{dstack};
      WbTangentSink ts{t1, t2, e, h, dfeet, k.w, k.z, k.a, k.y, k.grf, k.wvz};
      WbInSeed seed{x, nq, d};
      wb_dfs<Dual>(md, seed, ds, ts);
    }
    float* cp = col.data();
    wb_s3_tangent<NV, NC>(md, k, contact, is_u, is_u ? d - nq - nv : d, t1, t2, e, h, dfeet,
                          [cp](int i, float v) { cp[i] = v; });
    for (int i = 0; i < n; ++i) {
      if (!is_u) A[i * n + d] = col[i];
      else B[i * m + d - nq - nv] = col[i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = nq + nv; j < n; ++j) A[i * n + j] = 0.f;
}

// x+ and [A | B] through the v3 path of the linearise kernel: per-body tree quantities (wb_tree_*), the cooperative
// contact solve, ANALYTIC folded tangents (wb_tree_fold) and wb_s3_tangent.  One worker.
template <int NB, int NV, int NC, int NI>
int tree_t(const WbModel& md, const float* x, const float* u, const float* contact, float* xn, float* A, float* B) {
  const int n = md.n, m = md.m, nj = md.nj, nv = md.nv, nq = 7 + nj;
  if (!md.tree_ok || md.nitems > NI) return -3;
  static WbKnot<NV, NC> k;
  static WbTree<NB, NV, NC, NI> tr;
  auto nosync = [] {};
  const float* qpos = x;
  const float* qvel = x + nq;
  wb_tree_fk(md, tr, qpos, qvel, 0, 1);
  wb_tree_bodies(md, tr, qpos, k.feet, 0, 1);
  wb_tree_composite(md, tr, 0, 1);
  wb_tree_dofs(md, tr, k.D, 0, 1);
  wb_tree_mass_jac(md, tr, k.M, k.J, 0, 1);
  if constexpr (NV <= MPXB_INV_MAXNV) wb_s3_primal_inv<NV, NC>(md, k, x, u, contact, 0, 1, nosync);
  else wb_s3_primal<NV, NC>(md, k, x, u, contact, 0, 1, nosync);
  for (int i = 0; i < n; ++i) xn[i] = k.xn[i];
  wb_tree_acc(md, tr, k.w, k.a, k.z, 0, 1, nosync);
  wb_tree_items(md, tr, 0, 1, nosync);
  constexpr int NF = 3 * NC, FW = 2 * NV + 1, FWN = NV + 2, SZN = 2 * NV * FWN, SZH = 2 * NF * FWN;
  static WbDirW<NV> dw[NV + 1];
  // t1 of every x-direction in `fold`; t2 | e (foldn) and h | dfeet (foldh) of the position directions (+ the zero column
  // nq) in the narrow blocks -- placed over the tree's scratch block where they fit, exactly as the linearise kernel
  // does, so the host run also proves that nothing reads the scratch block after the direction frames are built
  static float fold[NV * FW], fold_plain[NV * FW], foldn_own[SZN], foldh_own[SZH], foldn_plain[SZN], foldh_plain[SZH];
  constexpr int SCR = WbTree<NB, NV, NC, NI>::SCRATCH;
  constexpr bool n_over = SZN <= SCR, h_over = n_over && SZN + SZH <= SCR;
  float* foldn = n_over ? tr.scratch() : foldn_own;
  float* foldh = h_over ? tr.scratch() + SZN : foldh_own;
  // the device runs the table-driven item order (wb_build_fold_table); both enumerations must fill the blocks identically
  static WbFoldTable ft;
  wb_build_fold_table(md, &ft);
  static WbTree<NB, NV, NC, NI> tr_copy;
  tr_copy = tr;
  wb_tree_fold_coop(md, tr_copy, k.J, k.w, k.a, k.z, k.y, k.grf, k.wvz, dw, fold_plain, FW, foldn_plain, foldh_plain, FWN, 0, 1, nosync);
  wb_tree_fold_coop(md, tr, k.J, k.w, k.a, k.z, k.y, k.grf, k.wvz, dw, fold, FW, foldn, foldh, FWN, 0, 1, nosync, &ft);
  for (int r = 0; r < NV; ++r)
    for (int c2 = 0; c2 < nq + nv; ++c2)
      if (!(fold[r * FW + c2] == fold_plain[r * FW + c2])) return -100;
  for (int c2 = 0; c2 <= nq; ++c2) {
    for (int r = 0; r < 2 * NV; ++r)
      if (!(foldn[r * FWN + c2] == foldn_plain[r * FWN + c2])) return -101;
    for (int r = 0; r < 2 * NF; ++r)
      if (!(foldh[r * FWN + c2] == foldh_plain[r * FWN + c2])) return -102;
  }
  std::vector<float> col(n);
  for (int d = 0; d < nq + nv + m; ++d) {
    float t1[NV] = {0};
    const int is_u = d >= nq + nv;
    const int cc = is_u ? 0 : d;
    if (!is_u) for (int i = 0; i < NV; ++i) t1[i] = fold[i * FW + d];
    const bool posdir = !is_u && cc < nq;
    const int cn = posdir ? cc : nq;
    const WbStrided t2 = posdir ? WbStrided{foldn + cc, FWN} : WbStrided{fold + cc, FW};
    const WbStrided e{foldn + NV * FWN + cn, FWN}, h{foldh + cn, FWN}, dfeet{foldh + NF * FWN + cn, FWN};
    float* cp = col.data();
    if constexpr (NV <= MPXB_INV_MAXNV) wb_s3_tangent_inv<NV, NC>(md, k, contact, is_u, is_u ? d - nq - nv : d, t1, t2, e, h, dfeet,
                          [cp](int i, float v) { cp[i] = v; });
    else wb_s3_tangent<NV, NC>(md, k, contact, is_u, is_u ? d - nq - nv : d, t1, t2, e, h, dfeet,
                          [cp](int i, float v) { cp[i] = v; });
    for (int i = 0; i < n; ++i) {
      if (!is_u) A[i * n + d] = col[i];
      else B[i * m + d - nq - nv] = col[i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = nq + nv; j < n; ++j) A[i * n + j] = 0.f;
  return 0;
}

}  // namespace

extern "C" {

int h_wb_dynamics_tree(const mpxb_model_desc* d, const float* x, const float* u, const float* contact, float* xn,
                       float* A, float* B) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  if (md.nv == 18 && md.nc == 4) return tree_t<14, 18, 4, 192>(md, x, u, contact, xn, A, B);
  if (md.nv == 25 && md.nc == 4) return tree_t<21, 25, 4, 320>(md, x, u, contact, xn, A, B);
  if (md.nv == 28 && md.nc == 8) return tree_t<32, 28, 8, 640>(md, x, u, contact, xn, A, B);
  return -2;
}

// tables of the analytic linearisation, for inspection
int h_wb_tree_info(const mpxb_model_desc* d, int* out) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  out[0] = md.tree_ok; out[1] = md.nleaf; out[2] = md.ndir; out[3] = md.nitems; out[4] = md.maxdepth;
  return 0;
}

// x+ = f(x, u) and (optionally) A = df/dx [n,n], B = df/du [n,m] by forward mode, as the linearise kernel does
int h_wb_dynamics(const mpxb_model_desc* d, const float* x, const float* u, const float* contact, float* xn, float* A,
                  float* B) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  if (md.nbody <= 14 && md.nv <= 18 && md.nc <= 4) dyn_t<14, 18, 4>(md, x, u, contact, xn, A, B);
  else if (md.nbody <= 21 && md.nv <= 25 && md.nc <= 4) dyn_t<21, 25, 4>(md, x, u, contact, xn, A, B);
  else dyn_t<32, 28, 8>(md, x, u, contact, xn, A, B);
  return 0;
}

int h_wb_cost(const mpxb_model_desc* d, const float* W, const float* ref, const float* x, const float* u, int terminal,
              float* cost, float* grad) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  if (md.nv <= 18 && md.nc <= 4) cost_t<12, 4>(md, W, ref, x, u, terminal, cost, grad);
  else if (md.nv <= 25 && md.nc <= 4) cost_t<19, 4>(md, W, ref, x, u, terminal, cost, grad);
  else cost_t<22, 8>(md, W, ref, x, u, terminal, cost, grad);
  return 0;
}

// mass matrix, bias force, contact Jacobian and foot positions at (qpos, qvel) -- the MJX quantities the dynamics read
int h_wb_terms(const mpxb_model_desc* d, const float* qpos, const float* qvel, float* M, float* bias, float* J,
               float* feet) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  constexpr int NB = 32, NV = 28, NC = 8;
  static float xpos[NB][3], xquat[NB][4], cinert[NB][10], cdof[NV][6], crb[NB][10], cvel[NB][6], cacc[NB][6], cfrc[NB][6];
  float com[3];
  wb_geometry<float, NB, NV, NC>(md, qpos, xpos, xquat, cinert, cdof, com, feet);
  wb_mass_matrix<float, NB, NV>(md, cinert, cdof, crb, M, md.nv);
  wb_bias<float, float, float, NB, NV>(md, cinert, cdof, qvel, cvel, cacc, cfrc, bias);
  wb_contact_jacobian<float, NV, NC>(md, cdof, com, feet, J, 3 * md.nc);
  return 0;
}

// the same terms through the single-pass path-stack evaluation (wb_dfs) that the v2 linearise kernel uses
int h_wb_terms_dfs(const mpxb_model_desc* d, const float* qpos, const float* qvel, float* M, float* bias, float* J,
                   float* feet) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  static float stack[WB_MAXB * WB_SLOT];
  const int nv = md.nv, nf = 3 * md.nc;
  for (int i = 0; i < nv * nv; ++i) M[i] = 0.f;
  for (int i = 0; i < nv * nf; ++i) J[i] = 0.f;
  WbStackPlain<float> stk{stack};
  WbStoreSink<float> sink{M, nv, bias, J, nf, feet};
  WbInPlain<float> in{qpos, qvel};
  wb_dfs<float>(md, in, stk, sink);
  return 0;
}

int h_wb_dynamics_hybrid(const mpxb_model_desc* d, const float* x, const float* u, const float* contact, float* xn,
                         float* A, float* B) {
  WbModel md;
  int rc = wb_build_model(*d, &md);
  if (rc) return rc;
  if (md.nv == 18 && md.nc == 4) hybrid_t<18, 4>(md, x, u, contact, xn, A, B);
  else if (md.nv == 25 && md.nc == 4) hybrid_t<25, 4>(md, x, u, contact, xn, A, B);
  else if (md.nv == 28 && md.nc == 8) hybrid_t<28, 8>(md, x, u, contact, xn, A, B);
  else return -2;
  return 0;
}

int h_srbd_dynamics(float mass, float dt, const float* inertia, const float* inertia_inv, const float* x, const float* u,
                    const float* par, float* xn) {
  SrbdParams P;
  P.mass = mass; P.dt = dt;
  for (int i = 0; i < 9; ++i) { P.inertia[i] = inertia[i]; P.inertia_inv[i] = inertia_inv[i]; }
  srbd_dynamics<float>(P, x, u, par, xn);
  return 0;
}
}
