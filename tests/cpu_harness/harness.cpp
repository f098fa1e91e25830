// harness.cpp -- TEST INFRASTRUCTURE.  Compiles the kernels' __host__ __device__ arithmetic (common.cuh, srbd.cuh,
// wb.cuh) with g++ so that the `-m "not gpu"` tests can compare exactly the code the CUDA kernels run against the
// float64 oracle on a machine without a GPU.  Nothing in the product path loads this library.
#include <vector>

#include "../../mpx_b200/csrc/srbd.cuh"
#include "../../mpx_b200/csrc/wb.cuh"

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
}  // namespace

extern "C" {

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

int h_srbd_dynamics(float mass, float dt, const float* inertia, const float* inertia_inv, const float* x, const float* u,
                    const float* par, float* xn) {
  SrbdParams P;
  P.mass = mass; P.dt = dt;
  for (int i = 0; i < 9; ++i) { P.inertia[i] = inertia[i]; P.inertia_inv[i] = inertia_inv[i]; }
  srbd_dynamics<float>(P, x, u, par, xn);
  return 0;
}
}
