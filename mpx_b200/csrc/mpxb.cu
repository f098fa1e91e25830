// mpxb.cu -- C ABI of libmpxb.so (include/mpxb.h): model handles, workspace layout, and the stream-ordered
// launch sequence of one primal-dual iLQR iteration:
//   K1 <model>_linearize : evaluate + linearise + quadratise, one warp per knot per env   (optimizers.py:625,150-173)
//   K2 lq_seq / lq_scan  : Riccati backward pass, rollout, dual update                     (primal_tvlqr.py, dual_tvlqr.py)
//   K3 <model>_trial     : 11 trial points of the filter line search                       (optimizers.py:518-540)
//   K4 select_update     : filter decision, iterate update, convergence flag               (optimizers.py:541-578)
// There is no host synchronisation and no allocation on this path.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "../../include/mpxb.h"
#include "lq.cuh"
#include "lq_small.cuh"
#include "sqp.cuh"
#include "srbd_kernels.cuh"
#include "wb_launch.h"
#include "lq_scan.cuh"
#include "lq_pscan.cuh"
#include "lq3.cuh"
#include "refgen.cuh"

using namespace mpxb;

struct mpxb_profile {
  int max_calls, calls;
  cudaEvent_t* ev;  // 5 per call
};

struct mpxb_model {
  mpxb_model_desc desc;
  mpxb_profile* prof;
  SrbdParams srbd;
  WbModel* wb_dev;  // device copy of the whole-body tables (null for SRBD)
  int eager_line_search;  // 1: evaluate all 11 trial points every call (default 0: lazy, same result)
  int merit_line_search;  // MPXB_LS_*: 0 filter, 1 parallel merit search, 2 sequential Armijo semantics (legacy slope)
  int stop_rule;          // MPXB_STOP_*
  int structured;         // MPXB_OPT_COST_STRUCTURE: 1 = W is diagonal (checked by the host) -> structured LQ path
  bool s3_ok;             // the structured kernels are compiled for this model's dimensions
  Lq3Layout s3;
  WbModel wb_host;
  int device;
  // LQ tail / line-search overlap (mpxb_mpc_step): a side stream and two events owned by the handle, SM count
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int sm_count;
};

// Entry points may be called with another device current (one handle per GPU, several GPUs per process): launch on
// the model's device and restore the caller's.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

namespace {

constexpr size_t MPXB_LQ_SMALL_MAX_BYTES = 232448 - 1024;  // 227 KB of shared memory per CTA on sm_100

template <int n, int m> cudaError_t launch_lq_seq(const LqArgs& a, cudaStream_t st) {
  constexpr int NT = 256;
  if constexpr (n <= 16 && m <= 16) {
    // small systems (SRBD): the whole horizon in shared memory, no global traffic inside the time loop (lq_small.cuh);
    // the streaming kernel below takes over when the horizon does not fit.  MPXB_LQ_SMALL=0: measurement switch
    static const int off = [] { const char* v = getenv("MPXB_LQ_SMALL"); return (v && v[0] == '0') ? 1 : 0; }();
    const size_t bytes = lq_small_bytes<n, m>(a.N);
    // one CTA per SM: up to ~6 waves it beats the streaming kernel's four resident CTAs (E = 296: 0.14 vs 0.24 ms)
    if (!off && bytes <= MPXB_LQ_SMALL_MAX_BYTES && a.E <= 6 * 148) {
      lq_small_kernel<n, m, NT><<<a.E, NT, bytes, st>>>(a);
      return cudaGetLastError();
    }
  }
  auto kern = lq_seq_kernel<n, m, NT>;
  kern<<<a.E, NT, lq_rt_bytes<n, m>(a.nzc), st>>>(a);
  return cudaGetLastError();
}
// > 48 KB of dynamic shared memory is a per-device function attribute.  It is set for every LQ kernel of a problem
// size at model creation (and by the stand-alone mpxb_tvlqr), never inside a launch sequence: the library keeps no
// mutable global state and a captured launch sequence contains launches only.
template <int n, int m> cudaError_t configure_lq() {
  cudaError_t err = cudaFuncSetAttribute(lq_seq_kernel<n, m, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lq_rt_bytes<n, m>(0)));
  if constexpr (n <= 16 && m <= 16) {
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(lq_small_kernel<n, m, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(MPXB_LQ_SMALL_MAX_BYTES));
  }
  if (err == cudaSuccess) err = configure_lq_scan<n, m>();
  if (err == cudaSuccess) err = configure_lq_pscan<n, m>();
  return err;
}
int configure_lq_dispatch(int n, int m) {
#define MPXB_LQ_CASE(NN, MM) if (n == NN && m == MM) return configure_lq<NN, MM>() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
  MPXB_LQ_CASE(13, 12)
  MPXB_LQ_CASE(61, 12)
  MPXB_LQ_CASE(75, 19)
  MPXB_LQ_CASE(81, 46)
  MPXB_LQ_CASE(2, 1)
  MPXB_LQ_CASE(4, 1)
  MPXB_LQ_CASE(8, 2)
#undef MPXB_LQ_CASE
  return MPXB_EUNSUPPORTED;
}

// structured whole-body LQ kernel (lq3.cuh): quadrupeds, H1, Talos
template <int NH, int NTL, int M, int NCONE, int NT> cudaError_t launch_lq3(const Lq3Args& a, cudaStream_t st) {
  using D = Lq3Dims<NH, NTL, M, NCONE>;
  auto kern = lq3_kernel<NH, NTL, M, NCONE, NT>;
  if (a.E == 0)  // configure only (mpxb_model_create)
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Lq3Smem<D>::bytes));
  kern<<<a.E, NT, Lq3Smem<D>::bytes, st>>>(a);
  return cudaGetLastError();
}
bool lq3_supported(const Lq3Layout& s) {
  return (s.nh == 37 && s.ntl == 24 && s.m == 12 && s.ncone == 4) || (s.nh == 51 && s.ntl == 24 && s.m == 19 && s.ncone == 0) ||
         (s.nh == 57 && s.ntl == 24 && s.m == 46 && s.ncone == 0);
}
// threads per CTA of the structured LQ kernel: the first of NTS... is the default, MPXB_LQ3_NT=<n> selects another
// instantiated count (measurement switch); a.E == 0 configures every instantiation (mpxb_model_create)
template <int NH, int NTL, int M, int NCONE, int... NTS> cudaError_t launch_lq3_nt(const Lq3Args& a, cudaStream_t st) {
  static const int want = [] { const char* v = getenv("MPXB_LQ3_NT"); return v ? atoi(v) : 0; }();
  constexpr int nts[] = {NTS...};
  int pick = nts[0];
  for (int x : nts) if (x == want) pick = x;
  cudaError_t err = cudaSuccess;
  if (a.E == 0) {
    ((err = (err == cudaSuccess) ? launch_lq3<NH, NTL, M, NCONE, NTS>(a, st) : err), ...);
    return err;
  }
  ((pick == NTS ? (void)(err = launch_lq3<NH, NTL, M, NCONE, NTS>(a, st)) : (void)0), ...);
  return err;
}
int dispatch_lq3(const Lq3Layout& s, const Lq3Args& a, cudaStream_t st) {
  cudaError_t err = cudaErrorInvalidValue;
  // Go2 / Aliengo: 4 CTAs of 128 threads per SM (51 KB each; x1024: 0.888 ms, 96 threads 0.893, 160 threads 0.951).
  // H1 (94 KB, 2 CTAs per SM): 192 threads at 168 registers, x4096: 7.32 ms vs 7.56 (160), 8.05 (256 threads at the
  // 128-register cap), 8.20 (128) and 8.27 (224), r02y
  if (s.nh == 37 && s.m == 12) err = launch_lq3_nt<37, 24, 12, 4, 128, 96>(a, st);
  else if (s.nh == 51 && s.m == 19) err = launch_lq3_nt<51, 24, 19, 0, 192, 128, 256>(a, st);
  else if (s.nh == 57 && s.m == 46) err = launch_lq3<57, 24, 46, 0, 128>(a, st);
  else return MPXB_EUNSUPPORTED;
  return err == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

// resident CTAs per SM of the structured LQ kernel a layout dispatches to (shared-memory bound)
int lq3_resident(const Lq3Layout& s) {
  if (s.nh == 37 && s.m == 12) return lq3_min_blocks<Lq3Dims<37, 24, 12, 4>>();
  if (s.nh == 51 && s.m == 19) return lq3_min_blocks<Lq3Dims<51, 24, 19, 0>>();
  return 1;
}

int dispatch_lq_seq(int n, int m, const LqArgs& a, cudaStream_t st) {
  cudaError_t err;
#define MPXB_LQ_CASE(NN, MM) if (n == NN && m == MM) { err = launch_lq_seq<NN, MM>(a, st); return err == cudaSuccess ? MPXB_OK : MPXB_ECUDA; }
  MPXB_LQ_CASE(13, 12)
  MPXB_LQ_CASE(61, 12)
  MPXB_LQ_CASE(75, 19)
  MPXB_LQ_CASE(81, 46)
  MPXB_LQ_CASE(2, 1)
  MPXB_LQ_CASE(4, 1)
  MPXB_LQ_CASE(8, 2)
#undef MPXB_LQ_CASE
  return MPXB_EUNSUPPORTED;
}

bool lq_supported(int n, int m) {
  const int s[][2] = {{13, 12}, {61, 12}, {75, 19}, {81, 46}, {2, 1}, {4, 1}, {8, 2}};
  for (auto& p : s) if (p[0] == n && p[1] == m) return true;
  return false;
}

template <int NB, int NV, int NC>
__global__ void wb_feet_kernel(const WbModel* __restrict__ mdp, int E, const float* __restrict__ qpos, float* __restrict__ feet) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const WbModel& md = *mdp;
  float xpos[NB][3], xquat[NB][4], cinert[NB][10], cdof[NV][6], com[3], f[3 * NC];
  wb_geometry<float, NB, NV, NC>(md, qpos + size_t(e) * (7 + md.nj), xpos, xquat, cinert, cdof, com, f);
  for (int i = 0; i < 3 * md.nc; ++i) feet[size_t(e) * 3 * md.nc + i] = f[i];
}

}  // namespace

extern "C" {

const char* mpxb_version_string(void) { return "mpxb 0.1 (sm_100a)"; }

int mpxb_model_create(const mpxb_model_desc* d, mpxb_model** out) {
  if (!d || !out) return MPXB_EINVAL;
  if (d->version != MPXB_VERSION) return MPXB_EINVAL;
  if (d->kind < MPXB_KIND_SRBD || d->kind > MPXB_KIND_TALOS_WB) return MPXB_EINVAL;
  if (d->n <= 0 || d->m <= 0 || d->N <= 0) return MPXB_EINVAL;
  if (!lq_supported(d->n, d->m)) return MPXB_EUNSUPPORTED;
  mpxb_model* mdl = new (std::nothrow) mpxb_model();
  if (!mdl) return MPXB_EINVAL;
  mdl->desc = *d;
  mdl->wb_dev = nullptr;
  mdl->prof = nullptr;
  mdl->eager_line_search = 0;
  mdl->merit_line_search = 0;
  mdl->stop_rule = MPXB_STOP_OFFLINE;
  mdl->structured = 0;
  mdl->s3_ok = false;
  if (d->kind != MPXB_KIND_SRBD) {
    const int nh = 13 + 2 * d->n_joints;
    mdl->s3 = lq3_layout(nh, d->n - nh, d->m, d->kind == MPXB_KIND_QUAD_WB ? d->n_contact : 0);
    mdl->s3_ok = d->n > nh && lq3_supported(mdl->s3);
  }
  cudaGetDevice(&mdl->device);
  mdl->side = nullptr; mdl->ev_fork = nullptr; mdl->ev_join = nullptr; mdl->sm_count = 148;
  cudaDeviceGetAttribute(&mdl->sm_count, cudaDevAttrMultiProcessorCount, mdl->device);
  if (d->kind != MPXB_KIND_SRBD) {
    if (cudaStreamCreateWithFlags(&mdl->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&mdl->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&mdl->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      mpxb_model_destroy(mdl);
      return MPXB_ECUDA;
    }
  }
  if (d->kind == MPXB_KIND_SRBD) {
    if (d->n != SRBD_N || d->m != SRBD_M || d->nref != SRBD_NREF || d->npar != SRBD_NPAR || d->nw != SRBD_NW) {
      delete mdl;
      return MPXB_EINVAL;
    }
    mdl->srbd.mass = d->mass;
    mdl->srbd.dt = d->dt;
    memcpy(mdl->srbd.inertia, d->inertia, sizeof(float) * 9);
    memcpy(mdl->srbd.inertia_inv, d->inertia_inv, sizeof(float) * 9);
  } else {
    int rc = wb_build_model(*d, &mdl->wb_host);
    if (rc != MPXB_OK) { delete mdl; return rc; }
    // device image: the model tables, then the work-item table of the folded-tangent stage (wb_build_fold_table)
    WbFoldTable* ft = new WbFoldTable;
    wb_build_fold_table(mdl->wb_host, ft);
    bool up = cudaMalloc(&mdl->wb_dev, sizeof(WbModel) + sizeof(WbFoldTable)) == cudaSuccess;
    if (!up) { delete ft; delete mdl; return MPXB_ECUDA; }
    up = cudaMemcpy(mdl->wb_dev, &mdl->wb_host, sizeof(WbModel), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(mdl->wb_dev + 1, ft, sizeof(WbFoldTable), cudaMemcpyHostToDevice) == cudaSuccess;
    delete ft;
    if (!up) {
      cudaFree(mdl->wb_dev);
      delete mdl;
      return MPXB_ECUDA;
    }
  }
  // per-device function attributes of every kernel this model can launch
  int rc = configure_lq_dispatch(d->n, d->m);
  if (rc == MPXB_OK && d->kind != MPXB_KIND_SRBD) rc = wb_launch_configure(mdl->wb_host);
  if (rc == MPXB_OK && mdl->s3_ok) {
    Lq3Args none{};
    rc = dispatch_lq3(mdl->s3, none, nullptr);
  }
  if (rc != MPXB_OK) { mpxb_model_destroy(mdl); return rc; }
  *out = mdl;
  return MPXB_OK;
}

int mpxb_model_set_option(mpxb_model* m, int option, int value) {
  if (!m) return MPXB_EINVAL;
  if (option == MPXB_OPT_EAGER_LINE_SEARCH) { m->eager_line_search = value ? 1 : 0; return MPXB_OK; }
  if (option == MPXB_OPT_COST_STRUCTURE) {
    if (value != MPXB_COST_DENSE && value != MPXB_COST_DIAGONAL_W) return MPXB_EINVAL;
    m->structured = (value == MPXB_COST_DIAGONAL_W);
    return MPXB_OK;
  }
  if (option == MPXB_OPT_LINE_SEARCH) {
    if (value != MPXB_LS_FILTER && value != MPXB_LS_MERIT && value != MPXB_LS_ARMIJO) return MPXB_EINVAL;
    m->merit_line_search = value;
    return MPXB_OK;
  }
  if (option == MPXB_OPT_STOP_RULE) {
    if (value != MPXB_STOP_OFFLINE && value != MPXB_STOP_LEGACY) return MPXB_EINVAL;
    m->stop_rule = value;
    return MPXB_OK;
  }
  return MPXB_EINVAL;
}

void mpxb_model_destroy(mpxb_model* m) {
  if (!m) return;
  if (m->wb_dev) cudaFree(m->wb_dev);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->side) cudaStreamDestroy(m->side);
  delete m;
}

// MPXB_LQ_SCAN picks the schedule of the associative scan by batch size: the parallel (Hillis-Steele, one CTA per
// knot) schedule up to PSCAN_MAX_ENVS environments, the serial one-CTA-per-environment schedule above;
// MPXB_LQ_SCAN_SERIAL / MPXB_LQ_SCAN_PARALLEL force one (the parallel one only where its scratch is provided for).
static bool lq_mode_valid(int lq_mode, int E) {
  if (lq_mode == MPXB_LQ_SEQUENTIAL || lq_mode == MPXB_LQ_SCAN || lq_mode == MPXB_LQ_SCAN_SERIAL) return true;
  return lq_mode == MPXB_LQ_SCAN_PARALLEL && pscan_applies(E);
}
static size_t lq_scan_ws_floats(int lq_mode, int E, int N, int n) {
  const bool par = (lq_mode == MPXB_LQ_SCAN || lq_mode == MPXB_LQ_SCAN_PARALLEL) && pscan_applies(E);
  return par ? lq_pscan_floats(E, N, n) : 0;
}
static int dispatch_lq(int lq_mode, int n, int m, const LqArgs& la, float* scan_ws, cudaStream_t st) {
  if (lq_mode == MPXB_LQ_SEQUENTIAL) return dispatch_lq_seq(n, m, la, st);
  if (lq_mode == MPXB_LQ_SCAN_PARALLEL || (lq_mode == MPXB_LQ_SCAN && pscan_applies(la.E)))
    return dispatch_lq_pscan(n, m, la, scan_ws, st);
  return dispatch_lq_scan(n, m, la, scan_ws, st);
}

// MPXB_SPLIT_LQ=0 switches the LQ tail / line-search overlap off (measurement switch)
static bool split_enabled() {
  static const bool on = [] { const char* v = getenv("MPXB_SPLIT_LQ"); return !(v && v[0] == '0'); }();
  return on;
}

static WsLayout layout_of(const mpxb_model* mdl, int E, int lq_mode) {
  const mpxb_model_desc& d = mdl->desc;
  size_t scan = lq_scan_ws_floats(lq_mode, E, d.N, d.n);
  return make_layout(E, d.N, d.n, d.m, scan, mdl->s3_ok ? &mdl->s3 : nullptr);
}
// the structured path serves limited_memory=True (tvlqr + rollout) once the host has vouched for a diagonal W
static bool use_structured(const mpxb_model* mdl, int lq_mode) {
  return mdl->structured && mdl->s3_ok && lq_mode == MPXB_LQ_SEQUENTIAL;
}

size_t mpxb_workspace_bytes(const mpxb_model* mdl, int E, int lq_mode) {
  if (!mdl || E <= 0) return 0;
  return layout_of(mdl, E, lq_mode).total * sizeof(float);
}

int mpxb_mpc_step(const mpxb_model* mdl, void* stream, int E, int lq_mode, int max_iter, const float* reference,
                  const float* parameter, const float* W, size_t w_env_stride, const float* x0, const float* X_in,
                  const float* U_in, const float* V_in, float* X, float* U, float* V, int32_t* alpha_index,
                  int32_t* iters, int32_t* status, void* workspace, size_t workspace_bytes) {
  if (!mdl || !reference || !parameter || !W || !x0 || !X_in || !U_in || !V_in || !X || !U || !V || !workspace)
    return MPXB_EINVAL;
  if (E <= 0 || max_iter <= 0) return MPXB_EINVAL;
  if (mdl->stop_rule == MPXB_STOP_LEGACY && mdl->merit_line_search == MPXB_LS_FILTER) return MPXB_EINVAL;
  if (!lq_mode_valid(lq_mode, E)) return MPXB_EINVAL;
  if (X == X_in || U == U_in || V == V_in) return MPXB_EINVAL;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return MPXB_EINVAL;  // bulk copies need 16-byte aligned blocks
  DeviceGuard guard(mdl->device);
  const mpxb_model_desc& d = mdl->desc;
  const bool structured = use_structured(mdl, lq_mode);
  const Lq3Layout& s3 = mdl->s3;
  const WsLayout L = layout_of(mdl, E, lq_mode);
  if (workspace_bytes < L.total * sizeof(float)) return MPXB_EWORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  const int N = d.N, n = d.n, m = d.m;
  int32_t* done = (max_iter > 1) ? reinterpret_cast<int32_t*>(ws + L.done) : nullptr;
  if (done) cudaMemsetAsync(done, 0, sizeof(int32_t) * E, st);

  // ping-pong so that the last iteration writes the caller's X/U/V
  const float *Xc = X_in, *Uc = U_in, *Vc = V_in;
  for (int it = 0; it < max_iter; ++it) {
    const bool to_user = ((max_iter - 1 - it) % 2) == 0;
    float* Xn = to_user ? X : ws + L.Xt;
    float* Un = to_user ? U : ws + L.Ut;
    float* Vn = to_user ? V : ws + L.Vt;
    ModelIO io{E, N, reference, parameter, W, w_env_stride, x0, Xc, Uc, Vc, done, 0};
    LinOut lo{ws + L.A, ws + L.B, ws + L.Q, ws + L.R, ws + L.M, ws + L.q, ws + L.r, ws + L.c, ws + L.cost,
              n, m, 0, 0, 0, nullptr, 0, 0, 0, 0, 0, 0, 0, 0};
    if (structured) {
      lo.A = ws + L.ABc; lo.B = ws + L.ABc + s3.LH; lo.ldA = s3.LAB; lo.ldB = s3.LAB; lo.compact = 1; lo.PK = ws + L.PK;
      lo.padA = s3.LH - s3.nh; lo.padB = s3.LM8 - s3.m;
      lo.pks = s3.PKS; lo.oQd = s3.oQd; lo.oq = s3.oq; lo.oc = s3.oc; lo.oQq = s3.oQq; lo.oQc = s3.oQc; lo.oRd = s3.oRd;
      lo.oR = s3.oR;
    }
    int32_t* need = reinterpret_cast<int32_t*>(ws + L.need);
    TrialOut to{ws + L.dX, ws + L.dU, ws + L.ls_theta, ws + L.ls_phi, ws + L.c, ws + L.cost, ws + L.slope,
                ws + L.theta_k, ws + L.phi_k, need, 0, NALPHA, mdl->merit_line_search, ws + L.dV, ws + L.ls_vc,
                ws + L.vc_k, ws + L.dv2, ws + L.dvc, ws + L.th_part, ws + L.ph_part, ws + L.vc_part,
                reinterpret_cast<int32_t*>(ws + L.need_count), 0, 0, 0};
    if (!mdl->eager_line_search) cudaMemsetAsync(to.need_count, 0, sizeof(int32_t), st);
    mpxb_profile* pf = (it == max_iter - 1 && mdl->prof && mdl->prof->calls < mdl->prof->max_calls) ? mdl->prof : nullptr;
    cudaEvent_t* pev = pf ? pf->ev + 5 * pf->calls : nullptr;
    if (pev) cudaEventRecord(pev[0], st);
    // K1
    if (d.kind == MPXB_KIND_SRBD) {
      const int wpb = 4;
      const int blocks = (E * (N + 1) + wpb - 1) / wpb;
      srbd_linearize_kernel<<<blocks, wpb * 32, 0, st>>>(mdl->srbd, io, lo);
    } else {
      int rc = wb_launch_linearize(mdl->wb_host, mdl->wb_dev, io, lo, st);
      if (rc != MPXB_OK) return rc;
    }
    if (pev) cudaEventRecord(pev[1], st);
    // K2
    LqArgs la{E, N, ws + L.A, ws + L.B, ws + L.Q, ws + L.R, ws + L.M, ws + L.q, ws + L.r, ws + L.c,
              ws + L.K, ws + L.k, ws + L.P, ws + L.p, ws + L.dX, ws + L.dU, ws + L.dV, ws + L.slope, done,
              // whole-body models: the foot / GRF tail of the state never enters f, so those columns of A are zero
              d.kind == MPXB_KIND_SRBD ? 0 : 13 + 2 * d.n_joints};
    int rc;
    // LQ tail / line-search overlap: one CTA per environment over `slots` resident CTAs leaves the last wave partly
    // empty (Go2 x1024: 592 + 432).  The full waves (slice A) and the remainder (slice B) are launched separately, and
    // the alpha = 1 trial points of slice A run on the handle's side stream while slice B's Riccati sweep is still
    // going, filling the SMs the short wave leaves idle.  Same kernels on the same data: results are unchanged.
    int EA = 0;
    if (structured && !mdl->eager_line_search && mdl->side && split_enabled()) {
      const int slots = mdl->sm_count * lq3_resident(s3);
      const int rem = E % slots;
      if (E > slots && rem > 0 && 10 * rem <= 9 * slots) EA = E - rem;
    }
    if (structured) {
      Lq3Args l3{E, N, ws + L.ABc, ws + L.PK, ws + L.Kc, ws + L.Pc, ws + L.OKc, ws + L.dX, ws + L.dU, ws + L.dV,
                 ws + L.slope, done, 0};
      if (EA > 0) {
        l3.E = EA;
        rc = dispatch_lq3(s3, l3, st);
        if (rc == MPXB_OK && cudaEventRecord(mdl->ev_fork, st) != cudaSuccess) rc = MPXB_ECUDA;
        l3.e0 = EA; l3.E = E - EA;
      }
      if (EA == 0 || rc == MPXB_OK) rc = dispatch_lq3(s3, l3, st);
    } else {
      rc = dispatch_lq(lq_mode, n, m, la, ws + L.scan, st);
    }
    if (rc != MPXB_OK) return rc;
    if (pev) cudaEventRecord(pev[2], st);
    // K3: the 11 trial points.  Lazy (default): alpha = 1 first; the ten smaller steps only for the environments
    // whose alpha = 1 point the filter rejects -- the first accepted point is all the reference keeps
    // (optimizers.py:568-572), so X, U, V and alpha_index are unchanged.
    // Robots whose batches reject alpha = 1 often (H1: ~30 % of the environments) get alpha = 1/2 as a pass of its own,
    // so the last nine points are evaluated only where both were rejected.
    const bool three = !mdl->eager_line_search && d.kind != MPXB_KIND_SRBD && d.nv > 18 && E >= 256;
    const int npass = mdl->eager_line_search ? 1 : (three ? 3 : 2);
    for (int pass = 0; pass < npass; ++pass) {
      if (mdl->eager_line_search) { to.ia0 = 0; to.nia = NALPHA; to.need = nullptr; to.need_count = nullptr; }
      else if (three) { to.ia0 = pass; to.nia = pass < 2 ? 1 : NALPHA - 2; to.update_need = (pass == 1); }
      else { to.ia0 = pass; to.nia = pass == 0 ? 1 : NALPHA - 1; }
      if (pass == 0 && EA > 0) {
        // slice A behind its own Riccati sweep on the side stream, slice B behind the whole sweep on `st`, then join
        ModelIO ioA = io, ioB = io;
        ioA.E = EA; ioA.e0 = 0; ioB.E = E - EA; ioB.e0 = EA;
        if (cudaStreamWaitEvent(mdl->side, mdl->ev_fork, 0) != cudaSuccess) return MPXB_ECUDA;
        rc = wb_launch_trial(mdl->wb_host, mdl->wb_dev, ioA, to, mdl->side);
        if (rc != MPXB_OK) return rc;
        if (cudaEventRecord(mdl->ev_join, mdl->side) != cudaSuccess) return MPXB_ECUDA;
        rc = wb_launch_trial(mdl->wb_host, mdl->wb_dev, ioB, to, st);
        if (rc != MPXB_OK) return rc;
        if (cudaStreamWaitEvent(st, mdl->ev_join, 0) != cudaSuccess) return MPXB_ECUDA;
        continue;
      }
      if (d.kind == MPXB_KIND_SRBD) {
        const int wpb = 4;
        const int blocks = (E * to.nia + wpb - 1) / wpb;
        srbd_trial_kernel<<<blocks, wpb * 32, 0, st>>>(mdl->srbd, io, to);
      } else {
        rc = wb_launch_trial(mdl->wb_host, mdl->wb_dev, io, to, st);
        if (rc != MPXB_OK) return rc;
      }
    }
    if (pev) cudaEventRecord(pev[3], st);
    // K4
    SelectArgs sa{E, N, n, m, Xc, Uc, Vc, ws + L.dX, ws + L.dU, ws + L.dV, ws + L.theta_k, ws + L.phi_k, ws + L.ls_theta,
                  ws + L.ls_phi, ws + L.slope, Xn, Un, Vn, ws + L.g, ws + L.last_cost, done, alpha_index, iters,
                  status, max_iter > 1 ? 1 : 0, it == 0 ? 1 : 0, mdl->merit_line_search, ws + L.ls_vc, ws + L.vc_k,
                  ws + L.dv2, ws + L.dvc, mdl->stop_rule};
    select_update_kernel<<<E, 128, 0, st>>>(sa);
    if (pev) { cudaEventRecord(pev[4], st); pf->calls++; }
    Xc = Xn; Uc = Un; Vc = Vn;
  }
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

const float* mpxb_workspace_view(const mpxb_model* mdl, int E, int lq_mode, void* workspace, const char* name,
                                 size_t* count) {
  if (!mdl || !workspace || !name) return nullptr;
  const mpxb_model_desc& d = mdl->desc;
  const WsLayout L = layout_of(mdl, E, lq_mode);
  const size_t e = size_t(E), N = d.N, n = d.n, m = d.m;
  struct Ent { const char* nm; size_t off, cnt; } ents[] = {
      {"A", L.A, e * N * n * n}, {"B", L.B, e * N * n * m}, {"Q", L.Q, e * (N + 1) * n * n},
      {"R", L.R, e * N * m * m}, {"M", L.M, e * N * n * m}, {"q", L.q, e * (N + 1) * n}, {"r", L.r, e * N * m},
      {"c", L.c, e * (N + 1) * n}, {"K", L.K, e * N * m * n}, {"k", L.k, e * N * m}, {"P", L.P, e * (N + 1) * n * n},
      {"p", L.p, e * (N + 1) * n}, {"dX", L.dX, e * (N + 1) * n}, {"dU", L.dU, e * N * m},
      {"dV", L.dV, e * (N + 1) * n}, {"ls_theta", L.ls_theta, e * NALPHA}, {"ls_phi", L.ls_phi, e * NALPHA},
      {"g", L.g, e}, {"cost", L.cost, e * (N + 1)}, {"slope", L.slope, e}, {"ls_vc", L.ls_vc, e * NALPHA}};
  for (auto& en : ents)
    if (strcmp(en.nm, name) == 0) {
      if (count) *count = en.cnt;
      return static_cast<float*>(workspace) + en.off;
    }
  return nullptr;
}

int mpxb_workspace_expand(const mpxb_model* mdl, void* stream, int E, int lq_mode, void* workspace) {
  if (!mdl || !workspace || E <= 0) return MPXB_EINVAL;
  if (!use_structured(mdl, lq_mode)) return MPXB_OK;  // the dense path keeps the dense arrays up to date itself
  DeviceGuard guard(mdl->device);
  const WsLayout L = layout_of(mdl, E, lq_mode);
  float* ws = static_cast<float*>(workspace);
  ExpandArgs a{E, mdl->desc.N, mdl->s3, ws + L.ABc, ws + L.PK, ws + L.Kc, ws + L.Pc, ws + L.OKc,
               ws + L.A, ws + L.B, ws + L.Q, ws + L.R, ws + L.M, ws + L.q, ws + L.r, ws + L.K, ws + L.k, ws + L.P, ws + L.p};
  lq3_expand_kernel<<<E * (mdl->desc.N + 1), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

int mpxb_shift_extract(const mpxb_model* mdl, void* stream, int E, int shift, const float* X, const float* U,
                       const float* V, const float* x0, const float* u_ref, float* X0, float* U0, float* V0,
                       float* tau, float* q, float* dq) {
  if (!mdl || !X || !U || !V || !x0 || !u_ref || !X0 || !U0 || !V0 || E <= 0 || shift < 0) return MPXB_EINVAL;
  DeviceGuard guard(mdl->device);
  const mpxb_model_desc& d = mdl->desc;
  const int wb = d.kind != MPXB_KIND_SRBD;
  shift_extract_kernel<<<E, 128, 0, static_cast<cudaStream_t>(stream)>>>(E, d.N, d.n, d.m, d.n_joints, shift, wb, X, U,
                                                                         V, x0, u_ref, X0, U0, V0, tau, q, dq);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

int mpxb_whole_body_interface(const mpxb_model* mdl, void* stream, int E, const float* Kp, const float* Kd,
                              const float* qpos, const float* qvel, const float* grf, const float* foot_ref,
                              const float* foot_ref_dot, const float* contact, float* tau, float* J) {
  if (!mdl || !Kp || !Kd || !qpos || !qvel || !grf || !foot_ref || !foot_ref_dot || !contact || !tau || E <= 0)
    return MPXB_EINVAL;
  if (mdl->desc.kind != MPXB_KIND_QUAD_WB) return MPXB_EINVAL;
  DeviceGuard guard(mdl->device);
  WbiArgs a{E, Kp, Kd, qpos, qvel, grf, foot_ref, foot_ref_dot, contact, tau, J};
  return wb_quad_interface(mdl->wb_host, mdl->wb_dev, a, static_cast<cudaStream_t>(stream));
}

int mpxb_wb_foot_positions(const mpxb_model* mdl, void* stream, int E, const float* qpos, float* feet) {
  if (!mdl || !qpos || !feet || E <= 0) return MPXB_EINVAL;
  if (mdl->desc.kind == MPXB_KIND_SRBD) return MPXB_EINVAL;
  DeviceGuard guard(mdl->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const WbModel& h = mdl->wb_host;
  const int nt = 64, nb = (E + nt - 1) / nt;
  if (h.nbody <= 14 && h.nv <= 18 && h.nc <= 4) wb_feet_kernel<14, 18, 4><<<nb, nt, 0, st>>>(mdl->wb_dev, E, qpos, feet);
  else if (h.nbody <= 21 && h.nv <= 25 && h.nc <= 4) wb_feet_kernel<21, 25, 4><<<nb, nt, 0, st>>>(mdl->wb_dev, E, qpos, feet);
  else wb_feet_kernel<32, 28, 8><<<nb, nt, 0, st>>>(mdl->wb_dev, E, qpos, feet);
  return cudaGetLastError() == cudaSuccess ? MPXB_OK : MPXB_ECUDA;
}

int mpxb_reference_generator(const mpxb_model* mdl, void* stream, int E, const mpxb_gait* g, const float* x,
                             const float* foot, int foot_stride, const float* input, const float* contact,
                             const float* timer_in, float* timer_out, float* liftoff, float* reference,
                             float* parameter, float* foot_ref_dot) {
  if (!mdl || !g || !x || !foot || !input || !timer_in || !liftoff || !reference || !parameter || E <= 0) return MPXB_EINVAL;
  const mpxb_model_desc& d = mdl->desc;
  if (g->use_terrain_estimation && d.n_contact != 4) return MPXB_EINVAL;
  if (foot_stride < 3 * d.n_contact || d.n_contact > MPXB_MAX_CONTACT) return MPXB_EINVAL;
  DeviceGuard guard(mdl->device);
  RefGenArgs a;
  a.E = E; a.N = d.N; a.nc = d.n_contact; a.srbd = d.kind == MPXB_KIND_SRBD; a.nj = a.srbd ? 0 : d.n_joints; a.n = d.n;
  a.use_terrain = g->use_terrain_estimation; a.advance_first = g->advance_timer_first;
  a.dt = d.dt; a.mpc_dt = g->mpc_dt; a.duty_factor = g->duty_factor; a.step_freq = g->step_freq;
  a.step_height = g->step_height; a.mass = g->mass; a.clearence_speed = g->clearence_speed;
  memcpy(a.foot0, g->foot0, sizeof(a.foot0));
  memcpy(a.q0, g->q0, sizeof(a.q0));
  a.x = x; a.foot = foot; a.foot_stride = foot_stride; a.input = input; a.contact = contact; a.timer_in = timer_in;
  a.timer_out = timer_out; a.liftoff = liftoff; a.reference = reference; a.parameter = parameter; a.foot_ref_dot = foot_ref_dot;
  return launch_refgen(a, static_cast<cudaStream_t>(stream));
}

int mpxb_profile_create(int max_calls, mpxb_profile** out) {
  if (!out || max_calls <= 0) return MPXB_EINVAL;
  mpxb_profile* p = new (std::nothrow) mpxb_profile();
  if (!p) return MPXB_EINVAL;
  p->max_calls = max_calls;
  p->calls = 0;
  p->ev = new (std::nothrow) cudaEvent_t[size_t(5) * max_calls];
  if (!p->ev) { delete p; return MPXB_EINVAL; }
  for (int i = 0; i < 5 * max_calls; ++i)
    if (cudaEventCreate(&p->ev[i]) != cudaSuccess) { delete[] p->ev; delete p; return MPXB_ECUDA; }
  *out = p;
  return MPXB_OK;
}

void mpxb_profile_destroy(mpxb_profile* p) {
  if (!p) return;
  for (int i = 0; i < 5 * p->max_calls; ++i) cudaEventDestroy(p->ev[i]);
  delete[] p->ev;
  delete p;
}

int mpxb_profile_attach(mpxb_model* mdl, mpxb_profile* p) {
  if (!mdl) return MPXB_EINVAL;
  mdl->prof = p;
  if (p) p->calls = 0;
  return MPXB_OK;
}

int mpxb_profile_read(mpxb_profile* p, double mean_ms[4], int* calls) {
  if (!p || !mean_ms) return MPXB_EINVAL;
  for (int k = 0; k < 4; ++k) mean_ms[k] = 0.0;
  for (int c = 0; c < p->calls; ++c) {
    cudaEvent_t* ev = p->ev + 5 * c;
    if (cudaEventSynchronize(ev[4]) != cudaSuccess) return MPXB_ECUDA;
    for (int k = 0; k < 4; ++k) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) != cudaSuccess) return MPXB_ECUDA;
      mean_ms[k] += ms;
    }
  }
  if (p->calls > 0)
    for (int k = 0; k < 4; ++k) mean_ms[k] /= p->calls;
  if (calls) *calls = p->calls;
  return MPXB_OK;
}

// Standalone LQ solve (tvlqr + rollout + dual_lqr, or the scan forms): the reference functions a13-a16 of SURVEY 8a.
int mpxb_tvlqr(void* stream, int E, int N, int n, int m, int lq_mode, const float* Q, const float* q, const float* R,
               const float* r, const float* M, const float* A, const float* B, const float* c /*[E,N+1,n]: c[0]=x0 offset*/,
               float* K, float* k, float* P, float* p, float* dX, float* dU, float* dV, float* slope, void* scan_ws) {
  if (!Q || !q || !R || !r || !M || !A || !B || !c || !K || !k || !P || !p || !dX || !dU || !dV || !slope) return MPXB_EINVAL;
  LqArgs la{E, N, A, B, Q, R, M, q, r, c, K, k, P, p, dX, dU, dV, slope, nullptr, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (E <= 0 || N <= 0 || !lq_mode_valid(lq_mode, E)) return MPXB_EINVAL;
  if (lq_mode != MPXB_LQ_SEQUENTIAL && !scan_ws) return MPXB_EINVAL;
  int rc = configure_lq_dispatch(n, m);
  if (rc != MPXB_OK) return rc;
  return dispatch_lq(lq_mode, n, m, la, static_cast<float*>(scan_ws), st);
}

size_t mpxb_tvlqr_scan_bytes(int E, int N, int n, int m) {
  (void)m;
  return lq_scan_ws_floats(MPXB_LQ_SCAN, E, N, n) * sizeof(float);
}

__global__ void fp32_peak_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
        a7 = a0 + 7.f;
  const float b = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
      a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

double mpxb_fp32_peak_tflops(void* stream, int iters) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, threads = 256;
  float* out = nullptr;
  if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return 0.0;  // "no measurement", never an error code as a rate
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp32_peak_kernel<<<blocks, threads, 0, st>>>(out, 64);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    fp32_peak_kernel<<<blocks, threads, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * double(iters) * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // extern "C"
