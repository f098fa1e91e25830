/*
 * mpxb.h -- C ABI of libmpxb.so: the B200-native (sm_100a) primal-dual iLQR solve of iit-DLSLab/mpx.
 *
 * Drop-in boundary.  In the reference the path is the Python callable
 *     work = partial(optimizers.mpc, cost, dynamics, hessian_approx, limited_memory)
 *     X, U, V = jax.jit(work)(reference, parameter, W, x0, X_in, U_in, V_in)
 * (reference mpx/primal_dual_ilqr/primal_dual_ilqr/optimizers.py:603-683; built at
 * mpx/utils/mpc_wrapper.py:74,244,249 and mpx/utils/mpc_wrapper_srbd.py:68,77; called at
 * mpc_wrapper.py:150-158,346-354,434-442 and mpc_wrapper_srbd.py:124-132).
 * A JAX maintainer binds mpxb_mpc_step as an XLA-FFI custom call; this repo's own host
 * code binds it with ctypes (mpx_b200/_lib.py).
 *
 * Conventions
 *   - every array is float32, row-major C-contiguous, leading axis E = environments ([E, knot, feature]);
 *   - all pointers are DEVICE pointers owned by the caller (inputs read-only, outputs fully overwritten);
 *   - calls are asynchronous on `stream`, allocate nothing, never synchronise and can be captured in a CUDA graph;
 *   - return value: 0 on success, negative MPXB_E* for ARGUMENT errors only.  Numerical failure is reported
 *     in-band exactly like the reference (NaNs in the outputs, mpc_wrapper.py:96,280) plus status[e] != 0.
 *   - no global mutable state; a model handle is immutable after creation (options are set before the first
 *     step) and may be used on several streams of its device as long as each call has its own workspace and the
 *     calls on one handle are ISSUED from one host thread at a time: for batches above one wave of LQ CTAs
 *     mpxb_mpc_step forks part of its line search onto a side stream owned by the handle (event fork / join on
 *     `stream`, valid under stream capture), and those two events are per handle.  A handle belongs to the device that was current at mpxb_model_create; every entry point
 *     switches to that device for its launches and restores the caller's current device.
 *   - `workspace` must be 256-byte aligned.
 */
#ifndef MPXB_H_
#define MPXB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPXB_VERSION 1

/* dynamics/cost family == reference function triple (models.py / objectives.py) */
enum {
  MPXB_KIND_SRBD = 0,     /* quadruped_srbd_dynamics / _obj / _hessian_gn   (models.py:6-37,   objectives.py:15-75)   */
  MPXB_KIND_QUAD_WB = 1,  /* quadruped_wb_dynamics / _obj / _hessian_gn     (models.py:39-117, objectives.py:77-202)  */
  MPXB_KIND_H1_WB = 2,    /* h1_wb_dynamics / _obj / _hessian_gn            (models.py:119-169, objectives.py:204-323) */
  MPXB_KIND_TALOS_WB = 3  /* talos_wb_dynamics / _obj / _hessian_gn         (models.py:170-217, objectives.py:325-450) */
};

/* LQ sub-problem solver == the reference's `limited_memory` static argument (optimizers.py:175-180) */
enum {
  MPXB_LQ_SEQUENTIAL = 0, /* tvlqr + rollout          (primal_tvlqr.py:103-144, 478-498)  limited_memory=True  */
  MPXB_LQ_SCAN = 1,       /* tvlqr_gpu + rollout_gpu  (primal_tvlqr.py:337-475, 523-549)  limited_memory=False:
                             the associative scan, scheduled in parallel over the horizon (one CTA per knot,
                             ceil(log2(N+1)) levels) for small batches and serially per environment for large ones */
  MPXB_LQ_SCAN_SERIAL = 2,   /* force the one-CTA-per-environment schedule of the same scan */
  MPXB_LQ_SCAN_PARALLEL = 3  /* force the parallel-in-time schedule (E <= 64) */
};

enum {
  MPXB_OK = 0,
  MPXB_EINVAL = -1,      /* null pointer / bad dimension / unknown kind */
  MPXB_EUNSUPPORTED = -2,/* (n, m) or tree size outside the compiled kernel set */
  MPXB_EWORKSPACE = -3,  /* workspace_bytes too small */
  MPXB_ECUDA = -4        /* a CUDA runtime call failed at model creation (never returned by mpxb_mpc_step) */
};

#define MPXB_MAX_BODY 32
#define MPXB_MAX_DOF 32
#define MPXB_MAX_CONTACT 8

/* Flat model description (host memory).  The Python side fills it from the config module and the MJCF
 * model table (the reference does the same through mujoco.MjModel.from_xml_path + mjx.put_model,
 * mpc_wrapper.py:43-47).  Whole-body fields are ignored for MPXB_KIND_SRBD and vice versa. */
typedef struct mpxb_model_desc {
  int32_t version;        /* MPXB_VERSION */
  int32_t kind;           /* MPXB_KIND_* */
  int32_t n, m, N;        /* state dim, control dim, horizon (config.n, config.m, config.N) */
  int32_t n_joints;       /* config.n_joints */
  int32_t n_contact;      /* config.n_contact */
  int32_t nref, npar, nw; /* columns of reference / parameter, side of W */
  int32_t swing_tracking; /* the flag bound by partial(quadruped_wb_obj, True) (config_go2.py:67) */
  float dt;               /* config.dt */
  /* ---- SRBD (config_srbd.py:60-65) ---- */
  float mass;
  float inertia[9];
  float inertia_inv[9];
  /* ---- whole body: MuJoCo-named tree arrays, body 0 = world (mpx_b200/mjcf.py) ---- */
  int32_t nbody, nv;
  float gravity[3];
  float baumgarte_alpha;  /* models.py:93 (25) / :146 (5) */
  float contact_reg;      /* models.py:155 (1e-3) */
  int32_t body_parentid[MPXB_MAX_BODY];
  int32_t body_dofadr[MPXB_MAX_BODY];   /* first dof of the body's joint, -1 if none */
  int32_t body_jnttype[MPXB_MAX_BODY];  /* 0 free, 3 hinge, -1 none (one joint per body) */
  float body_pos[MPXB_MAX_BODY][3];
  float body_quat[MPXB_MAX_BODY][4];
  float body_ipos[MPXB_MAX_BODY][3];
  float body_iquat[MPXB_MAX_BODY][4];
  float body_mass[MPXB_MAX_BODY];
  float body_inertia[MPXB_MAX_BODY][3];
  float jnt_pos[MPXB_MAX_BODY][3];      /* indexed by body */
  float jnt_axis[MPXB_MAX_BODY][3];
  float jnt_qpos0[MPXB_MAX_BODY];       /* hinge reference angle */
  float dof_armature[MPXB_MAX_DOF];
  int32_t contact_body[MPXB_MAX_CONTACT]; /* body carrying foot point i, in the model's contact order */
  float contact_pos[MPXB_MAX_CONTACT][3]; /* foot point in that body's frame (geom_pos) */
} mpxb_model_desc;

typedef struct mpxb_model mpxb_model; /* opaque; owns its constant device tables */

/* Create / destroy a model on the CURRENT CUDA device.  Replaces the reference's per-wrapper setup
 * (mpc_wrapper.py:43-74, mpc_wrapper_srbd.py:33-68). */
int mpxb_model_create(const mpxb_model_desc* desc, mpxb_model** out);
void mpxb_model_destroy(mpxb_model* model);

/* Options of a model handle (set before the first mpxb_mpc_step; not thread-safe against a running call).
 * MPXB_OPT_EAGER_LINE_SEARCH (default 0): the reference evaluates all 11 trial points of the filter line search
 * in one vmap and keeps the first accepted one (optimizers.py:518-572).  By default libmpxb evaluates alpha = 1
 * first and the ten smaller steps only for environments whose alpha = 1 point is rejected -- X, U, V and
 * alpha_index are identical, but the workspace views "ls_theta" / "ls_phi" then hold entries 1..10 only for
 * those environments.  Set to 1 to evaluate all 11 points for every environment (diagnostics / parity tests).
 * MPXB_OPT_LINE_SEARCH (default MPXB_LS_FILTER): MPXB_LS_FILTER is parallel_filter_line_search, the search the
 * reference's mpc runs (optimizers.py:468-578, 670-681); MPXB_LS_MERIT is parallel_line_search on the augmented
 * Lagrangian merit g + sum((V + rho/2 c) c) with merit_rho / slope (optimizers.py:185-217, 295-354), the call the
 * reference keeps commented out at optimizers.py:637-669.  MPXB_LS_ARMIJO is the sequential backtracking `line_search`
 * of the legacy solver (old_optimazers.py:173-251; same function at optimizers.py:219-293) with alpha_0 = 1,
 * alpha_mult = 1/2 and alpha_min = 2^-10, i.e. on the same 11 trial points: the first point that passes the Armijo test
 * on the merit function with the legacy slope (old_optimazers.py:154-170: sum(dV c) once), else the last point tried
 * (alpha = 2^-10) with no_errors = false (status bit 1).
 * MPXB_OPT_STOP_RULE (default MPXB_STOP_OFFLINE), for max_iter > 1: MPXB_STOP_OFFLINE is runOffline's rule
 * (mpc_wrapper.py:456-467: last_cost - g^2 < 1e-3 and sum c^2 < 1e-5 at the new iterate); MPXB_STOP_LEGACY is the
 * legacy loop's continuation_criterion (old_optimazers.py:435-447), evaluated on the search direction BEFORE the step
 * is taken: continue while no_errors and (|merit slope| > 1e-4 and |d|^2 > 0  or  sum c^2 > 1e-4); iters[e] counts
 * the steps taken.  It needs the merit quantities, so it requires MPXB_LS_MERIT or MPXB_LS_ARMIJO (else MPXB_EINVAL
 * from mpxb_mpc_step).
 * MPXB_OPT_COST_STRUCTURE (default MPXB_COST_DENSE): MPXB_COST_DIAGONAL_W is the caller's promise that the weight
 * matrix W passed to mpxb_mpc_step is DIAGONAL (every shipped config: block_diag of diag(...), config_*.py).  The
 * Gauss-Newton blocks of the whole-body objectives (objectives.py:127-202, 252-323, 378-450) are then Q_t = diag + a
 * 4x4 quaternion block (+ 3x3 friction-cone blocks), R_t = diag, M_t = 0, and with limited_memory = True
 * (MPXB_LQ_SEQUENTIAL) the solve runs the structured kernels (csrc/lq3.cuh): same recursion, structural zeros never
 * multiplied, LQ data in packed blocks moved by bulk asynchronous copies.  X, U, V agree with the dense path to
 * rounding.  The dense views "A".."p" are then filled by mpxb_workspace_expand.  Ignored for MPXB_KIND_SRBD and the
 * scan modes.  A non-diagonal W with this option set gives wrong results (the off-diagonal entries are never read). */
enum { MPXB_OPT_EAGER_LINE_SEARCH = 1, MPXB_OPT_LINE_SEARCH = 2, MPXB_OPT_COST_STRUCTURE = 3, MPXB_OPT_STOP_RULE = 4 };
enum { MPXB_LS_FILTER = 0, MPXB_LS_MERIT = 1, MPXB_LS_ARMIJO = 2 };
enum { MPXB_STOP_OFFLINE = 0, MPXB_STOP_LEGACY = 1 };
enum { MPXB_COST_DENSE = 0, MPXB_COST_DIAGONAL_W = 1 };
int mpxb_model_set_option(mpxb_model* model, int option, int value);

/* Bytes of device scratch one mpxb_mpc_step call with E environments needs. */
size_t mpxb_workspace_bytes(const mpxb_model* model, int E, int lq_mode);

/* One call == `max_iter` SQP iterations of optimizers.mpc for each of E environments (max_iter = 1 is the
 * reference's real-time iteration).  With max_iter > 1 the loop runs stream-ordered on the device and an
 * environment stops early once the offline stopping rule of mpc_wrapper.py:463-467 holds
 * (last_cost - g*g < 1e-3 and sum(c*c) < 1e-5); iters[e] reports the iterations actually taken.
 *
 *   reference [E,N+1,nref]  parameter [E,N+1,npar]  W [nw,nw] if w_env_stride == 0 else [E,nw,nw]
 *   x0 [E,n]  X_in/X [E,N+1,n]  U_in/U [E,N,m]  V_in/V [E,N+1,n]
 *   alpha_index [E]  index i of the accepted step 2^-i (10 if none, optimizers.py:572)      (may be NULL)
 *   iters [E], status [E]  (status bit 0: non-finite value in X/U/V; bit 1: MPXB_LS_ARMIJO ran out of step sizes)                        (may be NULL)
 * X/U/V must not alias X_in/U_in/V_in. */
int mpxb_mpc_step(const mpxb_model* model, void* stream /* cudaStream_t */, int E, int lq_mode, int max_iter,
                  const float* reference, const float* parameter, const float* W, size_t w_env_stride,
                  const float* x0, const float* X_in, const float* U_in, const float* V_in,
                  float* X, float* U, float* V,
                  int32_t* alpha_index, int32_t* iters, int32_t* status,
                  void* workspace, size_t workspace_bytes);

/* Read-only views into the workspace of the LAST mpxb_mpc_step call (gains, value function, search
 * direction, linearisation) for parity tests: name in {"A","B","Q","R","M","q","r","c","K","k","P","p",
 * "dX","dU","dV","ls_theta","ls_phi","g"}.  Returns a device pointer inside `workspace` or NULL. */
const float* mpxb_workspace_view(const mpxb_model* model, int E, int lq_mode, void* workspace, const char* name,
                                 size_t* count);
/* With MPXB_COST_DIAGONAL_W the last call kept A, B, Q, R, M, q, r, K, k, P, p in packed form; this launches (stream-
 * ordered) the kernel that writes them out densely where mpxb_workspace_view points.  A no-op for the dense path. */
int mpxb_workspace_expand(const mpxb_model* model, void* stream, int E, int lq_mode, void* workspace);

/* Warm-start shift / NaN guard / command extraction of update_and_extract_helper
 * (mpc_wrapper.py:79-98, 261-282; mpc_wrapper_srbd.py:135-139), batched over E, stream-ordered.
 *   X,U,V: solution of this tick; X0,U0,V0: previous warm start (in) -> next warm start (out);
 *   tau [E,n_joints], q [E,n_joints], dq [E,n_joints] (whole-body kinds; NULL for SRBD). */
int mpxb_shift_extract(const mpxb_model* model, void* stream, int E, int shift,
                       const float* X, const float* U, const float* V, const float* x0, const float* u_ref,
                       float* X0, float* U0, float* V0, float* tau, float* q, float* dq);

/* Standalone LQ sub-problem solve, batched over E: the reference functions tvlqr + rollout + dual_lqr
 * (lq_mode = MPXB_LQ_SEQUENTIAL; primal_tvlqr.py:103-144, 478-498; dual_tvlqr.py:6-19) or tvlqr_gpu + rollout_gpu +
 * dual_lqr (MPXB_LQ_SCAN; primal_tvlqr.py:337-475, 523-549) as compute_search_direction calls them
 * (optimizers.py:175-182).  c is [E,N+1,n]: c[:,0] is the rollout's initial offset (optimizers.py:177 passes c[0]),
 * c[:,1:] the dynamics offsets.  Outputs K [E,N,m,n], k [E,N,m], P [E,N+1,n,n], p [E,N+1,n], dX, dU, dV and
 * slope [E] = sum(q*dX) + sum(r*dU).  scan_ws: mpxb_tvlqr_scan_bytes() bytes of device scratch (scan mode only). */
int mpxb_tvlqr(void* stream, int E, int N, int n, int m, int lq_mode, const float* Q, const float* q, const float* R,
               const float* r, const float* M, const float* A, const float* B, const float* c,
               float* K, float* k, float* P, float* p, float* dX, float* dU, float* dV, float* slope, void* scan_ws);
size_t mpxb_tvlqr_scan_bytes(int E, int N, int n, int m);

/* Foot (contact sphere) positions by forward kinematics, batched: qpos [E, 7 + n_joints] -> feet [E, 3 * n_contact].
 * Replaces the host-side MuJoCo mj_kinematics + geom_xpos read of mpc_wrapper.py:313-318 / 373-376 (whole-body kinds). */
int mpxb_wb_foot_positions(const mpxb_model* model, void* stream, int E, const float* qpos, float* feet);

/* whole_body_interface (mpx/utils/mpc_utils.py:275-313; bound by mpc_wrapper_srbd.py:72,153-154): turns the SRBD
 * controller's ground-reaction forces into joint torques at the whole-body rate.  `model` is a MPXB_KIND_QUAD_WB handle
 * of the same robot (three joints per leg, contact order = leg order).  qM, qfrc_bias, the foot positions and the foot
 * Jacobians J [nv, 3 nc] are evaluated at (qpos, qvel);  csa = Kp (foot_ref - feet) + Kd (foot_ref_dot - J' qvel);
 * tau = contact ? -(J grf)[6:] : qfrc_bias[6:] + (qM pinv(J') csa)[6:]  per leg, with pinv(J') = J (J'J)^-1.
 * Kp, Kd [3nc, 3nc] shared; qpos [E, nq]; qvel [E, nv]; grf, foot_ref, foot_ref_dot [E, 3nc]; contact [E, nc];
 * tau [E, nj]; J [E, nv, 3nc] or NULL. */
int mpxb_whole_body_interface(const mpxb_model* model, void* stream, int E, const float* Kp, const float* Kd,
                              const float* qpos, const float* qvel, const float* grf, const float* foot_ref,
                              const float* foot_ref_dot, const float* contact, float* tau, float* J);

/* Gait timer + swing / foothold reference generator, batched over E and stream-ordered: the step immediately before
 * the solve in every wrapper tick.  Replaces timer_run (mpx/utils/mpc_utils.py:7-14), reference_generator (38-144,
 * whole-body kinds) and reference_generator_srbd (146-270, MPXB_KIND_SRBD) as the wrappers call them
 * (mpc_wrapper.py:132-147, 329-343; mpc_wrapper_srbd.py:103-121).
 *   x [E,n] measured state; foot: rows of `foot_stride` floats whose first 3*n_contact entries are the measured foot
 *   positions (may point into x); input [E,7] command; contact [E,n_contact] measured contact flags or NULL (= 0, as the
 *   batched whole-body wrapper passes, mpc_wrapper.py:134);
 *   timer_in [E,n_contact] gait phase; timer_out (may be NULL or alias timer_in) receives the phase advanced by
 *   gait->mpc_dt * step_freq (timer_run); the generator itself starts from timer_in (advance_timer_first = 0, batched
 *   whole-body wrapper) or from the advanced phase (advance_timer_first = 1, single-instance and SRBD wrappers);
 *   liftoff [E,3*n_contact] updated in place;
 *   reference [E,N+1,nref], parameter [E,N+1,npar] outputs; foot_ref_dot [E,N+1,3*n_contact] output or NULL. */
typedef struct mpxb_gait {
  int32_t use_terrain_estimation; /* config.use_terrain_estimation (needs n_contact == 4) */
  int32_t advance_timer_first;
  float mpc_dt;                   /* 1 / config.mpc_frequency */
  float duty_factor, step_freq, step_height;
  float mass;                     /* robot mass used for the GRF reference */
  float clearence_speed;
  float foot0[3 * MPXB_MAX_CONTACT]; /* config.p_legs0 */
  float q0[MPXB_MAX_DOF];            /* config.q0 (whole-body kinds) */
} mpxb_gait;
int mpxb_reference_generator(const mpxb_model* model, void* stream, int E, const mpxb_gait* gait, const float* x,
                             const float* foot, int foot_stride, const float* input, const float* contact,
                             const float* timer_in, float* timer_out, float* liftoff, float* reference,
                             float* parameter, float* foot_ref_dot);

/* Per-kernel timing for benchmarks (bench.py roofline): once attached, every mpxb_mpc_step on `model` records CUDA
 * events on its stream at the boundaries of the four kernels of the LAST SQP iteration of the call
 * (linearise | LQ solve | line-search trials | select/update).  Up to `max_calls` calls are kept (later calls are
 * not recorded).  mpxb_profile_read synchronises on the recorded events and returns the MEAN duration in
 * milliseconds of each of the four kernels over the recorded calls.  Attaching mutates the handle: benchmarking
 * only, not thread-safe.  Pass profile = NULL to detach. */
typedef struct mpxb_profile mpxb_profile;
int mpxb_profile_create(int max_calls, mpxb_profile** out);
void mpxb_profile_destroy(mpxb_profile* profile);
int mpxb_profile_attach(mpxb_model* model, mpxb_profile* profile);
int mpxb_profile_read(mpxb_profile* profile, double mean_ms[4], int* calls);

/* FP32 FFMA throughput micro-benchmark (roofline denominator).  Returns TFLOP/s measured with
 * CUDA events on `stream`, or a negative MPXB_E*. */
double mpxb_fp32_peak_tflops(void* stream, int iters);

const char* mpxb_version_string(void);

#ifdef __cplusplus
}
#endif
#endif /* MPXB_H_ */
