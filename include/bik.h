/* bik.h -- C ABI of libbik: batched differential inverse kinematics on B200 (sm_100a).
 *
 * mink has no FFI layer of its own: its hot path is reached through the Python
 * classes Configuration / Task / Limit and solve_ik().  This header is what a replacement of that
 * path binds to: every entry point below names the reference function(s) it replaces.  Plain
 * pointers and sizes only; no torch / numpy / C++ types.
 *
 * Conventions
 *   - every function returns BIK_OK (0) or a negative bik_status; text via bik_last_error().
 *   - `const float* q` etc. are DEVICE pointers unless a parameter is documented "host".
 *   - descriptors (bik_task_desc, bik_limit_desc, bik_frame, model blob) are HOST memory and are
 *     copied at create time; the caller keeps ownership of everything it passes in.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls enqueue
 *     work and return; nothing synchronises unless stated.
 *   - batch layout is row-major with the instance index outermost: q[B][nq], J[B][K][nv], ...
 *   - quaternions are (w,x,y,z); poses are (qw,qx,qy,qz,x,y,z) like mink.SE3.wxyz_xyz
 *     (reference mink/lie/se3.py:24-29); twists are (v, omega) (ibid.).
 *   - model handles are immutable after creation and may be shared between streams.  A problem handle owns the K1 -> K2
 *     hand-off buffers of bik_step / bik_converge: calls on one problem handle are ordered (a call on another stream than
 *     the previous one first waits, on the device, for that call's kernels); use one problem handle per stream for overlap.
 */
#ifndef BIK_H
#define BIK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIK_VERSION 200

typedef enum bik_status {
  BIK_OK = 0,
  BIK_ERR_INVALID = -1,     /* bad argument / malformed blob or descriptor */
  BIK_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  BIK_ERR_UNSUPPORTED = -3, /* model or problem outside the supported subset */
  BIK_ERR_NOMEM = -4
} bik_status;

/* Per-instance status bits written by the solve/step/check entry points.  Inside a batch these
 * replace the reference's exceptions (NotWithinConfigurationLimits, configuration.py:97-105;
 * `assert dq is not None`, solve_ik.py:103). */
#define BIK_STATUS_OUT_OF_LIMITS 1u  /* q outside [range - tol, range + tol] */
#define BIK_STATUS_QP_MAXITER    2u  /* active-set iteration cap reached, or more general rows in play than the solver holds
                                         (24 active at once, 64 inside the detection distance); dq is the last iterate */
#define BIK_STATUS_NONFINITE     4u  /* NaN/Inf met in q, targets or the factorisation */
#define BIK_STATUS_QP_INFEASIBLE 8u  /* inequality set inconsistent (lower > upper for a dof, or no feasible start): the
                                      * reference's solver returns None there and solve_ik asserts (solve_ik.py:103) */
#define BIK_STATUS_NOT_CONVERGED 16u /* bik_converge: thresholds not met within max_iters */

typedef struct bik_model bik_model;     /* flattened kinematic tree resident on one device */
typedef struct bik_problem bik_problem; /* static task + limit layout bound to a model */

/* A frame rigidly attached to a tree node (node = -1: fixed in the world).  The front end
 * resolves (name, "body"|"geom"|"site") to this once, instead of the per-call mj_name2id of
 * reference configuration.py:133,170. */
typedef struct bik_frame {
  int32_t node;
  int32_t reserved;
  double pos[3];
  double quat[4];
} bik_frame;

enum { BIK_TASK_FRAME = 0, BIK_TASK_POSTURE = 1, BIK_TASK_COM = 2, BIK_TASK_RELATIVE_FRAME = 3 };

/* One kinematic task (reference mink/tasks/task.py:54-79 holds cost/gain/lm_damping).
 *   FRAME   : FrameTask   (frame_task.py:28-45)   rows = 6, cost = position xyz | orientation xyz
 *   POSTURE : PostureTask (posture_task.py:27-50) rows = nv (never materialised as a Jacobian:
 *             J = -I with free-joint dofs zeroed, posture_task.py:137-141); dof_cost[nv] host array.
 *             DampingTask (damping_task.py:19-20) = POSTURE with gain 0 and target qpos0.
 *   COM     : ComTask     (com_task.py:25-33)     rows = 3, cost[0..2]
 *   RELATIVE_FRAME : RelativeFrameTask (relative_frame_task.py:24-45) rows = 6; pose of `frame` expressed
 *             in `root`; its target (one slot of frame_targets, in task order) is T_root<-target. */
typedef struct bik_task_desc {
  int32_t kind;
  int32_t reserved;
  bik_frame frame;
  double cost[6];
  const double* dof_cost;
  double gain;
  double lm_damping;
  bik_frame root; /* RELATIVE_FRAME only */
} bik_task_desc;

enum { BIK_LIMIT_CONFIGURATION = 0, BIK_LIMIT_VELOCITY = 1, BIK_LIMIT_COLLISION = 2 };
enum { BIK_GEOM_PLANE = 0, BIK_GEOM_SPHERE = 2, BIK_GEOM_CAPSULE = 3, BIK_GEOM_BOX = 6 };   /* mjtGeom codes */

typedef struct bik_geom {
  int32_t type;
  int32_t reserved;
  bik_frame frame;
  double size[3];
} bik_geom;

/* One inequality limit  G(q) dq <= h(q)  (reference mink/limits/limit.py:11-57).
 *   CONFIGURATION (configuration_limit.py:69-124): for each listed dof i (slide/hinge):
 *        -gain*(q_i - lower_i) <= dq_i <= gain*(upper_i - q_i);  lower/upper already include
 *        min_distance_from_limits (configuration_limit.py:50-51).
 *   VELOCITY      (velocity_limit.py:71-101):      |dq_i| <= dt * vmax_i.
 *   COLLISION     (collision_avoidance_limit.py:187-210): one row per geom pair,
 *        -n^T (Jp2 - Jp1) dq <= gain*(dist - minimum_distance)/dt + bound_relaxation, rows whose
 *        distance is >= detection_distance are inactive.  Primitive pairs only (plane, sphere,
 *        capsule, box; no plane-plane and no box-box pair). */
typedef struct bik_limit_desc {
  int32_t kind;
  int32_t n;             /* CONFIGURATION/VELOCITY: number of listed dofs; COLLISION: number of pairs */
  const int32_t* dof;    /* [n] tangent-space indices */
  const double* lower;   /* [n] CONFIGURATION */
  const double* upper;   /* [n] CONFIGURATION */
  const double* vmax;    /* [n] VELOCITY */
  double gain;
  const bik_geom* geoms; /* COLLISION */
  int32_t ngeoms;
  int32_t reserved;
  const int32_t* pairs;  /* [n][2] indices into geoms */
  double minimum_distance;
  double detection_distance;
  double bound_relaxation;
} bik_limit_desc;

/* Per-call inputs of the task layer (device pointers).  Targets are what FrameTask.set_target /
 * PostureTask.set_target / ComTask.set_target store (frame_task.py:83, posture_task.py:77,
 * com_task.py:61), one per instance. */
typedef struct bik_inputs {
  const void* q;               /* [B][nq] */
  const void* frame_targets;   /* [B][F][7]  F = number of FRAME + RELATIVE_FRAME tasks, in task-list order */
  const void* posture_targets; /* [B or 1][P][nq]  P = number of POSTURE tasks */
  const void* com_targets;     /* [B][C][3]  C = number of COM tasks */
  int32_t posture_batched;     /* 0: one posture target shared by the batch, 1: per instance */
  int32_t f64;                 /* element type of the four buffers: 0 float (fp32 entry points), 1 double (the ...64 ones) */
} bik_inputs;

typedef struct bik_dims {
  int32_t nq, nv, nnode;
  int32_t nframe, nposture, ncom; /* F, P, C */
  int32_t nrows;                  /* K = 6F + 3C : rows of the stacked J / e */
  int32_t npairs;                 /* collision rows */
} bik_dims;

int bik_version(void);
const char* bik_last_error(void);

/* Model = what mink reads out of MjModel during mj_kinematics/mj_jac (configuration.py:63-64,145).
 * `blob` is the BIKM v1 buffer produced by mink_b200.flatten.FlatModel.to_blob(). */
int bik_model_create(const void* blob, size_t nbytes, int device, bik_model** out);
void bik_model_destroy(bik_model* model);

int bik_problem_create(const bik_model* model, const bik_task_desc* tasks, int ntasks,
                       const bik_limit_desc* limits, int nlimits, bik_problem** out);
void bik_problem_destroy(bik_problem* problem);
int bik_problem_dims(const bik_problem* problem, bik_dims* out);

/* Forward kinematics of arbitrary frames.
 * Replaces Configuration.update + get_transform_frame_to_world (configuration.py:53-64,157-185)
 * and data.subtree_com[1] (com_task.py:69).  poses [B][nframes][7]; com [B][3] or NULL.
 * `frames` is a HOST array. */
int bik_fk(const bik_model* model, int B, const float* q, const bik_frame* frames, int nframes,
           float* poses, float* com, void* stream);

/* Body-frame Jacobians of arbitrary frames: Configuration.get_frame_jacobian
 * (configuration.py:112-155).  J [B][nframes][6][nv], rows = (linear xyz, angular xyz). */
int bik_frame_jacobian(const bik_model* model, int B, const float* q, const bik_frame* frames,
                       int nframes, float* J, void* stream);

/* K1: FK + task errors + task Jacobians for every task of the problem.
 * Replaces Task.compute_error / Task.compute_jacobian of FrameTask (frame_task.py:95-146),
 * ComTask (com_task.py:71-97) and PostureTask.compute_error (posture_task.py:87-118).
 *   J         [B][K][nv]   stacked rows in task order (FRAME 6 rows, COM 3 rows)
 *   e         [B][K]
 *   e_posture [B][P][nv]   (NULL allowed when P = 0)
 *   G_coll    [B][npairs][nv], h_coll [B][npairs]  collision rows (NULL allowed when npairs = 0);
 *             inactive rows are zero with h = +inf (collision_avoidance_limit.py:192-199) */
int bik_fk_jac(const bik_problem* problem, int B, const bik_inputs* in, float dt, float* J, float* e,
               float* e_posture, float* G_coll, float* h_coll, void* stream);

/* QP objective  H = damping*I + sum_t H_t,  c = sum_t c_t  with
 * H_t = (W J)^T (W J) + mu I, c_t = -(W(-gain e))^T (W J), mu = lm_damping ||W(-gain e)||^2.
 * Replaces Task.compute_qp_objective (task.py:105-138) + _compute_qp_objective (solve_ik.py:13-22).
 *   H [B][nv][nv], c [B][nv]  (fp64) */
int bik_qp_objective(const bik_problem* problem, int B, const float* J, const float* e,
                     const float* e_posture, double damping, double* H, double* c, void* stream);

/* Box form of the configuration + velocity limits:  lo <= dq <= hi  per dof ([B][nv], +-inf where
 * unbounded).  Replaces ConfigurationLimit / VelocityLimit.compute_qp_inequalities
 * (configuration_limit.py:98-124, velocity_limit.py:99-101); G is the constant +-projection. */
int bik_limits_box(const bik_problem* problem, int B, const float* q, float dt, float* lo, float* hi,
                   void* stream);

/* K2: assemble the QP from (J, e) and solve it exactly (active set on a warp-per-problem Cholesky).
 * Replaces build_ik + qpsolvers.solve_problem (solve_ik.py:43-65,101).  dq [B][nv] is the
 * displacement BEFORE the division by dt of solve_ik.py:104.  status [B] (may be NULL). */
int bik_solve(const bik_problem* problem, int B, const float* q, const float* J, const float* e,
              const float* e_posture, const float* G_coll, const float* h_coll, float dt,
              double damping, float* dq, int32_t* status, void* stream);

/* bik_solve plus diagnostics: iters [B] (may be NULL) receives the number of active-set iterations. */
int bik_solve_ex(const bik_problem* problem, int B, const float* q, const float* J, const float* e,
                 const float* e_posture, const float* G_coll, const float* h_coll, float dt,
                 double damping, float* dq, int32_t* status, int32_t* iters, void* stream);

/* q <- q (+) dq : Configuration.integrate / integrate_inplace (configuration.py:214-236). */
int bik_integrate(const bik_model* model, int B, float* q, const float* dq, void* stream);

/* Configuration.check_limits (configuration.py:77-110): sets BIK_STATUS_OUT_OF_LIMITS per instance. */
int bik_check_limits(const bik_model* model, int B, const float* q, float tol, int32_t* status,
                     void* stream);

/* The whole solve_ik step (solve_ik.py:68-105) `nsteps` times, two launches per step:
 *   K1 (check_limits + FK + task rows, packed: only the non-zero Jacobian columns) -> K2 (QP + integrate if `integrate` != 0).
 * `in->q` is ignored; q [B][nq] is read and, when integrating, updated in place.
 * dq [B][nv] receives the last step's displacement. */
int bik_step(const bik_problem* problem, int B, float* q, const bik_inputs* in, float dt,
             double damping, int nsteps, int integrate, float* dq, int32_t* status, void* stream);

/* Same step with HOST buffers: copies q/targets to the device, runs bik_step, copies dq (and q when
 * integrating) back, and synchronises.  This is the entry point a host-only caller (the reference's
 * numpy world) binds to; h2d/d2h byte counts are returned for measurement. */
int bik_step_host(const bik_problem* problem, int B, float* q_host, const bik_inputs* in_host,
                  float dt, double damping, int nsteps, int integrate, float* dq_host,
                  int32_t* status_host, size_t* h2d_bytes, size_t* d2h_bytes);

/* Converge-until-threshold driver: the solve_ik + integrate inner loop of the reference's examples
 * (examples/quadruped_spot.py:89-104, examples/arm_aloha.py:146-169), per instance:
 *     for i in range(max_iters):  v = solve_ik(...); q = integrate(q, v, dt)
 *                                 if all frame tasks: |e[:3]| <= pos_threshold and |e[3:]| <= ori_threshold: break
 * q [B][nq] is updated in place; an instance that met the thresholds keeps its configuration while the
 * others go on.  iters [B] receives the number of steps each instance took (1..max_iters); instances
 * that never met the thresholds get BIK_STATUS_NOT_CONVERGED.  The batch loop stops as soon as no
 * instance is left (the device counter is read back every `check_every` steps).  Device buffers. */
int bik_converge(const bik_problem* problem, int B, float* q, const bik_inputs* in, float dt,
                 double damping, int max_iters, float pos_threshold, float ori_threshold,
                 int check_every, int32_t* iters, int32_t* status, void* stream);

/* ---- fp64 entry points --------------------------------------------------------------------------------------------
 * The reference is fp64 end to end (numpy; mink/solve_ik.py:13-22).  These mirror the fp32 entry points above with
 * double buffers (bik_inputs.f64 = 1): the fp64 instantiations of the kernels run, on the kinematic constants at full
 * precision, and reproduce the reference to solver tolerance (~1e-8 rad) also where cost / sqrt(damping) would amplify
 * fp32 rounding past 1e-4 rad (BASELINE config 5).  A single (numpy, B = 1) Configuration of the Python front end
 * always takes them.  The fp32 entry points pick the fp64 K1 by themselves for ill-conditioned problems
 * (bik_problem_describe says which; BIK_K1_PRECISION=f32|f64 overrides). */
int bik_fk64(const bik_model* model, int B, const double* q, const bik_frame* frames, int nframes,
             double* poses, double* com, void* stream);
int bik_frame_jacobian64(const bik_model* model, int B, const double* q, const bik_frame* frames,
                         int nframes, double* J, void* stream);
int bik_fk_jac64(const bik_problem* problem, int B, const bik_inputs* in, double dt, double* J, double* e,
                 double* e_posture, double* G_coll, double* h_coll, void* stream);
int bik_qp_objective64(const bik_problem* problem, int B, const double* J, const double* e,
                       const double* e_posture, double damping, double* H, double* c, void* stream);
int bik_limits_box64(const bik_problem* problem, int B, const double* q, double dt, double* lo, double* hi,
                     void* stream);
int bik_solve64(const bik_problem* problem, int B, const double* q, const double* J, const double* e,
                const double* e_posture, const double* G_coll, const double* h_coll, double dt,
                double damping, double* dq, int32_t* status, int32_t* iters, void* stream);
int bik_integrate64(const bik_model* model, int B, double* q, const double* dq, void* stream);
int bik_check_limits64(const bik_model* model, int B, const double* q, double tol, int32_t* status,
                       void* stream);
int bik_step64(const bik_problem* problem, int B, double* q, const bik_inputs* in, double dt,
               double damping, int nsteps, int integrate, double* dq, int32_t* status, void* stream);

/* Measurement aid for the K2 roofline (bench.py): sustained fused-multiply-add throughput of the CUDA cores of `device`
 * in TFLOP/s (fp32 or fp64; 16 independent chains per thread, every SM filled), best of `repeats` event-timed launches. */
int bik_measure_fma_peak(int device, int use_double, int repeats, double* tflops);

/* Bytes of device scratch a problem needs for a batch of B (K1's packed task rows, collision rows, ... between K1 and K2). */
size_t bik_workspace_bytes(const bik_problem* problem, int B);

/* One-line description of how a problem is mapped onto the device (lanes per instance and visited
 * nodes in K1, coupled block size, K2 path and precision chosen for `damping`), for logs and
 * benchmark labels.  Writes at most cap bytes including the terminator; returns the full length. */
int bik_problem_describe(const bik_problem* problem, double damping, char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* BIK_H */
