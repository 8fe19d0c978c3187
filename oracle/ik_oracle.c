/* ORACLE / TEST INFRASTRUCTURE -- fp64 C restatement of mink's solve_ik path.
 *
 * NOT the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; mink_b200/ never does.
 *
 * What is restated, and from where (all paths relative to /root/reference):
 *   FK, frame poses ............ mujoco mj_kinematics as called at mink/configuration.py:63,
 *                                frame pose read at configuration.py:157-185
 *   point Jacobians ............ mujoco mj_jac/mj_jacBody/Site/Geom via configuration.py:144-145,
 *                                rotated into the frame at configuration.py:150-153
 *   SE3/SO3 log, jlog, _getQ ... mink/lie/so3.py:176-226, mink/lie/se3.py:159-249, lie/base.py:113-156
 *   FrameTask e, J ............. mink/tasks/frame_task.py:95-146
 *   PostureTask e .............. mink/tasks/posture_task.py:87-118 (mj_differentiatePos)
 *   ComTask e, J ............... mink/tasks/com_task.py:71-97 (mj_jacSubtreeCom, body 1)
 *   H, c ....................... mink/tasks/task.py:105-138, mink/solve_ik.py:13-22
 *   limits ..................... mink/limits/configuration_limit.py:69-124, velocity_limit.py:71-101,
 *                                collision_avoidance_limit.py:187-210
 *   QP ......................... qpsolvers.solve_problem at mink/solve_ik.py:101; restated as the
 *                                Goldfarb-Idnani dual active-set method (quadprog's algorithm)
 *   integrate .................. mujoco mj_integratePos via configuration.py:225,235
 * MuJoCo (>=3.1.6) and qpsolvers (>=4.3.1) are third-party and absent from /root/reference; see
 * oracle/mujoco/__init__.py for the published-algorithm notes.
 *
 * PINNING: this file is checked in tests/test_oracle_golden.py against tests/golden/*.npz, which
 * were produced by the unmodified reference Python running on the numpy shims (oracle/gen_golden.py),
 * themselves gated by the reference's own test-suite (oracle/run_reference_tests.py).
 *
 * Build: make -C oracle   ->  oracle/_build/libikoracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/bik.h"

#define EPS64 1e-10 /* mink/lie/utils.py:4-8, float64 */
#define MJ_MINVAL 1e-15

/* ------------------------------------------------------------------------- */
/* blob access                                                                */
/* ------------------------------------------------------------------------- */
typedef struct {
  int nq, nv, nnode, ncom;
  const int32_t *node_parent, *node_type, *node_qadr, *node_dadr, *dof_node, *dof_qadr, *dof_limited,
      *com_node;
  const double *node_pos, *node_quat, *node_axis, *node_jpos, *qpos0, *dof_lo, *dof_hi, *com_pos,
      *com_mass;
} Model;

static const void* find_section(const uint8_t* blob, const char* name) {
  const uint32_t* h = (const uint32_t*)blob;
  uint32_t nsec = h[3];
  const uint8_t* t = blob + 48;
  for (uint32_t s = 0; s < nsec; ++s, t += 32) {
    if (strncmp((const char*)t, name, 20) == 0) {
      uint32_t off = *(const uint32_t*)(t + 28);
      return blob + off;
    }
  }
  return NULL;
}

static int model_from_blob(const void* blob_, Model* m) {
  const uint8_t* blob = (const uint8_t*)blob_;
  const uint32_t* h = (const uint32_t*)blob;
  if (h[0] != 0x4D4B4942u || h[1] != 1u) return -1;
  const int32_t* hi = (const int32_t*)(blob + 16);
  m->nq = hi[0]; m->nv = hi[1]; m->nnode = hi[2]; m->ncom = hi[3];
#define SEC(f) m->f = find_section(blob, #f); if (!m->f) return -1;
  SEC(node_parent) SEC(node_type) SEC(node_qadr) SEC(node_dadr) SEC(node_pos) SEC(node_quat)
  SEC(node_axis) SEC(node_jpos) SEC(qpos0) SEC(dof_node) SEC(dof_qadr) SEC(dof_limited) SEC(dof_lo)
  SEC(dof_hi) SEC(com_node) SEC(com_pos) SEC(com_mass)
#undef SEC
  return 0;
}

/* ------------------------------------------------------------------------- */
/* small vector / quaternion algebra                                          */
/* ------------------------------------------------------------------------- */
static void qmul(const double* a, const double* b, double* r) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void qnormalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MJ_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void q2mat(const double* q, double* R) { /* mju_quat2Mat, row major */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static void mat2quat(const double* m, double* q) { /* mju_mat2Quat */
  double tr = m[0] + m[4] + m[8];
  if (tr > 0) {
    q[0] = 0.5 * sqrt(1 + tr); double k = 0.25 / q[0];
    q[1] = k * (m[7] - m[5]); q[2] = k * (m[2] - m[6]); q[3] = k * (m[3] - m[1]);
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = 0.5 * sqrt(1 + m[0] - m[4] - m[8]); double k = 0.25 / q[1];
    q[0] = k * (m[7] - m[5]); q[2] = k * (m[3] + m[1]); q[3] = k * (m[2] + m[6]);
  } else if (m[4] > m[8]) {
    q[2] = 0.5 * sqrt(1 - m[0] + m[4] - m[8]); double k = 0.25 / q[2];
    q[0] = k * (m[2] - m[6]); q[1] = k * (m[3] + m[1]); q[3] = k * (m[7] + m[5]);
  } else {
    q[3] = 0.5 * sqrt(1 - m[0] - m[4] + m[8]); double k = 0.25 / q[3];
    q[0] = k * (m[3] - m[1]); q[1] = k * (m[2] + m[6]); q[2] = k * (m[7] + m[5]);
  }
  qnormalize(q);
}
static void mat_vec(const double* R, const double* v, double* r) {
  double a = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double b = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double c = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = a; r[1] = b; r[2] = c;
}
static void matT_vec(const double* R, const double* v, double* r) {
  double a = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double b = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double c = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = a; r[1] = b; r[2] = c;
}
static void cross(const double* a, const double* b, double* r) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static void skew(const double* w, double* S) {
  S[0] = 0; S[1] = -w[2]; S[2] = w[1]; S[3] = w[2]; S[4] = 0; S[5] = -w[0]; S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}
static void mm3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, t, sizeof t);
}
static void tr3(const double* A, double* T) {
  double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
  memcpy(T, t, sizeof t);
}

/* ------------------------------------------------------------------------- */
/* Lie algebra pieces of the FrameTask path                                   */
/* ------------------------------------------------------------------------- */
static void so3_log(const double* q, double* w) { /* so3.py:176-191 */
  double qw = q[0], nsq = q[1] * q[1] + q[2] * q[2] + q[3] * q[3], f;
  if (nsq < EPS64) {
    f = 2.0 / qw - 2.0 / 3.0 * nsq / (qw * qw * qw);
  } else {
    double n = sqrt(nsq);
    if (fabs(qw) < EPS64) f = (qw > 0 ? 1.0 : -1.0) * M_PI / n;
    else f = 2.0 * atan2(qw < 0 ? -n : n, fabs(qw)) / n;
  }
  w[0] = f * q[1]; w[1] = f * q[2]; w[2] = f * q[3];
}
static void se3_log(const double* q, const double* t, double* xi) { /* se3.py:159-185 */
  double w[3], S[9], S2[9], Vi[9];
  so3_log(q, w);
  double tsq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  skew(w, S); mm3(S, S, S2);
  double k;
  if (tsq < EPS64) k = 1.0 / 12.0;
  else { double th = sqrt(tsq), h = 0.5 * th; k = (1.0 - th * cos(h) / (2.0 * sin(h))) / tsq; }
  for (int i = 0; i < 9; ++i) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * S[i] + k * S2[i];
  mat_vec(Vi, t, xi);
  xi[3] = w[0]; xi[4] = w[1]; xi[5] = w[2];
}
static void so3_ljacinv(const double* w, double* Ji) { /* so3.py:215-226 (threshold on theta) */
  double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), A, S[9], S2[9];
  if (th < EPS64) { double t2 = th * th; A = (1.0 / 12.0) * (1.0 + t2 / 60.0 * (1.0 + t2 / 42.0 * (1.0 + t2 / 40.0))); }
  else A = (1.0 / (th * th)) * (1.0 - (th * sin(th) / (2.0 * (1.0 - cos(th)))));
  skew(w, S); mm3(S, S, S2);
  for (int i = 0; i < 9; ++i) Ji[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * S[i] + A * S2[i];
}
static void se3_getQ(const double* c, double* Q) { /* se3.py:222-249 */
  double tsq = c[3] * c[3] + c[4] * c[4] + c[5] * c[5], A = 0.5, Bc, Cc, Dc;
  if (tsq < EPS64) { Bc = 1.0 / 6.0 + tsq / 120.0; Cc = -1.0 / 24.0 + tsq / 720.0; Dc = -1.0 / 60.0; }
  else {
    double th = sqrt(tsq), s = sin(th), co = cos(th);
    Bc = (th - s) / (tsq * th); Cc = (1.0 - tsq / 2.0 - co) / (tsq * tsq);
    Dc = (2 * th - 3 * s + th * co) / (2 * tsq * tsq * th);
  }
  double V[9], W[9], VW[9], WV[9], WVW[9], VWW[9], VWWt[9], T1[9], T2[9];
  skew(c, V); skew(c + 3, W);
  mm3(V, W, VW); tr3(VW, WV); mm3(WV, W, WVW); mm3(VW, W, VWW); tr3(VWW, VWWt);
  mm3(WVW, W, T1); mm3(W, WVW, T2);
  for (int i = 0; i < 9; ++i)
    Q[i] = A * V[i] + Bc * (WV[i] + VW[i] + WVW[i]) - Cc * (VWW[i] - VWWt[i] - 3 * WVW[i]) + Dc * (T1[i] + T2[i]);
}
static void se3_ljacinv(const double* xi, double* J6) { /* se3.py:211-218; 6x6 row major */
  double tsq = xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5];
  memset(J6, 0, 36 * sizeof(double));
  if (tsq < EPS64) { for (int i = 0; i < 6; ++i) J6[7 * i] = 1.0; return; }
  double Q[9], Ji[9], T[9], M[9];
  se3_getQ(xi, Q); so3_ljacinv(xi + 3, Ji);
  mm3(Ji, Q, T); mm3(T, Ji, M);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      J6[6 * i + j] = Ji[3 * i + j]; J6[6 * (i + 3) + j + 3] = Ji[3 * i + j]; J6[6 * i + j + 3] = -M[3 * i + j];
    }
}

/* ------------------------------------------------------------------------- */
/* forward kinematics over the node list (mj_kinematics restated per node)    */
/* ------------------------------------------------------------------------- */
typedef struct { double *xpos, *xquat, *xmat, *anchor, *axis; } Kin; /* per node */

static void fk_nodes(const Model* m, const double* q, Kin* k) {
  for (int n = 0; n < m->nnode; ++n) {
    int t = m->node_type[n], a = m->node_qadr[n], p = m->node_parent[n];
    double pos[3], quat[4], R[9];
    if (t == 0) { /* free: pose straight from qpos, quaternion normalised */
      memcpy(pos, q + a, 3 * sizeof(double)); memcpy(quat, q + a + 3, 4 * sizeof(double)); qnormalize(quat);
      memcpy(k->anchor + 3 * n, pos, sizeof pos);
      k->axis[3 * n] = 0; k->axis[3 * n + 1] = 0; k->axis[3 * n + 2] = 1;
    } else {
      if (p >= 0) {
        double off[3]; mat_vec(k->xmat + 9 * p, m->node_pos + 3 * n, off);
        for (int i = 0; i < 3; ++i) pos[i] = k->xpos[3 * p + i] + off[i];
        qmul(k->xquat + 4 * p, m->node_quat + 4 * n, quat);
      } else {
        memcpy(pos, m->node_pos + 3 * n, sizeof pos); memcpy(quat, m->node_quat + 4 * n, sizeof quat);
      }
      q2mat(quat, R);
      double jp[3]; mat_vec(R, m->node_jpos + 3 * n, jp);
      for (int i = 0; i < 3; ++i) k->anchor[3 * n + i] = pos[i] + jp[i];
      mat_vec(R, m->node_axis + 3 * n, k->axis + 3 * n);
      if (t == 2) { /* slide */
        double d = q[a] - m->qpos0[a];
        for (int i = 0; i < 3; ++i) pos[i] += k->axis[3 * n + i] * d;
      } else {
        double ql[4];
        if (t == 3) { /* hinge */
          double ang = q[a] - m->qpos0[a], s = sin(0.5 * ang);
          ql[0] = cos(0.5 * ang); ql[1] = m->node_axis[3 * n] * s; ql[2] = m->node_axis[3 * n + 1] * s; ql[3] = m->node_axis[3 * n + 2] * s;
        } else { memcpy(ql, q + a, sizeof ql); qnormalize(ql); } /* ball */
        double qn[4]; qmul(quat, ql, qn); memcpy(quat, qn, sizeof qn);
        q2mat(quat, R); mat_vec(R, m->node_jpos + 3 * n, jp);
        for (int i = 0; i < 3; ++i) pos[i] = k->anchor[3 * n + i] - jp[i];
      }
    }
    qnormalize(quat);
    memcpy(k->xpos + 3 * n, pos, sizeof pos); memcpy(k->xquat + 4 * n, quat, sizeof quat);
    q2mat(quat, k->xmat + 9 * n);
  }
}

static void frame_pose(const Kin* k, const bik_frame* f, double* p, double* R) {
  if (f->node < 0) { memcpy(p, f->pos, 3 * sizeof(double)); double qq[4]; memcpy(qq, f->quat, sizeof qq); qnormalize(qq); q2mat(qq, R); return; }
  double off[3], qq[4];
  mat_vec(k->xmat + 9 * f->node, f->pos, off);
  for (int i = 0; i < 3; ++i) p[i] = k->xpos[3 * f->node + i] + off[i];
  qmul(k->xquat + 4 * f->node, f->quat, qq); qnormalize(qq); q2mat(qq, R);
}

/* world-aligned point Jacobian of `point` rigidly attached to `node` (mj_jac): jacp/jacr [3][nv] */
static void point_jac(const Model* m, const Kin* k, int node, const double* point, double* jacp, double* jacr) {
  int nv = m->nv;
  if (jacp) memset(jacp, 0, 3 * nv * sizeof(double));
  if (jacr) memset(jacr, 0, 3 * nv * sizeof(double));
  for (int n = node; n >= 0; n = m->node_parent[n]) {
    int t = m->node_type[n], d = m->node_dadr[n];
    double r[3] = {point[0] - k->anchor[3 * n], point[1] - k->anchor[3 * n + 1], point[2] - k->anchor[3 * n + 2]};
    if (t == 3) {
      double c[3]; cross(k->axis + 3 * n, r, c);
      for (int i = 0; i < 3; ++i) { if (jacp) jacp[i * nv + d] = c[i]; if (jacr) jacr[i * nv + d] = k->axis[3 * n + i]; }
    } else if (t == 2) {
      for (int i = 0; i < 3; ++i) if (jacp) jacp[i * nv + d] = k->axis[3 * n + i];
    } else {
      int ro = (t == 0) ? 3 : 0; /* free: 3 world-axis translations first */
      if (t == 0) for (int i = 0; i < 3; ++i) if (jacp) jacp[i * nv + d + i] = 1.0;
      for (int c3 = 0; c3 < 3; ++c3) { /* rotational dofs about the body-frame axes */
        double ax[3] = {k->xmat[9 * n + c3], k->xmat[9 * n + 3 + c3], k->xmat[9 * n + 6 + c3]}, c[3];
        cross(ax, r, c);
        for (int i = 0; i < 3; ++i) { if (jacp) jacp[i * nv + d + ro + c3] = c[i]; if (jacr) jacr[i * nv + d + ro + c3] = ax[i]; }
      }
    }
  }
}

static Kin kin_alloc(int nnode) {
  Kin k;
  k.xpos = (double*)malloc(sizeof(double) * nnode * (3 + 4 + 9 + 3 + 3));
  k.xquat = k.xpos + 3 * nnode; k.xmat = k.xquat + 4 * nnode; k.anchor = k.xmat + 9 * nnode; k.axis = k.anchor + 3 * nnode;
  return k;
}
static void kin_free(Kin* k) { free(k->xpos); }

static void com_of(const Model* m, const Kin* k, double* com, double* total) {
  double acc[3] = {0, 0, 0}, M = 0;
  for (int c = 0; c < m->ncom; ++c) {
    double p[3]; int n = m->com_node[c];
    if (n >= 0) { mat_vec(k->xmat + 9 * n, m->com_pos + 3 * c, p); for (int i = 0; i < 3; ++i) p[i] += k->xpos[3 * n + i]; }
    else memcpy(p, m->com_pos + 3 * c, sizeof p);
    for (int i = 0; i < 3; ++i) acc[i] += m->com_mass[c] * p[i];
    M += m->com_mass[c];
  }
  for (int i = 0; i < 3; ++i) com[i] = M > 0 ? acc[i] / M : 0.0;
  if (total) *total = M;
}

/* ------------------------------------------------------------------------- */
/* public: FK / frame Jacobian                                                */
/* ------------------------------------------------------------------------- */
int iko_fk(const void* blob, int B, const double* q, const bik_frame* frames, int nframes, double* poses, double* com) {
  Model m; if (model_from_blob(blob, &m)) return -1;
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode);
#pragma omp for
    for (int b = 0; b < B; ++b) {
      fk_nodes(&m, q + (size_t)b * m.nq, &k);
      for (int f = 0; f < nframes; ++f) {
        double p[3], R[9], qq[4]; frame_pose(&k, frames + f, p, R); mat2quat(R, qq);
        double* o = poses + ((size_t)b * nframes + f) * 7;
        memcpy(o, qq, sizeof qq); memcpy(o + 4, p, sizeof p);
      }
      if (com) com_of(&m, &k, com + (size_t)b * 3, NULL);
    }
    kin_free(&k);
  }
  return 0;
}

static void body_jacobian(const Model* m, const Kin* k, const bik_frame* f, double* Jb /*6 x nv*/, double* p, double* R, double* scratch /*6 nv*/) {
  int nv = m->nv;
  frame_pose(k, f, p, R);
  double *jp = scratch, *jr = scratch + 3 * nv;
  point_jac(m, k, f->node, p, jp, jr);
  for (int d = 0; d < nv; ++d) { /* blockdiag(R^T, R^T) [jacp; jacr]  (configuration.py:150-153) */
    double a[3] = {jp[d], jp[nv + d], jp[2 * nv + d]}, w[3] = {jr[d], jr[nv + d], jr[2 * nv + d]}, ra[3], rw[3];
    matT_vec(R, a, ra); matT_vec(R, w, rw);
    for (int i = 0; i < 3; ++i) { Jb[i * nv + d] = ra[i]; Jb[(i + 3) * nv + d] = rw[i]; }
  }
}

int iko_frame_jacobian(const void* blob, int B, const double* q, const bik_frame* frames, int nframes, double* J) {
  Model m; if (model_from_blob(blob, &m)) return -1;
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode); double* s = (double*)malloc(sizeof(double) * 6 * m.nv);
#pragma omp for
    for (int b = 0; b < B; ++b) {
      fk_nodes(&m, q + (size_t)b * m.nq, &k);
      for (int f = 0; f < nframes; ++f) { double p[3], R[9]; body_jacobian(&m, &k, frames + f, J + ((size_t)b * nframes + f) * 6 * m.nv, p, R, s); }
    }
    kin_free(&k); free(s);
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* tasks                                                                      */
/* ------------------------------------------------------------------------- */
typedef struct { int F, P, C, K; } Counts;
static Counts count_tasks(const bik_task_desc* t, int n) {
  Counts c = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i) { if (t[i].kind == BIK_TASK_FRAME || t[i].kind == BIK_TASK_RELATIVE_FRAME) c.F++; else if (t[i].kind == BIK_TASK_POSTURE) c.P++; else c.C++; }
  c.K = 6 * c.F + 3 * c.C; return c;
}

static void sub_quat(const double* qa, const double* qb, double* w) { /* mju_subQuat */
  double qc[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, d[4]; qmul(qc, qa, d);
  double s = sqrt(d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
  if (s < MJ_MINVAL) { w[0] = w[1] = w[2] = 0; return; }
  double ang = 2.0 * atan2(s, d[0]); if (ang > M_PI) ang -= 2.0 * M_PI;
  for (int i = 0; i < 3; ++i) w[i] = d[i + 1] / s * ang;
}

/* one instance: J [K][nv], e [K], e_posture [P][nv] */
static void tasks_instance(const Model* m, const Kin* k, const bik_task_desc* tasks, int ntasks, const double* q,
                           const double* ftgt, const double* ptgt, const double* ctgt, double* J, double* e, double* ep, double* scratch) {
  int nv = m->nv, row = 0, fi = 0, pi = 0, ci = 0;
  for (int t = 0; t < ntasks; ++t) {
    const bik_task_desc* T = tasks + t;
    if (T->kind == BIK_TASK_FRAME) {
      double p[3], R[9]; double* Jbp = scratch + 6 * nv;
      body_jacobian(m, k, &T->frame, Jbp, p, R, scratch);
      const double* tg = ftgt + 7 * fi;
      double tq[4] = {tg[0], tg[1], tg[2], tg[3]}, Rt[9], qb[4];
      q2mat(tq, Rt); mat2quat(R, qb); /* frame rotation goes through mju_mat2Quat (configuration.py:182) */
      /* e = log(T_wb^-1 T_wt)  (frame_task.py:119-122, base.py:113-114) */
      double qbi[4] = {qb[0], -qb[1], -qb[2], -qb[3]}, qbt[4], dp[3] = {tg[4] - p[0], tg[5] - p[1], tg[6] - p[2]}, tbt[3];
      qmul(qbi, tq, qbt);
      { double Rb[9]; q2mat(qb, Rb); matT_vec(Rb, dp, tbt); }
      se3_log(qbt, tbt, e + row);
      /* J = -jlog(T_wt^-1 T_wb) Jb  (frame_task.py:145-146, base.py:151-156) */
      double tqi[4] = {tq[0], -tq[1], -tq[2], -tq[3]}, qtb[4], dq_[3] = {p[0] - tg[4], p[1] - tg[5], p[2] - tg[6]}, ttb[3], xi[6], nxi[6], L[36];
      qmul(tqi, qb, qtb); matT_vec(Rt, dq_, ttb);
      se3_log(qtb, ttb, xi);
      for (int i = 0; i < 6; ++i) nxi[i] = -xi[i];
      se3_ljacinv(nxi, L); /* rjacinv(xi) = ljacinv(-xi) */
      for (int i = 0; i < 6; ++i)
        for (int d = 0; d < nv; ++d) {
          double s = 0; for (int j = 0; j < 6; ++j) s += L[6 * i + j] * Jbp[j * nv + d];
          J[(row + i) * nv + d] = -s;
        }
      row += 6; fi++;
    } else if (T->kind == BIK_TASK_RELATIVE_FRAME) { /* relative_frame_task.py:106-142 */
      double pf[3], Rf[9], pr[3], Rr[9]; double* Jf = scratch + 6 * nv; double* Jr = scratch + 12 * nv;
      body_jacobian(m, k, &T->frame, Jf, pf, Rf, scratch);
      body_jacobian(m, k, &T->root, Jr, pr, Rr, scratch);
      double qf[4], qr[4]; mat2quat(Rf, qf); mat2quat(Rr, qr);
      /* T_rf = T_wr^-1 T_wf */
      double qri[4] = {qr[0], -qr[1], -qr[2], -qr[3]}, q_rf[4], d_[3] = {pf[0] - pr[0], pf[1] - pr[1], pf[2] - pr[2]}, t_rf[3];
      qmul(qri, qf, q_rf); matT_vec(Rr, d_, t_rf);
      const double* tg = ftgt + 7 * fi;
      double tq[4] = {tg[0], tg[1], tg[2], tg[3]}, tqi[4] = {tg[0], -tg[1], -tg[2], -tg[3]}, Rt[9], q_tf[4], dd[3] = {t_rf[0] - tg[4], t_rf[1] - tg[5], t_rf[2] - tg[6]}, t_tf[3];
      q2mat(tq, Rt); qmul(tqi, q_rf, q_tf); matT_vec(Rt, dd, t_tf);
      double xi[6], nxi[6], L[36];
      se3_log(q_tf, t_tf, xi);                       /* e = T_rf.rminus(T_rt) = log(T_rt^-1 T_rf) */
      memcpy(e + row, xi, sizeof xi);
      for (int i = 0; i < 6; ++i) nxi[i] = -xi[i];
      se3_ljacinv(nxi, L);                           /* jlog(T_tf) */
      /* Ad(T_fr), T_fr = T_rf^-1: R = R_rf^T, t = -R_rf^T t_rf */
      double Rrf[9], Rfr[9], tfr[3], S[9], SR[9];
      q2mat(q_rf, Rrf); tr3(Rrf, Rfr); mat_vec(Rfr, t_rf, tfr); for (int i = 0; i < 3; ++i) tfr[i] = -tfr[i];
      skew(tfr, S); mm3(S, Rfr, SR);
      for (int d = 0; d < nv; ++d) {
        double vr[3] = {Jr[d], Jr[nv + d], Jr[2 * nv + d]}, wr[3] = {Jr[3 * nv + d], Jr[4 * nv + d], Jr[5 * nv + d]}, a1[3], a2[3], a3[3], tw[6];
        mat_vec(Rfr, vr, a1); mat_vec(SR, wr, a2); mat_vec(Rfr, wr, a3);
        for (int i = 0; i < 3; ++i) { tw[i] = Jf[i * nv + d] - (a1[i] + a2[i]); tw[3 + i] = Jf[(3 + i) * nv + d] - a3[i]; }
        for (int i = 0; i < 6; ++i) { double sacc = 0; for (int j = 0; j < 6; ++j) sacc += L[6 * i + j] * tw[j]; J[(row + i) * nv + d] = sacc; }
      }
      row += 6; fi++;
    } else if (T->kind == BIK_TASK_COM) {
      double com[3], M; com_of(m, k, com, &M);
      const double* tg = ctgt + 3 * ci;
      for (int i = 0; i < 3; ++i) e[row + i] = com[i] - tg[i]; /* com_task.py:82 */
      memset(J + row * nv, 0, sizeof(double) * 3 * nv);
      for (int c = 0; c < m->ncom; ++c) { /* mj_jacSubtreeCom(body 1) */
        int n = m->com_node[c]; if (n < 0) continue;
        double pnt[3]; mat_vec(k->xmat + 9 * n, m->com_pos + 3 * c, pnt); for (int i = 0; i < 3; ++i) pnt[i] += k->xpos[3 * n + i];
        point_jac(m, k, n, pnt, scratch, NULL);
        for (int i = 0; i < 3 * nv; ++i) J[row * nv + i] += scratch[i] * m->com_mass[c] / M;
      }
      row += 3; ci++;
    } else { /* posture: e = q* (-) q with free dofs zeroed (posture_task.py:107-118) */
      const double* tg = ptgt + (size_t)pi * m->nq; double* o = ep + (size_t)pi * nv;
      for (int n = 0; n < m->nnode; ++n) {
        int ty = m->node_type[n], a = m->node_qadr[n], d = m->node_dadr[n];
        if (ty == 0) for (int i = 0; i < 6; ++i) o[d + i] = 0.0;
        else if (ty == 1) sub_quat(tg + a, q + a, o + d);
        else o[d] = tg[a] - q[a];
      }
      pi++;
    }
  }
}

int iko_fk_jac(const void* blob, const bik_task_desc* tasks, int ntasks, int B, const double* q, const double* frame_targets,
               const double* posture_targets, int posture_batched, const double* com_targets, double* J, double* e, double* e_posture) {
  Model m; if (model_from_blob(blob, &m)) return -1;
  Counts c = count_tasks(tasks, ntasks);
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode); double* s = (double*)malloc(sizeof(double) * 18 * m.nv);
#pragma omp for
    for (int b = 0; b < B; ++b) {
      const double* qb = q + (size_t)b * m.nq;
      fk_nodes(&m, qb, &k);
      tasks_instance(&m, &k, tasks, ntasks, qb, frame_targets ? frame_targets + (size_t)b * c.F * 7 : NULL,
                     posture_targets ? posture_targets + (posture_batched ? (size_t)b * c.P * m.nq : 0) : NULL,
                     com_targets ? com_targets + (size_t)b * c.C * 3 : NULL, J + (size_t)b * c.K * m.nv, e + (size_t)b * c.K,
                     e_posture ? e_posture + (size_t)b * c.P * m.nv : NULL, s);
    }
    kin_free(&k); free(s);
  }
  return 0;
}

/* H = damping I + sum_t (WJ)^T(WJ) + mu I ; c = sum_t -(W(-gain e))^T (WJ)   (task.py:125-138) */
static void objective_instance(const Model* m, const bik_task_desc* tasks, int ntasks, const double* J, const double* e,
                               const double* ep, double damping, double* H, double* c) {
  int nv = m->nv, row = 0, pi = 0;
  memset(H, 0, sizeof(double) * nv * nv); memset(c, 0, sizeof(double) * nv);
  for (int i = 0; i < nv; ++i) H[i * nv + i] = damping;
  for (int t = 0; t < ntasks; ++t) {
    const bik_task_desc* T = tasks + t;
    if (T->kind == BIK_TASK_POSTURE) {
      const double* o = ep + (size_t)pi * nv; double mu = 0;
      for (int d = 0; d < nv; ++d) {
        int isfree = m->node_type[m->dof_node[d]] == 0;
        double w = T->dof_cost[d], we = w * (-T->gain * o[d]);
        mu += we * we;
        if (!isfree) { H[d * nv + d] += w * w; c[d] += -we * (-w); }
      }
      mu *= T->lm_damping;
      for (int d = 0; d < nv; ++d) H[d * nv + d] += mu;
      pi++; continue;
    }
    int k = (T->kind == BIK_TASK_FRAME || T->kind == BIK_TASK_RELATIVE_FRAME) ? 6 : 3; double we[6], mu = 0;
    for (int i = 0; i < k; ++i) { we[i] = T->cost[i] * (-T->gain * e[row + i]); mu += we[i] * we[i]; }
    mu *= T->lm_damping;
    for (int a = 0; a < nv; ++a) {
      for (int b2 = 0; b2 < nv; ++b2) {
        double s = 0; for (int i = 0; i < k; ++i) s += T->cost[i] * T->cost[i] * J[(row + i) * nv + a] * J[(row + i) * nv + b2];
        H[a * nv + b2] += s;
      }
      H[a * nv + a] += mu;
      double s = 0; for (int i = 0; i < k; ++i) s += we[i] * T->cost[i] * J[(row + i) * nv + a];
      c[a] -= s;
    }
    row += k;
  }
}

int iko_objective(const void* blob, const bik_task_desc* tasks, int ntasks, int B, const double* J, const double* e,
                  const double* e_posture, double damping, double* H, double* c) {
  Model m; if (model_from_blob(blob, &m)) return -1;
  Counts cn = count_tasks(tasks, ntasks);
#pragma omp parallel for
  for (int b = 0; b < B; ++b)
    objective_instance(&m, tasks, ntasks, J + (size_t)b * cn.K * m.nv, e + (size_t)b * cn.K,
                       e_posture ? e_posture + (size_t)b * cn.P * m.nv : NULL, damping, H + (size_t)b * m.nv * m.nv, c + (size_t)b * m.nv);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* limits                                                                     */
/* ------------------------------------------------------------------------- */
static void box_instance(const Model* m, const bik_limit_desc* limits, int nlimits, const double* q, double dt, double* lo, double* hi) {
  for (int d = 0; d < m->nv; ++d) { lo[d] = -INFINITY; hi[d] = INFINITY; }
  for (int l = 0; l < nlimits; ++l) {
    const bik_limit_desc* L = limits + l;
    for (int i = 0; i < L->n; ++i) {
      if (L->kind == BIK_LIMIT_COLLISION) break;
      int d = L->dof[i]; double up, dn;
      if (L->kind == BIK_LIMIT_CONFIGURATION) { /* configuration_limit.py:98-124 */
        double qi = q[m->dof_qadr[d]];
        up = L->gain * (L->upper[i] - qi); dn = L->gain * (qi - L->lower[i]);
      } else { up = dn = dt * L->vmax[i]; } /* velocity_limit.py:99-101 */
      if (up < hi[d]) hi[d] = up;
      if (-dn > lo[d]) lo[d] = -dn;
    }
  }
}

int iko_box(const void* blob, const bik_limit_desc* limits, int nlimits, int B, const double* q, double dt, double* lo, double* hi) {
  Model m; if (model_from_blob(blob, &m)) return -1;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) box_instance(&m, limits, nlimits, q + (size_t)b * m.nq, dt, lo + (size_t)b * m.nv, hi + (size_t)b * m.nv);
  return 0;
}

/* signed distance between primitive geoms (plane / sphere / capsule / box), fromto on (g1, g2) */
static void seg_closest(const double* p1, const double* d1, const double* p2, const double* d2, double* a, double* b) {
  double r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  double A = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2], E = d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2];
  double Bq = d1[0] * d2[0] + d1[1] * d2[1] + d1[2] * d2[2], C = d1[0] * r[0] + d1[1] * r[1] + d1[2] * r[2], F = d2[0] * r[0] + d2[1] * r[1] + d2[2] * r[2];
  double den = A * E - Bq * Bq, s = 0, t = 0;
  if (den > 1e-14) { s = (Bq * F - C * E) / den; if (s < -1) s = -1; if (s > 1) s = 1; }
  if (E > 1e-14) t = (Bq * s + F) / E;
  if (t < -1 || t > 1) { t = t < -1 ? -1 : 1; if (A > 1e-14) { s = (Bq * t - C) / A; if (s < -1) s = -1; if (s > 1) s = 1; } else s = 0; }
  for (int i = 0; i < 3; ++i) { a[i] = p1[i] + s * d1[i]; b[i] = p2[i] + t * d2[i]; }
}
/* squared distance from c + t d to the box |x_k| <= s_k (box frame) */
static double seg_box_f(const double* c, const double* d, const double* s, double t) {
  double f = 0;
  for (int k = 0; k < 3; ++k) { double e = fabs(c[k] + t * d[k]) - s[k]; if (e > 0) f += e * e; }
  return f;
}
/* its minimiser over t in [-1, 1]: piecewise quadratic between the face-plane crossings, each piece minimised in closed form */
static double seg_box_param(const double* c, const double* d, const double* s) {
  double t0 = -1, t1 = 1; int hit = 1;   /* segment through the box: the middle of the part inside */
  for (int k = 0; k < 3; ++k) {
    if (d[k] == 0) { if (fabs(c[k]) > s[k]) hit = 0; continue; }
    double a = (-s[k] - c[k]) / d[k], b = (s[k] - c[k]) / d[k];
    if (a > b) { double x = a; a = b; b = x; }
    if (a > t0) t0 = a;
    if (b < t1) t1 = b;
  }
  if (hit && t0 <= t1) return 0.5 * (t0 + t1);
  double cuts[8]; int nc = 0;
  cuts[nc++] = -1; cuts[nc++] = 1;
  for (int k = 0; k < 3; ++k) if (d[k] != 0) for (int sg = -1; sg <= 1; sg += 2) { double t = (sg * s[k] - c[k]) / d[k]; if (t > -1 && t < 1) cuts[nc++] = t; }
  for (int i = 1; i < nc; ++i) { double v = cuts[i]; int j = i - 1; while (j >= 0 && cuts[j] > v) { cuts[j + 1] = cuts[j]; --j; } cuts[j + 1] = v; }
  double bt = -1, bf = seg_box_f(c, d, s, -1);
  for (int i = 0; i + 1 < nc; ++i) {
    double a = cuts[i], b = cuts[i + 1], tm = 0.5 * (a + b), num = 0, den = 0;
    for (int k = 0; k < 3; ++k) { double pm = c[k] + tm * d[k]; if (fabs(pm) > s[k]) { double sg = pm < 0 ? -1 : 1; num += sg * d[k] * (sg * c[k] - s[k]); den += d[k] * d[k]; } }
    double t = tm; if (den > 0) { t = -num / den; if (t < a) t = a; if (t > b) t = b; }
    double f = seg_box_f(c, d, s, t); if (f < bf) { bf = f; bt = t; }
    f = seg_box_f(c, d, s, b); if (f < bf) { bf = f; bt = b; }
  }
  return bt;
}
static double geom_distance(const Kin* k, const bik_geom* g1, const bik_geom* g2, double distmax, double* fromto) {
  const bik_geom* G[2] = {g1, g2}; int swap = 0;
  if (g2->type == BIK_GEOM_PLANE || (g2->type == BIK_GEOM_BOX && g1->type != BIK_GEOM_PLANE)) { G[0] = g2; G[1] = g1; swap = 1; }
  double p[2][3], R[2][9], on1[3], on2[3], dist;
  for (int i = 0; i < 2; ++i) frame_pose(k, &G[i]->frame, p[i], R[i]);
  if (G[0]->type == BIK_GEOM_PLANE && G[1]->type == BIK_GEOM_BOX) {   /* lowest corner */
    double n[3] = {R[0][2], R[0][5], R[0][8]}, corner[3] = {p[1][0], p[1][1], p[1][2]};
    for (int a = 0; a < 3; ++a) {
      double ax[3] = {R[1][a] * G[1]->size[a], R[1][3 + a] * G[1]->size[a], R[1][6 + a] * G[1]->size[a]};
      double sg = ax[0] * n[0] + ax[1] * n[1] + ax[2] * n[2] >= 0 ? 1 : -1;
      for (int i = 0; i < 3; ++i) corner[i] -= sg * ax[i];
    }
    dist = 0; for (int i = 0; i < 3; ++i) dist += (corner[i] - p[0][i]) * n[i];
    for (int i = 0; i < 3; ++i) { on2[i] = corner[i]; on1[i] = corner[i] - n[i] * dist; }
  } else if (G[0]->type == BIK_GEOM_BOX) {   /* box vs sphere / capsule core, in the box frame */
    const double* s = G[0]->size; double c[3], d[3] = {0, 0, 0}, rel[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    for (int a = 0; a < 3; ++a) c[a] = R[0][a] * rel[0] + R[0][3 + a] * rel[1] + R[0][6 + a] * rel[2];
    if (G[1]->type == BIK_GEOM_CAPSULE) {
      double ax[3] = {R[1][2] * G[1]->size[1], R[1][5] * G[1]->size[1], R[1][8] * G[1]->size[1]};
      for (int a = 0; a < 3; ++a) d[a] = R[0][a] * ax[0] + R[0][3 + a] * ax[1] + R[0][6 + a] * ax[2];
    }
    double t = G[1]->type == BIK_GEOM_CAPSULE ? seg_box_param(c, d, s) : 0, pt[3], qc[3], v[3], nl[3] = {0, 0, 0}, Ln = 0, rb = G[1]->size[0];
    for (int a = 0; a < 3; ++a) { pt[a] = c[a] + t * d[a]; qc[a] = pt[a] < -s[a] ? -s[a] : (pt[a] > s[a] ? s[a] : pt[a]); v[a] = pt[a] - qc[a]; Ln += v[a] * v[a]; }
    Ln = sqrt(Ln);
    if (Ln > MJ_MINVAL) { for (int a = 0; a < 3; ++a) nl[a] = v[a] / Ln; }
    else {
      int kb = 0; double gb = s[0] - fabs(pt[0]);
      for (int a = 1; a < 3; ++a) if (s[a] - fabs(pt[a]) < gb) { gb = s[a] - fabs(pt[a]); kb = a; }
      nl[kb] = pt[kb] < 0 ? -1 : 1; Ln = -gb;
      for (int a = 0; a < 3; ++a) qc[a] = pt[a] - Ln * nl[a];
    }
    dist = Ln - rb;
    for (int i = 0; i < 3; ++i) {
      on1[i] = p[0][i]; on2[i] = p[0][i];
      for (int a = 0; a < 3; ++a) { on1[i] += R[0][3 * i + a] * qc[a]; on2[i] += R[0][3 * i + a] * (pt[a] - rb * nl[a]); }
    }
  } else if (G[0]->type == BIK_GEOM_PLANE) {
    double n[3] = {R[0][2], R[0][5], R[0][8]}, ends[2][3]; int ne = 1;
    memcpy(ends[0], p[1], sizeof p[1]);
    if (G[1]->type == BIK_GEOM_CAPSULE) {
      ne = 2; double ax[3] = {R[1][2] * G[1]->size[1], R[1][5] * G[1]->size[1], R[1][8] * G[1]->size[1]};
      for (int i = 0; i < 3; ++i) { ends[0][i] = p[1][i] + ax[i]; ends[1][i] = p[1][i] - ax[i]; }
    }
    int best = 0; double hb = 0;
    for (int e = 0; e < ne; ++e) { double h = 0; for (int i = 0; i < 3; ++i) h += (ends[e][i] - p[0][i]) * n[i]; if (e == 0 || h < hb) { hb = h; best = e; } }
    dist = hb - G[1]->size[0];
    for (int i = 0; i < 3; ++i) { on2[i] = ends[best][i] - n[i] * G[1]->size[0]; on1[i] = ends[best][i] - n[i] * hb; }
  } else {
    double d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0}, a[3], b[3];
    if (G[0]->type == BIK_GEOM_CAPSULE) { d1[0] = R[0][2] * G[0]->size[1]; d1[1] = R[0][5] * G[0]->size[1]; d1[2] = R[0][8] * G[0]->size[1]; }
    if (G[1]->type == BIK_GEOM_CAPSULE) { d2[0] = R[1][2] * G[1]->size[1]; d2[1] = R[1][5] * G[1]->size[1]; d2[2] = R[1][8] * G[1]->size[1]; }
    seg_closest(p[0], d1, p[1], d2, a, b);
    double v[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, Ln = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), nr[3] = {1, 0, 0};
    if (Ln > MJ_MINVAL) for (int i = 0; i < 3; ++i) nr[i] = v[i] / Ln;
    dist = Ln - G[0]->size[0] - G[1]->size[0];
    for (int i = 0; i < 3; ++i) { on1[i] = a[i] + nr[i] * G[0]->size[0]; on2[i] = b[i] - nr[i] * G[1]->size[0]; }
  }
  if (dist >= distmax) { memset(fromto, 0, 6 * sizeof(double)); return distmax; }
  if (swap) { memcpy(fromto, on2, sizeof on2); memcpy(fromto + 3, on1, sizeof on1); }
  else { memcpy(fromto, on1, sizeof on1); memcpy(fromto + 3, on2, sizeof on2); }
  return dist;
}

/* collision rows of one limit (collision_avoidance_limit.py:187-210): G [n][nv], h [n] */
static void collision_rows(const Model* m, const Kin* k, const bik_limit_desc* L, double dt, double* G, double* h, double* scratch) {
  int nv = m->nv;
  for (int r = 0; r < L->n; ++r) {
    const bik_geom *g1 = L->geoms + L->pairs[2 * r], *g2 = L->geoms + L->pairs[2 * r + 1];
    double ft[6], dist = geom_distance(k, g1, g2, L->detection_distance, ft);
    memset(G + (size_t)r * nv, 0, sizeof(double) * nv); h[r] = INFINITY;
    if (dist == L->detection_distance) continue;
    h[r] = dist > L->minimum_distance ? L->gain * (dist - L->minimum_distance) / dt + L->bound_relaxation : L->bound_relaxation;
    double nrm[3] = {ft[3] - ft[0], ft[4] - ft[1], ft[5] - ft[2]}, nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    if (nn < MJ_MINVAL) { nrm[0] = 1; nrm[1] = nrm[2] = 0; } else for (int i = 0; i < 3; ++i) nrm[i] /= nn; /* mju_normalize3 */
    double *j1 = scratch, *j2 = scratch + 3 * nv;
    point_jac(m, k, g1->frame.node, ft, j1, NULL); point_jac(m, k, g2->frame.node, ft + 3, j2, NULL);
    for (int d = 0; d < nv; ++d) { double s = 0; for (int i = 0; i < 3; ++i) s += nrm[i] * (j2[i * nv + d] - j1[i * nv + d]); G[(size_t)r * nv + d] = -s; }
  }
}

/* ------------------------------------------------------------------------- */
/* Goldfarb-Idnani dual active set on  min 1/2 x'Hx + c'x,  lo <= x <= hi,  Gx <= h */
/* ------------------------------------------------------------------------- */
static int chol(double* A, int n) { /* in place lower Cholesky, row major */
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j]; for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
    if (!(s > 0)) return -1;
    double d = sqrt(s); A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) { double t = A[i * n + j]; for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k]; A[i * n + j] = t / d; }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}

/* constraint r in ">= form": a_r^T x >= b_r.  Rows 0..nv-1: x_i <= hi_i ; nv..2nv-1: x_i >= lo_i ; then general rows. */
typedef struct { int n, mg; const double *lo, *hi, *G, *h; } Cons;
static int cons_count(const Cons* c) { return 2 * c->n + c->mg; }
static int cons_valid(const Cons* c, int r) {
  if (r < c->n) return isfinite(c->hi[r]);
  if (r < 2 * c->n) return isfinite(c->lo[r - c->n]);
  return isfinite(c->h[r - 2 * c->n]);
}
static void cons_normal(const Cons* c, int r, double* a) {
  int n = c->n; memset(a, 0, sizeof(double) * n);
  if (r < n) a[r] = -1.0; else if (r < 2 * n) a[r - n] = 1.0;
  else for (int i = 0; i < n; ++i) a[i] = -c->G[(size_t)(r - 2 * n) * n + i];
}
static double cons_slack(const Cons* c, int r, const double* x, double* norm) { /* a^T x - b */
  int n = c->n; *norm = 1.0;
  if (r < n) return c->hi[r] - x[r];
  if (r < 2 * n) return x[r - n] - c->lo[r - n];
  const double* g = c->G + (size_t)(r - 2 * n) * n; double s = c->h[r - 2 * n], nn = 0;
  for (int i = 0; i < n; ++i) { s -= g[i] * x[i]; nn += g[i] * g[i]; }
  *norm = nn > 0 ? sqrt(nn) : 1.0; return s;
}

/* returns 0 ok, 1 infeasible, 2 iteration cap, 3 H not PD.  `nactive` out. */
static int qp_solve(int n, const double* H, const double* c, const Cons* cons, double* x, int* nactive, int* iters_out, double* work) {
  const double tol = 1e-11;
  double* L = work;                 /* n*n */
  double* Hinv = L + n * n;         /* n*n */
  int maxa = n;                     /* at most n linearly independent active constraints */
  double* N = Hinv + n * n;         /* maxa*n : active normals (row per constraint) */
  double* HN = N + (size_t)maxa * n;/* maxa*n : H^-1 a */
  double* M = HN + (size_t)maxa * n;/* maxa*maxa */
  double* np_ = M + (size_t)maxa * maxa, *hnp = np_ + n, *z = hnp + n, *r = z + n, *u = r + n, *rhs = u + n + 1;
  int* act = (int*)(rhs + n + 1);
  memcpy(L, H, sizeof(double) * n * n);
  if (chol(L, n)) return 3;
  for (int j = 0; j < n; ++j) { double* col = Hinv + (size_t)j * n; memset(col, 0, sizeof(double) * n); col[j] = 1.0; chol_solve(L, n, col); } /* symmetric */
  for (int i = 0; i < n; ++i) x[i] = -c[i];
  chol_solve(L, n, x);
  int q = 0, iters = 0, m = cons_count(cons), maxit = 50 * (m + n);
  for (;;) {
    int p = -1; double worst = -tol;
    for (int rr = 0; rr < m; ++rr) {
      if (!cons_valid(cons, rr)) continue;
      int isact = 0; for (int k = 0; k < q; ++k) if (act[k] == rr) { isact = 1; break; }
      if (isact) continue;
      double nn, s = cons_slack(cons, rr, x, &nn) / nn;
      if (s < worst) { worst = s; p = rr; }
    }
    if (p < 0) { *nactive = q; *iters_out = iters; return 0; }
    cons_normal(cons, p, np_);
    for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += Hinv[(size_t)i * n + j] * np_[j]; hnp[i] = s; }
    double uplus = 0;
    for (;;) {
      if (++iters > maxit) { *nactive = q; *iters_out = iters; return 2; }
      /* r = (N H^-1 N^T)^-1 N H^-1 n_p ; z = H^-1 n_p - (H^-1 N^T) r */
      if (q > 0) {
        for (int a = 0; a < q; ++a) {
          for (int b2 = 0; b2 <= a; ++b2) { double s = 0; for (int i = 0; i < n; ++i) s += N[(size_t)a * n + i] * HN[(size_t)b2 * n + i]; M[a * q + b2] = M[b2 * q + a] = s; }
          double s = 0; for (int i = 0; i < n; ++i) s += HN[(size_t)a * n + i] * np_[i]; rhs[a] = s;
        }
        if (chol(M, q)) { /* dependent active set: should not happen (GI keeps N full rank) */ *nactive = q; *iters_out = iters; return 1; }
        memcpy(r, rhs, sizeof(double) * q); chol_solve(M, q, r);
      }
      for (int i = 0; i < n; ++i) { double s = hnp[i]; for (int a = 0; a < q; ++a) s -= HN[(size_t)a * n + i] * r[a]; z[i] = s; }
      double t1 = INFINITY; int kdrop = -1;
      for (int a = 0; a < q; ++a) if (r[a] > 1e-13) { double cand = u[a] / r[a]; if (cand < t1) { t1 = cand; kdrop = a; } }
      double zn = 0, nhn = 0, nn; for (int i = 0; i < n; ++i) { zn += z[i] * np_[i]; nhn += np_[i] * hnp[i]; }
      double sp = cons_slack(cons, p, x, &nn);
      double t2 = (zn > 1e-13 * (nhn > 1.0 ? nhn : 1.0)) ? -sp / zn : INFINITY;
      double t = t1 < t2 ? t1 : t2;
      if (!isfinite(t)) { *nactive = q; *iters_out = iters; return 1; }
      if (isfinite(t2)) for (int i = 0; i < n; ++i) x[i] += t * z[i];
      for (int a = 0; a < q; ++a) u[a] -= t * r[a];
      uplus += t;
      if (isfinite(t2) && t == t2) {
        if (q >= maxa) { *nactive = q; *iters_out = iters; return 1; }
        memcpy(N + (size_t)q * n, np_, sizeof(double) * n); memcpy(HN + (size_t)q * n, hnp, sizeof(double) * n);
        act[q] = p; u[q] = uplus; q++;
        break;
      }
      /* drop constraint kdrop */
      for (int a = kdrop; a + 1 < q; ++a) {
        memcpy(N + (size_t)a * n, N + (size_t)(a + 1) * n, sizeof(double) * n); memcpy(HN + (size_t)a * n, HN + (size_t)(a + 1) * n, sizeof(double) * n);
        act[a] = act[a + 1]; u[a] = u[a + 1];
      }
      q--;
    }
  }
}
static size_t qp_work_doubles(int n) { return (size_t)2 * n * n + (size_t)2 * n * n + (size_t)n * n + 8 * (size_t)n + 16 + (size_t)n; }

/* ------------------------------------------------------------------------- */
/* integrate (mj_integratePos)                                                */
/* ------------------------------------------------------------------------- */
static void quat_integrate(double* q, const double* w, double dt) {
  qnormalize(q);
  double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (n < MJ_MINVAL) return;
  double ang = dt * n, s = sin(0.5 * ang), r[4] = {cos(0.5 * ang), w[0] / n * s, w[1] / n * s, w[2] / n * s}, o[4];
  qmul(q, r, o); qnormalize(o); memcpy(q, o, sizeof o);
}
static void integrate_instance(const Model* m, double* q, const double* v, double dt) {
  for (int n = 0; n < m->nnode; ++n) {
    int t = m->node_type[n], a = m->node_qadr[n], d = m->node_dadr[n];
    if (t == 0) { for (int i = 0; i < 3; ++i) q[a + i] += dt * v[d + i]; quat_integrate(q + a + 3, v + d + 3, dt); }
    else if (t == 1) quat_integrate(q + a, v + d, dt);
    else q[a] += dt * v[d];
  }
}
int iko_integrate(const void* blob, int B, double* q, const double* dq) {
  Model m; if (model_from_blob(blob, &m)) return -1;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) integrate_instance(&m, q + (size_t)b * m.nq, dq + (size_t)b * m.nv, 1.0);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* whole step: solve_ik (solve_ik.py:68-105) [+ integrate_inplace], nsteps times */
/* ------------------------------------------------------------------------- */
int iko_step(const void* blob, const bik_task_desc* tasks, int ntasks, const bik_limit_desc* limits, int nlimits, int B, double* q,
             const double* frame_targets, const double* posture_targets, int posture_batched, const double* com_targets, double dt,
             double damping, int nsteps, int integrate, double* dq, int32_t* status, int32_t* nactive_out, int nthreads) {
  Model m; if (model_from_blob(blob, &m)) return -1;
  Counts cn = count_tasks(tasks, ntasks);
  int nv = m.nv, npairs = 0;
  for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) npairs += limits[l].n;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode);
    double* s = (double*)malloc(sizeof(double) * 18 * nv);
    double* J = (double*)malloc(sizeof(double) * ((size_t)(cn.K + 1) * nv + cn.K + (size_t)(cn.P + 1) * nv));
    double* e = J + (size_t)(cn.K + 1) * nv; double* ep = e + cn.K;
    double* H = (double*)malloc(sizeof(double) * ((size_t)nv * nv + 3 * nv + (size_t)(npairs + 1) * (nv + 1)));
    double *c = H + (size_t)nv * nv, *lo = c + nv, *hi = lo + nv, *G = hi + nv, *h = G + (size_t)(npairs + 1) * nv;
    double* work = (double*)malloc(sizeof(double) * qp_work_doubles(nv) + sizeof(int) * (nv + 4));
    double* x = (double*)malloc(sizeof(double) * nv);
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      double* qb = q + (size_t)b * m.nq; int st = 0, nact = 0, it = 0;
      for (int step = 0; step < nsteps; ++step) {
        fk_nodes(&m, qb, &k);
        tasks_instance(&m, &k, tasks, ntasks, qb, frame_targets ? frame_targets + (size_t)b * cn.F * 7 : NULL,
                       posture_targets ? posture_targets + (posture_batched ? (size_t)b * cn.P * m.nq : 0) : NULL,
                       com_targets ? com_targets + (size_t)b * cn.C * 3 : NULL, J, e, ep, s);
        objective_instance(&m, tasks, ntasks, J, e, ep, damping, H, c);
        box_instance(&m, limits, nlimits, qb, dt, lo, hi);
        int row = 0;
        for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) { collision_rows(&m, &k, limits + l, dt, G + (size_t)row * nv, h + row, s); row += limits[l].n; }
        Cons cons = {nv, npairs, lo, hi, G, h};
        int rc = qp_solve(nv, H, c, &cons, x, &nact, &it, work);
        if (rc == 1) st |= BIK_STATUS_QP_INFEASIBLE; else if (rc == 2) st |= BIK_STATUS_QP_MAXITER; else if (rc == 3) st |= BIK_STATUS_NONFINITE;
        if (integrate) integrate_instance(&m, qb, x, 1.0); /* v*dt = dq */
      }
      memcpy(dq + (size_t)b * nv, x, sizeof(double) * nv);
      if (status) status[b] = st;
      if (nactive_out) nactive_out[b] = nact;
    }
    kin_free(&k); free(s); free(J); free(H); free(work); free(x);
  }
  return 0;
}

/* Converge-until-threshold loop of the reference's examples (examples/arm_iiwa.py:63-70, examples/arm_ur5e_actuators.py:88-97,
 * examples/quadruped_spot.py:89-104), per instance: solve_ik + integrate, then the frame-task errors at the NEW configuration are
 * compared with the thresholds.  iters_out = steps taken (1..max_iters); converged_out = 1 when the thresholds were met. */
int iko_converge(const void* blob, const bik_task_desc* tasks, int ntasks, const bik_limit_desc* limits, int nlimits, int B, double* q,
                 const double* frame_targets, const double* posture_targets, int posture_batched, const double* com_targets, double dt,
                 double damping, int max_iters, double pos_thr, double ori_thr, int32_t* iters_out, int32_t* converged_out, int32_t* status,
                 int nthreads) {
  Model m; if (model_from_blob(blob, &m)) return -1;
  Counts cn = count_tasks(tasks, ntasks);
  int nv = m.nv, npairs = 0;
  for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) npairs += limits[l].n;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  (void)nthreads;
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode);
    double* s = (double*)malloc(sizeof(double) * 18 * nv);
    double* J = (double*)malloc(sizeof(double) * ((size_t)(cn.K + 1) * nv + cn.K + (size_t)(cn.P + 1) * nv));
    double* e = J + (size_t)(cn.K + 1) * nv; double* ep = e + cn.K;
    double* H = (double*)malloc(sizeof(double) * ((size_t)nv * nv + 3 * nv + (size_t)(npairs + 1) * (nv + 1)));
    double *c = H + (size_t)nv * nv, *lo = c + nv, *hi = lo + nv, *G = hi + nv, *h = G + (size_t)(npairs + 1) * nv;
    double* work = (double*)malloc(sizeof(double) * qp_work_doubles(nv) + sizeof(int) * (nv + 4));
    double* x = (double*)malloc(sizeof(double) * nv);
#pragma omp for schedule(dynamic, 16)
    for (int b = 0; b < B; ++b) {
      double* qb = q + (size_t)b * m.nq; int st = 0, nact = 0, it = 0, conv = 0, steps = 0;
      const double* ft = frame_targets ? frame_targets + (size_t)b * cn.F * 7 : NULL;
      const double* pt = posture_targets ? posture_targets + (posture_batched ? (size_t)b * cn.P * m.nq : 0) : NULL;
      const double* ct = com_targets ? com_targets + (size_t)b * cn.C * 3 : NULL;
      fk_nodes(&m, qb, &k);
      tasks_instance(&m, &k, tasks, ntasks, qb, ft, pt, ct, J, e, ep, s);
      for (int step = 0; step < max_iters; ++step) {
        objective_instance(&m, tasks, ntasks, J, e, ep, damping, H, c);
        box_instance(&m, limits, nlimits, qb, dt, lo, hi);
        int row = 0;
        for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) { collision_rows(&m, &k, limits + l, dt, G + (size_t)row * nv, h + row, s); row += limits[l].n; }
        Cons cons = {nv, npairs, lo, hi, G, h};
        int rc = qp_solve(nv, H, c, &cons, x, &nact, &it, work);
        if (rc == 1) st |= BIK_STATUS_QP_INFEASIBLE; else if (rc == 2) st |= BIK_STATUS_QP_MAXITER; else if (rc == 3) st |= BIK_STATUS_NONFINITE;
        integrate_instance(&m, qb, x, 1.0);
        steps = step + 1;
        fk_nodes(&m, qb, &k);
        tasks_instance(&m, &k, tasks, ntasks, qb, ft, pt, ct, J, e, ep, s);
        int ok = 1, erow = 0;
        for (int t = 0; t < ntasks; ++t) {
          if (tasks[t].kind == BIK_TASK_FRAME || tasks[t].kind == BIK_TASK_RELATIVE_FRAME) {
            const double* eb = e + erow;
            double pos = sqrt(eb[0] * eb[0] + eb[1] * eb[1] + eb[2] * eb[2]), ori = sqrt(eb[3] * eb[3] + eb[4] * eb[4] + eb[5] * eb[5]);
            if (!(pos <= pos_thr && ori <= ori_thr)) ok = 0;
            erow += 6;
          } else if (tasks[t].kind == BIK_TASK_COM) erow += 3;
        }
        if (ok) { conv = 1; break; }
      }
      if (iters_out) iters_out[b] = steps;
      if (converged_out) converged_out[b] = conv;
      if (status) status[b] = st;
    }
    kin_free(&k); free(s); free(J); free(H); free(work); free(x);
  }
  return 0;
}

/* collision rows for a batch (checker for bik_fk_jac's G_coll / h_coll) */
int iko_collision(const void* blob, const bik_limit_desc* limits, int nlimits, int B, const double* q, double dt, double* G, double* h) {
  Model m; if (model_from_blob(blob, &m)) return -1;
  int npairs = 0; for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) npairs += limits[l].n;
#pragma omp parallel
  {
    Kin k = kin_alloc(m.nnode); double* s = (double*)malloc(sizeof(double) * 18 * m.nv);
#pragma omp for
    for (int b = 0; b < B; ++b) {
      fk_nodes(&m, q + (size_t)b * m.nq, &k); int row = 0;
      for (int l = 0; l < nlimits; ++l) if (limits[l].kind == BIK_LIMIT_COLLISION) {
        collision_rows(&m, &k, limits + l, dt, G + ((size_t)b * npairs + row) * m.nv, h + (size_t)b * npairs + row, s); row += limits[l].n; }
    }
    kin_free(&k); free(s);
  }
  return 0;
}

int iko_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
