// bik_k1.h -- K1: forward kinematics + task errors + task Jacobians (+ collision rows).
//
// Mapping: G lanes cooperate on one robot instance (G in {1,2,4,8,16,32}); a warp of W lanes
// therefore carries IPW = W/G consecutive instances.  The tree is walked by a host-compiled lane
// program (bik_build.h: lane_program): at step s lane g composes node prog[s][g] from its parent's
// pose, which an earlier step left in the per-instance shared-memory state.  Jacobian columns are
// written into a per-warp staging tile and flushed as contiguous runs, so that global stores are
// coalesced no matter how the lanes were assigned.
//
// The same source is compiled for the host with W = G = 1 (tests/host_emu) -- every cross-lane
// primitive degenerates to a no-op there.
//
// Reference semantics: mujoco mj_kinematics / mj_jac via mink/configuration.py:63-64,144-153;
// FrameTask mink/tasks/frame_task.py:95-146; ComTask mink/tasks/com_task.py:71-97;
// PostureTask.compute_error mink/tasks/posture_task.py:87-118;
// CollisionAvoidanceLimit mink/limits/collision_avoidance_limit.py:59-72,187-210.
#pragma once
#include "bik_layout.h"
#include "bik_math.h"

#if defined(__CUDA_ARCH__)
#define BIK_SYNCWARP() __syncwarp()
#else
#define BIK_SYNCWARP() ((void)0)
#endif

#define BIK_INF_F (bik::bik_inf())

namespace bik {

// plain 8/16-byte carriers for vectorised shared/global moves (float2/float4 only exist under nvcc)
struct alignas(8) F2 { float x, y; };
struct alignas(16) F4 { float x, y, z, w; };

BIK_HD float bik_inf() {
#if defined(__CUDA_ARCH__)
  return __int_as_float(0x7f800000);
#else
  return INFINITY;
#endif
}

typedef V3<float> F3;
typedef Q4<float> FQ;
typedef M3<float> FM;

struct PView {
  const uint32_t* base;
  BIK_HD const PHeader& h() const { return *reinterpret_cast<const PHeader*>(base); }
  BIK_HD const NodeRec& node(int n) const { return *reinterpret_cast<const NodeRec*>(base + h().off_nodes + NODE_WORDS * n); }
  BIK_HD const FrameRec& frame(int f) const { return *reinterpret_cast<const FrameRec*>(base + h().off_frames + FRAME_WORDS * f); }
  BIK_HD const ComNodeRec& comnode(int n) const { return *reinterpret_cast<const ComNodeRec*>(base + h().off_comnodes + COMNODE_WORDS * n); }
  BIK_HD const GeomRec& geom(int g) const { return *reinterpret_cast<const GeomRec*>(base + h().off_geoms + GEOM_WORDS * g); }
  BIK_HD const float* f(int off) const { return reinterpret_cast<const float*>(base + off); }
  BIK_HD const int32_t* i(int off) const { return reinterpret_cast<const int32_t*>(base + off); }
};

struct K1Args {
  int B;
  const float* q;        // [B][nq]
  const float* ftgt;     // [B][F][7]
  const float* ptgt;     // [B or 1][P][nq]
  const float* ctgt;     // [B][C][3]
  int pbatched;
  float dt;
  float* J;              // [B][K][nv]
  float* e;              // [B][K]
  float* ep;             // [B][P][nv]
  float* Gc;             // [B][npairs][nv]
  float* hc;             // [B][npairs]
};

// ---- per-warp scratch layout (floats) ------------------------------------------------------
BIK_HD int k1_state_stride(const PHeader& h) {  // pose (7) + CoM first moment (3) per node, odd stride
  int s = 7 * h.nslots + (h.C > 0 ? 3 * h.nnode : 0);
  return s | 1;
}
BIK_HD int k1_stage_rows(const PHeader& h) { return 6; }
BIK_HD int k1_state_words(const PHeader& h, int ipw) { return (ipw * k1_state_stride(h) + 3) & ~3; }  // keeps the stage 16-byte aligned
BIK_HD int k1_fsc_stride(const PHeader& h) { return h.nrel > 0 ? 32 : 24; }
// Frame-task rows go straight to global memory when J holds nothing else (k1_frames_direct); the staging tile (6 rows
// per instance) is only needed for CoM rows, collision rows, and frame rows that share J with CoM rows.
BIK_HD bool k1_frames_direct(const PHeader& h) {
#ifdef BIK_K1_STAGED   // A/B switch: every row through the staging tile, as before
  return false;
#else
  return h.C == 0;
#endif
}
BIK_HD int k1_stage_words(const PHeader& h, int ipw) {
  if (k1_frames_direct(h) && h.npairs == 0) return 0;
  return (ipw * 6 * h.nv + 3) & ~3;
}
BIK_HD int k1_warp_words(const PHeader& h, int ipw) {
  int w = k1_state_words(h, ipw) + k1_stage_words(h, ipw) + ipw * (h.K > 0 ? h.K : 1) + ipw * (h.F > 0 ? h.F : 1) * k1_fsc_stride(h) +
          ((ipw * h.nq + 3) & ~3);
  return (w + 3) & ~3;
}

BIK_HD FQ ld_q(const float* p) { return q4<float>(p[0], p[1], p[2], p[3]); }
BIK_HD F3 ld_v(const float* p) { return v3<float>(p[0], p[1], p[2]); }

// One node of the tree: pose of the node frame after its joint (mj_kinematics, one joint per node).
BIK_HD void fk_node(const PView& P, int n, const float* q, float* xs) {
  const NodeRec& r = P.node(n);
  FQ quat;
  F3 pos;
  if (r.type == JNT_FREE) {
    pos = ld_v(q + r.qadr);
    quat = ld_q(q + r.qadr + 3);
  } else {
    if (r.parent >= 0) {
      FQ pq = ld_q(xs + 7 * r.pslot);
      pos = ld_v(xs + 7 * r.pslot + 4) + qrot(pq, ld_v(r.pos));
      quat = qmul(pq, ld_q(r.quat));
    } else {
      pos = ld_v(r.pos);
      quat = ld_q(r.quat);
    }
    const float* q0 = P.f(P.h().off_qpos0);
    if (r.type == JNT_SLIDE) {
      pos = pos + (q[r.qadr] - q0[r.qadr]) * qrot(quat, ld_v(r.axis));
    } else {
      F3 jp = ld_v(r.jpos);
      bool off_centre = (jp.x != 0.f) || (jp.y != 0.f) || (jp.z != 0.f);
      F3 anchor = pos;
      if (off_centre) anchor = pos + qrot(quat, jp);
      FQ ql;
      if (r.type == JNT_HINGE) {
        float s, c;
        bik_sincos<float>(0.5f * (q[r.qadr] - q0[r.qadr]), &s, &c);
        ql = q4<float>(c, s * r.axis[0], s * r.axis[1], s * r.axis[2]);
      } else {
        ql = qnormalize(ld_q(q + r.qadr));
      }
      quat = qmul(quat, ql);
      if (off_centre) pos = anchor - qrot(quat, jp);
    }
  }
  // mj_kinematics renormalises every body quaternion.  A product of unit quaternions leaves the unit sphere by ~1e-7
  // per level in fp32 (1e-6 at the deepest node of the BASELINE robots, against a 1e-4 parity budget), so only
  // quaternions that come straight from q (free / ball joints: callers may pass them unnormalised) are renormalised.
#ifdef BIK_K1_NORMALIZE_ALL
  quat = qnormalize(quat);
#else
  if (r.type == JNT_FREE) quat = qnormalize(quat);
#endif
  float* o = xs + 7 * r.slot;
  o[0] = quat.w; o[1] = quat.x; o[2] = quat.y; o[3] = quat.z; o[4] = pos.x; o[5] = pos.y; o[6] = pos.z;
}

// `node` is the STATE ROW of the frame's node (FrameRec::slot; the node id itself where slots are the identity), -1 = world
BIK_HD void frame_pose(int node, const float* lpos, const float* lquat, const float* xs, FQ* qf, F3* pf) {
  if (node < 0) { *qf = ld_q(lquat); *pf = ld_v(lpos); return; }
  FQ nq = ld_q(xs + 7 * node);
  *qf = qnormalize(qmul(nq, ld_q(lquat)));
  *pf = ld_v(xs + 7 * node + 4) + qrot(nq, ld_v(lpos));
}

// World-aligned point-Jacobian column of dof `d` (owned by node n) for a point p moving with a
// descendant of n: jp = linear part, jr = angular part (mj_jac restated per column).
BIK_HD void jac_column(const PView& P, int d, int n, const float* xs, F3 p, F3* jp, F3* jr) {
  const NodeRec& r = P.node(n);
  FQ nq = ld_q(xs + 7 * r.slot);
  F3 np = ld_v(xs + 7 * r.slot + 4);
  int k = d - r.dadr;
  if (r.type == JNT_HINGE) {
    F3 ax = qrot(nq, ld_v(r.axis));
    F3 anchor = np + qrot(nq, ld_v(r.jpos));
    *jr = ax; *jp = cross(ax, p - anchor);
  } else if (r.type == JNT_SLIDE) {
    *jr = v3<float>(0.f, 0.f, 0.f); *jp = qrot(nq, ld_v(r.axis));
  } else if (r.type == JNT_FREE && k < 3) {
    *jr = v3<float>(0.f, 0.f, 0.f); *jp = v3<float>(k == 0 ? 1.f : 0.f, k == 1 ? 1.f : 0.f, k == 2 ? 1.f : 0.f);
  } else {  // rotational dof of a free or ball joint: body-frame axis k
    int c = (r.type == JNT_FREE) ? k - 3 : k;
    F3 ax = qrot(nq, v3<float>(c == 0 ? 1.f : 0.f, c == 1 ? 1.f : 0.f, c == 2 ? 1.f : 0.f));
    F3 anchor = (r.type == JNT_FREE) ? np : np + qrot(nq, ld_v(r.jpos));
    *jr = ax; *jp = cross(ax, p - anchor);
  }
}

// FrameTask error and the two 3x3 blocks that map world-aligned (jp, jr) columns to task-Jacobian
// columns:  J[:,d] = [A1 jp + A2 jr ; A1 jr]   with A1 = -Ji R^T, A2 = -Mi R^T.
BIK_HD void frame_task(FQ qf, F3 pf, const float* tgt, F3* ev, F3* ew, FM* A1, FM* A2) {
  FQ tq = qnormalize(ld_q(tgt));
  F3 tp = ld_v(tgt + 4);
  // e = log(T_wb^-1 T_wt)                                (frame_task.py:119-122)
  se3_log<float>(qmul(qconj(qf), tq), qrot_inv(qf, tp - pf), ev, ew);
  // jlog(T_wt^-1 T_wb) = ljacinv(-log(T_tb))            (frame_task.py:145-146, lie/base.py:151-156)
  // and log(T_tb) = log(T_bt^-1) = -e, so the argument of ljacinv is e itself: no second logarithm.
  FM Ji, Mi;
  se3_ljacinv_blocks<float>(*ev, *ew, &Ji, &Mi);
  FM Rt = mtrans(q2mat(qf));
  *A1 = mmul(Ji, Rt);
  *A2 = mmul(Mi, Rt);
  for (int i = 0; i < 9; ++i) { A1->m[i] = -A1->m[i]; A2->m[i] = -A2->m[i]; }
}

// RelativeFrameTask (reference mink/tasks/relative_frame_task.py:106-142): pose of frame f in root r,
//   e = log(T_rt^-1 T_rf),   J = jlog(T_tf) (J_f - Ad(T_rf^-1) J_r),   T_tf = T_rt^-1 T_rf.
// Returns e and the blocks (Ji, Mi) of jlog(T_tf) = ljacinv(-log T_tf).
BIK_HD void relative_frame_task(FQ qf, F3 pf, FQ qr, F3 pr, const float* tgt, F3* ev, F3* ew, FM* Ji, FM* Mi) {
  FQ q_rf = qmul(qconj(qr), qf);
  F3 t_rf = qrot_inv(qr, pf - pr);
  FQ tq = qnormalize(ld_q(tgt));
  F3 tp = ld_v(tgt + 4);
  se3_log<float>(qmul(qconj(tq), q_rf), qrot_inv(tq, t_rf - tp), ev, ew);
  se3_ljacinv_blocks<float>(v3<float>(-ev->x, -ev->y, -ev->z), v3<float>(-ew->x, -ew->y, -ew->z), Ji, Mi);
}

// ---- primitive geom distance (plane / sphere / capsule) ------------------------------------
BIK_HD void seg_closest(F3 p1, F3 d1, F3 p2, F3 d2, F3* a, F3* b) {
  F3 r = p1 - p2;
  float A = dot(d1, d1), E = dot(d2, d2), Bq = dot(d1, d2), C = dot(d1, r), F = dot(d2, r);
  float den = A * E - Bq * Bq, s = 0.f, t = 0.f;
  if (den > 1e-12f) s = fminf(fmaxf((Bq * F - C * E) / den, -1.f), 1.f);
  if (E > 1e-12f) t = (Bq * s + F) / E;
  if (t < -1.f || t > 1.f) {
    t = t < -1.f ? -1.f : 1.f;
    s = (A > 1e-12f) ? fminf(fmaxf((Bq * t - C) / A, -1.f), 1.f) : 0.f;
  }
  *a = p1 + s * d1; *b = p2 + t * d2;
}
BIK_HD float geom_distance(const GeomRec& g1, const GeomRec& g2, const float* xs, float distmax, F3* on1, F3* on2) {
  const GeomRec* A = &g1; const GeomRec* Bg = &g2;
  bool swap = false;
  if (g2.type == 0) { A = &g2; Bg = &g1; swap = true; }
  FQ qa, qb; F3 pa, pb;
  frame_pose(A->node, A->lpos, A->lquat, xs, &qa, &pa);
  frame_pose(Bg->node, Bg->lpos, Bg->lquat, xs, &qb, &pb);
  F3 oa, ob; float dist;
  F3 zb = qrot(qb, v3<float>(0.f, 0.f, 1.f));
  if (A->type == 0) {  // plane vs sphere/capsule
    F3 n = qrot(qa, v3<float>(0.f, 0.f, 1.f));
    F3 end = pb;
    if (Bg->type == 3) {
      F3 e1 = pb + Bg->size[1] * zb, e2 = pb - Bg->size[1] * zb;
      end = dot(e1 - pa, n) <= dot(e2 - pa, n) ? e1 : e2;
    }
    float hgt = dot(end - pa, n);
    dist = hgt - Bg->size[0];
    ob = end - Bg->size[0] * n; oa = end - hgt * n;
  } else {
    F3 d1 = v3<float>(0.f, 0.f, 0.f), d2 = d1, a, b;
    if (A->type == 3) d1 = A->size[1] * qrot(qa, v3<float>(0.f, 0.f, 1.f));
    if (Bg->type == 3) d2 = Bg->size[1] * zb;
    seg_closest(pa, d1, pb, d2, &a, &b);
    F3 v = b - a;
    float L = sqrtf(dot(v, v));
    F3 nr = L > 1e-15f ? (1.f / L) * v : v3<float>(1.f, 0.f, 0.f);
    dist = L - A->size[0] - Bg->size[0];
    oa = a + A->size[0] * nr; ob = b - Bg->size[0] * nr;
  }
  if (dist >= distmax) return distmax;
  if (swap) { *on1 = ob; *on2 = oa; } else { *on1 = oa; *on2 = ob; }
  return dist;
}

// ---- staging flush: IPW x (R*nv) floats -> global rows [inst][row0 .. row0+R) ----------------
template <int W>
BIK_HD void flush_rows(const float* stage, int chunk, int nvalid, float* out, long long inst_stride, int lane) {
  const bool vec2 = ((chunk | (int)(inst_stride & 1)) & 1) == 0 && ((reinterpret_cast<size_t>(out) | reinterpret_cast<size_t>(stage)) & 7) == 0;
  if (vec2) {  // 8-byte stores: every run starts on an 8-byte boundary
    const int c2 = chunk >> 1;
    for (int li = 0; li < nvalid; ++li) {
      F2* o = reinterpret_cast<F2*>(out + li * inst_stride);
      const F2* s = reinterpret_cast<const F2*>(stage + li * chunk);
      for (int k = lane; k < c2; k += W) o[k] = s[k];
    }
    return;
  }
  for (int li = 0; li < nvalid; ++li) {
    float* o = out + li * inst_stride;
    const float* s = stage + li * chunk;
    for (int k = lane; k < chunk; k += W) o[k] = s[k];
  }
}
template <int W>
BIK_HD void zero_words(float* p, int count, int lane) {  // p is 16-byte aligned, count a multiple of 4 or handled by the tail
  F4* p4 = reinterpret_cast<F4*>(p);
  const int c4 = count >> 2;
  const F4 z = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < c4; k += W) p4[k] = z;
  for (int k = (c4 << 2) + lane; k < count; k += W) p[k] = 0.f;
}

// One warp tile: instances [inst0, inst0 + IPW) clipped to B.
template <int G, int W>
BIK_HD void k1_warp_tile(const PView& P, const K1Args& a, int inst0, float* wsm, int lane) {
  const PHeader& h = P.h();
  constexpr int IPW = W / G;
  const int nv = h.nv, nq = h.nq, K = h.K;
  const int li = lane / G, g = lane % G;
  const int b = inst0 + li;
  const int nvalid = (a.B - inst0) < IPW ? (a.B - inst0) : IPW;
  const bool valid = li < nvalid;
  const int SS = k1_state_stride(h);
  float* state = wsm;
  float* stage = wsm + k1_state_words(h, IPW);
  float* estage = stage + k1_stage_words(h, IPW);
  float* fsc = estage + IPW * (K > 0 ? K : 1);  // per (instance, frame): per-frame SE(3) results
  const int FS = k1_fsc_stride(h);
  float* qtile = fsc + IPW * (h.F > 0 ? h.F : 1) * FS;  // q of this tile: one contiguous, coalesced run of global memory
  float* xs = state + li * SS;
  {
    const float* gq = a.q + (long long)inst0 * nq;
    for (int k = lane; k < nvalid * nq; k += W) qtile[k] = gq[k];
  }
  BIK_SYNCWARP();
  const float* qb = qtile + (valid ? li : 0) * nq;

  // ---- forward kinematics over the lane program --------------------------------------------
  const int32_t* prog = P.i(h.off_prog);
  for (int s = 0; s < h.nsteps; ++s) {
    int n = prog[s * G + g];
    if (valid && n >= 0) fk_node(P, n, qb, xs);
    BIK_SYNCWARP();
  }

  // ---- frame tasks -------------------------------------------------------------------------
  // (1) the G lanes of an instance evaluate different frames' SE(3) algebra (error, jlog blocks) in parallel
  const int32_t* cols = P.i(h.off_cols);
  if (valid) {
    for (int f = g; f < h.F; f += G) {
      const FrameRec& fr = P.frame(f);
      FQ qf; F3 pf, ev, ew; FM A1, A2;
      frame_pose(fr.slot, fr.lpos, fr.lquat, xs, &qf, &pf);
      float* sc = fsc + (li * h.F + f) * FS;
      if (!fr.relative) {
        frame_task(qf, pf, a.ftgt + ((long long)b * h.F + f) * 7, &ev, &ew, &A1, &A2);
        sc[0] = pf.x; sc[1] = pf.y; sc[2] = pf.z;
      } else {   // A1 = Ji, A2 = Mi of jlog(T_tf); plus both frames' poses for the column phase
        FQ qr; F3 pr;
        frame_pose(fr.rslot, fr.rlpos, fr.rlquat, xs, &qr, &pr);
        relative_frame_task(qf, pf, qr, pr, a.ftgt + ((long long)b * h.F + f) * 7, &ev, &ew, &A1, &A2);
        sc[0] = pf.x; sc[1] = pf.y; sc[2] = pf.z;
        sc[21] = qf.w; sc[22] = qf.x; sc[23] = qf.y; sc[24] = qf.z;
        sc[25] = qr.w; sc[26] = qr.x; sc[27] = qr.y; sc[28] = qr.z;
        sc[29] = pr.x; sc[30] = pr.y; sc[31] = pr.z;
      }
      float* eo = estage + li * K + fr.row0;
      eo[0] = ev.x; eo[1] = ev.y; eo[2] = ev.z; eo[3] = ew.x; eo[4] = ew.y; eo[5] = ew.z;
      for (int k = 0; k < 9; ++k) { sc[3 + k] = A1.m[k]; sc[12 + k] = A2.m[k]; }
    }
  }
  BIK_SYNCWARP();
  // (2) per frame: Jacobian columns into the staging tile, then one coalesced flush.
  // (A whole-tile variant -- all K rows staged, one cp.async.bulk store per tile, no per-tile zero-fill --
  //  was measured slower: 0.198 vs 0.153 ms at G = 8, the 3x larger tile halves the resident warps;
  //  profiles/r1_kernels.md.)
  // Direct variant (J holds frame rows only): the tile's K x nv blocks are one contiguous run of global memory; it is
  // zero-filled with 16-byte stores straight from registers, then each lane scatters the columns it computes (a frame
  // touches 12 of G1's 43 columns).  Zero-fill and scatter are ordered by the warp barrier; the partial sectors merge in
  // L2.  No staging tile: 1 032 of the 1 932 shared-memory words a warp needed for G1, and the flush loop (12 % of the
  // kernel's instructions) are gone.
  const bool direct = k1_frames_direct(h);
  if (direct && h.F > 0) {
    float* Jt = a.J + (long long)inst0 * K * nv;
    const int total = nvalid * K * nv;
    if ((reinterpret_cast<size_t>(Jt) & 15) == 0) zero_words<W>(Jt, total, lane);
    else for (int k = lane; k < total; k += W) Jt[k] = 0.f;
    BIK_SYNCWARP();
  }
  for (int f = 0; f < h.F; ++f) {
    const FrameRec& fr = P.frame(f);
    if (!direct) {
      zero_words<W>(stage, IPW * 6 * nv, lane);
      BIK_SYNCWARP();
    }
    if (valid) {
      const float* sc = fsc + (li * h.F + f) * FS;
      F3 pf = ld_v(sc);
      FM A1, A2;
      for (int k = 0; k < 9; ++k) { A1.m[k] = sc[3 + k]; A2.m[k] = sc[12 + k]; }
      float* st = direct ? a.J + ((long long)b * K + fr.row0) * nv : stage + li * 6 * nv;
      if (!fr.relative) {
        for (int c = g; c < fr.ncols; c += G) {
          int ent = cols[fr.col_off + c], d = ent & 0xffff, n = (ent >> 16) & 0x7fff;
          F3 jp, jr;
          jac_column(P, d, n, xs, pf, &jp, &jr);
          F3 top = mmul(A1, jp) + mmul(A2, jr), bot = mmul(A1, jr);
          st[d] = top.x; st[nv + d] = top.y; st[2 * nv + d] = top.z;
          st[3 * nv + d] = bot.x; st[4 * nv + d] = bot.y; st[5 * nv + d] = bot.z;
        }
      } else {
        FQ qf = ld_q(sc + 21), qr = ld_q(sc + 25);
        F3 pr = ld_v(sc + 29);
        FQ q_fr = qmul(qconj(qf), qr);          // rotation of T_rf^-1 = T_fr
        F3 t_fr = qrot_inv(qf, pr - pf);        // translation of T_fr
        for (int c = g; c < fr.ncols; c += G) {
          int ent = cols[fr.col_off + c], d = ent & 0xffff, n = (ent >> 16) & 0x7fff;
          bool root_side = ent < 0;
          F3 jp, jr, tv, tw;
          if (!root_side) {                       // column of J_f (body frame of f)
            jac_column(P, d, n, xs, pf, &jp, &jr);
            tv = qrot_inv(qf, jp); tw = qrot_inv(qf, jr);
          } else {                                // minus Ad(T_fr) applied to the column of J_r
            jac_column(P, d, n, xs, pr, &jp, &jr);
            F3 bv = qrot(q_fr, qrot_inv(qr, jp)), bw = qrot(q_fr, qrot_inv(qr, jr));
            F3 x = cross(t_fr, bw);
            tv = v3<float>(-(bv.x + x.x), -(bv.y + x.y), -(bv.z + x.z));
            tw = v3<float>(-bw.x, -bw.y, -bw.z);
          }
          F3 top = mmul(A1, tv) + mmul(A2, tw), bot = mmul(A1, tw);
          st[d] = top.x; st[nv + d] = top.y; st[2 * nv + d] = top.z;
          st[3 * nv + d] = bot.x; st[4 * nv + d] = bot.y; st[5 * nv + d] = bot.z;
        }
      }
    }
    if (!direct) {
      BIK_SYNCWARP();
      flush_rows<W>(stage, 6 * nv, nvalid, a.J + ((long long)inst0 * K + fr.row0) * nv, (long long)K * nv, lane);
      BIK_SYNCWARP();
    }
  }

  // ---- centre-of-mass tasks (mj_comPos + mj_jacSubtreeCom for body 1) -------------------------
  if (h.C > 0) {
    float* S = xs + 7 * h.nslots;  // first moments per node (a CoM task keeps slots == node ids)
    if (valid && g == 0) {
      for (int n = 0; n < h.nnode; ++n) {
        const ComNodeRec& cn = P.comnode(n);
        F3 m = v3<float>(0.f, 0.f, 0.f);   // nodes without mass below them are not visited by the lane program
        if (cn.sub_m > 0.f) m = cn.own_m * ld_v(xs + 7 * n + 4) + qrot(ld_q(xs + 7 * n), ld_v(cn.own_c));
        S[3 * n] = m.x; S[3 * n + 1] = m.y; S[3 * n + 2] = m.z;
      }
      for (int n = h.nnode - 1; n >= 0; --n) {
        int p = P.node(n).parent;
        if (p >= 0) { S[3 * p] += S[3 * n]; S[3 * p + 1] += S[3 * n + 1]; S[3 * p + 2] += S[3 * n + 2]; }
      }
    }
    BIK_SYNCWARP();
    const float invM = 1.f / h.com_total_mass;
    for (int c = 0; c < h.C; ++c) {
      const float* cr = P.f(h.off_com) + 8 * c;
      const int row0 = reinterpret_cast<const int32_t*>(cr)[5];
      zero_words<W>(stage, IPW * 3 * nv, lane);
      BIK_SYNCWARP();
      if (valid) {
        if (g == 0) {
          F3 tot = ld_v(h.com_fixed);
          for (int n = 0; n < h.nnode; ++n) if (P.node(n).parent < 0) tot = tot + ld_v(S + 3 * n);
          const float* tg = a.ctgt + ((long long)b * h.C + c) * 3;
          float* eo = estage + li * K + row0;  // e = com - target (com_task.py:82)
          eo[0] = tot.x * invM - tg[0]; eo[1] = tot.y * invM - tg[1]; eo[2] = tot.z * invM - tg[2];
        }
        float* st = stage + li * 3 * nv;
        for (int ci = g; ci < h.com_ncols; ci += G) {
          int ent = cols[h.com_cols_off + ci], d = ent & 0xffff, n = (ent >> 16) & 0x7fff;
          const NodeRec& r = P.node(n);
          float Ms = P.comnode(n).sub_m;
          FQ nq_ = ld_q(xs + 7 * n); F3 np = ld_v(xs + 7 * n + 4), Sn = ld_v(S + 3 * n), col;
          int k = d - r.dadr;
          if (r.type == JNT_SLIDE) col = (Ms * invM) * qrot(nq_, ld_v(r.axis));
          else if (r.type == JNT_FREE && k < 3) col = v3<float>(k == 0 ? Ms * invM : 0.f, k == 1 ? Ms * invM : 0.f, k == 2 ? Ms * invM : 0.f);
          else {
            F3 ax, anchor;
            if (r.type == JNT_HINGE) { ax = qrot(nq_, ld_v(r.axis)); anchor = np + qrot(nq_, ld_v(r.jpos)); }
            else {
              int cc = (r.type == JNT_FREE) ? k - 3 : k;
              ax = qrot(nq_, v3<float>(cc == 0 ? 1.f : 0.f, cc == 1 ? 1.f : 0.f, cc == 2 ? 1.f : 0.f));
              anchor = (r.type == JNT_FREE) ? np : np + qrot(nq_, ld_v(r.jpos));
            }
            col = invM * cross(ax, Sn - Ms * anchor);
          }
          st[d] = col.x; st[nv + d] = col.y; st[2 * nv + d] = col.z;
        }
      }
      BIK_SYNCWARP();
      flush_rows<W>(stage, 3 * nv, nvalid, a.J + ((long long)inst0 * K + row0) * nv, (long long)K * nv, lane);
      BIK_SYNCWARP();
    }
  }

  // ---- task errors out (contiguous IPW x K run) ------------------------------------------------
  if (K > 0)
    for (int k = lane; k < nvalid * K; k += W) a.e[(long long)inst0 * K + k] = estage[k];

  // ---- posture errors: e = q* (-) q, free-joint dofs zeroed (posture_task.py:107-118) ------------
  if (h.P > 0) {
    const int32_t* dofnode = P.i(h.off_dofnode);
    const int32_t* dofqadr = P.i(h.off_dofqadr);
    for (int l2 = 0; l2 < nvalid; ++l2) {
      const float* qq = qtile + l2 * nq;
      for (int p = 0; p < h.P; ++p) {
        const float* tg = a.ptgt + ((long long)(a.pbatched ? (inst0 + l2) : 0) * h.P + p) * nq;
        float* o = a.ep + ((long long)(inst0 + l2) * h.P + p) * nv;
        for (int d = lane; d < nv; d += W) {
          float val = 0.f;
          int qa = dofqadr[d];
          if (qa >= 0) val = tg[qa] - qq[qa];
          else {
            const NodeRec& r = P.node(dofnode[d]);
            if (r.type == JNT_BALL) {
              F3 w = quat_sub<float>(qnormalize(ld_q(tg + r.qadr)), qnormalize(ld_q(qq + r.qadr)));
              int c = d - r.dadr;
              val = c == 0 ? w.x : (c == 1 ? w.y : w.z);
            }
          }
          o[d] = val;
        }
      }
    }
  }

  // ---- collision rows (collision_avoidance_limit.py:187-210), 6 pairs per staging pass -----------
  if (h.npairs > 0) {
    const int32_t* pairs = P.i(h.off_pairs);
    for (int p0 = 0; p0 < h.npairs; p0 += 6) {
      int np_ = h.npairs - p0 < 6 ? h.npairs - p0 : 6;
      zero_words<W>(stage, IPW * np_ * nv, lane);
      BIK_SYNCWARP();
      if (valid) {
        for (int pi = g; pi < np_; pi += G) {
          const GeomRec& g1 = P.geom(pairs[2 * (p0 + pi)]);
          const GeomRec& g2 = P.geom(pairs[2 * (p0 + pi) + 1]);
          F3 o1, o2;
          float dist = geom_distance(g1, g2, xs, h.coll_ddet, &o1, &o2);
          float hval = BIK_INF_F;
          if (dist != h.coll_ddet) {
            hval = dist > h.coll_dmin ? h.coll_gain * (dist - h.coll_dmin) / a.dt + h.coll_relax : h.coll_relax;
            F3 nr = o2 - o1;
            float L = sqrtf(dot(nr, nr));
            nr = L > 1e-15f ? (1.f / L) * nr : v3<float>(1.f, 0.f, 0.f);
            float* st = stage + (li * np_ + pi) * nv;
            for (int side = 0; side < 2; ++side) {  // row = -n.(Jp2 - Jp1)
              const GeomRec& gg = side ? g2 : g1;
              F3 pt = side ? o2 : o1;
              float sgn = side ? -1.f : 1.f;
              for (int n = gg.node; n >= 0; n = P.node(n).parent) {
                const NodeRec& r = P.node(n);
                int nd = r.type == JNT_FREE ? 6 : (r.type == JNT_BALL ? 3 : 1);
                for (int k = 0; k < nd; ++k) {
                  F3 jp, jr;
                  jac_column(P, r.dadr + k, n, xs, pt, &jp, &jr);
                  st[r.dadr + k] += sgn * dot(nr, jp);
                }
              }
            }
          }
          a.hc[(long long)b * h.npairs + p0 + pi] = hval;
        }
      }
      BIK_SYNCWARP();
      flush_rows<W>(stage, np_ * nv, nvalid, a.Gc + ((long long)inst0 * h.npairs + p0) * nv, (long long)h.npairs * nv, lane);
      BIK_SYNCWARP();
    }
  }
}

// ---- FK of arbitrary frames / body-frame Jacobians (Configuration API) -------------------------
struct FkArgs {
  int B, nframes;
  const float* q;
  float* poses;   // [B][nframes][7] or null
  float* com;     // [B][3] or null
  float* J;       // [B][nframes][6][nv] or null
  FrameRec frames[16];
};

template <int G, int W>
BIK_HD void fk_warp_tile(const PView& P, const FkArgs& a, int inst0, float* wsm, int lane) {
  const PHeader& h = P.h();
  constexpr int IPW = W / G;
  const int li = lane / G, g = lane % G, b = inst0 + li, nv = h.nv;
  const int nvalid = (a.B - inst0) < IPW ? (a.B - inst0) : IPW;
  const bool valid = li < nvalid;
  const int SS = 7 * h.nnode | 1;
  float* xs = wsm + li * SS;
  const float* qb = a.q + (long long)(valid ? b : inst0) * h.nq;
  const int32_t* prog = P.i(h.off_prog);
  for (int s = 0; s < h.nsteps; ++s) {
    int n = prog[s * G + g];
    if (valid && n >= 0) fk_node(P, n, qb, xs);
    BIK_SYNCWARP();
  }
  if (!valid) return;
  for (int f = g; f < a.nframes; f += G) {
    const FrameRec& fr = a.frames[f];
    FQ qf; F3 pf;
    frame_pose(fr.node, fr.lpos, fr.lquat, xs, &qf, &pf);
    FM R = q2mat(qf);
    if (a.poses) {
      FQ qs = mat2quat(R);  // sign convention of SO3.from_matrix (configuration.py:182)
      float* o = a.poses + ((long long)b * a.nframes + f) * 7;
      o[0] = qs.w; o[1] = qs.x; o[2] = qs.y; o[3] = qs.z; o[4] = pf.x; o[5] = pf.y; o[6] = pf.z;
    }
    if (a.J) {  // blockdiag(R^T, R^T) [jacp; jacr]  (configuration.py:143-153)
      float* o = a.J + ((long long)b * a.nframes + f) * 6 * nv;
      for (int k = 0; k < 6 * nv; ++k) o[k] = 0.f;
      FM Rt = mtrans(R);
      for (int n = fr.node; n >= 0; n = P.node(n).parent) {
        const NodeRec& r = P.node(n);
        int nd = r.type == JNT_FREE ? 6 : (r.type == JNT_BALL ? 3 : 1);
        for (int k = 0; k < nd; ++k) {
          F3 jp, jr;
          int d = r.dadr + k;
          jac_column(P, d, n, xs, pf, &jp, &jr);
          F3 top = mmul(Rt, jp), bot = mmul(Rt, jr);
          o[d] = top.x; o[nv + d] = top.y; o[2 * nv + d] = top.z; o[3 * nv + d] = bot.x; o[4 * nv + d] = bot.y; o[5 * nv + d] = bot.z;
        }
      }
    }
  }
  if (a.com && g == 0) {
    F3 tot = ld_v(h.com_fixed);
    for (int n = 0; n < h.nnode; ++n) {
      const ComNodeRec& cn = P.comnode(n);
      tot = tot + cn.own_m * ld_v(xs + 7 * n + 4) + qrot(ld_q(xs + 7 * n), ld_v(cn.own_c));
    }
    float invM = h.com_total_mass > 0.f ? 1.f / h.com_total_mass : 0.f;
    a.com[(long long)b * 3] = tot.x * invM; a.com[(long long)b * 3 + 1] = tot.y * invM; a.com[(long long)b * 3 + 2] = tot.z * invM;
  }
}

// ---- integrate / check_limits / box limits: one thread per instance ----------------------------
BIK_HD void integrate_instance(const PView& P, float* q, const float* dq) {  // mj_integratePos, dt folded into dq
  const PHeader& h = P.h();
  for (int n = 0; n < h.nnode; ++n) {
    const NodeRec& r = P.node(n);
    if (r.type == JNT_FREE) {
      for (int k = 0; k < 3; ++k) q[r.qadr + k] += dq[r.dadr + k];
      FQ o = quat_integrate<float>(ld_q(q + r.qadr + 3), ld_v(dq + r.dadr + 3));
      q[r.qadr + 3] = o.w; q[r.qadr + 4] = o.x; q[r.qadr + 5] = o.y; q[r.qadr + 6] = o.z;
    } else if (r.type == JNT_BALL) {
      FQ o = quat_integrate<float>(ld_q(q + r.qadr), ld_v(dq + r.dadr));
      q[r.qadr] = o.w; q[r.qadr + 1] = o.x; q[r.qadr + 2] = o.y; q[r.qadr + 3] = o.z;
    } else {
      q[r.qadr] += dq[r.dadr];
    }
  }
}
BIK_HD int check_limits_instance(const PView& P, const float* q, float tol) {  // configuration.py:77-110
  const PHeader& h = P.h();
  const int32_t* dofqadr = P.i(h.off_dofqadr);
  const float* rng = P.f(h.off_range);
  int bad = 0;
  for (int d = 0; d < h.nv; ++d) {
    int qa = dofqadr[d];
    if (qa < 0) continue;
    float v = q[qa];
    if (v < rng[d] - tol || v > rng[h.nv + d] + tol) bad = 1;
    if (!(v == v)) bad |= 4;
  }
  return bad;
}
// lo <= dq <= hi from ConfigurationLimit(s) and VelocityLimit (configuration_limit.py:98-124, velocity_limit.py:99-101)
BIK_HD void box_dof(const PView& P, int d, const float* q, float dt, float* lo, float* hi) {
  const PHeader& h = P.h();
  float l = -BIK_INF_F, u = BIK_INF_F;
  int qa = P.i(h.off_dofqadr)[d];
  if (qa >= 0) {
    float qi = q[qa];
    for (int c = 0; c < h.ncfg; ++c) {
      const float* p = P.f(h.off_cfg) + c * (2 + 2 * h.nv);
      float up = p[0] * (p[2 + h.nv + d] - qi), dn = p[0] * (qi - p[2 + d]);
      u = fminf(u, up); l = fmaxf(l, -dn);
    }
  }
  if (h.has_vel) { float vm = dt * P.f(h.off_vmax)[d]; u = fminf(u, vm); l = fmaxf(l, -vm); }
  *lo = l; *hi = u;
}

}  // namespace bik
