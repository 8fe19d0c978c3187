// TEST INFRASTRUCTURE -- compiles the kernel headers (mink_b200/csrc/bik_k1.h, bik_k2.h) for the host
// with one lane per instance (G = W = 1) so the device math can be checked against the oracle on a
// machine without a GPU.  Never shipped, never loaded by mink_b200/: the product path is CUDA only.
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../mink_b200/csrc/bik_build.h"
#include "../../mink_b200/csrc/bik_k2lr.h"
#include "../../mink_b200/csrc/bik_k2x.h"

using namespace bik;

static std::string g_err;
extern "C" const char* emu_last_error() { return g_err.c_str(); }

struct EmuProblem {
  std::vector<uint32_t> image;        // problem image (tasks + limits; FK visits only the nodes they need)
  std::vector<uint32_t> model_image;  // model-only image, as bik_model_create builds it (FK visits every node)
};

extern "C" void* emu_problem_create(const void* blob, size_t nbytes, const bik_task_desc* tasks, int ntasks, const bik_limit_desc* limits, int nlimits) {
  HostModel m;
  if (!parse_model_blob(blob, nbytes, &m, &g_err)) return nullptr;
  EmuProblem* p = new EmuProblem;
  if (!build_image(m, tasks, ntasks, limits, nlimits, 1, &p->image, &g_err)) { delete p; return nullptr; }
  if (!build_image(m, nullptr, 0, nullptr, 0, 1, &p->model_image, &g_err)) { delete p; return nullptr; }
  return p;
}
// header fields of the problem image the CPU tests pin: nv, nu, nfree, nnode, nneeded, nslots, G, nsteps
extern "C" void emu_header(void* prob, int32_t* out) {
  const PHeader& h = PView{static_cast<EmuProblem*>(prob)->image.data()}.h();
  out[0] = h.nv; out[1] = h.nu; out[2] = h.nfree; out[3] = h.nnode; out[4] = h.nneeded; out[5] = h.nslots; out[6] = h.G; out[7] = h.nsteps;
}
extern "C" void emu_set_knobs(void* prob, int sweeps, int rule) {   // what BIK_K2_SWEEPS / BIK_K2_RULE do in bik_problem_create
  PHeader* h = reinterpret_cast<PHeader*>(static_cast<EmuProblem*>(prob)->image.data());
  h->k2_sweeps = sweeps; h->k2_rule = rule;
}
extern "C" void emu_problem_destroy(void* p) { delete static_cast<EmuProblem*>(p); }

extern "C" int emu_lane_program(const void* blob, size_t nbytes, int G, int32_t* out, int cap) {
  HostModel m;
  if (!parse_model_blob(blob, nbytes, &m, &g_err)) return -1;
  std::vector<int32_t> prog; int nsteps;
  lane_program(m, G, &prog, &nsteps);
  if ((int)prog.size() > cap) return -2;
  memcpy(out, prog.data(), prog.size() * 4);
  return nsteps;
}

extern "C" int emu_fk_jac(void* prob, int B, const float* q, const float* ftgt, const float* ptgt, int pbatched, const float* ctgt, float dt,
                          float* J, float* e, float* ep, float* Gc, float* hc) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->image.data()};
  K1Args a{B, q, ftgt, ptgt, ctgt, pbatched, dt, J, e, ep, Gc, hc};
  std::vector<float> wsm(k1_warp_words(P.h(), 1) + 16);
  for (int b = 0; b < B; ++b) k1_warp_tile<1, 1>(P, a, b, wsm.data(), 0);
  return 0;
}

// Small-group fp64 path with the rollout state of bik_step (nsteps > 1): `warm` [B][nu] bytes and the dq of the previous
// step in `dq` (in/out), as step_core passes them from its second step on.
extern "C" int emu_solve_warm(void* prob, int B, const float* q, const float* J, const float* e, const float* ep, float dt, double damping,
                              float* dq, int32_t* status, int32_t* iters, signed char* warm) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->image.data()};
  if (P.h().npairs != 0 || P.h().nu < 1 || P.h().nu > K2T_NMAX) { g_err = "thread path not applicable"; return -1; }
  K2Args a{B, q, J, e, ep, nullptr, nullptr, dt, damping, dq, status, iters, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, warm, nullptr, nullptr};
  std::vector<double> tw(k2t_warp_bytes(P, 8, 1) / 8 + 16);
  for (int b = 0; b < B; ++b) {
    if (status) status[b] = 0;
    k2t_warp_tile<double, 1, 1>(P, a, b, tw.data(), 0);
  }
  return 0;
}

extern "C" int emu_solve(void* prob, int B, const float* q, const float* J, const float* e, const float* ep, const float* Gc, const float* hc,
                         float dt, double damping, int use_double, float* dq, int32_t* status, int32_t* iters, double* H, double* c, float* lo, float* hi) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->image.data()};
  K2Args a{B, q, J, e, ep, Gc, hc, dt, damping, dq, status, iters, H, c, lo, hi, 0, 0, 0, nullptr, nullptr, nullptr};
  std::vector<double> wsm((k2_warp_bytes(P.h(), 8) + k2lr_warp_bytes(P.h())) / 8 + 16);
  if (use_double == 3 || use_double == 4) {   // small-group path (3: fp64, 4: fp32), one lane per problem on the host
    if (P.h().npairs != 0 || P.h().nu < 1 || P.h().nu > K2T_NMAX) { g_err = "thread path not applicable"; return -1; }
    std::vector<double> tw(k2t_warp_bytes(P, 8, 1) / 8 + 16);
    for (int b = 0; b < B; ++b) {
      if (status) status[b] = 0;
      if (use_double == 3) k2t_warp_tile<double, 1, 1>(P, a, b, tw.data(), 0);
      else k2t_warp_tile<float, 1, 1>(P, a, b, tw.data(), 0);
    }
    return 0;
  }
  if (use_double == 5 || use_double == 6 || use_double == 7) {   // fixed-size thread-per-problem path (5: mixed fp32/fp64 without
    const int nu = P.h().nu;                                      // hand-over, 6: fp32, 7: mixed with fp64 hand-over of flagged instances)
    if (P.h().npairs != 0 || nu < 1 || nu > 24) { g_err = "fixed-size path not applicable"; return -1; }
    const int N = nu <= 6 ? 6 : (nu <= 12 ? 12 : (nu <= 18 ? 18 : 24));
    std::vector<double> tw(k2x_warp_smem_bytes(P, 8, 8, N, 1) / 8 + 16), sc(k2x_warp_scratch_bytes(8, N, 1) / 8 + 16);
    std::vector<double> gw(k2t_warp_bytes(P, 8, 1) / 8 + 16);
    int nflag = 0;
    for (int b = 0; b < B; ++b) {
      if (status) status[b] = 0;
      int32_t flag = 0;
      K2Args ab = a;
      ab.flag_out = use_double == 7 ? &flag - b : nullptr;   // indexed by instance
      if (use_double != 6) {
        if (N == 6) k2x_warp_tile<float, double, 6, 1>(P, ab, b, tw.data(), sc.data(), 0);
        else if (N == 12) k2x_warp_tile<float, double, 12, 1>(P, ab, b, tw.data(), sc.data(), 0);
        else if (N == 18) k2x_warp_tile<float, double, 18, 1>(P, ab, b, tw.data(), sc.data(), 0);
        else k2x_warp_tile<float, double, 24, 1>(P, ab, b, tw.data(), sc.data(), 0);
        if (flag) { ++nflag; ab.flag_out = nullptr; k2t_warp_tile<double, 1, 1>(P, ab, b, gw.data(), 0); }
      } else {
        if (N == 6) k2x_warp_tile<float, float, 6, 1>(P, a, b, tw.data(), sc.data(), 0);
        else if (N == 12) k2x_warp_tile<float, float, 12, 1>(P, a, b, tw.data(), sc.data(), 0);
        else if (N == 18) k2x_warp_tile<float, float, 18, 1>(P, a, b, tw.data(), sc.data(), 0);
        else k2x_warp_tile<float, float, 24, 1>(P, a, b, tw.data(), sc.data(), 0);
      }
    }
    return nflag;
  }
  for (int b = 0; b < B; ++b) {
    if (status) status[b] = 0;
    if (use_double == 2) {   // low-rank (Woodbury) path
      std::vector<uint16_t> pairtab(tri(P.h().K) + 1);
      for (int p = 0; p < tri(P.h().K); ++p) { int r, s; tri_unflatten(p, &r, &s); pairtab[p] = (uint16_t)((r << 8) | s); }
      k2lr_warp<1, 65>(P, a, b, wsm.data(), pairtab.data(), 0);
    }
    else if (use_double) k2_warp<double, 1, 65>(P, a, b, wsm.data(), 0);
    else k2_warp<float, 1, 65>(P, a, b, wsm.data(), 0);
  }
  return 0;
}

extern "C" int emu_fk(void* prob, int B, const float* q, const bik_frame* frames, int nframes, float* poses, float* com, float* J) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->model_image.data()};
  if (nframes > 16) return -1;
  FkArgs a; memset(&a, 0, sizeof a);
  a.B = B; a.nframes = nframes; a.q = q; a.poses = poses; a.com = com; a.J = J;
  for (int f = 0; f < nframes; ++f) put_frame(frames[f], &a.frames[f].node, a.frames[f].lpos, a.frames[f].lquat);
  std::vector<float> wsm(7 * P.h().nnode + 16);
  for (int b = 0; b < B; ++b) fk_warp_tile<1, 1>(P, a, b, wsm.data(), 0);
  return 0;
}

extern "C" int emu_integrate(void* prob, int B, float* q, const float* dq) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->image.data()};
  for (int b = 0; b < B; ++b) integrate_instance(P, q + (size_t)b * P.h().nq, dq + (size_t)b * P.h().nv);
  return 0;
}
extern "C" int emu_check_limits(void* prob, int B, const float* q, float tol, int32_t* status) {
  EmuProblem* p = static_cast<EmuProblem*>(prob);
  PView P{p->image.data()};
  for (int b = 0; b < B; ++b) status[b] = check_limits_instance(P, q + (size_t)b * P.h().nq, tol);
  return 0;
}
