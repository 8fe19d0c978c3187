// bik_k2lr.h -- K2, low-rank path: exact box-QP solve through the task-space capacitance matrix.
//
// For the workloads mink is used on, the QP Hessian (reference mink/tasks/task.py:125-138,
// mink/solve_ik.py:13-22) is a diagonal plus a low-rank term:
//     H = D + A^T A,   D = (damping + sum_t mu_t) I + sum_posture diag(cost^2),   A = stacked W_t J_t  (K x nv)
// with K = 6 #FrameTasks + 3 #ComTasks (G1 config: K = 18, nv = 43).  After the change of variables
// x~ = D^{1/2} x the problem is   min 1/2 x~^T (I + U^T U) x~ + c~^T x~,  lo~ <= x~ <= hi~,  U = A D^{-1/2}.
// For a guess of the active set (B at bounds, F free) the reduced system is solved EXACTLY by the
// Sherman-Morrison-Woodbury identity,
//     x~_F = z - U_F^T M^{-1} U_F z,   z = -c~_F - U_F^T (U_B x~_B),   M = I_K + U_F U_F^T   (K x K, SPD)
// so one pivoting iteration costs a K x K Cholesky plus a few K x nv mat-vecs instead of an
// nv x nv factorisation, and moving index i between F and B is the rank-one change M -+= u_i u_i^T.
// The pivoting rule, tolerances and termination are those of the dense path (bik_k2.h): the result is
// the same KKT point.  Everything is fp64: Woodbury subtracts nearly equal terms, which fp64 absorbs
// (cond(H) eps64 ~ 1e-11) and fp32 would not.
//
// Code shape: the six mat-vecs of an iteration go through two small NON-inlined helpers so that the
// pivoting loop stays inside the instruction cache (the first version, fully inlined, was fetch-bound:
// profiles/r1_k2lr_v1.md).  Inactive entries are handled by zero-masked vectors, not by branches.
//
// Chosen by the dispatcher when K is well below nv, there are no general (collision) rows and
// `damping` keeps D away from zero; otherwise the dense path runs.
#pragma once
#include "bik_k2.h"

#ifndef BIK_LR_UNROLL
#define BIK_LR_UNROLL 2
#endif
#define BIK_PRAGMA_(x) _Pragma(#x)
#define BIK_UNROLL(n) BIK_PRAGMA_(unroll n)
#ifndef BIK_LR_U64
typedef float lr_u_t;   // U stored in fp32, products accumulated in fp64 (more warps per SM, one convert per load)
#else
typedef double lr_u_t;
#endif

namespace bik {

enum { K2LR_MAX_PAIRS = 8 };  // packed-triangle entries of M per lane: tri(K) <= 32 * 8  (K <= 22); larger K loops

BIK_HD int k2lr_ld(const PHeader& h) { return h.nv | 1; }
BIK_HD int k2lr_warp_bytes(const PHeader& h) {
  int n = h.nv, K = h.K;
  int words_T = tri(K) + tri(K + 1) + 4 * K + 8 * n + 8;
  int bytes = words_T * 8 + ((K * k2lr_ld(h) * (int)sizeof(lr_u_t) + 15) & ~15) + 4 * (2 * n + 8);
  return (bytes + 15) & ~15;
}

BIK_HD void tri_unflatten(int p, int* r, int* s) {
  int rr = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while (tri(rr + 1) <= p) ++rr;
  while (tri(rr) > p) --rr;
  *r = rr; *s = p - tri(rr);
}

template <int W> BIK_HD double warp_sum_d(double v) {
#if defined(__CUDA_ARCH__)
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// out[r] = sum_j U[r][j] v[j]            (lane <-> row r)
template <int W>
BIK_NOINLINE void lr_rowop(const lr_u_t* __restrict__ U, int ld, int K, int n, const double* __restrict__ v, double* __restrict__ out, int lane) {
  for (int r = lane; r < K; r += W) {
    const lr_u_t* Ur = U + r * ld;
    double a0 = 0, a1 = 0;
    int j = 0;
BIK_UNROLL(BIK_LR_UNROLL)
    for (; j + 1 < n; j += 2) { a0 += double(Ur[j]) * v[j]; a1 += double(Ur[j + 1]) * v[j + 1]; }
    if (j < n) a0 += double(Ur[j]) * v[j];
    out[r] = a0 + a1;
  }
  BIK_SYNCWARP();
}
// out[i] = sum_r U[r][i] w[r]            (lane <-> column i)
template <int W>
BIK_NOINLINE void lr_colop(const lr_u_t* __restrict__ U, int ld, int K, int n, const double* __restrict__ w, double* __restrict__ out, int lane) {
  for (int i = lane; i < n; i += W) {
    const lr_u_t* Ui = U + i;
    double a0 = 0, a1 = 0;
    int r = 0;
BIK_UNROLL(BIK_LR_UNROLL)
    for (; r + 1 < K; r += 2) { a0 += double(Ui[r * ld]) * w[r]; a1 += double(Ui[(r + 1) * ld]) * w[r + 1]; }
    if (r < K) a0 += double(Ui[r * ld]) * w[r];
    out[i] = a0 + a1;
  }
  BIK_SYNCWARP();
}
// M[p] += sgn * U[r][i] U[s][i] over this lane's packed-triangle entries; pairtab[p] = r << 8 | s
template <int W>
BIK_NOINLINE void lr_rank1(double* __restrict__ M, const lr_u_t* __restrict__ U, const uint16_t* __restrict__ pairtab, int ld, int npairs, int i,
                           double sgn, int lane) {
  for (int p = lane; p < npairs; p += W) {
    int rs = pairtab[p];
    M[p] += sgn * double(U[(rs >> 8) * ld + i]) * double(U[(rs & 0xff) * ld + i]);
  }
}

template <int W, int SLOTS>
BIK_HD void k2lr_warp(const PView& P, const K2Args& a, int b, void* wsm, const uint16_t* pairtab, int lane) {
  typedef double T;
  const PHeader& h = P.h();
  const int n = h.nv, K = h.K, ld = k2lr_ld(h), NP = tri(K);
  // ---- carve ------------------------------------------------------------------------------------
  T* M0 = reinterpret_cast<T*>(wsm);           // tri(K): I + U_F U_F^T, maintained across iterations
  T* Lp = M0 + NP;                              // tri(K+1): factor workspace, row K = right-hand side
  T* dM = Lp + tri(K + 1);                      // K: inverse diagonal of the factor
  T* tv = dM + K; T* yv = tv + K; T* we = yv + K;   // K each
  T* ct = we + K;                               // n: c~
  T* lo = ct + n; T* hi = lo + n; T* x = hi + n; T* sd = x + n; T* xb = sd + n; T* zf = xb + n; T* tmp = zf + n;
  lr_u_t* U = reinterpret_cast<lr_u_t*>(tmp + n + 8);   // K x ld: scaled weighted Jacobian rows
  int* st = reinterpret_cast<int*>(reinterpret_cast<char*>(U) + ((K * ld * (int)sizeof(lr_u_t) + 15) & ~15));
  int* nst = st + n;

  const float* Jb = a.J + (long long)b * K * n;
  const float* eb = a.e + (long long)b * K;
  const float* rowcost = P.f(h.off_rowinfo);
  const float* rowgain = rowcost + K;
  const float* rowlm = rowgain + K;
  // ---- W(-gain e) and mu = damping + sum_t lm_t ||W(-gain e)||^2 (task.py:128-131) --------------------
  T mu_part = 0;
  for (int r = lane; r < K; r += W) {
    T w = T(rowcost[r]) * (-T(rowgain[r]) * T(eb[r]));
    we[r] = w;
    mu_part += T(rowlm[r]) * w * w;
  }
  for (int p = 0; p < h.P; ++p) {
    const float* pr = P.f(h.off_posture) + p * (2 + n);
    if (pr[1] != 0.f) {
      const float* epb = a.ep + ((long long)b * h.P + p) * n;
      for (int d = lane; d < n; d += W) { T v = T(pr[2 + d]) * T(pr[0]) * T(epb[d]); mu_part += T(pr[1]) * v * v; }
    }
  }
  const T mu = T(a.damping) + warp_sum_d<W>(mu_part);
  BIK_SYNCWARP();
  // ---- per dof: diagonal d, linear term, box, scaled column of U -------------------------------------
  int status = 0;
  for (int d = lane; d < n; d += W) {
    T hd = mu, cd = 0;
    for (int p = 0; p < h.P; ++p) {
      const float* pr = P.f(h.off_posture) + p * (2 + n);
      T wgt = T(pr[2 + d]);
      hd += wgt * wgt;
      cd -= T(pr[0]) * wgt * wgt * T(a.ep[((long long)b * h.P + p) * n + d]);
    }
    if (!(hd > T(0))) { status |= 4; hd = T(1); }
    const T is = bik_rsqrt<T>(hd), s = hd * is;
    for (int r = 0; r < K; ++r) {
      T wj = T(rowcost[r]) * T(Jb[r * n + d]);
      cd -= we[r] * wj;
      U[r * ld + d] = lr_u_t(wj * is);
    }
    float blo, bhi;
    box_dof(P, d, a.q + (long long)b * h.nq, a.dt, &blo, &bhi);
    sd[d] = s; ct[d] = cd * is; lo[d] = T(blo) * s; hi[d] = T(bhi) * s;
    st[d] = 0; xb[d] = T(0);
  }
  BIK_SYNCWARP();
  // ---- M0 = I + U U^T over all dofs (everything free) -------------------------------------------
  for (int p = lane; p < NP; p += W) {
    const int rs = pairtab[p], r = rs >> 8, s = rs & 0xff;
    const lr_u_t* Ur = U + r * ld; const lr_u_t* Us = U + s * ld;
    T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = 0;
    for (; i + 3 < n; i += 4) {
      T u0 = T(Ur[i]), u1 = T(Ur[i + 1]), u2 = T(Ur[i + 2]), u3 = T(Ur[i + 3]), v0 = T(Us[i]), v1 = T(Us[i + 1]), v2 = T(Us[i + 2]), v3 = T(Us[i + 3]);
      a0 += u0 * v0; a1 += u1 * v1; a2 += u2 * v2; a3 += u3 * v3;
    }
    for (; i < n; ++i) a0 += T(Ur[i]) * T(Us[i]);
    M0[p] = (a0 + a1) + (a2 + a3) + (r == s ? T(1) : T(0));
  }
  BIK_SYNCWARP();

  const int MAXIT = 400, PATIENCE = 3;   // the single-pivot fallback is finite but slow: stalled instances of a rollout need up to ~150 pivots (was 60: 1-5 per thousand flagged)
  const T tolx = K2Tol<T>::x(), tolg = K2Tol<T>::g();
  int best = n + 1, patience = PATIENCE, it = 0, nactive = 0;
  for (; it < MAXIT; ++it) {
    // t = U xb  (xb = bound value on the active set, 0 elsewhere);  z = -c~ - U^T t on the free set
    if (nactive) {
      lr_rowop<W>(U, ld, K, n, xb, tv, lane);
      lr_colop<W>(U, ld, K, n, tv, tmp, lane);
    }
    for (int i = lane; i < n; i += W) zf[i] = st[i] == 0 ? -ct[i] - (nactive ? tmp[i] : T(0)) : T(0);
    BIK_SYNCWARP();
    // right-hand side r = U zf (row K of the factor workspace) and a fresh copy of M0
    lr_rowop<W>(U, ld, K, n, zf, Lp + tri(K), lane);
    for (int p = lane; p < NP; p += W) Lp[p] = M0[p];
    BIK_SYNCWARP();
    if (k2_factor<T, W, SLOTS>(Lp, dM, K, lane)) status |= 4;
    k2_backsub<T, W, SLOTS>(Lp, dM, K, yv, lane);
    // x~ = zf - U^T y on the free set, bounds on the active set
    lr_colop<W>(U, ld, K, n, yv, tmp, lane);
    for (int i = lane; i < n; i += W) x[i] = st[i] == 0 ? zf[i] - tmp[i] : xb[i];
    BIK_SYNCWARP();
    // gradient on the active set: g~ = c~ + x~ + U^T (U x~)
    if (nactive) {
      lr_rowop<W>(U, ld, K, n, x, tv, lane);
      lr_colop<W>(U, ld, K, n, tv, tmp, lane);
    }
    int ninf = 0, last = -1;
    for (int i = lane; i < n; i += W) {
      int cur = st[i], ns = cur;
      if (cur == 0) {
        T xi = x[i];
        if (xi < lo[i] - tolx * (T(1) + (lo[i] < 0 ? -lo[i] : lo[i]))) ns = 1;
        else if (xi > hi[i] + tolx * (T(1) + (hi[i] < 0 ? -hi[i] : hi[i]))) ns = 2;
      } else {
        T gi = ct[i] + x[i] + tmp[i];
        if (cur == 1 && gi < -tolg) ns = 0;
        else if (cur == 2 && gi > tolg) ns = 0;
      }
      nst[i] = ns;
      if (ns != cur) { ++ninf; last = i > last ? i : last; }
    }
    ninf = warp_sum_i<W>(ninf);
    last = warp_max_i<W>(last);
    if (ninf == 0) break;
    bool block;
    if (ninf < best) { best = ninf; patience = PATIENCE; block = true; }
    else if (patience > 0) { --patience; block = true; }
    else block = false;
    BIK_SYNCWARP();
    // apply the flips: rank-one up/down-dates of M0 (lanes over packed entries), then the state itself
    nactive = 0;
    for (int i = 0; i < n; ++i) {
      int cur = st[i], ns = nst[i];
      if (ns != cur && (block || i == last)) {
        if ((cur == 0) != (ns == 0)) lr_rank1<W>(M0, U, pairtab, ld, NP, i, ns == 0 ? T(1) : T(-1), lane);
        cur = ns;
      }
      nactive += cur != 0;
    }
    BIK_SYNCWARP();
    for (int i = lane; i < n; i += W) {
      if (block || i == last) st[i] = nst[i];
      xb[i] = st[i] == 1 ? lo[i] : (st[i] == 2 ? hi[i] : T(0));
    }
    BIK_SYNCWARP();
  }
  if (it >= MAXIT) status |= 2;
  for (int d = lane; d < n; d += W) {
    T v = x[d] / sd[d];
    if (!(v == v)) status |= 4;
    a.dq[(long long)b * n + d] = float(v);
  }
  status = warp_max_i<W>(status & 2) | warp_max_i<W>(status & 4);
  if (lane == 0) {
    if (a.status) a.status[b] |= status;
    if (a.iters) a.iters[b] = it + 1;
  }
  BIK_SYNCWARP();
}

}  // namespace bik
