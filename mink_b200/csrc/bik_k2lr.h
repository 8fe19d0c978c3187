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
// the same KKT point.  Arithmetic is fp64 except the storage of U (fp32, like J itself): Woodbury
// subtracts nearly equal terms, which fp64 absorbs (cond(H) eps64 ~ 1e-11) and fp32 would not.
//
// Chosen by the dispatcher when K is well below nv, there are no general (collision) rows and
// `damping` keeps D away from zero; otherwise the dense path runs.
#pragma once
#include "bik_k2.h"

namespace bik {

BIK_HD int k2lr_warp_bytes(const PHeader& h) {
  int n = h.nv, K = h.K;
  int words_T = tri(K) + tri(K + 1) + 3 * K + 5 * n + 4;
  int bytes = words_T * 8 + 4 * (K * n + K + 4) + 4 * (2 * n + 4);
  return (bytes + 15) & ~15;
}

BIK_HD void tri_unflatten(int p, int* r, int* s) {
  int rr = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while (tri(rr + 1) <= p) ++rr;
  while (tri(rr) > p) --rr;
  *r = rr; *s = p - tri(rr);
}

template <int W, int SLOTS>
BIK_HD void k2lr_warp(const PView& P, const K2Args& a, int b, void* wsm, int lane) {
  typedef double T;
  const PHeader& h = P.h();
  const int n = h.nv, K = h.K;
  // ---- carve ------------------------------------------------------------------------------------
  T* M0 = reinterpret_cast<T*>(wsm);          // tri(K): I + U_F U_F^T, kept up to date across iterations
  T* Lp = M0 + tri(K);                        // tri(K+1): factor workspace, row K = right-hand side
  T* dM = Lp + tri(K + 1);                    // K: inverse diagonal of the factor
  T* tv = dM + K;                             // K
  T* yv = tv + K;                             // K
  T* ct = yv + K;                             // n: c~
  T* lo = ct + n; T* hi = lo + n; T* x = hi + n; T* sd = x + n;   // n each
  float* U = reinterpret_cast<float*>(sd + n + 4);
  float* we = U + K * n;
  int* st = reinterpret_cast<int*>(we + K + 4);
  int* nst = st + n;

  const float* Jb = a.J + (long long)b * K * n;
  const float* eb = a.e + (long long)b * K;
  // ---- weighted rows W J and W(-gain e) (task.py:128-129) -------------------------------------------
  for (int f = 0; f < h.F; ++f) {
    const FrameRec& fr = P.frame(f);
    for (int k = lane; k < 6 * n; k += W) { int r = k / n; U[fr.row0 * n + k] = fr.cost[r] * Jb[fr.row0 * n + k]; }
    for (int r = lane; r < 6; r += W) we[fr.row0 + r] = fr.cost[r] * (-fr.gain * eb[fr.row0 + r]);
  }
  for (int c = 0; c < h.C; ++c) {
    const float* cr = P.f(h.off_com) + 8 * c;
    int row0 = reinterpret_cast<const int32_t*>(cr)[5];
    for (int k = lane; k < 3 * n; k += W) { int r = k / n; U[row0 * n + k] = cr[r] * Jb[row0 * n + k]; }
    for (int r = lane; r < 3; r += W) we[row0 + r] = cr[r] * (-cr[3] * eb[row0 + r]);
  }
  BIK_SYNCWARP();
  // ---- mu = damping + sum_t lm_t ||W(-gain e)||^2 (task.py:131), every lane in the same order ---------
  T mu = T(a.damping);
  for (int f = 0; f < h.F; ++f) {
    const FrameRec& fr = P.frame(f);
    if (fr.lm != 0.f) { T s = 0; for (int r = 0; r < 6; ++r) s += T(we[fr.row0 + r]) * T(we[fr.row0 + r]); mu += T(fr.lm) * s; }
  }
  for (int c = 0; c < h.C; ++c) {
    const float* cr = P.f(h.off_com) + 8 * c;
    int row0 = reinterpret_cast<const int32_t*>(cr)[5];
    if (cr[4] != 0.f) { T s = 0; for (int r = 0; r < 3; ++r) s += T(we[row0 + r]) * T(we[row0 + r]); mu += T(cr[4]) * s; }
  }
  for (int p = 0; p < h.P; ++p) {
    const float* pr = P.f(h.off_posture) + p * (2 + n);
    if (pr[1] != 0.f) {
      const float* epb = a.ep + ((long long)b * h.P + p) * n;
      T s = 0;
      for (int d = 0; d < n; ++d) { T v = T(pr[2 + d]) * T(pr[0]) * T(epb[d]); s += v * v; }
      mu += T(pr[1]) * s;
    }
  }
  // ---- per dof: diagonal d, linear term, box; then scale to x~ = sqrt(d) x --------------------------
  int status = 0;
  for (int d = lane; d < n; d += W) {
    T cd = 0;
    for (int r = 0; r < K; ++r) cd -= T(we[r]) * T(U[r * n + d]);
    T hd = mu;
    for (int p = 0; p < h.P; ++p) {
      const float* pr = P.f(h.off_posture) + p * (2 + n);
      T wgt = T(pr[2 + d]);
      hd += wgt * wgt;
      cd -= T(pr[0]) * wgt * wgt * T(a.ep[((long long)b * h.P + p) * n + d]);
    }
    if (!(hd > T(0))) { status |= 4; hd = T(1); }
    T s = bik_sqrt<T>(hd), is = T(1) / s;
    float blo, bhi;
    box_dof(P, d, a.q + (long long)b * h.nq, a.dt, &blo, &bhi);
    sd[d] = s; ct[d] = cd * is; lo[d] = T(blo) * s; hi[d] = T(bhi) * s;
    st[d] = 0;
    for (int r = 0; r < K; ++r) U[r * n + d] = float(T(U[r * n + d]) * is);
  }
  BIK_SYNCWARP();
  // ---- M0 = I + U U^T over all dofs (everything free) -------------------------------------------
  for (int p = lane; p < tri(K); p += W) {
    int r, s;
    tri_unflatten(p, &r, &s);
    const float* Ur = U + r * n; const float* Us = U + s * n;
    T a0 = 0, a1 = 0;
    int i = 0;
    for (; i + 1 < n; i += 2) { a0 += T(Ur[i]) * T(Us[i]); a1 += T(Ur[i + 1]) * T(Us[i + 1]); }
    if (i < n) a0 += T(Ur[i]) * T(Us[i]);
    M0[p] = a0 + a1 + (r == s ? T(1) : T(0));
  }
  BIK_SYNCWARP();

  const int MAXIT = 60, PATIENCE = 3;
  const T tolx = K2Tol<T>::x(), tolg = K2Tol<T>::g();
  int best = n + 1, patience = PATIENCE, it = 0, nactive = 0;
  for (; it < MAXIT; ++it) {
    // t = U_B x~_B
    for (int r = lane; r < K; r += W) {
      T t = 0;
      if (nactive) { const float* Ur = U + r * n; for (int j = 0; j < n; ++j) if (st[j]) t += T(Ur[j]) * (st[j] == 1 ? lo[j] : hi[j]); }
      tv[r] = t;
    }
    BIK_SYNCWARP();
    // z on the free set (kept in x), bounds on the active set
    for (int i = lane; i < n; i += W) {
      T v;
      if (st[i] == 0) { v = -ct[i]; if (nactive) for (int r = 0; r < K; ++r) v -= T(U[r * n + i]) * tv[r]; }
      else v = st[i] == 1 ? lo[i] : hi[i];
      x[i] = v;
    }
    BIK_SYNCWARP();
    // right-hand side r = U_F z and a fresh copy of M0 for the factorisation
    for (int r = lane; r < K; r += W) {
      const float* Ur = U + r * n;
      T acc = 0;
      for (int j = 0; j < n; ++j) if (st[j] == 0) acc += T(Ur[j]) * x[j];
      Lp[tri(K) + r] = acc;
      for (int s = 0; s <= r; ++s) Lp[tri(r) + s] = M0[tri(r) + s];
    }
    BIK_SYNCWARP();
    if (k2_factor<T, W, SLOTS>(Lp, dM, K, lane)) status |= 4;
    k2_backsub<T, W, SLOTS>(Lp, dM, K, yv, lane);
    // x~_F = z - U_F^T y
    for (int i = lane; i < n; i += W)
      if (st[i] == 0) { T v = x[i]; for (int r = 0; r < K; ++r) v -= T(U[r * n + i]) * yv[r]; x[i] = v; }
    BIK_SYNCWARP();
    // s = U x~ for the gradient on the active set
    if (nactive) {
      for (int r = lane; r < K; r += W) { const float* Ur = U + r * n; T acc = 0; for (int j = 0; j < n; ++j) acc += T(Ur[j]) * x[j]; tv[r] = acc; }
      BIK_SYNCWARP();
    }
    int ninf = 0, last = -1;
    for (int i = lane; i < n; i += W) {
      int cur = st[i], ns = cur;
      if (cur == 0) {
        T xi = x[i];
        if (xi < lo[i] - tolx * (T(1) + (lo[i] < 0 ? -lo[i] : lo[i]))) ns = 1;
        else if (xi > hi[i] + tolx * (T(1) + (hi[i] < 0 ? -hi[i] : hi[i]))) ns = 2;
      } else {
        T gi = ct[i] + x[i];
        for (int r = 0; r < K; ++r) gi += T(U[r * n + i]) * tv[r];
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
    // apply the flips: rank-one up/down-dates of M0, then the state itself
    for (int i = 0; i < n; ++i) {
      int cur = st[i], ns = nst[i];
      if (ns == cur || !(block || i == last)) continue;
      if ((cur == 0) != (ns == 0)) {
        T sgn = (ns == 0) ? T(1) : T(-1);
        for (int p = lane; p < tri(K); p += W) { int r, s; tri_unflatten(p, &r, &s); M0[p] += sgn * T(U[r * n + i]) * T(U[s * n + i]); }
      }
    }
    BIK_SYNCWARP();
    nactive = 0;
    for (int i = 0; i < n; ++i) {
      int v = (nst[i] != st[i] && (block || i == last)) ? nst[i] : st[i];
      nactive += v != 0;
    }
    BIK_SYNCWARP();
    for (int i = lane; i < n; i += W) if (block || i == last) st[i] = nst[i];
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
}

}  // namespace bik
