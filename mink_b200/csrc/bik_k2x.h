// bik_k2x.h -- K2, thread-per-problem path with the coupled size fixed at compile time.
//
// Same QP, same block principal pivoting rule and tolerances as bik_k2.h / bik_k2t.h.  One THREAD owns one
// problem; the coupled block is padded to a template size N (padding dofs: unit diagonal, zero linear term,
// infinite box) and every loop over the matrix dimension is unrolled with NO run-time guards, so a pivoting
// iteration is a single straight-line instruction stream:
//
//   * rows being produced and all vectors (x, multipliers, right-hand side) live in registers with compile-time
//     indices; the factor L (packed, reciprocal-sqrt diagonal) lives in shared memory, word w of lane l at
//     [w*W + l] (conflict-free, immediate offsets); two rows are produced per pass so that each loaded L entry
//     feeds two FMAs, and the right-hand side rides along as the last row (forward substitution for free);
//   * H, c and the box are kept in a per-warp global scratch, same lane-interleaved layout (coalesced, L2
//     resident, read with ld.global.cg): one pass per iteration copies H into the factor buffer with the active
//     set masked in (clamped dofs become identity rows/columns, so all lanes run the same code whatever their
//     active sets are) and accumulates H_FA x_A on the way, one more pass evaluates the multipliers;
//   * shared memory per problem is just the factor, which is what bounds the number of resident problems;
//   * instance data from K1 is staged a task at a time through an instance-major tile (bik_k2t.h staging),
//     6x6 column blocks of (W J)^T (W J) are accumulated in registers and added to the scratch in batches.
//
// Two instantiations: all-fp32 (BIK_SOLVE_PRECISION=f32), and MIXED -- fp32 factorisations of an fp64 H with the final
// iterate polished by iterative refinement and re-judged with tight tolerances (see k2x_warp_tile); problems the fp32
// factorisation cannot carry are handed to the fp64 small-group kernel.
// W = 32 on the device, 1 in the host emulation (tests/host_emu).
#pragma once
#include "bik_k2t.h"

#if defined(__CUDA_ARCH__)
#define BIK_LDCG(p) __ldcg(p)
#else
#define BIK_LDCG(p) (*(p))
#endif

namespace bik {

// per-warp shared memory: the factor (tri(N) words of the factor type per lane), aliased by the staging tiles during assembly
BIK_HD int k2x_warp_smem_bytes(const PView& P, int tl, int th, int N, int W) {
  const PHeader& h = P.h();
  size_t f = (size_t)tri(N) * W * tl;
  size_t t = (size_t)k2t_task_tile_words(P) * W * th;
  size_t q = (size_t)((h.nq + h.P * h.nv) | 1) * W * 4;
  size_t m = f > t ? f : t;
  m = m > q ? m : q;
  return (int)((m + 15) & ~(size_t)15);
}
// per-warp global scratch: H tri(N) | c N (words of the matrix type per lane), then lo N | hi N (floats per lane)
BIK_HD size_t k2x_warp_scratch_bytes(int th, int N, int W) { return (((size_t)(tri(N) + N) * th + (size_t)2 * N * 4) * W + 255) & ~(size_t)255; }

// decision tolerances once the iterate has been polished in the wide type (mixed precision only)
struct K2TolPolished { static BIK_HD double x() { return 1e-9; } static BIK_HD double g() { return 1e-7; } };

// (W J)^T (W J) and the linear term of one task, 6x6 column blocks in registers, added to the scratch in batches.
template <typename T, int W>
BIK_HD void k2x_task_accumulate(const T* tr, int nr, int nc, const int32_t* cols, const int32_t* umap, T* Hg, T* cg, float lm, T* mu) {
  constexpr int CB = 6;
  const T* wev = tr + nr * nc;
  if (lm != 0.f) { T s = T(0); for (int r = 0; r < nr; ++r) s += wev[r] * wev[r]; *mu += T(lm) * s; }
  for (int a0 = 0; a0 < nc; a0 += CB) {
    int ua[CB];
#pragma unroll
    for (int ia = 0; ia < CB; ++ia) ua[ia] = (a0 + ia < nc) ? umap[cols[a0 + ia] & 0xffff] : -1;
    for (int b0 = 0; b0 <= a0; b0 += CB) {
      int ub[CB];
#pragma unroll
      for (int ib = 0; ib < CB; ++ib) ub[ib] = (b0 + ib < nc) ? umap[cols[b0 + ib] & 0xffff] : -1;
      T acc[CB][CB], cacc[CB];
#pragma unroll
      for (int ia = 0; ia < CB; ++ia) { cacc[ia] = T(0);
#pragma unroll
        for (int ib = 0; ib < CB; ++ib) acc[ia][ib] = T(0); }
      for (int r = 0; r < nr; ++r) {
        const T* row = tr + r * nc;
        T A[CB], Bv[CB];
#pragma unroll
        for (int ia = 0; ia < CB; ++ia) A[ia] = (a0 + ia < nc) ? row[a0 + ia] : T(0);
#pragma unroll
        for (int ib = 0; ib < CB; ++ib) Bv[ib] = (b0 + ib < nc) ? row[b0 + ib] : T(0);
#pragma unroll
        for (int ia = 0; ia < CB; ++ia)
#pragma unroll
          for (int ib = 0; ib < CB; ++ib) acc[ia][ib] += A[ia] * Bv[ib];
        if (b0 == 0) {
          const T we = wev[r];
#pragma unroll
          for (int ia = 0; ia < CB; ++ia) cacc[ia] += we * A[ia];
        }
      }
      // batched read-modify-write: all loads first (independent, in flight together), then the stores
      int idx[CB][CB];
      T old[CB][CB];
#pragma unroll
      for (int ia = 0; ia < CB; ++ia)
#pragma unroll
        for (int ib = 0; ib < CB; ++ib) {
          const bool use = ua[ia] >= 0 && ub[ib] >= 0 && !(b0 == a0 && ib > ia);
          const int hi = ua[ia] > ub[ib] ? ua[ia] : ub[ib], lo = ua[ia] > ub[ib] ? ub[ib] : ua[ia];
          idx[ia][ib] = use ? (tri(hi) + lo) * W : -1;
          old[ia][ib] = use ? Hg[idx[ia][ib]] : T(0);
        }
#pragma unroll
      for (int ia = 0; ia < CB; ++ia)
#pragma unroll
        for (int ib = 0; ib < CB; ++ib)
          if (idx[ia][ib] >= 0) Hg[idx[ia][ib]] = old[ia][ib] + acc[ia][ib];
      if (b0 == 0) {
        T oc[CB];
#pragma unroll
        for (int ia = 0; ia < CB; ++ia) oc[ia] = ua[ia] >= 0 ? cg[ua[ia] * W] : T(0);
#pragma unroll
        for (int ia = 0; ia < CB; ++ia) if (ua[ia] >= 0) cg[ua[ia] * W] = oc[ia] - cacc[ia];
      }
    }
  }
}

// One pivoting iteration's linear algebra: masked copy of H into the factor buffer (+ H_FA x_A), two-row
// left-looking Cholesky with the right-hand side as last row, back substitution.  x = solution of the masked system.
// TL = type of the factorisation (storage and arithmetic), TH = type of H, c and the vectors.
template <typename TL, typename TH, int N, int W>
BIK_HD int k2x_solve_masked(const TH* __restrict__ Hg, const TH* __restrict__ cg, TL* __restrict__ Lp, uint32_t act, const TH (&xa)[N], TH (&x)[N]) {
  int bad = 0;
  TH rhs[N];
  {
    TH hv[N];
#pragma unroll
    for (int k = 0; k < N; ++k) hv[k] = TH(0);
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int k = 0; k <= i; ++k) {
        const TH hik = BIK_LDCG(Hg + (tri(i) + k) * W);
        if (k < i) {
          hv[i] += hik * xa[k];
          hv[k] += hik * xa[i];
          Lp[(tri(i) + k) * W] = (((act >> i) | (act >> k)) & 1u) ? TL(0) : TL(hik);
        } else {
          hv[i] += hik * xa[i];
          Lp[(tri(i) + i) * W] = ((act >> i) & 1u) ? TL(1) : TL(hik);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) rhs[k] = ((act >> k) & 1u) ? xa[k] : -(BIK_LDCG(cg + k * W) + hv[k]);
  }
  TL y[N];
#pragma unroll
  for (int i = 0; i <= N; i += 2) {
    const bool rhs0 = (i == N), has1 = (i + 1 <= N), rhs1 = (i + 1 == N);
    TL r0[N], r1[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { r0[k] = TL(0); r1[k] = TL(0); }
    TL ss0 = TL(0), s01 = TL(0), ss1 = TL(0);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k < i) {
        TL t0 = rhs0 ? TL(rhs[k]) : Lp[(tri(i) + k) * W];
        TL t1 = rhs1 ? TL(rhs[k]) : (has1 ? Lp[(tri(i + 1) + k) * W] : TL(0));
        TL e0 = TL(0), e1 = TL(0);
#pragma unroll
        for (int m = 0; m < k; ++m) {
          const TL lkm = Lp[(tri(k) + m) * W];
          if (m & 1) { e0 -= r0[m] * lkm; e1 -= r1[m] * lkm; } else { t0 -= r0[m] * lkm; t1 -= r1[m] * lkm; }
        }
        const TL dk = Lp[(tri(k) + k) * W];
        t0 = (t0 + e0) * dk; t1 = (t1 + e1) * dk;
        r0[k] = t0; r1[k] = t1;
        if (!rhs0) Lp[(tri(i) + k) * W] = t0;
        if (has1 && !rhs1) Lp[(tri(i + 1) + k) * W] = t1;
        ss0 += t0 * t0; s01 += t0 * t1; ss1 += t1 * t1;
      } else if (k == i) {
        TL dinv0 = TL(1);
        if (!rhs0) {
          TL d = Lp[(tri(i) + i) * W] - ss0;
          if (!(d > TL(0))) { bad = 1; d = TL(1e-30); }
          dinv0 = bik_rsqrt<TL>(d);
          Lp[(tri(i) + i) * W] = dinv0;
        }
        if (has1) {
          TL t1 = rhs1 ? TL(rhs[k]) : Lp[(tri(i + 1) + k) * W];
          t1 = (t1 - s01) * dinv0;
          r1[k] = t1;
          if (!rhs1) {
            Lp[(tri(i + 1) + k) * W] = t1;
            ss1 += t1 * t1;
            TL d = Lp[(tri(i + 1) + i + 1) * W] - ss1;
            if (!(d > TL(0))) { bad = 1; d = TL(1e-30); }
            Lp[(tri(i + 1) + i + 1) * W] = bik_rsqrt<TL>(d);
          }
        }
      }
    }
    if (rhs0) {
#pragma unroll
      for (int k = 0; k < N; ++k) y[k] = r0[k];
    } else if (rhs1) {
#pragma unroll
      for (int k = 0; k < N; ++k) y[k] = r1[k];
    }
  }
#pragma unroll
  for (int m = N - 1; m >= 0; --m) {
    const TL xm = y[m] * Lp[(tri(m) + m) * W];
    x[m] = TH(xm);
#pragma unroll
    for (int k = 0; k < m; ++k) y[k] -= Lp[(tri(m) + k) * W] * xm;
  }
  return bad;
}
// r <- (L L^T)^-1 r with the factor left in Lp by k2x_solve_masked (polishing step of the mixed-precision path)
template <typename TL, int N, int W>
BIK_HD void k2x_tri_solve(const TL* __restrict__ Lp, TL (&r)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) {
    TL s0 = r[k], s1 = TL(0);
#pragma unroll
    for (int m = 0; m < k; ++m) { if (m & 1) s1 -= Lp[(tri(k) + m) * W] * r[m]; else s0 -= Lp[(tri(k) + m) * W] * r[m]; }
    r[k] = (s0 + s1) * Lp[(tri(k) + k) * W];
  }
#pragma unroll
  for (int m = N - 1; m >= 0; --m) {
    const TL xm = r[m] * Lp[(tri(m) + m) * W];
    r[m] = xm;
#pragma unroll
    for (int k = 0; k < m; ++k) r[k] -= Lp[(tri(m) + k) * W] * xm;
  }
}
// g = H x + c, one pass over the scratch
template <typename TH, int N, int W>
BIK_HD void k2x_gradient(const TH* __restrict__ Hg, const TH* __restrict__ cg, const TH (&x)[N], TH (&g)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) g[k] = BIK_LDCG(cg + k * W);
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k <= i; ++k) {
      const TH hik = BIK_LDCG(Hg + (tri(i) + k) * W);
      g[i] += hik * x[k];
      if (k < i) g[k] += hik * x[i];
    }
  }
}

// One tile of W problems per warp.  TL == TH: everything in one precision.  TL = float, TH = double: mixed precision --
// the pivoting search runs on fp32 factorisations of the fp64 H; an iterate that passes the (loose) fp32 test is polished by
// one step of iterative refinement (fp64 residual = the multiplier pass, fp32 correction solve) and judged again with tight
// tolerances, and from then on every iterate is polished before it is judged.  Instances whose factorisation breaks down,
// whose refinement does not contract (cond(H) eps32 > 1e-3) or that hit the iteration cap are flagged in a.flag_out and left
// to the fp64 small-group kernel (launched behind this one on the flagged instances only).
template <typename TL, typename TH, int N, int W>
BIK_HD void k2x_warp_tile(const PView& P, const K2Args& a, long long b0, void* wsm, void* wscratch, int lane) {
  constexpr bool MIXED = sizeof(TL) < sizeof(TH);
  const PHeader& h = P.h();
  const int nv = h.nv, nu = h.nu, K = h.K, nq = h.nq, NP = h.P;
  const int32_t* cols = P.i(h.off_cols);
  const int32_t* umap = P.i(h.off_umap);
  const int32_t* ucols = P.i(h.off_ucols);
  const int cnt = (a.B - b0) < W ? (int)(a.B - b0) : W;
  const bool live = lane < cnt;
  const long long b = b0 + lane;
  TH* const U = reinterpret_cast<TH*>(wsm);
  TL* const Lp = reinterpret_cast<TL*>(wsm) + lane;
  TH* const Hg = reinterpret_cast<TH*>(wscratch) + lane;
  TH* const cg = Hg + (size_t)tri(N) * W;
  float* const log_ = reinterpret_cast<float*>(reinterpret_cast<TH*>(wscratch) + (size_t)(tri(N) + N) * W) + lane;
  float* const hig = log_ + (size_t)N * W;

  // ---- assembly into the scratch ----
#pragma unroll 4
  for (int k = 0; k < tri(N); ++k) Hg[k * W] = TH(0);
  for (int k = 0; k < N; ++k) { cg[k * W] = TH(0); log_[k * W] = -BIK_INF_F; hig[k * W] = BIK_INF_F; }
  for (int k = nu; k < N; ++k) Hg[(tri(k) + k) * W] = TH(1);   // padding dofs
  TH mu = TH(a.damping);
  const int St = k2t_task_tile_words(P);
  for (int t = 0; t < h.F + h.C; ++t) {
    int row0, nr, nc, coff; const float* cost; float gain, lm;
    if (t < h.F) { const FrameRec& fr = P.frame(t); row0 = fr.row0; nr = 6; nc = fr.ncols; coff = fr.col_off; cost = fr.cost; gain = fr.gain; lm = fr.lm; }
    else { const float* cr = P.f(h.off_com) + 8 * (t - h.F); row0 = reinterpret_cast<const int32_t*>(cr)[5]; nr = 3; nc = h.com_ncols; coff = h.com_cols_off; cost = cr; gain = cr[3]; lm = cr[4]; }
    BIK_SYNCWARP();
    k2t_stage_task<TH, W, W>(U, St, lane, cnt, b0, a, K, nv, cols + coff, row0, nr, nc, cost, gain);
    BIK_SYNCWARP();
    k2x_task_accumulate<TH, W>(U + (size_t)lane * St, nr, nc, cols + coff, umap, Hg, cg, lm, &mu);
  }
  BIK_SYNCWARP();
  float* const ft = reinterpret_cast<float*>(wsm);
  const int Sq = (nq + NP * nv) | 1;
  k2t_stage_rows<W, W>(ft, Sq, 0, lane, cnt, a.q + b0 * nq, nq, nq);
  if (NP > 0) k2t_stage_rows<W, W>(ft, Sq, nq, lane, cnt, a.ep + b0 * NP * nv, (long long)NP * nv, NP * nv);
  BIK_SYNCWARP();
  int st = 0;
  {
    const float* qrow = ft + (size_t)lane * Sq;
    const float* eprow = qrow + nq;
    for (int p = 0; p < NP; ++p) {
      const float* pr = P.f(h.off_posture) + p * (2 + nv);
      if (pr[1] != 0.f) {
        TH s = TH(0);
        for (int d = 0; d < nv; ++d) { TH v = TH(pr[2 + d]) * TH(pr[0]) * TH(eprow[p * nv + d]); s += v * v; }
        mu += TH(pr[1]) * s;
      }
    }
    for (int d = 0; d < nv; ++d) {
      const int u = umap[d];
      TH hd = mu, cd = TH(0);
      for (int p = 0; p < NP; ++p) {
        const float* pr = P.f(h.off_posture) + p * (2 + nv);
        const TH wgt = TH(pr[2 + d]);
        hd += wgt * wgt;
        cd -= TH(pr[0]) * wgt * wgt * TH(eprow[p * nv + d]);
      }
      float bl, bu;
      box_dof(P, d, qrow, a.dt, &bl, &bu);
      if (u >= 0) { Hg[(tri(u) + u) * W] += hd; cg[u * W] += cd; log_[u * W] = bl; hig[u * W] = bu; }
      else {
        TH v = -cd / hd;
        v = v < TH(bl) ? TH(bl) : (v > TH(bu) ? TH(bu) : v);
        if (!(v == v)) st |= 4;
        if (live) a.dq[b * nv + d] = float(v);
      }
    }
  }
  BIK_SYNCWARP();   // the tile is dead; the shared region becomes each lane's factor

  // ---- block principal pivoting (per thread; a lane that has converged waits for its warp) ----
  const int MAXIT = 60, PATIENCE = 3;
  uint32_t lom = 0u, upm = 0u;
  signed char* wm = (a.warm && live) ? a.warm + b * nu : nullptr;
  TH x[N];
  float lo[N], hi[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { lo[k] = log_[k * W]; hi[k] = hig[k * W]; x[k] = TH(0); }
  if (wm) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k < nu) {
        const int s0 = wm[k] & 3;
        if (s0 == 1 && lo[k] > -1e30f) lom |= 1u << k;
        else if (s0 == 2 && hi[k] < 1e30f) upm |= 1u << k;
      }
    }
  }
  int best = N + 1, patience = PATIENCE, it = 0, flag = 0;
  bool done = false, precise = false;
  for (; it < MAXIT && !done; ++it) {
    const uint32_t act = lom | upm;
    {
      TH xa[N];
#pragma unroll
      for (int k = 0; k < N; ++k) xa[k] = ((lom >> k) & 1u) ? TH(lo[k]) : (((upm >> k) & 1u) ? TH(hi[k]) : TH(0));
      if (k2x_solve_masked<TL, TH, N, W>(Hg, cg, Lp, act, xa, x)) { st |= 4; flag = 1; }
    }
    TH g[N];
    k2x_gradient<TH, N, W>(Hg, cg, x, g);   // multipliers (and, on free rows, minus the residual of the masked system)
    bool polished = false, need_polish = MIXED && precise;
    int ninf = 0, last = -1;
    uint32_t nlo = 0u, nup = 0u;
    for (;;) {
      if (need_polish) {   // one step of iterative refinement: wide residual, narrow correction solve
        TL r[N];
        TH mx = TH(0), md = TH(0);
#pragma unroll
        for (int k = 0; k < N; ++k) r[k] = ((act >> k) & 1u) ? TL(0) : TL(-g[k]);
        k2x_tri_solve<TL, N, W>(Lp, r);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          const TH dx = TH(r[k]), ax = x[k] < 0 ? -x[k] : x[k], ad = dx < 0 ? -dx : dx;
          x[k] += dx;
          mx = ax > mx ? ax : mx; md = ad > md ? ad : md;
        }
        if (md > TH(1e-3) * mx && md > TH(1e-7)) flag = 1;   // no contraction: this problem needs the wide factorisation
        k2x_gradient<TH, N, W>(Hg, cg, x, g);
        polished = true; need_polish = false;
      }
      const TH tolx = (MIXED && polished) ? TH(K2TolPolished::x()) : TH(K2Tol<TL>::x());
      const TH tolg = (MIXED && polished) ? TH(K2TolPolished::g()) : TH(K2Tol<TL>::g());
      ninf = 0; last = -1; nlo = 0u; nup = 0u;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const int cur = ((lom >> k) & 1u) ? 1 : (((upm >> k) & 1u) ? 2 : 0);
        int ns = cur;
        if (cur == 0) {
          const TH xi = x[k], bl = TH(lo[k]), bu = TH(hi[k]);
          if (xi < bl - tolx * (TH(1) + (bl < 0 ? -bl : bl))) ns = 1;
          else if (xi > bu + tolx * (TH(1) + (bu < 0 ? -bu : bu))) ns = 2;
        } else {
          if (cur == 1 && g[k] < -tolg) ns = 0;
          else if (cur == 2 && g[k] > tolg) ns = 0;
        }
        if (ns == 1) nlo |= 1u << k; else if (ns == 2) nup |= 1u << k;
        if (ns != cur) { ++ninf; last = k; }
      }
      if (MIXED && ninf == 0 && !polished) { need_polish = true; precise = true; continue; }
      break;
    }
    if (ninf == 0) { done = true; continue; }
    bool block;
    if (ninf < best) { best = ninf; patience = PATIENCE; block = true; }
    else if (patience > 0) { --patience; block = true; }
    else block = false;
    if (block) { lom = nlo; upm = nup; }
    else { const uint32_t bit = 1u << last; lom = (lom & ~bit) | (nlo & bit); upm = (upm & ~bit) | (nup & bit); }
  }
  if (!done) { st |= 2; flag = 1; }
  // ---- outputs ----
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k < nu) {
      const float v = float(x[k]);
      if (!(v == v)) { st |= 4; flag = 1; }
      if (live) a.dq[b * nv + ucols[k]] = v;
    }
  }
  if (!(MIXED && a.flag_out)) flag = 0;   // nobody behind us to pick flagged instances up: report what we have
  if (live) {
    if (MIXED && a.flag_out) a.flag_out[b] = flag;
    if (!flag) {
      if (a.status) a.status[b] |= st;
      if (a.iters) a.iters[b] = it;
      if (wm) {
#pragma unroll
        for (int k = 0; k < N; ++k) if (k < nu) wm[k] = (signed char)(((lom >> k) & 1u) ? 1 : (((upm >> k) & 1u) ? 2 : 0));
      }
    }
  }
  BIK_SYNCWARP();
}

}  // namespace bik
