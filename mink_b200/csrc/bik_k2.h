// bik_k2.h -- K2: per-instance QP assembly and exact solve, one warp per problem.
//
//   H = damping I + sum_t (W_t J_t)^T (W_t J_t) + mu_t I,   c = sum_t -(W_t(-gain_t e_t))^T W_t J_t
//       (reference mink/tasks/task.py:105-138, mink/solve_ik.py:13-22)
//   min 1/2 dq^T H dq + c^T dq   s.t.  lo <= dq <= hi  [+ general rows G dq <= h]
//       (reference mink/solve_ik.py:43-65,101 -> qpsolvers)
//
// Structure exploited: frame-task Jacobians only have their ancestor-dof columns
// non-zero, so each task adds a small dense block to H; the posture task is diagonal; configuration
// and velocity limits share their index set and collapse to one box per dof.
//
// Solver: block principal pivoting (Kunisch-Rendl infeasible active set with the Judice-Pires
// single-pivot safeguard) -- every iteration guesses the active set, solves the reduced system with
// a fresh packed Cholesky of the free block and flips every violated index at once.  It terminates
// at the exact KKT point of the strictly convex QP, i.e. the same optimum quadprog/daqp return.
// General rows (collision) join the same pivoting through a Schur complement on the factor.
//
// Lanes own rows (row i -> lane i % W); reductions over lanes use shuffles.  W = 1 on the host.
#pragma once
#include "bik_k1.h"

#if defined(__CUDACC__)
#define BIK_NOINLINE __host__ __device__ __noinline__
#else
#define BIK_NOINLINE __attribute__((noinline))
#endif

// CTA-wide "does any warp still have work" vote (lock-step mode keeps the warps of a CTA in the same
// solver phase so that they share instruction-cache lines); identity on the host.
#if defined(__CUDA_ARCH__)
#define BIK_BLOCK_ANY(x) __syncthreads_or(x)
#else
#define BIK_BLOCK_ANY(x) (x)
#endif

namespace bik {

struct K2Args {
  int B;
  const float* q;    // [B][nq]
  const float* J;    // [B][K][nv]
  const float* e;    // [B][K]
  const float* ep;   // [B][P][nv]
  const float* Gc;   // [B][npairs][nv]
  const float* hc;   // [B][npairs]
  float dt;
  double damping;
  float* dq;         // [B][nv]
  int32_t* status;   // [B] or null (OR-ed into)
  int32_t* iters;    // [B] or null: active-set iterations (diagnostics)
  double* Hout;      // [B][nv][nv] or null  (bik_qp_objective)
  double* cout;      // [B][nv] or null
  float* lo_out;     // [B][nv] or null      (bik_limits_box)
  float* hi_out;
  int skip_objective;  // bik_limits_box: J/e/ep are not read
  int skip_box;        // bik_qp_objective: q is not read
  int lockstep;        // warps of a CTA advance through the pivoting iterations together (block barriers)
  signed char* warm;   // [B][nu] or null: active-set guess in (0 free, 1 lower, 2 upper; +4 = dq still holds the previous step's result), read at entry, updated at exit
  int32_t* flag_out;   // [B] or null: mixed-precision path marks the instances it could not finish (bik_k2x.h)
  const int32_t* only; // [B] or null: small-group path processes only the instances marked here (bik_k2t.h)
};

enum { K2_MAX_GEN = 16 };  // general (collision) rows that may be active at once

BIK_HD int tri(int i) { return (i * (i + 1)) >> 1; }

// per-warp scratch, in bytes, for scalar type of size `ts`.
// Layout: Hp | U | dinv c lo hi x | [general-row block] | ints.   U is a union: during assembly it
// holds the weighted Jacobian rows (fp32), afterwards the packed factor augmented by the rhs row.
BIK_HD int k2_union_bytes(const PHeader& h, int ts) {
  int n = h.nu;
  int lp = tri(n + 1) * ts + 16;                       // (n+1) rows: factor + fused right-hand side
  int wj = 4 * ((h.K > 0 ? h.K : 1) * (h.nv + 1)) + 16;  // wJ [K][nv] + we [K]
  return ((lp > wj ? lp : wj) + 15) & ~15;
}
BIK_HD int k2_warp_bytes(const PHeader& h, int ts) {
  int n = h.nu, np = h.npairs;
  int words_T = tri(n) + 5 * n + h.nv + (np > 0 ? (K2_MAX_GEN * n + K2_MAX_GEN * K2_MAX_GEN + 3 * K2_MAX_GEN + np) : 0);
  int bytes = ((words_T * ts + 15) & ~15) + k2_union_bytes(h, ts) + 4 * (2 * n + 3 * np + 12);
  return (bytes + 15) & ~15;
}

template <typename T> BIK_HD T bik_sqrt(T x);
template <> BIK_HD float bik_sqrt<float>(float x) { return sqrtf(x); }
template <> BIK_HD double bik_sqrt<double>(double x) { return sqrt(x); }
template <typename T> BIK_HD T bik_rsqrt(T x);
template <> BIK_HD float bik_rsqrt<float>(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
template <> BIK_HD double bik_rsqrt<double>(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

template <int W> BIK_HD int warp_sum_i(int v) {
#if defined(__CUDA_ARCH__)
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
template <int W> BIK_HD int warp_max_i(int v) {
#if defined(__CUDA_ARCH__)
  for (int o = W / 2; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
#endif
  return v;
}
template <typename T> BIK_HD T warp_bcast(T v, int src) {
#if defined(__CUDA_ARCH__)
  return __shfl_sync(0xffffffffu, v, src);
#else
  (void)src;
  return v;
#endif
}

template <typename T> struct K2Ws {
  T *Hp, *Lp, *dinv, *c, *lo, *hi, *x, *xfull;
  T *Y, *S, *lam, *rg, *hg, *sg;  // general rows: Y = L^-1 G_F^T (K2_MAX_GEN x n), S Schur, lam, rhs, h, slack
  int *st, *idx, *gst, *gidx, *gnew;
  float *wJ, *we;
};
template <typename T> BIK_HD K2Ws<T> k2_carve(const PHeader& h, void* mem) {
  K2Ws<T> w;
  int n = h.nu, np = h.npairs;
  char* base = reinterpret_cast<char*>(mem);
  T* p = reinterpret_cast<T*>(base);
  w.Hp = p; p += tri(n);
  w.dinv = p; p += n; w.c = p; p += n; w.lo = p; p += n; w.hi = p; p += n; w.x = p; p += n; w.xfull = p; p += h.nv;
  w.Y = w.S = w.lam = w.rg = w.hg = w.sg = nullptr;
  if (np > 0) { w.Y = p; p += K2_MAX_GEN * n; w.S = p; p += K2_MAX_GEN * K2_MAX_GEN; w.lam = p; p += K2_MAX_GEN; w.rg = p; p += K2_MAX_GEN; w.sg = p; p += K2_MAX_GEN; w.hg = p; p += np; }
  int words_T = (int)(p - reinterpret_cast<T*>(base));
  char* u = base + ((words_T * (int)sizeof(T) + 15) & ~15);
  w.Lp = reinterpret_cast<T*>(u);
  w.wJ = reinterpret_cast<float*>(u);
  w.we = w.wJ + (h.K > 0 ? h.K : 1) * h.nv;
  int* ip = reinterpret_cast<int*>(u + k2_union_bytes(h, sizeof(T)));
  w.st = ip; ip += n; w.idx = ip; ip += n; w.gst = ip; ip += np + 4; w.gidx = ip; ip += np + 4; w.gnew = ip; ip += np + 4;
  return w;
}

template <typename T> BIK_HD T Hsym(const T* Hp, int i, int j) { return i >= j ? Hp[tri(i) + j] : Hp[tri(j) + i]; }

// ---- assembly ------------------------------------------------------------------------------
template <typename T, int W>
BIK_HD void k2_assemble(const PView& P, const K2Args& a, int b, K2Ws<T>& w, int lane) {
  const PHeader& h = P.h();
  const int n = h.nv, nu = h.nu, K = h.K;
  const float* Jb = a.J + (long long)b * K * n;
  const float* eb = a.e + (long long)b * K;
  const int32_t* cols = P.i(h.off_cols);
  const int32_t* umap = P.i(h.off_umap);
  if (a.skip_objective) {  // bik_limits_box: only the box, for every dof
    for (int d = lane; d < n; d += W) {
      float lo, hi;
      box_dof(P, d, a.q + (long long)b * h.nq, a.dt, &lo, &hi);
      a.lo_out[(long long)b * n + d] = lo; a.hi_out[(long long)b * n + d] = hi;
    }
    return;
  }
  for (int k = lane; k < tri(nu); k += W) w.Hp[k] = T(0);
  // weighted rows  W J  and  W(-gain e)
  for (int f = 0; f < h.F; ++f) {
    const FrameRec& fr = P.frame(f);
    for (int k = lane; k < 6 * n; k += W) { int r = k / n; w.wJ[fr.row0 * n + k] = fr.cost[r] * Jb[fr.row0 * n + k]; }
    for (int r = lane; r < 6; r += W) w.we[fr.row0 + r] = fr.cost[r] * (-fr.gain * eb[fr.row0 + r]);
  }
  for (int c = 0; c < h.C; ++c) {
    const float* cr = P.f(h.off_com) + 8 * c;
    int row0 = reinterpret_cast<const int32_t*>(cr)[5];
    for (int k = lane; k < 3 * n; k += W) { int r = k / n; w.wJ[row0 * n + k] = cr[r] * Jb[row0 * n + k]; }
    for (int r = lane; r < 3; r += W) w.we[row0 + r] = cr[r] * (-cr[3] * eb[row0 + r]);
  }
  BIK_SYNCWARP();
  // block contributions (W J)^T (W J), lower triangle of the COUPLED block, per task over its non-zero columns
  for (int t = 0; t < h.F + h.C; ++t) {
    int row0, nr, nc, coff;
    if (t < h.F) { const FrameRec& fr = P.frame(t); row0 = fr.row0; nr = 6; nc = fr.ncols; coff = fr.col_off; }
    else { const float* cr = P.f(h.off_com) + 8 * (t - h.F); row0 = reinterpret_cast<const int32_t*>(cr)[5]; nr = 3; nc = h.com_ncols; coff = h.com_cols_off; }
    for (int p = lane; p < nc * nc; p += W) {
      int ia = p / nc, ib = p - ia * nc;
      if (ib > ia) continue;
      int ca = cols[coff + ia] & 0xffff, cb = cols[coff + ib] & 0xffff;
      T s = T(0);
      for (int r = 0; r < nr; ++r) s += T(w.wJ[(row0 + r) * n + ca]) * T(w.wJ[(row0 + r) * n + cb]);
      w.Hp[tri(umap[ca]) + umap[cb]] += s;   // column lists are ascending, so umap[ca] >= umap[cb]
    }
    BIK_SYNCWARP();
  }
  // Levenberg-Marquardt terms mu_t = lm_t ||W(-gain e)||^2 (task.py:131) -- every lane, same order
  T mu = T(a.damping);
  for (int f = 0; f < h.F; ++f) {
    const FrameRec& fr = P.frame(f);
    if (fr.lm != 0.f) { T s = T(0); for (int r = 0; r < 6; ++r) s += T(w.we[fr.row0 + r]) * T(w.we[fr.row0 + r]); mu += T(fr.lm) * s; }
  }
  for (int c = 0; c < h.C; ++c) {
    const float* cr = P.f(h.off_com) + 8 * c;
    int row0 = reinterpret_cast<const int32_t*>(cr)[5];
    if (cr[4] != 0.f) { T s = T(0); for (int r = 0; r < 3; ++r) s += T(w.we[row0 + r]) * T(w.we[row0 + r]); mu += T(cr[4]) * s; }
  }
  for (int p = 0; p < h.P; ++p) {
    const float* pr = P.f(h.off_posture) + p * (2 + n);
    if (pr[1] != 0.f) {
      const float* epb = a.ep + ((long long)b * h.P + p) * n;
      T s = T(0);
      for (int d = 0; d < n; ++d) { T v = T(pr[2 + d]) * T(pr[0]) * T(epb[d]); s += v * v; }
      mu += T(pr[1]) * s;
    }
  }
  // linear term, diagonal and box.  A DECOUPLED dof (zero column in every task Jacobian) only sees
  // the diagonal: its optimum is the clamp of -c/h, written straight to xfull.
  for (int d = lane; d < n; d += W) {
    const int u = umap[d];
    T cd = T(0);
    if (u >= 0) for (int r = 0; r < K; ++r) cd -= T(w.we[r]) * T(w.wJ[r * n + d]);
    T hd = mu;
    for (int p = 0; p < h.P; ++p) {
      const float* pr = P.f(h.off_posture) + p * (2 + n);
      T wgt = T(pr[2 + d]);
      hd += wgt * wgt;                                                        // (W J)^T (W J), J = -I
      cd -= T(pr[0]) * wgt * wgt * T(a.ep[((long long)b * h.P + p) * n + d]);  // -(W(-g e))^T W (-I)
    }
    float lo = -BIK_INF_F, hi = BIK_INF_F;
    if (!a.skip_box) box_dof(P, d, a.q + (long long)b * h.nq, a.dt, &lo, &hi);
    if (a.lo_out) { a.lo_out[(long long)b * n + d] = lo; a.hi_out[(long long)b * n + d] = hi; }
    if (u >= 0) {
      w.Hp[tri(u) + u] += hd;
      w.c[u] = cd; w.lo[u] = T(lo); w.hi[u] = T(hi);
    } else {
      T v = -cd / hd;
      v = v < T(lo) ? T(lo) : (v > T(hi) ? T(hi) : v);
      w.xfull[d] = v;
      if (a.Hout) { a.Hout[((long long)b * n + d) * n + d] = double(hd); a.cout[(long long)b * n + d] = double(cd); }
    }
  }
  BIK_SYNCWARP();
}

// ---- packed Cholesky of the free block, right-hand side fused as row nf -------------------------
// On entry Lp holds the lower triangle of the nf x nf block (compact indices) and, as row nf, the
// right-hand side y.  On exit row i (< nf) holds L[i][0..i-1] (the diagonal is kept as its inverse in
// dinv) and row nf holds L^-1 y: the forward substitution is just one more row of the same recurrence.
// Lane i % W owns row i; one warp barrier per column; no division or square root on the critical
// path (rsqrt once per column, by the lane that owns the next diagonal, from a running sum of squares).
template <typename T, int W, int SLOTS>
BIK_NOINLINE int k2_factor(T* __restrict__ Lp, T* __restrict__ dinv, int nf, int lane) {
  int bad = 0;
  T ss[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) ss[s] = T(0);
  if (lane == 0 && nf > 0) {
    T d = Lp[0];
    if (!(d > T(0))) { bad = 1; d = T(1e-30); }
    dinv[0] = bik_rsqrt<T>(d);
  }
  BIK_SYNCWARP();
#pragma unroll 1
  for (int j = 0; j < nf; ++j) {
    const T* Lj = Lp + tri(j);
    const T inv = dinv[j];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int i = lane + s * W;
      if (i > j && i <= nf) {
        T* Li = Lp + tri(i);
        T a0 = T(0), a1 = T(0);
        int k = 0;
#pragma unroll 2
        for (; k + 1 < j; k += 2) { a0 += Li[k] * Lj[k]; a1 += Li[k + 1] * Lj[k + 1]; }
        if (k < j) a0 += Li[k] * Lj[k];
        T l = (Li[j] - (a0 + a1)) * inv;
        Li[j] = l;
        ss[s] += l * l;
        if (i == j + 1 && i < nf) {
          T d = Li[i] - ss[s];
          if (!(d > T(0))) { bad = 1; d = T(1e-30); }
          dinv[i] = bik_rsqrt<T>(d);
        }
      }
    }
    BIK_SYNCWARP();
  }
  return bad;
}
// Back substitution x = L^-T y with y in row nf of Lp; each lane keeps its entries in registers and
// the pivot value travels by shuffle.  Writes x_k into out[k] (compact indices).
template <typename T, int W, int SLOTS>
BIK_NOINLINE void k2_backsub(const T* __restrict__ Lp, const T* __restrict__ dinv, int nf, T* out, int lane) {
  T y[SLOTS];
  const T* rhs = Lp + tri(nf);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) { int i = lane + s * W; y[s] = (i < nf) ? rhs[i] : T(0); }
  for (int k = nf - 1; k >= 0; --k) {
    const int ks = k / W, kl = k - ks * W;
    T cand = y[0];
#pragma unroll
    for (int s = 1; s < SLOTS; ++s) if (ks == s) cand = y[s];
    T xk = warp_bcast<T>(cand, kl) * dinv[k];
    if (lane == kl) out[k] = xk;
    const T* Lk = Lp + tri(k);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) { int i = lane + s * W; if (i < k) y[s] -= Lk[i] * xk; }
  }
  BIK_SYNCWARP();
}
// y <- L^-1 y for a separate vector (general-row path only)
template <typename T, int W>
BIK_NOINLINE void k2_forward(const T* Lp, const T* dinv, T* y, int nf, int lane) {
  for (int k = 0; k < nf; ++k) {
    if (k % W == lane) y[k] = y[k] * dinv[k];
    BIK_SYNCWARP();
    T yk = y[k];
    for (int i = lane; i < nf; i += W) if (i > k) y[i] -= Lp[tri(i) + k] * yk;
  }
  BIK_SYNCWARP();
}

// Active general rows R (collision): KKT  [H_FF G_RF^T; G_RF 0][x_F; lam] = [y; h_R - G_RA x_A] through the
// factor:  Y_r = L^-1 G_rF^T,  S = Y Y^T,  lam = S^-1 (Y (L^-1 y) - rhs),  rhs row <- L^-1 y - Y^T lam.
// Cold for box-only problems: kept out of line so it does not occupy the instruction cache.
template <typename T, int W>
BIK_NOINLINE void k2_general_rows(K2Ws<T>& w, const float* Gb, T* rhs, int n, int nf, int ng, int lane) {
  for (int r = 0; r < ng; ++r) {
    const float* Gr = Gb + (long long)w.gidx[r] * n;
    T* Yr = w.Y + r * n;
    for (int i = lane; i < nf; i += W) Yr[i] = T(Gr[w.idx[i]]);
    if (lane == 0) { T sv = w.hg[w.gidx[r]]; for (int j = 0; j < n; ++j) if (w.st[j]) sv -= T(Gr[j]) * w.x[j]; w.rg[r] = sv; }
    BIK_SYNCWARP();
    k2_forward<T, W>(w.Lp, w.dinv, Yr, nf, lane);
  }
  if (lane == 0) {  // tiny dense solve, serial
    for (int r = 0; r < ng; ++r) {
      for (int q2 = 0; q2 <= r; ++q2) { T v = T(0); for (int i = 0; i < nf; ++i) v += w.Y[r * n + i] * w.Y[q2 * n + i]; w.S[r * K2_MAX_GEN + q2] = v; w.S[q2 * K2_MAX_GEN + r] = v; }
      T v = -w.rg[r]; for (int i = 0; i < nf; ++i) v += w.Y[r * n + i] * rhs[i]; w.lam[r] = v;
    }
    for (int j = 0; j < ng; ++j) {  // Cholesky of S in place + solve
      T d = w.S[j * K2_MAX_GEN + j]; for (int k = 0; k < j; ++k) d -= w.S[j * K2_MAX_GEN + k] * w.S[j * K2_MAX_GEN + k];
      d = bik_sqrt<T>(d > T(0) ? d : T(1e-30)); w.S[j * K2_MAX_GEN + j] = d;
      for (int i = j + 1; i < ng; ++i) { T v = w.S[i * K2_MAX_GEN + j]; for (int k = 0; k < j; ++k) v -= w.S[i * K2_MAX_GEN + k] * w.S[j * K2_MAX_GEN + k]; w.S[i * K2_MAX_GEN + j] = v / d; }
    }
    for (int i = 0; i < ng; ++i) { T v = w.lam[i]; for (int k = 0; k < i; ++k) v -= w.S[i * K2_MAX_GEN + k] * w.lam[k]; w.lam[i] = v / w.S[i * K2_MAX_GEN + i]; }
    for (int i = ng - 1; i >= 0; --i) { T v = w.lam[i]; for (int k = i + 1; k < ng; ++k) v -= w.S[k * K2_MAX_GEN + i] * w.lam[k]; w.lam[i] = v / w.S[i * K2_MAX_GEN + i]; }
  }
  BIK_SYNCWARP();
  for (int i = lane; i < nf; i += W) { T v = rhs[i]; for (int r = 0; r < ng; ++r) v -= w.Y[r * n + i] * w.lam[r]; rhs[i] = v; }
  BIK_SYNCWARP();
}

template <typename T> struct K2Tol;
template <> struct K2Tol<double> { static BIK_HD double x() { return 1e-12; } static BIK_HD double g() { return 1e-9; } };
template <> struct K2Tol<float> { static BIK_HD float x() { return 1e-7f; } static BIK_HD float g() { return 1e-4f; } };

// Returns status bits.  On exit w.x holds dq.
template <typename T, int W, int SLOTS>
BIK_HD int k2_solve(const PView& P, const K2Args& a, int b, K2Ws<T>& w, int lane, int* iters_out, bool active = true) {
  const PHeader& h = P.h();
  const int n = h.nu, np = h.npairs;   // coupled dofs only (n == nv whenever there are general rows)
  const int MAXIT = 400, PATIENCE = 3;   // the single-pivot fallback is finite but slow: stalled instances of a rollout need up to ~150 pivots (was 60: 1-5 per thousand flagged)
  const T tolx = K2Tol<T>::x(), tolg = K2Tol<T>::g();
  const float* Gb = np > 0 ? a.Gc + (long long)b * np * n : nullptr;
  for (int i = lane; i < n; i += W) {
    int s0 = (a.warm && active) ? (a.warm[(long long)b * n + i] & 3) : 0;   // warm start: last step's active set (bit 2 is the small-group path's marker)
    if ((s0 == 1 && !(w.lo[i] > T(-1e30))) || (s0 == 2 && !(w.hi[i] < T(1e30)))) s0 = 0;
    w.st[i] = s0;
  }
  for (int r = lane; r < np; r += W) { w.gst[r] = 0; w.hg[r] = T(a.hc[(long long)b * np + r]); }
  BIK_SYNCWARP();
  int status = 0, best = n + np + 1, patience = PATIENCE, it = 0;
  bool done = !active || n == 0;
  for (;; ++it) {
    if (a.lockstep) { if (!BIK_BLOCK_ANY(!done && it < MAXIT)) break; }
    else if (done || it >= MAXIT) break;
    if (done || it >= MAXIT) continue;
    // compact free list / active general rows (every lane writes the same values)
    int nf = 0, ng = 0;
    for (int i = 0; i < n; ++i) if (w.st[i] == 0) { w.idx[nf] = i; ++nf; }
    for (int r = 0; r < np; ++r) if (w.gst[r]) { if (ng < K2_MAX_GEN) { w.gidx[ng] = r; ++ng; } }
    BIK_SYNCWARP();
    // x on the bounds, rhs of the reduced system, copy of H_FF
    for (int i = lane; i < n; i += W) w.x[i] = w.st[i] == 1 ? w.lo[i] : (w.st[i] == 2 ? w.hi[i] : T(0));
    BIK_SYNCWARP();
    T* rhs = w.Lp + tri(nf);
    for (int i = lane; i < nf; i += W) {
      int ii = w.idx[i];
      const T* Hrow = w.Hp + tri(ii);
      T r = -w.c[ii];
      for (int j = 0; j < ii; ++j) if (w.st[j]) r -= Hrow[j] * w.x[j];
      for (int j = ii + 1; j < n; ++j) if (w.st[j]) r -= w.Hp[tri(j) + ii] * w.x[j];
      rhs[i] = r;
      T* Li = w.Lp + tri(i);
      for (int j = 0; j <= i; ++j) Li[j] = Hrow[w.idx[j]];
    }
    BIK_SYNCWARP();
    if (k2_factor<T, W, SLOTS>(w.Lp, w.dinv, nf, lane)) status |= 4;
    status = warp_max_i<W>(status);
    if (ng > 0) k2_general_rows<T, W>(w, Gb, rhs, n, nf, ng, lane);
    // x_F = L^-T (.), overwriting the rhs row in place
    k2_backsub<T, W, SLOTS>(w.Lp, w.dinv, nf, rhs, lane);
    for (int i = lane; i < nf; i += W) w.x[w.idx[i]] = rhs[i];
    BIK_SYNCWARP();
    // gradient on the active bounds, feasibility of free variables and of general rows.
    // Proposed new states go to w.idx (free after the scatter above) and w.gnew.
    int ninf = 0, last = -1;
    for (int i = lane; i < n; i += W) {
      int cur = w.st[i], ns = cur;
      if (cur == 0) {
        T xi = w.x[i];
        if (xi < w.lo[i] - tolx * (T(1) + (w.lo[i] < 0 ? -w.lo[i] : w.lo[i]))) ns = 1;
        else if (xi > w.hi[i] + tolx * (T(1) + (w.hi[i] < 0 ? -w.hi[i] : w.hi[i]))) ns = 2;
      } else {
        T gi = w.c[i];
        const T* Hrow = w.Hp + tri(i);
        for (int j = 0; j <= i; ++j) gi += Hrow[j] * w.x[j];
        for (int j = i + 1; j < n; ++j) gi += w.Hp[tri(j) + i] * w.x[j];
        for (int r = 0; r < ng; ++r) gi += T(Gb[(long long)w.gidx[r] * n + i]) * w.lam[r];
        if (cur == 1 && gi < -tolg) ns = 0;
        else if (cur == 2 && gi > tolg) ns = 0;
      }
      w.idx[i] = ns;
      if (ns != cur) { ++ninf; last = i > last ? i : last; }
    }
    for (int r = lane; r < np; r += W) {
      int cur = w.gst[r], ns = cur;
      T hr = w.hg[r];
      if (!(hr < T(1e30))) ns = 0;  // inactive row: h = +inf (collision_avoidance_limit.py:192-199)
      else if (cur == 0) {
        T sv = -hr;
        for (int j = 0; j < n; ++j) sv += T(Gb[(long long)r * n + j]) * w.x[j];
        if (sv > tolx * (T(1) + (hr < 0 ? -hr : hr))) ns = 1;
      } else {
        int k = 0;
        while (k < ng && w.gidx[k] != r) ++k;
        if (k < ng && w.lam[k] < -tolg) ns = 0;
      }
      w.gnew[r] = ns;
      if (ns != cur) { ++ninf; last = n + r > last ? n + r : last; }
    }
    ninf = warp_sum_i<W>(ninf);
    last = warp_max_i<W>(last);
    if (ninf == 0) { done = true; continue; }
    bool block;
    if (ninf < best) { best = ninf; patience = PATIENCE; block = true; }
    else if (patience > 0) { --patience; block = true; }
    else block = false;
    BIK_SYNCWARP();
    for (int i = lane; i < n; i += W) if (block || i == last) w.st[i] = w.idx[i];
    for (int r = lane; r < np; r += W) if (block || n + r == last) w.gst[r] = w.gnew[r];
    BIK_SYNCWARP();
  }
  if (!done) status |= 2;
  if (a.warm && active) for (int i = lane; i < n; i += W) a.warm[(long long)b * n + i] = (signed char)w.st[i];
  if (iters_out) *iters_out = it;
  return status;
}

// One instance per warp: assemble, optionally dump (H, c) / (lo, hi), solve, write dq.
template <typename T, int W, int SLOTS>
BIK_HD void k2_warp(const PView& P, const K2Args& a, int b, void* wsm, int lane, bool active = true) {
  const PHeader& h = P.h();
  const int n = h.nv, nu = h.nu;
  const int32_t* umap = P.i(h.off_umap);
  K2Ws<T> w = k2_carve<T>(h, wsm);
  if (!active) {  // lock-step tail: no instance for this warp, but it must take part in the CTA votes
    if (a.dq && nu > 0) { int it = 0; k2_solve<T, W, SLOTS>(P, a, b, w, lane, &it, false); }
    return;
  }
  k2_assemble<T, W>(P, a, b, w, lane);
  if (a.skip_objective) return;
  if (a.Hout) {
    for (int k = lane; k < n * n; k += W) {
      int i = k / n, j = k - i * n, ui = umap[i], uj = umap[j];
      if (ui >= 0 && uj >= 0) a.Hout[(long long)b * n * n + k] = double(Hsym(w.Hp, ui, uj));
      else if (i != j) a.Hout[(long long)b * n * n + k] = 0.0;
    }
    for (int d = lane; d < n; d += W) if (umap[d] >= 0) a.cout[(long long)b * n + d] = double(w.c[umap[d]]);
  }
  if (!a.dq) return;
  int iters = 0, st = 0;
  if (nu > 0) st = k2_solve<T, W, SLOTS>(P, a, b, w, lane, &iters);
  for (int d = lane; d < n; d += W) {
    T v = umap[d] >= 0 ? w.x[umap[d]] : w.xfull[d];
    if (!(v == v)) st |= 4;
    a.dq[(long long)b * n + d] = float(v);
  }
  st = warp_max_i<W>(st & 2) | warp_max_i<W>(st & 4) | warp_max_i<W>(st & 8);
  if (lane == 0) {
    if (a.status) a.status[b] |= st;
    if (a.iters) a.iters[b] = iters;
  }
  BIK_SYNCWARP();
}

}  // namespace bik
