"""ORACLE / TEST INFRASTRUCTURE -- exact dense QP solve standing in for `qpsolvers`.

Not the product; never imported by `mink_b200/`.  Restates the third-party dependency
`qpsolvers[quadprog] >= 4.3.1` (reference pyproject.toml:29), whose source is not under
/root/reference.  Reference call sites: `qpsolvers.Problem(P, q, G, h)` (mink/solve_ik.py:65)
and `qpsolvers.solve_problem(problem, solver=solver, **kwargs)` (mink/solve_ik.py:101).

Algorithm: Goldfarb & Idnani (1983) dual active-set method for strictly convex QPs -- the
algorithm behind `quadprog` (the solver every reference test uses) -- in fp64 numpy:
start at the unconstrained minimiser, repeatedly add the most violated constraint, taking
partial steps that drop constraints whose multiplier reaches zero.  "daqp"/"osqp"/"quadprog"
all map to this one exact solver: the parity target is the exact QP optimum.

Pinned by: reference tests/test_solve_ik.py (convergence < 20 steps under active velocity
bounds) via oracle/run_reference_tests.py, and by KKT-residual + scipy BVLS cross-checks in
tests/test_oracle_qp.py.
"""

from __future__ import annotations

import numpy as np

available_solvers = ["quadprog", "daqp", "osqp", "oracle"]


class Problem:
    def __init__(self, P, q, G=None, h=None, A=None, b=None, lb=None, ub=None):
        self.P, self.q, self.G, self.h = P, q, G, h
        self.A, self.b, self.lb, self.ub = A, b, lb, ub

    def unpack(self):
        return self.P, self.q, self.G, self.h, self.A, self.b, self.lb, self.ub


class Solution:
    def __init__(self, problem, x, z=None, found=True, active=None, iterations=0):
        self.problem = problem
        self.x = x
        self.z = z
        self.found = found
        self.active = active
        self.iterations = iterations


def goldfarb_idnani(H, c, G=None, h=None, max_iter=None, tol=1e-11):
    """min 1/2 x'Hx + c'x  s.t.  Gx <= h.  Returns (x, multipliers, active list, iters) or (None,...)."""
    H = np.asarray(H, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    n = c.shape[0]
    Lc = np.linalg.cholesky(0.5 * (H + H.T))  # raises LinAlgError if not PD, as quadprog does

    def hsolve(b):
        y = np.linalg.solve(Lc, b)
        return np.linalg.solve(Lc.T, y)

    x = -hsolve(c)
    if G is None or len(h) == 0:
        return x, np.zeros(0), [], 0
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    keep = np.isfinite(h)  # rows with h = +inf can never be active (collision_avoidance_limit.py:192)
    idx_map = np.nonzero(keep)[0]
    G, h = G[keep], h[keep]
    m = h.shape[0]
    if m == 0:   # every row inactive (all collision pairs beyond the detection distance): the unconstrained minimiser
        return x, np.zeros(len(keep)), [], 0
    rown = np.maximum(np.linalg.norm(G, axis=1), 1e-300)
    active: list[int] = []
    u = np.zeros(0)
    iters = 0
    max_iter = max_iter or 50 * (m + n)
    while True:
        slack = h - G @ x
        viol = slack / rown
        if active:
            viol[active] = np.inf
        p = int(np.argmin(viol))
        if viol[p] >= -tol:
            lam = np.zeros(len(keep))
            lam[idx_map[active]] = u
            return x, lam, [int(idx_map[a]) for a in active], iters
        n_p = -G[p]
        u_plus = 0.0
        while True:
            iters += 1
            if iters > max_iter:
                return None, None, None, iters
            hinv_np = hsolve(n_p)
            if active:
                N = -G[active].T
                hinv_N = hsolve(N)
                M = N.T @ hinv_N
                r = np.linalg.lstsq(M, hinv_N.T @ n_p, rcond=None)[0]
                z = hinv_np - hinv_N @ r
            else:
                r = np.zeros(0)
                z = hinv_np
            # Partial step length: largest t keeping all multipliers nonnegative.
            t1, k = np.inf, -1
            for j in range(len(active)):
                if r[j] > 1e-13:
                    cand = u[j] / r[j]
                    if cand < t1:
                        t1, k = cand, j
            zn = float(z @ n_p)
            s_p = float(h[p] - G[p] @ x)
            if zn > 1e-13 * max(1.0, float(n_p @ hinv_np)):
                t2 = -s_p / zn
            else:
                t2 = np.inf
            t = min(t1, t2)
            if not np.isfinite(t):
                return None, None, None, iters  # infeasible
            if not np.isfinite(t2):
                u = u - t * r
                u_plus += t
                u = np.delete(u, k)
                active.pop(k)
                continue
            x = x + t * z
            u = u - t * r
            u_plus += t
            if t == t2:
                active.append(p)
                u = np.append(u, u_plus)
                break
            u = np.delete(u, k)
            active.pop(k)


def solve_problem(problem: Problem, solver: str = "quadprog", initvals=None, verbose=False, **kwargs) -> Solution:
    P, q, G, h = problem.P, problem.q, problem.G, problem.h
    if problem.A is not None or problem.lb is not None or problem.ub is not None:
        raise NotImplementedError("oracle qpsolvers: only (P, q, G, h) problems are restated")
    x, lam, active, iters = goldfarb_idnani(P, q, G, h)
    return Solution(problem, x, z=lam, found=x is not None, active=active, iterations=iters)


def solve_qp(P, q, G=None, h=None, A=None, b=None, lb=None, ub=None, solver="quadprog", **kwargs):
    return solve_problem(Problem(P, q, G, h, A, b, lb, ub), solver=solver, **kwargs).x
