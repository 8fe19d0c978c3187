"""ORACLE PINNING -- the QP solver the oracle stands on (oracle/qpsolvers: Goldfarb-Idnani, the algorithm of quadprog) and the
golden dq it produced, checked against two things neither of them produced:

  * scipy.optimize.lsq_linear(method="bvls") -- an exact active-set solver for box-constrained least squares -- on every
    box-only configuration (the QP 1/2 x'Hx + c'x, lo <= x <= hi is the BVLS problem ||L'x + L^-1 c||^2 with H = L L');
  * explicit KKT residuals (stationarity, primal / dual feasibility, complementarity <= 1e-9) on the configurations with
    general rows (collision avoidance), where no second solver is available offline.
"""

import os
import sys

import numpy as np
import pytest
from scipy.optimize import lsq_linear

from tests.helpers import load_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from qpsolvers import goldfarb_idnani  # noqa: E402  (oracle shim; test infrastructure)

BOX_CASES = ["ur5e", "g1", "shadow", "g1_hands", "g1_full", "g1_rel", "ur5e_damp", "iiwa", "h1", "go1", "stretch", "tidybot", "aloha", "leap"]
ROW_CASES = ["spot", "edge"]


def _bvls(H, c, lo, hi):
    L = np.linalg.cholesky(H)
    b = -np.linalg.solve(L, c)
    r = lsq_linear(L.T, b, bounds=(lo, hi), method="bvls", tol=1e-15, max_iter=10000)
    assert r.status > 0 or r.success
    return r.x


@pytest.mark.parametrize("name", BOX_CASES)
def test_golden_dq_is_the_bvls_optimum(name):
    wl, fm, spec, g = load_case(name)
    worst = 0.0
    for b in range(g["q"].shape[0]):
        H, c = g["H"][b], g["c"][b]
        lo = np.where(np.isfinite(g["box_lo"][b]), g["box_lo"][b], -np.inf)
        hi = np.where(np.isfinite(g["box_hi"][b]), g["box_hi"][b], np.inf)
        x = _bvls(H, c, lo, hi)
        # objective values agree to rounding and the minimisers to 1e-9 (H is well conditioned on these configurations)
        f = lambda v: 0.5 * v @ H @ v + c @ v
        assert f(g["dq"][b]) <= f(x) + 1e-12 * max(1.0, abs(f(x)))
        worst = max(worst, np.abs(x - g["dq"][b]).max())
    print(name, "max |dq_golden - dq_bvls| =", worst)
    assert worst < 1e-8


@pytest.mark.parametrize("name", BOX_CASES)
def test_goldfarb_idnani_on_the_box_rows_matches_bvls(name):
    """The shim fed with the box written as general rows G = [+P; -P] (what mink hands to qpsolvers, solve_ik.py:25-40)."""
    wl, fm, spec, g = load_case(name)
    if "G" not in g:
        pytest.skip("no inequality rows in this golden")
    for b in range(min(6, g["q"].shape[0])):
        x, lam, active, iters = goldfarb_idnani(g["H"][b], g["c"][b], g["G"][b], g["h"][b])
        assert x is not None
        lo = np.where(np.isfinite(g["box_lo"][b]), g["box_lo"][b], -np.inf)
        hi = np.where(np.isfinite(g["box_hi"][b]), g["box_hi"][b], np.inf)
        np.testing.assert_allclose(x, _bvls(g["H"][b], g["c"][b], lo, hi), atol=1e-8)
        assert (lam >= -1e-10).all()


@pytest.mark.parametrize("name", ROW_CASES + ["g1", "shadow"])
def test_kkt_residuals_of_the_golden_solutions(name):
    """min 1/2 x'Px + q'x, Gx <= h at x = dq_golden: multipliers from the active rows by least squares, then
    stationarity |Px + q + G'lam|, primal feasibility max(Gx - h)+, dual feasibility min(lam)-, complementarity |lam (Gx - h)|."""
    wl, fm, spec, g = load_case(name)
    worst = dict(stat=0.0, prim=0.0, dual=0.0, comp=0.0)
    for b in range(g["q"].shape[0]):
        P, q, G, h, x = g["H"][b], g["c"][b], g["G"][b], g["h"][b], g["dq"][b]
        fin = np.isfinite(h)
        G, h = G[fin], h[fin]
        slack = h - G @ x
        scale = max(1.0, np.abs(P @ x).max(), np.abs(q).max())
        act = slack < 1e-9 * (1 + np.abs(h))
        lam = np.zeros(len(h))
        grad = P @ x + q
        if act.any():
            lam[act] = np.linalg.lstsq(G[act].T, -grad, rcond=None)[0]
        worst["stat"] = max(worst["stat"], np.abs(grad + G.T @ lam).max() / scale)
        worst["prim"] = max(worst["prim"], max(0.0, (-slack).max()))
        worst["dual"] = max(worst["dual"], max(0.0, (-lam).max()) / scale)
        worst["comp"] = max(worst["comp"], np.abs(lam * slack).max() / scale)
    print(name, worst)
    assert worst["stat"] < 1e-9 and worst["prim"] < 1e-9 and worst["dual"] < 1e-9 and worst["comp"] < 1e-9


def test_shim_returns_none_on_an_infeasible_problem_and_unconstrained_minimiser_without_rows():
    H = np.diag([2.0, 3.0])
    c = np.array([-1.0, 1.0])
    x, *_ = goldfarb_idnani(H, c, None, None)
    np.testing.assert_allclose(x, [0.5, -1.0 / 3.0])
    # x <= -1 and x >= 1 at once
    G = np.array([[1.0, 0.0], [-1.0, 0.0]])
    h = np.array([-1.0, -1.0])
    x, *_ = goldfarb_idnani(H, c, G, h)
    assert x is None
    # every row inactive (h = +inf): the unconstrained minimiser again (collision rows beyond the detection distance)
    x, lam, active, _ = goldfarb_idnani(H, c, G, np.array([np.inf, np.inf]))
    np.testing.assert_allclose(x, [0.5, -1.0 / 3.0])
    assert active == []
