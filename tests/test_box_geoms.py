"""Box geoms in the collision rows (capsule / sphere / plane against a box).

Three independent restatements are compared: the numpy shim's closed-form piecewise minimiser (what the reference ran on when
the golden `ur5e_wall` was generated), the C oracle's port of it, and the device code's bisection (host emulation here, the
kernel itself in test_gpu_parity / test_gpu_api).
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle import ikoracle  # noqa: E402
from helpers import load_case  # noqa: E402
from emu_lib import Emu  # noqa: E402


def test_shim_segment_box_minimiser_against_dense_sampling():
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    try:
        from mujoco import _seg_box_param   # the oracle shim, not MuJoCo
    finally:
        sys.path.pop(0)
    rng = np.random.default_rng(5)
    ts = np.linspace(-1, 1, 4001)
    for i in range(400):
        s = rng.uniform(0.05, 0.5, 3); c = rng.normal(0, 0.6, 3); d = rng.normal(0, 0.4, 3)
        if i % 5 == 0:
            d[rng.integers(3)] = 0.0
        t = _seg_box_param(c, d, s)
        assert -1.0 <= t <= 1.0
        e = np.maximum(np.abs(c + t * d) - s, 0.0)
        E = np.maximum(np.abs(c[None] + ts[:, None] * d[None]) - s, 0.0)
        assert e @ e <= (E * E).sum(1).min() + 1e-15


def test_device_box_rows_match_the_c_oracle_on_random_configurations():
    wl, fm, spec, g = load_case("ur5e_wall")
    blob = fm.to_blob()
    emu = Emu(blob, spec, fm.nq, fm.nv)
    orc = ikoracle.Oracle(blob, spec, fm.nq, fm.nv)
    rng = np.random.default_rng(11)
    B = 3000
    q = rng.uniform(-np.pi, np.pi, (B, fm.nq))
    dt = float(g["dt"])
    Gr, hr = orc.collision(q, dt)
    out = emu.fk_jac(q, np.tile(g["frame_targets"][:1], (B, 1, 1)), g["posture_target"], None, dt=dt, prec="f64")
    Gc, hc = out[3], out[4]
    fin = np.isfinite(hr)
    assert fin[:, 1].sum() > 300 and (hr[:, 1] == 0).sum() > 15     # wall rows: active ones, and penetrating ones
    assert np.array_equal(np.isfinite(hc), fin)
    np.testing.assert_allclose(hc[fin], hr[fin], rtol=1e-9, atol=1e-9)
    # the witness point is non-unique when the wrist axis is parallel to a face at equal distance; none of the samples is
    np.testing.assert_allclose(Gc, Gr, atol=1e-7)
    out32 = emu.fk_jac(q, np.tile(g["frame_targets"][:1], (B, 1, 1)), g["posture_target"], None, dt=dt, prec="f32")
    fin32 = np.isfinite(out32[4])
    same = fin32 == fin   # a pair exactly at the detection distance may flip in fp32
    assert same.mean() > 0.999
    both = (fin & fin32).all(axis=1)
    assert np.abs(out32[3][both] - Gr[both]).max() < 2e-3     # fp32 witness points near face edges
    assert np.median(np.abs(out32[3][both] - Gr[both]).max(axis=(1, 2))) < 2e-6
