"""Shared test helpers: load golden fixtures + flattened models (no /root/reference needed)."""

import os

import numpy as np

from mink_b200._abi import spec_from_workload
from mink_b200.workloads import WORKLOADS, load_flat, task_frames  # noqa: F401  (re-exported for the tests)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name: str):
    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return wl, fm, spec, g


def quat_align(a, b):
    """Flip sign of quaternions in a (...,4) to match b's hemisphere."""
    s = np.sign(np.sum(a * b, axis=-1, keepdims=True))
    s[s == 0] = 1
    return a * s
