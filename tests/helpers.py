"""Shared test helpers: load golden fixtures + flattened models (no /root/reference needed)."""

import os

import numpy as np

from mink_b200._abi import spec_from_workload
from mink_b200.flatten import FlatModel
from mink_b200.workloads import WORKLOADS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_flat(robot: str) -> FlatModel:
    with open(os.path.join(GOLDEN, "models", robot + ".bikm"), "rb") as f:
        blob = f.read()
    with open(os.path.join(GOLDEN, "models", robot + ".json")) as f:
        meta = f.read()
    return FlatModel.from_blob(blob, meta)


def load_case(name: str):
    wl = WORKLOADS[name]
    fm = load_flat(wl["robot"])
    spec = spec_from_workload(fm, wl)
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return wl, fm, spec, g


def task_frames(wl, fm):
    return [fm.frame(f["name"], f["type"]) for f in wl["frames"]]


def quat_align(a, b):
    """Flip sign of quaternions in a (...,4) to match b's hemisphere."""
    s = np.sign(np.sum(a * b, axis=-1, keepdims=True))
    s[s == 0] = 1
    return a * s
