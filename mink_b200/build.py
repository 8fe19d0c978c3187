"""Build libbik.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Four translation units (K1, general K2, small-group K2, C ABI) are compiled in parallel and linked into
mink_b200/lib/libbik.so.  A unit is recompiled when it is older than any source it includes; `build()` returns a report
saying what was compiled, so that "did the build check build anything" is visible in the driver's log.
"""

from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import time

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libbik.so")
UNITS = ["bik_k1.cu", "bik_k2.cu", "bik_k2t.cu", "bik_api.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(PKG), "include", "bik.h")]


def _headers_mtime() -> float:
    return max(os.path.getmtime(s) for s in sources() if not s.endswith(".cu"))


def _obj(unit: str) -> str:
    return os.path.join(OBJDIR, unit[:-3] + ".o")


def _unit_stale(unit: str, hm: float) -> bool:
    o = _obj(unit)
    return not os.path.exists(o) or os.path.getmtime(o) < max(hm, os.path.getmtime(os.path.join(CSRC, unit)))


def is_stale() -> bool:
    hm = _headers_mtime()
    if any(_unit_stale(u, hm) for u in UNITS):
        return True
    return not os.path.exists(LIB) or any(os.path.getmtime(_obj(u)) > os.path.getmtime(LIB) for u in UNITS)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> dict:
    """Returns {"lib": path, "compiled": [units], "linked": bool, "seconds": s}."""
    t0 = time.time()
    os.makedirs(OBJDIR, exist_ok=True)
    hm = _headers_mtime()
    todo = [u for u in UNITS if force or _unit_stale(u, hm)]
    nvcc = _nvcc()

    def compile_unit(u):
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(u), os.path.join(CSRC, u)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {u}:\n{r.stderr[-4000:]}")
        return u, r.stderr

    logs = {}
    if todo:
        with cf.ThreadPoolExecutor(max_workers=len(todo)) as ex:
            for u, err in ex.map(compile_unit, todo):
                logs[u] = err
    link = bool(todo) or not os.path.exists(LIB) or any(os.path.getmtime(_obj(u)) > os.path.getmtime(LIB) for u in UNITS)
    if link:
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(u) for u in UNITS])
    rep = {"lib": LIB, "compiled": todo, "linked": link, "seconds": round(time.time() - t0, 1)}
    if verbose:
        for u, err in logs.items():
            sys.stderr.write(f"==== {u} ====\n{err}\n")
    return rep


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
