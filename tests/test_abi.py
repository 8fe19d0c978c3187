"""The C-ABI library loads on a CPU-only box and exports every symbol include/bik.h declares."""

import os
import re

from mink_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    header = open(os.path.join(REPO, "include", "bik.h")).read()
    declared = set(re.findall(r"\b(bik_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libbik.so does not export {name}"
    assert declared == set(_lib.EXPORTS)
    assert lib.bik_version() == 200


def test_no_cpu_fallback():
    """Without a GPU, model creation must fail loudly instead of falling back to anything."""
    import ctypes as C

    import torch

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    blob = open(os.path.join(REPO, "mink_b200", "models", "ur5e.bikm"), "rb").read()
    h = C.c_void_p()
    rc = lib.bik_model_create(blob, len(blob), 0, C.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.bik_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "mink_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "ikoracle" not in src, f
