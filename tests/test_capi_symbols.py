"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/nsfs.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nsfs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsfs_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("nsfs_create", "nsfs_set_map", "nsfs_plan", "nsfs_plan_batch", "nsfs_field", "nsfs_field_path",
                 "nsfs_find_better_point", "nsfs_los_batch", "nsfs_any_angle", "nsfs_destroy", "nsfs_strerror"):
        assert must in syms


def test_library_exports_every_declared_symbol(gpu_lib):
    L = gpu_lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = gpu_lib.lib().nsfs_create(C.byref(h), 16, 16, 0)
    assert rc == gpu_lib.E_CUDA and not h.value
    with pytest.raises(gpu_lib.NsfsError):
        gpu_lib.Planner(16, 16)


def test_strerror_covers_codes(gpu_lib):
    L = gpu_lib.lib()
    for code in range(0, -9, -1):
        assert L.nsfs_strerror(code).decode() != "unknown nsfs error"


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "nav-stack-from-scratch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("oracle/ is", ""), os.path.join(dirpath, f)
