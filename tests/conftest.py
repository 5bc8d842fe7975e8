"""Shared pytest plumbing: import paths, the ``gpu`` marker, golden fixtures and the CPU checkers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nav-stack-from-scratch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
GOLDEN_MAPS = ("m0", "m1", "m2", "m3")
COMBOS = (("astar", 0, "f"), ("astar", 0, "fg"), ("astar", 0, "g"), ("djikstra", 1, "g"), ("djikstra", 1, "f"))


def same_cost(a, b):
    """g(goal) from the BATCH kernel (search_team.cu): it keeps labels as exact (cardinal, diagonal) step counts and reports
    n_card + n_diag*sqrt(2), which differs from the reference's running fp64 sum by rounding only (< 1e-12 relative; the
    stated bar is 1e-5).  Single-query calls use the wide kernel, whose fp64 sums are the reference's own: those are
    compared with ``==``."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return bool(np.all(np.abs(a - b) <= 1e-12 * np.maximum(np.abs(b), 1.0)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(GOLDEN)


def golden_map(golden, name):
    return golden[f"{name}/infl"].astype(np.int32), golden[f"{name}/lo"].astype(np.int32)


def split_by(flat, lens):
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + n])
        o += n
    return out


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA C-ABI.  Fails (not skips) when the extension is missing: there is no fallback path."""
    from nsfs import capi
    capi.lib()
    return capi
