"""The map producer's inflation step: reference loco_mapping keeps grid_inflation_ incrementally
(occupancy_grid.cpp:166-213); the restatement (oracle/nsfs_oracle.c nso_occ_inflation) and the CUDA kernel (csrc/inflate.cu)
compute it in bulk from the log-odds plane.  Golden planes come from the reference's own OccupancyGrid class run on
synthetic lidar scans and on raw update events (tests/golden/make_golden_occ.py).  Integer work: bit-exact."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLDEN_OCC = os.path.join(ROOT, "tests", "golden", "golden_occ_v1.npz")


@pytest.fixture(scope="module")
def gocc():
    return np.load(GOLDEN_OCC)


def test_mask_is_generateMask(oracle_mod, gocc):
    assert np.array_equal(oracle_mod.occ_mask(0.2, 0.05), gocc["mask"])
    assert len(gocc["mask"]) == 49                                   # the radius-4 disc a^2 + b^2 <= 16


@pytest.mark.parametrize("case", ["scan", "edge"])
def test_bulk_formula_reproduces_the_incremental_reference(oracle_mod, gocc, case):
    lo, want = gocc[f"{case}/lo"].astype(np.int32), gocc[f"{case}/infl"].astype(np.int32)
    rc, infl = oracle_mod.occ_inflation(lo, 10, 0.2, 0.05)
    assert rc == 0 and np.array_equal(infl, want)
    if case == "edge":
        # row 0 is "out of bounds" for updates, yet it IS written: a disc around (1, 0..3) spills over the column edge and
        # its flat keys land at the end of row 0 (the row guard looked at to_inflate.i = 1)
        assert (want[0, :-4] == 0).all() and want[0, -4:].any()


def test_reference_exception_is_reported(oracle_mod, gocc):
    rc, _ = oracle_mod.occ_inflation(gocc["throw/lo_at_throw"].astype(np.int32), 10, 0.2, 0.05)
    assert gocc["throw/rc"][0] == 1 and rc == oracle_mod.E_RANGE


def test_against_the_live_reference_on_random_event_streams(oracle_mod):
    if not oracle_mod.have_reference_occ():
        pytest.skip("oracle/_ref/libnsfs_ref_occ.so is only built where /root/reference exists")
    rng = np.random.RandomState(11)
    for trial in range(4):
        R = oracle_mod.ReferenceOcc(min_xy=(0.0, 0.0), max_xy=(3.6, 3.4), cell=0.05)       # 72 x 68
        ev_ij, ev_oc = [], []
        for _ in range(1500):
            c = (int(rng.randint(0, R.H - 5)), int(rng.randint(0, R.W)))                   # clear of the last rows: no throw
            n = int(rng.randint(1, 25))
            ev_ij += [c] * n
            ev_oc += [int(rng.rand() < 0.55)] * n
        assert R.update(np.array(ev_ij, np.int32), np.array(ev_oc, np.uint8)) == 0
        lo, infl = R.planes()
        rc, got = oracle_mod.occ_inflation(lo, 10, 0.2, 0.05)
        assert rc == 0 and np.array_equal(got, infl)


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["scan", "edge"])
def test_gpu_inflation_matches_the_reference_planes(gpu_lib, gocc, case):
    lo, want = gocc[f"{case}/lo"].astype(np.int32), gocc[f"{case}/infl"].astype(np.int32)
    P = gpu_lib.Planner(*lo.shape)
    assert P.set_map_logodds(lo, lo_thresh=10, occ_thresh=10) == 0
    assert np.array_equal(P.inflation(), want)
    P.close()


@pytest.mark.gpu
def test_gpu_inflation_random_planes_edges_and_exception(gpu_lib, oracle_mod, gocc):
    rng = np.random.RandomState(12)
    # the last three: discs of 149 / 709 cells in dense planes, so counts pass 63, 127, ... (a 6-plane counter would wrap there)
    for (H, W, dens, radius) in ((97, 131, 0.03, 0.2), (64, 64, 0.2, 0.2), (200, 90, 0.01, 0.35), (70, 300, 0.05, 0.1),
                                 (120, 150, 0.6, 0.35), (90, 200, 0.95, 0.35), (100, 128, 0.5, 0.75)):
        lo = np.where(rng.rand(H, W) < dens, 15, rng.randint(-20, 10, (H, W))).astype(np.int32)
        R = max(int(round(radius / 0.05)), 7)
        lo[H - 1, W - R - 1:] = 0                                    # keep clear of the throwing corner ...
        lo[H - R - 2:, W - R - 1:] = np.minimum(lo[H - R - 2:, W - R - 1:], 0)
        rc, want = oracle_mod.occ_inflation(lo, 10, radius, 0.05)
        P = gpu_lib.Planner(H, W)
        assert P.set_map_logodds(lo, 10, 10, radius, 0.05) == rc == 0
        assert np.array_equal(P.inflation(), want)
        lo[H - 1, W - 2] = 12                                        # ... then step into it
        rc, want = oracle_mod.occ_inflation(lo, 10, radius, 0.05)
        assert rc == oracle_mod.E_RANGE and P.set_map_logodds(lo, 10, 10, radius, 0.05) == gpu_lib.E_RANGE
        assert np.array_equal(P.inflation(), want)                   # the in-range part of the plane still agrees
        P.close()


@pytest.mark.gpu
def test_planner_sees_the_same_map_either_way(gpu_lib, oracle_mod, gocc):
    """Upload log-odds only (inflation on the device) vs upload both planes: identical planning results."""
    from nsfs import mapgen
    lo = gocc["scan/lo"].astype(np.int32)
    infl = gocc["scan/infl"].astype(np.int32)
    A = gpu_lib.Planner(*lo.shape); A.set_map_logodds(lo, 10, 10)
    B = gpu_lib.Planner(*lo.shape); B.set_map(infl, lo)
    qs = mapgen.sample_queries(infl, lo, 32, 5)
    pa, pb = A.plan_batch("astar", "f", qs, path_cap=512), B.plan_batch("astar", "f", qs, path_cap=512)
    assert np.array_equal(pa["paths"], pb["paths"]) and np.array_equal(pa["status"], pb["status"])
    bad = mapgen.sample_blocked_interior(infl, lo, 16, 5)
    assert np.array_equal(A.find_better_point_batch(bad)["cells"], B.find_better_point_batch(bad)["cells"])
    A.close(); B.close()
