"""The path consumer's safety check: reference Commander::checkTrajectorySafety with its own pos2idx /
checkCell / oob (commander.cpp:316-402).  Golden results come from RUNNING the reference's Commander class, compiled
unmodified (tests/golden/make_golden_cmd.py); the restatement (oracle/nsfs_oracle.c nso_cmd_check_trajectory), the CUDA path
(nsfs_check_trajectory through the C-ABI) and the C++ host mirror (GpuPlannerBase::checkTrajectorySafety) must return the same
(safe, first unsafe index) and the same std::out_of_range cases.  Integer / index work: exact."""
import os

import numpy as np
import pytest

from conftest import ROOT
from nsfs import mapgen

GOLDEN_CMD = os.path.join(ROOT, "tests", "golden", "golden_cmd_v1.npz")
CELL = 0.05


@pytest.fixture(scope="module")
def gcmd():
    return np.load(GOLDEN_CMD)


def _cases(g):
    maps = {}
    for k in range(int(g["n_cases"][0])):
        H, W, seed, inflated, frame = (int(v) for v in g[f"c{k}_map"])
        p, ox, oy = (float(v) for v in g[f"c{k}_p"])
        key = (H, W, p, seed, inflated, frame)
        if key not in maps:
            maps[key] = mapgen.synth_map(H, W, p, seed, bool(inflated), bool(frame))
        t_id, rc, safe, bad = (int(v) for v in g[f"c{k}_out"])
        yield k, key, maps[key], (ox, oy), g[f"c{k}_xy"], g[f"c{k}_ij"].astype(np.int32), t_id, rc, bool(safe), bad


def test_golden_set_covers_every_outcome(gcmd):
    out = np.array([gcmd[f"c{k}_out"] for k in range(int(gcmd["n_cases"][0]))])
    assert (out[:, 1] != 0).sum() >= 5                       # the reference threw std::out_of_range
    assert ((out[:, 1] == 0) & (out[:, 2] == 1)).sum() >= 10  # safe
    assert ((out[:, 1] == 0) & (out[:, 3] == -2)).sum() >= 1  # empty trajectory
    assert ((out[:, 1] == 0) & (out[:, 2] == 0) & (out[:, 3] >= 0)).sum() >= 30


def test_restatement_matches_the_reference_commander(oracle_mod, gcmd):
    for k, _, (infl, lo), origin, xy, ij, t_id, rc, safe, bad in _cases(gcmd):
        O = oracle_mod.Oracle(infl, lo, cell=CELL, origin=origin)
        if len(xy):
            assert np.array_equal(O.pos2idx(xy), ij), k           # Commander::pos2idx == core's rounding
        got = O.cmd_check_trajectory(ij, t_id)
        assert got[0] == (oracle_mod.E_RANGE if rc else 0), k
        if rc == 0:
            assert got[1:] == (safe, bad), (k, got, safe, bad)


def test_restatement_against_the_live_reference(oracle_mod):
    if not oracle_mod.have_reference_cmd():
        pytest.skip("oracle/_ref/libnsfs_ref_cmd.so is only built where /root/reference exists")
    rng = np.random.default_rng(7)
    infl, lo = mapgen.synth_map(80, 70, 0.08, 41, True, False)
    O = oracle_mod.Oracle(infl, lo, cell=CELL, origin=(1.0, -2.0))
    for t in range(200):
        n = int(rng.integers(0, 60))
        ij = np.stack([rng.integers(-3, 84, n), rng.integers(-80, 150, n)], 1).astype(np.int32)
        xy = ij * CELL + np.array([1.0, -2.0]) + rng.uniform(-0.02, 0.02, (n, 2))
        t_id = int(rng.integers(-2, n + 2))
        want = oracle_mod.cmd_check_trajectory_reference(infl, lo, xy, t_id, cell=CELL, origin=(1.0, -2.0))
        got = O.cmd_check_trajectory(O.pos2idx(xy) if n else ij, t_id)
        assert got[0] == want[0], t
        if want[0] == 0:
            assert got[1:] == want[1:], (t, got, want)


@pytest.mark.gpu
def test_cuda_check_matches_the_reference_commander(gpu_lib, gcmd):
    planners = {}
    for k, key, (infl, lo), origin, xy, ij, t_id, rc, safe, bad in _cases(gcmd):
        if key not in planners:
            planners[key] = gpu_lib.Planner(*infl.shape)
            planners[key].set_map(infl, lo)
        got = planners[key].check_trajectory(ij, t_id)
        assert got[0] == (gpu_lib.E_RANGE if rc else 0), (k, got)
        if rc == 0:
            assert got[1:] == (safe, bad), (k, got, safe, bad)
    for P in planners.values():
        P.close()


@pytest.mark.gpu
def test_host_mirror_takes_world_points(gcmd):
    from nsfs import host
    sessions = {}
    for k, key, (infl, lo), origin, xy, ij, t_id, rc, safe, bad in _cases(gcmd):
        skey = key + origin
        if skey not in sessions:
            sessions[skey] = host.Session("astar", "f", infl, lo, cell=CELL, origin=origin)
        got = sessions[skey].check_trajectory(xy, t_id)
        if rc:
            assert got["status"] == "out_of_range", k
        else:
            assert got["status"] == "ok" and (got["safe"], got["bad_idx"]) == (safe, bad), (k, got, safe, bad)
    for S in sessions.values():
        S.close()


@pytest.mark.gpu
def test_long_trajectory_first_unsafe_point_from_the_end(gpu_lib, oracle_mod):
    """200 000 points (many blocks): the scan runs from t_id down, so the LARGEST unsafe index at or below t_id is reported."""
    infl, lo = mapgen.synth_map(512, 512, 0.0, 3)
    P = gpu_lib.Planner(512, 512)
    P.set_map(infl, lo)
    O = oracle_mod.Oracle(infl, lo)
    rng = np.random.default_rng(5)
    n = 200000
    ij = np.stack([rng.integers(1, 511, n), rng.integers(1, 511, n)], 1).astype(np.int32)      # the frame is occupied, the inside is free
    assert P.check_trajectory(ij, n - 1) == (0, True, -1)
    for bad_at in (0, 31, 32, 255, 256, 77777, n - 1):
        q = ij.copy()
        q[bad_at] = (0, 5)                                   # row 0: outside for the commander
        q[max(bad_at - 1000, 0)] = (511, 5)                  # an earlier (lower) unsafe point must not win
        for t_id in (n - 1, bad_at, bad_at - 1, bad_at + 1):
            if 0 <= t_id < n:
                got = P.check_trajectory(q, t_id)
                assert got == O.cmd_check_trajectory(q, t_id), (bad_at, t_id, got)
    assert P.check_trajectory(ij, n + 5)[0] == gpu_lib.E_RANGE          # trajectory_.at(t_id) past the end
    P.close()
