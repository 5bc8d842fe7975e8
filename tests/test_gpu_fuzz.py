"""Seeded differential fuzzing: random map shapes, densities, frames, cost modes and planners; CUDA (batch kernel and,
for a subset, the one-query-per-warp kernel) against the CPU oracle on every query: status, settled cells, full path."""
import numpy as np
import pytest

from conftest import same_cost
from nsfs import mapgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(gpu_lib, oracle_mod, seed):
    rng = np.random.RandomState(1000 + seed)
    H, W = int(rng.randint(8, 180)), int(rng.randint(8, 180))
    p = float(rng.choice([0.0, 0.02, 0.1, 0.25, 0.4]))
    inflated = bool(rng.rand() < 0.3) and p <= 0.02
    frame = bool(rng.rand() < 0.7)
    infl, lo = mapgen.synth_map(H, W, p, 2000 + seed, inflated, frame)
    algo, aid = [("astar", 0), ("djikstra", 1)][int(rng.randint(2))]
    mode = ["f", "g", "fg"][int(rng.randint(3))]
    free = np.argwhere(mapgen.free_mask(infl, lo))
    if len(free) < 2:
        pytest.skip("no free cells")
    nq = 40
    qs = np.concatenate([free[rng.randint(len(free), size=nq)], free[rng.randint(len(free), size=nq)]], 1).astype(np.int32)
    qs[0, 2:] = qs[0, :2]
    blocked = np.argwhere(~mapgen.free_mask(infl, lo))
    if len(blocked):
        qs[1, 2:] = blocked[rng.randint(len(blocked))]                     # unreachable goal
        qs[2, :2] = blocked[rng.randint(len(blocked))]                     # blocked start (plan() pushes it regardless)
    O = oracle_mod.Oracle(infl, lo)
    P = gpu_lib.Planner(H, W)
    P.set_map(infl, lo)
    b = P.plan_batch(algo, mode, qs, path_cap=H * W + 1)
    P.set_option("search_kernel", 1)
    w = P.plan_batch(algo, mode, qs[:12], path_cap=H * W + 1)
    for qi, q in enumerate(qs):
        o = O.plan(aid, mode, q[:2], q[2:])
        ctx = (seed, H, W, p, inflated, frame, algo, mode, qi)
        assert b["status"][qi] == o["status"], ctx
        if o["status"] == 0:
            n = len(o["path"])
            assert b["settled"][qi] == o["settled"] and b["path_len"][qi] == n, ctx
            assert np.array_equal(b["paths"][qi][:n], o["path"]) and same_cost(b["g_goal"][qi], o["g_goal"]), ctx
        if qi < 12:
            assert w["status"][qi] == o["status"], ctx
            if o["status"] == 0:
                assert w["g_goal"][qi] == o["g_goal"] and np.array_equal(w["paths"][qi][:len(o["path"])], o["path"]), ctx
    P.close()
