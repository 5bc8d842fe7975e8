"""Degenerate and awkward grid shapes through every search entry point, against the CPU oracle: widths below 3 (two
NB_LUT directions alias one flat key), single rows / columns, sizes that are not multiples of the 64-cell tile or the
32-cell bit word, fully blocked and fully free maps."""
import numpy as np
import pytest

from conftest import same_cost
from nsfs import mapgen

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (2, 9), (9, 2), (3, 3), (5, 33), (33, 5), (63, 65), (65, 63), (129, 31)]


def _maps(H, W, seed):
    rng = np.random.RandomState(seed)
    occ = rng.rand(H, W) < 0.2
    yield "random", occ
    yield "free", np.zeros((H, W), bool)
    yield "blocked", np.ones((H, W), bool)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_small_and_ragged_shapes(gpu_lib, oracle_mod, shape):
    H, W = shape
    for name, occ in _maps(H, W, H * 100 + W):
        infl = occ.astype(np.int32)
        lo = np.where(occ, 20, 0).astype(np.int32)
        O = oracle_mod.Oracle(infl, lo)
        P = gpu_lib.Planner(H, W)
        P.set_map(infl, lo)
        cells = np.stack(np.meshgrid(np.arange(-1, H + 1), np.arange(-1, W + 1), indexing="ij"), -1).reshape(-1, 2)
        want = np.array([O.test_pos(i, j) for i, j in cells])
        got = P.test_pos(cells).astype(np.int64)
        assert np.array_equal(np.where(want < 0, -1, want), got), (shape, name)
        rng = np.random.RandomState(7)
        qs = np.stack([rng.randint(0, H, 12), rng.randint(0, W, 12), rng.randint(0, H, 12), rng.randint(0, W, 12)], 1).astype(np.int32)
        for algo, aid, mode in (("astar", 0, "f"), ("djikstra", 1, "g"), ("astar", 0, "fg")):
            b = P.plan_batch(algo, mode, qs, path_cap=H * W + 1)
            for qi, q in enumerate(qs):
                o = O.plan(aid, mode, q[:2], q[2:])
                assert b["status"][qi] == o["status"], (shape, name, algo, mode, qi)
                if o["status"] == 0:
                    n = len(o["path"])
                    assert b["path_len"][qi] == n and np.array_equal(b["paths"][qi][:n], o["path"]), (shape, name, algo, mode, qi)
                    assert same_cost(b["g_goal"][qi], o["g_goal"]) and b["settled"][qi] == o["settled"]
                s1 = P.plan(algo, mode, q[:2], q[2:])
                assert s1["status"] == o["status"] and (o["status"] != 0 or np.array_equal(s1["path"], o["path"]))
        src = (int(rng.randint(0, H)), int(rng.randint(0, W)))
        f, fo = P.field(src), O.field(src)
        assert f["status"] == fo["status"], (shape, name)
        if fo["status"] == 0:
            assert np.array_equal(f["g"] < 1e5, fo["g"] < 1e5)
            assert np.allclose(f["g"], fo["g"], rtol=1e-5, atol=1e-5)
        bad = np.argwhere(occ)[:6].astype(np.int32)
        if len(bad):
            fb = P.find_better_point_batch(bad)
            for t, c in enumerate(bad):
                o = O.find_better_point(c)
                assert fb["status"][t] == o["status"] and (o["status"] != 0 or tuple(fb["cells"][t]) == o["cell"]), (shape, name, t)
        segs = qs.copy()
        los = P.los_batch(segs)
        for l, sgm in zip(los, segs):
            o = O.is_los(sgm[:2], sgm[2:])
            assert int(l) == (o if o >= 0 else -1), (shape, name)
        P.close()
