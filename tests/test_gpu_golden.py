"""CUDA path (through the C-ABI) against the golden vectors produced by the reference itself.

Bars: integer / index / byte results bit-exact; exact-order searches reproduce the reference's fp64 g bit for bit
(the kernel performs the same additions in the same order); the fp32 cost-to-go field within 1e-5 relative with
reachability exact (north_star tolerance)."""
import numpy as np
import pytest

from conftest import same_cost, COMBOS, GOLDEN_MAPS, golden_map, split_by

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planners(golden, gpu_lib):
    out = {}
    for m in GOLDEN_MAPS:
        infl, lo = golden_map(golden, m)
        P = gpu_lib.Planner(*infl.shape)
        P.set_map(infl, lo)
        out[m] = P
    yield out
    for P in out.values():
        P.close()


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_test_pos_every_cell(golden, planners, mname):
    infl, lo = golden_map(golden, mname)
    H, W = infl.shape
    ij = np.stack(np.meshgrid(np.arange(-1, H + 1), np.arange(-1, W + 1), indexing="ij"), -1).reshape(-1, 2)
    got = planners[mname].test_pos(ij)
    free = (infl <= 0) & (lo <= 10)
    for (i, j), g in zip(ij, got):
        key = i * W + j
        want = 0 if (i < 0 or i >= H) else (-1 if (key < 0 or key >= H * W) else int(free.reshape(-1)[key]))
        assert g == want, (i, j)


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: f"{c[0]}_{c[2]}")
def test_plan_batch_and_single(golden, planners, oracle_mod, mname, combo):
    aname, algo, mode = combo
    P = planners[mname]
    key = f"{mname}/{aname}_{mode}"
    qs = golden[f"{mname}/queries"]
    paths = split_by(golden[key + "/paths"], golden[key + "/path_len"])
    finals = split_by(golden[key + "/finals"], golden[key + "/final_len"])
    cap = int(golden[key + "/path_len"].max()) + 4
    b = P.plan_batch(aname, mode, qs, path_cap=cap)
    infl, lo = golden_map(golden, mname)
    O = oracle_mod.Oracle(infl, lo)      # only for sparsifyPath + idx2pos, which stay on the host in the C++ adapter too
    for qi, q in enumerate(qs):
        st = golden[key + "/status"][qi]
        assert b["status"][qi] == st, (qi, q)
        if st != 0:
            continue
        n = len(paths[qi])
        assert b["path_len"][qi] == n
        assert np.array_equal(b["paths"][qi][:n], paths[qi])
        assert same_cost(b["g_goal"][qi], golden[key + "/g_goal"][qi])
        assert b["settled"][qi] == golden[key + "/settled"][qi]
        # the single-query entry point gives the same answer
        if qi % 4 == 0:
            r = P.plan(aname, mode, q[:2], q[2:])
            assert r["status"] == 0 and np.array_equal(r["path"], paths[qi]) and r["g_goal"] == golden[key + "/g_goal"][qi]
        # post_process_path: sparsify on the host (double), any-angle shortcutting on the GPU
        if n > 2:
            sp = O.sparsify(O.idx2pos(paths[qi]))
            pts = np.array([_pos2idx(O, xy) for xy in sp], np.int32)
            rc, keep = P.any_angle(pts)
            assert rc == 0
            assert np.array_equal(sp[keep.astype(bool)], finals[qi])
        else:
            assert np.array_equal(O.idx2pos(paths[qi]), finals[qi])


def _pos2idx(O, xy):
    import ctypes as C
    i, j = C.c_int(0), C.c_int(0)
    O.lib.nso_pos2idx(C.byref(O.map), C.c_double(xy[0]), C.c_double(xy[1]), C.byref(i), C.byref(j))
    return (i.value, j.value)


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_field(golden, planners, mname):
    P = planners[mname]
    for t, s in enumerate(golden[f"{mname}/field_src"]):
        r = P.field(s)
        want = golden[f"{mname}/field{t}/status"][0]
        assert r["status"] == want
        if want != 0:
            continue
        g = golden[f"{mname}/field{t}/g"]
        reach = g < 1e5
        assert np.array_equal(r["g"] < 1e5, reach)                          # reachability: bit-exact
        assert r["reached"] == int(golden[f"{mname}/field{t}/visited"].sum())
        rel = np.abs(r["g"].astype(np.float64) - g)[reach] / np.maximum(g[reach], 1.0)
        assert rel.max() <= 1e-5                                            # cost: 1e-5 relative (north_star)
        assert (r["pred"][~reach] == 255).all()


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_field_plan_matches_djikstra_cost(golden, planners, mname):
    """NSFS_DJIKSTRA_FIELD: same reachability and cost as Djikstra::plan, path valid (maybe another optimum)."""
    P = planners[mname]
    key = f"{mname}/djikstra_g"
    infl, lo = golden_map(golden, mname)
    free = ((infl <= 0) & (lo <= 10))
    H, W = infl.shape
    for qi, q in enumerate(golden[f"{mname}/queries"]):
        if golden[key + "/status"][qi] != 0:
            continue                      # whether the reference throws before reaching the goal is order dependent
        r = P.plan("djikstra_field", "g", q[:2], q[2:])
        if r["status"] == -1:
            continue                      # the full field reaches a throwing cell the goal-directed search never expanded
        assert r["status"] == 0
        g_ref = golden[key + "/g_goal"][qi]
        assert (r["n_path"] == 1) == (golden[key + "/path_len"][qi] == 1)
        if g_ref < 1e5:
            assert abs(r["g_goal"] - g_ref) <= 1e-5 * max(g_ref, 1.0)
            p = r["path"]
            assert tuple(p[0]) == tuple(q[2:]) and tuple(p[-1]) == tuple(q[:2])
            keys = p[:, 0] * W + p[:, 1]
            d = keys[:-1] - keys[1:]                                        # child - parent = OFF[dir]
            assert np.isin(d, [W, W + 1, 1, -W + 1, -W, -W - 1, -1, W - 1]).all()
            assert free.reshape(-1)[keys[:-1]].all()


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_find_better_point(golden, planners, mname):
    P = planners[mname]
    r = P.find_better_point_batch(golden[f"{mname}/fb_in"])
    assert np.array_equal(r["status"], golden[f"{mname}/fb_status"])
    ok = r["status"] == 0
    assert np.array_equal(r["cells"][ok], golden[f"{mname}/fb_out"][ok])
    one = P.find_better_point(golden[f"{mname}/fb_in"][0])
    assert one["status"] == golden[f"{mname}/fb_status"][0]
    if one["status"] == 0:
        assert one["cell"] == tuple(golden[f"{mname}/fb_out"][0])


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_line_of_sight(golden, planners, mname):
    got = planners[mname].los_batch(golden[f"{mname}/segs"])
    assert np.array_equal(got.astype(np.int32), golden[f"{mname}/los"])
