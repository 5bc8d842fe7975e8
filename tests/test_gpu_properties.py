"""Size-independent properties at BASELINE.json's full sizes (the CPU oracle is too slow to replay these fully)."""
import numpy as np
import pytest

from conftest import same_cost
from nsfs import mapgen

pytestmark = pytest.mark.gpu

SQRT2F = np.float32(1.41421356237309504880)


def test_c2_field_2048_is_a_fixed_point_and_matches_probe_count(gpu_lib):
    """BASELINE config 1.  Properties: the survey's probe settled-cell count, idempotence,
    Bellman consistency of (g, pred), and predecessor chains that end at the source with cost g."""
    infl, lo = mapgen.synth_map(2048, 2048, 0.10, 5)
    H, W = infl.shape
    P = gpu_lib.Planner(H, W)
    P.set_map(infl, lo)
    src = mapgen.field_source(infl, lo)
    r = P.field(src)
    assert r["status"] == 0
    assert r["reached"] == 3767918                      # reference Djikstra settled count measured in SURVEY.md §6
    r2 = P.field(src)
    assert np.array_equal(r["g"], r2["g"]) and np.array_equal(r["pred"], r2["pred"])      # idempotent, deterministic
    g, pred = r["g"].reshape(-1), r["pred"].reshape(-1)
    free = mapgen.free_mask(infl, lo).reshape(-1)
    reach = g < 1e5
    assert not reach[~free].any() and g[src[0] * W + src[1]] == 0
    OFF = np.array([W, W + 1, 1, -W + 1, -W, -W - 1, -1, W - 1])
    k = np.nonzero(reach & (pred != 255))[0]
    par = k - OFF[pred[k]]
    # every reached cell hangs off a reached parent at exactly one fp32 step of cost 1 or sqrt(2)
    assert reach[par].all()
    is1 = g[k] == g[par] + np.float32(1.0)
    is2 = g[k] == g[par] + SQRT2F
    assert (is1 | is2).all()
    assert (reach & (pred == 255)).sum() == 1            # only the source has no predecessor
    # Bellman: on a framed map every free neighbour u of a free cell v has an edge u -> v of cost <= sqrt(2),
    # so no reached neighbour may undercut v by more than that (np.roll's wrap only touches the blocked frame)
    for d in range(8):
        nb = np.roll(g, OFF[d])                            # nb[v] = g[v - OFF[d]]
        assert not (reach & (nb < 1e5) & (nb + SQRT2F < g)).any()
    for goal in mapgen.sample_free_cells(infl, lo, 8, 99):
        fp = P.field_path(goal)
        if g[goal[0] * W + goal[1]] < 1e5:
            assert abs(fp["cost"] - float(g[goal[0] * W + goal[1]])) <= 1e-5 * fp["cost"]
            assert tuple(fp["path"][-1]) == tuple(src) and tuple(fp["path"][0]) == tuple(goal)
    P.close()


def test_c4_batch_1024_subset_against_oracle(gpu_lib, oracle_mod):
    """BASELINE config 3 (C4): batched Astar f on the 1024x1024 map; a 256-query slice of the 65,536-query
    workload runs on the GPU and 12 of them are replayed by the CPU oracle (136 ms each on one core)."""
    infl, lo = mapgen.synth_map(1024, 1024, 0.10, 7)
    P = gpu_lib.Planner(1024, 1024)
    P.set_map(infl, lo)
    qs = mapgen.sample_queries(infl, lo, 256, 7)
    b = P.plan_batch("astar", "f", qs, path_cap=0)
    assert (b["status"] == 0).all()
    b2 = P.plan_batch("astar", "f", qs[::-1].copy(), path_cap=0)                   # order of the batch is irrelevant
    assert np.array_equal(b["g_goal"], b2["g_goal"][::-1]) and np.array_equal(b["settled"], b2["settled"][::-1])
    O = oracle_mod.Oracle(infl, lo)
    for qi in range(0, 256, 22):
        o = O.plan(0, "f", qs[qi][:2], qs[qi][2:])
        assert same_cost(b["g_goal"][qi], o["g_goal"]) and b["settled"][qi] == o["settled"] and b["path_len"][qi] == len(o["path"])
    P.close()


def test_c4_every_query_of_a_768_slice_against_the_cpu_planner(gpu_lib, oracle_mod):
    """SURVEY.md 8d C4 asks for parity on a fixed subset of the 65,536 queries when the host is small: here EVERY query of a
    768-query slice is replayed by the CPU planner (the reference's own Astar when oracle/_ref is present, else the
    restatement), one planner instance per host thread, and compared: status, settled cells, path length, g(goal), and
    the full path for every 8th query."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    infl, lo = mapgen.synth_map(1024, 1024, 0.10, 7)
    nq = 768
    qs = mapgen.sample_queries(infl, lo, 4096, 7)[1000:1000 + nq].copy()          # a slice from the middle of the workload
    P = gpu_lib.Planner(1024, 1024)
    P.set_map(infl, lo)
    b = P.plan_batch("astar", "f", qs, path_cap=4096)
    P.close()
    threads = max(1, min(os.cpu_count() or 1, 32))
    if oracle_mod.have_reference():
        r = oracle_mod.Reference(infl, lo).plan_batch(0, "f", qs, nthreads=threads)      # the reference's own Astar::plan
        assert r["status"] == 0 and (b["status"] == 0).all()
        assert np.array_equal(b["settled"], r["settled"]) and np.array_equal(b["path_len"], r["path_len"])
        assert same_cost(b["g_goal"], r["g_goal"])
    O = [oracle_mod.Oracle(infl, lo) for _ in range(threads)]

    def run(t):
        return [(qi, O[t].plan(0, "f", qs[qi][:2], qs[qi][2:])) for qi in range(t, nq, threads)]

    with ThreadPoolExecutor(threads) as ex:                                           # ctypes releases the GIL inside nso_plan
        for chunk in ex.map(run, range(threads)):
            for qi, o in chunk:
                assert b["status"][qi] == o["status"] and b["settled"][qi] == o["settled"], qi
                assert b["path_len"][qi] == len(o["path"]) and same_cost(b["g_goal"][qi], o["g_goal"]), qi
                if qi % 8 == 0:
                    assert np.array_equal(b["paths"][qi][:len(o["path"])], o["path"]), qi


def test_c3_thetastar_4096_paths_are_line_of_sight_chains(gpu_lib):
    """BASELINE config 2 (C3): ThetaStar on the 4096x4096 cluttered map.  No reference exists (thetastar.cpp is
    empty), so the checks are structural: consecutive path vertices see each other, the chain is no longer than
    the goal's label, reachability agrees with Astar, and both ends are right."""
    infl, lo = mapgen.synth_map(4096, 4096, 0.20, 9)
    P = gpu_lib.Planner(4096, 4096)
    P.set_map(infl, lo)
    # starts from the C3 sampler; goals re-drawn within ~100 cells of each start to keep the test short
    starts = mapgen.sample_free_cells(infl, lo, 4, 10)
    free = mapgen.free_mask(infl, lo)
    qs = np.zeros((4, 4), np.int32)
    for t, s0 in enumerate(starts):
        i0, j0 = int(np.clip(s0[0], 100, 3995)), int(np.clip(s0[1], 100, 3995))
        near = mapgen.sample_cells(free[i0 - 100:i0 + 100, j0 - 100:j0 + 100], 1, 5 + t)[0]
        qs[t] = (s0[0], s0[1], i0 - 100 + near[0], j0 - 100 + near[1])
    t = P.plan_batch("thetastar", "f", qs, path_cap=4096)
    a = P.plan_batch("astar", "f", qs, path_cap=0)
    for qi, q in enumerate(qs):
        assert t["status"][qi] == 0
        n = t["path_len"][qi]
        path = t["paths"][qi][:n]
        if n == 1:
            assert a["path_len"][qi] == 1
            continue
        assert tuple(path[0]) == tuple(q[2:]) and tuple(path[-1]) == tuple(q[:2])
        segs = np.concatenate([path[1:], path[:-1]], 1)      # parent -> child, the direction the search tested
        assert (P.los_batch(segs) == 1).all()
        length = np.sqrt(((path[1:] - path[:-1]).astype(np.float64) ** 2).sum(1)).sum()
        # g(goal) is the label the goal held when it was popped.  Like the reference's plan() loop (astar.cpp:116-121)
        # the search re-labels closed cells without re-expanding them, so an ancestor's g may have dropped after the
        # goal's label was formed: the chain's geometric length can only be shorter than g(goal), never longer.
        assert length <= t["g_goal"][qi] * (1 + 1e-9) + 1e-9
    P.close()
