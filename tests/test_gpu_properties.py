"""BASELINE.json's full sizes: the reference's own planner where it finishes in seconds (C2 field, the C5 generator at
4096^2, C4 slices, the 16 C3 pairs against the restatement), size-independent properties and pinned counts / checksums where
it does not (the 16384^2 field: 153 s and ~13 GB on one host core)."""
import numpy as np
import pytest

from conftest import same_cost
from nsfs import mapgen

pytestmark = pytest.mark.gpu

SQRT2F = np.float32(1.41421356237309504880)


def _cpu_field(oracle_mod, infl, lo, src):
    """The reference's own Djikstra::plan run to exhaustion (oracle/_ref) when it was built, else the C restatement."""
    impl = oracle_mod.Reference(infl, lo) if oracle_mod.have_reference() else oracle_mod.Oracle(infl, lo)
    return impl.field(src)


def _assert_field_matches_cpu(r, o):
    """reachability and settled count exact, cost within 1e-5 relative of the fp64 reference (north_star's rule)."""
    reach = o["g"] < 1e5
    assert r["status"] == 0 and o["status"] == 0
    assert np.array_equal(r["g"] < 1e5, reach)
    assert r["reached"] == int(reach.sum()) == o["settled"]
    rel = np.abs(r["g"].astype(np.float64) - o["g"])[reach] / np.maximum(o["g"][reach], 1.0)
    assert rel.max() <= 1e-5, rel.max()


def test_c2_field_2048_is_a_fixed_point_and_matches_probe_count(gpu_lib, oracle_mod):
    """BASELINE config 1.  g and reachability against the reference's own Djikstra::plan on the same map
    (djikstra.cpp:27-145, ~1 s), the survey's probe settled-cell count, idempotence, Bellman consistency of (g, pred) with
    the per-edge quirk weights, and predecessor chains that end at the source with cost g."""
    infl, lo = mapgen.synth_map(2048, 2048, 0.10, 5)
    H, W = infl.shape
    P = gpu_lib.Planner(H, W)
    P.set_map(infl, lo)
    src = mapgen.field_source(infl, lo)
    r = P.field(src)
    assert r["status"] == 0
    assert r["reached"] == 3767918                      # reference Djikstra settled count measured in SURVEY.md §6
    _assert_field_matches_cpu(r, _cpu_field(oracle_mod, infl, lo, src))
    r2 = P.field(src)
    assert np.array_equal(r["g"], r2["g"]) and np.array_equal(r["pred"], r2["pred"])      # idempotent, deterministic
    g, pred = r["g"].reshape(-1), r["pred"].reshape(-1)
    free = mapgen.free_mask(infl, lo).reshape(-1)
    reach = g < 1e5
    assert not reach[~free].any() and g[src[0] * W + src[1]] == 0
    OFF = np.array([W, W + 1, 1, -W + 1, -W, -W - 1, -1, W - 1])
    k = np.nonzero(reach & (pred != 255))[0]
    par = k - OFF[pred[k]]
    # every reached cell hangs off a reached parent at exactly one step of cost 1 or sqrt(2): the solver works in Q17.15 integers
    # (1 -> 32768, sqrt(2) -> 46341) and hands out fp32(q / 32768), so q is recovered to within the fp32 rounding of g (< 2^-23 relative)
    assert reach[par].all()
    q = np.rint(g.astype(np.float64) * 32768.0)
    tol = 2.0 + q[k] * 2.0 ** -22
    step = q[k] - q[par]
    assert ((np.abs(step - 32768.0) <= tol) | (np.abs(step - 46341.0) <= tol)).all()
    assert (reach & (pred == 255)).sum() == 1            # only the source has no predecessor
    # Bellman with the per-edge quirk weights (astar.cpp:88-123): the edge u -> u + OFF[d] exists iff the target is free and
    # costs 1 if an even number of PASSABLE directions precede d at u, else sqrt(2); no in-edge may undercut g[v] (in the solver's
    # own Q17.15 units, up to the fp32 rounding of the values handed out).  On a framed map np.roll's wrap only touches the blocked frame.
    npass = np.zeros(g.shape, np.int32)
    for d in range(8):
        tgt_free = np.roll(free, -OFF[d])                  # tgt_free[u] = free[u + OFF[d]]
        w = np.where(npass % 2 == 0, 32768.0, 46341.0)
        cand = np.roll(np.where(free & tgt_free & reach, q + w, 1e18), OFF[d])     # cand[v] from u = v - OFF[d]
        assert not (reach & (cand < q - (2.0 + q * 2.0 ** -21))).any(), d
        assert not ((~reach) & free & (cand < 1e17)).any(), d         # an unreached free cell has no reached in-neighbour
        npass += tgt_free & free
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


def test_c3_thetastar_4096_the_16_sampler_pairs_against_the_restatement(gpu_lib, oracle_mod):
    """BASELINE config 2 (C3): ThetaStar on the 4096x4096 cluttered map, the 16 un-shortened (start, goal) pairs of the sampler.
    No reference implementation exists (thetastar.cpp is empty), so the checker is the restatement nso_plan(NSO_THETASTAR)
    (parity unpinned by design); what IS pinned underneath, isLos / bresenham_los, is re-checked on the returned chains."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    infl, lo = mapgen.synth_map(4096, 4096, 0.20, 9)
    qs = mapgen.sample_queries(infl, lo, 16, 9)
    P = gpu_lib.Planner(4096, 4096)
    P.set_map(infl, lo)
    t = P.plan_batch("thetastar", "f", qs, path_cap=8192)
    a = P.plan_batch("astar", "f", qs, path_cap=0)
    threads = max(1, min(os.cpu_count() or 1, 16))
    O = [oracle_mod.Oracle(infl, lo) for _ in range(threads)]

    def run(k):
        return [(qi, O[k].plan(oracle_mod.THETASTAR, "f", qs[qi][:2], qs[qi][2:])) for qi in range(k, 16, threads)]

    with ThreadPoolExecutor(threads) as ex:                      # ctypes releases the GIL inside nso_plan
        for chunk in ex.map(run, range(threads)):
            for qi, o in chunk:
                n = len(o["path"])
                assert t["status"][qi] == o["status"] == 0, qi
                assert t["settled"][qi] == o["settled"] and t["path_len"][qi] == n, qi
                assert t["g_goal"][qi] == o["g_goal"], qi                       # the wide kernel's fp64 sums are the restatement's own
                assert np.array_equal(t["paths"][qi][:n], o["path"]), qi
    for qi, q in enumerate(qs):
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


def test_c5_generator_4096_field_against_the_reference(gpu_lib, oracle_mod):
    """The C5 map generator (p_occ 0.10, seed 11) at 4096^2, where the reference's Djikstra::plan (djikstra.cpp:27-145) still
    finishes in seconds: reachability and settled count exact, cost within 1e-5 relative."""
    infl, lo = mapgen.synth_map(4096, 4096, 0.10, 11)
    P = gpu_lib.Planner(4096, 4096)
    P.set_map(infl, lo)
    src = mapgen.field_source(infl, lo)
    r = P.field(src)
    P.close()
    _assert_field_matches_cpu(r, _cpu_field(oracle_mod, infl, lo, src))


def test_c5_field_16384_count_checksum_and_fallback_calls(gpu_lib, oracle_mod):
    """BASELINE config 4 (C5) at full size.  (i) The field: the reference settles 241 537 616 cells on this map (SURVEY.md §6
    probe, 153 s on one core), and the position-weighted checksum of g / pred equals the committed single-GPU one
    (tests/golden/c5_field_checksum.json: what bench.py's multi-GPU run is compared with).  (ii) 128 find_better_point calls
    from random non-free interior cells against the CPU planner (djikstra.cpp:165-273) run on a window of 129 full-width rows
    around each cell: the search touches a few dozen cells, the flat keys inside the window differ from the full map's by a
    constant, so the popped order and the returned cell are the same as long as the search stays off the window's edge rows
    (asserted)."""
    import json
    import os

    import torch
    from nsfs import sharded_field as sf

    H = W = 16384
    infl, lo = mapgen.synth_map_rows(H, W, 0.10, 11, 0, H)
    P = gpu_lib.Planner(H, W)
    P.set_map(infl, lo)
    fm = mapgen.free_mask(infl[H // 2:H // 2 + 1], lo[H // 2:H // 2 + 1])[0]
    sj = W // 2
    while not fm[sj]:
        sj += 1
    r = P.field_device((H // 2, sj))
    assert r["status"] == 0 and r["reached"] == 241537616
    dev = torch.device("cuda", 0)
    g = sf._tensor_from_ptr(torch, r["d_g"], (H, W), "<f4", dev)
    pred = sf._tensor_from_ptr(torch, r["d_pred"], (H, W), "|u1", dev)
    sg, sp = sf.field_checksum(torch, g, pred, 0)
    # a row-block split of the same field sums to the same values (what the multi-GPU bench relies on)
    parts = [sf.field_checksum(torch, g[a:b], pred[a:b], a * W) for a, b in ((0, 5000), (5000, 16384))]
    assert (parts[0][0] + parts[1][0]) % (1 << 64) == sg and (parts[0][1] + parts[1][1]) % (1 << 64) == sp
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c5_field_checksum.json")))
    assert f"{sg:016x}" == want["checksum_g"] and f"{sp:016x}" == want["checksum_pred"]
    # the predecessor of every reached cell, sampled: one exact fp32 step below it
    gs, ps = g[8000:8064].cpu().numpy().reshape(-1), pred[8000:8064].cpu().numpy().reshape(-1)
    del g, pred
    bad = mapgen.sample_blocked_interior(infl, lo, 128, 11, margin=80)
    fb = P.find_better_point_batch(bad)
    P.close()
    assert (fb["status"] == 0).all()
    for t, (bi, bj) in enumerate(bad):
        r0 = int(bi) - 64
        O = oracle_mod.Oracle(infl[r0:r0 + 129], lo[r0:r0 + 129])
        o = O.find_better_point((64, int(bj)))
        assert o["status"] == 0 and 2 <= o["cell"][0] <= 126 and o["settled"] < 2000, (t, o)
        assert tuple(fb["cells"][t]) == (o["cell"][0] + r0, o["cell"][1]), (t, tuple(fb["cells"][t]), o["cell"])
    assert ((gs < 1e5) == (ps != 255)).all()


