"""The opt-in textbook graph (``graph: "textbook"`` / nsfs_set_option(h, "graph", 1)): the same kernels on
edge masks of the plain 8-connected grid (rows AND columns bounds-checked, diagonal steps sqrt 2) instead of the reference's
quirk graph.  There is no reference behaviour to match (parity unpinned by design, like ThetaStar): the checker is a
textbook fp64 Dijkstra (oracle/nsfs_oracle.c nso_textbook_field), itself checked against scipy's.  Bars: reachability exact,
cost within 1e-5 relative (fp32 field) / 1e-9 (fp64 exact-order kernel), every path step a true 8-neighbour over free cells.
The default graph stays the reference's: switching the option off restores the golden results bit for bit."""
import numpy as np
import pytest

from nsfs import mapgen


def _scipy_field(infl, lo, src):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import dijkstra
    H, W = infl.shape
    free = mapgen.free_mask(infl, lo)
    rows, cols, w = [], [], []
    DI, DJ = (1, 1, 0, -1, -1, -1, 0, 1), (0, 1, 1, 1, 0, -1, -1, -1)
    ii, jj = np.mgrid[0:H, 0:W]
    for d in range(8):
        ni, nj = ii + DI[d], jj + DJ[d]
        ok = (ni >= 0) & (ni < H) & (nj >= 0) & (nj < W)
        ok &= free[np.clip(ni, 0, H - 1), np.clip(nj, 0, W - 1)]
        ok &= free | ((ii == src[0]) & (jj == src[1]))          # only free cells (and the source) are expanded
        rows.append((ii * W + jj)[ok]); cols.append((ni * W + nj)[ok]); w.append(np.full(ok.sum(), np.sqrt(2.0) if d & 1 else 1.0))
    g = coo_matrix((np.concatenate(w), (np.concatenate(rows), np.concatenate(cols))), shape=(H * W, H * W)).tocsr()
    return dijkstra(g, indices=src[0] * W + src[1]).reshape(H, W)


def test_textbook_checker_against_scipy(oracle_mod):
    infl, lo = mapgen.synth_map(90, 70, 0.15, 61, False, False)
    src = tuple(mapgen.sample_free_cells(infl, lo, 1, 61)[0])
    got = oracle_mod.Oracle(infl, lo).textbook_field(src)
    want = _scipy_field(infl, lo, src)
    reach = np.isfinite(want)
    assert np.array_equal(got < 1e5, reach)
    assert np.allclose(got[reach], want[reach], rtol=1e-12, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("spec", [(300, 260, 0.10, 71, True), (200, 333, 0.20, 72, False), (65, 64, 0.05, 73, False)],
                         ids=["framed", "unframed-no-wrap", "tile-edge"])
def test_textbook_field_and_paths(gpu_lib, oracle_mod, spec):
    H, W, p, seed, frame = spec
    infl, lo = mapgen.synth_map(H, W, p, seed, False, frame)
    free = mapgen.free_mask(infl, lo)
    src = tuple(mapgen.sample_free_cells(infl, lo, 1, seed)[0])
    O = oracle_mod.Oracle(infl, lo)
    want = O.textbook_field(src)
    P = gpu_lib.Planner(H, W)
    P.set_map(infl, lo)
    quirk = P.field(src)
    P.set_option("graph", 1)
    got = P.field(src)
    assert got["status"] == 0                                    # nothing throws on this graph
    reach = want < 1e5
    assert np.array_equal(got["g"] < 1e5, reach) and got["reached"] == int(reach.sum())
    assert np.all(np.abs(got["g"][reach] - want[reach]) <= 1e-5 * np.maximum(want[reach], 1.0))
    goals = mapgen.sample_free_cells(infl, lo, 24, seed + 1)
    for gl in goals:
        gl = tuple(gl)
        for algo, tol in (("djikstra_field", 1e-5), ("djikstra", 1e-9)):
            r = P.plan(algo, "g", src, gl)
            assert r["status"] == 0
            if not reach[gl]:
                assert r["n_path"] == 1 and tuple(r["path"][0]) == src
                continue
            path = r["path"]
            assert tuple(path[0]) == gl and tuple(path[-1]) == src               # goal -> start, like the reference
            step = np.abs(np.diff(path.astype(np.int64), axis=0))
            assert step.max() <= 1 and (step.sum(1) >= 1).all()                 # true 8-neighbours: no column wrap
            assert free[path[:, 0], path[:, 1]].all()
            cost = float(np.where(step.sum(1) == 2, np.sqrt(2.0), 1.0).sum())
            assert abs(cost - want[gl]) <= tol * max(want[gl], 1.0), (algo, gl, cost, want[gl])
            assert abs(r["g_goal"] - want[gl]) <= tol * max(want[gl], 1.0)
    P.set_option("graph", 0)                                      # back to the reference's graph: the very same bits as before
    again = P.field(src)
    assert np.array_equal(again["g"], quirk["g"]) and np.array_equal(again["pred"], quirk["pred"])
    P.close()


@pytest.mark.gpu
def test_host_planner_with_the_textbook_graph(oracle_mod):
    from nsfs import host
    infl, lo = mapgen.synth_map(160, 160, 0.12, 81)
    O = oracle_mod.Oracle(infl, lo)
    pts = mapgen.sample_free_cells(infl, lo, 8, 82)
    S = host.Session("djikstra", "g", infl, lo, backend="field")
    S.set_graph(True)
    for a, b in zip(pts[:4], pts[4:]):
        want = O.textbook_field(tuple(a))[tuple(b)]
        r = S.generate_path_idx(tuple(a), tuple(b))
        assert r["status"] == "ok"
        xy = r["path"]
        if want >= 1e5:
            assert len(xy) == 1
            continue
        ij = np.rint(xy / 0.05).astype(int)
        assert tuple(ij[0]) == tuple(b) and tuple(ij[-1]) == tuple(a)
        # the any-angle post-processing only shortens: the polyline is no longer than the grid optimum
        length = float(np.hypot(*np.diff(ij, axis=0).T).sum())
        assert length <= want * (1 + 1e-9)
    S.close()
