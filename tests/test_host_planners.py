"""The C++17 GridPlannerCore children (host/gpu_planners.cpp): interface, runtime selection, and - on the GPU -
generatePath() / find_better_point() against the golden vectors produced by the reference's own classes.

generatePath = plan + post_process_path (grid_planner_core.cpp:603-615); the golden ``finals`` are the reference's
world-coordinate outputs, so equality of doubles (``==``) is the bar: the index<->world conversions and the
collinearity filter are the same fp64 expressions, and the search underneath is order-exact."""
import os

import numpy as np
import pytest

from conftest import COMBOS, GOLDEN_MAPS, PKG, ROOT, golden_map, split_by


# our children compiled against the reference's own headers + base-class sources (oracle/Makefile, reference present only)
DROPIN_SO = os.path.join(ROOT, "oracle", "_ref", "libnsfs_dropin.so")


@pytest.fixture(scope="module")
def host_mod():
    from nsfs import host
    host.lib()
    return host


def test_host_library_exports_the_shim(host_mod):
    L = host_mod.lib()
    for sym in host_mod.SYMBOLS:
        assert hasattr(L, sym), sym


def test_shipped_yaml_selects_like_the_reference(host_mod, tmp_path):
    cfg = host_mod.load_config(os.path.join(PKG, "global_planner.yaml"))
    assert cfg["planner_name"] == "astar" and cfg["cost_mode"] == "f" and cfg["gp_rate"] == 25.0 and cfg["verbose_planner"]
    # defaults of loadParams (global_planner.cpp:482-490) when the keys are absent: astar / fg
    p = tmp_path / "empty.yaml"
    p.write_text("com_rate: 25 #rate of commander\n")
    cfg = host_mod.load_config(p)
    assert cfg["planner_name"] == "astar" and cfg["cost_mode"] == "fg"
    # the upstream file's own formatting: quoted values with trailing comments (core_turtle_config.yaml:33-36)
    p = tmp_path / "core.yaml"
    p.write_text('gp_rate: 12.5 #rate\nverbose_planner: false #verbosity of logs\nplanner_name: "djikstra" #name of planner --> Options:\n'
                 'cost_mode: "g" #cost function to use\ndjikstra_backend: field\ngpu_device: 3\n')
    cfg = host_mod.load_config(p)
    assert cfg == dict(planner_name="djikstra", cost_mode="g", djikstra_backend="field", gp_rate=12.5, verbose_planner=False, device=3)


def test_sparsify_matches_the_oracle_restatement(host_mod, oracle_mod, golden):
    """sparsifyPath is host arithmetic (no device): run it on every golden raw path."""
    for mname in GOLDEN_MAPS:
        O = oracle_mod.Oracle(*golden_map(golden, mname))
        key = f"{mname}/astar_f"
        for path in split_by(golden[key + "/paths"], golden[key + "/path_len"]):
            xy = np.array([O.idx2pos(c) for c in path], np.float64)
            assert np.array_equal(host_mod.sparsify(xy), O.sparsify(xy))


def test_reference_header_build_exists_where_the_reference_is(host_mod):
    if not os.path.isdir("/root/reference/src/global_planner"):
        pytest.skip("no /root/reference on this machine")
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    L = host_mod.lib(DROPIN_SO)          # children compiled against the reference's own grid_planner_core.h
    for sym in host_mod.SYMBOLS:
        assert hasattr(L, sym), sym


# ---- GPU ------------------------------------------------------------------------------------------------------
def _variants(host_mod):
    v = [None]
    if os.path.exists(DROPIN_SO):
        v.append(DROPIN_SO)
    return v


@pytest.mark.gpu
@pytest.mark.parametrize("mname", GOLDEN_MAPS)
@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: f"{c[0]}_{c[2]}")
def test_generate_path_matches_reference_finals(host_mod, golden, mname, combo):
    aname, _, mode = combo
    infl, lo = golden_map(golden, mname)
    key = f"{mname}/{aname}_{mode}"
    qs = golden[f"{mname}/queries"]
    finals = split_by(golden[key + "/finals"], golden[key + "/final_len"])
    for so in _variants(host_mod):
        S = host_mod.Session(aname, mode, infl, lo, so_path=so)
        assert S.kind() == aname
        for qi, q in enumerate(qs):
            r = S.generate_path_idx(q[:2], q[2:])
            if golden[key + "/status"][qi] != 0:
                assert r["status"] == "out_of_range", (mname, combo, qi)       # std::out_of_range, like vector::at
                continue
            assert r["status"] == "ok"
            assert np.array_equal(r["path"], finals[qi]), (mname, combo, qi, so)
            # the Pos2D overload converts with pos2idx first (grid_planner_core.cpp:610-615)
            r2 = S.generate_path_pos(np.array(q[:2]) * 0.05 + 0.01, np.array(q[2:]) * 0.05 - 0.01)
            assert r2["status"] == "ok" and np.array_equal(r2["path"], finals[qi])
        S.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_fallback_planner_matches_reference(host_mod, golden, mname):
    infl, lo = golden_map(golden, mname)
    for so in _variants(host_mod):
        S = host_mod.Session("astar", "f", infl, lo, so_path=so)
        for b, st, out in zip(golden[f"{mname}/fb_in"], golden[f"{mname}/fb_status"], golden[f"{mname}/fb_out"]):
            r = S.find_better_point(b)
            if st != 0:
                assert r["status"] == "out_of_range"
                continue
            assert r["pos"] == (out[0] * 0.05 + 0.0, out[1] * 0.05 + 0.0)     # idx2pos, grid_planner_core.cpp:416-421
            assert S.find_better_point((b[0] * 0.05, b[1] * 0.05), as_pos=True)["pos"] == r["pos"]
        S.close()


@pytest.mark.gpu
def test_selection_rule_and_field_backend(host_mod, golden, oracle_mod):
    infl, lo = golden_map(golden, "m0")
    S = host_mod.Session("no-such-planner", "zz", infl, lo)          # unknown name => astar, unknown mode => fg
    assert S.kind() == "astar"
    O = oracle_mod.Oracle(infl, lo)
    q = golden["m0/queries"][0]
    assert np.array_equal(S.generate_path_idx(q[:2], q[2:])["path"], O.generate_path(0, "fg", q[:2], q[2:])["final_xy"])
    S.close()
    assert host_mod.Session("thetastar", "f", infl, lo).kind() == "thetastar"
    # djikstra with the bulk field backend: same reachability and cost as the exact-order search
    F = host_mod.Session("djikstra", "g", infl, lo, backend="field")
    E = host_mod.Session("djikstra", "g", infl, lo, backend="exact")
    for q in golden["m0/queries"][:8]:
        pf, pe = F.generate_path_idx(q[:2], q[2:])["path"], E.generate_path_idx(q[:2], q[2:])["path"]
        assert (len(pf) == 1) == (len(pe) == 1)
        assert np.array_equal(pf[0], pe[0]) and np.array_equal(pf[-1], pe[-1])
    F.close(); E.close()


@pytest.mark.gpu
def test_map_replaced_between_calls_like_the_ros_callbacks(host_mod, golden, oracle_mod):
    """global_planner.cpp:65-73 swaps the grid vectors whenever a message arrives; the next plan must see them."""
    infl, lo = golden_map(golden, "m0")
    infl2, lo2 = golden_map(golden, "m2")
    S = host_mod.Session("astar", "f", infl, lo)
    q = golden["m0/queries"][1]
    a = S.generate_path_idx(q[:2], q[2:])["path"]
    assert np.array_equal(a, oracle_mod.Oracle(infl, lo).generate_path(0, "f", q[:2], q[2:])["final_xy"])
    lo_b = lo.copy()
    lo_b[q[2], q[3]] = 20                                             # the goal cell becomes occupied
    S.set_map(infl, lo_b)
    b = S.generate_path_idx(q[:2], q[2:])["path"]
    assert np.array_equal(b, oracle_mod.Oracle(infl, lo_b).generate_path(0, "f", q[:2], q[2:])["final_xy"])
    S.close()


@pytest.mark.gpu
def test_plan_request_four_cases_like_global_planner_run(host_mod, oracle_mod):
    """global_planner.cpp:164-273: both ok -> plan; bad goal / bad robot / both -> fallback repair(s), then plan.  The pieces
    are checked against the CPU planner: find_better_point's cell, then generatePath from the repaired ends."""
    from nsfs import mapgen
    infl, lo = mapgen.synth_map(96, 120, 0.004, 81, inflated=True)
    O = oracle_mod.Oracle(infl, lo)
    free = mapgen.sample_free_cells(infl, lo, 4, 81)
    bad = mapgen.sample_blocked_interior(infl, lo, 4, 81, margin=6)
    S = host_mod.Session("astar", "f", infl, lo)
    pos = lambda c: (c[0] * 0.05, c[1] * 0.05)
    for case, robot, goal in ((1, free[0], free[1]), (2, free[2], bad[0]), (3, bad[1], free[3]), (4, bad[2], bad[3])):
        r = S.plan_request(pos(robot), pos(goal))
        assert r["status"] == "ok" and r["case"] == case
        assert r["robot_ok"] == (case in (1, 2)) and r["goal_ok"] == (case in (1, 3))
        start = robot if case in (1, 2) else O.find_better_point(robot)["cell"]
        end = goal if case in (1, 3) else O.find_better_point(goal)["cell"]
        assert r["goal_updated"] == (case in (2, 4)) and r["backup_mode"] == (case == 3)
        assert r["goal"] == pos(end) and (case in (1, 2) or r["backup_robot"] == pos(start))
        assert np.array_equal(r["path"], O.generate_path(0, "f", start, end)["final_xy"])
    S.close()


@pytest.mark.gpu
def test_dirty_rows_update_equals_a_full_upload(gpu_lib):
    """nsfs_update_map_rows: only the changed rows travel and only the masks that can see them are rebuilt; every
    downstream result must equal a planner that received the whole updated map."""
    from nsfs import mapgen
    infl, lo = mapgen.synth_map(300, 260, 0.10, 82)
    infl2, lo2 = infl.copy(), lo.copy()
    rng = np.random.RandomState(82)
    for (a, b) in ((0, 3), (130, 141), (296, 300)):                       # top edge, middle, bottom edge
        blk = rng.rand(b - a, 260) < 0.2
        infl2[a:b] = np.where(blk, 1, 0); lo2[a:b] = np.where(blk, 20, 0)
        infl2[a:b, 0] = infl2[a:b, -1] = 1; lo2[a:b, 0] = lo2[a:b, -1] = 20
    infl2[0] = infl2[-1] = 1; lo2[0] = lo2[-1] = 20
    A = gpu_lib.Planner(300, 260); A.set_map(infl, lo)
    for (a, b) in ((0, 3), (130, 141), (296, 300)):
        A.update_map_rows(a, infl2[a:b], lo2[a:b])
    B = gpu_lib.Planner(300, 260); B.set_map(infl2, lo2)
    every = np.stack(np.meshgrid(np.arange(300), np.arange(260), indexing="ij"), -1).reshape(-1, 2)
    assert np.array_equal(A.test_pos(every), B.test_pos(every))
    src = mapgen.sample_free_cells(infl2, lo2, 1, 82)[0]
    fa, fb = A.field(src), B.field(src)
    assert np.array_equal(fa["g"], fb["g"]) and np.array_equal(fa["pred"], fb["pred"])
    qs = mapgen.sample_queries(infl2, lo2, 24, 82)
    pa, pb = A.plan_batch("astar", "f", qs, path_cap=1024), B.plan_batch("astar", "f", qs, path_cap=1024)
    assert np.array_equal(pa["paths"], pb["paths"]) and np.array_equal(pa["settled"], pb["settled"])
    A.close(); B.close()
