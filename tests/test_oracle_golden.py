"""The plain-C oracle (oracle/nsfs_oracle.c) against the golden vectors produced by the reference itself.

Everything here is integer/byte/index work or the reference's own fp64 sums in the reference's own order, so
the bar is bit-exact equality (``==`` on doubles included)."""
import numpy as np
import pytest

from conftest import COMBOS, GOLDEN_MAPS, golden_map, split_by


@pytest.fixture(scope="module")
def oracles(golden, oracle_mod):
    return {m: oracle_mod.Oracle(*golden_map(golden, m)) for m in GOLDEN_MAPS}


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: f"{c[0]}_{c[2]}")
def test_plan_and_generate_path(golden, oracles, mname, combo):
    aname, algo, mode = combo
    O = oracles[mname]
    key = f"{mname}/{aname}_{mode}"
    qs = golden[f"{mname}/queries"]
    paths = split_by(golden[key + "/paths"], golden[key + "/path_len"])
    finals = split_by(golden[key + "/finals"], golden[key + "/final_len"])
    for qi, q in enumerate(qs):
        r = O.generate_path(algo, mode, q[:2], q[2:])
        assert r["status"] == golden[key + "/status"][qi], (mname, combo, qi)
        if r["status"] != 0:
            continue
        assert np.array_equal(r["path"], paths[qi])
        assert r["g_goal"] == golden[key + "/g_goal"][qi]
        assert r["settled"] == golden[key + "/settled"][qi]
        assert np.array_equal(r["final_xy"], finals[qi])     # world coordinates, bit-exact doubles


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_field(golden, oracles, mname):
    O = oracles[mname]
    for t, s in enumerate(golden[f"{mname}/field_src"]):
        r = O.field(s)
        assert r["status"] == golden[f"{mname}/field{t}/status"][0]
        if r["status"] == 0:
            assert np.array_equal(r["g"], golden[f"{mname}/field{t}/g"])
            assert np.array_equal(r["visited"], golden[f"{mname}/field{t}/visited"])
            # order independence: a FIFO label-correcting fixed point lands on the same field
            rc, gfp = O.field_fixed_point(s)
            assert rc == 0
            assert np.array_equal(gfp < 1e5, r["g"] < 1e5)
            assert np.max(np.abs(gfp - r["g"])) < 1e-9


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_find_better_point(golden, oracles, mname):
    O = oracles[mname]
    for b, st, out in zip(golden[f"{mname}/fb_in"], golden[f"{mname}/fb_status"], golden[f"{mname}/fb_out"]):
        r = O.find_better_point(b)
        assert r["status"] == st
        if st == 0:
            assert r["cell"] == tuple(out)


@pytest.mark.parametrize("mname", GOLDEN_MAPS)
def test_rays_and_los(golden, oracles, mname):
    O = oracles[mname]
    segs = golden[f"{mname}/segs"]
    rays = split_by(golden[f"{mname}/rays"], golden[f"{mname}/ray_len"])
    for s, ray, los, do, de in zip(segs, rays, golden[f"{mname}/los"], golden[f"{mname}/dist_oct"], golden[f"{mname}/dist_euc"]):
        assert np.array_equal(O.bresenham(s[:2], s[2:]), ray)
        assert np.array_equal(O.bresenham(s[:2], s[2:], closed=True), ray)     # closed form == recurrence
        assert O.is_los(s[:2], s[2:]) == los
        assert O.dist_oct(s[:2], s[2:]) == do
        assert O.dist_euc(s[:2], s[2:]) == de


@pytest.mark.parametrize("mode", ["f", "g", "fg"])
def test_open_list_trace(golden, oracles, mode):
    got = oracles["m0"].heap_trace(mode, golden[f"heap/{mode}/ops"])
    assert np.array_equal(got, golden[f"heap/{mode}/pops"])


def test_closed_form_bresenham_exhaustive(oracles):
    O = oracles["m0"]
    for di in range(-9, 10):
        for dj in range(-9, 10):
            a = O.bresenham((20, 20), (20 + di, 20 + dj))
            b = O.bresenham((20, 20), (20 + di, 20 + dj), closed=True)
            assert np.array_equal(a, b), (di, dj)
