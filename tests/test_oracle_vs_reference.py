"""The plain-C oracle against the reference's own code run live (oracle/_ref, built from /root/reference).

Skipped where the reference build is absent and cannot be rebuilt; the committed golden vectors
(test_oracle_golden.py) carry the same pinning to such machines."""
import ctypes as C

import numpy as np
import pytest

from nsfs import mapgen


@pytest.fixture(scope="module")
def ref_available(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return True


def test_mapgen_matches_std_mt19937(oracle_mod, ref_available):
    lib = C.CDLL(oracle_mod.REF_SO)
    for seed in (1, 5, 7, 1234567):
        n = 4096
        raw = np.zeros(n, np.uint32)
        lib.nsref_mt_raw(C.c_uint(seed), n, raw.ctypes.data_as(C.c_void_p))
        assert np.array_equal(raw, mapgen._mt_raw(seed, n))
        u = np.zeros(n, np.float64)
        lib.nsref_uniform01(C.c_uint(seed), n, u.ctypes.data_as(C.c_void_p))
        assert np.array_equal(u, mapgen.uniform01(seed, n))


@pytest.mark.parametrize("spec", [(200, 160, 0.10, 21, False), (150, 150, 0.006, 22, True), (90, 220, 0.30, 23, False)])
def test_plans_match_reference(oracle_mod, ref_available, spec):
    H, W, p, seed, inflated = spec
    infl, lo = mapgen.synth_map(H, W, p, seed, inflated)
    O, R = oracle_mod.Oracle(infl, lo), oracle_mod.Reference(infl, lo)
    qs = mapgen.sample_queries(infl, lo, 12, seed)
    for algo, mode in ((0, "f"), (0, "fg"), (1, "g")):
        for q in qs:
            a, b = O.generate_path(algo, mode, q[:2], q[2:]), R.generate_path(algo, mode, q[:2], q[2:])
            assert a["status"] == b["status"] == 0
            assert np.array_equal(a["path"], b["path"])
            assert a["g_goal"] == b["g_goal"] and a["settled"] == b["settled"]
            assert np.array_equal(a["final_xy"], b["final_xy"])
    src = mapgen.field_source(infl, lo)
    fa, fb = O.field(src), R.field(src)
    assert np.array_equal(fa["g"], fb["g"]) and fa["settled"] == fb["settled"]
    for bad in mapgen.sample_blocked_interior(infl, lo, 40, seed):
        a, b = O.find_better_point(bad), R.find_better_point(bad)
        assert a["status"] == b["status"]              # -1 where the search expands a cell whose testPos throws
        if a["status"] == 0:
            assert a["cell"] == b["cell"] and a["settled"] == b["settled"]


def test_survey_probe_counts(oracle_mod, ref_available):
    """SURVEY.md §8d C1 map: free fraction ~74 %, so the generator is the one the survey probed."""
    infl, lo = mapgen.synth_map(384, 384, 0.005, 1, inflated=True)
    assert abs(mapgen.free_mask(infl, lo).mean() - 0.74) < 0.01
