import numpy as np

from nsfs import mapgen


def test_deterministic_and_framed():
    a = mapgen.synth_map(40, 56, 0.2, 9)
    b = mapgen.synth_map(40, 56, 0.2, 9)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    free = mapgen.free_mask(*a)
    assert not free[0].any() and not free[-1].any() and not free[:, 0].any() and not free[:, -1].any()
    assert 0.7 < free[1:-1, 1:-1].mean() < 0.9


def test_inflation_counts_disc():
    H, W = 30, 34
    infl, lo = mapgen.synth_map(H, W, 0.03, 4, inflated=True)
    occ = lo > mapgen.LO_THRESH
    brute = np.zeros((H, W), np.int32)
    for i in range(H):
        for j in range(W):
            for a in range(-4, 5):
                for b in range(-4, 5):
                    if a * a + b * b <= 16 and 0 <= i + a < H and 0 <= j + b < W and occ[i + a, j + b]:
                        brute[i, j] += 1
    assert np.array_equal(infl, brute)


def test_samplers():
    infl, lo = mapgen.synth_map(64, 64, 0.1, 3)
    free = mapgen.free_mask(infl, lo)
    qs = mapgen.sample_queries(infl, lo, 50, 3)
    assert qs.shape == (50, 4) and free[qs[:, 0], qs[:, 1]].all() and free[qs[:, 2], qs[:, 3]].all()
    assert np.array_equal(qs, mapgen.sample_queries(infl, lo, 50, 3))
    assert np.array_equal(qs[:10], mapgen.sample_queries(infl, lo, 10, 3))      # prefix-stable
    bad = mapgen.sample_blocked_interior(infl, lo, 20, 3)
    assert not free[bad[:, 0], bad[:, 1]].any() and bad.min() >= 2
    s = mapgen.field_source(infl, lo)
    assert free[s]


def test_row_range_generator_matches_the_whole_map():
    infl, lo = mapgen.synth_map(301, 257, 0.1, 77)
    for a, b in ((0, 301), (0, 17), (120, 200), (290, 301)):
        i2, l2 = mapgen.synth_map_rows(301, 257, 0.1, 77, a, b)
        assert np.array_equal(i2, infl[a:b]) and np.array_equal(l2, lo[a:b])
