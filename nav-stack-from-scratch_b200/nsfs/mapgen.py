"""Synthetic occupancy grids and start/goal samplers, shared by tests and bench.py.

The generator is specified in terms of ``std::mt19937`` + ``std::uniform_real_distribution<double>(0,1)``
so that a C++ harness and this module produce the same map from the same seed:

* ``numpy.random.RandomState(seed)`` seeds MT19937 with the same ``init_genrand`` recurrence as
  ``std::mt19937(seed)``, and ``randint(0, 2**32, dtype=uint32)`` hands back the raw 32-bit outputs;
* libstdc++'s ``generate_canonical<double, 53>`` consumes two outputs per double:
  ``u = (x1 + x2 * 2**32) / 2**64`` evaluated in double (clamped below 1.0).

Layout contract of the produced arrays (reference ``occupancy_grid.cpp:215-218``, ``map_data.h:10-11``):
int32, row-major ``key = i * W + j``; a cell is free iff ``infl <= 0 and lo <= lo_thresh``.
"""
from __future__ import annotations

import numpy as np

LO_OCCUPIED = 20
LO_THRESH = 10
CELL_SIZE = 0.05
INFLATION_RADIUS = 4  # cells; disc a^2 + b^2 <= 16 (reference occupancy_grid.cpp:73-89)


def _mt_raw(seed: int, n: int) -> np.ndarray:
    rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)
    return rs.randint(0, 2**32, size=n, dtype=np.uint32)


def uniform01(seed: int, n: int) -> np.ndarray:
    """n draws of std::uniform_real_distribution<double>(0,1) over std::mt19937(seed)."""
    raw = _mt_raw(seed, 2 * n).astype(np.float64)
    u = (raw[0::2] + raw[1::2] * 4294967296.0) / 18446744073709551616.0
    return np.where(u >= 1.0, np.nextafter(1.0, 0.0), u)


def synth_map(H: int, W: int, p_occ: float, seed: int, inflated: bool = False, frame: bool = True):
    """Return (infl, lo) int32 [H, W].

    occupied iff u < p_occ (u drawn row-major for EVERY cell) or the cell lies on the 1-cell frame.
    occupied => lo = 20.  ``inflated``: infl counts occupied cells within the radius-4 disc (clipped at
    the border, no wrap); otherwise infl = 1 on occupied cells only.
    """
    u = uniform01(seed, H * W).reshape(H, W)
    occ = u < p_occ
    if frame:
        occ[0, :] = occ[-1, :] = True
        occ[:, 0] = occ[:, -1] = True
    lo = np.where(occ, LO_OCCUPIED, 0).astype(np.int32)
    if not inflated:
        infl = occ.astype(np.int32)
    else:
        R = INFLATION_RADIUS
        infl = np.zeros((H, W), np.int32)
        o = occ.astype(np.int32)
        for a in range(-R, R + 1):
            for b in range(-R, R + 1):
                if a * a + b * b > R * R:
                    continue
                src = o[max(0, -a): H - max(0, a), max(0, -b): W - max(0, b)]
                infl[max(0, a): H - max(0, -a), max(0, b): W - max(0, -b)] += src
    return infl, lo


def synth_map_rows(H: int, W: int, p_occ: float, seed: int, row_lo: int, row_hi: int, frame: bool = True):
    """Rows [row_lo, row_hi) of the raw (not inflated) ``synth_map(H, W, p_occ, seed)`` without building the whole map.

    MT19937 cannot jump ahead, so the stream is generated from the start in chunks and dropped until ``row_lo``; memory
    stays at a few rows' worth plus the result (a 16384 x 16384 map would otherwise need ~10 GB of temporaries per rank).
    Returns (infl, lo) int32 [row_hi - row_lo, W], bit-identical to the same rows of synth_map (tests/test_mapgen.py)."""
    rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)
    rows_per_chunk = max(1, (1 << 22) // max(W, 1))
    occ = np.zeros((row_hi - row_lo, W), bool)
    r = 0
    while r < row_hi:
        nr = min(rows_per_chunk, row_hi - r)
        raw = rs.randint(0, 2**32, size=2 * nr * W, dtype=np.uint32)
        a, b = max(r, row_lo), r + nr
        if b > a:
            blk = raw[2 * (a - r) * W:].astype(np.float64)
            u = (blk[0::2] + blk[1::2] * 4294967296.0) / 18446744073709551616.0
            occ[a - row_lo:b - row_lo] = (np.minimum(u, np.nextafter(1.0, 0.0)) < p_occ).reshape(b - a, W)
        r += nr
    if frame:
        occ[:, 0] = occ[:, -1] = True
        if row_lo == 0:
            occ[0, :] = True
        if row_hi == H:
            occ[-1, :] = True
    lo = np.where(occ, LO_OCCUPIED, 0).astype(np.int32)
    return occ.astype(np.int32), lo


def free_mask(infl: np.ndarray, lo: np.ndarray, lo_thresh: int = LO_THRESH) -> np.ndarray:
    return (infl <= 0) & (lo <= lo_thresh)


def sample_cells(mask: np.ndarray, n: int, seed: int) -> np.ndarray:
    """First n cells accepted by: mt19937(seed); i = g() % H; j = g() % W; reject where mask is False."""
    H, W = mask.shape
    out = np.zeros((0, 2), np.int32)
    draws = max(4 * n, 64)
    while True:
        raw = _mt_raw(seed, 2 * draws).astype(np.int64)
        ij = np.stack([raw[0::2] % H, raw[1::2] % W], 1)
        ok = mask[ij[:, 0], ij[:, 1]]
        out = ij[ok].astype(np.int32)
        if len(out) >= n or draws > 1 << 28:
            break
        draws *= 2
    if len(out) < n:
        raise ValueError("mask has too few accepted cells")
    return np.ascontiguousarray(out[:n])


def sample_free_cells(infl, lo, n, seed, lo_thresh=LO_THRESH):
    return sample_cells(free_mask(infl, lo, lo_thresh), n, seed)


def sample_queries(infl, lo, nq, seed, lo_thresh=LO_THRESH) -> np.ndarray:
    """int32 [nq, 4] = (si, sj, gi, gj): consecutive accepted free cells from the sampler at seed + 1."""
    cells = sample_free_cells(infl, lo, 2 * nq, seed + 1, lo_thresh)
    return np.ascontiguousarray(cells.reshape(nq, 4))


def sample_blocked_interior(infl, lo, n, seed, lo_thresh=LO_THRESH, margin=2) -> np.ndarray:
    """Non-free cells at least ``margin`` away from the border (inputs for find_better_point)."""
    m = ~free_mask(infl, lo, lo_thresh)
    if margin > 0:
        m[:margin, :] = m[-margin:, :] = False
        m[:, :margin] = m[:, -margin:] = False
    return sample_cells(m, n, seed + 1)


def field_source(infl, lo, lo_thresh=LO_THRESH):
    """(H/2, W/2) advanced in +j to the first free cell."""
    H, W = infl.shape
    fm = free_mask(infl, lo, lo_thresh)
    i, j = H // 2, W // 2
    while not fm[i, j]:
        j += 1
    return i, j
