"""Multi-rank host logic on CPU: row-partitioned field (C5) and query sharding (C4), in process and under
torch.distributed gloo with world_size 2.  The local solver is the oracle (tests/shard_helpers.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from nsfs import mapgen, sharded_field as sf, sharded_queries as sq
from shard_helpers import INF, make_oracle_shards, monolithic


def test_row_partition_bookkeeping():
    p = sf.RowPartition(103, 50, 4)
    assert p.bounds == [0, 26, 52, 78, 103]
    assert p.owned(1) == (26, 52) and p.held(1) == (22, 56) and p.active(1) == (24, 54)
    assert p.held(0) == (0, 30) and p.held(3) == (74, 103) and p.active(3) == (76, 103)
    assert [p.owner_of_row(i) for i in (0, 25, 26, 102)] == [0, 0, 1, 3]
    with pytest.raises(ValueError):
        sf.RowPartition(5, 50, 4)


def test_query_slices_cover_the_batch_once():
    for nq, world in ((65536, 8), (10, 4), (3, 8), (0, 2)):
        got = [sq.query_slice(nq, world, r) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == nq
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("spec", [(96, 64, 0.12, 31, True), (61, 47, 0.25, 32, True), (40, 23, 0.10, 33, False)],
                         ids=["framed", "dense", "unframed-wrap"])
@pytest.mark.parametrize("world,band", [(1, None), (2, None), (3, None), (5, None), (2, 12.5), (3, 40.0)],
                         ids=["1", "2", "3", "5", "2-band12.5", "3-band40"])
def test_sharded_field_equals_monolithic(oracle_mod, spec, world, band):
    H, W, p, seed, frame = spec
    if band is not None:                                 # every banded round re-queues the whole shard in the CPU checker
        H, W = max(H // 2, 6 * world), max(W // 2, 12)
    infl, lo = mapgen.synth_map(H, W, p, seed, False, frame)
    if not frame:
        # no frame => the flat-key stencil wraps around the column edges.  Block only the four cells whose
        # expansion throws (R6) so the solve runs to completion THROUGH the wrap instead of stopping at the exception.
        for c in ((0, 0), (1, 0), (H - 1, W - 1), (H - 2, W - 1)):
            infl[c] = 1
            lo[c] = 20
    src = mapgen.sample_free_cells(infl, lo, 1, seed)[0]
    rc, want = monolithic(oracle_mod, infl, lo, src)
    part, shards = make_oracle_shards(oracle_mod, infl, lo, world)
    out = sf.solve_in_process(shards, src, band=band)
    statuses = [r["status"] for r in out["results"]]
    assert rc >= 0 and all(s == 0 for s in statuses)
    got = np.concatenate([s.owned_rows() for s in shards])
    assert np.array_equal(got, want)                     # slack 0 => unique fixed point => bit-identical
    assert sum(r["reached"] for r in out["results"]) == int((want < INF).sum())
    if not frame:                                        # the wrap really was used: some cell is closer than any in-row walk allows
        assert (want < INF).sum() > 0.5 * H * W
    # ghost rows agree with their owners after convergence
    for s in shards[1:]:
        assert np.array_equal(s.solver.rows(s.loc(s.r0 - sf.GHOST), sf.GHOST), want[s.r0 - sf.GHOST:s.r0])


def test_sharded_field_reports_the_reference_exception(oracle_mod):
    """Unframed map, corner cells reachable: expanding (0,0) indexes key -1 and the reference throws."""
    infl, lo = mapgen.synth_map(40, 23, 0.10, 33, False, False)
    src = (20, 11)
    assert monolithic(oracle_mod, infl, lo, src)[0] < 0
    part, shards = make_oracle_shards(oracle_mod, infl, lo, 3)
    out = sf.solve_in_process(shards, src)
    assert any(r["status"] < 0 for r in out["results"])


def _run_workers(tmp_path, mode, world=2):
    port = 29500 + (os.getpid() % 2000)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world))
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), mode, str(tmp_path)],
                                      env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    return outs


def test_gloo_world2_sharded_field(tmp_path, oracle_mod):
    outs = _run_workers(tmp_path, "field")
    assert "FIELD_OK" in outs[0]


def test_gloo_world2_sharded_queries(tmp_path, oracle_mod):
    outs = _run_workers(tmp_path, "queries")
    assert "QUERIES_OK" in outs[0]
