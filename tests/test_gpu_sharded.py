"""Row-partitioned field on the GPU (SURVEY.md §8e / BASELINE C5): several shards of one map, each its own nsfs handle
over its held rows, solved in lockstep on one device, against the single-handle field.  The field is solved in exact integer
arithmetic, so the bar is bit-identical g, identical predecessors and reached counts."""
import numpy as np
import pytest

from nsfs import mapgen, sharded_field as sf

pytestmark = pytest.mark.gpu


def _mono(gpu_lib, infl, lo, src):
    P = gpu_lib.Planner(*infl.shape)
    P.set_map(infl, lo)
    r = P.field(src)
    P.close()
    return r


@pytest.mark.parametrize("spec", [(700, 520, 0.10, 51, True, 3), (260, 300, 0.22, 52, True, 4), (300, 130, 0.12, 53, False, 2)],
                         ids=["3-shards", "4-shards-dense", "2-shards-unframed-wrap"])
@pytest.mark.parametrize("field_mode", [0, 1], ids=["residency-by-size", "8-tiles-per-sm"])
def test_sharded_field_is_bit_identical_to_one_handle(gpu_lib, spec, field_mode, monkeypatch):
    monkeypatch.setenv("NSFS_FIELD_MODE", str(field_mode))       # read by nsfs_create: every handle below uses that solver
    H, W, p, seed, frame, world = spec
    infl, lo = mapgen.synth_map(H, W, p, seed, False, frame)
    if not frame:
        for c in ((0, 0), (1, 0), (H - 1, W - 1), (H - 2, W - 1)):     # keep the wrap, drop the four throwing cells
            infl[c] = 1
            lo[c] = 20
    src = mapgen.sample_free_cells(infl, lo, 1, seed)[0]
    want = _mono(gpu_lib, infl, lo, src)
    assert want["status"] == 0
    part = sf.RowPartition(H, W, world)
    shards = [sf.make_gpu_shard(part, r, infl, lo) for r in range(world)]
    out = sf.solve_in_process(shards, src, band=[None, 64.0, 300.0][seed % 3])
    assert all(r["status"] == 0 for r in out["results"])
    g = np.concatenate([s.owned_rows().cpu().numpy() for s in shards])
    pred = np.concatenate([s.solver.pred_rows(s.loc(s.r0), s.r1 - s.r0).cpu().numpy() for s in shards])
    assert np.array_equal(g, want["g"])
    assert np.array_equal(pred, want["pred"])
    assert sum(r["reached"] for r in out["results"]) == want["reached"]
    assert out["rounds"] >= 2
    for s in shards:
        s.solver.close()


def test_shard_boundaries_on_tile_boundaries(gpu_lib):
    """Shard edges that coincide with the solver's 64-cell tile edges: an injected ghost row must wake the tiles that
    read it as halo, not only the tile that holds it (16384 rows over 2/4/8 ranks is exactly this case)."""
    infl, lo = mapgen.synth_map(256, 192, 0.10, 55)
    src = tuple(mapgen.sample_free_cells(infl[130:250], lo[130:250], 1, 55)[0] + np.array([130, 0]))
    want = _mono(gpu_lib, infl, lo, src)
    for world in (2, 4):
        part = sf.RowPartition(256, 192, world)
        assert all(b % 64 == 0 for b in part.bounds)
        shards = [sf.make_gpu_shard(part, r, infl, lo) for r in range(world)]
        out = sf.solve_in_process(shards, src)
        g = np.concatenate([s.owned_rows().cpu().numpy() for s in shards])
        assert np.array_equal(g, want["g"])
        assert sum(r["reached"] for r in out["results"]) == want["reached"]
        for s in shards:
            s.solver.close()


def test_source_on_a_tile_edge_with_its_tile_walled_off(gpu_lib, oracle_mod):
    """The source cell never changes during a visit, so the tiles that read it as halo must be seeded explicitly."""
    infl, lo = mapgen.synth_map(192, 192, 0.0, 56)
    src = (64, 70)                                   # first row of tile row 1
    for c in ((64, 69), (64, 71), (65, 69), (65, 70), (65, 71)):      # every in-tile neighbour blocked: only row 63 is open
        infl[c] = 1
        lo[c] = 20
    P = gpu_lib.Planner(192, 192)
    P.set_map(infl, lo)
    r = P.field(src)
    o = oracle_mod.Oracle(infl, lo).field(src)
    assert r["status"] == 0 and o["status"] == 0
    assert np.array_equal(r["g"] < 1e5, o["g"] < 1e5) and r["reached"] == int((o["g"] < 1e5).sum()) > 30000
    assert np.abs(r["g"] - o["g"]).max() <= 1e-5 * o["g"][o["g"] < 1e5].max()
    P.close()


def test_shard_without_source_and_unreachable_rows(gpu_lib):
    """A wall across the map: the shards beyond it stay at 1e5 and the loop still terminates."""
    infl, lo = mapgen.synth_map(200, 160, 0.05, 54)
    infl[120, :] = 1
    lo[120, :] = 20
    src = (30, 80) if mapgen.free_mask(infl, lo)[30, 80] else tuple(mapgen.sample_free_cells(infl[:100], lo[:100], 1, 54)[0])
    want = _mono(gpu_lib, infl, lo, src)
    part = sf.RowPartition(200, 160, 4)
    shards = [sf.make_gpu_shard(part, r, infl, lo) for r in range(4)]
    out = sf.solve_in_process(shards, src)
    g = np.concatenate([s.owned_rows().cpu().numpy() for s in shards])
    assert np.array_equal(g, want["g"]) and (g[121:] >= 1e5).all()
    assert out["results"][3]["reached"] == 0
    for s in shards:
        s.solver.close()


def test_peer_to_peer_shards_over_nvlink():
    """N >= 2 GPUs: one solve kernel per GPU, boundary rows and tile keys posted into the neighbour's memory with remote
    atomics, one shared pending counter (nsfs_field_relax_p2p).  Bit-identical to the single-handle field."""
    import os
    import subprocess
    import sys

    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs on one node")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tests", "p2p_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "P2P_OK" in out.stdout, out.stdout[-3000:]


def test_group_entry_points_over_nccl():
    """N >= 2 GPUs: nsfs_group_* (map broadcast, sharded batch with gathered results, row-block field with CUDA IPC attach) and
    the C++ nsfs_host::DeviceGroup / planBatchSharded above them, against the single-GPU calls: identical arrays."""
    import os
    import subprocess
    import sys

    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs on one node")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29300 + os.getpid() % 300), os.path.join(root, "tests", "group_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "GROUP_OK" in out.stdout, out.stdout[-3000:]
