"""One rank of the multi-GPU group test (launched through torchrun by tests/test_gpu_sharded.py, >= 2 GPUs): the C-ABI
nsfs_group_* entry points (NCCL bound inside libnsfs_gpu.so) and the C++ nsfs_host::DeviceGroup / planBatchSharded above them.
torch.distributed only carries the 128-byte group id and gathers arrays for the comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np, torch, torch.distributed as dist
from nsfs import mapgen, capi, host as nhost, sharded_field as sf
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
G = capi.Group(dist, rank, world, local)

# ---- batched queries: map on rank 0 only, broadcast, slices planned per rank, results gathered everywhere -----------------
H, W = 256, 200
infl, lo = mapgen.synth_map(H, W, 0.12, 61)
qs = mapgen.sample_queries(infl, lo, 301, 61)                      # not divisible by the world size: ragged slices
P = capi.Planner(H, W, device=local)
if rank == 0:
    P.set_map(infl, lo)
G.broadcast_map(P, root=0)
for algo, mode in (("astar", "f"), ("astar", "fg"), ("djikstra", "g")):
    got = G.plan_batch(P, algo, mode, qs)
    if rank == 0:
        want = P.plan_batch(algo, mode, qs)
        same = all(np.array_equal(got[k], want[k]) for k in ("status", "g_goal", "path_len", "settled"))
        ok = ok and same
        print(f"group plan_batch {algo}/{mode} world {world}: identical={same}", flush=True)
P.close()
# the same through the C++ child (GpuPlannerBase::planBatchSharded); rank != 0 hold an EMPTY map: only the root's MapData counts
so = os.path.join(ROOT, "oracle", "_ref", "libnsfs_dropin.so")
so = so if os.path.exists(so) else None
blank = np.zeros_like(infl)
S = nhost.Session("astar", "f", infl if rank == 0 else blank, lo if rank == 0 else blank, device=local, so_path=so)
GC = nhost.make_group(dist, rank, world, local, so_path=so)
got = S.plan_batch_sharded(GC, qs)
if rank == 0:
    want = S.plan_batch_idx(qs)
    same = all(np.array_equal(got[k], want[k]) for k in ("status", "g_goal", "path_len", "settled"))
    ok = ok and same
    print(f"C++ planBatchSharded world {world}: identical={same}", flush=True)
nhost.destroy_group(GC, so_path=so); S.close()

# ---- row-block field: IPC attach + one solve kernel per GPU through nsfs_group_field_* -------------------------------------
for (H, W, p, seed) in [(512, 384, 0.10, 71), (640, 200, 0.2, 73), (1024, 1536, 0.10, 75)]:
    infl, lo = mapgen.synth_map(H, W, p, seed)
    src = tuple(int(x) for x in mapgen.sample_free_cells(infl, lo, 1, seed)[0])
    part = sf.RowPartition(H, W, world)
    if seed == 73:                                    # the source in the last owned row of rank 0 (see tests/p2p_worker.py)
        row = part.bounds[1] - 1
        src = (row, int(np.nonzero(mapgen.free_mask(infl, lo)[row])[0][5]))
    shard = sf.make_gpu_shard(part, rank, infl, lo, device=local)
    Ps = shard.solver.P
    G.field_attach(Ps, shard.lo, shard.r0, shard.r1)
    for rep in range(2):
        r = G.field_solve(Ps, src)
    g = shard.owned_rows().cpu().numpy(); pred = shard.solver.pred_rows(shard.loc(shard.r0), shard.r1 - shard.r0).cpu().numpy()
    parts = [None] * world
    dist.all_gather_object(parts, (g, pred, r["reached_total"], r["status"]))
    if rank == 0:
        P1 = capi.Planner(H, W, device=local); P1.set_map(infl, lo); want = P1.field(src); P1.close()
        Gs = np.concatenate([x[0] for x in parts]); PR = np.concatenate([x[1] for x in parts])
        same = np.array_equal(Gs, want["g"]) and np.array_equal(PR, want["pred"]) and all(x[2] == want["reached"] for x in parts)
        ok = ok and same
        print(f"group field {H}x{W} world {world}: identical={same} statuses={[x[3] for x in parts]}", flush=True)
    Ps.field_peer_detach(); shard.solver.close()
    dist.barrier()
G.close()
if rank == 0:
    print("GROUP_OK" if ok else "GROUP_MISMATCH", flush=True)
dist.destroy_process_group()
