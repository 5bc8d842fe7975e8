"""One rank of the multi-GPU peer-to-peer test (launched through torchrun by tests/test_gpu_sharded.py, >= 2 GPUs):
the row-partitioned field whose shards post into each other's memory over NVLink, against the single-handle field."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np, torch, torch.distributed as dist
from nsfs import mapgen, capi, sharded_field as sf
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for (H, W, p, seed, frame) in [(512, 384, 0.10, 71, True), (256, 192, 0.12, 72, True), (640, 200, 0.2, 73, True), (300, 130, 0.12, 74, False), (1024, 1536, 0.10, 75, True)]:
    infl, lo = mapgen.synth_map(H, W, p, seed, False, frame)
    if not frame:
        for c in ((0, 0), (1, 0), (H - 1, W - 1), (H - 2, W - 1)):
            infl[c] = 1; lo[c] = 20
    src = tuple(mapgen.sample_free_cells(infl, lo, 1, seed)[0])
    part = sf.RowPartition(H, W, world)
    if seed in (71, 75):
        # the source in the FIRST owned row of rank 1 (a shard's boundary row: it never "changes" in a visit, so it reaches the
        # neighbour's ghost row only through the boundary push that precedes the solve; 16384 rows over 2 ranks is this case)
        row = part.bounds[1] + (1 if seed == 75 else 0)
        src = (row, int(np.nonzero(mapgen.free_mask(infl, lo)[row])[0][W // 3 % max(1, int(mapgen.free_mask(infl, lo)[row].sum()))]))
    shard = sf.make_gpu_shard(part, rank, infl, lo, device=local)
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        out = sf.solve_p2p(shard, src, dist, dev)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    g = shard.owned_rows().cpu().numpy(); pred = shard.solver.pred_rows(shard.loc(shard.r0), shard.r1 - shard.r0).cpu().numpy()
    parts = [None] * world
    dist.all_gather_object(parts, (g, pred, out["result"]["reached"], out["result"]["status"], shard.solver.stats[-1]["main_kernel_ms"]))
    if rank == 0:
        P = capi.Planner(H, W, device=local); P.set_map(infl, lo); want = P.field(src); P.close()
        G = np.concatenate([x[0] for x in parts]); PR = np.concatenate([x[1] for x in parts])
        same = np.array_equal(G, want["g"]) and np.array_equal(PR, want["pred"]) and sum(x[2] for x in parts) == want["reached"]
        ok = ok and same
        if not same:
            bad = np.argwhere(G != want["g"])
            print(f"  g differs in {len(bad)} cells, rows {bad[:, 0].min() if len(bad) else -1}..{bad[:, 0].max() if len(bad) else -1}, cols "
                  f"{bad[:, 1].min() if len(bad) else -1}..{bad[:, 1].max() if len(bad) else -1}; shard bounds {part.bounds}; src {src}; first: "
                  + ", ".join(f"({r},{c}) got {G[r, c]:.4f} want {want['g'][r, c]:.4f}" for r, c in bad[:6])
                  + f"; pred differs in {(PR != want['pred']).sum()}; reached {sum(x[2] for x in parts)} vs {want['reached']}", flush=True)
        print(f"{H}x{W} world {world}: identical={same} statuses={[x[3] for x in parts]} wall {dt*1e3:.2f} ms kernel_ms {[round(x[4],2) for x in parts]} single {want['stats']['main_kernel_ms']:.2f} ms", flush=True)
    shard.solver.P.field_peer_detach(); shard.solver.close()
    dist.barrier()
if rank == 0:
    print("P2P_OK" if ok else "P2P_MISMATCH", flush=True)
dist.destroy_process_group()
