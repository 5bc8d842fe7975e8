"""One rank of the gloo world_size-2 tests (launched by tests/test_sharded_cpu.py; reads RANK / WORLD_SIZE / MASTER_*)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nav-stack-from-scratch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

from nsfs import mapgen, sharded_field as sf, sharded_queries as sq  # noqa: E402
from oracle import oracle as orc  # noqa: E402   (checker standing in for the GPU solver)
from shard_helpers import INF, OracleShardSolver, monolithic  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if mode == "field":
        infl, lo = mapgen.synth_map(90, 70, 0.15, 41)
        src = mapgen.sample_free_cells(infl, lo, 1, 41)[0]
        part = sf.RowPartition(90, 70, world)
        shard = sf.FieldShard(part, rank, OracleShardSolver(orc, part, rank, infl, lo))
        out = sf.solve_distributed(shard, src, dist, band=25.0)
        parts = [None] * world
        dist.all_gather_object(parts, (np.array(shard.owned_rows()), out["result"]["reached"], out["rounds"]))
        if rank == 0:
            rc, want = monolithic(orc, infl, lo, src)
            got = np.concatenate([p[0] for p in parts])
            assert rc >= 0 and np.array_equal(got, want)
            assert sum(p[1] for p in parts) == int((want < INF).sum())
            assert len({p[2] for p in parts}) == 1           # every rank left the loop in the same round
            print("FIELD_OK rounds", parts[0][2])
    else:
        infl, lo = mapgen.synth_map(64, 64, 0.12, 42)
        qs = mapgen.sample_queries(infl, lo, 11, 42)           # 11 queries over 2 ranks: 6 + 5
        O = orc.Oracle(infl, lo)

        def plan_fn(sg4):
            rs = [O.plan(orc.ASTAR, "f", q[:2], q[2:]) for q in sg4]
            return dict(g_goal=np.array([r["g_goal"] for r in rs]), settled=np.array([r["settled"] for r in rs], np.int64),
                        path_len=np.array([len(r["path"]) for r in rs], np.int32))

        got = sq.plan_sharded(plan_fn, qs, dist, rank, world)
        if rank == 0:
            want = plan_fn(qs)
            assert all(np.array_equal(got[k], want[k]) for k in want)
            print("QUERIES_OK", len(qs))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
