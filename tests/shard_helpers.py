"""CPU stand-ins that let the multi-rank host logic (nsfs/sharded_field.py, nsfs/sharded_queries.py) run without a GPU.

TEST INFRASTRUCTURE: the local solver here is the oracle's warm-started fixed point (oracle/nsfs_oracle.c
nso_field_relax_window) in fp64 with slack 0, so a sharded solve and a monolithic solve have the same unique fixed
point and can be compared bit for bit."""
import numpy as np

from nsfs import sharded_field as sf

INF = 1e5


class OracleShardSolver:
    def __init__(self, oracle_mod, part, rank, infl, lo):
        a, b = part.held(rank)
        self.O = oracle_mod.Oracle(np.ascontiguousarray(infl[a:b]), np.ascontiguousarray(lo[a:b]))
        self.a0, self.a1 = (x - a for x in part.active(rank))
        self.r0, self.r1 = (x - a for x in part.owned(rank))
        self.g = None
        self.status = 0

    def begin(self, src):
        self.g = np.full((self.O.H, self.O.W), INF, np.float64)
        if src is not None:
            self.g[src[0], src[1]] = 0.0

    def inject(self, row0, rows):
        rows = np.asarray(rows.numpy() if hasattr(rows, "numpy") else rows, np.float64)
        cur = self.g[row0:row0 + len(rows)]
        better = rows < cur
        cur[better] = rows[better]
        return int(better.sum())

    def relax(self, cap=None):
        rc = self.O.field_relax_window(self.g, self.a0, self.a1, 0.0, 1e300 if cap is None else cap)
        if rc < 0:
            self.status = int(rc)

    def min_pending(self):
        m = self.O.field_min_pending(self.g, self.a0, self.a1, 0.0)
        return m if m < 1e299 else float("inf")

    def finish(self):
        return dict(status=self.status, reached=int((self.g[self.r0:self.r1] < INF).sum()))

    def rows(self, row0, n):
        return self.g[row0:row0 + n]


def make_oracle_shards(oracle_mod, infl, lo, world):
    part = sf.RowPartition(infl.shape[0], infl.shape[1], world)
    return part, [sf.FieldShard(part, r, OracleShardSolver(oracle_mod, part, r, infl, lo)) for r in range(world)]


def monolithic(oracle_mod, infl, lo, src):
    O = oracle_mod.Oracle(infl, lo)
    g = np.full(infl.shape, INF, np.float64)
    g[src[0], src[1]] = 0.0
    rc = O.field_relax_window(g, 0, infl.shape[0], 0.0, 1e300)
    return rc, g
