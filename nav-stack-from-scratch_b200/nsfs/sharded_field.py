"""Row-partitioned cost-to-go field across GPUs (BASELINE config C5).

A map too large (or too slow) for one GPU is cut into contiguous row blocks, one per rank.  The reference's stencil
works on the FLAT key (grid_planner_core.cpp:404-407, 423-426: only the row is bounds-checked), so an edge from
(i, W-1) in direction (+1,+1) lands on (i+2, 0): a shard needs **2 ghost rows** of g on either side.  The direction
costs of a ghost-row cell depend on which of ITS neighbours are passable (astar.cpp:88-123), again up to 2 rows away,
so each shard holds **4 extra rows of map** per side; the outer 2 take no in-edges (``active`` window) and exist only so
that the ghost rows' edge masks equal the ones a monolithic solve would build.

Per solve, all shards advance together band by band (a delta-stepping loop ACROSS GPUs; inside a shard the solver's
own tile scheduler runs to the band's local fixed point):

    begin (the source is seeded on its owner)
    repeat:  send my first / last 2 OWNED rows to the upper / lower neighbour        (2*W floats each way)
             inject what arrived into my ghost rows: g = min(g, received)
             M = all-reduce(min) of every shard's smallest pending value; stop when nothing is pending anywhere
             relax every pending tile up to the cost cap M + band                     (shards with nothing below the cap skip)
    finish (predecessors, reached count over the owned rows)

With band = +inf this degenerates to "each shard to its fixed point, then exchange": correct, but a wave that starts
in one shard is then solved shard after shard.  A finite band keeps every shard that the wavefront currently crosses
busy in the same round.  The exchange is the path's only collective: point-to-point rows + one scalar all-reduce per
round.  Relaxation is monotone and the fp32 fixed point is unique (every value is the min over in-edges of
fl(g[u] + w)), so the assembled result is bit-identical to the single-GPU field whatever the band.

``solver`` objects implement: begin(src|None), inject(row0, rows)->improved, relax(cap|None), min_pending()->float,
finish()->dict, rows(row0, n),
where ``rows`` are 2-D arrays/tensors [n, W] in the solver's own memory space (torch CUDA tensors for the GPU solver).
This module contains no search code: the GPU solver is nsfs.capi.Planner; tests plug a CPU checker in to exercise the
partition / exchange / termination logic under gloo.
"""
from __future__ import annotations

import ctypes

import numpy as np

GHOST = 2        # rows of g exchanged per side
MAP_HALO = 4     # rows of map held per side


class RowPartition:
    """Contiguous, balanced row blocks; every block owns at least GHOST rows."""

    def __init__(self, H, W, world):
        if world < 1 or H < GHOST * world:
            raise ValueError(f"cannot split {H} rows over {world} shards (each needs >= {GHOST} rows)")
        self.H, self.W, self.world = int(H), int(W), int(world)
        base, extra = divmod(self.H, self.world)
        self.bounds = [0]
        for r in range(self.world):
            self.bounds.append(self.bounds[-1] + base + (1 if r < extra else 0))

    def owned(self, r):
        return self.bounds[r], self.bounds[r + 1]

    def held(self, r):
        r0, r1 = self.owned(r)
        return max(0, r0 - MAP_HALO), min(self.H, r1 + MAP_HALO)

    def active(self, r):
        r0, r1 = self.owned(r)
        return max(0, r0 - GHOST), min(self.H, r1 + GHOST)

    def owner_of_row(self, i):
        for r in range(self.world):
            if self.bounds[r] <= i < self.bounds[r + 1]:
                return r
        raise ValueError(i)


class FieldShard:
    """One rank's share: bookkeeping between global rows and the solver's local rows."""

    def __init__(self, part: RowPartition, rank: int, solver):
        self.part, self.rank, self.solver = part, rank, solver
        self.r0, self.r1 = part.owned(rank)
        self.lo, self.hi = part.held(rank)
        self.a0, self.a1 = part.active(rank)
        self.relaxes = 0

    # global row -> local row
    def loc(self, i):
        return i - self.lo

    def start(self, src):
        mine = src is not None and self.r0 <= src[0] < self.r1
        self.solver.begin((src[0] - self.lo, src[1]) if mine else None)

    def boundary_rows(self):
        """(rows for the upper neighbour, rows for the lower neighbour): my first / last GHOST owned rows."""
        up = self.solver.rows(self.loc(self.r0), GHOST) if self.rank > 0 else None
        down = self.solver.rows(self.loc(self.r1 - GHOST), GHOST) if self.rank < self.part.world - 1 else None
        return up, down

    def absorb(self, from_up, from_down):
        """from_up = the upper neighbour's last GHOST owned rows = my ghost rows [r0-GHOST, r0)."""
        improved = 0
        if from_up is not None:
            improved += self.solver.inject(self.loc(self.r0 - GHOST), from_up)
        if from_down is not None:
            improved += self.solver.inject(self.loc(self.r1), from_down)
        return improved

    def relax(self, cap=None):
        self.solver.relax(cap)
        self.relaxes += 1

    def min_pending(self):
        return self.solver.min_pending()

    def failed(self):
        """A reached cell whose expansion throws in the reference (NSFS_E_RANGE): the whole solve stops, like the exception."""
        return getattr(self.solver, "status", 0) != 0

    def finish(self):
        return self.solver.finish()

    def owned_rows(self):
        return self.solver.rows(self.loc(self.r0), self.r1 - self.r0)


# ---- drivers --------------------------------------------------------------------------------------------------
INF = float("inf")
DEFAULT_BAND = 2048.0      # cost units a round may advance; see the module docstring


def _cap(m, band):
    return None if band is None or band == INF else m + band


def solve_in_process(shards, src, band=DEFAULT_BAND, max_rounds=1000000):
    """All shards live in this process (one GPU or a CPU test): lockstep rounds, neighbours read directly."""
    for s in shards:
        s.start(src)
    rounds = 0
    while rounds < max_rounds:
        rows = [s.boundary_rows() for s in shards]
        for k, s in enumerate(shards):
            from_up = rows[k - 1][1] if k > 0 else None
            from_down = rows[k + 1][0] if k + 1 < len(shards) else None
            s.absorb(_snapshot(from_up), _snapshot(from_down))
        pend = [s.min_pending() for s in shards]
        m = min(pend)
        if not m < INF or any(s.failed() for s in shards):
            break
        rounds += 1
        cap = _cap(m, band)
        for s, p in zip(shards, pend):
            if p < INF and (cap is None or p <= cap):
                s.relax(cap)
    return dict(rounds=rounds, results=[s.finish() for s in shards])


def _snapshot(x):
    if x is None:
        return None
    return x.clone() if hasattr(x, "clone") else np.array(x, copy=True)


def solve_distributed(shard: FieldShard, src, dist, device=None, band=DEFAULT_BAND, max_rounds=1000000):
    """One shard per process (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
    import torch

    rank, world = shard.rank, shard.part.world
    on_cuda = device is not None and device.type == "cuda"
    shard.start(src)
    rounds = 0
    as_t = (lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)))
    while rounds < max_rounds:
        up, down = shard.boundary_rows()
        ops, recv_up, recv_down, keep = [], None, None, []
        if rank > 0:
            t = as_t(up).contiguous(); keep.append(t)
            recv_up = torch.empty_like(t)
            ops += [dist.P2POp(dist.isend, t, rank - 1), dist.P2POp(dist.irecv, recv_up, rank - 1)]
        if rank < world - 1:
            t = as_t(down).contiguous(); keep.append(t)
            recv_down = torch.empty_like(t)
            ops += [dist.P2POp(dist.isend, t, rank + 1), dist.P2POp(dist.irecv, recv_down, rank + 1)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if on_cuda:
            torch.cuda.current_stream(device).synchronize()
        shard.absorb(recv_up, recv_down)
        pend = shard.min_pending()
        mine = -1.0 if shard.failed() else (pend if pend < INF else 3.0e38)      # a failed shard stops everybody
        flag = torch.tensor([mine], dtype=torch.float32, device=device if device is not None else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        m = float(flag.item())
        if m >= 3.0e38 or m < 0.0:
            break
        rounds += 1
        cap = _cap(m, band)
        if pend < INF and (cap is None or pend <= cap):
            shard.relax(cap)
    return dict(rounds=rounds, result=shard.finish())


def solve_p2p(shard: FieldShard, src, dist, device):
    """One shard per process / GPU, NO exchange rounds: every shard's solve kernel pushes its boundary rows into the
    neighbour's memory over NVLink and wakes the neighbour's tiles itself (remote atomics); one shared counter of pending
    tiles ends all kernels together.  Host-side work per solve: seed, one all-reduce to prime the counter, two barriers."""
    import torch

    rank, world = shard.rank, shard.part.world
    P = shard.solver.P
    if not getattr(shard, "_p2p_ready", False):
        infos = [None] * world
        dist.all_gather_object(infos, P.field_peer_export())
        lo = [shard.part.held(r)[0] for r in range(world)]
        if rank > 0:
            P.field_peer_attach(0, infos[rank - 1], lo[rank] - lo[rank - 1])
        if rank < world - 1:
            P.field_peer_attach(1, infos[rank + 1], lo[rank] - lo[rank + 1])
        P.field_peer_counter(None if rank == 0 else infos[0])
        P.field_peer_rows(shard.loc(shard.r0), shard.loc(shard.r1))
        shard._p2p_ready = True
    shard.start(src)
    n = torch.tensor([P.field_pending_count()], dtype=torch.int64, device=device)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if rank == 0:
        P.field_set_counter(int(n.item()))
    dist.barrier()                       # the counter is primed before any kernel may decrement it
    r = P.field_relax_p2p()
    shard.solver.stats.append(r["stats"])
    shard.solver.status = shard.solver.status or r["status"]
    shard.relaxes += 1
    dist.barrier()                       # every shard is at the global fixed point
    return dict(rounds=1, result=shard.finish())


# ---- the GPU solver: one nsfs handle over the shard's held rows ----------------------------------------------------
class GpuShardSolver:
    """nsfs.capi.Planner over rows [lo, hi) of the global map; row values travel as torch CUDA tensors."""

    def __init__(self, part: RowPartition, rank: int, infl_rows, lo_rows, device=0, lo_thresh=10):
        import torch
        from . import capi

        self.torch = torch
        self.dev = torch.device("cuda", device)
        lo, hi = part.held(rank)
        r0, r1 = part.owned(rank)
        a0, a1 = part.active(rank)
        self.Hl, self.W = hi - lo, part.W
        assert infl_rows.shape == (self.Hl, self.W)
        self.P = capi.Planner(self.Hl, self.W, device=device)
        # windows are in local rows; "hi == 0" means unset, and a window ending at the last local row is the same as unset
        self.P.set_option("active_row_lo", a0 - lo)
        self.P.set_option("active_row_hi", a1 - lo)
        self.P.set_option("count_row_lo", r0 - lo)
        self.P.set_option("count_row_hi", r1 - lo)
        self.P.set_map(infl_rows, lo_rows, lo_thresh)
        self._g = None
        self.stats = []
        self.status = 0

    def _gview(self):
        if self._g is None:
            ptr, _ = self.P.field_pointers()
            self._g = _tensor_from_ptr(self.torch, ptr, (self.Hl, self.W), "<f4", self.dev)
        return self._g

    def begin(self, src):
        self.P.field_begin(src)
        self._g = None
        self.status = 0

    def inject(self, row0, rows):
        t = rows if isinstance(rows, self.torch.Tensor) else self.torch.from_numpy(np.ascontiguousarray(rows, np.float32))
        t = t.to(self.dev, self.torch.float32).contiguous()
        self.torch.cuda.current_stream(self.dev).synchronize()      # the nsfs handle launches on its own stream
        return self.P.field_inject_rows(row0, t.shape[0], t.data_ptr())

    def relax(self, cap=None):
        r = self.P.field_relax(cap)
        self.stats.append(r["stats"])
        self.status = self.status or r["status"]
        return r

    def min_pending(self):
        return self.P.field_min_pending()

    def finish(self):
        r = self.P.field_finish()
        r["status"] = r["status"] or self.status
        return r

    def rows(self, row0, n):
        return self._gview()[row0:row0 + n]

    def pred_rows(self, row0, n):
        _, ptr = self.P.field_pointers()
        return _tensor_from_ptr(self.torch, ptr, (self.Hl, self.W), "|u1", self.dev)[row0:row0 + n]

    def close(self):
        self._g = None
        self.P.close()


def _tensor_from_ptr(torch, ptr, shape, typestr, device):
    class _Raw:
        pass
    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(raw, device=device)


def field_checksum(torch, g_rows, pred_rows, key0):
    """Position-weighted sums mod 2^64 of the fp32 bit patterns of g and of (pred + 1) over the global keys key0, key0 + 1, ...
    (torch CUDA tensors).  The sums of disjoint row blocks add up (mod 2^64) to the whole field's, so N shards can be checked
    against the single-GPU field without gathering it; integer arithmetic wraps, so the order of summation is irrelevant.
    Returns two Python ints."""
    n = g_rows.numel()
    gi = g_rows.reshape(-1).view(torch.int32)
    pr = pred_rows.reshape(-1)
    sg, sp, chunk = 0, 0, 1 << 25
    for o in range(0, n, chunk):
        e = min(o + chunk, n)
        w = torch.arange(key0 + o, key0 + e, device=gi.device, dtype=torch.int64) * 2 + 1
        sg += int((gi[o:e].to(torch.int64) * w).sum().item())
        sp += int(((pr[o:e].to(torch.int64) + 1) * w).sum().item())
    return sg % (1 << 64), sp % (1 << 64)


def make_gpu_shard(part: RowPartition, rank: int, infl, lo, device=0, lo_thresh=10):
    """Slice the held rows out of global int32 planes (numpy [H, W]) and build the shard."""
    a, b = part.held(rank)
    solver = GpuShardSolver(part, rank, np.ascontiguousarray(infl[a:b]), np.ascontiguousarray(lo[a:b]), device, lo_thresh)
    return FieldShard(part, rank, solver)
