"""Batched start/goal queries sharded across ranks (BASELINE config C4).

Queries are independent, so there is no exchange during the search: rank r plans the contiguous slice
``query_slice(nq, world, r)`` on its own GPU against the same map (rank 0 packs the map once and broadcasts the two
bit planes, N/4 bytes), and the per-query results are concatenated in rank order.  ``plan_fn(sg4_slice) -> dict of
per-query numpy arrays`` is the local planner (nsfs.capi.Planner.plan_batch on a GPU rank).
"""
from __future__ import annotations

import numpy as np


def query_slice(nq, world, rank):
    """Contiguous, balanced: the first ``nq % world`` ranks take one extra query."""
    base, extra = divmod(int(nq), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_packed_map(planner, dist, device, rank, src=0):
    """Ship the packed free / inflated bit planes of ``planner`` on rank ``src`` to every other rank's planner."""
    import torch
    from .sharded_field import _tensor_from_ptr

    words = planner.map_words()
    planes = torch.empty(2 * words, dtype=torch.int32, device=device)
    if rank == src:
        f_ptr, i_ptr = planner.packed_map_device()
        planes[:words].copy_(_tensor_from_ptr(torch, f_ptr, (words,), "<i4", device))
        planes[words:].copy_(_tensor_from_ptr(torch, i_ptr, (words,), "<i4", device))
    dist.broadcast(planes, src=src)
    torch.cuda.current_stream(device).synchronize()
    if rank != src:
        planner.set_packed_map_device(planes.data_ptr(), planes[words:].data_ptr())
    return planes


def plan_sharded(plan_fn, sg4, dist=None, rank=0, world=1, gather=True):
    """Run this rank's slice; with ``gather`` every rank returns the full per-query arrays in query order."""
    sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
    a, b = query_slice(len(sg4), world, rank)
    local = plan_fn(sg4[a:b])
    if world == 1 or dist is None or not gather:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, {k: v for k, v in local.items() if isinstance(v, np.ndarray)})
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
