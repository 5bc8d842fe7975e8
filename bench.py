#!/usr/bin/env python
"""bench.py — headline measurement of the B200 grid planner (one JSON line on stdout from rank 0).

Workloads (SURVEY.md §8d / BASELINE.json configs; ``--workload``):
  c2 (default)  Djikstra full cost-to-go field, 2048x2048 raw map p=0.10 seed 5, source = centre      [configs[1]]
  c4            batched Astar "f" start/goal queries, 1024x1024 raw map p=0.10 seed 7                  [configs[3]]
  c1            Astar "f" single queries, 384x384 inflated map seed 1 (latency)                        [configs[0]]
  c3            ThetaStar, 4096x4096 raw map p=0.20 seed 9, 16 start/goal pairs run as one batch        [configs[2]]
  c5            Djikstra field on the 16384x16384 raw map p=0.10 seed 11; with N > 1 ranks the map is
                partitioned by row blocks and 2-row halos are exchanged over NCCL                      [configs[4]]

A step = one pass of the hot path over one batch of synthetic input: one field (c2, c5), one batch of ``--batch``
queries (c4), one query (c1), the 16 queries (c3).  ``value`` is measured with inputs resident in HBM (device time
of the step's kernels, CUDA events on the launching stream, L2 flushed between steps); ``e2e`` goes through the
host-pointer C-ABI calls with host<->device copies (pinned host memory) inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c1|c3|c5] [--impl ours|reference]
N > 1 is launched by torchrun (one rank per GPU, NCCL).  c2/c4/c1/c3: the packed map is broadcast once from rank 0 and
the work units are independent per rank (no data-path collective; weak scaling).  c5: one field, strong scaling.
Timing is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "nav-stack-from-scratch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from nsfs import mapgen  # noqa: E402

BYTES_PER_SETTLED = 9.125        # read g (4) + write g (4) + write predecessor (1) + occupancy bit (0.125)
RELAX_PER_SETTLED = 8            # one settled cell = 8 cell-relaxations (reference loop astar.cpp:89)

WORKLOADS = {
    "c2": dict(H=2048, W=2048, p=0.10, seed=5, inflated=False, algo="djikstra",
               name="Djikstra full cost-to-go field, 2048x2048 raw map p_occ=0.10 seed 5, source = centre (C2)"),
    "c4": dict(H=1024, W=1024, p=0.10, seed=7, inflated=False, algo="astar",
               name="batched Astar f start/goal queries, 1024x1024 raw map p_occ=0.10 seed 7 (C4)"),
    "c1": dict(H=384, W=384, p=0.005, seed=1, inflated=True, algo="astar",
               name="Astar f single query, 384x384 inflated map p_occ=0.005 seed 1 (C1)"),
    "c3": dict(H=4096, W=4096, p=0.20, seed=9, inflated=False, algo="thetastar",
               name="ThetaStar f, 4096x4096 raw map p_occ=0.20 seed 9, 16 start/goal pairs as one batch (C3)"),
    "c5": dict(H=16384, W=16384, p=0.10, seed=11, inflated=False, algo="djikstra",
               name="Djikstra full cost-to-go field, 16384x16384 raw map p_occ=0.10 seed 11, source = centre (C5)"),
}


# ---------------------------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(kernel, units=1):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if there is one.  Kernels whose launch size
    varies (the batch search kernel) are recorded per unit (per query) and scaled by the units of this launch."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        if kernel in d:
            return d[kernel]
        if kernel + "_per_unit" in d:
            return int(d[kernel + "_per_unit"] * units)
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.t0, self.t1, self.proc = [], None, None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def finish(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.08)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 - 0.03 <= t <= self.t1 + 0.06] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_REAL_STDOUT = None


def guard_stdout():
    """Libraries (NCCL prints its version banner) must not write to stdout: the driver reads ONE JSON line from it.
    File descriptor 1 is pointed at stderr for the whole run; emit() writes the JSON line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def c3_queries(infl, lo, seed, n=16):
    """The C3 pairs: starts from the sampler, goals re-drawn a few hundred cells away so a query stays in the seconds."""
    starts = mapgen.sample_free_cells(infl, lo, n, seed + 1)
    free = mapgen.free_mask(infl, lo)
    H, W = infl.shape
    R = 300
    qs = np.zeros((n, 4), np.int32)
    for t, s0 in enumerate(starts):
        i0, j0 = int(np.clip(s0[0], R, H - R - 1)), int(np.clip(s0[1], R, W - R - 1))
        g = mapgen.sample_cells(free[i0 - R:i0 + R, j0 - R:j0 + R], 1, seed + 100 + t)[0]
        qs[t] = (s0[0], s0[1], i0 - R + g[0], j0 - R + g[1])
    return qs


# ---------------------------------------------------------------------------------------------------------------
def cpu_impl(infl, lo):
    from oracle import oracle as orc        # bench.py may execute oracle/ only in the cpu_baseline / --impl reference legs (spec ④)
    have_ref = orc.have_reference()
    return orc, (orc.Reference(infl, lo) if have_ref else orc.Oracle(infl, lo)), ("reference" if have_ref else "port"), have_ref


def reference_arm(args):
    """The reference's own CPU planners (oracle/_ref when present, else the C restatement) on all host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    note = None
    if args.workload == "c5":
        # 153 s per field on one core: the bounded sample is the same generator at 1/16 of the area
        H = W = 4096
        note = "bounded sample: the C5 generator (p_occ=0.10, seed 11) at 4096x4096, one full field per host thread"
    else:
        H, W = wl["H"], wl["W"]
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    orc, impl, kind, have_ref = cpu_impl(infl, lo)
    times, units = [], 0
    if args.workload in ("c2", "c5"):
        # a step = `threads` independent full fields, one per host thread (a single field cannot use more than one core)
        threads = min(cores, 64) if have_ref else 1
        srcs = mapgen.sample_free_cells(infl, lo, threads, wl["seed"] + 17)
        srcs[0] = mapgen.field_source(infl, lo)
        for it in range(args.warmup + args.steps):
            if have_ref:
                r = impl.field_batch(srcs, nthreads=threads)
                sec, settled = r["seconds"], int(r["settled"].sum())
            else:
                t0 = time.perf_counter(); r = impl.field(srcs[0]); sec = time.perf_counter() - t0
                settled = int(r["settled"])
            if it >= args.warmup:
                times.append(sec); units += settled
        value = RELAX_PER_SETTLED * units / sum(times) / 1e9
        metric, unit = "gcell_relax_per_s", "Gcell-relax/s"
        sample = note or f"{threads} full fields per step (one per host thread) on the C2 map, {args.steps} steps"
    else:
        threads = cores if have_ref else 1
        if args.workload == "c3":
            qs = c3_queries(infl, lo, wl["seed"])
            note = ("ThetaStar has no reference implementation (thetastar.cpp is empty): the reference's Astar f + "
                    "post_process_path on the same 16 pairs is the nearest reference behaviour")
        else:
            qs = mapgen.sample_queries(infl, lo, max(cores, 8) if args.workload == "c4" else 8, wl["seed"])
        nq = len(qs)
        for it in range(args.warmup + args.steps):
            if have_ref:
                sec = impl.plan_batch(0, "f", qs, nthreads=threads)["seconds"]
            else:
                t0 = time.perf_counter()
                for q in qs:
                    impl.plan(0, "f", q[:2], q[2:])
                sec = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(sec); units += nq
        value = units / sum(times)
        metric, unit = "planner_queries_per_s", "queries/s"
        sample = note or f"{nq} Astar f queries per step over {threads} host threads, {args.steps} steps"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "strong" if args.workload == "c5" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "threads": threads},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_baseline(workload, infl, lo):
    """Bounded sample of the same workload on the host cores (rank 0, N = 1 only)."""
    wl = WORKLOADS[workload]
    if workload == "c5":
        infl, lo = mapgen.synth_map(4096, 4096, wl["p"], wl["seed"], False)
    orc, impl, kind, have_ref = cpu_impl(infl, lo)
    if workload in ("c2", "c5"):
        src = mapgen.field_source(infl, lo)
        reps, secs, settled = (6 if workload == "c2" else 2), [], 0
        for _ in range(reps):
            t0 = time.perf_counter(); r = impl.field(src); dt = time.perf_counter() - t0
            secs.append(r.get("seconds", dt)); settled = int(r["settled"])
        v = RELAX_PER_SETTLED * settled / float(np.median(secs)) / 1e9
        what = "the C2 source" if workload == "c2" else "the centre of a 4096x4096 map from the C5 generator (1/16 of the area)"
        return {"value": v, "unit": "Gcell-relax/s", "cores": 1, "kind": kind, "settled": settled,
                "ms_per_field": 1e3 * float(np.median(secs)),
                "sample": f"{reps} full Djikstra::plan fields from {what} on one core (median)"}
    if workload == "c3":
        qs = c3_queries(infl, lo, wl["seed"])[:6]
        O = orc.Oracle(infl, lo)
        secs_t, secs_a = [], []
        for q in qs:
            t0 = time.perf_counter(); O.plan(orc.THETASTAR, "f", q[:2], q[2:]); secs_t.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); r = impl.plan(0, "f", q[:2], q[2:]); dt = time.perf_counter() - t0
            secs_a.append(r.get("seconds", dt))
        return {"value": len(qs) / float(np.sum(secs_t)), "unit": "queries/s", "cores": 1, "kind": "port",
                "p50_ms": 1e3 * float(np.median(secs_t)),
                "reference_astar_f_p50_ms": 1e3 * float(np.median(secs_a)), "reference_astar_kind": kind,
                "sample": f"{len(qs)} of the 16 C3 pairs, one at a time on one core: the ThetaStar restatement (no reference "
                          "implementation exists) and, beside it, the reference's Astar f on the same pairs"}
    nq = 64 if workload == "c4" else 200
    qs = mapgen.sample_queries(infl, lo, nq, wl["seed"])
    secs = []
    for q in qs:
        t0 = time.perf_counter(); r = impl.plan(0, "f", q[:2], q[2:]); dt = time.perf_counter() - t0
        secs.append(r.get("seconds", dt))
    return {"value": nq / float(np.sum(secs)), "unit": "queries/s", "cores": 1, "kind": kind,
            "p50_ms": 1e3 * float(np.median(secs)), "p90_ms": 1e3 * float(np.percentile(secs, 90)),
            "sample": f"{nq} Astar::plan f queries one at a time on one core"}


# ---------------------------------------------------------------------------------------------------------------
def ours_c5(args, torch, dist, rank, world, local, dev):
    """One field on the 16384^2 map: a single handle at N = 1, row-block shards with NCCL halo exchange at N > 1."""
    from nsfs import capi, sharded_field as sf

    wl = WORKLOADS["c5"]
    H, W = wl["H"], wl["W"]
    part = sf.RowPartition(H, W, world)
    a, b = part.held(rank)
    infl, lo = mapgen.synth_map_rows(H, W, wl["p"], wl["seed"], a, b)
    # the source: (H/2, W/2) advanced in +j to the first free cell
    mid_infl, mid_lo = (infl[H // 2 - a:H // 2 - a + 1], lo[H // 2 - a:H // 2 - a + 1]) if a <= H // 2 < b else \
        mapgen.synth_map_rows(H, W, wl["p"], wl["seed"], H // 2, H // 2 + 1)
    fm = mapgen.free_mask(mid_infl, mid_lo)[0]
    sj = W // 2
    while not fm[sj]:
        sj += 1
    src = (H // 2, sj)
    shard = sf.FieldShard(part, rank, sf.GpuShardSolver(part, rank, infl, lo, device=local))
    P = shard.solver.P
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        shard.solver.stats.clear()
        if world == 1:
            r = P.field_device((src[0] - a, src[1]))
            return dict(reached=r["reached"], rounds=1, kernel_ms=r["stats"]["main_kernel_ms"], dev_ms=r["stats"]["gpu_ms"],
                        visits=r["stats"]["tile_visits"])
        if args.exchange == "p2p":
            out = sf.solve_p2p(shard, src, dist, dev)
        else:
            out = sf.solve_distributed(shard, src, dist, dev, band=args.band if args.band > 0 else None)
        return dict(reached=out["result"]["reached"], rounds=out["rounds"],
                    kernel_ms=sum(s["main_kernel_ms"] for s in shard.solver.stats),
                    dev_ms=sum(s["gpu_ms"] for s in shard.solver.stats) + out["result"]["stats"]["gpu_ms"],
                    visits=sum(s["tile_visits"] for s in shard.solver.stats))

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = P.launches()
    if sampler:
        sampler.mark_start()
    wall, kernel_ms, reached, last = 0.0, 0.0, 0, None
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        last = step()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
        kernel_ms += last["kernel_ms"]; reached += last["reached"]
    if sampler:
        sampler.mark_stop()
    launches = P.launches() - launches0
    if world > 1:
        t = torch.tensor([wall, kernel_ms, float(reached), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        wall, kernel_ms, reached, launches = tmax[0].item(), tmax[1].item(), tsum[2].item(), int(tsum[3].item())
    # e2e: upload this rank's int32 planes through the host-pointer call, solve, read the owned rows of g and pred back
    pin_i, pin_l = torch.from_numpy(infl).pin_memory(), torch.from_numpy(lo).pin_memory()
    own = part.owned(rank)
    host_g = torch.empty((own[1] - own[0], W), dtype=torch.float32).pin_memory()
    host_p = torch.empty((own[1] - own[0], W), dtype=torch.uint8).pin_memory()

    def step_e2e():
        P.set_map(pin_i.numpy(), pin_l.numpy())
        r = step()
        host_g.copy_(shard.owned_rows()); host_p.copy_(shard.solver.pred_rows(shard.loc(own[0]), own[1] - own[0]))
        torch.cuda.synchronize()
        return r["reached"]

    e2e_steps = max(1, min(args.steps, 3))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_reached = sum(step_e2e() for _ in range(e2e_steps))
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec, float(e2e_reached)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_sec, e2e_reached = tmax[0].item(), tsum[1].item()
    if rank != 0:
        return
    peak, peak_src = peaks()
    settled = reached / args.steps
    alg_bytes = BYTES_PER_SETTLED * settled / world            # per GPU: the dominant kernel of each rank covers its own rows
    kernel_s = kernel_ms / args.steps * 1e-3
    achieved = alg_bytes / kernel_s / 1e9
    main_kernel = "field_solve_async_kernel"
    line = {
        "metric": "gcell_relax_per_s", "value": RELAX_PER_SETTLED * reached / wall / 1e9, "unit": "Gcell-relax/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "units_per_step": "1 field over all ranks",
                   "partition": f"{world} row blocks, 2 ghost rows of g + 4 rows of map per side, " + ("peer-to-peer posts over NVLink inside the solve kernel (no rounds)" if args.exchange == "p2p" else f"NCCL send/recv + 1 all-reduce per round, band {args.band:g}") if world > 1 else "single handle",
                   "l2": "flushed between timed steps (256 MiB fill)",
                   "timing": "wall clock between cuda synchronize + barrier on both sides, max over ranks (the step spans several kernels, NCCL and host round control)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": main_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": kernel_ms / args.steps,
                     "note": "per GPU: that rank's share of the settled cells over the summed time of its solve launches (max over ranks)"},
        "e2e": {"value": RELAX_PER_SETTLED * e2e_reached / e2e_sec / 1e9, "unit": "Gcell-relax/s",
                "h2d_bytes_per_step": int(2 * infl.size * 4 * world), "d2h_bytes_per_step": int(H * W * 5),
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps},
        "gpu_launches": launches, "clocks": sampler.finish() if sampler else None,
        "extra": {"settled_per_field": settled, "exchange_rounds": last["rounds"], "tile_visits_rank0": last["visits"],
                  "fields_per_s": args.steps / wall},
    }
    if world == 1:
        # C5 (ii): 1024 find_better_point calls from random non-free interior cells (replicas only: each search is local)
        bad = mapgen.sample_blocked_interior(infl, lo, 1024, wl["seed"])
        P.find_better_point_batch(bad[:64])
        fb = P.find_better_point_batch(bad)
        line["extra"]["fallback_1024"] = {"queries_per_s": 1024 / (fb["stats"]["gpu_ms"] * 1e-3), "gpu_ms": fb["stats"]["gpu_ms"],
                                          "settled": fb["stats"]["expansions"]}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline("c5", None, None)
    emit(line)


def ours(args):
    import torch
    from nsfs import capi, sharded_queries as sq

    rank, world, local = dist_env()
    guard_stdout()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if args.workload == "c5":
        ours_c5(args, torch, dist, rank, world, local, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    wl = WORKLOADS[args.workload]
    H, W = wl["H"], wl["W"]
    P = capi.Planner(H, W, device=local)

    # ---- inputs: rank 0 packs the synthetic map; other ranks receive the two bit planes over NCCL ----------------
    # (every rank also keeps the int32 host planes: the e2e leg uploads them through nsfs_set_map each step)
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    if rank == 0:
        d_infl = torch.from_numpy(infl).to(dev)
        d_lo = torch.from_numpy(lo).to(dev)
        P.set_map_device(d_infl.data_ptr(), d_lo.data_ptr())
    if world > 1:
        sq.broadcast_packed_map(P, dist, dev, rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- per-rank work units ---------------------------------------------------------------------------------
    if args.workload == "c2":
        if rank == 0:
            srcs = mapgen.sample_free_cells(infl, lo, max(world, 1), wl["seed"] + 17)
            srcs[0] = mapgen.field_source(infl, lo)
            srcs_t = torch.from_numpy(srcs.astype(np.int32)).to(dev)
        else:
            srcs_t = torch.zeros((world, 2), dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(srcs_t, src=0)
        src = tuple(int(x) for x in srcs_t[rank].tolist())

        def step_resident():
            r = P.field_device(src)
            return r["stats"], r["reached"]

        host_infl = torch.from_numpy(infl).pin_memory(); host_lo = torch.from_numpy(lo).pin_memory()
        host_g = torch.empty((H, W), dtype=torch.float32).pin_memory()
        host_pred = torch.empty((H, W), dtype=torch.uint8).pin_memory()

        def step_e2e():
            P.set_map(host_infl.numpy(), host_lo.numpy())
            st = capi.Stats()
            rc = P.L.nsfs_field(P.h, src[0], src[1], host_g.numpy().ctypes.data, host_pred.numpy().ctypes.data, st)
            assert rc == 0
            return st.expansions

        nq = 1
        h2d, d2h = 2 * H * W * 4 + 8, H * W * 4 + H * W
        main_kernel = "field_solve_async_kernel"
    else:
        algo = wl["algo"]
        per_step = {"c4": args.batch, "c1": 1, "c3": 16}[args.workload]
        nq = per_step
        rotating = args.workload == "c1"                    # c1: a different query every step
        total = nq * world * ((args.steps + args.warmup) if rotating else 1)
        if rank == 0:
            allq = c3_queries(infl, lo, wl["seed"], 16 * world) if args.workload == "c3" else mapgen.sample_queries(infl, lo, total, wl["seed"])
            allq_t = torch.from_numpy(np.ascontiguousarray(allq)).to(dev)
        else:
            allq_t = torch.zeros((total, 4), dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(allq_t, src=0)
        a, b = sq.query_slice(total, world, rank)
        myq = allq_t[a:b].contiguous()
        per = myq.shape[0]
        d_status = torch.zeros(nq, dtype=torch.int32, device=dev); d_g = torch.zeros(nq, dtype=torch.float64, device=dev)
        d_len = torch.zeros(nq, dtype=torch.int32, device=dev); d_set = torch.zeros(nq, dtype=torch.int64, device=dev)
        counter = {"i": 0}

        def pick(t):
            if not rotating:
                return t
            k = counter["i"] % per
            counter["i"] += 1
            return t[k:k + 1]

        def step_resident():
            q = pick(myq)
            st = P.plan_batch_device(algo, "f", nq, q.data_ptr(), d_status.data_ptr(), d_g.data_ptr(), d_len.data_ptr(),
                                     d_set.data_ptr(), want_stats=True)
            return st, int(d_set.sum().item())

        host_q = myq.cpu().pin_memory().numpy()

        def step_e2e():
            r = P.plan_batch(algo, "f", pick(host_q), path_cap=0)
            return int(r["settled"].sum())

        h2d, d2h = nq * 16, nq * (4 + 8 + 4 + 8)
        main_kernel = "search_team_kernel" if args.workload == "c4" else ("search_team_kernel(batch)" if nq > 1 and algo != "thetastar" else "search_kernel")

    # ---- warm-up, then exactly K timed steps (device time of each step's kernels; L2 flushed between steps) ----
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = P.launches()
    if sampler:
        sampler.mark_start()
    t_wall0 = time.perf_counter()
    dev_ms, main_ms, settled_total, last_stats, per_step_ms = 0.0, 0.0, 0, None, []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        st, settled = step_resident()
        dev_ms += st["gpu_ms"]; main_ms += st["main_kernel_ms"]; settled_total += settled; last_stats = st
        per_step_ms.append(st["gpu_ms"])
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.mark_stop()
    launches = P.launches() - launches0
    if world > 1:
        dist.barrier()
        t = torch.tensor([dev_ms, main_ms, float(settled_total), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms_max, main_ms_max = tmax[0].item(), tmax[1].item()
        settled_all, launches_all = tsum[2].item(), int(tsum[3].item())
    else:
        dev_ms_max, main_ms_max, settled_all, launches_all = dev_ms, main_ms, float(settled_total), launches

    # ---- e2e through the host-pointer API: every rank uploads its inputs and reads its results each step --------
    e2e_warm = 2 if args.workload in ("c2", "c1") else 0
    for _ in range(e2e_warm):
        step_e2e()
    torch.cuda.synchronize()
    e2e_steps = max(3, min(args.steps, 20)) if args.workload in ("c2", "c1") else max(1, min(args.steps, 2))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_settled = 0
    for _ in range(e2e_steps):
        e2e_settled += step_e2e()
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec, float(e2e_settled)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_sec, e2e_settled = tmax[0].item(), tsum[1].item()
    if rank != 0:
        dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    ms_per_step = dev_ms_max / args.steps
    if args.workload == "c2":
        metric, unit = "gcell_relax_per_s", "Gcell-relax/s"
        value = RELAX_PER_SETTLED * settled_all / (dev_ms_max * 1e-3) / 1e9
        e2e_value = RELAX_PER_SETTLED * e2e_settled / e2e_sec / 1e9
        extra = {"fields_per_s": world * args.steps / (dev_ms_max * 1e-3), "settled_per_field": settled_all / (args.steps * world),
                 "rounds": last_stats["rounds"], "tile_visits": last_stats["tile_visits"], "p50_ms_single_query": float(np.median(per_step_ms))}
    else:
        metric, unit = "planner_queries_per_s", "queries/s"
        value = world * nq * args.steps / (dev_ms_max * 1e-3)
        e2e_value = world * nq * e2e_steps / e2e_sec
        extra = {"gcell_relax_per_s": RELAX_PER_SETTLED * settled_all / (dev_ms_max * 1e-3) / 1e9,
                 "settled_per_query": settled_all / (world * nq * args.steps), "batch": nq, "pops": last_stats["pops"],
                 "los_checks": last_stats["los_checks"], "p50_ms_per_step": float(np.median(per_step_ms))}
    alg_bytes = BYTES_PER_SETTLED * settled_all / (args.steps * world)          # per launch of the dominant kernel
    achieved = alg_bytes / (main_ms_max / args.steps * 1e-3) / 1e9
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.workload == "c2" else "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "units_per_step_per_gpu": "1 field" if args.workload == "c2" else f"{nq} queries",
                   "l2": "flushed between timed steps (256 MiB fill)", "map_broadcast": "NCCL, packed bit planes" if world > 1 else "n/a",
                   "timing": "sum of per-step device time (CUDA events on the launching stream), max over ranks",
                   "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(main_kernel, nq), "kernel": main_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": main_ms_max / args.steps},
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps},
        "gpu_launches": launches_all,
        "clocks": sampler.finish() if sampler else None,
        "extra": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, infl, lo)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=65536, help="queries per step per GPU (c4): the 65 536 queries BASELINE names, about four waves of the 16 576 query slots (28 warps x 4 teams x 148 SMs)")
    ap.add_argument("--band", type=float, default=2048.0, help="c5 at N > 1: cost units all shards advance per exchange round (0 = unbounded)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "rounds"],
                    help="c5 at N > 1: p2p = shards post boundary rows into each other's memory over NVLink from inside the solve "
                         "kernel (one launch per GPU); rounds = NCCL send/recv + all-reduce per band")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # defaults that finish within minutes: steps take 3.5 ms (c2), ~0.1 s (c1), seconds (c4, c3, c5)
    if args.steps is None:
        args.steps = {"c2": 30, "c1": 30}.get(args.workload, 2) if args.impl == "ours" else 2
    if args.warmup is None:
        args.warmup = {"c2": 5, "c1": 5}.get(args.workload, 3) if args.impl == "ours" else 1
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
