#!/usr/bin/env python
"""bench.py — headline measurement of the B200 grid planner (one JSON line on stdout from rank 0).

Default (what the driver runs at every N): BASELINE config C4, STRONG scaling —
  65 536 Astar "f" start/goal queries on the 1024x1024 raw map (p=0.10, seed 7), `query_slice` per rank, the packed map
  broadcast once over NCCL, every rank the same share of the same workload, results gathered on rank 0.
  metric planner_queries_per_s at every N.  The same line carries, under "extra":
    c2   Djikstra full field 2048x2048 (Gcell-relax/s, p50 ms per field, plugin-path e2e pageable / pinned)   [rank 0]
    c5   Djikstra full field 16384x16384, one handle at N = 1, row blocks + NVLink peer-to-peer posts at N > 1,
         with a position-weighted checksum of g and pred that must equal the committed single-GPU one

Other workloads for single measurements (``--workload``):
  c2   Djikstra full cost-to-go field, 2048x2048 raw map p=0.10 seed 5, source = centre                 [configs[1]]
  c1   Astar "f" single queries, 384x384 inflated map seed 1 (latency)                                  [configs[0]]
  c3   ThetaStar, 4096x4096 raw map p=0.20 seed 9, the 16 sampler pairs run as one batch                [configs[2]]
  c5   Djikstra field on the 16384x16384 raw map p=0.10 seed 11 (N > 1: row blocks, p2p or NCCL rounds) [configs[4]]
  c4w  the C4 queries with 65 536 queries PER GPU (weak scaling; round-1 lines)
  lat  single-query latency on the 2048x2048 map (north_star's latency target): the three ways a child serves ONE query

A step = one pass of the hot path over one batch of synthetic input.  ``value`` is measured with inputs resident in HBM
(device time of the step's kernels, CUDA events on the launching stream, L2 flushed between steps); ``e2e`` goes through
the reference-facing C++ child (GpuAstar::planBatch / GpuDjikstra::generatePath over the nsfsh_* shim) with the map in
pageable std::vector<int> planes, host<->device copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ...] [--impl ours|reference]
N > 1 is launched by torchrun (one rank per GPU, NCCL).  Timing is the max over ranks.
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
C4_TOTAL = 65536
C5_SETTLED = 241537616           # reference Djikstra settled count on the C5 map (SURVEY.md §6 probe)
C5_CRC_FILE = os.path.join(ROOT, "tests", "golden", "c5_field_checksum.json")

WORKLOADS = {
    "c2": dict(H=2048, W=2048, p=0.10, seed=5, inflated=False, algo="djikstra",
               name="Djikstra full cost-to-go field, 2048x2048 raw map p_occ=0.10 seed 5, source = centre (C2)"),
    "c4": dict(H=1024, W=1024, p=0.10, seed=7, inflated=False, algo="astar",
               name="batched 65,536 Astar f start/goal queries, 1024x1024 raw map p_occ=0.10 seed 7, sharded over the GPUs (C4)"),
    "c4w": dict(H=1024, W=1024, p=0.10, seed=7, inflated=False, algo="astar",
                name="batched Astar f start/goal queries, 1024x1024 raw map p_occ=0.10 seed 7, 65,536 per GPU (C4, weak)"),
    "c1": dict(H=384, W=384, p=0.005, seed=1, inflated=True, algo="astar",
               name="Astar f single query, 384x384 inflated map p_occ=0.005 seed 1 (C1)"),
    "c3": dict(H=4096, W=4096, p=0.20, seed=9, inflated=False, algo="thetastar",
               name="ThetaStar f, 4096x4096 raw map p_occ=0.20 seed 9, the 16 sampler start/goal pairs as one batch (C3)"),
    "lat": dict(H=2048, W=2048, p=0.10, seed=5, inflated=False, algo="djikstra",
                name="single start/goal query latency on the 2048x2048 raw map p_occ=0.10 seed 5: Djikstra (field backend / exact order) and Astar f (exact order)"),
    "c5": dict(H=16384, W=16384, p=0.10, seed=11, inflated=False, algo="djikstra",
               name="Djikstra full cost-to-go field, 16384x16384 raw map p_occ=0.10 seed 11, source = centre (C5)"),
}


def shared_config(workload):
    """The SAME dict in both arms (the driver compares them): what is run; how each arm measures it is spelled out once."""
    wl = WORKLOADS[workload]
    cfg = {"workload": wl["name"],
           "l2": "GPU arm: flushed between timed steps (256 MiB fill); CPU arm: n/a",
           "timing": "GPU arm: device time of each step's kernels (CUDA events on the launching stream), max over ranks; "
                     "reference arm: steady_clock around the reference's own plan() calls on all host threads"}
    if workload == "c4":
        cfg["queries_per_step"] = "GPU arm: all 65,536, query_slice per rank; reference arm: a bounded sample of the same query stream per step"
    return cfg


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
            if len(self.rows) > 4000:                        # long runs: keep every other sample
                self.rows = self.rows[::2]

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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def c3_queries(infl, lo, seed, n=16):
    """The C3 pairs as BASELINE states them: 16 consecutive (start, goal) pairs of the free-cell sampler on the 4096^2 map."""
    return mapgen.sample_queries(infl, lo, n, seed)


def host_so():
    """The C++ children the plugin-path e2e goes through: the build against the reference's OWN grid_planner_core.h when it
    exists (oracle/_ref/libnsfs_dropin.so), else the in-repo build against the interface mirror."""
    p = os.path.join(ROOT, "oracle", "_ref", "libnsfs_dropin.so")
    return (p, "GridPlannerCore children built against the reference's own headers (libnsfs_dropin.so)") if os.path.exists(p) else \
        (None, "GridPlannerCore children built against the interface mirror (libnsfs_planners.so)")


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
    if args.workload == "c5":
        # one full field is 153 s and ~13 GB on one core and a field cannot use a second core: K + W of them do
        # not fit the run's time limit.  No shrunken stand-in is reported as if it were C5.
        emit({"impl": "reference", "unavailable": "C5 unmeasured here: one 16384^2 Djikstra::plan field takes ~153 s on one host core "
                                                  "(SURVEY.md §6); the default arm (C4) carries the reference comparison"})
        return
    H, W = wl["H"], wl["W"]
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    orc, impl, kind, have_ref = cpu_impl(infl, lo)
    times, units = [], 0
    if args.workload == "c2":
        # a step = `threads` independent full fields from the C2 source, one per host thread (a single field cannot use
        # more than one core); every thread does the same unit of work
        threads = min(cores, 64) if have_ref else 1
        srcs = np.repeat(np.asarray([mapgen.field_source(infl, lo)], np.int32), threads, 0)
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
        sample = f"{threads} full fields from the C2 source per step (one per host thread), {args.steps} steps"
    else:
        threads = cores if have_ref else 1
        note = None
        if args.workload == "c3":
            qs = c3_queries(infl, lo, wl["seed"])
            note = ("ThetaStar has no reference implementation (thetastar.cpp is empty): the reference's Astar f on the same 16 "
                    "pairs is the nearest reference behaviour")
        elif args.workload in ("c4", "c4w"):
            # bounded sample of the SAME query stream: 4 queries per host thread per step (~0.14 s each), a fresh slice every step
            per_step = 4 * threads
            allq = mapgen.sample_queries(infl, lo, per_step * (args.warmup + args.steps), wl["seed"])
            qs = None
        else:
            qs = mapgen.sample_queries(infl, lo, 8 * threads, wl["seed"])
        for it in range(args.warmup + args.steps):
            q = allq[it * per_step:(it + 1) * per_step] if qs is None else qs
            if have_ref:
                sec = impl.plan_batch(0, "f", q, nthreads=threads)["seconds"]
            else:
                t0 = time.perf_counter()
                for one in q:
                    impl.plan(0, "f", one[:2], one[2:])
                sec = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(sec); units += len(q)
        value = units / sum(times)
        metric, unit = "planner_queries_per_s", "queries/s"
        sample = note or (f"{units // max(args.steps, 1)} Astar f queries of the C4 stream per step over {threads} host threads "
                          f"(one planner instance per thread), {args.steps} steps")
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "strong" if args.workload == "c4" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.workload),
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_baseline(workload, infl, lo, gpu_check=None):
    """Bounded sample of the same workload on the host cores (rank 0, N = 1 only).  ``gpu_check`` (per-query arrays or the
    field from the GPU arm) is compared with what the CPU planner returns: the bench line's own parity check."""
    wl = WORKLOADS[workload]
    orc, impl, kind, have_ref = cpu_impl(infl, lo)
    if workload == "c2":
        src = mapgen.field_source(infl, lo)
        reps, secs, settled, r = 4, [], 0, None
        for _ in range(reps):
            t0 = time.perf_counter(); r = impl.field(src); dt = time.perf_counter() - t0
            secs.append(r.get("seconds", dt)); settled = int(r["settled"])
        v = RELAX_PER_SETTLED * settled / float(np.median(secs)) / 1e9
        out = {"value": v, "unit": "Gcell-relax/s", "cores": 1, "kind": kind, "settled": settled,
               "ms_per_field": 1e3 * float(np.median(secs)),
               "sample": f"{reps} full Djikstra::plan fields from the C2 source on one core (median)"}
        if gpu_check is not None:
            g_ref = r["g"]
            reach = g_ref < 1e5
            g_gpu = gpu_check["g"].astype(np.float64)
            same_reach = bool(np.array_equal(g_gpu < 1e5, reach))
            rel = float((np.abs(g_gpu - g_ref)[reach] / np.maximum(g_ref[reach], 1.0)).max())
            out["parity_checked"] = {"reachability_identical": same_reach, "max_rel_cost_error": rel, "bar": 1e-5,
                                     "ok": bool(same_reach and rel <= 1e-5 and gpu_check["reached"] == settled)}
        return out
    if workload == "c3":
        qs = c3_queries(infl, lo, wl["seed"])[:4]
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
    nq = 64 if workload in ("c4", "c4w") else 200
    qs = mapgen.sample_queries(infl, lo, nq, wl["seed"])
    secs, res = [], []
    for q in qs:
        t0 = time.perf_counter(); r = impl.plan(0, "f", q[:2], q[2:]); dt = time.perf_counter() - t0
        secs.append(r.get("seconds", dt)); res.append(r)
    out = {"value": nq / float(np.sum(secs)), "unit": "queries/s", "cores": 1, "kind": kind,
           "p50_ms": 1e3 * float(np.median(secs)), "p90_ms": 1e3 * float(np.percentile(secs, 90)),
           "sample": f"the first {nq} queries of the workload, Astar::plan f one at a time on one core"}
    if gpu_check is not None:
        g = np.array([r["g_goal"] for r in res]); st = np.array([r["settled"] for r in res]); pl = np.array([len(r["path"]) for r in res])
        gg = gpu_check["g_goal"][:nq]
        ok = bool(np.array_equal(gpu_check["settled"][:nq], st) and np.array_equal(gpu_check["path_len"][:nq], pl)
                  and np.all(np.abs(gg - g) <= 1e-12 * np.maximum(np.abs(g), 1.0)) and np.all(gpu_check["status"][:nq] == 0))
        out["parity_checked"] = {"queries": nq, "settled_path_len_identical_cost_within_1e-12": ok, "ok": ok}
    return out


# ---------------------------------------------------------------------------------------------------------------
def run_c5(torch, dist, rank, world, local, dev, steps, warmup, exchange="p2p", band=2048.0, want_e2e=True, want_fallback=True):
    """One field on the 16384^2 map: a single handle at N = 1, row-block shards at N > 1 (peer-to-peer posts over NVLink inside
    the solve kernel, or NCCL rounds).  Returns the measurements (every rank) — the caller formats them."""
    from nsfs import sharded_field as sf
    field_checksum = sf.field_checksum

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
    group = None
    if world > 1 and exchange == "p2p":
        # the C-ABI group entry (NCCL bound inside libnsfs_gpu.so): CUDA IPC handles exchanged over NCCL, neighbours attached once
        from nsfs import capi
        group = capi.Group(dist, rank, world, local)
        group.field_attach(P, shard.lo, shard.r0, shard.r1)

    def step():
        shard.solver.stats.clear()
        if world == 1:
            r = P.field_device((src[0] - a, src[1]))
            return dict(reached=r["reached"], rounds=1, kernel_ms=r["stats"]["main_kernel_ms"], dev_ms=r["stats"]["gpu_ms"],
                        visits=r["stats"]["tile_visits"])
        if exchange == "p2p":
            r = group.field_solve(P, src)              # seed, prime the shared counter, one solve kernel per GPU, finish
            return dict(reached=r["stats"]["expansions"], rounds=1, kernel_ms=r["stats"]["main_kernel_ms"], dev_ms=r["stats"]["gpu_ms"],
                        visits=r["stats"]["tile_visits"], status=r["status"])
        else:
            out = sf.solve_distributed(shard, src, dist, dev, band=band if band > 0 else None)
        return dict(reached=out["result"]["reached"], rounds=out["rounds"],
                    kernel_ms=sum(s["main_kernel_ms"] for s in shard.solver.stats),
                    dev_ms=sum(s["gpu_ms"] for s in shard.solver.stats) + out["result"]["stats"]["gpu_ms"],
                    visits=sum(s["tile_visits"] for s in shard.solver.stats))

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = P.launches()
    t_start = time.time()
    wall, kernel_ms, reached, last, walls = 0.0, 0.0, 0, None, []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        last = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        wall += dt; walls.append(dt)
        kernel_ms += last["kernel_ms"]; reached += last["reached"]
    t_stop = time.time()
    launches = P.launches() - launches0
    # parity: the checksum of the owned rows of g and pred, summed over the ranks, against the committed single-GPU one
    own = part.owned(rank)
    g_rows = shard.owned_rows()
    p_rows = shard.solver.pred_rows(shard.loc(own[0]), own[1] - own[0])
    sg, sp = field_checksum(torch, g_rows, p_rows, own[0] * W)
    if world > 1:
        t = torch.tensor([wall, kernel_ms, float(reached), float(launches)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        wall, kernel_ms, reached, launches = tmax[0].item(), tmax[1].item(), tsum[2].item(), int(tsum[3].item())
        # checksums as four 16-bit limbs each so that the all-reduce sum stays exact in int64
        limbs = torch.tensor([(sg >> s) & 0xffff for s in (0, 16, 32, 48)] + [(sp >> s) & 0xffff for s in (0, 16, 32, 48)],
                             dtype=torch.int64, device=dev)
        dist.all_reduce(limbs, op=dist.ReduceOp.SUM)
        lv = [int(x) for x in limbs.tolist()]
        sg = sum(lv[k] << (16 * k) for k in range(4)) % (1 << 64)
        sp = sum(lv[4 + k] << (16 * k) for k in range(4)) % (1 << 64)
    out = dict(wall=wall, kernel_ms=kernel_ms, reached=reached, launches=launches, last=last, walls=walls, t_start=t_start,
               t_stop=t_stop, checksum_g=f"{sg:016x}", checksum_pred=f"{sp:016x}", steps=steps, H=H, W=W, P=P, shard=shard,
               part=part, infl=infl, lo=lo, src=src, group=group)
    want = None
    try:
        want = json.load(open(C5_CRC_FILE))
    except Exception:
        pass
    per_field = reached / steps
    if want:
        out["parity_checked"] = {"checksum_g": out["checksum_g"], "checksum_pred": out["checksum_pred"],
                                 "committed": {k: want[k] for k in ("checksum_g", "checksum_pred")},
                                 "settled": per_field, "reference_settled": C5_SETTLED,
                                 "ok": bool(out["checksum_g"] == want["checksum_g"] and out["checksum_pred"] == want["checksum_pred"]
                                            and int(per_field) == C5_SETTLED)}
    else:
        out["parity_checked"] = {"checksum_g": out["checksum_g"], "checksum_pred": out["checksum_pred"], "committed": None,
                                 "settled": per_field, "reference_settled": C5_SETTLED, "ok": bool(int(per_field) == C5_SETTLED)}
    if want_e2e:
        # e2e: upload this rank's int32 planes through the host-pointer call, solve, read the owned rows of g and pred back
        pin_i, pin_l = torch.from_numpy(infl).pin_memory(), torch.from_numpy(lo).pin_memory()
        host_g = torch.empty((own[1] - own[0], W), dtype=torch.float32).pin_memory()
        host_p = torch.empty((own[1] - own[0], W), dtype=torch.uint8).pin_memory()

        def step_e2e():
            P.set_map(pin_i.numpy(), pin_l.numpy())
            r = step()
            host_g.copy_(shard.owned_rows()); host_p.copy_(shard.solver.pred_rows(shard.loc(own[0]), own[1] - own[0]))
            torch.cuda.synchronize()
            return r["reached"]

        e2e_steps = max(1, min(steps, 2))
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
        out.update(e2e_sec=e2e_sec, e2e_reached=e2e_reached, e2e_steps=e2e_steps,
                   h2d=int(2 * infl.size * 4 * world), d2h=int(H * W * 5))
    if want_fallback and world == 1:
        # C5 (ii): 1024 find_better_point calls from random non-free interior cells (replicas only: each search is local)
        bad = mapgen.sample_blocked_interior(infl, lo, 1024, wl["seed"])
        P.find_better_point_batch(bad[:64])
        fb = P.find_better_point_batch(bad)
        out["fallback_1024"] = {"queries_per_s": 1024 / (fb["stats"]["gpu_ms"] * 1e-3), "gpu_ms": fb["stats"]["gpu_ms"],
                                "settled": fb["stats"]["expansions"]}
    return out


def c5_summary(r, world, exchange):
    """The C5 measurements as the dict that goes under extra["c5"] (default run) or into the c5 line."""
    alg_bytes = BYTES_PER_SETTLED * (r["reached"] / r["steps"]) / world
    peak, _ = peaks()
    achieved = alg_bytes / (r["kernel_ms"] / r["steps"] * 1e-3) / 1e9
    d = {"workload": WORKLOADS["c5"]["name"], "gcell_relax_per_s": RELAX_PER_SETTLED * r["reached"] / r["wall"] / 1e9,
         "ms_per_field": 1e3 * r["wall"] / r["steps"], "p50_ms_per_field": 1e3 * float(np.median(r["walls"])), "steps": r["steps"],
         "partition": (f"{world} row blocks, 2 ghost rows of g + 4 rows of map per side, " +
                       ("peer-to-peer posts over NVLink inside the solve kernel (no rounds), driven through nsfs_group_field_*" if exchange == "p2p" else "NCCL rounds"))
         if world > 1 else "single handle",
         "timing": "wall clock between cuda synchronize + barrier on both sides, max over ranks",
         "solve_kernel_ms_per_field_max_rank": r["kernel_ms"] / r["steps"], "roofline_frac": achieved / peak,
         "tile_visits_rank0": r["last"]["visits"], "parity_checked": r["parity_checked"]}
    if "fallback_1024" in r:
        d["fallback_1024"] = r["fallback_1024"]
    if "e2e_sec" in r:
        d["e2e_ms_per_field"] = 1e3 * r["e2e_sec"] / r["e2e_steps"]
    return d


def ours_c5(args, torch, dist, rank, world, local, dev):
    sampler = ClockSampler(local) if rank == 0 else None
    r = run_c5(torch, dist, rank, world, local, dev, args.steps, args.warmup, args.exchange, args.band)
    if rank != 0:
        return
    sampler.t0, sampler.t1 = r["t_start"], r["t_stop"]
    peak, peak_src = peaks()
    alg_bytes = BYTES_PER_SETTLED * (r["reached"] / r["steps"]) / world            # per GPU: the dominant kernel of each rank covers its own rows
    achieved = alg_bytes / (r["kernel_ms"] / r["steps"] * 1e-3) / 1e9
    s = c5_summary(r, world, args.exchange)
    line = {
        "metric": "gcell_relax_per_s", "value": s["gcell_relax_per_s"], "unit": "Gcell-relax/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["ms_per_field"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(shared_config("c5"), units_per_step="1 field over all ranks", partition=s["partition"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "field_solve_async_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": r["kernel_ms"] / r["steps"],
                     "note": "per GPU: that rank's share of the settled cells over the summed time of its solve launches (max over ranks)"},
        "e2e": {"value": RELAX_PER_SETTLED * r["e2e_reached"] / r["e2e_sec"] / 1e9, "unit": "Gcell-relax/s",
                "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "ms_per_step": 1e3 * r["e2e_sec"] / r["e2e_steps"], "steps": r["e2e_steps"]},
        "gpu_launches": r["launches"], "clocks": sampler.finish(),
        "extra": {"settled_per_field": r["reached"] / r["steps"], "exchange_rounds": r["last"]["rounds"],
                  "tile_visits_rank0": r["last"]["visits"], "fields_per_s": r["steps"] / r["wall"],
                  "parity_checked": r["parity_checked"], "fallback_1024": r.get("fallback_1024")},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
def extra_c2(torch, local, dev, steps=20, warmup=5, want_cpu=True):
    """Rank 0: the C2 field line (device-timed, L2 flushed), the plugin-path e2e (GpuDjikstra::generatePath, field backend,
    pageable and pinned MapData) and the parity of the field against the reference's own Djikstra::plan."""
    from nsfs import capi, host as nhost

    wl = WORKLOADS["c2"]
    H, W = wl["H"], wl["W"]
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], False)
    src = mapgen.field_source(infl, lo)
    P = capi.Planner(H, W, device=local)
    P.set_map(infl, lo)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        P.field_device(src)
    ms, kms, reached = [], [], 0
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        r = P.field_device(src)
        ms.append(r["stats"]["gpu_ms"]); kms.append(r["stats"]["main_kernel_ms"]); reached = r["reached"]
    peak, _ = peaks()
    out = {"workload": wl["name"], "gcell_relax_per_s": RELAX_PER_SETTLED * reached * steps / (sum(ms) * 1e-3) / 1e9,
           "ms_per_field": float(np.mean(ms)), "p50_ms_per_field": float(np.median(ms)), "steps": steps,
           "solve_kernel_ms": float(np.mean(kms)), "settled": reached,
           "roofline_frac": BYTES_PER_SETTLED * reached / (float(np.mean(kms)) * 1e-3) / 1e9 / peak,
           "timing": "device time of the call's kernels (CUDA events on the handle's stream), L2 flushed between steps"}
    # host-pointer C-ABI (pinned buffers): what round 1 reported as e2e
    pin_i, pin_l = torch.from_numpy(infl).pin_memory(), torch.from_numpy(lo).pin_memory()
    host_g = torch.empty((H, W), dtype=torch.float32).pin_memory(); host_p = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    st = capi.Stats()

    def cabi():
        P.set_map(pin_i.numpy(), pin_l.numpy())
        assert P.L.nsfs_field(P.h, src[0], src[1], host_g.numpy().ctypes.data, host_p.numpy().ctypes.data, st) == 0

    for _ in range(2):
        cabi()
    t0 = time.perf_counter()
    for _ in range(10):
        cabi()
    out["e2e_cabi_pinned_ms"] = 1e3 * (time.perf_counter() - t0) / 10
    field = dict(g=host_g.numpy().copy(), reached=reached)
    P.close()
    # the call a GlobalPlanner makes: GpuDjikstra::generatePath (field backend) on a MapData with pageable std::vector<int> planes
    so, what = host_so()
    goal = tuple(int(x) for x in mapgen.sample_free_cells(infl, lo, 1, 99)[0])
    try:
        S = nhost.Session("djikstra", "g", infl, lo, backend="field", device=local, so_path=so)
        for mode in ("every_call", "every_call_pinned", "on_notify"):
            S.set_map_sync(mode)
            for _ in range(2):
                S.generate_path_idx(src, goal)
            t0 = time.perf_counter()
            for _ in range(10):
                r = S.generate_path_idx(src, goal)
            out[f"e2e_generatePath_{mode}_ms"] = 1e3 * (time.perf_counter() - t0) / 10
        out["e2e_generatePath"] = (f"{what}; one Djikstra plan request = upload of both int32 planes from MapData's std::vector<int> "
                                   "(every_call: pageable, the reference's contract; every_call_pinned: the same vectors page-locked once; "
                                   "on_notify: no upload unless the map changed) + field + predecessor walk + post_process_path")
        S.close()
    except Exception as e:          # noqa: BLE001 — the line must still be printed
        out["e2e_generatePath_error"] = repr(e)
    if want_cpu:
        out["cpu_baseline"] = cpu_baseline("c2", infl, lo, gpu_check=field)
    return out


def ours_c4(args, torch, dist, rank, world, local, dev):
    """Default: the 65 536 C4 queries sharded over the ranks (strong scaling) + the C2 / C5 lines under extra."""
    from nsfs import capi, host as nhost, sharded_queries as sq

    wl = WORKLOADS["c4"]
    H, W = wl["H"], wl["W"]
    total = args.batch
    P = capi.Planner(H, W, device=local)
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    if rank == 0:
        d_infl = torch.from_numpy(infl).to(dev); d_lo = torch.from_numpy(lo).to(dev)
        P.set_map_device(d_infl.data_ptr(), d_lo.data_ptr())
    if world > 1:
        sq.broadcast_packed_map(P, dist, dev, rank)               # N/4 bytes over NCCL, once
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    if rank == 0:
        allq = mapgen.sample_queries(infl, lo, total, wl["seed"])
        allq_t = torch.from_numpy(np.ascontiguousarray(allq)).to(dev)
    else:
        allq_t = torch.zeros((total, 4), dtype=torch.int32, device=dev)
    if world > 1:
        dist.broadcast(allq_t, src=0)
    a, b = sq.query_slice(total, world, rank)
    myq = allq_t[a:b].contiguous()
    nq = b - a
    d_status = torch.zeros(nq, dtype=torch.int32, device=dev); d_g = torch.zeros(nq, dtype=torch.float64, device=dev)
    d_len = torch.zeros(nq, dtype=torch.int32, device=dev); d_set = torch.zeros(nq, dtype=torch.int64, device=dev)

    def step_resident():
        st = P.plan_batch_device("astar", "f", nq, myq.data_ptr(), d_status.data_ptr(), d_g.data_ptr(), d_len.data_ptr(),
                                 d_set.data_ptr(), want_stats=True)
        return st, int(d_set.sum().item())

    sampler = ClockSampler(local) if rank == 0 else None
    for w in range(args.warmup):
        st, _ = step_resident()
        if rank == 0:
            log(f"warm-up {w}: {st['gpu_ms']:.0f} ms for {nq} queries on this rank")
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
    pops_rank = float(last_stats["pops"])
    if world > 1:
        dist.barrier()
        t = torch.tensor([dev_ms, main_ms, float(settled_total), float(launches), pops_rank], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dev_ms_max, main_ms_max, dev_ms_min = tmax[0].item(), tmax[1].item(), tmin[0].item()
        settled_all, launches_all, pops_all = tsum[2].item(), int(tsum[3].item()), tsum[4].item()
    else:
        dev_ms_max, main_ms_max, dev_ms_min = dev_ms, main_ms, dev_ms
        settled_all, launches_all, pops_all = float(settled_total), launches, pops_rank
    gpu_results = dict(status=d_status.cpu().numpy(), g_goal=d_g.cpu().numpy(), path_len=d_len.cpu().numpy(), settled=d_set.cpu().numpy())
    P.close()                                                      # the query slots (85 % of HBM) go back before the plugin-path leg
    del P

    # ---- e2e: the reference-facing C++ child over the nsfsh_* shim.  N = 1: GpuAstar::planBatch.  N > 1: GpuAstar::planBatchSharded
    # over a nsfs_host::DeviceGroup: rank 0 uploads its pageable MapData, the packed map is broadcast (NCCL, called from C++ below
    # the C-ABI), every rank plans its slice and every rank gets all results back ----
    so, what = host_so()
    all_host_q = np.ascontiguousarray(allq_t.cpu().numpy())
    host_q = np.ascontiguousarray(all_host_q[a:b])
    blank = np.zeros_like(infl)
    S = nhost.Session("astar", "f", infl if rank == 0 else blank, lo if rank == 0 else blank, device=local, so_path=so)
    cgroup = nhost.make_group(dist, rank, world, local, so_path=so) if world > 1 else None

    def step_e2e():
        if world > 1:
            r = S.plan_batch_sharded(cgroup, all_host_q)
            return {k: (v[a:b] if isinstance(v, np.ndarray) else v) for k, v in r.items()}
        return S.plan_batch_idx(host_q)                            # uploads both planes from pageable vectors, queries in, results out

    e2e_steps = max(1, min(args.steps, 2))
    step_e2e()                                                     # allocates the session's own query slots
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    e2e_same = bool(np.array_equal(r_e2e["settled"], gpu_results["settled"]) and np.array_equal(r_e2e["g_goal"], gpu_results["g_goal"]))
    if cgroup is not None:
        nhost.destroy_group(cgroup, so_path=so)
    S.close()
    if world > 1:
        t = torch.tensor([e2e_sec, 0.0 if e2e_same else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec, e2e_same = t[0].item(), t[1].item() == 0.0

    # ---- the other two configs BASELINE names, under extra -----------------------------------------------------------------
    extra_lines = {}
    if not args.no_extra:
        if rank == 0:
            try:
                extra_lines["c2"] = extra_c2(torch, local, dev, want_cpu=(world == 1 and not args.no_cpu_baseline))
            except Exception as e:      # noqa: BLE001
                extra_lines["c2"] = {"error": repr(e)}
        if world > 1:
            dist.barrier()
        try:
            r5 = run_c5(torch, dist, rank, world, local, dev, steps=3, warmup=2, exchange="p2p", want_e2e=False)
            extra_lines["c5"] = c5_summary(r5, world, "p2p")
            if r5["group"] is not None:
                r5["P"].field_peer_detach(); r5["group"].close()
            r5["shard"].solver.close()
        except Exception as e:          # noqa: BLE001
            extra_lines["c5"] = {"error": repr(e)}
    if rank != 0:
        return

    peak, peak_src = peaks()
    value = total * args.steps / (dev_ms_max * 1e-3)
    # capi.cu use_team_kernel: a rank's share of up to 2 x SMs x 32 queries (9 472) runs one query per warp (no ncu traffic capture of that one)
    main_kernel = "search_team_kernel" if nq > 2 * torch.cuda.get_device_properties(dev).multi_processor_count * 32 else "search_kernel"
    alg_bytes = BYTES_PER_SETTLED * settled_all / (args.steps * world)          # per launch of the dominant kernel (one per rank and step)
    achieved = alg_bytes / (main_ms_max / args.steps * 1e-3) / 1e9
    line = {
        "metric": "planner_queries_per_s", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config("c4"),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(main_kernel, nq), "kernel": main_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": main_ms_max / args.steps,
                     "note": "per GPU: 9.125 B x the reference-settled cells of that rank's queries over its kernel time (max over ranks)"},
        "e2e": {"value": total * e2e_steps / e2e_sec, "unit": "queries/s",
                "h2d_bytes_per_step": int(2 * H * W * 4 + total * 16), "d2h_bytes_per_step": int(total * 24 * world),
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps,
                "path": f"{what}: " + ("GpuAstar::planBatch" if world == 1 else "GpuAstar::planBatchSharded over nsfs_host::DeviceGroup (map broadcast + "
                        "result all-gather over NCCL, called from C++)") + " on a MapData whose std::vector<int> planes are uploaded on every call (page-locked once by the child: MapSync::EveryCallPinned, its default), "
                        "queries from host memory, per-query results back to the host",
                "l2": "not flushed between e2e steps (the resident line flushes it; at 10 s per step the difference is within the noise)",
                "same_results_as_resident_path": e2e_same},
        "gpu_launches": launches_all,
        "clocks": sampler.finish() if sampler else None,
        "extra": dict({"queries_per_gpu": nq, "map_broadcast": "NCCL, packed bit planes (N/4 bytes), once" if world > 1 else "n/a",
                       "gcell_relax_per_s": RELAX_PER_SETTLED * settled_all / (dev_ms_max * 1e-3) / 1e9,
                       "settled_per_query": settled_all / (total * args.steps), "pops_per_step": pops_all,
                       "p50_ms_per_step": float(np.median(per_step_ms)), "slowest_rank_ms_per_step": dev_ms_max / args.steps,
                       "fastest_rank_ms_per_step": dev_ms_min / args.steps,
                       "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps}, **extra_lines),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline("c4", infl, lo, gpu_check=gpu_results)
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
def ours_replicas(args, torch, dist, rank, world, local, dev):
    """c2 / c1 / c3 / c4w: every rank runs the same unit of work on its own GPU (replicas; no data-path collective)."""
    from nsfs import capi, sharded_queries as sq

    wl = WORKLOADS[args.workload]
    H, W = wl["H"], wl["W"]
    P = capi.Planner(H, W, device=local)
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    if rank == 0:
        d_infl = torch.from_numpy(infl).to(dev)
        d_lo = torch.from_numpy(lo).to(dev)
        P.set_map_device(d_infl.data_ptr(), d_lo.data_ptr())
    if world > 1:
        sq.broadcast_packed_map(P, dist, dev, rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    field = None

    if args.workload == "c2":
        src = mapgen.field_source(infl, lo)                             # the same source on every rank: the same unit of work

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
        per_step = {"c4w": args.batch, "c1": 1, "c3": 16}[args.workload]
        nq = per_step
        rotating = args.workload == "c1"                    # c1: a different query every step (the same sequence on every rank)
        total = nq * ((args.steps + args.warmup) if rotating else 1)
        allq = c3_queries(infl, lo, wl["seed"], 16) if args.workload == "c3" else mapgen.sample_queries(infl, lo, total, wl["seed"])
        myq = torch.from_numpy(np.ascontiguousarray(allq)).to(dev)
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
        main_kernel = "search_team_kernel" if args.workload == "c4w" else "search_kernel"

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
    if args.workload == "c2":
        field = dict(g=host_g.numpy().copy(), reached=int(e2e_settled / e2e_steps))
    if world > 1:
        t = torch.tensor([e2e_sec, float(e2e_settled)], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        e2e_sec, e2e_settled = tmax[0].item(), tsum[1].item()
    if rank != 0:
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
                 "los_checks": last_stats["los_checks"], "p50_ms_per_step": float(np.median(per_step_ms)),
                 "p90_ms_per_step": float(np.percentile(per_step_ms, 90))}
    alg_bytes = BYTES_PER_SETTLED * settled_all / (args.steps * world)          # per launch of the dominant kernel
    achieved = alg_bytes / (main_ms_max / args.steps * 1e-3) / 1e9
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.workload == "c2" else "f64", "data": "synthetic",
        "config": dict(shared_config(args.workload), units_per_step_per_gpu="1 field" if args.workload == "c2" else f"{nq} queries",
                       wall_ms_per_step_incl_flush=1e3 * t_wall / args.steps),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(main_kernel, nq), "kernel": main_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": main_ms_max / args.steps},
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps,
                "path": "host-pointer C-ABI (nsfs_set_map + nsfs_field / nsfs_plan_batch), pinned host buffers"},
        "gpu_launches": launches_all,
        "clocks": sampler.finish() if sampler else None,
        "extra": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, infl, lo, gpu_check=field)
    emit(line)


def ours_latency(args, torch, local, dev):
    """One query at a time on the 2048^2 map (north_star: "single-query latency below the CPU planner on 2048^2 grids"): the
    plugin path (GpuDjikstra / GpuAstar::generatePath through the nsfsh_* shim, MapData uploaded once: on_notify) per planner."""
    from nsfs import capi, host as nhost

    wl = WORKLOADS["lat"]
    H, W = wl["H"], wl["W"]
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], False)
    nq = max(4, min(args.steps, 16))
    qs = mapgen.sample_queries(infl, lo, nq, wl["seed"])
    so, what = host_so()
    table = {}
    for name, planner, mode, backend in (("djikstra_field_backend", "djikstra", "g", "field"), ("djikstra_exact_order", "djikstra", "g", "exact"),
                                         ("astar_f_exact_order", "astar", "f", "exact")):
        S = nhost.Session(planner, mode, infl, lo, backend=backend, device=local, so_path=so)
        S.set_map_sync("on_notify")
        S.generate_path_idx(qs[0][:2], qs[0][2:])
        ms = []
        for q in qs:
            t0 = time.perf_counter(); S.generate_path_idx(q[:2], q[2:]); ms.append(1e3 * (time.perf_counter() - t0))
        table[name] = {"p50_ms": float(np.median(ms)), "max_ms": float(np.max(ms))}
        S.close()
    line = {"metric": "single_query_p50_ms", "value": table["djikstra_field_backend"]["p50_ms"], "unit": "ms", "n_gpus": 1, "steps": nq, "warmup": 1,
            "ms_per_step": table["djikstra_field_backend"]["p50_ms"], "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (Q17.15) / f64", "data": "synthetic", "config": shared_config("lat"),
            "e2e": {"value": table["djikstra_field_backend"]["p50_ms"], "unit": "ms", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 8 * 4096,
                    "path": f"{what}: generatePath, map resident (on_notify), wall clock per call"},
            "gpu_launches": None, "extra": {"gpu": table, "queries": nq,
                                            "note": "value = the Djikstra child with djikstra_backend: field (same reachability and cost as Djikstra::plan, "
                                                    "an equal-cost path); the exact-order replays return the reference's identical path and are a serial chain"}}
    if not args.no_cpu_baseline:
        orc, impl, kind, have_ref = cpu_impl(infl, lo)
        cpu = {}
        for name, algo, mode in (("djikstra_g", 1, "g"), ("astar_f", 0, "f")):
            ms = []
            for q in qs:
                t0 = time.perf_counter(); r = impl.plan(algo, mode, q[:2], q[2:]); dt = time.perf_counter() - t0
                ms.append(1e3 * r.get("seconds", dt))
            cpu[name] = {"p50_ms": float(np.median(ms)), "max_ms": float(np.max(ms))}
        line["cpu_baseline"] = {"value": cpu["djikstra_g"]["p50_ms"], "unit": "ms", "cores": 1, "kind": kind, "table": cpu,
                                "sample": f"the same {nq} queries, plan() one at a time on one core"}
    emit(line)


def ours(args):
    import torch

    rank, world, local = dist_env()
    guard_stdout()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    try:
        if args.workload == "c5":
            ours_c5(args, torch, dist, rank, world, local, dev)
        elif args.workload == "c4":
            ours_c4(args, torch, dist, rank, world, local, dev)
        elif args.workload == "lat":
            if rank == 0:
                ours_latency(args, torch, local, dev)
        else:
            ours_replicas(args, torch, dist, rank, world, local, dev)
    finally:
        if world > 1:
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=C4_TOTAL, help="c4: queries per step over ALL ranks (65 536 = BASELINE); c4w: per GPU")
    ap.add_argument("--band", type=float, default=2048.0, help="c5 at N > 1 with --exchange rounds: cost units all shards advance per round (0 = unbounded)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "rounds"],
                    help="c5 at N > 1: p2p = shards post boundary rows into each other's memory over NVLink from inside the solve "
                         "kernel (one launch per GPU); rounds = NCCL send/recv + all-reduce per band")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="c4: skip the C2 / C5 lines under extra")
    args = ap.parse_args()
    # defaults that finish within minutes: a C4 step is seconds (65 536 queries), c2 3.5 ms, c1 ~0.1 s, c3 / c5 seconds
    if args.steps is None:
        args.steps = {"c2": 30, "c1": 30, "lat": 8}.get(args.workload, 2) if args.impl == "ours" else 2
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
