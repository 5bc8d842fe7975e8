#!/usr/bin/env python
"""bench.py — headline measurement of the B200 grid planner (one JSON line on stdout from rank 0).

Workloads (SURVEY.md §8d / BASELINE.json configs; ``--workload``):
  c2 (default)  Djikstra full cost-to-go field, 2048x2048 raw map p=0.10 seed 5, source = centre      [configs[1]]
  c4            batched Astar "f" start/goal queries, 1024x1024 raw map p=0.10 seed 7                  [configs[3]]
  c1            Astar "f" single queries, 384x384 inflated map seed 1 (latency)                        [configs[0]]

A step = one pass of the hot path over one batch of synthetic input: one field (c2), one batch of ``--batch``
queries (c4), one query (c1).  ``value`` is measured with inputs resident in HBM (device time of the step's
kernels, CUDA events on the launching stream, L2 flushed between steps); ``e2e`` goes through the host-pointer
C-ABI calls with host<->device copies inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c4|c1] [--impl ours|reference]
N > 1 is launched by torchrun (one rank per GPU, NCCL): the packed map is broadcast once from rank 0, the work
units are independent per rank (no data-path collective), timing is the max over ranks.
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
C2_SETTLED = 3767918             # reference Djikstra settled cells on the C2 input (asserted in tests/)

WORKLOADS = {
    "c2": dict(H=2048, W=2048, p=0.10, seed=5, inflated=False,
               name="Djikstra full cost-to-go field, 2048x2048 raw map p_occ=0.10 seed 5, source = centre (C2)"),
    "c4": dict(H=1024, W=1024, p=0.10, seed=7, inflated=False,
               name="batched Astar f start/goal queries, 1024x1024 raw map p_occ=0.10 seed 7 (C4)"),
    "c1": dict(H=384, W=384, p=0.005, seed=1, inflated=True,
               name="Astar f single query, 384x384 inflated map p_occ=0.005 seed 1 (C1)"),
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


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if there is one."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
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


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU planners (oracle/_ref when present, else the C restatement) on all host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import oracle as orc        # the one other place bench.py may execute oracle/ (spec ④)
    wl = WORKLOADS[args.workload]
    infl, lo = mapgen.synth_map(wl["H"], wl["W"], wl["p"], wl["seed"], wl["inflated"])
    cores = os.cpu_count() or 1
    have_ref = orc.have_reference()
    impl = orc.Reference(infl, lo) if have_ref else orc.Oracle(infl, lo)
    kind = "reference" if have_ref else "port"
    times, units = [], 0
    if args.workload == "c2":
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
        sample = f"{threads} full fields per step (one per host thread) on the C2 map, {args.steps} steps"
    else:
        nq = max(cores, 8) if args.workload == "c4" else 8
        threads = cores if have_ref else 1
        qs = mapgen.sample_queries(infl, lo, nq, wl["seed"])
        for it in range(args.warmup + args.steps):
            if have_ref:
                r = impl.plan_batch(0, "f", qs, nthreads=threads)
                sec = r["seconds"]
            else:
                t0 = time.perf_counter()
                for q in qs:
                    impl.plan(0, "f", q[:2], q[2:])
                sec = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(sec); units += nq
        value = units / sum(times)
        metric, unit = "planner_queries_per_s", "queries/s"
        sample = f"{nq} Astar f queries per step over {threads} host threads, {args.steps} steps"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "threads": threads},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(workload, infl, lo):
    """Bounded sample of the same workload on the host cores (rank 0, N = 1 only)."""
    from oracle import oracle as orc        # cpu_baseline leg: the checker timed beside the GPU, never the product path
    have_ref = orc.have_reference()
    impl = orc.Reference(infl, lo) if have_ref else orc.Oracle(infl, lo)
    kind = "reference" if have_ref else "port"
    if workload == "c2":
        src = mapgen.field_source(infl, lo)
        reps, secs, settled = 6, [], 0
        for _ in range(reps):
            t0 = time.perf_counter(); r = impl.field(src); dt = time.perf_counter() - t0
            secs.append(r.get("seconds", dt)); settled = int(r["settled"])
        v = RELAX_PER_SETTLED * settled / float(np.median(secs)) / 1e9
        return {"value": v, "unit": "Gcell-relax/s", "cores": 1, "kind": kind, "settled": settled,
                "ms_per_field": 1e3 * float(np.median(secs)),
                "sample": f"{reps} full Djikstra::plan fields from the C2 source on one core (median)"}
    wl = WORKLOADS[workload]
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
def ours(args):
    import torch
    from nsfs import capi

    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    wl = WORKLOADS[args.workload]
    H, W = wl["H"], wl["W"]
    P = capi.Planner(H, W, device=local)
    words = P.map_words()

    # ---- inputs: rank 0 builds the synthetic map; other ranks receive the packed planes over NCCL -------------
    # (every rank also keeps the int32 host planes: the e2e leg uploads them through nsfs_set_map each step)
    infl, lo = mapgen.synth_map(H, W, wl["p"], wl["seed"], wl["inflated"])
    if rank == 0:
        d_infl = torch.from_numpy(infl).to(dev)
        d_lo = torch.from_numpy(lo).to(dev)
        P.set_map_device(d_infl.data_ptr(), d_lo.data_ptr())
    if world > 1:
        planes = torch.empty(2 * words, dtype=torch.int32, device=dev)
        if rank == 0:
            f_ptr, i_ptr = P.packed_map_device()
            src_planes = torch.empty(2 * words, dtype=torch.int32, device=dev)
            _copy_from_ptr(src_planes[:words], f_ptr, words * 4)
            _copy_from_ptr(src_planes[words:], i_ptr, words * 4)
            planes.copy_(src_planes)
        dist.broadcast(planes, src=0)
        if rank != 0:
            P.set_packed_map_device(planes.data_ptr(), planes[words:].data_ptr())
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
        units_per_step = None      # settled cells, read back from the kernel (== reference count, asserted in tests/)

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

        h2d, d2h = 2 * H * W * 4 + 8, H * W * 4 + H * W
        main_kernel = "field_solve_async_kernel"
    else:
        nq = args.batch if args.workload == "c4" else 1
        if rank == 0:
            allq = mapgen.sample_queries(infl, lo, nq * world * (1 if args.workload == "c4" else args.steps + args.warmup), wl["seed"])
            allq_t = torch.from_numpy(allq).to(dev)
        else:
            allq_t = torch.zeros((nq * world * (1 if args.workload == "c4" else args.steps + args.warmup), 4), dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(allq_t, src=0)
        per = allq_t.shape[0] // world
        myq = allq_t[rank * per:(rank + 1) * per].contiguous()
        d_status = torch.zeros(nq, dtype=torch.int32, device=dev); d_g = torch.zeros(nq, dtype=torch.float64, device=dev)
        d_len = torch.zeros(nq, dtype=torch.int32, device=dev); d_set = torch.zeros(nq, dtype=torch.int64, device=dev)
        counter = {"i": 0}

        def step_resident():
            q = myq if args.workload == "c4" else myq[counter["i"] % per: counter["i"] % per + 1]
            counter["i"] += 1
            st = P.plan_batch_device("astar", "f", nq, q.data_ptr(), d_status.data_ptr(), d_g.data_ptr(), d_len.data_ptr(),
                                     d_set.data_ptr(), want_stats=True)
            return st, int(d_set.sum().item())

        host_q = myq.cpu().numpy()

        def step_e2e():
            q = host_q if args.workload == "c4" else host_q[counter["i"] % per: counter["i"] % per + 1]
            counter["i"] += 1
            r = P.plan_batch("astar", "f", q, path_cap=0)
            return int(r["settled"].sum())

        h2d, d2h = nq * 16, nq * (4 + 8 + 4 + 8)
        main_kernel = "search_kernel"

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
    dev_ms, main_ms, settled_total, last_stats = 0.0, 0.0, 0, None
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        st, settled = step_resident()
        dev_ms += st["gpu_ms"]; main_ms += st["main_kernel_ms"]; settled_total += settled; last_stats = st
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
    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    e2e_steps = max(3, min(args.steps, 20))
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
                 "rounds": last_stats["rounds"], "tile_visits": last_stats["tile_visits"], "p50_ms_single_query": ms_per_step}
    else:
        metric, unit = "planner_queries_per_s", "queries/s"
        value = world * nq * args.steps / (dev_ms_max * 1e-3)
        e2e_value = world * nq * e2e_steps / e2e_sec
        extra = {"gcell_relax_per_s": RELAX_PER_SETTLED * settled_all / (dev_ms_max * 1e-3) / 1e9,
                 "settled_per_query": settled_all / (world * nq * args.steps), "batch": nq, "pops": last_stats["pops"]}
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
                     "traffic": ncu_traffic(main_kernel), "kernel": main_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_per_launch": main_ms_max / args.steps},
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps},
        "gpu_launches": launches_all,
        "clocks": sampler.finish() if sampler else None,
        "extra": extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, infl, lo)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _copy_from_ptr(dst_tensor, src_ptr, nbytes):
    """Device-to-device copy from a raw device pointer into a torch tensor (cudaMemcpy via the CUDA runtime)."""
    import torch

    # torch exposes no raw-pointer copy; wrap the pointer with the __cuda_array_interface__ protocol instead
    class _Raw:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n // 4,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    src = torch.as_tensor(_Raw(src_ptr, nbytes), device=dst_tensor.device)
    dst_tensor.copy_(src)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="queries per step per GPU (c4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
