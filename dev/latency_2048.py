"""Single-query latency on a 2048 x 2048 map (north_star: "single-query latency below the CPU planner on 2048^2 grids"):
the three ways a GridPlannerCore child can serve one query on the GPU, next to the reference planners on one host core."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
from oracle import oracle as orc
H = W = 2048
infl, lo = mapgen.synth_map(H, W, 0.10, 5)
qs = mapgen.sample_queries(infl, lo, 8, 5)
P = capi.Planner(H, W); P.set_map(infl, lo)
R = orc.Reference(infl, lo) if orc.have_reference() else orc.Oracle(infl, lo)
out = {"map": "2048x2048 raw p_occ=0.10 seed 5", "queries": 8, "cpu_kind": R.kind if hasattr(R, "kind") else "reference"}
def gpu(algo, mode):
    ms = []
    for q in qs:
        P.plan(algo, mode, q[:2], q[2:], want_path=False)            # warm
        t0 = time.perf_counter(); r = P.plan(algo, mode, q[:2], q[2:], cap=8192); ms.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ms)), float(np.max(ms))
def cpu(algo, mode):
    ms = []
    for q in qs:
        t0 = time.perf_counter(); r = R.plan(algo, mode, q[:2], q[2:]); dt = time.perf_counter() - t0
        ms.append(1e3 * r.get("seconds", dt))
    return float(np.median(ms)), float(np.max(ms))
out["gpu_astar_f_exact_ms_p50_max"] = gpu("astar", "f")
out["gpu_djikstra_g_exact_ms_p50_max"] = gpu("djikstra", "g")
out["gpu_djikstra_field_backend_ms_p50_max"] = gpu("djikstra_field", "g")
out["cpu_astar_f_ms_p50_max"] = cpu(0, "f")
out["cpu_djikstra_g_ms_p50_max"] = cpu(1, "g")
print(json.dumps(out))
