"""Dev: one batch of Astar f queries (python dev/prof_search.py [H] [nq] [reps] [search_kernel] [max_slots_warps]).
search_kernel: 0 auto, 1 one query per warp (search.cu), 2 four queries per warp (search_team.cu)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
kern = int(sys.argv[4]) if len(sys.argv) > 4 else 0
slots = int(sys.argv[5]) if len(sys.argv) > 5 else 0
seed, p, infl_mode = (7, 0.10, False) if H != 384 else (1, 0.005, True)
infl, lo = mapgen.synth_map(H, H, p, seed, infl_mode)
P = capi.Planner(H, H); P.set_map(infl, lo)
P.set_option("search_kernel", kern)
if slots:
    P.set_option("max_slots", slots)
qs = mapgen.sample_queries(infl, lo, nq, seed)
for r in range(reps):
    b = P.plan_batch("astar", "f", qs, path_cap=0)
    st = b["stats"]
    print(f"rep {r}: kernel {kern} nq {nq} gpu_ms {st['gpu_ms']:.1f} q/s {nq / st['gpu_ms'] * 1e3:.1f} pops {st['pops']} max settled {b['settled'].max()} "
          f"mean settled {b['settled'].mean():.0f} heap_peak {st['heap_peak']} statuses {np.unique(b['status']).tolist()}", flush=True)
