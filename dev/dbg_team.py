import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
H, W, p, seed = 320, 288, 0.12, 61
infl, lo = mapgen.synth_map(H, W, p, seed, False)
qs = mapgen.sample_queries(infl, lo, 96, seed)
P = capi.Planner(H, W); P.set_map(infl, lo)
P.set_option("search_kernel", 1)
wide = P.plan_batch("astar", "f", qs, path_cap=2048)
P.set_option("search_kernel", 2)
P.set_option("team_heap_cap", 64)
small = P.plan_batch("astar", "f", qs, path_cap=2048)
bad = [q for q in range(96) if not np.array_equal(small["paths"][q], wide["paths"][q])]
print("differ:", bad)
for q in bad[:4]:
    a, b = small["paths"][q], wide["paths"][q]
    k = np.argmax((a != b).any(1))
    print(q, "len", small["path_len"][q], wide["path_len"][q], "first diff at", k, a[k-1:k+2].tolist(), b[k-1:k+2].tolist(), "g", small["g_goal"][q], wide["g_goal"][q], "settled", small["settled"][q], wide["settled"][q])
