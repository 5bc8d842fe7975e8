import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
infl, lo = mapgen.synth_map(H, W, 0.10, 5)
P = capi.Planner(H, W); P.set_map(infl, lo)
src = mapgen.field_source(infl, lo)
ref = None
P.set_option('field_mode', int(os.environ.get('MODE', '0')))
inners = [int(x) for x in os.environ.get('INNERS', '1,2,4,8').split(',')]
deltas = [float(x) for x in os.environ.get('DELTAS', '16,32,64,128,1000000').split(',')]
idles = [int(x) for x in os.environ.get('IDLES', '100').split(',')]
alphas = [int(x) for x in os.environ.get('ALPHAS', '100').split(',')]
dins = [int(x) for x in os.environ.get('DINS', '12').split(',')]
groups = [int(x) for x in os.environ.get('TGROUPS', '0').split(',')]
gates = [int(x) for x in os.environ.get('GATES', '0').split(',')]
earlies = [int(x) for x in os.environ.get('EARLY', '0').split(',')]   # 0/1 on, 2 off
merges = [int(x) for x in os.environ.get('MERGE', '0').split(',')]    # 0/1 on, 2 off
fences = [int(x) for x in os.environ.get('FENCE', '0').split(',')]    # 1: release-atomic key posts
P.set_option('field_sched_ns', int(os.environ.get('SCHED', '0')))
waits = [int(x) for x in os.environ.get('WAITS', '200').split(',')]
for inner, delta, alpha, idle, din, wait in itertools.product(inners, deltas, alphas, idles, dins, waits):
  for grp, gate, early, merge, fence in itertools.product(groups, gates, earlies, merges, fences):
    P.set_option('field_groups', grp); P.set_option('field_gate', gate)
    P.set_option('field_early', early); P.set_option('field_merge', merge); P.set_option('field_fence', fence)
    P.set_option('field_delta_in', din); P.set_option('field_wait_ns', wait)
    P.set_option('field_idle_ns', idle)
    P.set_option('field_alpha', alpha)
    P.set_option('field_inner', inner); P.set_option('field_delta', int(delta))
    best = None
    for _ in range(4):
        r = P.field_device(src)
        st = r['stats']
        if best is None or st['gpu_ms'] < best['gpu_ms']: best = st
    g = P.field(src)['g']
    if ref is None: ref = g
    same = np.array_equal(ref, g)
    print(f"early {early} merge {merge} fence {fence} gate {gate} groups {grp} din {din} wait {wait} idle {idle} alpha {alpha} inner {inner} delta {delta:9.0f}: gpu_ms {best['gpu_ms']:.3f} solve_ms {best['main_kernel_ms']:.3f} rounds {best['rounds']} visits {best['tile_visits']} sweeps {best['pops']} reached {r['reached']} same_bits {same}", flush=True)
