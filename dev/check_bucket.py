"""A/B of the field solvers: field_mode 0 (pull relaxation) against field_mode 3 (bucketed push), same bits expected."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi

cases = [
    (64, 64, 0.10, 1, True), (100, 130, 0.10, 2, True), (65, 63, 0.2, 3, False), (7, 300, 0.05, 4, False), (300, 5, 0.05, 5, False),
    (384, 384, 0.10, 6, True), (384, 384, 0.30, 7, False), (1024, 1024, 0.10, 7, True), (2048, 2048, 0.10, 5, True),
]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1:]]
deltas = [int(x) for x in os.environ.get('DELTAS', '0').split(',')]
groups = [int(x) for x in os.environ.get('BPS', '0').split(',')]
ok_all = True
for (H, W, p, seed, frame) in cases:
    infl, lo = mapgen.synth_map(H, W, p, seed, frame=frame)
    P = capi.Planner(H, W); P.set_map(infl, lo)
    src = mapgen.field_source(infl, lo)
    P.set_option('field_mode', 0)
    try:
        r0 = P.field(src)
    except capi.NsfsError as e:
        print(f"{H}x{W}: mode 0 raised {e}"); r0 = None
    t0 = None
    if r0 is not None:
        t0 = min(P.field_device(src)['stats']['main_kernel_ms'] for _ in range(3))
    P.set_option('field_mode', 3)
    for d in deltas:
        for bps in groups:
            P.set_option('field_delta', d); P.set_option('field_groups', bps)
            try:
                r3 = P.field(src)
            except capi.NsfsError as e:
                print(f"{H}x{W}: mode 3 raised {e}")
                if r0 is not None: ok_all = False
                continue
            best = None
            for _ in range(4):
                st = P.field_device(src)['stats']
                if best is None or st['main_kernel_ms'] < best['main_kernel_ms']: best = st
            same = r0 is not None and np.array_equal(r0['g'], r3['g']) and np.array_equal(r0['pred'], r3['pred']) and r0['reached'] == r3['reached']
            ok_all &= bool(same)
            nbad = int((r0['g'] != r3['g']).sum()) if r0 is not None else -1
            print(f"{H}x{W} p={p} frame={frame} delta={d} bps={bps}: mode0 {t0:.3f} ms | mode3 solve {best['main_kernel_ms']:.3f} ms gpu {best['gpu_ms']:.3f} visits {best['tile_visits']} "
                  f"steps {best['pops']} merges {best['pushes']} empty {best['heap_peak']} reached {r3['reached']} same_bits {same} diff_cells {nbad}", flush=True)
    P.close()
print("ALL_OK" if ok_all else "MISMATCH")
