"""Sweep the field solver's scheduling knobs on one map (same bits expected everywhere).
    python dev/tune_field.py [size]     env: MODES=0,1,2,3,4 TILES=0 DELTAS=128 INNERS=16 IDLES=1000 WAITS=200 SCHED=0 MERGE=0 FENCE=0 GATES=0"""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
seed = 11 if H == 16384 else 5
infl, lo = mapgen.synth_map_rows(H, W, 0.10, seed, 0, H)
P = capi.Planner(H, W); P.set_map(infl, lo)
fm = mapgen.free_mask(infl[H // 2:H // 2 + 1], lo[H // 2:H // 2 + 1])[0]
sj = W // 2
while not fm[sj]: sj += 1
src = (H // 2, sj)
ev = lambda k, d: [int(x) for x in os.environ.get(k, d).split(',')]
ref = None
for mode, tiles, delta, inner, idle, wait, merge, fence, gate in itertools.product(ev('MODES', '0,1,2,3,4'), ev('TILES', '0'), ev('DELTAS', '128'), ev('INNERS', '16'),
                                                                           ev('IDLES', '1000'), ev('WAITS', '200'), ev('MERGE', '0'), ev('FENCE', '0'), ev('GATES', '0')):
    for k, v in (('field_mode', mode), ('field_tiles_per_sm', tiles), ('field_delta', delta), ('field_inner', inner), ('field_idle_ns', idle),
                 ('field_wait_ns', wait), ('field_merge', merge), ('field_fence', fence), ('field_gate', gate), ('field_sched_ns', int(os.environ.get('SCHED', '0')))):
        P.set_option(k, v)
    best = None
    for _ in range(4):
        r = P.field_device(src)
        st = r['stats']
        if best is None or st['main_kernel_ms'] < best['main_kernel_ms']: best = st
    import torch
    from nsfs import sharded_field as sf
    g = sf._tensor_from_ptr(torch, r['d_g'], (H, W), '<f4', torch.device('cuda', 0))
    chk = sf.field_checksum(torch, g, sf._tensor_from_ptr(torch, r['d_pred'], (H, W), '|u1', torch.device('cuda', 0)), 0)
    if ref is None: ref = chk
    print(f"mode {mode} tiles/sm {tiles} delta {delta} inner {inner} idle {idle} wait {wait} merge {merge} fence {fence} gate {gate}: solve_ms {best['main_kernel_ms']:.3f} "
          f"gpu_ms {best['gpu_ms']:.3f} visits {best['tile_visits']} reached {r['reached']} same_bits {chk == ref}", flush=True)
