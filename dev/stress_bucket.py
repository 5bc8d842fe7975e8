"""Repeat the bucket solver (field_mode 3) under several options and count cells that differ from the pull solver."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
H, W = 700, 900
infl, lo = mapgen.synth_map(H, W, 0.12, 93)
src = mapgen.field_source(infl, lo)
P = capi.Planner(H, W); P.set_map(infl, lo)
P.set_option('field_mode', 0)
want = P.field(src)
P.set_option('field_mode', 3)
for early in (0, 2):
    for delta in (0, 6, 40, 100):
        for groups in (0, 1, 2):
            P.set_option('field_early', early); P.set_option('field_delta', delta); P.set_option('field_groups', groups)
            bad = []
            for rep in range(6):
                got = P.field(src)
                nb = int((got['g'] != want['g']).sum())
                if nb:
                    d = np.argwhere(got['g'] != want['g'])
                    bad.append((nb, tuple(d[0]), float(got['g'][tuple(d[0])]), float(want['g'][tuple(d[0])]), int((got['g'] > want['g']).sum())))
            print(f"fence {early == 2} delta {delta} blocks/SM {groups}: mismatching runs {len(bad)}/6 {bad[:2]}", flush=True)
