import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np
from nsfs import mapgen, capi
from oracle import oracle as orc
infl, lo = mapgen.synth_map(400, 360, 0.08, 91, inflated=True)
infl[150:190, 100:260] = 3; lo[150:190, 100:260] = 20
infl[250:270, 40:80] = 2
bad = mapgen.sample_blocked_interior(infl, lo, 600, 91, margin=3)
bad = np.concatenate([bad, [[170, 180], [169, 181], [160, 120], [260, 60], [255, 45]]]).astype(np.int32)
P = capi.Planner(400, 360); P.set_map(infl, lo)
small = P.find_better_point_batch(bad)
nz = np.nonzero(small["status"])[0]
print("nonzero:", nz, small["status"][nz], bad[nz], small["cells"][nz], small["stats"])
O = orc.Oracle(infl, lo)
for t in nz[:5]:
    print(t, O.find_better_point(bad[t]))
