import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np, torch
from nsfs import capi, mapgen
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
infl, lo = mapgen.synth_map_rows(H, W, 0.10, 11, 0, H)
d_i, d_l = torch.from_numpy(infl).cuda(), torch.from_numpy(lo).cuda()
P = capi.Planner(H, W)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); P.set_map_device(d_i.data_ptr(), d_l.data_ptr()); dt = time.perf_counter() - t0
    print(f"set_map_device {H}x{W}: {dt*1e3:.2f} ms", flush=True)
segs = np.random.RandomState(3).randint(1, H - 1, (1 << 20, 4)).astype(np.int32)
for r in range(2):
    t0 = time.perf_counter(); out = P.los_batch(segs); dt = time.perf_counter() - t0
    print(f"los_batch {len(segs)} rays: {dt*1e3:.2f} ms wall incl. copies, {100*(out==1).mean():.2f}% clear", flush=True)
