import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'nav-stack-from-scratch_b200'))
import numpy as np, torch
from nsfs import capi
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.RandomState(1)
lo = torch.from_numpy(np.where(rng.rand(H, W) < 0.1, 20, 0).astype(np.int32)).pin_memory()
lo.numpy()[-12:, -12:] = 0
P = capi.Planner(H, W)
for r in range(3):
    t0 = time.perf_counter(); rc = P.set_map_logodds(lo.numpy(), 10, 10); dt = time.perf_counter() - t0
    print(f"set_map_logodds {H}x{W}: rc {rc} {dt*1e3:.2f} ms wall (H2D {H*W*4/1e9:.2f} GB + inflate + pack + masks)", flush=True)
