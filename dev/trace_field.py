"""Wave-speed trace of the field solve kernel (instrumented build: python dev/build_variant.py trace -DNSFS_FIELD_TRACE=1).
Writes gpurun_out/trace_<size>.bin (raw) and prints a summary; analyse with dev/analyse_trace.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NSFS_GPU_SO"] = os.path.join(ROOT, "dev", "_build", "libnsfs_gpu_trace.so")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "nav-stack-from-scratch_b200"))
import numpy as np
from nsfs import mapgen, capi
size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
seed, p = (5, 0.10) if size != 16384 else (11, 0.10)
infl, lo = mapgen.synth_map_rows(size, size, p, seed, 0, size)
P = capi.Planner(size, size); P.set_map(infl, lo)
fm = mapgen.free_mask(infl[size // 2:size // 2 + 1], lo[size // 2:size // 2 + 1])[0]
sj = size // 2
while not fm[sj]: sj += 1
src = (size // 2, sj)
for k, v in [x.split("=") for x in sys.argv[2:]]:
    P.set_option(k, int(v))
for _ in range(3):
    r = P.field_device(src)
out = os.path.join(ROOT, "gpurun_out", f"trace_{size}.bin")
os.makedirs(os.path.dirname(out), exist_ok=True)
os.environ["NSFS_FIELD_TRACE_OUT"] = out
r = P.field_device(src)
del os.environ["NSFS_FIELD_TRACE_OUT"]
print("traced run:", r["stats"]["main_kernel_ms"], "ms solve,", r["stats"]["tile_visits"], "visits, reached", r["reached"])
r = P.field_device(src)
print("untraced run of the same build:", r["stats"]["main_kernel_ms"], "ms")
# tile min g for the analysis: the cost distance of every tile
import torch
from nsfs import sharded_field as sf
g = sf._tensor_from_ptr(torch, r["d_g"], (size, size), "<f4", torch.device("cuda", 0))
T = size // 64
gm = g.reshape(T, 64, T, 64).permute(0, 2, 1, 3).reshape(T, T, -1)
gmin = gm.min(-1).values.cpu().numpy()
gmax = torch.where(gm < 1e5, gm, torch.zeros_like(gm)).max(-1).values.cpu().numpy()
np.savez(os.path.join(ROOT, "gpurun_out", f"trace_{size}_tiles.npz"), gmin=gmin, gmax=gmax, src=np.array(src))
