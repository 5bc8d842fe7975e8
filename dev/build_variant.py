"""Build an instrumented / experimental variant of libnsfs_gpu.so next to the product library (dev/_build/<name>.so).
    python dev/build_variant.py trace -DNSFS_FIELD_TRACE=1
Select it with NSFS_GPU_SO=dev/_build/libnsfs_gpu_<name>.so (nsfs/capi.py reads the variable)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nav-stack-from-scratch_b200")
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "dev", "_build", f"libnsfs_gpu_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
srcs = sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")))
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
       *flags, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"), "-o", out, *srcs]
subprocess.run(cmd, check=True)
print(out)
