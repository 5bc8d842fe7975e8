"""Generate tests/golden/golden_v1.npz by RUNNING THE REFERENCE ITSELF.

The reference repository ships no tests, fixtures or known-answer vectors, so the golden set is
produced here: oracle/_ref/libnsfs_ref.so is the reference's own planner sources, compiled unmodified from
/root/reference by oracle/Makefile, and every array below is an output of that library.  /root/reference is not
available on the GPU box, which is why the vectors are committed.

    python tests/golden/make_golden.py          # needs /root/reference (this container)

Case list (small on purpose, the whole file is a few hundred kB):
  m0  48x48  raw       p=0.15  framed     m1  64x80 inflated p=0.01 framed
  m2  40x56  raw       p=0.25  framed     m3  24x31 raw      p=0.08 NO frame (flat-key wrap + std::out_of_range)
per map: 16 start/goal pairs x {astar f, astar fg, astar g, djikstra g, djikstra f}: status, raw path, g(goal),
settled count, post-processed path; 2 full Djikstra fields; 24 find_better_point calls; 64 Bresenham rays and
isLos results; plus OpenList pop traces for the three comparators and dist_oct / dist_euc samples.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nav-stack-from-scratch_b200"))

from nsfs import mapgen  # noqa: E402
from oracle.oracle import ASTAR, DJIKSTRA, Reference  # noqa: E402

MAPS = {
    "m0": dict(H=48, W=48, p=0.15, seed=11, inflated=False, frame=True),
    "m1": dict(H=64, W=80, p=0.01, seed=12, inflated=True, frame=True),
    "m2": dict(H=40, W=56, p=0.25, seed=13, inflated=False, frame=True),
    "m3": dict(H=24, W=31, p=0.08, seed=14, inflated=False, frame=False),
}
COMBOS = [("astar", ASTAR, "f"), ("astar", ASTAR, "fg"), ("astar", ASTAR, "g"), ("djikstra", DJIKSTRA, "g"),
          ("djikstra", DJIKSTRA, "f")]
NQ, NFB, NSEG = 16, 24, 64


def make_map(spec):
    return mapgen.synth_map(spec["H"], spec["W"], spec["p"], spec["seed"], spec["inflated"], spec["frame"])


def queries_for(name, spec, infl, lo):
    qs = mapgen.sample_queries(infl, lo, NQ, spec["seed"])
    if name == "m3":
        # unframed map: force cases around the four cells whose expansion throws
        H, W = spec["H"], spec["W"]
        free = mapgen.free_mask(infl, lo)
        corner = [(0, 0), (1, 0), (H - 1, W - 1), (H - 2, W - 1), (0, W - 1), (H - 1, 0)]
        k = 0
        for c in corner:
            if free[c]:
                qs[k, :2] = c
                k += 1
        qs[k, 2:] = (0, 0)
    # one unreachable-goal case (a blocked goal cell) and one start == goal
    blocked = mapgen.sample_blocked_interior(infl, lo, 1, spec["seed"] + 5, margin=1)[0]
    qs[-1, 2:] = blocked
    qs[-2, 2:] = qs[-2, :2]
    return qs


def main():
    out = {}
    rng = np.random.RandomState(2024)
    for name, spec in MAPS.items():
        infl, lo = make_map(spec)
        R = Reference(infl, lo)
        H, W = spec["H"], spec["W"]
        out[f"{name}/infl"] = infl.astype(np.int16)
        out[f"{name}/lo"] = lo.astype(np.int16)
        qs = queries_for(name, spec, infl, lo)
        out[f"{name}/queries"] = qs
        for aname, algo, mode in COMBOS:
            key = f"{name}/{aname}_{mode}"
            status, g, settled, plen, flen = [], [], [], [], []
            paths, finals = [], []
            for q in qs:
                r = R.generate_path(algo, mode, q[:2], q[2:])
                status.append(r["status"])
                ok = r["status"] == 0
                g.append(r["g_goal"] if ok else -1.0)
                settled.append(r["settled"] if ok else -1)
                p = r["path"] if ok else np.zeros((0, 2), np.int32)
                f = r["final_xy"] if ok else np.zeros((0, 2), np.float64)
                plen.append(len(p)); flen.append(len(f))
                paths.append(p); finals.append(f)
            out[key + "/status"] = np.array(status, np.int32)
            out[key + "/g_goal"] = np.array(g, np.float64)
            out[key + "/settled"] = np.array(settled, np.int64)
            out[key + "/path_len"] = np.array(plen, np.int32)
            out[key + "/final_len"] = np.array(flen, np.int32)
            out[key + "/paths"] = np.concatenate(paths).astype(np.int32)
            out[key + "/finals"] = np.concatenate(finals).astype(np.float64)
        # fields
        srcs = mapgen.sample_free_cells(infl, lo, 2, spec["seed"] + 7)
        out[f"{name}/field_src"] = srcs
        for t, s in enumerate(srcs):
            f = R.field(s)
            out[f"{name}/field{t}/status"] = np.array([f["status"]], np.int32)
            if f["status"] == 0:
                out[f"{name}/field{t}/g"] = f["g"]
                out[f"{name}/field{t}/visited"] = f["visited"]
        # fallback
        bads = mapgen.sample_blocked_interior(infl, lo, NFB, spec["seed"] + 9, margin=0 if name == "m3" else 2)
        out[f"{name}/fb_in"] = bads
        fb = [R.find_better_point(b) for b in bads]
        out[f"{name}/fb_status"] = np.array([r["status"] for r in fb], np.int32)
        out[f"{name}/fb_out"] = np.array([r["cell"] if r["status"] == 0 else (-1, -1) for r in fb], np.int32)
        # rays (end points may lie outside the grid: testPos then reports blocked or throws)
        segs = np.stack([rng.randint(-2, H + 2, NSEG), rng.randint(-2, W + 2, NSEG),
                         rng.randint(-2, H + 2, NSEG), rng.randint(-2, W + 2, NSEG)], 1).astype(np.int32)
        segs[: NSEG // 2] = np.clip(segs[: NSEG // 2], 1, [H - 2, W - 2, H - 2, W - 2])
        segs[0] = (5, 5, 5, 5)
        out[f"{name}/segs"] = segs
        rays = [R.bresenham(s[:2], s[2:]) for s in segs]
        out[f"{name}/ray_len"] = np.array([len(r) for r in rays], np.int32)
        out[f"{name}/rays"] = np.concatenate(rays).astype(np.int32)
        out[f"{name}/los"] = np.array([R.is_los(s[:2], s[2:]) for s in segs], np.int32)
        out[f"{name}/dist_oct"] = np.array([R.dist_oct(s[:2], s[2:]) for s in segs], np.float64)
        out[f"{name}/dist_euc"] = np.array([R.dist_euc(s[:2], s[2:]) for s in segs], np.float64)
        del R
    # OpenList traces: 3000 operations, ~60 % pushes, small key range so ties dominate (as in the planners)
    infl, lo = make_map(MAPS["m0"])
    R = Reference(infl, lo)
    for mode in ("f", "g", "fg"):
        n = 3000
        ops = np.zeros((n, 4), np.int32)
        ops[:, 0] = rng.rand(n) < 0.6
        ops[:, 1] = rng.randint(0, 12, n)
        ops[:, 2] = rng.randint(0, 8, n)
        ops[:, 3] = np.arange(n)
        out[f"heap/{mode}/ops"] = ops
        out[f"heap/{mode}/pops"] = R.heap_trace(mode, ops)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
