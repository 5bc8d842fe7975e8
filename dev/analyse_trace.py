"""Read a trace written by dev/trace_field.py (NSFS_FIELD_TRACE build) and print where a field solve's time goes:
visit phases, claim latency (first key post -> the owner block starts the visit), block occupancy, and the wave's progress
(cost distance reached vs time).   python dev/analyse_trace.py gpurun_out/trace_2048.bin [tiles.npz]"""
import sys
import numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.uint64)
tx, ty, nrec, grid = (int(x) for x in raw[:4])
body = raw[4:]
nt = tx * ty
off = (16 + nt + 7) & ~7
first_post = body[16:16 + nt].astype(np.int64)
cyc = body[1:8].astype(np.float64)
if cyc.sum() > 0 and cyc[0] < 1e18:
    names = ["due scan + idle", "claim + set-up", "passes", "in-tile posts", "write-back", "fence + tile posts"]
    tot = cyc[:6].sum()
    print("warp-cycles inside the relax phase, all visits: " + ", ".join(f"{n} {100*c/tot:.1f}%" for n, c in zip(names, cyc[:6])) +
          f" | sub-block turns {cyc[6]:.0f}, cycles per turn excluding idle {(tot-cyc[0])/max(cyc[6],1):.0f} (set-up {cyc[1]/max(cyc[6],1):.0f}, passes {cyc[2]/max(cyc[6],1):.0f}, posts {cyc[3]/max(cyc[6],1):.0f}, write-back {cyc[4]/max(cyc[6],1):.0f}, fence {cyc[5]/max(cyc[6],1):.0f})")
nrec = min(nrec, (len(body) - off) // 8)
rec = body[off:off + nrec * 8].reshape(nrec, 8)
tile = (rec[:, 0] & 0xffffffff).astype(np.int64); block = (rec[:, 0] >> 32).astype(np.int64)
t_claim, t_staged, t_relax, t_end = (rec[:, k].astype(np.int64) for k in (1, 2, 3, 4))
key0 = (rec[:, 5] & 0xffffffff).astype(np.float64) / 32768.0; thr = (rec[:, 5] >> 32).astype(np.float64) / 32768.0   # Q17.15
it_sum = (rec[:, 6] & 0xffffffff).astype(np.int64); it_max = ((rec[:, 6] >> 32) & 0xffff).astype(np.int64); merges = (rec[:, 6] >> 48).astype(np.int64)
first_chg = (rec[:, 7] & 0xffffffff).astype(np.int64); last_chg = (rec[:, 7] >> 32).astype(np.int64)
t0 = t_claim.min(); T = t_end.max() - t0
print(f"tiles {tx}x{ty}, blocks {grid}, visits {nrec}, span {T/1e3:.1f} us")
d_stage, d_relax, d_pub = t_staged - t_claim, t_relax - t_staged, t_end - t_relax
print(f"per visit (us): stage {d_stage.mean()/1e3:.2f}  relax {d_relax.mean()/1e3:.2f} (p50 {np.median(d_relax)/1e3:.2f}, p90 {np.percentile(d_relax,90)/1e3:.2f}, max {d_relax.max()/1e3:.1f})  publish {d_pub.mean()/1e3:.2f}")
print(f"sub-block passes per visit: sum {it_sum.mean():.0f}, busiest warp {it_max.mean():.0f}; merges per visit {merges.mean():.2f}")
print(f"relax phase: first change after {np.median(first_chg[first_chg>0])/1e3:.2f} us (p50), last change at {np.median(last_chg[last_chg>0])/1e3:.2f} us, "
      f"idle tail (relax end - last change - stage) p50 {np.median((t_relax - t_claim - last_chg)[last_chg>0])/1e3:.2f} us")
busy = np.zeros(grid)
np.add.at(busy, block, t_end - t_claim)
print(f"block busy fraction: mean {busy.mean()/T:.2f}, min {busy.min()/T:.2f}, max {busy.max()/T:.2f}; tiles in flight on average {busy.sum()/T:.1f}")
# claim latency: from the first post to a tile to the start of its first visit
first_visit = np.full(nt, np.iinfo(np.int64).max); np.minimum.at(first_visit, tile, t_claim)
ok = (first_post < np.iinfo(np.int64).max) & (first_visit < np.iinfo(np.int64).max) & (first_post > 0)
lat = (first_visit - first_post)[ok]
print(f"claim latency, first post -> first visit of the tile (us): p50 {np.median(lat)/1e3:.2f} p90 {np.percentile(lat,90)/1e3:.2f} max {lat.max()/1e3:.1f}  (tiles {ok.sum()})")
# every visit: was the owner block busy with ANOTHER tile when this tile's key was posted?  gap between a block's visits
order = np.lexsort((t_claim, block))
gaps = []
for b in range(grid):
    idx = order[block[order] == b]
    if len(idx) > 1: gaps.append(t_claim[idx][1:] - t_end[idx][:-1])
gaps = np.concatenate(gaps) if gaps else np.array([0])
print(f"gap between consecutive visits of a block (us): p50 {np.median(gaps)/1e3:.2f} p90 {np.percentile(gaps,90)/1e3:.2f}; visits per tile {nrec/nt:.2f}")
# wave progress: the smallest key0 among visits running at time t ~ where the front is
print("time (us) | visits running | min / median key0 of running visits | max thr reached so far")
for f in np.linspace(0.02, 0.98, 13):
    t = t0 + f * T
    run = (t_claim <= t) & (t_end >= t)
    done = t_end <= t
    print(f"{f*T/1e3:9.1f} | {run.sum():5d} | {key0[run].min() if run.any() else -1:8.1f} / {np.median(key0[run]) if run.any() else -1:8.1f} | {thr[done].max() if done.any() else 0:8.1f}")
if len(sys.argv) > 2:
    z = np.load(sys.argv[2]); gmin = z["gmin"].reshape(-1)
    reach = gmin < 1e5
    fv = (first_visit - t0).astype(np.float64)
    sel = reach & (first_visit < np.iinfo(np.int64).max)
    # wave speed: first-visit time of a tile against the tile's smallest cost
    A = np.vstack([gmin[sel], np.ones(sel.sum())]).T
    slope, icpt = np.linalg.lstsq(A, fv[sel], rcond=None)[0]
    print(f"first visit time vs tile cost distance: {slope:.1f} ns per cost unit (+{icpt/1e3:.1f} us); the whole solve: {T/ (gmin[reach].max()):.1f} ns per cost unit")
    # per-tile latency = first visit - (time the wave 'should' be there at the fastest observed speed)
    bins = np.linspace(0, gmin[reach].max(), 11)
    for a, b in zip(bins[:-1], bins[1:]):
        m = sel & (gmin >= a) & (gmin < b)
        if m.any(): print(f"  cost {a:7.0f}..{b:7.0f}: tiles {m.sum():5d} first visit p10 {np.percentile(fv[m],10)/1e3:8.1f} p50 {np.median(fv[m])/1e3:8.1f} p90 {np.percentile(fv[m],90)/1e3:8.1f} us")
