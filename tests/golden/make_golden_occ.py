"""Generate tests/golden/golden_occ_v1.npz by RUNNING THE REFERENCE'S MAP PRODUCER ITSELF (oracle/_ref/libnsfs_ref_occ.so =
src/loco_mapping/src/occupancy_grid.cpp compiled unmodified; needs /root/reference, i.e. this container).

Three cases, each = the log-odds and inflation planes the node would publish on grid/mapinfo/* plus what produced them:
  scan   a 6.4 m x 4.8 m room (128 x 96 cells), 40 synthetic lidar scans from a robot walking through it, integrated by
         OccupancyGrid::run() (occupancy_grid.cpp:91-164): ray casting, log-odds caps, threshold crossings both ways
  edge   raw updateLogOdds events on and around the rows / columns where the reference's index quirks bite (row 0 is
         "out of bounds", discs spill across column edges into the neighbouring row)
  throw  an occupied cell in the last row at the last columns: updateInflation indexes past the grids (std::out_of_range)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import ReferenceOcc  # noqa: E402


def room_scans(n_scans, seed):
    """Poses inside a rectangular room with a few box obstacles; ranges by exact ray / segment intersection."""
    rng = np.random.RandomState(seed)
    W_, H_ = 6.4, 4.8
    boxes = [(1.0, 1.0, 1.6, 1.5), (3.0, 2.5, 3.4, 3.6), (4.6, 0.8, 5.4, 1.2)]
    segs = []
    for (x0, y0, x1, y1) in [(0.3, 0.3, W_ - 0.3, H_ - 0.3)] + boxes:
        segs += [((x0, y0), (x1, y0)), ((x1, y0), (x1, y1)), ((x1, y1), (x0, y1)), ((x0, y1), (x0, y0))]
    poses, ranges = [], []
    for s in range(n_scans):
        t = s / max(n_scans - 1, 1)
        px, py = 0.9 + 4.2 * t, 2.1 + 0.9 * np.sin(3.0 * t)
        hd = rng.uniform(-3.0, 3.0)
        r = np.full(360, 10.0, np.float32)
        for a in range(360):
            ang = np.pi * a / 180.0 + hd
            dx, dy = np.cos(ang), np.sin(ang)
            best = 10.0
            for (ax, ay), (bx, by) in segs:
                ex, ey = bx - ax, by - ay
                den = dx * ey - dy * ex
                if abs(den) < 1e-12:
                    continue
                tt = ((ax - px) * ey - (ay - py) * ex) / den
                uu = ((ax - px) * dy - (ay - py) * dx) / den
                if tt > 1e-9 and 0.0 <= uu <= 1.0 and tt < best:
                    best = tt
            r[a] = best + rng.normal(0, 0.004)
        poses.append((px, py, hd)); ranges.append(r)
    return np.array(poses, np.float64), np.array(ranges, np.float32)


def main():
    out = {}
    R = ReferenceOcc(min_xy=(0.0, 0.0), max_xy=(6.4, 4.8), cell=0.05)
    poses, ranges = room_scans(40, 5)
    assert R.integrate(poses, ranges) == 0
    lo, infl = R.planes()
    out["scan/poses"], out["scan/ranges"], out["scan/lo"], out["scan/infl"] = poses, ranges, lo.astype(np.int8), infl.astype(np.int16)
    out["mask"] = R.mask()
    assert (lo >= 10).sum() > 200 and infl.max() > 5

    R = ReferenceOcc(min_xy=(0.0, 0.0), max_xy=(3.2, 4.0), cell=0.05)          # 64 x 80
    H, W = R.H, R.W
    rng = np.random.RandomState(7)
    cells = [(0, 5), (1, 0), (1, W - 1), (2, W - 2), (3, 1), (H - 2, 0), (H - 6, W - 1), (H // 2, 0), (H // 2, W - 1), (4, 40), (4, 41), (5, 40)]
    ev = []
    for c in cells:
        ev += [(c, 1)] * 12                              # up through the threshold
    for c in cells[::3]:
        ev += [(c, 0)] * 5                               # a few drop below it again
    for _ in range(400):
        c = (int(rng.randint(-2, H + 2)), int(rng.randint(0, W)))            # rows may be out of range (oob returns), columns are the cell's own
        ev += [(c, int(rng.rand() < 0.6))] * int(rng.randint(1, 14))
    ev = [e for e in ev if not (e[0][0] >= H - 1 and e[0][1] >= W - 5)]      # keep clear of the throwing corner here
    ij = np.array([e[0] for e in ev], np.int32); oc = np.array([e[1] for e in ev], np.uint8)
    assert R.update(ij, oc) == 0
    lo, infl = R.planes()
    out["edge/events_ij"], out["edge/events_occ"], out["edge/lo"], out["edge/infl"] = ij, oc, lo.astype(np.int8), infl.astype(np.int16)

    R = ReferenceOcc(min_xy=(0.0, 0.0), max_xy=(3.2, 4.0), cell=0.05)
    ij = np.array([(R.H - 1, R.W - 2)] * 10, np.int32); oc = np.ones(10, np.uint8)
    out["throw/rc"] = np.array([R.update(ij, oc)], np.int32)
    assert out["throw/rc"][0] == 1
    out["throw/events_ij"], out["throw/events_occ"] = ij, oc
    lo = np.zeros((R.H, R.W), np.int8); lo[R.H - 1, R.W - 2] = 10
    out["throw/lo_at_throw"] = lo
    path = os.path.join(HERE, "golden_occ_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
