"""TEST INFRASTRUCTURE ONLY — ctypes front-ends for the two CPU checkers.

* ``Oracle``    -> oracle/_build/libnsfs_oracle.so, the plain-C restatement (nsfs_oracle.c)
* ``Reference`` -> oracle/_ref/libnsfs_ref.so, the UNMODIFIED reference sources behind ref_harness.cpp

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Both classes expose the same method names so a test can run either against the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libnsfs_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnsfs_ref.so")

ASTAR, DJIKSTRA, THETASTAR = 0, 1, 2
MODE_F, MODE_G, MODE_FG = 0, 1, 2
MODES = {"f": MODE_F, "g": MODE_G, "fg": MODE_FG}
E_RANGE = -1

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """(Re)build both checkers with oracle/Makefile (the _ref target is a no-op without /root/reference)."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "nsfs_oracle.c"))
    ):
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if os.path.isdir("/root/reference/src/global_planner") and (
        force or not os.path.exists(REF_SO)
        or os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(HERE, "ref_harness.cpp"))
    ):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


REF_OCC_SO = os.path.join(HERE, "_ref", "libnsfs_ref_occ.so")


def have_reference_occ() -> bool:
    return os.path.exists(REF_OCC_SO)


REF_CMD_SO = os.path.join(HERE, "_ref", "libnsfs_ref_cmd.so")


def have_reference_cmd() -> bool:
    return os.path.exists(REF_CMD_SO)


def cmd_check_trajectory_reference(infl, lo, xy, t_id, lo_thresh=10, cell=0.05, origin=(0.0, 0.0)):
    """The reference's own Commander::checkTrajectorySafety (commander.cpp:316-357, compiled unmodified): (status, safe, bad_idx),
    status E_RANGE where it throws std::out_of_range."""
    lib = C.CDLL(REF_CMD_SO)
    infl = np.ascontiguousarray(infl, np.int32); lo = np.ascontiguousarray(lo, np.int32)
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    H, W = infl.shape
    lib.nsref_cmd_check_trajectory.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                               C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    safe, bad = C.c_int(0), C.c_int(-1)
    rc = lib.nsref_cmd_check_trajectory(H, W, infl.ctypes.data, lo.ctypes.data, lo_thresh, cell, origin[0], origin[1], len(xy),
                                        xy.ctypes.data, int(t_id), C.byref(safe), C.byref(bad))
    if rc == 2:
        raise ValueError("the reference derived another map size from these extents")
    return (E_RANGE if rc == 1 else 0), bool(safe.value), int(bad.value)


def cmd_pos2idx_reference(xy, cell=0.05, origin=(0.0, 0.0)):
    lib = C.CDLL(REF_CMD_SO)
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    ij = np.zeros((len(xy), 2), np.int32)
    lib.nsref_cmd_pos2idx.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib.nsref_cmd_pos2idx(cell, origin[0], origin[1], len(xy), xy.ctypes.data, ij.ctypes.data)
    return ij


def occ_mask(inflation_radius=0.2, cell_size=0.05):
    """generateMask restated (oracle/nsfs_oracle.c nso_occ_mask): int32 [n, 2] offsets."""
    build()
    lib = C.CDLL(ORACLE_SO)
    lib.nso_occ_mask.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int]
    ab = np.zeros((4096, 2), np.int32)
    n = lib.nso_occ_mask(float(inflation_radius), float(cell_size), ab.ctypes.data_as(C.c_void_p), 4096)
    return ab[:n].copy()


def occ_inflation(lo, thresh=10, inflation_radius=0.2, cell_size=0.05):
    """(status, infl): the inflation plane the reference holds for this log-odds plane (nso_occ_inflation)."""
    build()
    lib = C.CDLL(ORACLE_SO)
    lo = np.ascontiguousarray(lo, np.int32)
    H, W = lo.shape
    ab = occ_mask(inflation_radius, cell_size)
    infl = np.zeros((H, W), np.int32)
    lib.nso_occ_inflation.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    rc = lib.nso_occ_inflation(H, W, lo.ctypes.data_as(C.c_void_p), int(thresh), ab.ctypes.data_as(C.c_void_p), len(ab),
                               infl.ctypes.data_as(C.c_void_p))
    return rc, infl


class ReferenceOcc:
    """The reference's OccupancyGrid node class, unmodified, behind oracle/ref_occ_harness.cpp (oracle/_ref only)."""

    def __init__(self, min_xy=(0.0, 0.0), max_xy=(10.0, 10.0), cell=0.05, inflation_radius=0.2, thresh=10, cap=20, max_range=3.499999):
        self.lib = C.CDLL(REF_OCC_SO)
        L = self.lib
        L.nsref_occ_create.restype = C.c_void_p
        L.nsref_occ_create.argtypes = [C.c_double] * 6 + [C.c_int, C.c_int, C.c_double]
        L.nsref_occ_destroy.argtypes = [C.c_void_p]
        L.nsref_occ_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.nsref_occ_integrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.nsref_occ_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.nsref_occ_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.nsref_occ_mask.argtypes = [C.c_void_p, C.c_void_p]
        self.ctx = L.nsref_occ_create(min_xy[0], min_xy[1], max_xy[0], max_xy[1], cell, inflation_radius, thresh, cap, max_range)
        h, w, n = C.c_int(0), C.c_int(0), C.c_int(0)
        L.nsref_occ_size(self.ctx, C.byref(h), C.byref(w), C.byref(n))
        self.H, self.W, self.n_mask = h.value, w.value, n.value

    def __del__(self):
        try:
            self.lib.nsref_occ_destroy(self.ctx)
        except Exception:
            pass

    def integrate(self, poses, ranges):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
        ranges = np.ascontiguousarray(ranges, np.float32).reshape(len(poses), 360)
        return self.lib.nsref_occ_integrate(self.ctx, len(poses), poses.ctypes.data_as(C.c_void_p), ranges.ctypes.data_as(C.c_void_p))

    def update(self, ij, occupy):
        ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
        occupy = np.ascontiguousarray(occupy, np.uint8)
        return self.lib.nsref_occ_update(self.ctx, len(ij), ij.ctypes.data_as(C.c_void_p), occupy.ctypes.data_as(C.c_void_p))

    def planes(self):
        lo = np.zeros((self.H, self.W), np.int32)
        infl = np.zeros((self.H, self.W), np.int32)
        self.lib.nsref_occ_get(self.ctx, lo.ctypes.data_as(C.c_void_p), infl.ctypes.data_as(C.c_void_p))
        return lo, infl

    def mask(self):
        ab = np.zeros((self.n_mask, 2), np.int32)
        self.lib.nsref_occ_mask(self.ctx, ab.ctypes.data_as(C.c_void_p))
        return ab


class _Stats(C.Structure):
    _fields_ = [("pops", C.c_longlong), ("expansions", C.c_longlong), ("pushes", C.c_longlong),
                ("heap_peak", C.c_longlong), ("los_checks", C.c_longlong)]


class _Map(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("infl", C.c_void_p), ("lo", C.c_void_p),
                ("lo_thresh", C.c_int), ("cell", C.c_double), ("ox", C.c_double), ("oy", C.c_double)]


class Oracle:
    """Plain-C restatement.  ``infl``/``lo`` are int32 [H, W] arrays (kept alive by the object)."""

    kind = "port"

    def __init__(self, infl, lo, lo_thresh=10, cell=0.05, origin=(0.0, 0.0)):
        build()
        self.lib = C.CDLL(ORACLE_SO)
        self.infl = np.ascontiguousarray(infl, dtype=np.int32)
        self.lo = np.ascontiguousarray(lo, dtype=np.int32)
        self.H, self.W = self.infl.shape
        self.map = _Map(self.H, self.W, self.infl.ctypes.data, self.lo.ctypes.data, lo_thresh, cell,
                        origin[0], origin[1])
        L = self.lib
        L.nso_dist_oct.restype = C.c_double
        L.nso_dist_euc.restype = C.c_double

    # -- planners ---------------------------------------------------------------------------------
    def plan(self, algo, mode, start, goal, want_field=False, cap=None):
        N = self.H * self.W
        cap = cap or N + 1
        path = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        gg = C.c_double(0)
        st = _Stats()
        g = np.zeros(N, np.float64) if want_field else None
        par = np.zeros(N, np.int32) if want_field else None
        vis = np.zeros(N, np.uint8) if want_field else None
        rc = self.lib.nso_plan(C.byref(self.map), algo, MODES.get(mode, mode) if isinstance(mode, str) else mode,
                               int(start[0]), int(start[1]), int(goal[0]), int(goal[1]),
                               path.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(gg),
                               g.ctypes.data_as(C.c_void_p) if want_field else None,
                               par.ctypes.data_as(C.c_void_p) if want_field else None,
                               vis.ctypes.data_as(C.c_void_p) if want_field else None, C.byref(st))
        out = dict(status=rc, path=path[: n.value].copy(), g_goal=gg.value, settled=st.expansions,
                   pops=st.pops, pushes=st.pushes, heap_peak=st.heap_peak, los_checks=st.los_checks)
        if want_field:
            out.update(g=g.reshape(self.H, self.W), parent=par.reshape(self.H, self.W),
                       visited=vis.reshape(self.H, self.W))
        return out

    def field(self, start, mode="g"):
        r = self.plan(DJIKSTRA, mode, start, (-7, -7), want_field=True, cap=4)
        return r

    def field_fixed_point(self, start):
        g = np.zeros(self.H * self.W, np.float64)
        rc = self.lib.nso_field_fixed_point(C.byref(self.map), int(start[0]), int(start[1]),
                                            g.ctypes.data_as(C.c_void_p))
        return rc, g.reshape(self.H, self.W)

    def field_relax_window(self, g, row_lo=0, row_hi=None, slack=0.0, cap=1e300):
        """Warm-started label-correcting relaxation of one row-block shard (g is float64 [H, W], updated in place);
        candidates above ``cap`` are left pending."""
        assert g.dtype == np.float64 and g.flags.c_contiguous and g.shape == (self.H, self.W)
        self.lib.nso_field_relax_window.restype = C.c_longlong
        self.lib.nso_field_relax_window.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        return self.lib.nso_field_relax_window(C.addressof(self.map), g.ctypes.data_as(C.c_void_p), int(row_lo),
                                               int(self.H if row_hi is None else row_hi), float(slack), float(cap))

    def field_min_pending(self, g, row_lo=0, row_hi=None, slack=0.0):
        self.lib.nso_field_min_pending.restype = C.c_double
        self.lib.nso_field_min_pending.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double]
        return self.lib.nso_field_min_pending(C.addressof(self.map), g.ctypes.data_as(C.c_void_p), int(row_lo),
                                              int(self.H if row_hi is None else row_hi), float(slack))

    def find_better_point(self, bad):
        fi, fj = C.c_int(0), C.c_int(0)
        st = _Stats()
        rc = self.lib.nso_find_better_point(C.byref(self.map), int(bad[0]), int(bad[1]), C.byref(fi), C.byref(fj),
                                            C.byref(st))
        return dict(status=rc, cell=(fi.value, fj.value), settled=st.expansions, pops=st.pops)

    # -- primitives -------------------------------------------------------------------------------
    def test_pos(self, i, j):
        return self.lib.nso_test_pos(C.byref(self.map), int(i), int(j))

    def bresenham(self, s, t, closed=False):
        n = int(max(abs(int(t[0]) - int(s[0])), abs(int(t[1]) - int(s[1])))) + 1
        out = np.zeros((n, 2), np.int32)
        fn = self.lib.nso_bresenham_closed if closed else self.lib.nso_bresenham
        m = fn(int(s[0]), int(s[1]), int(t[0]), int(t[1]), out.ctypes.data_as(C.c_void_p), n)
        assert m == n
        return out

    def is_los(self, s, t):
        return self.lib.nso_is_los(C.byref(self.map), int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def dist_oct(self, s, t):
        return self.lib.nso_dist_oct(int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def dist_euc(self, s, t):
        return self.lib.nso_dist_euc(int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def idx2pos(self, ij):
        ij = np.asarray(ij, np.float64).reshape(-1, 2)
        # grid_planner_core.cpp:416-421, evaluated in double exactly as written there
        return np.stack([ij[:, 0] * self.map.cell + self.map.ox, ij[:, 1] * self.map.cell + self.map.oy], 1)

    def _xy_call(self, fn, xy, with_map=True):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        out = np.zeros((max(len(xy), 2) + 2, 2), np.float64)
        args = [xy.ctypes.data_as(C.c_void_p), len(xy), out.ctypes.data_as(C.c_void_p), len(out)]
        n = fn(C.byref(self.map), *args) if with_map else fn(*args)
        return out[:n].copy() if n >= 0 else n

    def sparsify(self, xy):
        return self._xy_call(self.lib.nso_sparsify, xy, with_map=False)

    def any_angle(self, xy):
        return self._xy_call(self.lib.nso_any_angle, xy)

    def post_process(self, xy):
        return self._xy_call(self.lib.nso_post_process, xy)

    def generate_path(self, algo, mode, start, goal):
        """generatePath == plan + post_process_path, world coordinates (grid_planner_core.cpp:603-608)."""
        r = self.plan(algo, mode, start, goal)
        if r["status"] != 0:
            return r
        r["final_xy"] = self.post_process(self.idx2pos(r["path"]))
        return r

    def textbook_field(self, start):
        """Plain 8-connected Dijkstra (cardinal 1, diagonal sqrt 2, rows and columns bounds-checked): fp64 [H, W], 1e5 = unreached."""
        g = np.zeros(self.H * self.W, np.float64)
        rc = self.lib.nso_textbook_field(C.byref(self.map), int(start[0]), int(start[1]), g.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return g.reshape(self.H, self.W)

    def cmd_check_trajectory(self, ij, t_id):
        """Commander::checkTrajectorySafety restated (nso_cmd_check_trajectory): (status, safe, bad_idx)."""
        ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
        safe, bad = C.c_int(0), C.c_int(-1)
        rc = self.lib.nso_cmd_check_trajectory(C.byref(self.map), len(ij), ij.ctypes.data_as(C.c_void_p), int(t_id), C.byref(safe), C.byref(bad))
        return rc, bool(safe.value), int(bad.value)

    def pos2idx(self, xy):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        out = np.zeros((len(xy), 2), np.int32)
        i, j = C.c_int(0), C.c_int(0)
        for t in range(len(xy)):
            self.lib.nso_pos2idx(C.byref(self.map), C.c_double(xy[t, 0]), C.c_double(xy[t, 1]), C.byref(i), C.byref(j))
            out[t] = (i.value, j.value)
        return out

    def heap_trace(self, mode, ops):
        ops = np.ascontiguousarray(ops, np.int32).reshape(-1, 4)
        out = np.zeros(len(ops), np.int32)
        n = self.lib.nso_heap_trace(MODES[mode], len(ops), ops.ctypes.data_as(C.c_void_p),
                                    out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()


class Reference:
    """The reference's own planners (compiled from /root/reference, unmodified)."""

    kind = "reference"

    def __init__(self, infl, lo, lo_thresh=10, cell=0.05, origin=(0.0, 0.0)):
        if not have_reference():
            build()
        if not have_reference():
            raise FileNotFoundError(REF_SO)
        self.lib = L = C.CDLL(REF_SO)
        self.infl = np.ascontiguousarray(infl, dtype=np.int32)
        self.lo = np.ascontiguousarray(lo, dtype=np.int32)
        self.H, self.W = self.infl.shape
        self.cell, self.origin = cell, origin
        L.nsref_create.restype = C.c_void_p
        L.nsref_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
        L.nsref_destroy.argtypes = [C.c_void_p]
        L.nsref_dist_oct.restype = C.c_double
        L.nsref_dist_euc.restype = C.c_double
        self.ctx = C.c_void_p(L.nsref_create(self.H, self.W, cell, origin[0], origin[1], self.infl.ctypes.data,
                                             self.lo.ctypes.data, lo_thresh))

    def __del__(self):
        try:
            self.lib.nsref_destroy(self.ctx)
        except Exception:
            pass

    def plan(self, algo, mode, start, goal, post=False, cap=None):
        N = self.H * self.W
        cap = cap or N + 1
        xy = np.zeros((cap, 2), np.float64)
        ij = np.zeros((cap, 2), np.int32)
        n = C.c_int(0)
        gg = C.c_double(0)
        settled = C.c_longlong(0)
        sec = C.c_double(0)
        rc = self.lib.nsref_plan(self.ctx, algo, MODES.get(mode, mode) if isinstance(mode, str) else mode,
                                 int(start[0]), int(start[1]), int(goal[0]), int(goal[1]), int(post),
                                 xy.ctypes.data_as(C.c_void_p), ij.ctypes.data_as(C.c_void_p), cap,
                                 C.byref(n), C.byref(gg), C.byref(settled), C.byref(sec))
        return dict(status=rc, path=ij[: n.value].copy(), xy=xy[: n.value].copy(), g_goal=gg.value,
                    settled=settled.value, seconds=sec.value)

    def generate_path(self, algo, mode, start, goal):
        raw = self.plan(algo, mode, start, goal, post=False)
        fin = self.plan(algo, mode, start, goal, post=True)
        raw["final_xy"] = fin["xy"]
        raw["seconds"] = fin["seconds"]
        return raw

    def field(self, start, mode="g"):
        N = self.H * self.W
        g = np.zeros(N, np.float64)
        par = np.zeros(N, np.int32)
        vis = np.zeros(N, np.uint8)
        settled = C.c_longlong(0)
        sec = C.c_double(0)
        rc = self.lib.nsref_field(self.ctx, MODES[mode], int(start[0]), int(start[1]), g.ctypes.data_as(C.c_void_p),
                                  par.ctypes.data_as(C.c_void_p), vis.ctypes.data_as(C.c_void_p),
                                  C.byref(settled), C.byref(sec))
        return dict(status=rc, g=g.reshape(self.H, self.W), parent=par.reshape(self.H, self.W),
                    visited=vis.reshape(self.H, self.W), settled=settled.value, seconds=sec.value)

    def find_better_point(self, bad):
        fi, fj = C.c_int(0), C.c_int(0)
        x, y = C.c_double(0), C.c_double(0)
        settled = C.c_longlong(0)
        sec = C.c_double(0)
        rc = self.lib.nsref_find_better_point(self.ctx, int(bad[0]), int(bad[1]), C.byref(fi), C.byref(fj),
                                              C.byref(x), C.byref(y), C.byref(settled), C.byref(sec))
        return dict(status=rc, cell=(fi.value, fj.value), xy=(x.value, y.value), settled=settled.value,
                    seconds=sec.value)

    def test_pos(self, i, j):
        return self.lib.nsref_test_pos(self.ctx, int(i), int(j))

    def bresenham(self, s, t, closed=False):
        n = int(max(abs(int(t[0]) - int(s[0])), abs(int(t[1]) - int(s[1])))) + 1
        out = np.zeros((n, 2), np.int32)
        m = self.lib.nsref_bresenham(int(s[0]), int(s[1]), int(t[0]), int(t[1]), out.ctypes.data_as(C.c_void_p), n)
        assert m == n
        return out

    def is_los(self, s, t):
        return self.lib.nsref_is_los(self.ctx, int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def dist_oct(self, s, t):
        return self.lib.nsref_dist_oct(int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def dist_euc(self, s, t):
        return self.lib.nsref_dist_euc(int(s[0]), int(s[1]), int(t[0]), int(t[1]))

    def idx2pos(self, ij):
        ij = np.asarray(ij, np.int64).reshape(-1, 2)
        out = np.zeros((len(ij), 2), np.float64)
        x, y = C.c_double(0), C.c_double(0)
        for t, (i, j) in enumerate(ij):
            self.lib.nsref_idx2pos(self.ctx, int(i), int(j), C.byref(x), C.byref(y))
            out[t] = (x.value, y.value)
        return out

    def _xy_call(self, fn, xy, *pre):
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        out = np.zeros((max(len(xy), 2) + 2, 2), np.float64)
        n = fn(self.ctx, *pre, xy.ctypes.data_as(C.c_void_p), len(xy), out.ctypes.data_as(C.c_void_p), len(out))
        return out[:n].copy() if n >= 0 else n

    def sparsify(self, xy):
        return self._xy_call(self.lib.nsref_sparsify, xy)

    def any_angle(self, xy):
        return self._xy_call(self.lib.nsref_any_angle, xy)

    def post_process(self, xy, algo=ASTAR):
        return self._xy_call(self.lib.nsref_post_process, xy, algo)

    def heap_trace(self, mode, ops):
        ops = np.ascontiguousarray(ops, np.int32).reshape(-1, 4)
        out = np.zeros(len(ops), np.int32)
        n = self.lib.nsref_heap_trace(MODES[mode], len(ops), ops.ctypes.data_as(C.c_void_p),
                                      out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def field_batch(self, srcs, nthreads=1):
        srcs = np.ascontiguousarray(srcs, np.int32).reshape(-1, 2)
        settled = np.zeros(len(srcs), np.int64)
        sec = C.c_double(0)
        rc = self.lib.nsref_field_batch(self.ctx, len(srcs), srcs.ctypes.data_as(C.c_void_p), int(nthreads),
                                        settled.ctypes.data_as(C.c_void_p), C.byref(sec))
        return dict(status=rc, settled=settled, seconds=sec.value)

    def plan_batch(self, algo, mode, sg4, nthreads=1):
        sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
        nq = len(sg4)
        g = np.zeros(nq, np.float64)
        plen = np.zeros(nq, np.int32)
        settled = np.zeros(nq, np.int64)
        sec = C.c_double(0)
        rc = self.lib.nsref_plan_batch(self.ctx, algo, MODES[mode], nq, sg4.ctypes.data_as(C.c_void_p), int(nthreads),
                                       g.ctypes.data_as(C.c_void_p), plen.ctypes.data_as(C.c_void_p),
                                       settled.ctypes.data_as(C.c_void_p), C.byref(sec))
        return dict(status=rc, g_goal=g, path_len=plen, settled=settled, seconds=sec.value)
