"""ctypes binding of libnsfs_gpu.so (include/nsfs.h) for the Python test and bench drivers.

This is a thin 1:1 wrapper: numpy arrays in, numpy arrays out, every call goes through the C-ABI.  There is no
Python or CPU implementation behind it; if the library is missing the import of a planner fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_SO = os.environ.get("NSFS_GPU_SO") or os.path.join(PKG, "libnsfs_gpu.so")   # override: instrumented dev builds only

OK, E_RANGE, E_CUDA, E_TRUNC, E_ARG, E_NOMEM, E_HEAP, E_NOMAP, E_NOFIELD = 0, -1, -2, -3, -4, -5, -6, -7, -8
ASTAR, DJIKSTRA, THETASTAR, DJIKSTRA_FIELD = 0, 1, 2, 3
MODE_F, MODE_G, MODE_FG = 0, 1, 2
MODES = {"f": MODE_F, "g": MODE_G, "fg": MODE_FG}
ALGOS = {"astar": ASTAR, "djikstra": DJIKSTRA, "thetastar": THETASTAR, "djikstra_field": DJIKSTRA_FIELD}


class NsfsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nsfs error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("pops", C.c_longlong), ("expansions", C.c_longlong), ("pushes", C.c_longlong),
                ("heap_peak", C.c_longlong), ("los_checks", C.c_longlong), ("rounds", C.c_longlong),
                ("tile_visits", C.c_longlong), ("launches", C.c_longlong), ("gpu_ms", C.c_float),
                ("main_kernel_ms", C.c_float)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PeerInfo(C.Structure):
    """nsfs_peer_info: CUDA IPC handles of a shard's g / tile-key / counter arrays (picklable through bytes())."""
    _fields_ = [("g", C.c_ubyte * 64), ("key", C.c_ubyte * 64), ("counts", C.c_ubyte * 64),
                ("rows", C.c_int), ("W", C.c_int), ("tiles_x", C.c_int), ("tiles_y", C.c_int)]


_lib = None


def lib():
    """Load libnsfs_gpu.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(GPU_SO):
            raise FileNotFoundError(f"{GPU_SO} not built; run python nav-stack-from-scratch_b200/build.py")
        L = C.CDLL(GPU_SO)
        vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
        L.nsfs_create.argtypes = [C.POINTER(vp), i32, i32, i32]
        L.nsfs_destroy.argtypes = [vp]
        L.nsfs_destroy.restype = None
        L.nsfs_set_map.argtypes = [vp, vp, vp, i32]
        L.nsfs_set_map_device.argtypes = [vp, vp, vp, i32]
        L.nsfs_set_map_logodds.argtypes = [vp, vp, i32, i32, dbl, dbl]
        L.nsfs_get_inflation.argtypes = [vp, vp]
        L.nsfs_update_map_rows.argtypes = [vp, i32, i32, vp, vp]
        L.nsfs_map_words.argtypes = [vp]
        L.nsfs_get_packed_map_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.nsfs_set_packed_map_device.argtypes = [vp, vp, vp]
        L.nsfs_test_pos_batch.argtypes = [vp, i32, vp, vp]
        L.nsfs_check_trajectory.argtypes = [vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.nsfs_plan.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32, C.POINTER(i32), C.POINTER(dbl), C.POINTER(Stats)]
        L.nsfs_last_path.argtypes = [vp, vp, i32, C.POINTER(i32)]
        L.nsfs_pin_host.argtypes = [vp, vp, C.c_size_t]
        L.nsfs_unpin_host.argtypes = [vp, vp]
        L.nsfs_group_unique_id.argtypes = [vp]
        L.nsfs_group_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32]
        L.nsfs_group_destroy.argtypes = [vp]
        L.nsfs_group_destroy.restype = None
        L.nsfs_group_rank.argtypes = [vp]
        L.nsfs_group_world.argtypes = [vp]
        L.nsfs_group_barrier.argtypes = [vp]
        L.nsfs_group_error.argtypes = [vp]
        L.nsfs_group_error.restype = C.c_char_p
        L.nsfs_group_broadcast_map.argtypes = [vp, vp, i32]
        L.nsfs_group_plan_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, C.POINTER(Stats)]
        L.nsfs_group_field_attach.argtypes = [vp, vp, i32, i32, i32]
        L.nsfs_group_field_solve.argtypes = [vp, vp, i32, i32, C.POINTER(C.c_longlong), C.POINTER(Stats)]
        L.nsfs_plan_batch.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, C.POINTER(Stats)]
        L.nsfs_plan_batch_device.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, C.POINTER(Stats)]
        L.nsfs_field.argtypes = [vp, i32, i32, vp, vp, C.POINTER(Stats)]
        L.nsfs_field_device.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(Stats)]
        L.nsfs_field_begin.argtypes = [vp, i32, i32]
        L.nsfs_field_inject_rows.argtypes = [vp, i32, i32, vp, C.POINTER(i32)]
        L.nsfs_field_relax.argtypes = [vp, C.POINTER(Stats)]
        L.nsfs_field_relax_until.argtypes = [vp, C.c_float, C.POINTER(C.c_float), C.POINTER(Stats)]
        L.nsfs_field_min_pending.argtypes = [vp, C.POINTER(C.c_float)]
        L.nsfs_field_finish.argtypes = [vp, C.POINTER(Stats)]
        L.nsfs_field_pointers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.nsfs_field_peer_export.argtypes = [vp, C.POINTER(PeerInfo)]
        L.nsfs_field_peer_attach.argtypes = [vp, i32, C.POINTER(PeerInfo), i32]
        L.nsfs_field_peer_counter.argtypes = [vp, C.POINTER(PeerInfo)]
        L.nsfs_field_peer_rows.argtypes = [vp, i32, i32]
        L.nsfs_field_peer_detach.argtypes = [vp]
        L.nsfs_field_pending_count.argtypes = [vp, C.POINTER(i32)]
        L.nsfs_field_set_counter.argtypes = [vp, i32]
        L.nsfs_field_relax_p2p.argtypes = [vp, C.POINTER(Stats)]
        L.nsfs_field_path.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32), C.POINTER(dbl)]
        L.nsfs_find_better_point.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(Stats)]
        L.nsfs_find_better_point_batch.argtypes = [vp, i32, vp, vp, vp, C.POINTER(Stats)]
        L.nsfs_los_batch.argtypes = [vp, i32, vp, vp]
        L.nsfs_any_angle.argtypes = [vp, i32, vp, vp]
        L.nsfs_sync.argtypes = [vp]
        L.nsfs_kernel_launches.argtypes = [vp]
        L.nsfs_kernel_launches.restype = C.c_longlong
        L.nsfs_stream.argtypes = [vp]
        L.nsfs_stream.restype = vp
        L.nsfs_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
        L.nsfs_strerror.argtypes = [i32]
        L.nsfs_strerror.restype = C.c_char_p
        L.nsfs_last_cuda_error.argtypes = [vp]
        L.nsfs_last_cuda_error.restype = C.c_char_p
        L.nsfs_version.restype = C.c_char_p
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Group:
    """nsfs_group_*: this rank of a one-process-per-GPU group (NCCL bound inside libnsfs_gpu.so).  ``dist`` (torch.distributed)
    only carries the 128-byte id from rank 0 to the others."""

    def __init__(self, dist, rank, world, device):
        import torch
        self.L = lib()
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            rc = self.L.nsfs_group_unique_id(buf.ctypes.data_as(C.c_void_p))
            if rc != OK:
                raise NsfsError(rc, "nsfs_group_unique_id (is libnccl.so.2 loadable?)")
        t = torch.from_numpy(buf).to(torch.device("cuda", device))
        dist.broadcast(t, src=0)
        buf = np.ascontiguousarray(t.cpu().numpy())
        self.g = C.c_void_p()
        rc = self.L.nsfs_group_create(C.byref(self.g), buf.ctypes.data_as(C.c_void_p), int(rank), int(world), int(device))
        if rc != OK:
            self.g = None
            raise NsfsError(rc, "nsfs_group_create")
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "g", None):
            self.L.nsfs_group_destroy(self.g)
            self.g = None

    def _check(self, rc, allow=()):
        if rc != OK and rc not in allow:
            raise NsfsError(rc, self.L.nsfs_strerror(rc).decode() + " | " + self.L.nsfs_group_error(self.g).decode())
        return rc

    def barrier(self):
        self._check(self.L.nsfs_group_barrier(self.g))

    def broadcast_map(self, planner, root=0):
        self._check(self.L.nsfs_group_broadcast_map(self.g, planner.h, int(root)))

    def plan_batch(self, planner, algo, mode, sg4):
        sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
        nq = len(sg4)
        status, plen = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        g, settled = np.zeros(nq, np.float64), np.zeros(nq, np.int64)
        st = Stats()
        self._check(self.L.nsfs_group_plan_batch(self.g, planner.h, ALGOS.get(algo, algo), MODES.get(mode, mode), nq, _ptr(sg4), _ptr(status),
                                                 _ptr(g), _ptr(plen), _ptr(settled), C.byref(st)))
        return dict(status=status, g_goal=g, path_len=plen, settled=settled, stats=st.asdict())

    def field_attach(self, planner, held_lo, own_lo, own_hi):
        self._check(self.L.nsfs_group_field_attach(self.g, planner.h, int(held_lo), int(own_lo), int(own_hi)))

    def field_solve(self, planner, src):
        reached, st = C.c_longlong(0), Stats()
        rc = self.L.nsfs_group_field_solve(self.g, planner.h, int(src[0]), int(src[1]), C.byref(reached), C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, reached_total=reached.value, stats=st.asdict())


class Planner:
    """One nsfs handle (one CUDA stream) bound to an H x W grid on ``device``."""

    def __init__(self, H, W, device=0):
        self.L = lib()
        self.H, self.W, self.N = int(H), int(W), int(H) * int(W)
        self.h = C.c_void_p()
        rc = self.L.nsfs_create(C.byref(self.h), self.H, self.W, int(device))
        if rc != OK:
            self.h = None
            raise NsfsError(rc, self.L.nsfs_strerror(rc).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.nsfs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != OK and rc not in allow:
            msg = self.L.nsfs_strerror(rc).decode()
            if rc == E_CUDA:
                msg += " | " + self.L.nsfs_last_cuda_error(self.h).decode()
            raise NsfsError(rc, msg)
        return rc

    # -- map ---------------------------------------------------------------------------------------
    def set_map(self, infl, lo, lo_thresh=10):
        """Host int32 arrays (any array exposing a contiguous buffer, e.g. a pinned torch tensor's numpy view)."""
        infl = np.ascontiguousarray(infl, np.int32)
        lo = np.ascontiguousarray(lo, np.int32)
        assert infl.size == self.N and lo.size == self.N
        self._check(self.L.nsfs_set_map(self.h, _ptr(infl), _ptr(lo), int(lo_thresh)))

    def set_map_logodds(self, lo, lo_thresh=10, occ_thresh=10, inflation_radius=0.2, cell_size=0.05):
        """Only the log-odds plane travels; the inflation plane is computed on the device (the map producer's last step)."""
        lo = np.ascontiguousarray(lo, np.int32)
        assert lo.size == self.N
        rc = self.L.nsfs_set_map_logodds(self.h, _ptr(lo), int(lo_thresh), int(occ_thresh), float(inflation_radius), float(cell_size))
        self._check(rc, allow=(E_RANGE,))
        return rc

    def inflation(self):
        out = np.zeros((self.H, self.W), np.int32)
        self._check(self.L.nsfs_get_inflation(self.h, _ptr(out)))
        return out

    def update_map_rows(self, row0, infl_rows, lo_rows):
        """Rows [row0, row0 + len(infl_rows)) changed since set_map: upload and rebuild only what can see them."""
        infl_rows = np.ascontiguousarray(infl_rows, np.int32).reshape(-1, self.W)
        lo_rows = np.ascontiguousarray(lo_rows, np.int32).reshape(-1, self.W)
        assert infl_rows.shape == lo_rows.shape
        self._check(self.L.nsfs_update_map_rows(self.h, int(row0), len(infl_rows), _ptr(infl_rows), _ptr(lo_rows)))

    def set_map_device(self, d_infl_ptr, d_lo_ptr, lo_thresh=10):
        self._check(self.L.nsfs_set_map_device(self.h, C.c_void_p(d_infl_ptr), C.c_void_p(d_lo_ptr), int(lo_thresh)))

    def map_words(self):
        return self.L.nsfs_map_words(self.h)

    def packed_map_device(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.L.nsfs_get_packed_map_device(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_packed_map_device(self, d_free_ptr, d_infl_ptr):
        self._check(self.L.nsfs_set_packed_map_device(self.h, C.c_void_p(d_free_ptr), C.c_void_p(d_infl_ptr)))

    def test_pos(self, ij):
        ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
        out = np.zeros(len(ij), np.int8)
        self._check(self.L.nsfs_test_pos_batch(self.h, len(ij), _ptr(ij), _ptr(out)))
        return out

    def check_trajectory(self, ij, t_id):
        """Commander::checkTrajectorySafety on trajectory cells: (status, safe, bad_idx); status E_RANGE where the reference throws."""
        ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
        safe, bad = C.c_int32(0), C.c_int32(-1)
        rc = self.L.nsfs_check_trajectory(self.h, len(ij), _ptr(ij) if len(ij) else None, int(t_id), C.byref(safe), C.byref(bad))
        self._check(rc, allow=(E_RANGE,))
        return rc, bool(safe.value), int(bad.value)

    # -- searches ----------------------------------------------------------------------------------
    def plan(self, algo, mode, start, goal, cap=None, want_path=True):
        cap = int(cap if cap is not None else self.N + 1)
        path = np.zeros((cap, 2), np.int32) if want_path else None
        n, g, st = C.c_int(0), C.c_double(0), Stats()
        rc = self.L.nsfs_plan(self.h, ALGOS.get(algo, algo), MODES.get(mode, mode), int(start[0]), int(start[1]),
                              int(goal[0]), int(goal[1]), _ptr(path), cap, C.byref(n), C.byref(g), C.byref(st))
        self._check(rc, allow=(E_RANGE, E_TRUNC, E_HEAP))
        out = dict(status=rc, g_goal=g.value, n_path=n.value, settled=st.expansions, stats=st.asdict())
        if want_path:
            out["path"] = path[: min(n.value, cap)].copy()
        return out

    def last_path(self, cap):
        """The path of the last plan() again without searching (status E_NOFIELD when the search state is gone)."""
        path = np.zeros((int(cap), 2), np.int32)
        n = C.c_int(0)
        rc = self.L.nsfs_last_path(self.h, _ptr(path), int(cap), C.byref(n))
        self._check(rc, allow=(E_TRUNC, E_NOFIELD))
        return dict(status=rc, path=path[: min(n.value, int(cap))].copy(), n_path=n.value)

    def plan_batch(self, algo, mode, sg4, path_cap=0):
        sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
        nq = len(sg4)
        status = np.zeros(nq, np.int32)
        g = np.zeros(nq, np.float64)
        plen = np.zeros(nq, np.int32)
        settled = np.zeros(nq, np.int64)
        paths = np.zeros((nq, path_cap, 2), np.int32) if path_cap > 0 else None
        st = Stats()
        self._check(self.L.nsfs_plan_batch(self.h, ALGOS.get(algo, algo), MODES.get(mode, mode), nq, _ptr(sg4), _ptr(status),
                                           _ptr(g), _ptr(plen), _ptr(settled), _ptr(paths), int(path_cap), C.byref(st)))
        return dict(status=status, g_goal=g, path_len=plen, settled=settled, paths=paths, stats=st.asdict())

    def plan_batch_device(self, algo, mode, nq, d_sg4, d_status, d_g, d_len, d_settled, d_paths=0, path_cap=0, want_stats=False):
        st = Stats()
        self._check(self.L.nsfs_plan_batch_device(self.h, ALGOS.get(algo, algo), MODES.get(mode, mode), int(nq),
                                                  C.c_void_p(d_sg4), C.c_void_p(d_status), C.c_void_p(d_g), C.c_void_p(d_len),
                                                  C.c_void_p(d_settled), C.c_void_p(d_paths) if d_paths else None, int(path_cap),
                                                  C.byref(st) if want_stats else None))
        return st.asdict() if want_stats else None

    def field(self, start, want_g=True, want_pred=True):
        g = np.zeros((self.H, self.W), np.float32) if want_g else None
        pred = np.zeros((self.H, self.W), np.uint8) if want_pred else None
        st = Stats()
        rc = self.L.nsfs_field(self.h, int(start[0]), int(start[1]), _ptr(g), _ptr(pred), C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, g=g, pred=pred, reached=st.expansions, stats=st.asdict())

    def field_device(self, start):
        dg, dp, st = C.c_void_p(), C.c_void_p(), Stats()
        rc = self.L.nsfs_field_device(self.h, int(start[0]), int(start[1]), C.byref(dg), C.byref(dp), C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, d_g=dg.value, d_pred=dp.value, reached=st.expansions, stats=st.asdict())

    # -- row-partitioned field: caller-driven phases (nsfs/sharded_field.py) -----------------------
    def field_begin(self, start=None):
        si, sj = (-1, -1) if start is None else (int(start[0]), int(start[1]))
        self._check(self.L.nsfs_field_begin(self.h, si, sj))

    def field_inject_rows(self, row0, nrows, d_vals_ptr):
        n = C.c_int(0)
        self._check(self.L.nsfs_field_inject_rows(self.h, int(row0), int(nrows), C.c_void_p(d_vals_ptr), C.byref(n)))
        return n.value

    def field_relax(self, cap=None):
        """Relax to the local fixed point, or (cap given) only up to that cost; min_pending = smallest value left pending."""
        st, mp = Stats(), C.c_float(0)
        rc = self.L.nsfs_field_relax_until(self.h, C.c_float(3.4e38 if cap is None else float(cap)), C.byref(mp), C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, min_pending=mp.value, stats=st.asdict())

    def field_min_pending(self):
        mp = C.c_float(0)
        self._check(self.L.nsfs_field_min_pending(self.h, C.byref(mp)))
        return mp.value

    def field_finish(self):
        st = Stats()
        rc = self.L.nsfs_field_finish(self.h, C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, reached=st.expansions, stats=st.asdict())

    def field_pointers(self):
        dg, dp = C.c_void_p(), C.c_void_p()
        self._check(self.L.nsfs_field_pointers(self.h, C.byref(dg), C.byref(dp)))
        return dg.value, dp.value

    # -- peer-to-peer shards (NVLink, one process per GPU) -------------------------------------------
    def field_peer_export(self):
        info = PeerInfo()
        self._check(self.L.nsfs_field_peer_export(self.h, C.byref(info)))
        return bytes(info)

    def field_peer_attach(self, side, info_bytes, row_off):
        info = PeerInfo.from_buffer_copy(info_bytes)
        self._check(self.L.nsfs_field_peer_attach(self.h, int(side), C.byref(info), int(row_off)))

    def field_peer_counter(self, info_bytes=None):
        if info_bytes is None:
            self._check(self.L.nsfs_field_peer_counter(self.h, None))
        else:
            info = PeerInfo.from_buffer_copy(info_bytes)
            self._check(self.L.nsfs_field_peer_counter(self.h, C.byref(info)))

    def field_peer_rows(self, own_lo, own_hi):
        self._check(self.L.nsfs_field_peer_rows(self.h, int(own_lo), int(own_hi)))

    def field_peer_detach(self):
        self._check(self.L.nsfs_field_peer_detach(self.h))

    def field_pending_count(self):
        n = C.c_int(0)
        self._check(self.L.nsfs_field_pending_count(self.h, C.byref(n)))
        return n.value

    def field_set_counter(self, total):
        self._check(self.L.nsfs_field_set_counter(self.h, int(total)))

    def field_relax_p2p(self):
        st = Stats()
        rc = self.L.nsfs_field_relax_p2p(self.h, C.byref(st))
        self._check(rc, allow=(E_RANGE,))
        return dict(status=rc, min_pending=self.field_min_pending(), stats=st.asdict())

    def field_path(self, goal, cap=None):
        cap = int(cap if cap is not None else self.N + 1)
        path = np.zeros((cap, 2), np.int32)
        n, cost = C.c_int(0), C.c_double(0)
        rc = self.L.nsfs_field_path(self.h, int(goal[0]), int(goal[1]), _ptr(path), cap, C.byref(n), C.byref(cost))
        self._check(rc, allow=(E_TRUNC,))
        return dict(status=rc, path=path[: min(n.value, cap)].copy(), n_path=n.value, cost=cost.value)

    def find_better_point(self, bad):
        fi, fj, st = C.c_int(0), C.c_int(0), Stats()
        rc = self.L.nsfs_find_better_point(self.h, int(bad[0]), int(bad[1]), C.byref(fi), C.byref(fj), C.byref(st))
        self._check(rc, allow=(E_RANGE, E_HEAP))
        return dict(status=rc, cell=(fi.value, fj.value), settled=st.expansions, stats=st.asdict())

    def find_better_point_batch(self, bad_ij):
        bad_ij = np.ascontiguousarray(bad_ij, np.int32).reshape(-1, 2)
        out = np.zeros_like(bad_ij)
        status = np.zeros(len(bad_ij), np.int32)
        st = Stats()
        self._check(self.L.nsfs_find_better_point_batch(self.h, len(bad_ij), _ptr(bad_ij), _ptr(out), _ptr(status), C.byref(st)))
        return dict(cells=out, status=status, stats=st.asdict())

    # -- line of sight -----------------------------------------------------------------------------
    def los_batch(self, seg4):
        seg4 = np.ascontiguousarray(seg4, np.int32).reshape(-1, 4)
        out = np.zeros(len(seg4), np.int8)
        self._check(self.L.nsfs_los_batch(self.h, len(seg4), _ptr(seg4), _ptr(out)))
        return out

    def any_angle(self, pts_ij):
        pts = np.ascontiguousarray(pts_ij, np.int32).reshape(-1, 2)
        keep = np.zeros(len(pts), np.uint8)
        rc = self.L.nsfs_any_angle(self.h, len(pts), _ptr(pts), _ptr(keep))
        self._check(rc, allow=(E_RANGE,))
        return rc, keep

    # -- service -----------------------------------------------------------------------------------
    def sync(self):
        self._check(self.L.nsfs_sync(self.h))

    def launches(self):
        return self.L.nsfs_kernel_launches(self.h)

    def stream(self):
        return self.L.nsfs_stream(self.h)

    def set_option(self, key, value):
        self._check(self.L.nsfs_set_option(self.h, key.encode(), int(value)))
