"""ctypes front-end of the C++ planner children (host/gpu_planners.cpp) through their nsfsh_* test shim.

``Session`` is what GlobalPlanner holds (reference global_planner.h:52-53): a main planner picked by
planner_name / cost_mode and a Djikstra fallback prepared with cost mode "g".  The default library is
libnsfs_planners.so, the in-repo build against the interface mirror (host/planner_iface.hpp); ``so_path`` selects
another build of the same children, e.g. the one the tests compile against the reference's OWN grid_planner_core.h
where the reference tree exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(PKG)
HOST_SO = os.path.join(PKG, "libnsfs_planners.so")

SYMBOLS = ("nsfsh_create", "nsfsh_destroy", "nsfsh_error", "nsfsh_set_map", "nsfsh_generate_path_idx",
           "nsfsh_generate_path_pos", "nsfsh_find_better_point_pos", "nsfsh_find_better_point_idx", "nsfsh_load_config",
           "nsfsh_main_planner_kind", "nsfsh_sparsify", "nsfsh_plan_request", "nsfsh_check_trajectory", "nsfsh_set_graph", "nsfsh_set_map_sync",
           "nsfsh_notify_map_changed", "nsfsh_plan_batch_idx", "nsfsh_group_id", "nsfsh_group_create", "nsfsh_group_destroy",
           "nsfsh_plan_batch_sharded")

_libs = {}


def lib(so_path=None):
    path = so_path or HOST_SO
    if path not in _libs:
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built")
        L = C.CDLL(path)
        vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
        L.nsfsh_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, i32, i32, dbl, dbl, dbl, i32, vp, vp, i32, C.POINTER(i32)]
        L.nsfsh_create.restype = vp
        L.nsfsh_destroy.argtypes = [vp]
        L.nsfsh_destroy.restype = None
        L.nsfsh_error.argtypes = [vp]
        L.nsfsh_error.restype = C.c_char_p
        L.nsfsh_set_map.argtypes = [vp, vp, vp]
        L.nsfsh_generate_path_idx.argtypes = [vp, i32, i32, i32, i32, vp, i32, C.POINTER(i32)]
        L.nsfsh_generate_path_pos.argtypes = [vp, dbl, dbl, dbl, dbl, vp, i32, C.POINTER(i32)]
        L.nsfsh_find_better_point_pos.argtypes = [vp, dbl, dbl, C.POINTER(dbl), C.POINTER(dbl)]
        L.nsfsh_find_better_point_idx.argtypes = [vp, i32, i32, C.POINTER(dbl), C.POINTER(dbl)]
        L.nsfsh_load_config.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i32)]
        L.nsfsh_main_planner_kind.argtypes = [vp]
        L.nsfsh_sparsify.argtypes = [vp, i32, vp, i32]
        L.nsfsh_plan_request.argtypes = [vp, dbl, dbl, dbl, dbl, vp, i32, C.POINTER(i32), vp, vp, vp]
        L.nsfsh_set_graph.argtypes = [vp, i32]
        L.nsfsh_check_trajectory.argtypes = [vp, i32, vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.nsfsh_set_map_sync.argtypes = [vp, i32]
        L.nsfsh_notify_map_changed.argtypes = [vp]
        L.nsfsh_plan_batch_idx.argtypes = [vp, i32, vp, vp, vp, vp, vp, C.POINTER(dbl)]
        L.nsfsh_group_id.argtypes = [vp]
        L.nsfsh_group_create.argtypes = [vp, i32, i32, i32]
        L.nsfsh_group_create.restype = vp
        L.nsfsh_group_destroy.argtypes = [vp]
        L.nsfsh_group_destroy.restype = None
        L.nsfsh_plan_batch_sharded.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, C.POINTER(dbl)]
        _libs[path] = L
    return _libs[path]


def load_config(yaml_path, so_path=None):
    L = lib(so_path)
    a, b, c = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(32)
    rate, verbose, device = C.c_double(0), C.c_int(0), C.c_int(0)
    rc = L.nsfsh_load_config(str(yaml_path).encode(), a, b, c, C.byref(rate), C.byref(verbose), C.byref(device))
    if rc != 0:
        raise RuntimeError(f"cannot read {yaml_path}")
    return dict(planner_name=a.value.decode(), cost_mode=b.value.decode(), djikstra_backend=c.value.decode(),
                gp_rate=rate.value, verbose_planner=bool(verbose.value), device=device.value)


def sparsify(xy, so_path=None):
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    out = np.zeros_like(xy)
    n = lib(so_path).nsfsh_sparsify(xy.ctypes.data_as(C.c_void_p), len(xy), out.ctypes.data_as(C.c_void_p), len(xy))
    return out[:n].copy()


class PlannerError(RuntimeError):
    pass


def make_group(dist, rank, world, device, so_path=None):
    """A nsfs_host::DeviceGroup for this rank.  The NCCL id is made by rank 0 in C++ and travels over torch.distributed here
    only because the test / bench driver already has it; a deployment ships the 128 bytes any way it likes."""
    import torch
    L = lib(so_path)
    buf = np.zeros(128, np.uint8)
    if rank == 0 and L.nsfsh_group_id(buf.ctypes.data_as(C.c_void_p)) != 0:
        raise PlannerError("nsfs_group_unique_id failed: NCCL not available")
    t = torch.from_numpy(buf).to(torch.device("cuda", device))
    dist.broadcast(t, src=0)
    buf = t.cpu().numpy()
    g = L.nsfsh_group_create(buf.ctypes.data_as(C.c_void_p), rank, world, device)
    if not g:
        raise PlannerError("nsfs_group_create failed")
    return g


def destroy_group(group, so_path=None):
    lib(so_path).nsfsh_group_destroy(group)


class Session:
    """main_planner_ + fallback_planner_ over one MapData, as GlobalPlanner::setupPlanner builds them."""

    def __init__(self, planner_name, cost_mode, infl, lo, lo_thresh=10, cell=0.05, origin=(0.0, 0.0), backend="exact",
                 device=0, so_path=None):
        self.L = lib(so_path)
        infl = np.ascontiguousarray(infl, np.int32)
        lo = np.ascontiguousarray(lo, np.int32)
        self.H, self.W = infl.shape
        rc = C.c_int(0)
        self.s = self.L.nsfsh_create(planner_name.encode(), cost_mode.encode(), backend.encode(), self.H, self.W, cell,
                                     origin[0], origin[1], lo_thresh, infl.ctypes.data_as(C.c_void_p),
                                     lo.ctypes.data_as(C.c_void_p), device, C.byref(rc))
        if rc.value != 0:
            msg = self.L.nsfsh_error(self.s).decode()
            self.close()
            raise PlannerError(msg)

    def close(self):
        if getattr(self, "s", None):
            self.L.nsfsh_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def kind(self):
        return ("astar", "djikstra", "thetastar")[self.L.nsfsh_main_planner_kind(self.s)]

    def set_map(self, infl, lo):
        infl = np.ascontiguousarray(infl, np.int32)
        lo = np.ascontiguousarray(lo, np.int32)
        self.L.nsfsh_set_map(self.s, infl.ctypes.data_as(C.c_void_p), lo.ctypes.data_as(C.c_void_p))

    def set_map_sync(self, mode):
        """'every_call' (pageable upload per plan: the reference's contract, default) | 'on_notify' | 'every_call_pinned'."""
        self.L.nsfsh_set_map_sync(self.s, {"every_call": 0, "on_notify": 1, "every_call_pinned": 2}[mode])

    def notify_map_changed(self):
        self.L.nsfsh_notify_map_changed(self.s)

    def plan_batch_idx(self, sg4):
        """nq plan(Index, Index) calls of the main planner on one upload of the session's MapData (GpuPlannerBase::planBatch)."""
        sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
        nq = len(sg4)
        status, plen = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        g, settled = np.zeros(nq, np.float64), np.zeros(nq, np.int64)
        ms = C.c_double(0)
        vp = C.c_void_p
        rc = self.L.nsfsh_plan_batch_idx(self.s, nq, sg4.ctypes.data_as(vp), status.ctypes.data_as(vp), g.ctypes.data_as(vp),
                                         plen.ctypes.data_as(vp), settled.ctypes.data_as(vp), C.byref(ms))
        if rc != 0:
            raise PlannerError(self.L.nsfsh_error(self.s).decode())
        return dict(status=status, g_goal=g, path_len=plen, settled=settled, gpu_ms=ms.value)

    def plan_batch_sharded(self, group, sg4):
        """GpuPlannerBase::planBatchSharded over a DeviceGroup (see ``make_group``): all queries in, all results out on every rank."""
        sg4 = np.ascontiguousarray(sg4, np.int32).reshape(-1, 4)
        nq = len(sg4)
        status, plen = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        g, settled = np.zeros(nq, np.float64), np.zeros(nq, np.int64)
        ms = C.c_double(0)
        vp = C.c_void_p
        rc = self.L.nsfsh_plan_batch_sharded(self.s, group, nq, sg4.ctypes.data_as(vp), status.ctypes.data_as(vp), g.ctypes.data_as(vp),
                                             plen.ctypes.data_as(vp), settled.ctypes.data_as(vp), C.byref(ms))
        if rc != 0:
            raise PlannerError(self.L.nsfsh_error(self.s).decode())
        return dict(status=status, g_goal=g, path_len=plen, settled=settled, gpu_ms=ms.value)

    def _path(self, fn, *args):
        cap = 4 * (self.H + self.W) + 64
        while True:
            out = np.zeros((cap, 2), np.float64)
            n = C.c_int(0)
            rc = fn(self.s, *args, out.ctypes.data_as(C.c_void_p), cap, C.byref(n))
            if rc == 1:
                return dict(status="out_of_range", path=None)
            if rc != 0:
                raise PlannerError(self.L.nsfsh_error(self.s).decode())
            if n.value <= cap:
                return dict(status="ok", path=out[: n.value].copy())
            cap = n.value

    def generate_path_idx(self, start, goal):
        return self._path(self.L.nsfsh_generate_path_idx, int(start[0]), int(start[1]), int(goal[0]), int(goal[1]))

    def generate_path_pos(self, start_xy, goal_xy):
        return self._path(self.L.nsfsh_generate_path_pos, float(start_xy[0]), float(start_xy[1]), float(goal_xy[0]), float(goal_xy[1]))

    def plan_request(self, robot_xy, goal_xy):
        """GlobalPlanner::run's plan request (global_planner.cpp:164-273): test both cells, repair with the fallback, plan."""
        cap = 4 * (self.H + self.W) + 64
        while True:
            out = np.zeros((cap, 2), np.float64)
            n, info = C.c_int(0), np.zeros(4, np.int32)
            goal, backup = np.zeros(2, np.float64), np.zeros(2, np.float64)
            rc = self.L.nsfsh_plan_request(self.s, float(robot_xy[0]), float(robot_xy[1]), float(goal_xy[0]), float(goal_xy[1]),
                                           out.ctypes.data_as(C.c_void_p), cap, C.byref(n), info.ctypes.data_as(C.c_void_p),
                                           goal.ctypes.data_as(C.c_void_p), backup.ctypes.data_as(C.c_void_p))
            if rc == 1:
                return dict(status="out_of_range")
            if rc != 0:
                raise PlannerError(self.L.nsfsh_error(self.s).decode())
            if n.value <= cap:
                return dict(status="ok", path=out[: n.value].copy(), case=int(info[0]), goal_updated=bool(info[1]),
                            backup_mode=bool(info[2]), robot_ok=bool(info[3] & 1), goal_ok=bool(info[3] & 2),
                            goal=tuple(goal), backup_robot=tuple(backup))
            cap = n.value

    def set_graph(self, textbook):
        """Opt-in textbook 8-connected grid for the main planner (False = the reference's own graph)."""
        if self.L.nsfsh_set_graph(self.s, 1 if textbook else 0) != 0:
            raise PlannerError(self.L.nsfsh_error(self.s).decode())

    def check_trajectory(self, xy, t_id):
        """Commander::checkTrajectorySafety (commander.cpp:316-357) on world-metre trajectory points against the session's map."""
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        safe, bad = C.c_int(0), C.c_int(-1)
        rc = self.L.nsfsh_check_trajectory(self.s, len(xy), xy.ctypes.data_as(C.c_void_p), int(t_id), C.byref(safe), C.byref(bad))
        if rc == 1:
            return dict(status="out_of_range", safe=None, bad_idx=None)
        if rc != 0:
            raise PlannerError(self.L.nsfsh_error(self.s).decode())
        return dict(status="ok", safe=bool(safe.value), bad_idx=int(bad.value))

    def find_better_point(self, bad, as_pos=False):
        fx, fy = C.c_double(0), C.c_double(0)
        if as_pos:
            rc = self.L.nsfsh_find_better_point_pos(self.s, float(bad[0]), float(bad[1]), C.byref(fx), C.byref(fy))
        else:
            rc = self.L.nsfsh_find_better_point_idx(self.s, int(bad[0]), int(bad[1]), C.byref(fx), C.byref(fy))
        if rc == 1:
            return dict(status="out_of_range", pos=None)
        if rc != 0:
            raise PlannerError(self.L.nsfsh_error(self.s).decode())
        return dict(status="ok", pos=(fx.value, fy.value))
