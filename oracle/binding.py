"""oracle/binding.py -- TEST INFRASTRUCTURE. ctypes loaders for liboracle.so (the restated CPU
oracle) and _ref/libref.so (the unmodified reference + shim).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference leg may import this module."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
if _root not in sys.path:
    sys.path.insert(0, _root)
import dmpp_b200  # noqa: E402
from dmpp_b200 import abi  # noqa: E402

CALLS_CAP = 32


def build(ref_dir="/root/reference"):
    """compile liboracle.so and, when the reference sources are present, _ref/libref.so."""
    subprocess.check_call(["make", "-s", "-C", _here, "REF=" + ref_dir, "all"])


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


class _Runner:
    def __init__(self):
        self._keep = None

    @staticmethod
    def _alloc(n, cycles, want_paths, want_calls, want_trace):
        out = {"rec": np.zeros((cycles, n), abi.plan_record)}
        out["trace"] = np.zeros((cycles, n), abi.trace_record) if want_trace else None
        out["path_xy"] = np.zeros((cycles, n, 2, abi.PATH_POINTS)) if want_paths else None
        out["path_ll"] = np.zeros((cycles, n, 2, abi.OUT_POINTS)) if want_paths else None
        out["calls"] = np.zeros((cycles, n, CALLS_CAP), abi.ref_call) if want_calls else None
        out["n_calls"] = np.zeros((cycles, n), np.int32) if want_calls else None
        out["carry"] = np.zeros(n, abi.carry)
        out["last_path"] = np.zeros((n, 2, abi.PATH_POINTS))
        return out


def _closed_loop(lib, fn, params, wp, hdr, agents, cycles, paths, threads):
    n, max_obs = agents.shape
    o = {"hdr": hdr.copy(), "agents": agents.copy(), "ox": np.zeros((n, max_obs)), "oy": np.zeros((n, max_obs)),
         "rec": np.zeros((cycles, n), abi.plan_record), "hdr_log": np.zeros((cycles, n), abi.scene_hdr),
         "obs_log_x": np.zeros((cycles, n, max_obs)), "obs_log_y": np.zeros((cycles, n, max_obs)),
         "path_xy": np.zeros((cycles, n, 2, abi.PATH_POINTS)) if paths else None,
         "carry": np.zeros(n, abi.carry), "last_path": np.zeros((n, 2, abi.PATH_POINTS))}
    args = [C.byref(params), C.byref(wp), C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(o["hdr"]), abi.ptr(o["agents"]),
            abi.ptr(o["ox"]), abi.ptr(o["oy"]), abi.ptr(o["rec"]), abi.ptr(o["hdr_log"]), abi.ptr(o["obs_log_x"]), abi.ptr(o["obs_log_y"]),
            abi.ptr(o["path_xy"]), abi.ptr(o["carry"]), abi.ptr(o["last_path"])]
    if threads is not None:
        o["ub_scene"] = np.zeros(n, np.int32)
        args += [C.c_int(threads), abi.ptr(o["ub_scene"])]
    f = getattr(lib, fn)
    f.restype = C.c_longlong if threads is not None else C.c_int
    rc = f(*args)
    if rc < 0:
        raise RuntimeError("%s failed: %d" % (fn, rc))
    o["status"] = rc
    return o


def _v2x(fn, head, hdr, v2x, wp_lat, wp_lng, mode):
    n = hdr.shape[0]
    out = np.zeros(n, abi.v2x_flags)
    wl, wg = np.ascontiguousarray(wp_lat, np.float64), np.ascontiguousarray(wp_lng, np.float64)
    rc = fn(*head, C.c_int(n), abi.ptr(np.ascontiguousarray(hdr)), abi.ptr(np.ascontiguousarray(v2x)), abi.ptr(wl), abi.ptr(wg),
            C.c_int(mode), abi.ptr(out))
    if rc != 0:
        raise RuntimeError("v2x_event failed: %d" % rc)
    return out


class Oracle(_Runner):
    """the restated oracle (liboracle.so): re-entrant, multi-threaded."""

    def __init__(self):
        super().__init__()
        self.lib = _load(os.path.join(_here, "liboracle.so"))
        if self.lib is None:
            raise RuntimeError("oracle/liboracle.so missing: run __graft_entry__.build()")
        self.lib.oracle_run_batch.restype = C.c_longlong
        self.lib.oracle_run_batch_tracks.restype = C.c_longlong
        self.lib.oracle_atan.restype = C.c_double
        self.lib.oracle_atan.argtypes = [C.c_double]
        self.lib.oracle_calc_global_dir.restype = C.c_double
        self.lib.oracle_calc_global_dir.argtypes = [C.c_double] * 4
        self.lib.oracle_lat_dis.restype = C.c_double
        self.lib.oracle_lat_dis.argtypes = [C.c_double] * 6
        self.lib.oracle_calc_distance.restype = C.c_double
        self.lib.oracle_calc_distance.argtypes = [C.c_double] * 4
        self.params = abi.Params()
        self.lib.oracle_default_params(C.byref(self.params))

    def set_map(self, m):
        self._keep = m
        self._desc = m.desc()
        assert self.lib.oracle_set_map(C.byref(self._desc)) == 0

    def run(self, H, OX, OY, paths=True, calls=True, trace=True, exhaustive=True, threads=1):
        cycles, n = H.shape
        max_obs = OX.shape[2]
        o = self._alloc(n, cycles, paths, calls, trace)
        sec = C.c_double(0)
        traj = C.c_longlong(0)
        ub = self.lib.oracle_run_batch(
            C.byref(self.params), C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(H), abi.ptr(OX), abi.ptr(OY),
            abi.ptr(o["rec"]), abi.ptr(o["trace"]), abi.ptr(o["path_xy"]), abi.ptr(o["path_ll"]), abi.ptr(o["calls"]),
            abi.ptr(o["n_calls"]), C.c_int(CALLS_CAP), abi.ptr(o["carry"]), abi.ptr(o["last_path"]),
            C.c_int(1 if exhaustive else 0), C.c_int(threads), C.byref(sec), C.byref(traj))
        if ub < 0:
            raise RuntimeError("oracle_run_batch failed: %d" % ub)
        o["ub_hits"] = ub
        o["seconds"] = sec.value
        o["traj"] = traj.value
        return o

    def run_tracks(self, H, OX, OY, VX, VY, DTH, T, paths=True, trace=True, threads=1):
        """episodes with predicted agent tracks (BASELINE config 5): constant-turn-rate parameters per cycle / scene / agent"""
        cycles, n = H.shape
        max_obs = OX.shape[2]
        o = self._alloc(n, cycles, paths, False, trace)
        sec = C.c_double(0)
        ub = self.lib.oracle_run_batch_tracks(
            C.byref(self.params), C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(H), abi.ptr(OX), abi.ptr(OY), abi.ptr(VX),
            abi.ptr(VY), abi.ptr(DTH), C.c_int(T), abi.ptr(o["rec"]), abi.ptr(o["trace"]), abi.ptr(o["path_xy"]), abi.ptr(o["path_ll"]),
            abi.ptr(o["carry"]), abi.ptr(o["last_path"]), C.c_int(threads), C.byref(sec))
        if ub < 0:
            raise RuntimeError("oracle_run_batch_tracks failed: %d" % ub)
        o["ub_hits"] = ub
        o["seconds"] = sec.value
        return o

    # ---- closed-loop episodes and output frames (oracle/world_spec.h) ----
    def world_params(self):
        wp = abi.WorldParams()
        self.lib.oracle_world_default_params(C.byref(wp))
        return wp

    def world_step(self, hdr, agents, ox, oy, rec=None, last_path=None, wp=None):
        """in place; rec None: place the agents and localise only"""
        wp = wp or self.world_params()
        n, max_obs = ox.shape
        assert self.lib.oracle_world_step(C.byref(self.params), C.byref(wp), C.c_int(n), C.c_int(max_obs), abi.ptr(hdr), abi.ptr(agents),
                                          abi.ptr(ox), abi.ptr(oy), abi.ptr(rec), abi.ptr(last_path)) == 0

    def run_closed_loop(self, hdr, agents, cycles, wp=None, paths=False, threads=1):
        """hdr[n], agents[n][max_obs] (copied; the final world is returned)"""
        return _closed_loop(self.lib, "oracle_run_closed_loop", self.params, wp or self.world_params(), hdr, agents, cycles, paths, threads)

    def pack_frames(self, rec, path_xy):
        n = rec.shape[0]
        ctrl, status = np.zeros(n, abi.ctrl_frame), np.zeros(n, abi.status_frame)
        self.lib.oracle_pack_frames(C.byref(self.params), C.c_int(n), abi.ptr(np.ascontiguousarray(rec)),
                                    abi.ptr(np.ascontiguousarray(path_xy)), abi.ptr(ctrl), abi.ptr(status))
        return ctrl, status

    def v2x_event(self, hdr, v2x, wp_lat, wp_lng, mode=0):
        return _v2x(self.lib.oracle_v2x_event, [C.byref(self.params)], hdr, v2x, wp_lat, wp_lng, mode)

    def v2x_apply(self, flags, rec):
        out = np.ascontiguousarray(rec).copy()
        self.lib.oracle_v2x_apply(C.c_int(out.shape[0]), abi.ptr(np.ascontiguousarray(flags)), abi.ptr(out))
        return out

    def rollout_ctr(self, x0, y0, vx, vy, dth, T):
        ox, oy = np.zeros(T), np.zeros(T)
        self.lib.oracle_rollout_ctr.argtypes = [C.c_double] * 5 + [C.c_int, C.c_void_p, C.c_void_p]
        self.lib.oracle_rollout_ctr(x0, y0, vx, vy, dth, T, abi.ptr(ox), abi.ptr(oy))
        return ox, oy

    BRANCHES = ["b3_nav_1108", "b3_obs_1382", "b3_both_1618", "b3_both_1711", "enter_1596", "enter_1688", "aim_right_473",
                "aim_right_walk"]

    def branch_hits(self, reset=True):
        """hit counters of the right-lane-change sites since the last reset (planner_oracle.h BR_*)"""
        out = np.zeros(len(self.BRANCHES), np.int64)
        self.lib.oracle_branch_hits(abi.ptr(out), C.c_int(1 if reset else 0))
        return dict(zip(self.BRANCHES, out.tolist()))

    # operator-level
    def search_obstacle(self, px, py, ox, oy, lo, hi):
        px, py, ox, oy = (np.ascontiguousarray(a, np.float64) for a in (px, py, ox, oy))
        out = np.zeros(1, abi.search_slot)
        self.lib.oracle_search_obstacle(abi.ptr(px), abi.ptr(py), C.c_int(px.size), abi.ptr(ox), abi.ptr(oy),
                                        C.c_int(ox.size), C.c_double(lo), C.c_double(hi), abi.ptr(out))
        return out[0]

    def create_new_path(self, px, py, d):
        px, py = (np.ascontiguousarray(a, np.float64) for a in (px, py))
        ox, oy = np.zeros_like(px), np.zeros_like(py)
        self.lib.oracle_create_new_path(abi.ptr(px), abi.ptr(py), C.c_int(px.size), C.c_double(d), abi.ptr(ox), abi.ptr(oy))
        return ox, oy

    def bezier(self, poses6, n=abi.PATH_POINTS):
        poses6 = np.ascontiguousarray(poses6, np.float64)
        out = np.zeros((2, n))
        self.lib.oracle_bezier(abi.ptr(poses6), abi.ptr(out), C.c_int(n))
        return out

    def nearest_id(self, qx, qy, px, py):
        px, py = (np.ascontiguousarray(a, np.float64) for a in (px, py))
        return int(self.lib.oracle_nearest_id(C.c_double(qx), C.c_double(qy), abi.ptr(px), abi.ptr(py), C.c_int(px.size)))

    def mean_points(self, px, py, n_out=abi.PATH_POINTS):
        px, py = (np.ascontiguousarray(a, np.float64) for a in (px, py))
        out = np.zeros((2, n_out))
        self.lib.oracle_mean_points(abi.ptr(px), abi.ptr(py), C.c_int(px.size), abi.ptr(out), C.c_int(n_out))
        return out

    def score_candidates(self, base_x, base_y, offset, n_pts, ox, oy, dvx=None, dvy=None, lat_min=-0.9, lat_max=0.9,
                         clear_dis=25.0):
        bx, by = np.ascontiguousarray(base_x, np.float64), np.ascontiguousarray(base_y, np.float64)
        off = np.ascontiguousarray(offset, np.float64)
        npt = np.ascontiguousarray(n_pts, np.int32)
        ox, oy = np.ascontiguousarray(ox, np.float64), np.ascontiguousarray(oy, np.float64)
        dvx = None if dvx is None else np.ascontiguousarray(dvx, np.float64)
        dvy = None if dvy is None else np.ascontiguousarray(dvy, np.float64)
        out = np.zeros(off.size)
        best = self.lib.oracle_score_candidates(abi.ptr(bx), abi.ptr(by), C.c_int(bx.size), abi.ptr(off), abi.ptr(npt),
                                                C.c_int(off.size), abi.ptr(ox), abi.ptr(oy), abi.ptr(dvx), abi.ptr(dvy),
                                                C.c_int(ox.size), C.c_double(lat_min), C.c_double(lat_max),
                                                C.c_double(clear_dis), abi.ptr(out))
        return best, out

    def sincos_deg(self, a):
        c, s = C.c_double(0), C.c_double(0)
        self.lib.oracle_sincos_deg(C.c_double(a), C.byref(c), C.byref(s))
        return c.value, s.value


class Reference(_Runner):
    """the UNMODIFIED reference (oracle/_ref/libref.so); single-threaded by construction."""

    def __init__(self, gpu_share=False):
        """gpu_share=True loads _ref/libref_gpu.so: the same unmodified reference objects linked against the PRODUCT's
        drop-in Share.h (host/), i.e. every CShare operator call is a CUDA launch in libdmpp_b200.so."""
        super().__init__()
        name = "libref_gpu.so" if gpu_share else "libref.so"
        self.lib = _load(os.path.join(_here, "_ref", name))
        if self.lib is None:
            raise RuntimeError("oracle/_ref/%s missing (built by __graft_entry__.build() where /root/reference exists)" % name)
        self.lib.ref_run_batch.restype = C.c_longlong

    @staticmethod
    def available(gpu_share=False):
        return os.path.exists(os.path.join(_here, "_ref", "libref_gpu.so" if gpu_share else "libref.so"))

    def set_map(self, m):
        self._keep = m
        d = m.desc()
        assert self.lib.ref_set_map(C.byref(d)) == 0

    def run(self, H, OX, OY, paths=True, calls=True):
        cycles, n = H.shape
        max_obs = OX.shape[2]
        o = self._alloc(n, cycles, paths, calls, False)
        sec = C.c_double(0)
        traj = C.c_longlong(0)
        rc = self.lib.ref_run_batch(
            C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(H), abi.ptr(OX), abi.ptr(OY), abi.ptr(o["rec"]),
            abi.ptr(o["path_xy"]), abi.ptr(o["path_ll"]), abi.ptr(o["calls"]), abi.ptr(o["n_calls"]), C.c_int(CALLS_CAP),
            abi.ptr(o["carry"]), abi.ptr(o["last_path"]), C.byref(sec), C.byref(traj))
        if rc < 0:
            raise RuntimeError("ref_run_batch failed: %d" % rc)
        o["msgbox"] = rc
        o["seconds"] = sec.value
        o["traj"] = traj.value
        return o

    def run_with_frames(self, H, OX, OY):
        """run() that also captures what the Planning thread hands to SetUdpSendCtrl / SetPlanningStatus every cycle"""
        cycles, n = H.shape
        ctrl, status = np.zeros((cycles, n), abi.ctrl_frame), np.zeros((cycles, n), abi.status_frame)
        self.lib.ref_capture_frames(abi.ptr(ctrl), abi.ptr(status))
        try:
            o = self.run(H, OX, OY, paths=True, calls=False)
        finally:
            self.lib.ref_capture_frames(None, None)
        o["ctrl"], o["status"] = ctrl, status
        return o

    def v2x_event(self, hdr, v2x, wp_lat, wp_lng, mode=0):
        """the reference's own V2XEventDecision / V2XConstructionEventTemporal, one call per scene (only the flags are observable)"""
        return _v2x(self.lib.ref_v2x_event, [], hdr, v2x, wp_lat, wp_lng, mode)

    def run_closed_loop(self, params, wp, hdr, agents, cycles, paths=False):
        """closed-loop episodes of the unmodified reference with the world step of oracle/world_spec.cpp between its cycles"""
        return _closed_loop(self.lib, "ref_run_closed_loop", params, wp, hdr, agents, cycles, paths, None)
