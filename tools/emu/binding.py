"""tools/emu/binding.py -- TEST INFRASTRUCTURE: ctypes loader of the host emulation of the group kernel
(tools/emu/libdp_emu.so).  Imported by tests/ only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(os.path.dirname(_here))
if _root not in sys.path:
    sys.path.insert(0, _root)
import dmpp_b200  # noqa: E402,F401
from dmpp_b200 import abi  # noqa: E402


def build():
    subprocess.check_call(["make", "-s", "-C", _here, "all"])


class Emu:
    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(_here, "libdp_emu.so"))
        self.params = abi.Params()
        from oracle import binding as ob
        ob.Oracle().lib.oracle_default_params(C.byref(self.params))

    def set_map(self, m):
        self._keep = m
        self._desc = m.desc()
        assert self.lib.emu_set_map(C.byref(self._desc)) == 0

    def run_tracks(self, H, OX, OY, VX, VY, DTH, T, trace=True, paths=True, group=14):
        cycles, n = H.shape
        max_obs = OX.shape[2]
        o = {"rec": np.zeros((cycles, n), abi.plan_record),
             "trace": np.zeros((cycles, n), abi.trace_record) if trace else None,
             "path_xy": np.zeros((cycles, n, 2, abi.PATH_POINTS)) if paths else None,
             "path_ll": np.zeros((cycles, n, 2, abi.OUT_POINTS)) if paths else None,
             "carry": np.zeros(n, abi.carry), "last_path": np.zeros((n, 2, abi.PATH_POINTS))}
        rc = self.lib.emu_run_batch_tracks(C.byref(self.params), C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(H), abi.ptr(OX),
                                           abi.ptr(OY), abi.ptr(VX), abi.ptr(VY), abi.ptr(DTH), C.c_int(T), abi.ptr(o["rec"]),
                                           abi.ptr(o["trace"]), abi.ptr(o["path_xy"]), abi.ptr(o["path_ll"]), abi.ptr(o["carry"]),
                                           abi.ptr(o["last_path"]), C.c_int(group))
        assert rc == 0, rc
        return o

    def run(self, H, OX, OY, trace=True, paths=True, group=14):
        cycles, n = H.shape
        max_obs = OX.shape[2]
        o = {"rec": np.zeros((cycles, n), abi.plan_record),
             "trace": np.zeros((cycles, n), abi.trace_record) if trace else None,
             "path_xy": np.zeros((cycles, n, 2, abi.PATH_POINTS)) if paths else None,
             "path_ll": np.zeros((cycles, n, 2, abi.OUT_POINTS)) if paths else None,
             "carry": np.zeros(n, abi.carry), "last_path": np.zeros((n, 2, abi.PATH_POINTS))}
        rc = self.lib.emu_run_batch(C.byref(self.params), C.c_int(n), C.c_int(cycles), C.c_int(max_obs), abi.ptr(H), abi.ptr(OX),
                                    abi.ptr(OY), abi.ptr(o["rec"]), abi.ptr(o["trace"]), abi.ptr(o["path_xy"]), abi.ptr(o["path_ll"]),
                                    abi.ptr(o["carry"]), abi.ptr(o["last_path"]), C.c_int(group))
        assert rc == 0, rc
        return o
