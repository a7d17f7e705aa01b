"""the C-ABI boundary: struct layouts agree between include/dmpp_b200.h, the numpy mirrors and both
oracle libraries; libdmpp_b200.so loads WITHOUT a GPU and exports every symbol the header declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dmpp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from dmpp_b200 import planner
    assert header_symbols() == sorted(planner.SYMBOLS)


def test_library_exports_every_symbol():
    from dmpp_b200 import planner
    lib = planner.load()                       # dlopen only: no CUDA call
    for s in header_symbols():
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", planner.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dp_[a-z0-9_]+)", nm))
    assert exported == set(header_symbols())


def test_struct_sizes(oracle):
    from dmpp_b200 import abi
    want = [abi.scene_hdr, abi.plan_record, abi.carry, abi.ref_call, abi.trace_record]
    for i, dt in enumerate(want):
        assert oracle.lib.oracle_sizeof(i) == dt.itemsize
    assert oracle.lib.oracle_sizeof(5) == C.sizeof(abi.Params)
    assert oracle.lib.oracle_sizeof(6) == C.sizeof(abi.MapDesc)


def test_default_params_match_oracle(oracle):
    from dmpp_b200 import abi, planner
    p = planner.default_params()
    for name, _ in abi.Params._fields_:
        assert getattr(p, name) == getattr(oracle.params, name), name


def test_no_gpu_means_loud_failure():
    """the product has no CPU path: without a usable device dp_create must fail, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dmpp_b200.planner import DpError, Planner
    with pytest.raises(DpError):
        Planner(16, 10)
