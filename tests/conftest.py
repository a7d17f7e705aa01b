import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import ctypes
        return ctypes.CDLL("libcuda.so.1").cuInit(0) == 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of dying in dp_create"""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def the_map():
    from dmpp_b200 import scenes
    return scenes.Map()


@pytest.fixture(scope="session")
def oracle(the_map):
    from oracle import binding
    o = binding.Oracle()
    o.set_map(the_map)
    return o


@pytest.fixture(scope="session")
def reference(the_map):
    from oracle import binding
    if not binding.Reference.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    r = binding.Reference()
    r.set_map(the_map)
    return r


def same(a, b):
    """bit-for-bit equality, NaN == NaN"""
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return (a == b) | ((a != a) & (b != b))
    return a == b


def assert_records_equal(got, want, fields, mask=None, close=None, what=""):
    """bit-exact compare of structured-array fields; `close` maps field -> abs tolerance."""
    close = close or {}
    for f in fields:
        x, y = got[f], want[f]
        ok = same(x, y)
        if f in close:
            ok = ok | (np.abs(x - y) <= close[f])
        if mask is not None:
            ok = ok | ~mask
        if not ok.all():
            idx = np.argwhere(~ok)
            raise AssertionError("%s field %s: %d mismatches, first at %s: got %r want %r" % (
                what, f, len(idx), tuple(idx[0]), x[tuple(idx[0])], y[tuple(idx[0])]))
