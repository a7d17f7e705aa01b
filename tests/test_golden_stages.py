"""Golden vectors of the stages either side of the cycle (tests/golden/stages/, frozen from the UNMODIFIED reference by
tests/golden/make_golden_stages.py): closed-loop episodes, the published frames, the V2X flags.  The CPU oracle is checked
against them here without the reference being present; the CUDA path is checked against the same files (marked gpu)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_records_equal, same
from test_closed_loop import HDR
from test_oracle_vs_ref import REC

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden_stages as G  # noqa: E402

DIR = os.path.join(ROOT, "tests", "golden", "stages")
FLAGS = ["light_flag", "construction_flag", "pedestrian_flag"]


def rows(a):
    """frames as rows of bytes (a NaN in a frame must compare equal to itself)"""
    a = np.ascontiguousarray(a)
    return a.view("u1").reshape(a.shape + (-1,))


def golden(name, *inputs):
    g = np.load(os.path.join(DIR, name + ".npz"))
    assert np.array_equal(G.checksum(*inputs), g["inputs"]), "scene generator drifted: regenerate the fixtures deliberately"
    return g


def check_closed_loop(a, g, close=None):
    clean = g["ub_scene"] == 0
    assert_records_equal(a["rec"][:, clean], g["rec"][:, clean], REC, close=close, what="record")
    assert_records_equal(a["hdr_log"][:, clean], g["hdr_log"][:, clean], HDR, what="logged header")
    assert same(a["obs_log_x"][:, clean], g["obs_log_x"][:, clean]).all() and same(a["obs_log_y"][:, clean], g["obs_log_y"][:, clean]).all()
    assert_records_equal(a["hdr"][clean], g["hdr"][clean], HDR, what="final header")
    assert a["agents"][clean].tobytes() == g["agents"][clean].tobytes()
    assert same(a["last_path"][clean], g["last_path"][clean]).all()


def test_closed_loop_oracle_matches_golden(oracle, the_map):
    w, cycles = G.closed_loop_inputs(the_map)
    check_closed_loop(oracle.run_closed_loop(w.hdr, w.agents, cycles, threads=4), golden("closed_loop_24", w.hdr, w.agents))


@pytest.mark.parametrize("name", sorted(G.FRAMES))
def test_frames_oracle_matches_golden(name, oracle, the_map):
    H, OX, OY = G.frames_inputs(the_map, name)
    g = golden(name, H, OX, OY)
    a = oracle.run(H, OX, OY, exhaustive=False, threads=4)
    ctrl, status = oracle.pack_frames(a["rec"].reshape(-1), a["path_xy"].reshape(-1, 2, 200))
    clean = g["clean"]
    assert (rows(ctrl.reshape(clean.shape))[clean] == rows(g["ctrl"])[clean]).all()
    assert (rows(status.reshape(clean.shape))[clean] == rows(g["status"])[clean]).all()


def test_v2x_oracle_matches_golden(oracle, the_map):
    h, v, wl, wg = G.v2x_inputs(the_map)
    g = golden("v2x_2048", h, v, wl, wg)
    for mode in (0, 1):
        a = oracle.v2x_event(h, v, wl, wg, mode)
        ok = g["ub%d" % mode] == 0
        assert np.array_equal(a["ub"], g["ub%d" % mode])
        for i, f in enumerate(FLAGS):
            assert np.array_equal(a[f][ok], g["mode%d" % mode][i][ok]), (mode, f)


@pytest.mark.gpu
def test_cuda_matches_stage_goldens(the_map):
    from dmpp_b200.planner import Planner
    w, cycles = G.closed_loop_inputs(the_map)
    p = Planner(max_scenes=2048, max_obs=10)
    p.upload_map(the_map)
    # (path_dir_err: the reference calls libm atan there, the device evaluates the specification's polynomial -- 1e-9 degrees, as everywhere)
    check_closed_loop(p.run_closed_loop(w.hdr, w.agents, cycles), golden("closed_loop_24", w.hdr, w.agents), close={"path_dir_err": 1e-9})
    for name in sorted(G.FRAMES):
        H, OX, OY = G.frames_inputs(the_map, name)
        g = golden(name, H, OX, OY)
        p.reset(0, H.shape[1])
        for c in range(H.shape[0]):
            o = p.cycle(np.ascontiguousarray(H[c]), np.ascontiguousarray(OX[c]), np.ascontiguousarray(OY[c]))
            ctrl, status = p.pack_frames(o["rec"])
            k = g["clean"][c]
            assert (rows(ctrl)[k] == rows(g["ctrl"][c])[k]).all() and (rows(status)[k] == rows(g["status"][c])[k]).all(), (name, c)
    h, v, wl, wg = G.v2x_inputs(the_map)
    g = golden("v2x_2048", h, v, wl, wg)
    for mode in (0, 1):
        a = p.v2x_event(h, v, wl, wg, mode)
        ok = g["ub%d" % mode] == 0
        for i, f in enumerate(FLAGS):
            assert np.array_equal(a[f][ok], g["mode%d" % mode][i][ok]), (mode, f)
    p.close()
