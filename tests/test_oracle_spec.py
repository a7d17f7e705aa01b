"""Known-answer tests of the operator specification (oracle/cshare_spec.cpp): the reference ships no
tests for these, so the analytic cases below are what pins the frozen definitions."""
import math

import numpy as np
import pytest


def test_sincos_atan_accuracy(oracle):
    worst = 0.0
    pi_ld = np.longdouble("3.14159265358979323846264338327950288")
    for a in np.linspace(-720, 1080, 3601):
        c, s = oracle.sincos_deg(float(a))
        x = np.longdouble(a) * pi_ld / np.longdouble(180)          # 80-bit reference: the degree reduction itself is exact
        worst = max(worst, abs(float(np.longdouble(c) - np.cos(x))), abs(float(np.longdouble(s) - np.sin(x))))
    assert worst < 4e-16
    assert oracle.sincos_deg(90.0) == (0.0, 1.0) and oracle.sincos_deg(180.0) == (-1.0, 0.0) and oracle.sincos_deg(270.0) == (0.0, -1.0)
    for z in np.concatenate([np.linspace(-50, 50, 2001), [1e-12, 1e9, -1e9, 0.41421356237309503, 1.0]]):
        assert abs(oracle.lib.oracle_atan(float(z)) - math.atan(z)) < 5e-16


def test_heading_convention(oracle):
    """degrees, 0 = east, CCW, [0,360)  (Planning.cpp:712-750)"""
    g = oracle.lib.oracle_calc_global_dir
    assert g(0, 0, 1, 0) == 0.0
    assert g(0, 0, 0, 1) == 90.0
    assert abs(g(0, 0, -1, 0) - 180.0) < 1e-12
    assert g(0, 0, 0, -1) == 270.0
    assert abs(g(0, 0, 1, 1) - 45.0) < 1e-12 and abs(g(0, 0, 1, -1) - 315.0) < 1e-12
    assert g(3, 4, 3, 4) == 0.0


def test_lat_dis_left_positive(oracle):
    f = oracle.lib.oracle_lat_dis
    assert f(0.0, 2.0, 0, 0, 10, 0) == 2.0          # left of an east-bound segment
    assert f(0.0, -2.0, 0, 0, 10, 0) == -2.0
    assert f(1.0, 5.0, 0, 0, 0, 10) == -1.0         # vertical segment branch: right of a north-bound segment
    assert f(5.0, 1e-9, 0, 0, 10, 0) == 0.0         # below EPSILON


def straight(n=120, spacing=0.5, heading=0.0, x0=0.0, y0=0.0):
    s = np.arange(n) * spacing
    return x0 + s * math.cos(math.radians(heading)), y0 + s * math.sin(math.radians(heading))


def test_search_obstacle_straight_path(oracle):
    px, py = straight()
    # obstacle 20 m ahead, 0.4 m to the RIGHT (south of an east-bound path) -> d = +0.4, s = 20
    r = oracle.search_obstacle(px, py, [20.0], [-0.4], -0.9, 0.9)
    assert r["found"] and r["pathid"] == 40 and r["ob_index"] == 0
    assert r["dis_lng"] == 20.0 and abs(r["dis_lat"] - 0.4) < 1e-15
    # left of the path -> negative; outside the corridor -> not found with the 999 sentinel
    assert oracle.search_obstacle(px, py, [20.0], [0.4], -0.9, 0.9)["dis_lat"] < 0
    r = oracle.search_obstacle(px, py, [20.0], [-1.0], -0.9, 0.9)
    assert not r["found"] and r["dis_lng"] == 999.0 and r["dis_lat"] == 999.0 and r["pathid"] == 0 and r["ob_index"] == -1
    # nearest along the path wins, not nearest in space; ties on the path point go to the lower obstacle index
    r = oracle.search_obstacle(px, py, [30.0, 10.0, 10.0], [0.0, 0.5, -0.5], -0.9, 0.9)
    assert r["ob_index"] == 1 and r["pathid"] == 20 and r["dis_lng"] == 10.0
    # window edges are inclusive
    assert oracle.search_obstacle(px, py, [5.0], [-0.9], -0.9, 0.9)["found"]
    # obstacles behind the first point / beyond the last point are not on the path
    assert not oracle.search_obstacle(px, py, [-3.0], [0.0], -0.9, 0.9)["found"]
    assert not oracle.search_obstacle(px, py, [70.0], [0.0], -0.9, 0.9)["found"]
    assert oracle.search_obstacle(px, py, [0.1], [0.0], -0.9, 0.9)["dis_lng"] == 0.0
    # degenerate inputs
    assert not oracle.search_obstacle(px[:1], py[:1], [0.0], [0.0], -1, 1)["found"]
    assert not oracle.search_obstacle(px, py, [], [], -1, 1)["found"]


def test_search_obstacle_quarter_circle(oracle):
    R, n = 50.0, 158
    a = np.arange(n) * 0.01
    px, py = R * np.sin(a), R * (1 - np.cos(a))     # starts east-bound, turns LEFT
    k = 100
    # obstacle 1 m outside the arc at point k = to the RIGHT of travel -> d ~ +1
    ox, oy = (R + 1.0) * math.sin(a[k]), R - (R + 1.0) * math.cos(a[k])
    r = oracle.search_obstacle(px, py, [ox], [oy], -2.0, 2.0)
    assert r["found"] and r["pathid"] == k
    assert abs(r["dis_lat"] - 1.0) < 2e-3
    chord = 2 * R * math.sin(0.005)
    assert abs(r["dis_lng"] - k * chord) < 1e-9


def test_create_new_path(oracle):
    px, py = straight(n=5, spacing=1.0)
    ox, oy = oracle.create_new_path(px, py, 0.3)            # +d = right of east-bound = south
    assert np.array_equal(ox, px) and np.allclose(oy, -0.3, atol=0) and np.all(oy == -0.3)
    ox, oy = oracle.create_new_path(px, py, -3.75)          # Decision.cpp:629: -W is the LEFT lane
    assert np.all(oy == 3.75)
    qx, qy = straight(n=4, spacing=2.0, heading=90.0)       # north-bound: right = east
    ox, oy = oracle.create_new_path(qx, qy, 1.0)
    assert np.allclose(ox, 1.0, atol=1e-15) and np.allclose(oy, qy, atol=1e-15)
    # reversed traversal flips the side (the reference's rear paths run backwards, Decision.cpp:590)
    ox, oy = oracle.create_new_path(px[::-1].copy(), py[::-1].copy(), 0.3)
    assert np.all(oy == 0.3)
    # last point reuses the previous segment; single point is copied
    ox, oy = oracle.create_new_path(np.array([0.0, 1.0, 1.0]), np.array([0.0, 0.0, 1.0]), 1.0)
    assert (ox[2], oy[2]) == (2.0, 1.0) and (ox[1], oy[1]) == (2.0, 0.0) and (ox[0], oy[0]) == (0.0, -1.0)
    ox, oy = oracle.create_new_path([7.0], [8.0], 5.0)
    assert (ox[0], oy[0]) == (7.0, 8.0)


def test_bezier_endpoints_and_tangents(oracle):
    out = oracle.bezier([0.0, 0.0, 0.0, 30.0, 10.0, 45.0])
    assert out.shape == (2, 200)
    assert (out[0, 0], out[1, 0]) == (0.0, 0.0) and (out[0, -1], out[1, -1]) == (30.0, 10.0)
    t0 = math.degrees(math.atan2(out[1, 1] - out[1, 0], out[0, 1] - out[0, 0]))
    t1 = math.degrees(math.atan2(out[1, -1] - out[1, -2], out[0, -1] - out[0, -2]))
    assert abs(t0 - 0.0) < 0.5 and abs(t1 - 45.0) < 0.5
    # collinear poses give a straight line
    out = oracle.bezier([0.0, 0.0, 90.0, 0.0, 60.0, 90.0])
    assert np.allclose(out[0], 0.0, atol=1e-12) and np.all(np.diff(out[1]) > 0)


def test_mean_points_345_polyline(oracle):
    # legs of length 3 and 4 (total 7): uniform resample to 200 points
    out = oracle.mean_points([0.0, 3.0, 3.0], [0.0, 0.0, 4.0])
    assert (out[0, 0], out[1, 0]) == (0.0, 0.0) and (out[0, -1], out[1, -1]) == (3.0, 4.0)
    step = 7.0 / 199
    d = np.hypot(np.diff(out[0]), np.diff(out[1]))
    corner = int(3.0 / step)
    ok = np.ones(199, bool); ok[corner] = False             # the step that turns the corner is shorter in chord length
    assert np.allclose(d[ok], step, atol=1e-12)
    assert np.all(out[1, :corner + 1] == 0.0) and np.allclose(out[0, corner + 1:], 3.0, atol=1e-12)
    # degenerate inputs
    assert np.all(oracle.mean_points([], []) == 0.0)
    one = oracle.mean_points([2.0], [5.0])
    assert np.all(one[0] == 2.0) and np.all(one[1] == 5.0)
    dup = oracle.mean_points([1.0, 1.0, 2.0], [0.0, 0.0, 0.0])   # zero-length first segment
    assert dup[0, 0] == 1.0 and dup[0, -1] == 2.0 and np.all(np.diff(dup[0]) >= 0)


def test_ctr_rollout_and_tile_search(oracle):
    """constant-turn-rate rollout: a quarter turn per step walks the unit square; a tile of T = 1 is the static search"""
    import ctypes as C
    from dmpp_b200 import abi
    x, y = oracle.rollout_ctr(0.0, 0.0, 1.0, 0.0, 90.0, 5)
    assert np.allclose(x, [0, 1, 1, 0, 0], atol=1e-15) and np.allclose(y, [0, 0, 1, 1, 0], atol=1e-15)
    x, y = oracle.rollout_ctr(3.0, -2.0, 0.25, 0.1, 0.0, 9)            # no turn: exact arithmetic progression of binary fractions
    assert np.array_equal(x, 3.0 + 0.25 * np.arange(9))
    px, py = straight(120)
    ox = np.array([10.2, 30.1, 55.0]); oy = np.array([0.3, -0.5, 2.5])
    want = oracle.search_obstacle(px, py, ox, oy, -0.9, 0.9)
    out = np.zeros(1, abi.search_slot)
    oracle.lib.oracle_search_obstacle_tile(abi.ptr(np.ascontiguousarray(px)), abi.ptr(np.ascontiguousarray(py)), C.c_int(120),
                                           abi.ptr(ox), abi.ptr(oy), C.c_int(1), C.c_int(3), C.c_double(-0.9), C.c_double(0.9), abi.ptr(out))
    for f in ("dis_lat", "dis_lng", "ob_index", "pathid", "found"):
        assert out[0][f] == want[f], f
    # an agent that drives away along the path at the ego's pace is met later than a parked one
    T = 120
    tx = np.zeros((T, 1)); ty = np.zeros((T, 1))
    tx[:, 0] = 20.0 + 0.25 * np.arange(T)
    oracle.lib.oracle_search_obstacle_tile(abi.ptr(np.ascontiguousarray(px)), abi.ptr(np.ascontiguousarray(py)), C.c_int(120),
                                           abi.ptr(tx), abi.ptr(ty), C.c_int(T), C.c_int(1), C.c_double(-0.9), C.c_double(0.9), abi.ptr(out))
    assert out[0]["found"] == 1 and out[0]["pathid"] == 80 and out[0]["dis_lng"] == 40.0
