"""V2X event handlers (include/dmpp_b200.h section 10; Decision.cpp:1824-2434).

CPU side: the restatement (oracle/v2x_oracle.cpp) against the unmodified reference's own private methods -- V2XEventDecision as
SegmentDecision calls it (Decision.cpp:283) and V2XConstructionEventTemporal -- on seeded events that reach every branch.
GPU side: dp_v2x_event_batch against the oracle, flags and the two distances behind them bit for bit."""
import numpy as np
import pytest

FLAGS = ["light_flag", "construction_flag", "pedestrian_flag"]


def events(the_map, seed0, n, seed):
    from dmpp_b200 import scenes
    h = scenes.Episodes(the_map, np.arange(seed0, seed0 + n), cycles=1).hdr(0)
    return (h,) + scenes.v2x_events(the_map, h, seed=seed)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("seed0,n,seed", [(0, 4096, 1), (50000, 2048, 7)])
def test_v2x_oracle_equals_unmodified_reference(oracle, reference, the_map, mode, seed0, n, seed):
    h, v, wl, wg = events(the_map, seed0, n, seed)
    a = oracle.v2x_event(h, v, wl, wg, mode)
    b = reference.v2x_event(h, v, wl, wg, mode)
    ok = a["ub"] == 0                                        # (where the reference reads its path vector out of bounds: undefined there)
    assert ok.mean() > 0.9
    for f in FLAGS:
        assert np.array_equal(a[f][ok], b[f][ok]), f
    # every outcome of every handler occurs
    st = v["warn_status"]
    assert set(np.unique(a["light_flag"][ok & (st == 3)]).tolist()) == {0, 1, 2}
    for s, f in ((4, "construction_flag"), (5, "pedestrian_flag")):
        hit = a[f][ok & (st == s)]
        assert 50 < hit.sum() < hit.size - 50, (s, f)
    assert not a["construction_flag"][st != 4].any() and not a["pedestrian_flag"][st != 5].any() and not a["light_flag"][st != 3].any()


def test_v2x_float_quirk_and_last_index_quirk_matter(oracle, the_map):
    """the road-works handler keeps the nearest point in `float` variables (Decision.cpp:2021-2022) and tests it against the
    path index of the LAST list entry (:2118-2122): a restatement without either would decide differently on these inputs"""
    h, v, wl, wg = events(the_map, 0, 4096, 1)
    a = oracle.v2x_event(h, v, wl, wg, 0)
    sel = (v["warn_status"] == 4) & (a["ub"] == 0) & (a["lat_distance"] != 9999)
    lat = a["lat_distance"][sel]
    assert ((lat >= 0) & (lat < 1.875)).sum() > 50 and (lat < 0).sum() > 50 and (lat >= 1.875).sum() > 20
    # single-precision longitude: the lateral distance moves by decimetres, so it is not a multiple of the lateral grid drawn
    assert np.abs(lat - np.round(lat, 6)).max() > 0


def test_v2x_known_answers(oracle, the_map):
    """road 1 (straight east, 0.5 m spacing), ego on lane 2 at point 300"""
    from dmpp_b200 import abi
    m = the_map
    gl = m.lane_index(1, 2)
    off = int(m.lane_pt_off[gl])
    h = np.zeros(1, abi.scene_hdr)
    h["road_num"], h["lane_num"] = 1, 2
    h["id"][0, :3] = 300
    lat0, lng0, k_lat, k_lng = 23.0, 113.0, 1.0 / 110574.0, 1.0 / 102470.0

    def gps(i, left):
        return lat0 + (m.y[off + i] + left) * k_lat, lng0 + m.x[off + i] * k_lng

    v = np.zeros(1, abi.v2x_data)
    v["warn_status"], v["ped_distance"] = 5, 30.0
    v["ped_lat"], v["ped_lng"] = gps(360, 1.0)               # 30 m ahead, 1 m to the left: inside 1.5 m
    r = oracle.v2x_event(h, v, [], [])[0]
    assert r["pedestrian_flag"] == 1 and abs(r["lng_distance"] - 30.0) < 1e-6 and abs(r["lat_distance"] - 1.0) < 1e-6
    v["ped_lat"], v["ped_lng"] = gps(360, 3.0)               # 3 m to the left: the 'approaching' counter can never reach 5
    assert oracle.v2x_event(h, v, [], [])[0]["pedestrian_flag"] == 0
    v["ped_lat"], v["ped_lng"] = gps(360, -5.0)              # 5 m to the RIGHT: negative lateral distance <= 1.5 raises the flag
    assert oracle.v2x_event(h, v, [], [])[0]["pedestrian_flag"] == 1
    v["ped_distance"] = 130.0
    assert oracle.v2x_event(h, v, [], [])[0]["pedestrian_flag"] == 0
    v["warn_status"], v["spat_lane_occupied"], v["spat_state"] = 3, 1, 7
    assert oracle.v2x_event(h, v, [], [])[0]["light_flag"] == 1
    v["spat_state"] = 6
    assert oracle.v2x_event(h, v, [], [])[0]["light_flag"] == 2
    v["spat_lane_occupied"] = 0
    assert oracle.v2x_event(h, v, [], [])[0]["light_flag"] == 0
    v["warn_status"] = 4                                     # road works: the event point alone, 40 m ahead of point 308 = id + ID_MORE
    v["rsi_lat"], v["rsi_lng"] = gps(388, 0.5)
    r = oracle.v2x_event(h, v, [], [])[0]
    assert r["construction_flag"] == 1 and abs(r["lng_distance"] - 40.0) < 1e-6
    v["rsi_lat"] = 0.0
    assert oracle.v2x_event(h, v, [], [])[0]["construction_flag"] == 0


def test_v2x_layouts(oracle):
    from dmpp_b200 import abi
    assert oracle.lib.oracle_sizeof(11) == abi.v2x_data.itemsize == 80
    assert oracle.lib.oracle_sizeof(12) == abi.v2x_flags.itemsize == 24


@pytest.fixture(scope="module")
def planner(the_map):
    from dmpp_b200.planner import Planner
    p = Planner(max_scenes=64, max_obs=10)
    p.upload_map(the_map)
    yield p
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("seed0,n,seed", [(0, 8192, 1), (50000, 4096, 7), (123, 1, 3)])
def test_gpu_v2x_equals_oracle(planner, oracle, the_map, mode, seed0, n, seed):
    h, v, wl, wg = events(the_map, seed0, n, seed)
    want = oracle.v2x_event(h, v, wl, wg, mode)
    got = planner.v2x_event(h, v, wl, wg, mode)
    assert got.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_gpu_v2x_rejects_bad_slices(planner, the_map):
    from dmpp_b200.planner import DpError
    h, v, wl, wg = events(the_map, 0, 16, 1)
    v["wp_first"][3], v["wp_count"][3] = wl.size, 2
    with pytest.raises(DpError):
        planner.v2x_event(h, v, wl, wg)


def test_v2x_apply_spec(oracle):
    """the opt-in speed command: stop for a pedestrian or a red / yellow light, creep past road works, nothing else touched"""
    from dmpp_b200 import abi
    rec = np.zeros(5, abi.plan_record)
    rec["brakespeed"], rec["des_acc"], rec["behavior"] = [10.0, 8.0, 7.0, 2.0, 9.0], 0.0, 2
    f = np.zeros(5, abi.v2x_flags)
    f["pedestrian_flag"][0] = 1; f["light_flag"][1] = 1; f["construction_flag"][2] = 1; f["construction_flag"][3] = 1; f["light_flag"][4] = 2
    out = oracle.v2x_apply(f, rec)
    assert out["brakespeed"].tolist() == [0.0, 0.0, 3.0, 2.0, 9.0] and out["acc_flag"].tolist() == [1, 1, 0, 0, 0]
    assert out["des_acc"].tolist() == [-3.0, -3.0, 0.0, 0.0, 0.0] and (out["behavior"] == 2).all()


@pytest.mark.gpu
def test_gpu_v2x_apply_and_closed_loop_stop(planner, oracle, the_map):
    """CUDA == oracle on the adjusted records; and a world stepped with them brakes (a_max x dt per cycle) where a flag is up"""
    from dmpp_b200 import abi, scenes
    h, v, wl, wg = events(the_map, 0, 64, 1)
    flags = planner.v2x_event(h, v, wl, wg)
    ep = scenes.Episodes(the_map, np.arange(64), cycles=1)
    planner.reset(0, 64)
    rec = planner.cycle(np.ascontiguousarray(ep.hdr(0)), *ep.obstacles(0))["rec"]
    got = planner.v2x_apply(flags, rec)
    assert got.tobytes() == oracle.v2x_apply(flags, rec).tobytes()
    up = (flags["pedestrian_flag"] == 1) | (flags["light_flag"] == 1)
    assert up.sum() >= 5 and (got["brakespeed"][up] == 0).all() and (got["acc_flag"][up] == 1).all()
    assert got[~up & (flags["construction_flag"] == 0)].tobytes() == rec[~up & (flags["construction_flag"] == 0)].tobytes()
