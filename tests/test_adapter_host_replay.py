"""The HOST logic of the C++ adapter's full SearchByProjection (window construction on both grids, order-dependent acceptance,
stereo-fisheye cross assignments) EXECUTED without a GPU: tests/adapter_driver/host_replay_main.cc links adapter/*.cc with
test-only stand-ins for the three C-ABI calls on that path (the Hamming distances the GPU supplies; covered on the device by
tests/test_gpu_candidates.py) and its output is compared with the oracle, which is pinned to the reference function."""
import os
import subprocess

import numpy as np
import pytest

from test_ref_frame_pin import _local_points_case          # noqa: E402  (tests/ is on sys.path under pytest)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def M():
    from oracle import match_oracle
    match_oracle.build()
    return match_oracle


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    from rumi_slam_b200 import _lib
    _lib.lib()                                              # the other matcher entry points still resolve against the library
    ad = os.path.join(ROOT, "rumi_slam_b200", "adapter")
    exe = tmp_path_factory.mktemp("host_replay") / "host_replay"
    cmd = ["g++", "-std=c++14", "-O1", "-pthread", "-I", os.path.join(ROOT, "oracle", "cvstub"), "-I", ad,
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "adapter_driver", "host_replay_main.cc"),
           os.path.join(ad, "ORBmatcher_accel.cc"), "-o", str(exe),
           "-L", os.path.join(ROOT, "rumi_slam_b200"), "-lrumi_orb", "-Wl,-rpath," + os.path.join(ROOT, "rumi_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(exe)


def _write(f, a, dt):
    a = np.zeros(0, dt) if a is None else np.ascontiguousarray(a, dt)
    f.write(np.int32(a.size if dt != KP else len(a)).tobytes())
    f.write(a.tobytes())


KP = None


@pytest.mark.parametrize("mode", ["plain", "occupied", "stereo", "fisheye"])
@pytest.mark.parametrize("seed,th,ratio", [(6, 3.0, 0.8), (7, 1.0, 0.8), (9, 15.0, 0.6)])
def test_cpp_search_by_projection_host_logic_equals_oracle(oracle, M, driver, tmp_path, seed, th, ratio, mode):
    global KP
    from rumi_slam_b200 import KP_DTYPE
    KP = KP_DTYPE
    kL, dF, dMP, kw = _local_points_case(oracle, seed, "occupied" if mode == "plain" else mode)
    if mode == "plain":
        kw.pop("occupied")
    sf = oracle.tables()["scale"]
    proj, level, view_cos, has_obs = kw.pop("proj"), kw.pop("level"), kw.pop("view_cos"), kw.pop("has_obs")
    n, fm = M.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio, **kw)
    scen, out = tmp_path / "scenario.bin", tmp_path / "out.bin"
    with open(scen, "wb") as f:
        _write(f, [th, ratio, 0, 0, 640, 480], np.float32)
        _write(f, np.ascontiguousarray(kL, KP_DTYPE), KP_DTYPE)
        _write(f, None if kw.get("kR") is None else np.ascontiguousarray(kw["kR"], KP_DTYPE), KP_DTYPE)
        _write(f, dF, np.uint8)
        _write(f, dMP, np.uint8)
        _write(f, sf, np.float32)
        _write(f, proj, np.float32)
        _write(f, kw.get("proj_r"), np.float32)
        _write(f, level, np.int32)
        _write(f, kw.get("level_r"), np.int32)
        _write(f, view_cos, np.float32)
        _write(f, kw.get("view_cos_r"), np.float32)
        _write(f, has_obs, np.uint8)
        _write(f, kw.get("occupied"), np.uint8)
        _write(f, kw.get("in_view"), np.uint8)
        _write(f, kw.get("in_view_r"), np.uint8)
        _write(f, kw.get("u_right"), np.float32)
        _write(f, kw.get("l2r"), np.int32)
        _write(f, kw.get("r2l"), np.int32)
    r = subprocess.run([driver, str(scen), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    buf = open(out, "rb").read()
    gn, nfm, n0 = np.frombuffer(buf, np.int32, 3)
    gfm = np.frombuffer(buf, np.int32, nfm, 12)
    assert gn == n and np.array_equal(gfm, fm)
    assert n > 40
    if mode == "plain":                                    # no extras: the original entry point gives the same answer
        n0s = int(np.frombuffer(buf, np.int32, 1, 12 + 4 * nfm)[0])
        fm0 = np.frombuffer(buf, np.int32, n0s, 16 + 4 * nfm)
        assert n0 == n and np.array_equal(fm0, fm)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,direction", [(51, 15.0, "none"), (53, 15.0, "forward"), (54, 15.0, "backward")])
def test_cpp_last_frame_fisheye_host_logic_equals_oracle(oracle, M, driver, tmp_path, seed, th, direction, check_ori):
    global KP
    from rumi_slam_b200 import KP_DTYPE
    from test_ref_frame_pin import _last_frame_fisheye_case
    KP = KP_DTYPE
    k1, d1, kC, kR, dC, uv, depth, valid, has_obs, occupied, shift = _last_frame_fisheye_case(oracle, seed)
    sf = oracle.tables()["scale"]
    fw, bw = direction == "forward", direction == "backward"
    invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
    uvr = (uv + np.asarray(shift, np.float32)).astype(np.float32)
    n, cm = M.search_by_projection_last_fisheye(kC, kR, dC, sf, (0, 0, 640, 480), valid, uv, uvr, invz, k1["octave"], k1["angle"],
                                                d1, has_obs, th, occupied, fw, bw, check_ori)
    scen, out = tmp_path / "scenario.bin", tmp_path / "out.bin"
    with open(scen, "wb") as f:
        _write(f, [th, 0.9, 0, 0, 640, 480, 1, fw, bw, check_ori], np.float32)
        _write(f, np.ascontiguousarray(kC, KP_DTYPE), KP_DTYPE)
        _write(f, np.ascontiguousarray(kR, KP_DTYPE), KP_DTYPE)
        _write(f, dC, np.uint8)
        _write(f, d1, np.uint8)
        _write(f, sf, np.float32)
        _write(f, occupied, np.uint8)
        _write(f, valid, np.uint8)
        _write(f, has_obs, np.uint8)
        _write(f, uv, np.float32)
        _write(f, uvr, np.float32)
        _write(f, invz, np.float32)
        _write(f, k1["angle"], np.float32)
        _write(f, k1["octave"], np.int32)
    r = subprocess.run([driver, str(scen), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    buf = open(out, "rb").read()
    gn, ncm, _ = np.frombuffer(buf, np.int32, 3)
    assert gn == n and np.array_equal(np.frombuffer(buf, np.int32, ncm, 12), cm)
    assert n > 100
