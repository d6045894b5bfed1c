"""The C++ adapter EXECUTES on the GPU: a g++-built driver calls ORB_SLAM3::ORBextractor::operator() (two instances in
two threads, as Frame.cc:116-119 does) and ORBmatcherAccel::ComputeStereoMatches through rumi_slam_b200/adapter ->
C ABI -> CUDA kernels, and its output is compared bit for bit with the oracle: keypoints (order included), descriptors,
monoIndex, the public mvImagePyramid, mvuRight / mvDepth."""
import os
import subprocess

import numpy as np
import pytest

from rumi_slam_b200.synth import stereo_pair, synthetic_frame

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    from rumi_slam_b200 import _lib
    _lib.lib()
    ad = os.path.join(ROOT, "rumi_slam_b200", "adapter")
    exe = tmp_path_factory.mktemp("adapter") / "adapter_driver"
    # OpenCV headers are absent in this image: oracle/cvstub provides cv::Mat / cv::KeyPoint (containers only; with a
    # real OpenCV the same sources build unchanged)
    cmd = ["g++", "-std=c++14", "-O1", "-pthread", "-I", os.path.join(ROOT, "oracle", "cvstub"), "-I", ad,
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "adapter_driver", "main.cc"),
           os.path.join(ad, "ORBextractor.cc"), os.path.join(ad, "ORBmatcher_accel.cc"), "-o", str(exe),
           "-L", os.path.join(ROOT, "rumi_slam_b200"), "-lrumi_orb", "-Wl,-rpath," + os.path.join(ROOT, "rumi_slam_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(exe)


def read_frame(buf, off, kp_dtype):
    mono, nkp = np.frombuffer(buf, np.int32, 2, off); off += 8
    kps = np.frombuffer(buf, kp_dtype, nkp, off).copy(); off += 28 * nkp
    desc = np.frombuffer(buf, np.uint8, 32 * nkp, off).reshape(nkp, 32).copy(); off += 32 * nkp
    nl = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    pyr = []
    for _ in range(nl):
        w, h = np.frombuffer(buf, np.int32, 2, off); off += 8
        pyr.append(np.frombuffer(buf, np.uint8, w * h, off).reshape(h, w).copy()); off += w * h
    return int(mono), kps, desc, pyr, off


@pytest.mark.parametrize("lap", [(0, 0), (0, 1000)])
def test_operator_call_through_cpp_adapter(oracle, driver, tmp_path, lap):
    img = synthetic_frame(77, 640, 480)
    (tmp_path / "l.raw").write_bytes(img.tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([driver, str(out), "640", "480", "1000", str(lap[0]), str(lap[1]), str(tmp_path / "l.raw")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = out.read_bytes()
    mono, kps, desc, pyr, off = read_frame(buf, 0, oracle.KP_DTYPE)
    rk, rd, rmono = oracle.extract(img, lapping=lap)
    assert mono == rmono and np.array_equal(kps, rk) and np.array_equal(desc, rd)
    for got, want in zip(pyr, oracle.pyramid(img)):
        assert np.array_equal(got, want)
    assert np.frombuffer(buf, np.int32, 1, off)[0] == -1          # empty image -> -1


def test_stereo_frame_through_cpp_adapter(oracle, driver, tmp_path):
    left, right = stereo_pair(4, 752, 480)
    (tmp_path / "l.raw").write_bytes(left.tobytes())
    (tmp_path / "r.raw").write_bytes(right.tobytes())
    out = tmp_path / "out.bin"
    fx, bf = 435.2, 47.9
    r = subprocess.run([driver, str(out), "752", "480", "1200", "0", "0", str(tmp_path / "l.raw"), str(tmp_path / "r.raw"),
                        repr(bf), repr(bf / fx)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = out.read_bytes()
    lm, lk, ld, lp, off = read_frame(buf, 0, oracle.KP_DTYPE)
    rm, rk, rd, rp, off = read_frame(buf, off, oracle.KP_DTYPE)
    for (mono, k, d), img in (((lm, lk, ld), left), ((rm, rk, rd), right)):
        ok, od, omono = oracle.extract(img, nfeatures=1200)
        assert mono == omono and np.array_equal(k, ok) and np.array_equal(d, od)
    n = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    u = np.frombuffer(buf, np.float32, len(lk), off); off += 4 * len(lk)
    dep = np.frombuffer(buf, np.float32, len(lk), off)
    ru, rdp, rn = oracle.stereo_match(left, right, lk, ld, rk, rd, bf, bf / fx)      # double -> float at the C boundary, both sides
    assert n == rn and n > 100 and np.array_equal(u, ru) and np.array_equal(dep, rdp)
    off += 4 * len(lk)
    # ORBmatcherAccel::SearchForInitialization and ::AssociateSubmap, executed in C++
    from oracle import match_oracle as M
    ni = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    m12 = np.frombuffer(buf, np.int32, len(lk), off); off += 4 * len(lk)
    prev = np.stack([lk["x"], lk["y"]], 1).astype(np.float32)
    rn_, rm_, _ = M.search_for_initialization(lk, ld, rk, rd, (0, 0, 752, 480), prev, 100, 0.9, True)
    assert ni == rn_ and np.array_equal(m12, rm_) and rn_ > 20
    na, n1 = (int(v) for v in np.frombuffer(buf, np.int32, 2, off)); off += 8
    a12 = np.frombuffer(buf, np.int32, n1, off)
    k1, k2 = lk[lk["octave"] == 0], rk[rk["octave"] == 0]
    d1, d2 = oracle.describe(left, k1)[1], oracle.describe(right, k2)[1]
    i1, e1, e2 = oracle.hamming_top2(d1, d2)
    ok = (i1 >= 0) & (e1 <= 50) & (e1.astype(np.float32) < np.float32(0.75) * e2.astype(np.float32))
    assert n1 == len(k1) and na == int(ok.sum()) and np.array_equal(a12, np.where(ok, i1, -1))
    off += 4 * n1
    # ORBmatcherAccel::SearchByProjectionLastFrame / ::SearchByProjectionKeyFrame, executed in C++ on the same inputs
    nl, nk, nR = (int(v) for v in np.frombuffer(buf, np.int32, 3, off)); off += 12
    cm_last = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    cm_kf = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    i = np.arange(len(lk))
    valid, has_obs = i % 7 != 0, i % 5 != 0
    uv = np.stack([lk["x"] - np.float32(2.0), lk["y"] + np.float32(1.0)], 1).astype(np.float32)
    d3 = (1.0 + (i % 13)).astype(np.float32)
    invz = (np.float32(1.0) / d3).astype(np.float32)
    sf = oracle.tables()["scale"]
    rn1, rcm1 = M.search_by_projection_last(rk, rd, sf, (0, 0, 752, 480), valid, uv, invz, lk["octave"], lk["angle"], ld, has_obs,
                                            15.0, None, None, 0.0, False, False, True)
    rn2, rcm2 = M.search_by_projection_kf(rk, rd, sf, (0, 0, 752, 480), valid, uv, d3, np.full(len(lk), 0.5, np.float32),
                                          np.full(len(lk), 12.0, np.float32), lk["octave"], lk["angle"], ld, 10.0, 100, None, True)
    assert nR == len(rk) and nl == rn1 and np.array_equal(cm_last, rcm1) and rn1 > 50
    assert nk == rn2 and np.array_equal(cm_kf, rcm2) and rn2 > 30
    # ORBmatcherAccel::FuseSearch, executed in C++: map point i = key point (7 i) mod nR of the key frame, half a pixel off
    nf = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    f_idx = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    f_dist = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    inv_s2 = (np.float32(1.0) / (sf * sf).astype(np.float32)).astype(np.float32)
    ii = np.arange(nR)
    j = (ii * 7) % nR
    uv_f = np.stack([rk["x"][j] + np.float32(0.5), rk["y"][j] - np.float32(0.5)], 1).astype(np.float32)
    d3_f = (1.0 + (ii % 13)).astype(np.float32)
    ur_f = (uv_f[:, 0] - (np.float32(40.0) / d3_f).astype(np.float32)).astype(np.float32)
    rn3, rbest, rdist = M.fuse_search(rk, rd, sf, inv_s2, (0, 0, 752, 480), np.full(nR, -1.0, np.float32), ii % 9 != 0, uv_f, ur_f,
                                      d3_f, np.zeros(nR, np.float32), np.full(nR, 1e9, np.float32), rk["octave"][j], rd[j], 3.0)
    assert nf == rn3 and np.array_equal(f_idx, rbest) and np.array_equal(f_dist, rdist) and rn3 > 300
    # ORBmatcherAccel::SearchForTriangulation, executed in C++ ("node" of a feature = first descriptor byte mod 16)
    from oracle import bow_oracle as B
    B.build()
    nt = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    m12t = np.frombuffer(buf, np.int32, len(lk), off); off += 4 * len(lk)
    fv = lambda d: {n: [int(v) for v in np.flatnonzero(d[:, 0] % 16 == n)] for n in range(16) if np.any(d[:, 0] % 16 == n)}
    il, ir = np.arange(len(lk)), np.arange(nR)
    epi = (il[:, None] * 31 + ir[None, :] * 17) % 5 != 0
    rn4, rm4 = B.search_for_triangulation(ld, lk["angle"], il % 4 == 0, il % 3 == 0, fv(ld), rd, rk["angle"], ir % 5 == 0, ir % 2 == 0,
                                          rk["x"], rk["y"], rk["octave"], fv(rd), sf, (376.0, 240.0), epi, False, False, True)
    assert nt == rn4 and np.array_equal(m12t, rm4) and rn4 > 20
    # ORBmatcherAccel::SearchByProjectionSim3 / ::FuseSearchSim3 on the Fuse scene, executed in C++
    ns, nfs = (int(v) for v in np.frombuffer(buf, np.int32, 2, off)); off += 8
    km_s = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    f_idx_s = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    zeros, big = np.zeros(nR, np.float32), np.full(nR, 1e9, np.float32)
    rn5, rkm5 = M.search_by_projection_sim3(rk, rd, sf, (0, 0, 752, 480), ir % 3 == 0, ii % 9 != 0, uv_f, d3_f, zeros, big,
                                            rk["octave"][j], rd[j], 4, 0.8)
    rn6, rbest6, _ = M.fuse_search(rk, rd, sf, np.zeros(len(sf), np.float32), (0, 0, 752, 480), np.full(nR, -1.0, np.float32),
                                   ii % 9 != 0, uv_f, zeros, d3_f, zeros, big, rk["octave"][j], rd[j], 4.0)
    assert ns == rn5 and np.array_equal(km_s, rkm5) and rn5 > 100
    assert nfs == rn6 and np.array_equal(f_idx_s, rbest6) and rn6 > 300
    # ORBmatcherAccel::SearchBySim3 (right image against itself, half-pixel offsets), executed in C++
    nsim = int(np.frombuffer(buf, np.int32, 1, off)[0]); off += 4
    m12s = np.frombuffer(buf, np.int32, nR, off); off += 4 * nR
    uv1 = np.stack([rk["x"] + np.float32(0.5), rk["y"] - np.float32(0.5)], 1).astype(np.float32)
    uv2 = np.stack([rk["x"] - np.float32(0.25), rk["y"] + np.float32(0.25)], 1).astype(np.float32)
    rn7, rm7 = M.search_by_sim3(rk, rd, rk, rd, sf, (0, 0, 752, 480), ir % 4 != 0, uv1, d3_f, zeros, big, rk["octave"], ir % 5 != 0,
                                uv2, d3_f, zeros, big, rk["octave"], 7.5)
    assert nsim == rn7 and np.array_equal(m12s, rm7) and rn7 > 300
