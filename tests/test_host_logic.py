"""Host build of the kernels' RUMI_HD arithmetic (tests/host_harness) checked against the oracle WITHOUT a GPU:
the quad-tree core (closed-form level phase + std::sort replay), geometry tables, FAST score, fastAtan2, sincosf."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_frame

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hh():
    spec = importlib.util.spec_from_file_location("hh_build", os.path.join(HERE, "host_harness", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    L = C.CDLL(mod.build())
    L.hh_atan2.restype = C.c_float
    L.hh_atan2.argtypes = [C.c_float, C.c_float]
    L.hh_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return L


def u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def pack(xyr):
    x, y, r = (xyr[:, i].astype(np.uint32) for i in range(3))
    return (x | (y << 12) | (r << 24)).astype(np.uint32)


def run_octree(hh, W, H, nf, nlevels, level, xyr, N=-1):
    cand = pack(xyr)
    out = np.zeros(len(cand) + 64, np.uint32)
    n = hh.hh_octree(W, H, nf, C.c_float(1.2), nlevels, level, u32p(cand), len(cand), N, u32p(out), len(out))
    assert n >= 0
    o = out[:n]
    return np.stack([o & 0xFFF, (o >> 12) & 0xFFF, o >> 24], 1).astype(np.float32)


def test_geometry_matches_oracle(hh, oracle):
    for (W, H, nf) in [(640, 480, 1000), (752, 480, 1200), (1241, 376, 2000), (640, 480, 5000)]:
        info = np.zeros(12 * 8 + 32, np.int32)
        finfo = np.zeros(3 * 8, np.float32)
        assert hh.hh_level_info(W, H, nf, C.c_float(1.2), 8, 20, 7, info.ctypes.data_as(C.POINTER(C.c_int)),
                                finfo.ctypes.data_as(C.POINTER(C.c_float))) == 0
        t = oracle.tables(nf)
        ws, hs = oracle.level_sizes(W, H)
        lv = info[:96].reshape(8, 12)
        assert np.array_equal(lv[:, 0], ws) and np.array_equal(lv[:, 1], hs)
        assert np.array_equal(lv[:, 7], t["quota"])
        assert np.array_equal(info[96:112], t["umax"])
        assert np.array_equal(finfo.reshape(8, 3)[:, 1], t["scale"])
        assert np.array_equal(finfo.reshape(8, 3)[:, 2], np.floor(np.float32(31) * t["scale"]))
    small = np.zeros(200, np.int32)
    assert hh.hh_level_info(100, 100, 500, C.c_float(1.2), 8, 20, 7, small.ctypes.data_as(C.POINTER(C.c_int)),
                            np.zeros(24, np.float32).ctypes.data_as(C.POINTER(C.c_float))) < 0


def test_resize_coefficients(hh, oracle):
    img = synthetic_frame(1, 640, 480)
    ofs, a0, a1 = np.zeros(533, np.uint16), np.zeros(533, np.int16), np.zeros(533, np.int16)
    hh.hh_resize_coef(640, 533, ofs.ctypes.data_as(C.POINTER(C.c_uint16)), a0.ctypes.data_as(C.POINTER(C.c_int16)),
                      a1.ctypes.data_as(C.POINTER(C.c_int16)))
    yo, b0, b1 = np.zeros(400, np.uint16), np.zeros(400, np.int16), np.zeros(400, np.int16)
    hh.hh_resize_coef(480, 400, yo.ctypes.data_as(C.POINTER(C.c_uint16)), b0.ctypes.data_as(C.POINTER(C.c_int16)),
                      b1.ctypes.data_as(C.POINTER(C.c_int16)))
    s = img.astype(np.int64)
    x1 = np.minimum(ofs.astype(int) + 1, 639)
    hrow = s[:, ofs] * a0 + s[:, x1] * a1
    y1 = np.minimum(yo.astype(int) + 1, 479)
    out = ((((b0[:, None] * (hrow[yo] >> 4)) >> 16) + ((b1[:, None] * (hrow[y1] >> 4)) >> 16) + 2) >> 2)
    assert np.array_equal(out.astype(np.uint8), oracle.resize(img, 533, 400))


def test_std_sort_replay_matches_libstdcxx(hh):
    rng = np.random.default_rng(1)
    u64p = C.POINTER(C.c_uint64)
    for trial in range(600):
        n = int(rng.integers(0, 900))
        kmax = int(rng.choice([1, 2, 3, 5, 20, 1000]))
        v = (rng.integers(0, kmax, n).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        if trial % 5 == 0:
            v = np.sort(v)
        if trial % 7 == 0:
            v = np.sort(v)[::-1].copy()
        if trial % 11 == 0 and n > 4:                      # organ pipe: pushes introsort into its heapsort fallback
            k = np.concatenate([np.arange(n // 2), np.arange(n - n // 2)[::-1]]).astype(np.uint64)
            v = (k << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        a, b = v.copy(), v.copy()
        hh.hh_stdsort(a.ctypes.data_as(u64p), n)
        hh.hh_realsort(b.ctypes.data_as(u64p), n)
        assert np.array_equal(a, b)


def test_octree_core_on_real_candidates(hh, oracle):
    """Real FAST candidates of every level.  Also a census of the two ways the final phase is evaluated: the serial
    std::sort replay only runs when equal (count, UL.x) keys meet inside the part of the sorted array that is consumed;
    both ways must occur here (and both equal the oracle = the reference)."""
    phase_b = replays = 0
    for (W, H, nf) in [(640, 480, 1000), (752, 480, 1200), (1241, 376, 2000), (640, 480, 5000), (640, 480, 200)]:
        tb = oracle.tables(nf)
        for seed in range(4):
            py = oracle.pyramid(synthetic_frame(seed + 100, W, H))
            for l in range(8):
                xyr, _ = oracle.grid_fast(py[l])
                h, w = py[l].shape
                sel = oracle.octree(xyr, 16, w - 16, 16, h - 16, int(tb["quota"][l]))
                got = run_octree(hh, W, H, nf, 8, l, xyr)
                assert np.array_equal(xyr[sel], got), (W, H, nf, seed, l)
                st = hh.hh_octree_last_replays()
                phase_b += st >> 16
                replays += 1 if (st & 0xFFFF) else 0
    print("sorted final phase ran in %d problems, %d of them needed the serial std::sort replay" % (phase_b, replays))
    assert phase_b > 100 and 0 < replays < phase_b


def test_octree_core_adversarial(hh, oracle):
    rng = np.random.default_rng(5)
    for trial in range(800):
        W, H = [(640, 480), (752, 480), (1241, 376), (200, 150), (179, 134)][trial % 5]
        w, h = W - 32, H - 32
        M = int(rng.choice([0, 1, 2, 3, 5, 17, 100, 700, 3000]))
        mode = trial % 4
        if mode == 0:
            xs, ys = rng.integers(3, w - 3, M), rng.integers(3, h - 3, M)
        elif mode == 1:
            cx, cy, r = rng.integers(10, w - 10), rng.integers(10, h - 10), int(rng.choice([2, 5, 20]))
            xs = np.clip(cx + rng.integers(-r, r + 1, M), 3, w - 4)
            ys = np.clip(cy + rng.integers(-r, r + 1, M), 3, h - 4)
        elif mode == 2:
            xs, ys = rng.integers(3, w - 3, M), np.full(M, rng.integers(3, h - 3))
        else:
            xs, ys = np.full(M, rng.integers(3, w - 3)), rng.integers(3, h - 3, M)
        pts = np.unique(np.stack([xs, ys], 1), axis=0)
        rng.shuffle(pts)
        resp = rng.integers(7, int(rng.choice([9, 30, 255])), len(pts))
        xyr = np.concatenate([pts, resp[:, None]], 1).astype(np.float32)
        N = int(rng.choice([0, 1, 2, 3, 4, 5, 16, 17, 60, 217, 1086]))
        sel = oracle.octree(xyr, 16, W - 16, 16, H - 16, N)
        got = run_octree(hh, W, H, 1000, 2, 0, xyr, N)
        assert np.array_equal(xyr[sel], got), (trial, W, H, len(pts), N)


def test_scalar_math(hh, oracle):
    rng = np.random.default_rng(0)

    def ref_score(d):
        d2 = np.concatenate([d, d])
        return max(max(d2[k:k + 9].min(), (-d2[k:k + 9]).min()) for k in range(16)) - 1

    for _ in range(3000):
        d = rng.integers(-int(rng.choice([3, 20, 255])), 256, 16).astype(np.int32)
        assert hh.hh_fast_score(d.ctypes.data_as(C.POINTER(C.c_int))) == ref_score(d)
    for _ in range(20000):
        y, x = (int(v) for v in rng.integers(-300000, 300000, 2))
        assert hh.hh_atan2(y, x) == oracle.fast_atan2(y, x)
    assert hh.hh_atan2(0, 0) == 0.0
    s, c = C.c_float(0), C.c_float(0)
    for a in np.concatenate([np.linspace(0, 6.2832, 4000, dtype=np.float32),
                             np.float32(rng.uniform(0, 360, 4000)) * np.float32(np.pi / 180)]):
        hh.hh_sincos(float(a), C.byref(s), C.byref(c))
        rc, rs = oracle.sincos(a)
        assert (s.value, c.value) == (rs, rc)


def test_umma_bit_expansion_identity():
    """The arithmetic identity behind the tcgen05 matcher's operand expansion (rumi_slam_b200/csrc/match_umma.cu,
    expand_row): x * 0x8040201008040201 places bit i of the byte x at positions i + 9 j (all distinct: no carries), so
    masking with 0x80..80 leaves byte j = 0x80 iff bit 7 - j of x is set -- a fixed permutation of the 8 bits, applied to
    queries and train rows alike, which a dot product does not see."""
    C = 0x8040201008040201
    M = 0x8080808080808080
    for x in range(256):
        p = (x * C) & 0xFFFFFFFFFFFFFFFF & M
        for j in range(8):
            assert ((p >> (8 * j)) & 0xFF) == (0x80 if (x >> (7 - j)) & 1 else 0)
    # dot product of two expanded rows (queries 0/1, train 0/0x80) = 128 * popcount(a & b); Hamming from it
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
    ea = np.array([[(int(x) * C & M) >> (8 * j) & 0xFF for j in range(8)] for x in a]).ravel() >> 7
    eb = np.array([[(int(x) * C & M) >> (8 * j) & 0xFF for j in range(8)] for x in b]).ravel()
    acc = int((ea * eb).sum())
    pop = lambda v: int(np.unpackbits(v).sum())
    assert acc == 128 * pop(a & b)
    assert pop(a) + pop(b) - 2 * (acc >> 7) == pop(a ^ b)
