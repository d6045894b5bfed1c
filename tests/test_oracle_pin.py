"""Pins the oracle (CPU) against (1) the committed golden vectors (cv2 primitives + the unmodified reference
ORBextractor.cc compiled over the cv stub), (2) cv2 itself when importable, (3) oracle/_ref when built."""
import os

import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_frame

G = os.path.join(os.path.dirname(__file__), "golden")


def test_constructor_tables(oracle):
    t = oracle.tables(1000)
    assert list(t["quota"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(oracle.tables(1200)["quota"]) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert list(oracle.tables(2000)["quota"]) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(oracle.tables(5000)["quota"]) == [1086, 905, 754, 628, 524, 436, 364, 303]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    ws, hs = oracle.level_sizes(640, 480)
    assert list(zip(ws, hs)) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161),
                                 (179, 134)]
    assert [int(31 * s) for s in t["scale"]] == [31, 37, 44, 53, 64, 77, 92, 111]


def test_golden_cv2_primitives(oracle):
    g = np.load(os.path.join(G, "cv2_primitives.npz"))
    pyr = oracle.pyramid(g["img"], nlevels=4)
    for l in (1, 2, 3):
        assert np.array_equal(pyr[l], g["pyr%d" % l])
    for name in ("noise", "patch"):
        for th in (20, 7):
            assert np.array_equal(oracle.fast(g[name], th), g["%s_fast%d" % (name, th)])
        assert np.array_equal(oracle.blur(g[name]), g[name + "_blur"])
    assert np.array_equal(oracle.grid_fast(g["grid_img"])[0], g["grid_cand"])
    got = np.array([oracle.fast_atan2(y, x) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got, g["atan_deg"])


def test_golden_reference_extract(oracle):
    g = np.load(os.path.join(G, "reference_extract.npz"))
    for name in "abc":
        nf, nl, l0, l1, mono = (int(v) for v in g["meta_" + name])
        k, d, m = oracle.extract(g["img_" + name], nfeatures=nf, nlevels=nl, lapping=(l0, l1))
        assert m == mono
        assert np.array_equal(k.view(np.uint8).reshape(-1, 28), g["kps_" + name])
        assert np.array_equal(d, g["desc_" + name])
    sel = oracle.octree(g["oct_cand"], 16, 640 - 16, 16, 480 - 16, 217)
    assert np.array_equal(g["oct_cand"][sel], g["oct_sel"])


def test_golden_matcher(oracle):
    g = np.load(os.path.join(G, "matcher_kats.npz"))
    i1, d1, d2 = oracle.hamming_top2(g["Q"], g["T"])
    assert np.array_equal(d1, g["bf_d1"]) and np.array_equal(d2, g["bf_d2"])      # cv::BFMatcher distances
    # indices: ours keeps the EARLIEST index among ties (reference scan); check they point at equal distances
    T, Q = g["T"], g["Q"]
    for q in range(0, len(Q), 17):
        assert oracle.descriptor_distance(Q[q], T[i1[q]]) == d1[q]
        assert i1[q] == min(t for t in range(len(T)) if oracle.descriptor_distance(Q[q], T[t]) == d1[q])
    i1, d1, d2 = oracle.hamming_top2(np.zeros((1, 32), np.uint8), g["kat_T"])
    assert [i1[0], d1[0], d2[0]] == list(g["kat_expect"])
    i1, d1, d2 = oracle.hamming_top2(np.zeros((1, 32), np.uint8), np.zeros((0, 32), np.uint8))
    assert (i1[0], d1[0], d2[0]) == (-1, 256, 256)


def test_oracle_vs_cv2_live(oracle):
    from oracle import cv2_oracle as A
    if not A.HAVE_CV2:
        pytest.skip("cv2 not importable here; covered by the golden vectors")
    for (w, h, seed) in [(640, 480, 31), (752, 480, 32), (1241, 376, 33)]:
        img = synthetic_frame(seed, w, h)
        ws, hs = oracle.level_sizes(w, h)
        pa, pb = A.pyramid(img, ws, hs), oracle.pyramid(img)
        for l in range(8):
            assert np.array_equal(pa[l], pb[l])
            ca, na = A.grid_fast(pb[l])
            cb, nb = oracle.grid_fast(pb[l])
            assert na == nb and np.array_equal(ca, cb)
            assert np.array_equal(A.blur(pb[l]), oracle.blur(pb[l]))
    Q = oracle.extract(synthetic_frame(1))[1]
    T = oracle.extract(synthetic_frame(2))[1]
    _, d1, d2 = oracle.hamming_top2(Q, T)
    _, b1, b2 = A.knn2(Q, T)
    assert np.array_equal(d1, b1) and np.array_equal(d2, b2)


def test_oracle_vs_unmodified_reference_live(oracle):
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built (reference sources absent); covered by the golden vectors")
    for (w, h, nf, lap, seed) in [(640, 480, 1000, (0, 0), 3), (640, 480, 1000, (0, 1000), 4),
                                  (752, 480, 1200, (0, 0), 5), (1241, 376, 2000, (0, 1000), 6)]:
        img = synthetic_frame(seed, w, h)
        a = oracle.extract(img, nfeatures=nf, lapping=lap)
        b = ref_lib.extract(img, nfeatures=nf, lapping=lap)
        assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    ta, tb = oracle.tables(), ref_lib.tables()
    for k in tb:
        assert np.array_equal(ta[k], tb[k])
    assert oracle.extract(np.zeros((0, 0), np.uint8)) is None and ref_lib.extract(np.zeros((0, 0), np.uint8)) is None


WIDE_CONFIGS = [  # (w, h, nfeatures, lapping): the five BASELINE / Tracking.cc configurations
    (640, 480, 1000, (0, 0)), (640, 480, 2000, (0, 1000)), (640, 480, 5000, (0, 0)),
    (752, 480, 1200, (0, 0)), (1241, 376, 2000, (0, 1000))]


def test_oracle_equals_unmodified_reference_on_2500_frames(oracle):
    """500 distinct synthetic frames x the 5 configurations: the whole operator() output of the oracle (keypoints in
    order, angles, descriptors, monoIndex) is bit-identical to the UNMODIFIED reference ORBextractor.cc (oracle/_ref).
    SURVEY.md 7 measured ~1 octree problem in 20 to be tie-order sensitive: 2 500 frames x 8 levels = 20 000 quad-tree
    problems go through the reference's own std::list / std::sort here."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built (reference sources absent); covered by the golden vectors")
    nframes = int(os.environ.get("RUMI_PIN_FRAMES", "500"))
    ref_lib.lib(); oracle.lib()

    def one(job):
        (w, h, nf, lap), seed = job
        img = synthetic_frame(100000 + seed, w, h)
        a = oracle.extract(img, nfeatures=nf, lapping=lap)
        b = ref_lib.extract(img, nfeatures=nf, lapping=lap)
        return a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), len(a[0])

    jobs = [(cfg, 1000 * ci + s) for ci, cfg in enumerate(WIDE_CONFIGS) for s in range(nframes)]
    with ThreadPoolExecutor(max_workers=len(os.sched_getaffinity(0))) as ex:
        res = list(ex.map(one, jobs))
    bad = [jobs[i] for i, r in enumerate(res) if not r[0]]
    assert not bad, "oracle != reference on %d of %d frames, first: %s" % (len(bad), len(jobs), bad[:3])
    assert sum(r[1] for r in res) > 900 * len(jobs)


def test_octree_equals_reference_on_2400_problems_with_tie_census(oracle):
    """2 400 standalone DistributeOctTree problems (real FAST candidates of all 8 levels, plus clustered / collinear /
    duplicated-column candidate sets built to provoke compareNodes ties) through the reference's own function
    (ref_octree).  The census counts how many of them are tie-order sensitive -- their result changes when the final
    phase uses std::stable_sort instead of libstdc++'s introsort -- to show the pin exercises that behaviour."""
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(11)
    problems = []
    for seed in range(150):                                   # 150 frames x 8 levels = 1 200 real problems
        img = synthetic_frame(200000 + seed, 640, 480)
        pyr = oracle.pyramid(img)
        quota = oracle.tables(1000)["quota"] if seed % 3 else oracle.tables(2000)["quota"]
        for l in range(8):
            cand, _ = oracle.grid_fast(pyr[l])
            lh, lw = pyr[l].shape
            problems.append((cand, 16, lw - 16, 16, lh - 16, int(quota[l])))
    for i in range(1200):                                     # synthetic candidate sets
        w = int(rng.integers(80, 1300))
        h = int(rng.integers(60, min(500, int(w / 0.55))))     # nIni = round(w / h) >= 1 (0 roots: the reference crashes)
        n = int(rng.integers(1, 3000))
        kind = i % 4
        if kind == 0:                                         # uniform
            x, y = rng.integers(0, w, n), rng.integers(0, h, n)
        elif kind == 1:                                       # a few tight clusters: many equal node counts
            c = rng.integers(0, [w, h], (int(rng.integers(2, 12)), 2))
            p = c[rng.integers(0, len(c), n)] + rng.integers(-6, 7, (n, 2))
            x, y = np.clip(p[:, 0], 0, w - 1), np.clip(p[:, 1], 0, h - 1)
        elif kind == 2:                                       # lattice: equal counts AND equal UL.x in every column
            step = int(rng.integers(3, 17))
            gx, gy = np.meshgrid(np.arange(0, w, step), np.arange(0, h, step))
            x, y = gx.ravel(), gy.ravel()
        else:                                                 # vertical lines: nodes stacked over the same UL.x
            cols = rng.integers(0, w, int(rng.integers(1, 9)))
            x, y = cols[rng.integers(0, len(cols), n)], rng.integers(0, h, n)
        xyr = np.stack([x, y, rng.integers(7, 120, len(x))], 1).astype(np.float32)
        xyr = xyr[np.sort(np.unique(xyr[:, :2], axis=0, return_index=True)[1])]      # distinct pixels, original order
        problems.append((xyr, 0, w, 0, h, int(rng.integers(1, max(2, len(xyr))))))
    sensitive = 0
    for xyr, x0, x1, y0, y1, N in problems:
        sel = oracle.octree(xyr, x0, x1, y0, y1, N)
        ref = ref_lib.octree(xyr, x0, x1, y0, y1, N)
        assert np.array_equal(np.asarray(xyr, np.float32)[sel], ref), (len(xyr), x1 - x0, y1 - y0, N)
        if not np.array_equal(oracle.octree_stable(xyr, x0, x1, y0, y1, N), sel):
            sensitive += 1
    print("octree problems: %d, tie-order sensitive: %d" % (len(problems), sensitive))
    assert len(problems) >= 2400 and sensitive >= 50, sensitive


def test_cloud_frame_compute_descriptors_equals_reference(oracle):
    """a12: oracle.describe == the reference's own CloudFrameComputeDescriptors (R/lib_src/ORBextractor.cc:989-1011) for
    keypoints of every octave used as-is (level-0 coordinates, given angles), and -1 on an empty image."""
    from oracle import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(5)
    for seed, (w, h) in enumerate([(640, 480), (752, 480), (1241, 376), (333, 257)]):
        img = synthetic_frame(300 + seed, w, h)
        k = oracle.extract(img, nfeatures=1500)[0]
        k = k[(k["x"] >= 19) & (k["y"] >= 19) & (k["x"] < w - 19) & (k["y"] < h - 19)].copy()
        k["angle"][::3] = (rng.random(len(k[::3])) * 360).astype(np.float32)     # arbitrary given angles
        a, b = oracle.describe(img, k), ref_lib.describe(img, k)
        assert a[0] == b[0] == len(k) and np.array_equal(a[1], b[1])
    assert ref_lib.describe(np.zeros((0, 0), np.uint8), k)[0] == -1 == oracle.describe(np.zeros((0, 0), np.uint8), k)[0]


def test_sincos_is_glibc(oracle):
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.cosf.restype = libm.sinf.restype = ctypes.c_float
    libm.cosf.argtypes = libm.sinf.argtypes = [ctypes.c_float]
    for a in np.linspace(0, 6.2831, 500, dtype=np.float32):
        c, s = oracle.sincos(a)
        assert c == libm.cosf(float(a)) and s == libm.sinf(float(a))
