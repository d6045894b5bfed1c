"""GPU parity tests of the sparse pyramidal Lucas-Kanade flow (KFDSample::Step's calcOpticalFlowPyrLK, SURVEY.md 8f
rank 4) through the C ABI (rumi_flow_*) against the oracle.

The device accumulates the normal equations exactly like the oracle (integer sums) and runs the float32 tail with
explicit IEEE operations, so the comparison with the ORACLE is bit-exact: pyramid levels, Scharr derivatives, tracked
positions, status flags and err values (tolerance used: 0).  Against OpenCV itself (frozen cv2 outputs) the same
tolerance as the oracle pin applies (tests/test_flow_oracle.py)."""
import os

import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_kats.npz")


@pytest.fixture(scope="module")
def fo():
    from oracle import flow_oracle
    flow_oracle.build()
    return flow_oracle


def shifted(img, dx, dy, seed, noise=2):
    """Sub-pixel translation by bilinear blending of rolled copies (+ noise): synthetic motion without OpenCV."""
    ix, iy = int(np.floor(dx)), int(np.floor(dy))
    fx, fy = dx - ix, dy - iy
    f = img.astype(np.float64)
    r = lambda a, b: np.roll(np.roll(f, iy + a, 0), ix + b, 1)
    out = (1 - fx) * (1 - fy) * r(0, 0) + fx * (1 - fy) * r(0, 1) + (1 - fx) * fy * r(1, 0) + fx * fy * r(1, 1)
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(out) + rng.integers(-noise, noise + 1, img.shape), 0, 255).astype(np.uint8)


def points(w, h, n, seed):
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1).astype(np.float32)
    edge = [[0, 0], [w - 1, h - 1], [0.5, 0.5], [w - 1.5, 3], [5, h - 1], [w - 1, 0], [15.25, 15.75], [-40, 10],
            [w + 40, h + 40], [w / 2, h / 2]]
    pts[:len(edge)] = edge
    return pts


@pytest.mark.parametrize("w,h,n,dx,dy", [(640, 480, 1000, 2.3, -1.6), (752, 480, 1200, -4.4, 0.8),
                                         (333, 257, 400, 0.4, 3.7), (160, 120, 150, 1.2, 0.6),
                                         (1241, 376, 2000, 6.5, -0.3), (70, 66, 40, 0.7, 0.2)])
def test_track_equals_oracle(fo, w, h, n, dx, dy):
    from rumi_slam_b200 import SparsePyrLK
    prev = synthetic_batch(1, w, h, seed0=w + h)[0]
    nxt = shifted(prev, dx, dy, seed=n)
    pts = points(w, h, n, 3)
    lk = SparsePyrLK()
    gn, gs, ge = lk.calc(prev, nxt, pts)
    on, os_, oe = fo.lk(prev, nxt, pts)
    # integer stages
    a = prev
    for level in range(lk.levels()):
        assert np.array_equal(lk.pyramid_level(level, 0), a), "pyramid level %d" % level
        assert np.array_equal(lk.derivatives(level), fo.scharr(a)), "Scharr level %d" % level
        a = fo.pyr_down(a)
    assert np.array_equal(gs, os_)
    assert np.array_equal(gn.view(np.uint32), on.view(np.uint32)), np.abs(gn - on).max()
    assert np.array_equal(ge.view(np.uint32), oe.view(np.uint32))
    assert gs[10:].mean() > 0.8                    # the synthetic motion is actually tracked
    good = gs == 1
    assert np.abs((gn - pts)[good][10:] - [dx, dy]).mean() < 0.5


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_track_vs_frozen_opencv(case):
    from rumi_slam_b200 import SparsePyrLK
    g = np.load(GOLD)
    gn, gs, ge = SparsePyrLK().calc(g[case + "_prev"], g[case + "_next"], g[case + "_pts"])
    cn, cs, ce = g[case + "_cv_next"], g[case + "_cv_status"], g[case + "_cv_err"]
    assert np.array_equal(gs, cs)
    good = cs == 1
    d = np.abs(gn - cn).max(1)[good]
    assert d.max() <= 0.03 and np.median(d) <= 1e-3        # tolerance of the oracle pin (tests/test_flow_oracle.py)
    assert np.abs(ge - ce)[good].max() <= 0.05


@pytest.mark.parametrize("win,max_level,count,eps", [(21, 3, 30, 0.01), (15, 0, 10, 0.03), (31, 5, 100, 0.0)])
def test_other_parameters(fo, win, max_level, count, eps):
    from rumi_slam_b200 import SparsePyrLK
    prev = synthetic_batch(1, 400, 300, seed0=5)[0]
    nxt = shifted(prev, -1.7, 2.2, seed=1)
    pts = points(400, 300, 300, 8)
    gn, gs, ge = SparsePyrLK(win, max_level, count, eps).calc(prev, nxt, pts)
    on, os_, oe = fo.lk(prev, nxt, pts, win, max_level, count, eps)
    assert np.array_equal(gs, os_) and np.array_equal(gn, on) and np.array_equal(ge, oe)


def test_sequence_with_advance(fo):
    """KFDSample's steady state: every frame is uploaded once, the tracked frame becomes the previous one."""
    from rumi_slam_b200 import SparsePyrLK
    base = synthetic_batch(1, 640, 480, seed0=77)[0]
    frames = [base] + [shifted(base, 1.3 * k, -0.8 * k, seed=k) for k in range(1, 5)]
    pts = points(640, 480, 500, 4)
    lk = SparsePyrLK()
    lk.set_prev(frames[0])
    cur = pts
    for k in range(1, 5):
        gn, gs, ge = lk.track_next(frames[k], cur, advance=True)
        on, os_, oe = fo.lk(frames[k - 1], frames[k], cur)
        assert np.array_equal(gs, os_) and np.array_equal(gn, on) and np.array_equal(ge, oe), k
        cur = gn                                   # `old = next` (KFDSample.cc:166): lost points are carried along
    assert lk.launches() == 3 + 4 * 4              # set_prev: 2 pyrDown + Scharr; per frame: 2 pyrDown + LK + Scharr


def test_edge_cases(fo):
    from rumi_slam_b200 import SparsePyrLK, RumiError
    prev = synthetic_batch(1, 320, 240, seed0=1)[0]
    lk = SparsePyrLK()
    n, s, e = lk.calc(prev, prev, np.zeros((0, 2), np.float32))
    assert n.shape == (0, 2) and s.shape == (0,)
    flat = np.full_like(prev, 90)
    n, s, e = lk.calc(flat, flat, np.array([[80.0, 60.0], [-500.0, 3.0]], np.float32))
    assert list(s) == [0, 0] and list(e) == [0, 0]
    # non-contiguous rows (a cv::Mat ROI)
    big = synthetic_batch(1, 400, 300, seed0=2)[0]
    roi_p, roi_n = big[20:260, 30:350], shifted(big, 1.5, 0.5, 3)[20:260, 30:350]
    pts = points(320, 240, 200, 6)
    gn, gs, ge = lk.calc(roi_p, roi_n, pts)
    on, os_, oe = fo.lk(np.ascontiguousarray(roi_p), np.ascontiguousarray(roi_n), pts)
    assert np.array_equal(gs, os_) and np.array_equal(gn, on) and np.array_equal(ge, oe)
    with pytest.raises(RumiError):
        SparsePyrLK(win=17)
    with pytest.raises(RumiError):
        SparsePyrLK().track_next(prev, pts)        # no previous frame
    with pytest.raises(ValueError):
        lk.track_next(prev[:100], pts)


def test_kfdsample_step_equals_oracle_replay(fo, oracle):
    """The KFDSample mirror (device extraction + device flow + host PD selector) takes the same key-frame decisions,
    with the same flow magnitudes, as a replay of KFDSample::Step built from the oracles."""
    from rumi_slam_b200 import KFDSample
    base = synthetic_batch(1, 640, 480, seed0=123)[0]
    shifts = [0, 0.6, 1.5, 3.0, 6.0, 6.5, 7.5, 12.0, 12.4, 20.0]
    frames = [shifted(base, s, -0.5 * s, seed=i) for i, s in enumerate(shifts)]
    ts = [0.05 * i for i in range(len(frames))]
    kf = KFDSample(nfeatures=1000, th=2.0)
    # oracle replay of KFDSample.cc:88-175
    pd, old, prev, last = fo.PD(0.8, 0.005, 2.0), None, None, 0.0
    for i, (im, t) in enumerate(zip(frames, ts)):
        got = kf.Step(im, t)
        if old is None:
            k, _, _ = oracle.extract(im, nfeatures=1000)
            old = np.stack([k["x"], k["y"]], 1).astype(np.float32)
            want, last = True, t
        else:
            nxt, st, _ = fo.lk(prev, im, old)
            mag, _ = fo.mean_magnitude(old, nxt, st)
            th = np.float32(np.float32(mag) + np.float32(pd.update(mag, t - last)))
            want = bool(np.float32(mag) > th)
            assert np.float32(mag) == kf.moptf, i
            if want:
                k, _, _ = oracle.extract(im, nfeatures=1000)
                old = np.stack([k["x"], k["y"]], 1).astype(np.float32)
            else:
                old = nxt
            last = t
        prev = im
        assert got == want, i
        assert np.array_equal(kf.old, old), i
    assert 1 < len(kf.GetAllKF()) < len(frames)
