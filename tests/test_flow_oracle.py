"""The optical-flow oracle (oracle/flow_oracle.cpp: restatement of cv::calcOpticalFlowPyrLK as KFDSample::Step calls
it) pinned against OpenCV: frozen cv2 4.13.0 outputs in tests/golden/flow_kats.npz (generator:
tools/gen_golden_flow.py) and, where cv2 is importable, live.

Integer stages (pyrDown, Scharr) must be bit-exact.  The tracker's normal equations are sums of integer products that
OpenCV accumulates in a build-dependent float order; the oracle uses the exact integer sum (OpenCV's NEON variant), so
positions are compared with a tolerance: every status flag equal, |position difference| <= 0.03 px (one
termination-threshold flip), median <= 1e-3 px, err within 0.05."""
import os

import numpy as np
import pytest

from oracle import flow_oracle as fo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_kats.npz")
POS_TOL, POS_MEDIAN_TOL, ERR_TOL = 0.03, 1e-3, 0.05


def _check_against_cv(on, os_, oe, cn, cs, ce):
    assert np.array_equal(os_, cs), "status flags differ from OpenCV"
    good = cs == 1
    d = np.abs(on - cn).max(1)[good]
    assert d.max() <= POS_TOL, d.max()
    assert np.median(d) <= POS_MEDIAN_TOL
    assert np.abs(oe - ce)[good].max() <= ERR_TOL


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_oracle_vs_frozen_cv2(case):
    g = np.load(GOLD)
    prev, nxt, pts = g[case + "_prev"], g[case + "_next"], g[case + "_pts"]
    l1 = fo.pyr_down(prev)
    assert np.array_equal(l1, g[case + "_pyr1"])
    assert np.array_equal(fo.pyr_down(l1), g[case + "_pyr2"])
    if case + "_scharr" in g:
        assert np.array_equal(fo.scharr(prev), g[case + "_scharr"])
    on, os_, oe = fo.lk(prev, nxt, pts)
    _check_against_cv(on, os_, oe, g[case + "_cv_next"], g[case + "_cv_status"], g[case + "_cv_err"])


def test_oracle_vs_cv2_live():
    cv2 = pytest.importorskip("cv2")
    from rumi_slam_b200.synth import synthetic_batch
    cv2.setNumThreads(1)
    rng = np.random.default_rng(5)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.03)
    for i, (w, h) in enumerate([(640, 480), (333, 257)]):
        prev = synthetic_batch(1, w, h, seed0=20 + i)[0]
        M = np.array([[1.008, 0.01, 3.3 - i], [-0.008, 0.996, -2.1 + i]], np.float32)
        nxt = cv2.warpAffine(prev, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        nxt = np.clip(nxt.astype(int) + rng.integers(-3, 4, nxt.shape), 0, 255).astype(np.uint8)
        a = prev
        for _ in range(3):
            b = cv2.pyrDown(a)
            assert np.array_equal(fo.pyr_down(a), b)
            a = b
        sch = np.stack([cv2.Scharr(prev, cv2.CV_16S, 1, 0), cv2.Scharr(prev, cv2.CV_16S, 0, 1)], -1)
        assert np.array_equal(fo.scharr(prev), sch)
        pts = np.stack([rng.uniform(0, w - 1, 600), rng.uniform(0, h - 1, 600)], 1).astype(np.float32)
        cn, cs, ce = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, winSize=(31, 31), maxLevel=2,
                                              criteria=crit)
        ce = ce.ravel().copy()
        ce[cs.ravel() == 0] = 0
        on, os_, oe = fo.lk(prev, nxt, pts)
        _check_against_cv(on, os_, oe, cn.reshape(-1, 2), cs.ravel(), ce)


def test_oracle_edge_cases():
    g = np.load(GOLD)
    prev, nxt = g["b_prev"], g["b_next"]
    n, s, e = fo.lk(prev, nxt, np.zeros((0, 2), np.float32))
    assert n.shape == (0, 2) and s.shape == (0,)
    # a point far outside the image is reported lost, a constant image has no texture (minEig test)
    n, s, e = fo.lk(prev, nxt, np.array([[-100.0, -100.0], [1e4, 1e4]], np.float32))
    assert list(s) == [0, 0]
    flat = np.full_like(prev, 90)
    n, s, e = fo.lk(flat, flat, np.array([[80.0, 60.0]], np.float32))
    assert list(s) == [0]
    # identical frames: zero flow
    pts = np.array([[40.5, 30.25], [100.0, 90.0]], np.float32)
    n, s, e = fo.lk(prev, prev, pts)
    assert list(s) == [1, 1] and np.abs(n - pts).max() < 1e-3 and e.max() == 0


def test_selector_host_logic():
    """SelectGoodPts + Calmoptflmag + PD::update: the product's host mirror (rumi_slam_b200/flow.py, pure float32
    host logic, no device needed) equals the oracle's C restatement."""
    from rumi_slam_b200.flow import PD, mean_flow_magnitude
    rng = np.random.default_rng(9)
    old = rng.uniform(0, 600, (500, 2)).astype(np.float32)
    nxt = (old + rng.normal(0, 3, old.shape)).astype(np.float32)
    st = (rng.uniform(size=500) < 0.9).astype(np.uint8)
    want, ngood = fo.mean_magnitude(old, nxt, st)
    got = mean_flow_magnitude(old[st == 1], nxt[st == 1])
    assert ngood == int(st.sum()) and np.float32(want) == got
    a, b = fo.PD(0.8, 0.005, 10.0), PD(0.8, 0.005)
    b.setpoint = np.float32(10.0)
    for v, ts in [(3.2, 0.033), (7.9, 0.034), (12.5, 0.03), (300.0, 0.05), (0.5, 1e-3)]:
        assert np.float32(a.update(v, ts)) == b.update(v, ts)
