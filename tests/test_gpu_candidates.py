"""GPU parity of candidate-list matching (rumi_hamming_candidates) and of the matchers built on it: the reference's
SearchForInitialization and three SearchByProjection overloads (map points, last frame, key frame), and the descriptor-based key-point association of a submap merge
(40 front + 40 back key frames).  The oracle side is pinned to the unmodified reference functions
(tests/test_ref_frame_pin.py)."""
import numpy as np
import pytest

from rumi_slam_b200.synth import motion_sequence, synthetic_frame

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from oracle import match_oracle
    match_oracle.build()
    return match_oracle


def frame_pair(oracle, seed, shift=(3.0, -2.0)):
    seq = motion_sequence(2, 640, 480, seed=seed, vx=shift[0], vy=shift[1])
    k1, d1, _ = oracle.extract(seq[0])
    k2, d2, _ = oracle.extract(seq[1])
    return k1, d1, k2, d2


def test_candidate_distances_and_top2(oracle, M):
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, 11)
    bounds = (0, 0, 640, 480)
    off, idx = M.candidate_lists(k2, bounds, np.stack([k1["x"], k1["y"]], 1), 40.0)
    m = ORBmatcher()
    dist, (i1, e1, i2, e2) = m.candidates(d1, d2, off, idx, top2=True)
    assert len(dist) == off[-1] > 10 * len(k1)
    for q in range(len(k1)):
        lst = idx[off[q]:off[q + 1]]
        ref = np.array([oracle.descriptor_distance(d1[q], d2[j]) for j in lst], np.int64)
        assert np.array_equal(dist[off[q]:off[q + 1]], ref)
        b1, b2, bi, bi2 = 256, 256, -1, -1                         # the reference scan incl. the second-best index
        for j, d in zip(lst, ref):
            if d < b1:
                b2, bi2, b1, bi = b1, bi, d, j
            elif d < b2:
                b2, bi2 = d, j
        assert (i1[q], e1[q], e2[q]) == (bi, b1, b2) and (i2[q] == bi2 or e2[q] == 256)
    # empty lists, a single candidate, duplicated candidates (earliest list entry wins)
    off2 = np.array([0, 0, 1, 4], np.int32)
    idx2 = np.array([5, 7, 7, 3], np.int32)
    T = d2.copy(); T[3] = T[7]
    dist, (i1, e1, i2, e2) = m.candidates(d1[:3], T, off2, idx2, top2=True)
    assert (i1[0], e1[0], i2[0], e2[0]) == (-1, 256, -1, 256)
    assert i1[1] == 5 and i2[1] == -1 and e2[1] == 256
    assert i1[2] == 7 and i2[2] == 7 and e1[2] == e2[2]


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,window,ratio", [(2, 100, 0.9), (3, 30, 0.9), (4, 100, 0.6)])
def test_search_for_initialization_matches_oracle(oracle, M, seed, window, ratio, check_ori):
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, p = ORBmatcher(ratio, check_ori).SearchForInitialization(k1, d1, k2, d2, (0, 0, 640, 480), prev, window)
    rn, rm, rp = M.search_for_initialization(k1, d1, k2, d2, (0, 0, 640, 480), prev, window, ratio, check_ori)
    assert n == rn and np.array_equal(m12, rm) and np.array_equal(p, rp)
    assert rn > (20 if ratio > 0.8 else 0)


def _projection_case(oracle, seed, stereo):
    """Last frame / key frame = frame 1 with map points projected to where its features moved (+ noise), see
    tests/test_ref_frame_pin.py::_last_frame_case (same construction; the oracle side is pinned there)."""
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    rng = np.random.default_rng(seed)
    n1, n2 = len(k1), len(k2)
    uv = np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 1.5, (n1, 2)).astype(np.float32)
    uv[rng.random(n1) < 0.02] += np.float32(700.0)
    depth = rng.uniform(0.5, 20.0, n1).astype(np.float32)
    depth[rng.random(n1) < 0.03] *= np.float32(-1.0)
    valid = (rng.random(n1) < 0.85).astype(np.uint8)
    has_obs = (rng.random(n1) < 0.8).astype(np.uint8)
    occupied = (rng.random(n2) < 0.05).astype(np.uint8)
    u_right = np.where(rng.random(n2) < 0.7, k2["x"] - 40.0 / rng.uniform(0.5, 20.0, n2), -1.0).astype(np.float32) if stereo else None
    return k1, d1, k2, d2, uv, depth, valid, has_obs, occupied, u_right


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,mode", [(11, 15.0, "mono"), (12, 7.0, "mono"), (13, 15.0, "stereo"), (14, 15.0, "forward"),
                                          (15, 15.0, "backward")])
def test_search_by_projection_last_frame_matches_oracle(oracle, M, seed, th, mode, check_ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (R/lib_src/ORBmatcher.cc:1498-1684)."""
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2, uv, depth, valid, has_obs, occupied, u_right = _projection_case(oracle, seed, mode != "mono")
    sf = oracle.tables()["scale"]
    invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
    fwd, bwd = mode == "forward", mode == "backward"
    n, cm = ORBmatcher(0.9, check_ori).SearchByProjectionLastFrame(k2, d2, sf, (0, 0, 640, 480), valid, uv, invz, k1["octave"],
                                                                   k1["angle"], d1, has_obs, th, u_right, occupied, 40.0, fwd, bwd)
    rn, rcm = M.search_by_projection_last(k2, d2, sf, (0, 0, 640, 480), valid, uv, invz, k1["octave"], k1["angle"], d1, has_obs,
                                          th, u_right, occupied, 40.0, fwd, bwd, check_ori)
    assert n == rn and np.array_equal(cm, rcm)
    assert rn > 100


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,orb_dist", [(21, 10.0, 100), (22, 3.0, 64)])
def test_search_by_projection_keyframe_matches_oracle(oracle, M, seed, th, orb_dist, check_ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (R/lib_src/ORBmatcher.cc:1685-1794)."""
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2, uv, depth, valid, _, occupied, _ = _projection_case(oracle, seed, False)
    rng = np.random.default_rng(seed + 100)
    n1 = len(k1)
    sf = oracle.tables()["scale"]
    level = np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32)
    dist3d = np.abs(depth)
    min_d = rng.uniform(0.0, 6.0, n1).astype(np.float32)
    max_d = (min_d + rng.uniform(0.0, 30.0, n1)).astype(np.float32)
    n, cm = ORBmatcher(0.9, check_ori).SearchByProjectionKeyFrame(k2, d2, sf, (0, 0, 640, 480), valid, uv, dist3d, min_d, max_d,
                                                                  level, k1["angle"], d1, th, orb_dist, occupied)
    rn, rcm = M.search_by_projection_kf(k2, d2, sf, (0, 0, 640, 480), valid, uv, dist3d, min_d, max_d, level, k1["angle"], d1,
                                        th, orb_dist, occupied, check_ori)
    assert n == rn and np.array_equal(cm, rcm)
    assert rn > 30


@pytest.mark.parametrize("seed,th,stereo", [(31, 3.0, False), (32, 3.0, True), (33, 5.0, True)])
def test_fuse_search_matches_oracle(oracle, M, seed, th, stereo):
    """Matching core of ORBmatcher::Fuse(pKF, vpMapPoints, th) (R/lib_src/ORBmatcher.cc:1015-1147); the oracle side is pinned to
    the reference function in tests/test_ref_frame_pin.py::test_fuse_search_equals_reference (same construction)."""
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    rng = np.random.default_rng(seed + 200)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 0.8, (n1, 2)).astype(f32))
    uv[rng.random(n1) < 0.03] += f32(700.0)
    z = rng.uniform(0.5, 20.0, n1).astype(f32)
    tab = oracle.tables()
    sf, inv_sigma2 = tab["scale"], (1.0 / (tab["scale"].astype(np.float64) ** 2)).astype(f32)
    valid = rng.random(n1) < 0.8
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    bf = 40.0
    u_right = (np.where(rng.random(n2) < 0.6, k2["x"] - bf / rng.uniform(2.0, 20.0, n2), -1.0) if stereo else np.full(n2, -1.0)).astype(f32)
    dist3d = np.sqrt(((uv[:, 0] * uv[:, 0] + uv[:, 1] * uv[:, 1]).astype(f32) + z * z).astype(f32)).astype(f32)
    ur = (uv[:, 0] - (f32(bf) * (f32(1.0) / z).astype(f32)).astype(f32)).astype(f32)
    min_d = (dist3d * rng.choice([0.5, 0.9, 1.01], n1)).astype(f32)
    max_d = (dist3d * rng.choice([0.99, 1.1, 2.0], n1)).astype(f32)
    args = (k2, d2, sf, inv_sigma2, (0, 0, 640, 480), u_right, valid, uv, ur, dist3d, min_d, max_d, level, d1, th)
    n, best, bdist = ORBmatcher().FuseSearch(*args)
    rn, rbest, rdist = M.fuse_search(*args)
    assert n == rn and np.array_equal(best, rbest) and np.array_equal(bdist, rdist)
    assert rn > 40


@pytest.mark.parametrize("seed,th", [(71, 7.5), (72, 3.0)])
def test_search_by_sim3_matches_oracle(oracle, M, seed, th):
    """ORBmatcher::SearchBySim3 (R/lib_src/ORBmatcher.cc:1293-1497); oracle pinned in tests/test_ref_frame_pin.py."""
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    rng = np.random.default_rng(seed + 400)
    f32 = np.float32

    def side(ka, dx, dy):
        n = len(ka)
        uv = (np.stack([ka["x"] + dx, ka["y"] + dy], 1).astype(f32) + rng.normal(0, 0.8, (n, 2)).astype(f32))
        uv[rng.random(n) < 0.03] += f32(700.0)
        level = np.clip(ka["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
        dist = rng.uniform(1.0, 100.0, n).astype(f32)
        mn = rng.uniform(0.0, 20.0, n).astype(f32)
        return rng.random(n) < 0.8, uv, dist, mn, (mn + rng.uniform(10.0, 500.0, n)).astype(f32), level
    a, b = side(k1, -3.0, 2.0), side(k2, 3.0, -2.0)
    sf = oracle.tables()["scale"]
    n, m = ORBmatcher().SearchBySim3(k1, d1, k2, d2, sf, (0, 0, 640, 480), *a, *b, th)
    rn, rm = M.search_by_sim3(k1, d1, k2, d2, sf, (0, 0, 640, 480), *a, *b, th)
    assert n == rn and np.array_equal(m, rm) and rn > 30


@pytest.mark.parametrize("seed,th,ratio", [(51, 3, 1.0), (53, 10, 0.8), (54, 30, 1.5)])
def test_sim3_matchers_match_oracle(oracle, M, seed, th, ratio):
    """ORBmatcher::SearchByProjection(pKF, Scw, ...) (R/lib_src/ORBmatcher.cc:372-580) and Fuse(pKF, Scw, ...) (:1182-1292); the
    oracle side is pinned to the reference functions in tests/test_ref_frame_pin.py (..._sim3_equals_reference)."""
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    rng = np.random.default_rng(seed + 300)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 1.0, (n1, 2)).astype(f32))
    uv[rng.random(n1) < 0.03] += f32(700.0)
    valid = rng.random(n1) < 0.8
    occupied = rng.random(n2) < 0.15
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    dist3d = rng.uniform(1.0, 100.0, n1).astype(f32)
    min_d = rng.uniform(0.0, 20.0, n1).astype(f32)
    max_d = (min_d + rng.uniform(10.0, 500.0, n1)).astype(f32)
    sf = oracle.tables()["scale"]
    args = (k2, d2, sf, (0, 0, 640, 480), occupied, valid, uv, dist3d, min_d, max_d, level, d1, th, ratio)
    n, km = ORBmatcher().SearchByProjectionSim3(*args)
    rn, rkm = M.search_by_projection_sim3(*args)
    assert n == rn and np.array_equal(km, rkm) and rn > 50
    nf, best, bd = ORBmatcher().FuseSearchSim3(k2, d2, sf, (0, 0, 640, 480), valid, uv, dist3d, min_d, max_d, level, d1, float(th))
    rnf, rbest, rbd = M.fuse_search(k2, d2, sf, np.zeros(8, f32), (0, 0, 640, 480), np.full(n2, -1.0, f32), valid, uv,
                                    np.zeros(n1, f32), dist3d, min_d, max_d, level, d1, float(th))
    assert nf == rnf and np.array_equal(best, rbest) and np.array_equal(bd, rbd) and rnf > 50


@pytest.mark.parametrize("seed,th,ratio", [(6, 3.0, 0.8), (7, 1.0, 0.8), (8, 5.0, 0.9), (9, 15.0, 0.6)])
def test_search_by_projection_matches_oracle(oracle, M, seed, th, ratio):
    from rumi_slam_b200 import ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, seed)
    rng = np.random.default_rng(seed)
    sf = oracle.tables()["scale"]
    proj = np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 0.7, (len(k1), 2)).astype(np.float32)
    level = np.clip(k1["octave"] + rng.integers(-1, 2, len(k1)), 0, 7).astype(np.int32)
    view_cos = rng.choice([0.9, 0.9985, 1.0], len(k1)).astype(np.float32)
    has_obs = (rng.random(len(k1)) < 0.7).astype(np.uint8)
    n, fm = ORBmatcher(ratio).SearchByProjection(k2, d2, sf, (0, 0, 640, 480), proj, level, view_cos, d1, has_obs, th)
    rn, rfm = M.search_by_projection(k2, d2, sf, (0, 0, 640, 480), proj, level, view_cos, d1, has_obs, th, ratio)
    assert n == rn and np.array_equal(fm, rfm) and rn > 50


def test_associate_submap_40_plus_40_keyframes(oracle, M):
    """A synthetic submap merge: 40 front key frames, 40 back key frames observing the same places (slightly shifted views,
    independent noise), cloud key points WITHOUT descriptors (as in the reference).  AssociateSubmap == the oracle
    composition (CloudFrameComputeDescriptors -> top-2 scan -> SearchByBoW acceptance) pair by pair, and it finds most of
    what the reference's pixel-distance association (CloudMerging.cc:503-551) finds when the views coincide."""
    from rumi_slam_b200 import ORBextractor, ORBmatcher
    npairs, w, h = 40, 640, 480
    rng = np.random.default_rng(3)
    imgs1, imgs2, keys1, keys2, valid1, valid2 = [], [], [], [], [], []
    for p in range(npairs):
        seq = motion_sequence(2, w, h, seed=700 + p, vx=0.4, vy=-0.3, noise=2)
        k = oracle.extract(seq[0])[0]
        k = k[k["octave"] == 0].copy()                        # cloud key points live in level-0 coordinates
        k2 = k.copy()
        k2["x"] = np.rint(k["x"] - 0.4); k2["y"] = np.rint(k["y"] + 0.3)        # the same places in the second view
        keep = rng.random(len(k2)) < 0.9                      # the back submap misses some of them, and has extra ones
        extra = oracle.extract(seq[1])[0]
        extra = extra[extra["octave"] == 0][:40]
        k2 = np.concatenate([k2[keep], extra])
        k2 = k2[rng.permutation(len(k2))]
        imgs1.append(seq[0]); imgs2.append(seq[1]); keys1.append(k); keys2.append(k2)
        valid1.append(rng.random(len(k)) < 0.9); valid2.append(rng.random(len(k2)) < 0.9)
    imgs1, imgs2 = np.stack(imgs1), np.stack(imgs2)
    ex, m = ORBextractor(1000, 1.2, 8, 20, 7), ORBmatcher(0.75)
    match12, counts = m.AssociateSubmap(ex, imgs1, keys1, valid1, imgs2, keys2, valid2)
    total, agree, pix = 0, 0, 0
    for p in range(npairs):
        d1 = oracle.describe(imgs1[p], keys1[p])[1]
        d2 = oracle.describe(imgs2[p], keys2[p])[1]
        s1, s2 = np.flatnonzero(valid1[p]), np.flatnonzero(valid2[p])
        i1, e1, e2 = oracle.hamming_top2(d1[s1], d2[s2])
        ok = m.accept_bow(e1, e2) & (i1 >= 0)
        want = np.full(len(keys1[p]), -1, np.int32)
        want[s1[ok]] = s2[i1[ok]]
        assert np.array_equal(match12[p], want), p
        assert counts[p] == int(ok.sum())
        total += counts[p]
        npx, px = M.associate_pixels(keys1[p], valid1[p], keys2[p], valid2[p], (0, 0, w, h), 3.0)
        both = (px >= 0) & (want >= 0)
        pix += npx
        agree += int((px[both] == want[both]).sum())
    assert total > 40 * 100 and pix > 40 * 100
    assert agree > 0.6 * pix                                  # same physical points found by descriptors alone


def test_projection_matchers_edge_cases(oracle, M):
    """No usable map point, projections outside the image, an empty current frame: every projection matcher returns 0
    matches and all -1, like the reference loops that `continue` on every point."""
    from rumi_slam_b200 import KP_DTYPE, ORBmatcher
    k1, d1, k2, d2 = frame_pair(oracle, 5)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    sf = oracle.tables()["scale"]
    uv = np.stack([k1["x"], k1["y"]], 1).astype(f32)
    ones, zeros = np.ones(n1, f32), np.zeros(n1, f32)
    big = np.full(n1, 1e9, f32)
    m = ORBmatcher(0.9, True)
    none = np.zeros(n1, bool)
    outside = uv + f32(5000.0)
    for valid, pts in ((none, uv), (~none, outside)):
        n, cm = m.SearchByProjectionLastFrame(k2, d2, sf, (0, 0, 640, 480), valid, pts, ones, k1["octave"], k1["angle"], d1, ~none)
        assert n == 0 and np.all(cm == -1)
        n, cm = m.SearchByProjectionKeyFrame(k2, d2, sf, (0, 0, 640, 480), valid, pts, ones, zeros, big, k1["octave"], k1["angle"], d1)
        assert n == 0 and np.all(cm == -1)
        n, km = m.SearchByProjectionSim3(k2, d2, sf, (0, 0, 640, 480), np.zeros(n2, bool), valid, pts, ones, zeros, big, k1["octave"], d1)
        assert n == 0 and np.all(km == -1)
        n, best, bd = m.FuseSearchSim3(k2, d2, sf, (0, 0, 640, 480), valid, pts, ones, zeros, big, k1["octave"], d1)
        assert n == 0 and np.all(best == -1) and np.all(bd == 256)
    # every feature of the current frame already taken
    n, cm = m.SearchByProjectionKeyFrame(k2, d2, sf, (0, 0, 640, 480), ~none, uv, ones, zeros, big, k1["octave"], k1["angle"], d1,
                                         occupied=np.ones(n2, bool))
    assert n == 0 and np.all(cm == -1)
    # an empty current frame / key frame
    e_k, e_d = np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
    n, cm = m.SearchByProjectionLastFrame(e_k, e_d, sf, (0, 0, 640, 480), ~none, uv, ones, k1["octave"], k1["angle"], d1, ~none)
    assert n == 0 and len(cm) == 0
    n, best, bd = m.FuseSearchSim3(e_k, e_d, sf, (0, 0, 640, 480), ~none, uv, ones, zeros, big, k1["octave"], d1)
    assert n == 0 and np.all(best == -1)
