"""Pins the oracle's restatements of Frame::ComputeStereoMatches, MapPoint::ComputeDistinctiveDescriptors and both
ORBmatcher::SearchByBoW overloads against the UNMODIFIED reference functions (oracle/_ref/librefframe.so: the function
definitions cut out of /root/reference's Frame.cc / MapPoint.cc / ORBmatcher.cc at build time and compiled over class
stand-ins, oracle/ref_frame_shim.cpp).  Skipped where neither /root/reference nor the prebuilt library exists."""
import numpy as np
import pytest

from rumi_slam_b200.synth import descriptors_near_vocabulary, stereo_pair, synthetic_vocabulary


@pytest.fixture(scope="module")
def rf():
    from oracle import ref_frame_lib
    if not ref_frame_lib.available():
        pytest.skip("oracle/_ref/librefframe.so not available")
    return ref_frame_lib


@pytest.fixture(scope="module")
def B():
    from oracle import bow_oracle
    bow_oracle.build()
    return bow_oracle


def test_descriptor_distance_is_the_references(oracle, rf):
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert rf.descriptor_distance(a, b) == oracle.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())


@pytest.mark.parametrize("w,h,nf,fx,bf,seeds", [(752, 480, 1200, 435.2, 47.9, (1, 2, 3, 4)),
                                                (1241, 376, 2000, 718.856, 386.1448, (5, 6)),
                                                (640, 480, 1000, 517.3, 40.0, (7,))])
def test_compute_stereo_matches_equals_reference(oracle, rf, w, h, nf, fx, bf, seeds):
    """mvuRight / mvDepth of the oracle == the reference's own Frame::ComputeStereoMatches (R/lib_src/Frame.cc:828-985),
    float for float, on the BASELINE stereo shapes (EuRoC, KITTI) -- best-1 band search, SAD slide, parabola, median cut."""
    tb = oracle.tables(nf)
    total = 0
    for s in seeds:
        left, right = stereo_pair(s, w, h)
        lk, ld, _ = oracle.extract(left, nfeatures=nf)
        rk, rd, _ = oracle.extract(right, nfeatures=nf)
        u, d, n = oracle.stereo_match(left, right, lk, ld, rk, rd, bf, bf / fx)
        ru, rdp, rn = rf.stereo_match(oracle.pyramid(left), oracle.pyramid(right), lk, ld, rk, rd, tb["scale"],
                                      tb["inv_scale"], bf, bf / fx)
        assert n == rn and np.array_equal(u, ru) and np.array_equal(d, rdp), (s, n, rn)
        total += n
    assert total > 300 * len(seeds)


def test_distinctive_descriptor_equals_reference(rf, B):
    """The descriptor MapPoint::ComputeDistinctiveDescriptors (R/lib_src/MapPoint.cc:353-426) stores == the row the oracle
    picks, for 1..100 observations incl. ties (identical observations, two equally central rows)."""
    rng = np.random.default_rng(3)
    sizes = np.concatenate([rng.integers(1, 12, 300), rng.integers(30, 100, 20), [1, 2, 3, 33, 64, 65]])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    centre = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.concatenate([centre[p] ^ np.packbits(rng.random((sizes[p], 256)) < 0.08, axis=1) for p in range(len(sizes))])
    desc[off[5]:off[6]] = desc[off[5]]                       # all observations identical: first wins
    desc[off[7] + 1] = desc[off[7]]                          # duplicated rows inside a point
    best, med = B.distinctive(desc, off)
    for p in range(len(sizes)):
        d = desc[off[p]:off[p + 1]]
        assert np.array_equal(rf.distinctive(d), d[best[p]]), p


def _scene(B, seed, nk=700, nf=800, flip=0.03):
    k, L = 10, 4
    par, leaf, vdesc, w = synthetic_vocabulary(k, L, seed=11)
    V = B.Vocabulary(k, L, par, leaf, vdesc, w)
    rng = np.random.default_rng(seed)
    dk = descriptors_near_vocabulary(vdesc, leaf, nk, 20 + seed, flip=0.04)
    src = rng.integers(0, nk, nf - 150)
    df = np.concatenate([dk[src] ^ np.packbits(rng.random((nf - 150, 256)) < flip, axis=1),
                         descriptors_near_vocabulary(vdesc, leaf, 150, 30 + seed)])
    perm = rng.permutation(nf)
    df = df[perm]
    ak = (rng.random(nk) * 360).astype(np.float32)
    af = (rng.random(nf) * 360).astype(np.float32)
    af[np.argsort(perm)[:nf - 150]] = ((ak[src] - 15 + rng.normal(0, 2, nf - 150)) % 360).astype(np.float32)

    def featvec(desc):
        _, weight, node = V.transform(desc, 2)
        fv = {}
        for i in range(len(desc)):
            if weight[i] > 0:
                fv.setdefault(int(node[i]), []).append(i)
        return fv
    return dk, ak, featvec(dk), df, af, featvec(df), rng


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_search_by_bow_equals_reference(rf, B, seed, check_ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (R/lib_src/ORBmatcher.cc:198-370): the assignment of every frame
    feature equals the reference's -- order-dependent 'already matched' skip and rotation histogram included."""
    dk, ak, fvk, df, af, fvf, rng = _scene(B, seed)
    valid = (rng.random(len(dk)) < 0.8).astype(np.uint8)
    for ratio in (0.6, 0.75, 0.9):
        n, m = B.search_by_bow(dk, ak, valid, fvk, df, af, fvf, nnratio=ratio, check_ori=check_ori)
        rn, rm = rf.search_by_bow(dk, ak, valid, fvk, df, af, fvf, nnratio=ratio, check_ori=check_ori)
        assert n == rn and np.array_equal(m, rm), (seed, ratio)
    assert rn > 100


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,frac", [(7, 0.5), (8, 0.7), (9, 0.3)])
def test_search_by_bow_stereo_fisheye_branch_equals_reference(rf, B, seed, frac, check_ori):
    """The F.Nleft != -1 branches of SearchByBoW(KeyFrame*, Frame&, ...) (R/lib_src/ORBmatcher.cc:258-340): features
    [0, Nleft) are the left, [Nleft, N) the right fisheye camera; separate best / second best per side, the right match needs
    no ratio test but only counts when the left one passed TH_LOW."""
    dk, ak, fvk, df, af, fvf, rng = _scene(B, seed)
    valid = (rng.random(len(dk)) < 0.8).astype(np.uint8)
    n_left = int(len(df) * frac)
    for ratio in (0.6, 0.9):
        n, m = B.search_by_bow(dk, ak, valid, fvk, df, af, fvf, nnratio=ratio, check_ori=check_ori, n_left=n_left)
        rn, rm = rf.search_by_bow(dk, ak, valid, fvk, df, af, fvf, nnratio=ratio, check_ori=check_ori, n_left=n_left)
        assert n == rn and np.array_equal(m, rm), (seed, ratio)
    assert rn > 100 and (rm[n_left:] >= 0).sum() > 20


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed", [4, 5, 6])
def test_search_by_bow_keyframes_equals_reference(rf, B, seed, check_ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...) (R/lib_src/ORBmatcher.cc:682-804), strict '<' TH_LOW."""
    d1, a1, fv1, d2, a2, fv2, rng = _scene(B, seed)
    v1 = (rng.random(len(d1)) < 0.85).astype(np.uint8)
    v2 = (rng.random(len(d2)) < 0.85).astype(np.uint8)
    for ratio in (0.7, 0.8):
        n, m = B.search_by_bow_kf(d1, a1, v1, fv1, d2, a2, v2, fv2, nnratio=ratio, check_ori=check_ori)
        rn, rm = rf.search_by_bow_kf(d1, a1, v1, fv1, d2, a2, v2, fv2, nnratio=ratio, check_ori=check_ori)
        assert n == rn and np.array_equal(m, rm), (seed, ratio)
    assert rn > 100


# ---- the feature grid and the candidate-list matchers (oracle/match_oracle.cpp) ----
@pytest.fixture(scope="module")
def M():
    from oracle import match_oracle
    match_oracle.build()
    return match_oracle


def _frame_pair(oracle, seed, shift=(3.0, -2.0)):
    from rumi_slam_b200.synth import motion_sequence
    seq = motion_sequence(2, 640, 480, seed=seed, vx=shift[0], vy=shift[1])
    k1, d1, _ = oracle.extract(seq[0])
    k2, d2, _ = oracle.extract(seq[1])
    return k1, d1, k2, d2


def test_features_in_area_equals_reference(oracle, rf, M):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (R/lib_src/Frame.cc:441-466, 695-767): the same
    indices in the same order, with and without the level filter, near and beyond the image border."""
    k1, _, _, _ = _frame_pair(oracle, 1)
    bounds = (0, 0, 640, 480)
    rng = np.random.default_rng(0)
    for _ in range(300):
        x, y = rng.uniform(-40, 680), rng.uniform(-40, 520)
        r = float(rng.choice([1.5, 3, 10, 25, 100]))
        lv = int(rng.integers(-1, 8))
        a = M.features_in_area(k1, bounds, x, y, r, lv - 1 if lv >= 0 else -1, lv)
        b = rf.features_in_area(k1, bounds, x, y, r, lv - 1 if lv >= 0 else -1, lv)
        assert np.array_equal(a, b)
    off, idx = M.candidate_lists(k1, bounds, np.stack([k1["x"], k1["y"]], 1), 10.0, 0, 0)
    for q in (0, 17, len(k1) - 1):
        assert np.array_equal(idx[off[q]:off[q + 1]], rf.features_in_area(k1, bounds, k1["x"][q], k1["y"][q], 10.0, 0, 0))


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,window,ratio", [(2, 100, 0.9), (3, 30, 0.9), (4, 100, 0.6), (5, 10, 0.9)])
def test_search_for_initialization_equals_reference(oracle, rf, M, seed, window, ratio, check_ori):
    """ORBmatcher::SearchForInitialization (R/lib_src/ORBmatcher.cc:581-680): vnMatches12, the updated vbPrevMatched and
    the match count -- the vMatchedDistance skip, the re-assignment of a taken feature and the rotation histogram."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m, p = M.search_for_initialization(k1, d1, k2, d2, (0, 0, 640, 480), prev, window, ratio, check_ori)
    rn, rm, rp = rf.search_for_initialization(k1, d1, k2, d2, (0, 0, 640, 480), prev, window, ratio, check_ori)
    assert n == rn and np.array_equal(m, rm) and np.array_equal(p, rp)
    assert rn > (20 if window >= 30 and ratio > 0.8 else 0)


@pytest.mark.parametrize("seed,th,ratio", [(6, 3.0, 0.8), (7, 1.0, 0.8), (8, 5.0, 0.9), (9, 15.0, 0.6)])
def test_search_by_projection_equals_reference(oracle, rf, M, seed, th, ratio):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (R/lib_src/ORBmatcher.cc:39-118) on a mono frame: level
    window, level-aware ratio test and the order-dependent 'already has a map point with observations' skip."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    rng = np.random.default_rng(seed)
    sf = oracle.tables()["scale"]
    # map points = the features of frame 1, projected into frame 2 with the known image motion plus noise
    proj = np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 0.7, (len(k1), 2)).astype(np.float32)
    level = np.clip(k1["octave"] + rng.integers(-1, 2, len(k1)), 0, 7).astype(np.int32)
    view_cos = rng.choice([0.9, 0.9985, 1.0], len(k1)).astype(np.float32)
    has_obs = (rng.random(len(k1)) < 0.7).astype(np.uint8)
    n, fm = M.search_by_projection(k2, d2, sf, (0, 0, 640, 480), proj, level, view_cos, d1, has_obs, th, ratio)
    rn, rfm = rf.search_by_projection(k2, d2, sf, (0, 0, 640, 480), proj, level, view_cos, d1, has_obs, th, ratio)
    assert n == rn and np.array_equal(fm, rfm)
    assert rn > 50


def _right_camera(oracle, seed, dx=-6):
    """Key points / descriptors of a 'right camera' that sees the second frame of _frame_pair(seed) shifted by dx pixels."""
    from rumi_slam_b200.synth import motion_sequence
    img = np.roll(motion_sequence(2, 640, 480, seed=seed, vx=3.0, vy=-2.0)[1], dx, axis=1)
    kR, dR, _ = oracle.extract(np.ascontiguousarray(img))
    return kR, dR


def _last_frame_fisheye_case(oracle, seed):
    """_last_frame_case with a second extraction playing the right camera of the current frame; the right camera sees every
    point 6 px further left."""
    k1, d1, k2, d2, uv, depth, valid, has_obs, _, _ = _last_frame_case(oracle, seed, False)
    kR, dR = _right_camera(oracle, seed)
    rng = np.random.default_rng(seed + 500)
    occupied = (rng.random(len(k2) + len(kR)) < 0.05).astype(np.uint8)
    return k1, d1, k2, kR, np.concatenate([d2, dR]), uv, depth, valid, has_obs, occupied, (-6.0, 0.0)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,direction", [(51, 15.0, "none"), (52, 7.0, "none"), (53, 15.0, "forward"), (54, 15.0, "backward")])
def test_search_by_projection_last_frame_fisheye_equals_reference(oracle, rf, M, seed, th, direction, check_ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, false) with CurrentFrame.Nleft != -1 (R/lib_src/ORBmatcher.cc:
    1498-1684 incl. :1602-1656): the second, right-camera search of every last-frame point and its votes in the shared
    rotation histogram."""
    k1, d1, kC, kR, dC, uv, depth, valid, has_obs, occupied, shift = _last_frame_fisheye_case(oracle, seed)
    sf = oracle.tables()["scale"]
    fw, bw = direction == "forward", direction == "backward"
    rn, rcm = rf.search_by_projection_last_fisheye(kC, kR, dC, sf, (0, 0, 640, 480), valid, uv, shift, depth, k1["octave"],
                                                   k1["angle"], d1, has_obs, th, occupied, fw, bw, check_ori)
    invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
    uvr = (uv + np.asarray(shift, np.float32)).astype(np.float32)
    n, cm = M.search_by_projection_last_fisheye(kC, kR, dC, sf, (0, 0, 640, 480), valid, uv, uvr, invz, k1["octave"], k1["angle"],
                                                d1, has_obs, th, occupied, fw, bw, check_ori)
    assert n == rn and np.array_equal(cm, rcm)
    nL = len(kC)
    assert (rcm[:nL] >= 0).sum() > 20 and (rcm[nL:] >= 0).sum() > 20


def _local_points_case(oracle, seed, mode):
    """Tracking::SearchLocalPoints as the reference function sees it.  Frame 2 = the current frame; map points = the
    features of frame 1 projected with the known image motion plus noise.  mode 'occupied': the frame already holds matches
    with observations; 'stereo': mvuRight set on most features and mTrackProjXR near / far from it; 'fisheye': a third
    extraction plays the right camera (Nleft = n2), map points are in view of the left, the right or both cameras, some
    left / right features are stereo partners of each other."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    rng = np.random.default_rng(100 + seed)
    n1, n2 = len(k1), len(k2)
    kw = dict(proj=np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 0.7, (n1, 2)).astype(np.float32),
              level=np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32),
              view_cos=rng.choice([0.9, 0.9985, 1.0], n1).astype(np.float32),
              has_obs=(rng.random(n1) < 0.7).astype(np.uint8))
    kL, dF = k2, d2
    if mode in ("occupied", "stereo"):
        kw["occupied"] = (rng.random(n2) < 0.25).astype(np.uint8)
    if mode == "stereo":
        ur = (k2["x"] - rng.uniform(2, 40, n2)).astype(np.float32)
        ur[rng.random(n2) < 0.3] = -1.0                                     # no stereo match for that feature
        kw["u_right"] = ur
        # mTrackProjXR: mostly close to the right coordinate of the feature the point should hit, sometimes far off
        pr = np.zeros((n1, 2), np.float32)
        pr[:, 0] = kw["proj"][:, 0] - rng.uniform(2, 40, n1).astype(np.float32)
        pr[:, 1] = kw["proj"][:, 1]
        kw["proj_r"] = pr
    if mode == "fisheye":
        kR, dR = _right_camera(oracle, seed)
        nR = len(kR)
        kw["kR"] = kR
        dF = np.concatenate([d2, dR])
        kw["in_view"] = (rng.random(n1) < 0.8).astype(np.uint8)
        kw["in_view_r"] = (rng.random(n1) < 0.6).astype(np.uint8)
        kw["proj_r"] = np.stack([k1["x"] - 9.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 0.7, (n1, 2)).astype(np.float32)
        lr = np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32)
        lr[rng.random(n1) < 0.05] = -1                                      # mnTrackScaleLevelR == -1: right half skipped
        kw["level_r"] = lr
        kw["view_cos_r"] = rng.choice([0.9, 0.9985, 1.0], n1).astype(np.float32)
        # stereo partners: a partial one-to-one map between left and right features
        m = min(n2, nR) // 3
        li, ri = rng.permutation(n2)[:m], rng.permutation(nR)[:m]
        l2r, r2l = np.full(n2, -1, np.int32), np.full(nR, -1, np.int32)
        l2r[li] = ri
        r2l[ri] = li
        kw["l2r"], kw["r2l"] = l2r, r2l
        kw["occupied"] = (rng.random(n2 + nR) < 0.15).astype(np.uint8)
    return kL, dF, d1, kw


@pytest.mark.parametrize("mode", ["occupied", "stereo", "fisheye"])
@pytest.mark.parametrize("seed,th,ratio", [(6, 3.0, 0.8), (7, 1.0, 0.8), (8, 5.0, 0.9), (9, 15.0, 0.6)])
def test_search_by_projection_all_branches_equal_reference(oracle, rf, M, seed, th, ratio, mode):
    """The whole ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (R/lib_src/ORBmatcher.cc:39-189): features that
    already hold a map point with observations (:80-82), the right-image gate of rectified stereo / RGB-D frames (:84-88)
    and both halves of the stereo-fisheye loop with their cross assignments (:114-118, :125-185)."""
    kL, dF, dMP, kw = _local_points_case(oracle, seed, mode)
    sf = oracle.tables()["scale"]
    args = dict(kw)
    proj, level, view_cos, has_obs = args.pop("proj"), args.pop("level"), args.pop("view_cos"), args.pop("has_obs")
    n, fm = M.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio, **args)
    rn, rfm = rf.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio, **args)
    assert n == rn and np.array_equal(fm, rfm)
    assert rn > 40
    if mode == "fisheye":
        nL = len(kL)
        assert (rfm[nL:] >= 0).sum() > 20 and (rfm[:nL] >= 0).sum() > 20          # both cameras received matches
    if mode == "stereo":                                                         # the gate did reject candidates
        n0, _ = rf.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio,
                                           occupied=args["occupied"])
        assert n0 != rn
    # without any of the extras the extended entry equals the mono one
    if mode == "occupied":
        a = M.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio)
        b = M.search_by_projection(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


def _last_frame_case(oracle, seed, stereo):
    """Frame 1 = last frame (its features carry map points), frame 2 = current frame; map point i is projected to where
    frame 1's feature moved (known image motion + noise), with a depth; some features lack a map point / are outliers,
    some map points have no observations, a few projections fall outside the image or behind the camera."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    rng = np.random.default_rng(seed)
    n1, n2 = len(k1), len(k2)
    uv = np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(np.float32) + rng.normal(0, 1.5, (n1, 2)).astype(np.float32)
    uv[rng.random(n1) < 0.02] += np.float32(700.0)                       # out of the image bounds
    depth = rng.uniform(0.5, 20.0, n1).astype(np.float32)
    depth[rng.random(n1) < 0.03] *= np.float32(-1.0)                     # behind the camera: invzc < 0
    valid = (rng.random(n1) < 0.85).astype(np.uint8)
    has_obs = (rng.random(n1) < 0.8).astype(np.uint8)
    occupied = (rng.random(n2) < 0.05).astype(np.uint8)
    u_right = None
    if stereo:
        u_right = np.where(rng.random(n2) < 0.7, k2["x"] - 40.0 / rng.uniform(0.5, 20.0, n2), -1.0).astype(np.float32)
    return k1, d1, k2, d2, uv, depth, valid, has_obs, occupied, u_right


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,mode", [(11, 15.0, "mono"), (12, 7.0, "mono"), (13, 15.0, "stereo"), (14, 15.0, "forward"),
                                          (15, 15.0, "backward"), (16, 30.0, "stereo")])
def test_search_by_projection_last_frame_equals_reference(oracle, rf, M, seed, th, mode, check_ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (R/lib_src/ORBmatcher.cc:1498-1684), the matcher of
    TrackWithMotionModel: level windows for forward / backward / lateral motion, the right-image consistency test, the
    order-dependent 'feature already holds a point with observations' skip, overwriting, the rotation histogram."""
    k1, d1, k2, d2, uv, depth, valid, has_obs, occupied, u_right = _last_frame_case(oracle, seed, mode != "mono")
    sf = oracle.tables()["scale"]
    fwd, bwd = mode == "forward", mode == "backward"
    invz = (1.0 / depth.astype(np.float64)).astype(np.float32)          # const float invzc = 1.0 / x3Dc(2)  (:1527)
    n, cm = M.search_by_projection_last(k2, d2, sf, (0, 0, 640, 480), valid, uv, invz, k1["octave"], k1["angle"], d1, has_obs,
                                        th, u_right, occupied, 40.0, fwd, bwd, check_ori)
    rn, rcm = rf.search_by_projection_last(k2, d2, sf, (0, 0, 640, 480), valid, uv, depth, k1["octave"], k1["angle"], d1,
                                           has_obs, th, u_right, occupied, 40.0, fwd, bwd, check_ori)
    assert n == rn and np.array_equal(cm, rcm)
    assert rn > 100


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,orb_dist", [(21, 10.0, 100), (22, 3.0, 64), (23, 10.0, 50)])
def test_search_by_projection_keyframe_equals_reference(oracle, rf, M, seed, th, orb_dist, check_ori):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (R/lib_src/ORBmatcher.cc:1685-1794), the
    matcher of Relocalization: bad / already-found points, the distance-invariance window, the predicted level window,
    'any map point' occupancy, the rotation histogram."""
    k1, d1, k2, d2, uv, depth, _, _, occupied, _ = _last_frame_case(oracle, seed, False)
    depth = np.abs(depth)
    rng = np.random.default_rng(seed + 100)
    n1 = len(k1)
    sf = oracle.tables()["scale"]
    state = rng.choice([0, 1, 1, 1, 1, 1, 2, 3], n1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32)
    min_d = rng.uniform(0.0, 6.0, n1).astype(np.float32)
    max_d = (min_d + rng.uniform(0.0, 800.0, n1)).astype(np.float32)
    # the reference compares |x3Dw - Ow| of the stand-in world point (uv, depth): take that number from the reference run
    rn, rcm, dist3d = rf.search_by_projection_kf(k2, d2, sf, (0, 0, 640, 480), state, uv, depth, level, min_d, max_d, k1["angle"],
                                                 d1, th, orb_dist, occupied, check_ori)
    n, cm = M.search_by_projection_kf(k2, d2, sf, (0, 0, 640, 480), state == 1, uv, dist3d, min_d, max_d, level, k1["angle"], d1,
                                      th, orb_dist, occupied, check_ori)
    assert n == rn and np.array_equal(cm, rcm)
    assert rn > 30


@pytest.mark.parametrize("seed,th,stereo", [(31, 3.0, False), (32, 3.0, True), (33, 5.0, True), (34, 2.0, False), (35, 4.0, True)])
def test_fuse_search_equals_reference(oracle, rf, M, seed, th, stereo):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (R/lib_src/ORBmatcher.cc:1015-1181) up to the fuse decision: which key-frame
    feature each map point meets -- null / bad / already-observed / behind-camera / outside-image / out-of-range / oblique
    points, the level window, the chi-square reprojection gates (mono 5.99, stereo 7.8), TH_LOW."""
    k1, d1, k2, d2, uv, depth, _, _, _, _ = _last_frame_case(oracle, seed, False)
    rng = np.random.default_rng(seed + 200)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 0.8, (n1, 2)).astype(f32))
    uv[rng.random(n1) < 0.03] += f32(700.0)
    tab = oracle.tables()
    sf, inv_sigma2 = tab["scale"], (1.0 / (tab["scale"].astype(np.float64) ** 2)).astype(f32)
    state = rng.choice([0, 1, 1, 1, 1, 1, 1, 2, 3, 4], n1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    bf = 40.0
    u_right = (np.where(rng.random(n2) < 0.6, k2["x"] - bf / rng.uniform(2.0, 20.0, n2), -1.0) if stereo
               else np.full(n2, -1.0)).astype(f32)
    kf_has = rng.choice([0, 0, 0, 1, 2], n2).astype(np.uint8)
    n_obs = rng.integers(1, 10, n1).astype(np.int32)
    # what the reference computes from the stand-in world point (uv, depth) with Ow = 0: float32, operation by operation
    x, y, z = uv[:, 0], uv[:, 1], depth
    dist3d = np.sqrt(((x * x + y * y).astype(f32) + z * z).astype(f32)).astype(f32)
    invz = (f32(1.0) / z).astype(f32)
    ur = (x - (f32(bf) * invz).astype(f32)).astype(f32)
    min_d = (dist3d * rng.choice([0.5, 0.9, 1.01], n1)).astype(f32)
    max_d = (dist3d * rng.choice([0.99, 1.1, 2.0], n1)).astype(f32)
    rn, rbest = rf.fuse(k2, d2, sf, inv_sigma2, (0, 0, 640, 480), u_right, kf_has, bf, state, uv, depth, min_d, max_d, level, d1,
                        n_obs, th)
    valid = (state == 1) & ~(z < 0)
    n, best, bdist = M.fuse_search(k2, d2, sf, inv_sigma2, (0, 0, 640, 480), u_right, valid, uv, ur, dist3d, min_d, max_d, level,
                                   d1, th)
    assert rn >= 0 and n == rn
    traced = (best < 0) | (kf_has[np.maximum(best, 0)] != 2)         # a bad resident point leaves no trace in the reference run
    assert np.array_equal(best[traced], rbest[traced]) and np.all(rbest[~traced] == -1)
    assert rn > 40 and (best >= 0).sum() == rn


@pytest.mark.parametrize("seed,th", [(36, 3.0), (37, 5.0), (38, 2.0)])
def test_fuse_in_the_right_camera_is_fuse_search_on_the_right_arrays(oracle, rf, M, seed, th):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = true) (R/lib_src/ORBmatcher.cc:1015-1181 with mGridRight / mvKeysRight /
    descriptor rows NLeft + i): the same search as bRight = false on the right camera's arrays -- what the adapters' FuseSearch
    computes when it is handed mvKeysRight, the right rows of mDescriptors and a grid over mvKeysRight; bestIdx + NLeft."""
    k1, d1, k2, d2, uv, depth, _, _, _, _ = _last_frame_case(oracle, seed, False)
    kR, dR = _right_camera(oracle, seed)
    rng = np.random.default_rng(seed + 300)
    n1, nL, nR = len(k1), len(k2), len(kR)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 9.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 0.8, (n1, 2)).astype(f32))   # in the right image
    uv[rng.random(n1) < 0.03] += f32(700.0)
    tab = oracle.tables()
    sf, inv_sigma2 = tab["scale"], (1.0 / (tab["scale"].astype(np.float64) ** 2)).astype(f32)
    state = rng.choice([0, 1, 1, 1, 1, 1, 1, 2, 3, 4], n1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    bf = 40.0
    u_right_all = np.full(nL + nR, -1.0, f32)                     # fisheye frames carry no rectified right coordinate ...
    u_right_all[:nR][rng.random(nR) < 0.1] = f32(5.0)             # ... but the function would read entry [right index] if one were set
    kf_has = rng.choice([0, 0, 0, 1, 2], nL + nR).astype(np.uint8)
    n_obs = rng.integers(1, 10, n1).astype(np.int32)
    x, y, z = uv[:, 0], uv[:, 1], depth
    dist3d = np.sqrt(((x * x + y * y).astype(f32) + z * z).astype(f32)).astype(f32)
    invz = (f32(1.0) / z).astype(f32)
    ur = (x - (f32(bf) * invz).astype(f32)).astype(f32)
    min_d = (dist3d * rng.choice([0.5, 0.9, 1.01], n1)).astype(f32)
    max_d = (dist3d * rng.choice([0.99, 1.1, 2.0], n1)).astype(f32)
    rn, rbest = rf.fuse_right(nL, kR, np.concatenate([d2, dR]), sf, inv_sigma2, (0, 0, 640, 480), u_right_all, kf_has, bf, state, uv,
                              depth, min_d, max_d, level, d1, n_obs, th)
    valid = (state == 1) & ~(z < 0)
    n, best, _ = M.fuse_search(kR, dR, sf, inv_sigma2, (0, 0, 640, 480), u_right_all[:nR], valid, uv, ur, dist3d, min_d, max_d, level,
                               d1, th)
    best = np.where(best >= 0, best + nL, -1)
    assert rn >= 0 and n == rn
    traced = (best < 0) | (kf_has[np.maximum(best, 0)] != 2)
    assert np.array_equal(best[traced], rbest[traced]) and np.all(rbest[~traced] == -1)
    assert rn > 40


def _triangulation_case(B, seed):
    """Two key frames whose descriptors share vocabulary nodes (the SearchByBoW scene), with key points, map-point flags,
    stereo flags, an epipole inside image 2 and a pseudo-random epipolar-constraint table."""
    from rumi_slam_b200 import KP_DTYPE
    d1, a1, fv1, d2, a2, fv2, rng = _scene(B, seed)
    n1, n2 = len(d1), len(d2)

    def kps(n, ang):
        k = np.zeros(n, KP_DTYPE)
        k["x"], k["y"] = rng.uniform(20, 620, n), rng.uniform(20, 460, n)
        k["angle"], k["octave"], k["size"], k["class_id"] = ang, rng.integers(0, 8, n), 31.0, -1
        return k
    k1, k2 = kps(n1, a1), kps(n2, a2)
    has1, has2 = rng.random(n1) < 0.3, rng.random(n2) < 0.3
    ur1 = np.where(rng.random(n1) < 0.5, k1["x"] - 5.0, -1.0).astype(np.float32)
    ur2 = np.where(rng.random(n2) < 0.5, k2["x"] - 5.0, -1.0).astype(np.float32)
    epi = ((np.arange(n1)[:, None] * 31 + np.arange(n2)[None, :] * 17) % 5 != 0).astype(np.uint8)
    return d1, a1, fv1, k1, has1, ur1, d2, a2, fv2, k2, has2, ur2, epi, (320.0, 240.0)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,only_stereo,coarse", [(40, False, False), (41, True, False), (42, False, True), (43, False, False)])
def test_search_for_triangulation_equals_reference(oracle, rf, B, seed, only_stereo, coarse, check_ori):
    """ORBmatcher::SearchForTriangulation (R/lib_src/ORBmatcher.cc:806-1013), the matcher of LocalMapping::CreateNewMapPoints:
    common-node walk, 'already has a map point' skips, stereo-only mode, the running best distance that only an accepted
    (epipolar-consistent) candidate lowers, the epipole exclusion zone for mono pairs, the rotation histogram."""
    d1, a1, fv1, k1, has1, ur1, d2, a2, fv2, k2, has2, ur2, epi, ep = _triangulation_case(B, seed)
    sf = oracle.tables()["scale"]
    rn, rm = rf.search_for_triangulation(k1, d1, has1, ur1, fv1, k2, d2, has2, ur2, fv2, sf, ep, epi, only_stereo, coarse, check_ori)
    n, m = B.search_for_triangulation(d1, a1, has1, ur1 >= 0, fv1, d2, a2, has2, ur2 >= 0, k2["x"], k2["y"], k2["octave"], fv2, sf,
                                      ep, epi, only_stereo, coarse, check_ori)
    assert n == rn and np.array_equal(m, rm)
    assert rn > (20 if only_stereo else 60)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,coarse", [(44, False), (45, True), (46, False)])
def test_search_for_triangulation_on_a_stereo_fisheye_rig(oracle, rf, B, seed, coarse, check_ori):
    """The same function with mpCamera2 set on both key frames (R/lib_src/ORBmatcher.cc:870, :878-880, :908-916): no feature counts
    as stereo, the epipole exclusion zone is off, key points come from mvKeys / mvKeysRight by the flattened index.  That is the
    existing entry point with stereo flags all false, the flattened key points and an epipole nothing comes close to (the
    camera pair of the epipolar test is the caller's predicate either way)."""
    d1, a1, fv1, k1, has1, _, d2, a2, fv2, k2, has2, _, epi, ep = _triangulation_case(B, seed)
    sf = oracle.tables()["scale"]
    nl1, nl2 = len(k1) * 2 // 3, len(k2) // 2
    rn, rm = rf.search_for_triangulation_rig(k1, d1, has1, fv1, nl1, k2, d2, has2, fv2, nl2, sf, ep, epi, False, coarse, check_ori)
    none1, none2 = np.zeros(len(d1), bool), np.zeros(len(d2), bool)
    far = (3.0e9, 3.0e9)
    n, m = B.search_for_triangulation(d1, a1, has1, none1, fv1, d2, a2, has2, none2, k2["x"], k2["y"], k2["octave"], fv2, sf, far,
                                      epi, False, coarse, check_ori)
    assert n == rn and np.array_equal(m, rm)
    assert rn > 60
    # bOnlyStereo on a rig: nothing is stereo, nothing matches
    rn2, rm2 = rf.search_for_triangulation_rig(k1, d1, has1, fv1, nl1, k2, d2, has2, fv2, nl2, sf, ep, epi, True, coarse, check_ori)
    assert rn2 == 0 and np.all(rm2 == -1)


def _sim3_case(oracle, seed):
    """Loop-closing scene: key frame = frame 2, candidate map points = the features of frame 1 projected to where they moved;
    depths are powers of two so that the overload projecting with fx * x / z + cx reproduces the prescribed pixel exactly."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    rng = np.random.default_rng(seed + 300)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 1.0, (n1, 2)).astype(f32))
    uv[rng.random(n1) < 0.03] += f32(700.0)
    depth = rng.choice([0.5, 1.0, 2.0, 4.0, 8.0, -1.0], n1, p=[0.2, 0.2, 0.2, 0.2, 0.17, 0.03]).astype(f32)
    state = rng.choice([1, 1, 1, 1, 1, 2, 3, 4], n1).astype(np.uint8)
    already_at = np.zeros(n1, np.int64)                               # distinct features for the already-found points
    found = np.flatnonzero(state == 3)
    already_at[found] = rng.permutation(n2)[:len(found)]
    occupied = (rng.random(n2) < 0.1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    return k1, d1, k2, d2, uv, depth, state, already_at.astype(np.int32), occupied, level, rng


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("seed,th,ratio", [(51, 3, 1.0), (52, 8, 1.0), (53, 10, 0.8), (54, 30, 1.5)])
def test_search_by_projection_sim3_equals_reference(oracle, rf, M, seed, th, ratio, variant):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (R/lib_src/ORBmatcher.cc:372-471) and its
    overload with the points' key frames (:473-580), the matchers of LoopClosing: bad / already-found points, image and
    invariance-window tests, the [level - 1, level] window, 'feature already matched' skips incl. matches made earlier in
    the call, the float threshold TH_LOW * ratioHamming."""
    k1, d1, k2, d2, uv, depth, state, already_at, occupied, level, rng = _sim3_case(oracle, seed)
    sf = oracle.tables()["scale"]
    n1 = len(k1)
    min_d = rng.uniform(0.0, 50.0, n1).astype(np.float32)
    max_d = (min_d + rng.uniform(100.0, 5000.0, n1)).astype(np.float32)
    rn, rkm, dist3d = rf.search_by_projection_sim3(variant, k2, d2, sf, (0, 0, 640, 480), occupied, state, already_at, uv, depth,
                                                   min_d, max_d, level, d1, th, ratio)
    occ = occupied.astype(bool).copy()
    occ[already_at[state == 3]] = True
    n, km = M.search_by_projection_sim3(k2, d2, sf, (0, 0, 640, 480), occ, (state == 1) & ~(depth < 0), uv, dist3d, min_d, max_d,
                                        level, d1, th, ratio)
    assert rn >= 0 and n == rn and np.array_equal(km, rkm)
    assert rn > 50


@pytest.mark.parametrize("seed,th", [(61, 3.0), (62, 4.0), (63, 10.0)])
def test_fuse_sim3_equals_reference(oracle, rf, M, seed, th):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (R/lib_src/ORBmatcher.cc:1182-1292): the fuse decision per
    candidate == the Fuse matching core without the reprojection gates (zero inverse sigma, no right coordinates)."""
    k1, d1, k2, d2, uv, depth, state, already_at, occupied, level, rng = _sim3_case(oracle, seed)
    sf = oracle.tables()["scale"]
    n1, n2 = len(k1), len(k2)
    min_d = rng.uniform(0.0, 50.0, n1).astype(np.float32)
    max_d = (min_d + rng.uniform(100.0, 5000.0, n1)).astype(np.float32)
    rn, rbest, dist3d = rf.fuse_sim3(k2, d2, sf, (0, 0, 640, 480), occupied, state, already_at, uv, depth, min_d, max_d, level, d1, th)
    n, best, _ = M.fuse_search(k2, d2, sf, np.zeros(8, np.float32), (0, 0, 640, 480), np.full(n2, -1.0, np.float32),
                               (state == 1) & ~(depth < 0), uv, np.zeros(n1, np.float32), dist3d, min_d, max_d, level, d1, th)
    assert rn >= 0 and n == rn and np.array_equal(best, rbest)
    assert rn > 50


def _sim3_pair_case(oracle, seed):
    """Two key frames of a loop: frame 1 / frame 2 of a known image motion; every feature's map point projects into the other
    image where the feature moved (+ noise); depths are powers of two (see _sim3_case)."""
    k1, d1, k2, d2 = _frame_pair(oracle, seed)
    rng = np.random.default_rng(seed + 400)
    f32 = np.float32

    def side(ka, dx, dy):
        n = len(ka)
        uv = (np.stack([ka["x"] + dx, ka["y"] + dy], 1).astype(f32) + rng.normal(0, 0.8, (n, 2)).astype(f32))
        uv[rng.random(n) < 0.03] += f32(700.0)
        depth = rng.choice([0.5, 1.0, 2.0, 4.0, 8.0, -1.0], n, p=[0.2, 0.2, 0.2, 0.2, 0.17, 0.03]).astype(f32)
        state = rng.choice([0, 1, 1, 1, 1, 1, 2], n).astype(np.uint8)
        level = np.clip(ka["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32)
        mn = rng.uniform(0.0, 50.0, n).astype(f32)
        mx = (mn + rng.uniform(100.0, 5000.0, n)).astype(f32)
        return uv, depth, state, level, mn, mx
    s1, s2 = side(k1, -3.0, 2.0), side(k2, 3.0, -2.0)
    pre = np.full(len(k1), -2, np.int32)                              # vpMatches12 on entry
    some = rng.random(len(k1)) < 0.1
    pre[some] = rng.integers(-1, len(k2), int(some.sum()))
    return k1, d1, k2, d2, s1, s2, pre


@pytest.mark.parametrize("seed,th", [(71, 7.5), (72, 3.0), (73, 15.0)])
def test_search_by_sim3_equals_reference(oracle, rf, M, seed, th):
    """ORBmatcher::SearchBySim3 (R/lib_src/ORBmatcher.cc:1293-1497): both search directions, 'already matched' features on
    both sides, TH_HIGH, and the mutual-agreement test."""
    k1, d1, k2, d2, (uv12, z1, st1, lv12, mn1, mx1), (uv21, z2, st2, lv21, mn2, mx2), pre = _sim3_pair_case(oracle, seed)
    sf = oracle.tables()["scale"]
    rn, rm, q12, q21 = rf.search_by_sim3(k1, d1, k2, d2, sf, (0, 0, 640, 480), st1, pre, uv12, z1, mn1, mx1, lv12, st2, uv21, z2,
                                         mn2, mx2, lv21, th)
    matched2 = np.zeros(len(k2), bool)
    matched2[pre[pre >= 0]] = True
    v1 = (st1 == 1) & (pre == -2) & ~(z1 < 0)
    v2 = (st2 == 1) & ~matched2 & ~(z2 < 0)
    n, m = M.search_by_sim3(k1, d1, k2, d2, sf, (0, 0, 640, 480), v1, uv12, q12, mn1, mx1, lv12, v2, uv21, q21, mn2, mx2, lv21, th)
    assert n == rn and np.array_equal(m, rm)
    assert rn > 30
