"""GPU parity tests of the bag-of-words path (vocabulary tree descent, Frame::ComputeBoW vectors, SearchByBoW) against
the oracle, through the C ABI.  Integer results (word / node ids, match indices) are bit-exact; BowVector values are
the same doubles (the device only looks weights up, the arithmetic is the host's)."""
import numpy as np
import pytest

from rumi_slam_b200.synth import synthetic_vocabulary, descriptors_near_vocabulary

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bow_oracle():
    from oracle import bow_oracle as B
    B.build()
    return B


def _features(desc, leaf, n, seed):
    rng = np.random.default_rng(seed)
    near = descriptors_near_vocabulary(desc, leaf, n - n // 4, seed)
    return np.concatenate([near, rng.integers(0, 256, (n // 4, 32), dtype=np.uint8)])


@pytest.mark.parametrize("k,L,n", [(10, 4, 3000), (4, 6, 2000), (17, 2, 1000), (20, 3, 777), (2, 9, 500)])
def test_transform_matches_oracle(bow_oracle, k, L, n):
    from rumi_slam_b200 import ORBVocabulary
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=3 * k + L, stop_every=5)
    O = bow_oracle.Vocabulary(k, L, par, leaf, desc, w)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    assert V.size() == int(leaf.sum())
    f = _features(desc, leaf, n, 7)
    for levelsup in (0, 1, 4, L, L + 3):
        word, weight, node = V.transform_features(f, levelsup)
        rw, rwt, rn = O.transform(f, levelsup)
        assert np.array_equal(word, rw) and np.array_equal(node, rn) and np.array_equal(weight, rwt)
        assert V.transform(f, levelsup) == O.vectors(f, levelsup)
    assert V.transform_features(np.zeros((0, 32), np.uint8)) [0].shape == (0,)


def test_transform_full_size_vocabulary_and_device_path(bow_oracle):
    import torch
    from rumi_slam_b200 import ORBVocabulary
    k, L = 10, 6                                   # the shape of ORBvoc.txt: 1 111 111 nodes, 10^6 words
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=1)
    O = bow_oracle.Vocabulary(k, L, par, leaf, desc, w)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    f = _features(desc, leaf, 4000, 9)
    word, weight, node = V.transform_features(f, 4)
    rw, rwt, rn = O.transform(f, 4)
    assert np.array_equal(word, rw) and np.array_equal(node, rn) and np.array_equal(weight, rwt)
    dw, dwt, dn = V.transform_features_device(torch.from_numpy(f).cuda(), 4)
    assert np.array_equal(dw.cpu().numpy(), rw) and np.array_equal(dn.cpu().numpy(), rn)
    assert np.array_equal(dwt.cpu().numpy(), rwt)


def test_unbalanced_tree_and_ties(bow_oracle):
    from rumi_slam_b200 import ORBVocabulary
    # root -> {1, 2, 3}; 1 is a leaf, 2 -> {4, 5}, 3 -> {6, 7, 8}; children 4 and 5 are IDENTICAL (tie -> first wins)
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (9, 32), dtype=np.uint8)
    d[5] = d[4]
    d[2] = d[4]                                    # so that a feature equal to d[4] goes down through node 2
    par = np.array([0, 0, 0, 0, 2, 2, 3, 3, 3], np.int32)
    leaf = np.array([0, 1, 0, 0, 1, 1, 1, 1, 1], np.uint8)
    w = np.array([0, 1.5, 0, 0, 2.5, 3.5, 0.0, 4.5, 5.5])
    O = bow_oracle.Vocabulary(3, 2, par, leaf, d, w)
    V = ORBVocabulary(3, 2, par, leaf, d, w)
    f = np.concatenate([d[1:], rng.integers(0, 256, (200, 32), dtype=np.uint8)])
    for levelsup in (0, 1, 2):
        a, b = V.transform_features(f, levelsup), O.transform(f, levelsup)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    word, _, _ = V.transform_features(d[5:6], 0)
    assert word[0] == 1                            # node 4 (the first of the tied children) is word 1, node 5 is word 2


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_search_by_bow_matches_oracle(bow_oracle, check_ori, seed):
    from rumi_slam_b200 import ORBVocabulary, ORBmatcher
    k, L = 10, 4
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=11)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    rng = np.random.default_rng(seed)
    nk, nf = 1000, 1100
    dk = descriptors_near_vocabulary(desc, leaf, nk, 20 + seed, flip=0.04)
    # frame features: noisy copies of 800 keyframe features (some duplicated) + unrelated ones, shuffled
    src = rng.integers(0, nk, 900)
    noise = np.packbits(rng.random((900, 256)) < 0.03, axis=1)
    df = np.concatenate([dk[src] ^ noise, descriptors_near_vocabulary(desc, leaf, nf - 900, 30 + seed)])
    perm = rng.permutation(nf)
    df = df[perm]
    ak = rng.random(nk).astype(np.float32) * 360
    af = np.zeros(nf, np.float32)
    af[np.argsort(perm)[:900]] = (ak[src] - 15 + rng.normal(0, 2, 900)).astype(np.float32) % 360
    valid = (rng.random(nk) < 0.8).astype(np.uint8)
    _, fv_k = V.transform(dk, 2)
    _, fv_f = V.transform(df, 2)
    m = ORBmatcher(0.75, check_ori)
    n_gpu, match_gpu = m.SearchByBoW(dk, ak, valid, fv_k, df, af, fv_f)
    n_ref, match_ref = bow_oracle.search_by_bow(dk, ak, valid, fv_k, df, af, fv_f, nnratio=0.75, check_ori=check_ori)
    assert n_gpu == n_ref and n_ref > 100
    assert np.array_equal(match_gpu, match_ref)


def test_search_by_bow_edge_cases(bow_oracle):
    from rumi_slam_b200 import ORBmatcher
    m = ORBmatcher(0.75, True)
    d = np.zeros((3, 32), np.uint8)
    n, match = m.SearchByBoW(d, np.zeros(3), np.ones(3, np.uint8), {1: [0, 1, 2]}, d, np.zeros(3), {2: [0, 1, 2]})
    assert n == 0 and list(match) == [-1, -1, -1]                 # no common node
    n, match = m.SearchByBoW(d, np.zeros(3), np.zeros(3, np.uint8), {1: [0, 1, 2]}, d, np.zeros(3), {1: [0, 1, 2]})
    assert n == 0                                                 # no valid map point
    n, match = m.SearchByBoW(d, np.zeros(3), np.ones(3, np.uint8), {}, d, np.zeros(3), {1: [0]})
    assert n == 0


def test_distinctive_descriptors_match_oracle(bow_oracle):
    from rumi_slam_b200 import ORBmatcher
    rng = np.random.default_rng(3)
    sizes = np.concatenate([rng.integers(0, 12, 400), rng.integers(30, 100, 20), [1, 2, 0, 33, 64, 65]])
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    centre = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    parts = [centre[p] ^ np.packbits(rng.random((sizes[p], 256)) < 0.08, axis=1) for p in range(len(sizes)) if sizes[p]]
    desc = np.concatenate(parts)
    desc[off[5]:off[6]] = desc[off[5]] if sizes[5] else 0            # a point whose observations are all identical
    m = ORBmatcher()
    best, med = m.ComputeDistinctiveDescriptors(desc, off)
    rb, rm = bow_oracle.distinctive(desc, off)
    assert np.array_equal(best, rb) and np.array_equal(med, rm)
    assert m.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.array([0], np.int32))[0].shape == (0,)


@pytest.mark.parametrize("check_ori", [False, True])
def test_search_by_bow_keyframe_keyframe_matches_oracle(bow_oracle, check_ori):
    from rumi_slam_b200 import ORBVocabulary, ORBmatcher
    k, L = 10, 4
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=12)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    rng = np.random.default_rng(4)
    n1, n2 = 900, 1000
    d1 = descriptors_near_vocabulary(desc, leaf, n1, 40, flip=0.04)
    src = rng.integers(0, n1, 800)
    d2 = np.concatenate([d1[src] ^ np.packbits(rng.random((800, 256)) < 0.03, axis=1),
                         descriptors_near_vocabulary(desc, leaf, n2 - 800, 41)])
    perm = rng.permutation(n2)
    d2 = d2[perm]
    a1 = rng.random(n1).astype(np.float32) * 360
    a2 = rng.random(n2).astype(np.float32) * 360
    a2[np.argsort(perm)[:800]] = (a1[src] + 20 + rng.normal(0, 2, 800)).astype(np.float32) % 360
    v1 = (rng.random(n1) < 0.85).astype(np.uint8)
    v2 = (rng.random(n2) < 0.85).astype(np.uint8)
    _, fv1 = V.transform(d1, 2)
    _, fv2 = V.transform(d2, 2)
    m = ORBmatcher(0.8, check_ori)
    n_gpu, match_gpu = m.SearchByBoW_KF(d1, a1, v1, fv1, d2, a2, v2, fv2)
    n_ref, match_ref = bow_oracle.search_by_bow_kf(d1, a1, v1, fv1, d2, a2, v2, fv2, nnratio=0.8, check_ori=check_ori)
    assert n_gpu == n_ref and n_ref > 100 and np.array_equal(match_gpu, match_ref)
    assert all(v2[j] for j in match_gpu if j >= 0)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,only_stereo,coarse", [(40, False, False), (41, True, False), (42, False, True)])
def test_search_for_triangulation_matches_oracle(bow_oracle, seed, only_stereo, coarse, check_ori):
    """ORBmatcher::SearchForTriangulation (R/lib_src/ORBmatcher.cc:806-1013); the oracle side is pinned to the reference function
    in tests/test_ref_frame_pin.py::test_search_for_triangulation_equals_reference."""
    from rumi_slam_b200 import KP_DTYPE, ORBVocabulary, ORBmatcher
    k, L = 10, 4
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=12)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    rng = np.random.default_rng(seed)
    n1, n2 = 900, 1000
    d1 = descriptors_near_vocabulary(desc, leaf, n1, seed, flip=0.04)
    src = rng.integers(0, n1, 800)
    d2 = np.concatenate([d1[src] ^ np.packbits(rng.random((800, 256)) < 0.03, axis=1),
                         descriptors_near_vocabulary(desc, leaf, n2 - 800, seed + 1)])
    perm = rng.permutation(n2)
    d2 = d2[perm]
    a1 = rng.random(n1).astype(np.float32) * 360
    a2 = rng.random(n2).astype(np.float32) * 360
    a2[np.argsort(perm)[:800]] = (a1[src] + 20 + rng.normal(0, 2, 800)).astype(np.float32) % 360
    k2 = np.zeros(n2, KP_DTYPE)
    k2["x"], k2["y"], k2["octave"] = rng.uniform(20, 620, n2), rng.uniform(20, 460, n2), rng.integers(0, 8, n2)
    has1, has2 = rng.random(n1) < 0.3, rng.random(n2) < 0.3
    st1, st2 = rng.random(n1) < 0.5, rng.random(n2) < 0.5
    epi = ((np.arange(n1)[:, None] * 31 + np.arange(n2)[None, :] * 17) % 5 != 0)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    _, fv1 = V.transform(d1, 2)
    _, fv2 = V.transform(d2, 2)
    calls = []

    def epipolar_ok(i1, i2):                       # the caller's geometry: only asked for pairs that survive the distance tests
        calls.append((i1, i2))
        return bool(epi[i1, i2])
    n, m = ORBmatcher(0.6, check_ori).SearchForTriangulation(d1, a1, has1, st1, fv1, d2, a2, has2, st2, k2, fv2, sf, (320.0, 240.0),
                                                            epipolar_ok, only_stereo, coarse)
    rn, rm = bow_oracle.search_for_triangulation(d1, a1, has1, st1, fv1, d2, a2, has2, st2, k2["x"], k2["y"], k2["octave"], fv2, sf,
                                                 (320.0, 240.0), epi, only_stereo, coarse, check_ori)
    assert n == rn and np.array_equal(m, rm) and rn > 50
    assert (len(calls) == 0) == coarse and len(calls) < 20 * n1


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,frac", [(7, 0.5), (8, 0.7)])
def test_search_by_bow_stereo_fisheye_branch_matches_oracle(bow_oracle, seed, frac, check_ori):
    """The F.Nleft != -1 branches of SearchByBoW (R/lib_src/ORBmatcher.cc:258-340); oracle pinned in
    tests/test_ref_frame_pin.py::test_search_by_bow_stereo_fisheye_branch_equals_reference."""
    from rumi_slam_b200 import ORBVocabulary, ORBmatcher
    k, L = 10, 4
    par, leaf, desc, w = synthetic_vocabulary(k, L, seed=11)
    V = ORBVocabulary(k, L, par, leaf, desc, w)
    rng = np.random.default_rng(seed)
    nk, nf = 900, 1000
    dk = descriptors_near_vocabulary(desc, leaf, nk, 20 + seed, flip=0.04)
    src = rng.integers(0, nk, 800)
    df = np.concatenate([dk[src] ^ np.packbits(rng.random((800, 256)) < 0.03, axis=1),
                         descriptors_near_vocabulary(desc, leaf, nf - 800, 30 + seed)])[rng.permutation(nf)]
    ak = rng.random(nk).astype(np.float32) * 360
    af = rng.random(nf).astype(np.float32) * 360
    valid = (rng.random(nk) < 0.8).astype(np.uint8)
    _, fv_k = V.transform(dk, 2)
    _, fv_f = V.transform(df, 2)
    n_left = int(nf * frac)
    n, m = ORBmatcher(0.75, check_ori).SearchByBoW(dk, ak, valid, fv_k, df, af, fv_f, n_left=n_left)
    rn, rm = bow_oracle.search_by_bow(dk, ak, valid, fv_k, df, af, fv_f, nnratio=0.75, check_ori=check_ori, n_left=n_left)
    assert n == rn and np.array_equal(m, rm) and rn > 100 and (rm[n_left:] >= 0).sum() > 20


def test_stereo_fisheye_matches_equal_cv2_knn(bow_oracle):
    """Frame::ComputeStereoFishEyeMatches (R/lib_src/Frame.cc:1120-1161), matching core == cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2)
    + Lowe's ratio 0.7 on the lapping-area rows."""
    import cv2
    from rumi_slam_b200 import ORBmatcher
    rng = np.random.default_rng(5)
    nl, nr, mono_l, mono_r = 1200, 1100, 400, 350
    dr = rng.integers(0, 256, (nr, 32), dtype=np.uint8)
    dl = rng.integers(0, 256, (nl, 32), dtype=np.uint8)
    pick = rng.integers(mono_r, nr, nl - mono_l)                      # most lapping-area features have a noisy twin on the right
    dl[mono_l:] = dr[pick] ^ np.packbits(rng.random((nl - mono_l, 256)) < 0.06, axis=1)
    pairs = ORBmatcher().StereoFishEyeMatches(dl, mono_l, dr, mono_r)
    ref = []
    for mm in cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(dl[mono_l:], dr[mono_r:], 2):
        if len(mm) >= 2 and mm[0].distance < mm[1].distance * 0.7:
            ref.append((mm[0].queryIdx + mono_l, mm[0].trainIdx + mono_r))
    assert pairs == ref and len(ref) > 300
    assert ORBmatcher().StereoFishEyeMatches(dl, mono_l, dr[:mono_r + 1], mono_r) == []      # one candidate: size() >= 2 fails
