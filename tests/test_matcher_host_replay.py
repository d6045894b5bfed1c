"""Host logic of the Python mirror without a GPU: the acceptance replay of ORBmatcher.SearchByProjection (occupied features,
right-image gate, stereo-fisheye halves) against the oracle that is pinned to the reference function.  The ONE thing the GPU
does on this path -- the Hamming distances of the candidate lists (rumi_hamming_candidates, covered by tests/test_gpu_candidates.py)
-- is supplied by numpy HERE, IN THE TEST: the product has no such path."""
import numpy as np
import pytest

from test_ref_frame_pin import _local_points_case          # noqa: E402  (tests/ is on sys.path under pytest)


@pytest.fixture(scope="module")
def M():
    from oracle import match_oracle
    match_oracle.build()
    return match_oracle


def _mirror_with_numpy_distances():
    from rumi_slam_b200.matcher import ORBmatcher
    m = ORBmatcher.__new__(ORBmatcher)                      # no handle: nothing below may touch the library

    def candidates(Q, T, off, idx, top2=False):
        Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
        T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
        q_of = np.repeat(np.arange(len(Q)), np.diff(off))
        return np.unpackbits(Q[q_of] ^ T[idx], axis=1).sum(1).astype(np.uint16)
    m.candidates = candidates
    return m


@pytest.mark.parametrize("mode", ["occupied", "stereo", "fisheye"])
@pytest.mark.parametrize("seed,th,ratio", [(6, 3.0, 0.8), (7, 1.0, 0.8), (9, 15.0, 0.6)])
def test_search_by_projection_replay_equals_oracle(oracle, M, seed, th, ratio, mode):
    kL, dF, dMP, kw = _local_points_case(oracle, seed, mode)
    sf = oracle.tables()["scale"]
    proj, level, view_cos, has_obs = kw.pop("proj"), kw.pop("level"), kw.pop("view_cos"), kw.pop("has_obs")
    n, fm = M.search_by_projection_ex(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th, ratio, **kw)
    m = _mirror_with_numpy_distances()
    m.mfNNratio = np.float32(ratio)
    names = {"kR": "keys_right"}
    gn, gfm = m.SearchByProjection(kL, dF, sf, (0, 0, 640, 480), proj, level, view_cos, dMP, has_obs, th,
                                   **{names.get(k, k): v for k, v in kw.items()})
    assert gn == n and np.array_equal(gfm, fm)
    assert n > 40


def test_default_arguments_take_the_original_path(oracle, M):
    kL, dF, dMP, kw = _local_points_case(oracle, 6, "occupied")
    sf = oracle.tables()["scale"]
    m = _mirror_with_numpy_distances()
    m.mfNNratio = np.float32(0.8)
    n, fm = M.search_by_projection(kL, dF, sf, (0, 0, 640, 480), kw["proj"], kw["level"], kw["view_cos"], dMP, kw["has_obs"], 3.0, 0.8)
    gn, gfm = m.SearchByProjection(kL, dF, sf, (0, 0, 640, 480), kw["proj"], kw["level"], kw["view_cos"], dMP, kw["has_obs"], 3.0)
    assert gn == n and np.array_equal(gfm, fm)


@pytest.mark.parametrize("check_ori", [False, True])
@pytest.mark.parametrize("seed,th,direction", [(51, 15.0, "none"), (52, 7.0, "none"), (53, 15.0, "forward"), (54, 15.0, "backward")])
def test_last_frame_fisheye_replay_equals_oracle(oracle, M, seed, th, direction, check_ori):
    from test_ref_frame_pin import _last_frame_fisheye_case
    k1, d1, kC, kR, dC, uv, depth, valid, has_obs, occupied, shift = _last_frame_fisheye_case(oracle, seed)
    sf = oracle.tables()["scale"]
    fw, bw = direction == "forward", direction == "backward"
    invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
    uvr = (uv + np.asarray(shift, np.float32)).astype(np.float32)
    n, cm = M.search_by_projection_last_fisheye(kC, kR, dC, sf, (0, 0, 640, 480), valid, uv, uvr, invz, k1["octave"], k1["angle"],
                                                d1, has_obs, th, occupied, fw, bw, check_ori)
    m = _mirror_with_numpy_distances()
    m.mbCheckOrientation = check_ori
    gn, gcm = m.SearchByProjectionLastFrameFisheye(kC, kR, dC, sf, (0, 0, 640, 480), valid, uv, uvr, invz, k1["octave"],
                                                   k1["angle"], d1, has_obs, th, occupied, fw, bw)
    assert gn == n and np.array_equal(gcm, cm)
    assert n > 100
