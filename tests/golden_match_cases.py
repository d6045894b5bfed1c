"""Seeded scenes for the frozen matcher-function vectors (tests/golden/match_functions.npz).  Every case builds its inputs
from the scene helpers of test_ref_frame_pin.py and runs either the REFERENCE function (tools/gen_golden_match.py, dev
container) or the ORACLE (tests/test_golden_match.py, anywhere) on them; both return the same named outputs."""
import numpy as np

import test_ref_frame_pin as T

BOUNDS = (0, 0, 640, 480)


def _f32(a):
    return np.asarray(a, np.float32)


def last_frame(oracle, B, rf, reference, seed=11, stereo=True, th=15.0):
    from oracle import match_oracle as M
    k1, d1, k2, d2, uv, depth, valid, has_obs, occupied, u_right = T._last_frame_case(oracle, seed, stereo)
    sf = oracle.tables()["scale"]
    if reference:
        n, cm = rf.search_by_projection_last(k2, d2, sf, BOUNDS, valid, uv, depth, k1["octave"], k1["angle"], d1, has_obs, th,
                                             u_right, occupied, 40.0, False, False, True)
    else:
        invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
        n, cm = M.search_by_projection_last(k2, d2, sf, BOUNDS, valid, uv, invz, k1["octave"], k1["angle"], d1, has_obs, th,
                                            u_right, occupied, 40.0, False, False, True)
    return {"n": n, "match": cm}


def keyframe(oracle, B, rf, reference, seed=21, th=10.0, orb_dist=100):
    from oracle import match_oracle as M
    k1, d1, k2, d2, uv, depth, _, _, occupied, _ = T._last_frame_case(oracle, seed, False)
    depth = np.abs(depth)
    rng = np.random.default_rng(seed + 100)
    n1 = len(k1)
    sf = oracle.tables()["scale"]
    state = rng.choice([0, 1, 1, 1, 1, 1, 2, 3], n1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(-1, 2, n1), 0, 7).astype(np.int32)
    min_d = rng.uniform(0.0, 6.0, n1).astype(np.float32)
    max_d = (min_d + rng.uniform(0.0, 800.0, n1)).astype(np.float32)
    if reference:
        n, cm, dist3d = rf.search_by_projection_kf(k2, d2, sf, BOUNDS, state, uv, depth, level, min_d, max_d, k1["angle"], d1, th,
                                                   orb_dist, occupied, True)
        return {"n": n, "match": cm, "dist3d": dist3d}
    dist3d = reference_dist(oracle, "keyframe/dist3d")
    n, cm = M.search_by_projection_kf(k2, d2, sf, BOUNDS, state == 1, uv, dist3d, min_d, max_d, level, k1["angle"], d1, th, orb_dist,
                                      occupied, True)
    return {"n": n, "match": cm, "dist3d": dist3d}


_FROZEN = {}


def reference_dist(oracle, key):
    return _FROZEN[key]


def fuse(oracle, B, rf, reference, seed=32, th=3.0):
    from oracle import match_oracle as M
    k1, d1, k2, d2, uv, depth, _, _, _, _ = T._last_frame_case(oracle, seed, False)
    rng = np.random.default_rng(seed + 200)
    n1, n2 = len(k1), len(k2)
    f32 = np.float32
    uv = (np.stack([k1["x"] - 3.0, k1["y"] + 2.0], 1).astype(f32) + rng.normal(0, 0.8, (n1, 2)).astype(f32))
    tab = oracle.tables()
    sf, inv_s2 = tab["scale"], (1.0 / (tab["scale"].astype(np.float64) ** 2)).astype(f32)
    state = rng.choice([0, 1, 1, 1, 1, 1, 1, 2, 3, 4], n1).astype(np.uint8)
    level = np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7).astype(np.int32)
    bf = 40.0
    u_right = np.where(rng.random(n2) < 0.6, k2["x"] - bf / rng.uniform(2.0, 20.0, n2), -1.0).astype(f32)
    kf_has = rng.choice([0, 0, 0, 1], n2).astype(np.uint8)
    n_obs = rng.integers(1, 10, n1).astype(np.int32)
    x, y, z = uv[:, 0], uv[:, 1], depth
    dist3d = np.sqrt(((x * x + y * y).astype(f32) + z * z).astype(f32)).astype(f32)
    ur = (x - (f32(bf) * (f32(1.0) / z).astype(f32)).astype(f32)).astype(f32)
    min_d = (dist3d * rng.choice([0.5, 0.9, 1.01], n1)).astype(f32)
    max_d = (dist3d * rng.choice([0.99, 1.1, 2.0], n1)).astype(f32)
    if reference:
        n, best = rf.fuse(k2, d2, sf, inv_s2, BOUNDS, u_right, kf_has, bf, state, uv, depth, min_d, max_d, level, d1, n_obs, th)
    else:
        n, best, _ = M.fuse_search(k2, d2, sf, inv_s2, BOUNDS, u_right, (state == 1) & ~(z < 0), uv, ur, dist3d, min_d, max_d, level,
                                   d1, th)
    return {"n": n, "best": best}


def triangulation(oracle, B, rf, reference, seed=40):
    d1, a1, fv1, k1, has1, ur1, d2, a2, fv2, k2, has2, ur2, epi, ep = T._triangulation_case(B, seed)
    sf = oracle.tables()["scale"]
    if reference:
        n, m = rf.search_for_triangulation(k1, d1, has1, ur1, fv1, k2, d2, has2, ur2, fv2, sf, ep, epi, False, False, True)
    else:
        n, m = B.search_for_triangulation(d1, a1, has1, ur1 >= 0, fv1, d2, a2, has2, ur2 >= 0, k2["x"], k2["y"], k2["octave"], fv2, sf,
                                          ep, epi, False, False, True)
    return {"n": n, "match": m}


def triangulation_rig(oracle, B, rf, reference, seed=47):
    """Two key frames of a stereo-fisheye rig: the existing entry point with flattened key points, no stereo flags, far epipole."""
    d1, a1, fv1, k1, has1, _, d2, a2, fv2, k2, has2, _, epi, ep = T._triangulation_case(B, seed)
    sf = oracle.tables()["scale"]
    if reference:
        n, m = rf.search_for_triangulation_rig(k1, d1, has1, fv1, len(k1) * 2 // 3, k2, d2, has2, fv2, len(k2) // 2, sf, ep, epi,
                                               False, False, True)
    else:
        n, m = B.search_for_triangulation(d1, a1, has1, np.zeros(len(d1), bool), fv1, d2, a2, has2, np.zeros(len(d2), bool),
                                          k2["x"], k2["y"], k2["octave"], fv2, sf, (3.0e9, 3.0e9), epi, False, False, True)
    return {"n": n, "match": m}


def sim3_pair(oracle, B, rf, reference, seed=71, th=7.5):
    from oracle import match_oracle as M
    k1, d1, k2, d2, (uv12, z1, st1, lv12, mn1, mx1), (uv21, z2, st2, lv21, mn2, mx2), pre = T._sim3_pair_case(oracle, seed)
    sf = oracle.tables()["scale"]
    if reference:
        n, m, q12, q21 = rf.search_by_sim3(k1, d1, k2, d2, sf, BOUNDS, st1, pre, uv12, z1, mn1, mx1, lv12, st2, uv21, z2, mn2, mx2,
                                           lv21, th)
        return {"n": n, "match": m, "q12": q12, "q21": q21}
    q12, q21 = reference_dist(oracle, "sim3_pair/q12"), reference_dist(oracle, "sim3_pair/q21")
    matched2 = np.zeros(len(k2), bool)
    matched2[pre[pre >= 0]] = True
    n, m = M.search_by_sim3(k1, d1, k2, d2, sf, BOUNDS, (st1 == 1) & (pre == -2) & ~(z1 < 0), uv12, q12, mn1, mx1, lv12,
                            (st2 == 1) & ~matched2 & ~(z2 < 0), uv21, q21, mn2, mx2, lv21, th)
    return {"n": n, "match": m, "q12": q12, "q21": q21}


def initialization(oracle, B, rf, reference, seed=2, window=100, ratio=0.9):
    from oracle import match_oracle as M
    k1, d1, k2, d2 = T._frame_pair(oracle, seed)
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    side = rf if reference else M
    n, m, p = side.search_for_initialization(k1, d1, k2, d2, BOUNDS, prev, window, ratio, True)
    return {"n": n, "match": m, "prev": p}


def last_frame_fisheye(oracle, B, rf, reference, seed=61, th=15.0):
    from oracle import match_oracle as M
    k1, d1, kC, kR, dC, uv, depth, valid, has_obs, occupied, shift = T._last_frame_fisheye_case(oracle, seed)
    sf = oracle.tables()["scale"]
    if reference:
        n, cm = rf.search_by_projection_last_fisheye(kC, kR, dC, sf, BOUNDS, valid, uv, shift, depth, k1["octave"], k1["angle"], d1,
                                                     has_obs, th, occupied, False, False, True)
    else:
        invz = (1.0 / depth.astype(np.float64)).astype(np.float32)
        uvr = (uv + np.asarray(shift, np.float32)).astype(np.float32)
        n, cm = M.search_by_projection_last_fisheye(kC, kR, dC, sf, BOUNDS, valid, uv, uvr, invz, k1["octave"], k1["angle"], d1,
                                                    has_obs, th, occupied, False, False, True)
    return {"n": n, "match": cm}


def _local_points(mode, seed, th, ratio):
    def case(oracle, B, rf, reference):
        from oracle import match_oracle as M
        kL, dF, dMP, kw = T._local_points_case(oracle, seed, mode)
        sf = oracle.tables()["scale"]
        proj, level, view_cos, has_obs = kw.pop("proj"), kw.pop("level"), kw.pop("view_cos"), kw.pop("has_obs")
        fn = (rf if reference else M).search_by_projection_ex
        n, fm = fn(kL, dF, sf, BOUNDS, proj, level, view_cos, dMP, has_obs, th, ratio, **kw)
        return {"n": n, "match": fm}
    return case


CASES = {"last_frame": last_frame, "keyframe": keyframe, "fuse": fuse, "triangulation": triangulation, "sim3_pair": sim3_pair,
         "initialization": initialization,
         "local_points_occupied": _local_points("occupied", 41, 3.0, 0.8), "local_points_stereo": _local_points("stereo", 42, 5.0, 0.8),
         "local_points_fisheye": _local_points("fisheye", 43, 3.0, 0.8), "last_frame_fisheye": last_frame_fisheye,
         "triangulation_rig": triangulation_rig}
