"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/match_oracle.cpp (feature grid, SearchForInitialization,
SearchByProjection, CloudMerging's pixel association).  Pinned against oracle/_ref/librefframe.so."""
import ctypes as C
import os
import subprocess

import numpy as np

from .orb_oracle import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmatch_oracle.so")
_u8p, _f32p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "match_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libmatch_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(max(len(kps), 1), np.int32)
    L = lib()
    L.mo_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_int, _i32p, C.c_int]
    n = L.mo_features_in_area(kps.ctypes.data, len(kps), *[int(b) for b in bounds], float(x), float(y), float(r),
                              int(min_level), int(max_level), _p(out, _i32p), len(out))
    return out[:n].copy()


def candidate_lists(kps, bounds, qxy, qr, qmin=None, qmax=None):
    """CSR candidate lists (off[nq+1], idx) of GetFeaturesInArea for many queries on one frame grid."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    qxy = np.ascontiguousarray(qxy, np.float32).reshape(-1, 2)
    nq = len(qxy)
    qr = np.ascontiguousarray(np.broadcast_to(np.asarray(qr, np.float32), (nq,)))
    qmin = None if qmin is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qmin, np.int32), (nq,)))
    qmax = None if qmax is None else np.ascontiguousarray(np.broadcast_to(np.asarray(qmax, np.int32), (nq,)))
    off = np.zeros(nq + 1, np.int32)
    L = lib()
    L.mo_candidate_lists.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _i32p, _i32p, C.c_int,
                                     _i32p, _i32p, C.c_int]
    cap = max(64 * nq, 1024)
    while True:
        idx = np.zeros(cap, np.int32)
        tot = L.mo_candidate_lists(kps.ctypes.data, len(kps), *[int(b) for b in bounds], _p(qxy, _f32p), _p(qr, _f32p),
                                   None if qmin is None else _p(qmin, _i32p), None if qmax is None else _p(qmax, _i32p), nq,
                                   _p(off, _i32p), _p(idx, _i32p), cap)
        if tot <= cap:
            return off, idx[:tot].copy()
        cap = tot


def search_for_initialization(k1, d1, k2, d2, bounds, prev_matched, window=100, nnratio=0.9, check_ori=True):
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.zeros(max(len(k1), 1), np.int32)
    L = lib()
    L.mo_search_for_initialization.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_int, _f32p, C.c_int, C.c_float, C.c_int, _i32p]
    n = L.mo_search_for_initialization(k1.ctypes.data, _p(d1, _u8p), len(k1), k2.ctypes.data, _p(d2, _u8p), len(k2),
                                       *[int(b) for b in bounds], _p(prev, _f32p), int(window), float(nnratio),
                                       1 if check_ori else 0, _p(m12, _i32p))
    return n, m12[:len(k1)], prev


def search_by_projection(kF, dF, scale_factors, bounds, proj, level, view_cos, dMP, has_obs, th=3.0, nnratio=0.8):
    kF = np.ascontiguousarray(kF, KP_DTYPE)
    dF, dMP = np.ascontiguousarray(dF, np.uint8).reshape(-1, 32), np.ascontiguousarray(dMP, np.uint8).reshape(-1, 32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    proj = np.ascontiguousarray(proj, np.float32)
    level = np.ascontiguousarray(level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32)
    ho = np.ascontiguousarray(has_obs, np.uint8)
    out = np.zeros(max(len(kF), 1), np.int32)
    L = lib()
    L.mo_search_by_projection.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _i32p,
                                          _f32p, _u8p, _u8p, C.c_int, C.c_float, C.c_float, _i32p]
    n = L.mo_search_by_projection(kF.ctypes.data, _p(dF, _u8p), len(kF), _p(sf, _f32p), *[int(b) for b in bounds],
                                  _p(proj, _f32p), _p(level, _i32p), _p(vc, _f32p), _p(dMP, _u8p), _p(ho, _u8p), len(dMP),
                                  float(th), float(nnratio), _p(out, _i32p))
    return n, out[:len(kF)]


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def search_by_projection_ex(kL, dF, scale_factors, bounds, proj, level, view_cos, dMP, has_obs, th=3.0, nnratio=0.8, kR=None, u_right=None,
                            occupied=None, l2r=None, r2l=None, in_view=None, in_view_r=None, proj_r=None, level_r=None,
                            view_cos_r=None):
    """Oracle restatement of the whole ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th): occupied features, mvuRight
    gate, stereo-fisheye halves.  (nmatches, frameMatch[nL + nR])."""
    kL = np.ascontiguousarray(kL, KP_DTYPE)
    kR = np.zeros(0, KP_DTYPE) if kR is None else np.ascontiguousarray(kR, KP_DTYPE)
    nL, nR = len(kL), len(kR)
    dF, dMP = np.ascontiguousarray(dF, np.uint8).reshape(-1, 32), np.ascontiguousarray(dMP, np.uint8).reshape(-1, 32)
    assert len(dF) == nL + nR
    sf = np.ascontiguousarray(scale_factors, np.float32)
    keep = []                                            # keeps the converted arrays alive during the call

    def opt(a, dt, ptr):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return a.ctypes.data_as(ptr)
    out = np.zeros(max(nL + nR, 1), np.int32)
    L = lib()
    fn = L.mo_search_by_projection_ex
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                   _i32p, _i32p, _u8p, _u8p, _f32p, _f32p, _i32p, _i32p, _f32p, _f32p, _u8p, _u8p, C.c_int, C.c_float,
                   C.c_float, _i32p]
    n = fn(kL.ctypes.data, nL, kR.ctypes.data if nR else None, nR, _p(dF, _u8p), _p(sf, _f32p),
           *[int(b) for b in bounds], opt(u_right, np.float32, _f32p), opt(occupied, np.uint8, _u8p), opt(l2r, np.int32, _i32p),
           opt(r2l, np.int32, _i32p), opt(in_view, np.uint8, _u8p), opt(in_view_r, np.uint8, _u8p),
           opt(proj, np.float32, _f32p), opt(proj_r, np.float32, _f32p), opt(level, np.int32, _i32p),
           opt(level_r, np.int32, _i32p), opt(view_cos, np.float32, _f32p), opt(view_cos_r, np.float32, _f32p),
           _p(dMP, _u8p), opt(has_obs, np.uint8, _u8p), len(dMP), float(th), float(nnratio), _p(out, _i32p))
    return n, out[:nL + nR]


def search_by_projection_last(kC, dC, scale_factors, bounds, valid, uv, invz, octave, angle_last, dMP, mp_has_obs, th=15.0,
                              u_right=None, occupied=None, mbf=0.0, forward=False, backward=False, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1498-1684, Nleft == -1):
    (nmatches, curMatch[j] = last-frame feature or -1)."""
    kC = np.ascontiguousarray(kC, KP_DTYPE)
    dC, dMP = _u8(dC).reshape(-1, 32), _u8(dMP).reshape(-1, 32)
    sf, uv, invz, ang = _f32(scale_factors), _f32(uv), _f32(invz), _f32(angle_last)
    octv = np.ascontiguousarray(octave, np.int32)
    occ = _u8(np.zeros(len(kC)) if occupied is None else occupied)
    ur = None if u_right is None else _f32(u_right)
    out = np.zeros(max(len(kC), 1), np.int32)
    L = lib()
    L.mo_search_by_projection_last.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                                               C.c_float, _u8p, _f32p, _f32p, _i32p, _f32p, _u8p, _u8p, C.c_int, C.c_float,
                                               C.c_int, C.c_int, C.c_int, _i32p]
    n = L.mo_search_by_projection_last(kC.ctypes.data, _p(dC, _u8p), len(kC), _p(sf, _f32p), *[int(b) for b in bounds],
                                       None if ur is None else _p(ur, _f32p), _p(occ, _u8p), float(mbf), _p(_u8(valid), _u8p),
                                       _p(uv, _f32p), _p(invz, _f32p), _p(octv, _i32p), _p(ang, _f32p), _p(dMP, _u8p),
                                       _p(_u8(mp_has_obs), _u8p), len(dMP), float(th), int(forward), int(backward),
                                       int(check_ori), _p(out, _i32p))
    return n, out[:len(kC)]


def search_by_projection_last_fisheye(kC, kR, dC, scale_factors, bounds, valid, uv, uv_r, invz, octave, angle_last, dMP,
                                      mp_has_obs, th=15.0, occupied=None, forward=False, backward=False, check_ori=True):
    """Oracle restatement of SearchByProjection(CurrentFrame, LastFrame, th, false) with a stereo-fisheye current frame:
    uv_r[i] = projection of point i into the right camera.  (nmatches, curMatch[len(kC) + len(kR)])."""
    kC, kR = np.ascontiguousarray(kC, KP_DTYPE), np.ascontiguousarray(kR, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dC, dMP, sf, uv, uvr, iz, al = (u8(dC).reshape(-1, 32), u8(dMP).reshape(-1, 32), f32(scale_factors), f32(uv), f32(uv_r),
                                    f32(invz), f32(angle_last))
    N = len(kC) + len(kR)
    occ = u8(np.zeros(N) if occupied is None else occupied)
    va, ho, oc = u8(valid), u8(mp_has_obs), np.ascontiguousarray(octave, np.int32)
    out = np.zeros(max(N, 1), np.int32)
    L = lib()
    fn = L.mo_search_by_projection_last_fisheye
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _f32p,
                   _f32p, _f32p, _i32p, _f32p, _u8p, _u8p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _i32p]
    n = fn(kC.ctypes.data, len(kC), kR.ctypes.data, len(kR), _p(dC, _u8p), _p(sf, _f32p), *[int(b) for b in bounds],
           _p(occ, _u8p), _p(va, _u8p), _p(uv, _f32p), _p(uvr, _f32p), _p(iz, _f32p), _p(oc, _i32p), _p(al, _f32p),
           _p(dMP, _u8p), _p(ho, _u8p), len(dMP), float(th), int(forward), int(backward), int(check_ori), _p(out, _i32p))
    return n, out[:N]


def search_by_projection_kf(kC, dC, scale_factors, bounds, valid, uv, dist3d, min_dist, max_dist, level, angle_kf, dMP, th=10.0,
                            orb_dist=100, occupied=None, check_ori=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1685-1794):
    (nmatches, curMatch[j] = key-frame feature or -1)."""
    kC = np.ascontiguousarray(kC, KP_DTYPE)
    dC, dMP = _u8(dC).reshape(-1, 32), _u8(dMP).reshape(-1, 32)
    occ = _u8(np.zeros(len(kC)) if occupied is None else occupied)
    out = np.zeros(max(len(kC), 1), np.int32)
    L = lib()
    L.mo_search_by_projection_kf.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p,
                                             _f32p, _f32p, _f32p, _f32p, _i32p, _f32p, _u8p, C.c_int, C.c_float, C.c_int,
                                             C.c_int, _i32p]
    n = L.mo_search_by_projection_kf(kC.ctypes.data, _p(dC, _u8p), len(kC), _p(_f32(scale_factors), _f32p),
                                     *[int(b) for b in bounds], _p(occ, _u8p), _p(_u8(valid), _u8p), _p(_f32(uv), _f32p),
                                     _p(_f32(dist3d), _f32p), _p(_f32(min_dist), _f32p), _p(_f32(max_dist), _f32p),
                                     _p(np.ascontiguousarray(level, np.int32), _i32p), _p(_f32(angle_kf), _f32p), _p(dMP, _u8p),
                                     len(dMP), float(th), int(orb_dist), int(check_ori), _p(out, _i32p))
    return n, out[:len(kC)]


def fuse_search(kK, dK, scale_factors, inv_level_sigma2, bounds, u_right, valid, uv, ur, dist3d, min_dist, max_dist, level, dMP,
                th=3.0, th_dist=50):
    """Matching core of ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:1015-1147): (nFused, bestIdx, bestDist)."""
    kK = np.ascontiguousarray(kK, KP_DTYPE)
    dK, dMP = _u8(dK).reshape(-1, 32), _u8(dMP).reshape(-1, 32)
    n = len(dMP)
    bi, bd = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    L = lib()
    L.mo_fuse_search.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p, _f32p,
                                 _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, C.c_int, C.c_float, C.c_int, _i32p, _i32p]
    nf = L.mo_fuse_search(kK.ctypes.data, _p(dK, _u8p), len(kK), _p(_f32(scale_factors), _f32p), _p(_f32(inv_level_sigma2), _f32p),
                          *[int(b) for b in bounds], _p(_f32(u_right), _f32p), _p(_u8(valid), _u8p), _p(_f32(uv), _f32p),
                          _p(_f32(ur), _f32p), _p(_f32(dist3d), _f32p), _p(_f32(min_dist), _f32p), _p(_f32(max_dist), _f32p),
                          _p(np.ascontiguousarray(level, np.int32), _i32p), _p(dMP, _u8p), n, float(th), int(th_dist), _p(bi, _i32p), _p(bd, _i32p))
    return nf, bi[:n], bd[:n]


def search_by_projection_sim3(kK, dK, scale_factors, bounds, occupied, valid, uv, dist3d, min_dist, max_dist, level, dMP, th=3,
                              ratio_hamming=1.0):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:372-471 / :473-580):
    (nmatches, kfMatch[j] = candidate map point or -1)."""
    kK = np.ascontiguousarray(kK, KP_DTYPE)
    dK, dMP = _u8(dK).reshape(-1, 32), _u8(dMP).reshape(-1, 32)
    out = np.zeros(max(len(kK), 1), np.int32)
    L = lib()
    L.mo_search_by_projection_sim3.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p,
                                               _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, C.c_int, C.c_int, C.c_float, _i32p]
    n = L.mo_search_by_projection_sim3(kK.ctypes.data, _p(dK, _u8p), len(kK), _p(_f32(scale_factors), _f32p),
                                       *[int(b) for b in bounds], _p(_u8(occupied), _u8p), _p(_u8(valid), _u8p), _p(_f32(uv), _f32p),
                                       _p(_f32(dist3d), _f32p), _p(_f32(min_dist), _f32p), _p(_f32(max_dist), _f32p),
                                       _p(np.ascontiguousarray(level, np.int32), _i32p), _p(dMP, _u8p), len(dMP), int(th),
                                       float(ratio_hamming), _p(out, _i32p))
    return n, out[:len(kK)]


def search_by_sim3(k1, d1, k2, d2, scale_factors, bounds, valid1, uv12, dist12, min1, max1, level12, valid2, uv21, dist21, min2,
                   max2, level21, th=7.5):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1293-1497): (nFound, match12[i1] = feature of key frame 2 or -1)."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, d2 = _u8(d1).reshape(-1, 32), _u8(d2).reshape(-1, 32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    out = np.zeros(max(len(k1), 1), np.int32)
    L = lib()
    L.mo_search_by_sim3.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _u8p, _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, _f32p, _f32p, _f32p, _f32p, _i32p, C.c_float, _i32p]
    n = L.mo_search_by_sim3(k1.ctypes.data, _p(d1, _u8p), len(k1), k2.ctypes.data, _p(d2, _u8p), len(k2), _p(_f32(scale_factors), _f32p),
                            *[int(b) for b in bounds], _p(_u8(valid1), _u8p), _p(_f32(uv12), _f32p), _p(_f32(dist12), _f32p),
                            _p(_f32(min1), _f32p), _p(_f32(max1), _f32p), _p(i32(level12), _i32p), _p(_u8(valid2), _u8p),
                            _p(_f32(uv21), _f32p), _p(_f32(dist21), _f32p), _p(_f32(min2), _f32p), _p(_f32(max2), _f32p),
                            _p(i32(level21), _i32p), float(th), _p(out, _i32p))
    return n, out[:len(k1)]


def associate_pixels(k1, valid1, k2, valid2, bounds, tol=3.0):
    """CloudMerging.cc:503-551 for one key-frame pair: (matchNum, match12)."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    v1, v2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
    out = np.zeros(max(len(k1), 1), np.int32)
    L = lib()
    L.mo_associate_pixels.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_float, _i32p]
    n = L.mo_associate_pixels(k1.ctypes.data, _p(v1, _u8p), len(k1), k2.ctypes.data, _p(v2, _u8p), len(k2),
                              *[int(b) for b in bounds], float(tol), _p(out, _i32p))
    return n, out[:len(k1)]
