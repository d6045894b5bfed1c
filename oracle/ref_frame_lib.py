"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/librefframe.so.

librefframe.so = reference functions of Frame.cc / MapPoint.cc / ORBmatcher.cc cut out of /root/reference at build time
and compiled UNMODIFIED over class stand-ins (oracle/ref_frame_shim.cpp, oracle/extract_ref.py, `make -C oracle refframe`).
Built only where /root/reference exists; travels to the GPU box as a prebuilt file.  Pins the oracle's restatements of
ComputeStereoMatches, ComputeDistinctiveDescriptors, both SearchByBoW overloads, GetFeaturesInArea,
SearchForInitialization and SearchByProjection.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .orb_oracle import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "librefframe.so")
REF_SRC = "/root/reference/src/rumi-slam/lib_src/Frame.cc"

_u8p, _f32p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)


def build():
    if os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "refframe"], stdout=subprocess.DEVNULL)
    return os.path.exists(_LIB)


def available():
    return os.path.exists(_LIB) or build()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/librefframe.so not built (reference sources absent)")
        _lib = C.CDLL(_LIB)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def stereo_match(pyrL, pyrR, Lk, Ld, Rk, Rd, scale, inv_scale, mbf, mb):
    """Frame::ComputeStereoMatches on the given pyramids (lists of dense u8 levels): (mvuRight, mvDepth, kept)."""
    n = len(pyrL)
    keepL = [np.ascontiguousarray(a, np.uint8) for a in pyrL]
    keepR = [np.ascontiguousarray(a, np.uint8) for a in pyrR]
    PL = (C.c_void_p * n)(*[a.ctypes.data for a in keepL])
    PR = (C.c_void_p * n)(*[a.ctypes.data for a in keepR])
    lw = np.array([a.shape[1] for a in keepL], np.int32)
    lh = np.array([a.shape[0] for a in keepL], np.int32)
    Lk, Rk = np.ascontiguousarray(Lk, KP_DTYPE), np.ascontiguousarray(Rk, KP_DTYPE)
    Ld, Rd = np.ascontiguousarray(Ld, np.uint8), np.ascontiguousarray(Rd, np.uint8)
    sc, isc = np.ascontiguousarray(scale, np.float32), np.ascontiguousarray(inv_scale, np.float32)
    u, d = np.zeros(len(Lk), np.float32), np.zeros(len(Lk), np.float32)
    L = lib()
    L.ref_stereo_match.argtypes = [C.c_void_p, C.c_void_p, _i32p, _i32p, C.c_int, C.c_void_p, _u8p, C.c_int, C.c_void_p,
                                   _u8p, C.c_int, _f32p, _f32p, C.c_float, C.c_float, _f32p, _f32p]
    kept = L.ref_stereo_match(PL, PR, _p(lw, _i32p), _p(lh, _i32p), n, Lk.ctypes.data, _p(Ld, _u8p), len(Lk),
                              Rk.ctypes.data, _p(Rd, _u8p), len(Rk), _p(sc, _f32p), _p(isc, _f32p), float(mbf), float(mb),
                              _p(u, _f32p), _p(d, _f32p))
    return u, d, kept


def distinctive(desc):
    """MapPoint::ComputeDistinctiveDescriptors for one map point: the 32 bytes the reference stores in mDescriptor."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    out = np.zeros(32, np.uint8)
    L = lib()
    L.ref_distinctive.argtypes = [_u8p, C.c_int, _u8p]
    rc = L.ref_distinctive(_p(desc, _u8p), len(desc), _p(out, _u8p))
    return None if rc else out


def descriptor_distance(a, b):
    a, b = np.ascontiguousarray(a, np.uint8).reshape(32), np.ascontiguousarray(b, np.uint8).reshape(32)
    L = lib()
    L.ref_descriptor_distance.argtypes = [_u8p, _u8p]
    return L.ref_descriptor_distance(_p(a, _u8p), _p(b, _u8p))


def _nodes(fv, n):
    node = np.full(max(n, 1), -1, np.int32)
    for nid, idx in fv.items():
        for i in idx:
            node[i] = nid
    return node


def search_by_bow(desc_kf, angle_kf, kf_valid, fv_kf, desc_f, angle_f, fv_f, nnratio=0.7, check_ori=True, n_left=-1):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...).  fv_*: {node id: [feature indices in ascending order]} (the order
    DBoW2::FeatureVector::addFeature produces when features are added by ascending index).  (nmatches, matchF)."""
    dk, df = np.ascontiguousarray(desc_kf, np.uint8).reshape(-1, 32), np.ascontiguousarray(desc_f, np.uint8).reshape(-1, 32)
    ak, af = np.ascontiguousarray(angle_kf, np.float32), np.ascontiguousarray(angle_f, np.float32)
    vk = np.ascontiguousarray(kf_valid, np.uint8)
    nk, nf = _nodes(fv_kf, len(dk)), _nodes(fv_f, len(df))
    match = np.zeros(max(len(df), 1), np.int32)
    L = lib()
    L.ref_search_by_bow_nleft.argtypes = [_u8p, _f32p, _u8p, _i32p, C.c_int, _u8p, _f32p, _i32p, C.c_int, C.c_float, C.c_int, C.c_int,
                                          _i32p]
    n = L.ref_search_by_bow_nleft(_p(dk, _u8p), _p(ak, _f32p), _p(vk, _u8p), _p(nk, _i32p), len(dk), _p(df, _u8p), _p(af, _f32p),
                                  _p(nf, _i32p), len(df), float(nnratio), 1 if check_ori else 0, int(n_left), _p(match, _i32p))
    return n, match[:len(df)]


def search_by_bow_kf(desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nnratio=0.8, check_ori=True):
    d1, d2 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32), np.ascontiguousarray(desc2, np.uint8).reshape(-1, 32)
    a1, a2 = np.ascontiguousarray(angle1, np.float32), np.ascontiguousarray(angle2, np.float32)
    v1, v2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
    n1, n2 = _nodes(fv1, len(d1)), _nodes(fv2, len(d2))
    match = np.zeros(max(len(d1), 1), np.int32)
    L = lib()
    L.ref_search_by_bow_kf.argtypes = [_u8p, _f32p, _u8p, _i32p, C.c_int, _u8p, _f32p, _u8p, _i32p, C.c_int, C.c_float, C.c_int,
                                       _i32p]
    n = L.ref_search_by_bow_kf(_p(d1, _u8p), _p(a1, _f32p), _p(v1, _u8p), _p(n1, _i32p), len(d1), _p(d2, _u8p), _p(a2, _f32p),
                               _p(v2, _u8p), _p(n2, _i32p), len(d2), float(nnratio), 1 if check_ori else 0, _p(match, _i32p))
    return n, match[:len(d1)]


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    """Frame::GetFeaturesInArea on the grid AssignFeaturesToGrid builds; bounds = (mnMinX, mnMinY, mnMaxX, mnMaxY)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(max(len(kps), 1), np.int32)
    L = lib()
    L.ref_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_int, C.c_int, _i32p, C.c_int]
    n = L.ref_features_in_area(kps.ctypes.data, len(kps), *[int(b) for b in bounds], float(x), float(y), float(r),
                               int(min_level), int(max_level), _p(out, _i32p), len(out))
    return out[:n].copy()


def search_for_initialization(k1, d1, k2, d2, bounds, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """ORBmatcher::SearchForInitialization: (nmatches, vnMatches12, updated vbPrevMatched)."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    d1, d2 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32), np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.zeros(max(len(k1), 1), np.int32)
    L = lib()
    L.ref_search_for_initialization.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int, _f32p, C.c_int, C.c_float, C.c_int, _i32p]
    n = L.ref_search_for_initialization(k1.ctypes.data, _p(d1, _u8p), len(k1), k2.ctypes.data, _p(d2, _u8p), len(k2),
                                        *[int(b) for b in bounds], _p(prev, _f32p), int(window), float(nnratio),
                                        1 if check_ori else 0, _p(m12, _i32p))
    return n, m12[:len(k1)], prev


def search_by_projection(kF, dF, scale_factors, bounds, proj, level, view_cos, dMP, has_obs, th=3.0, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) on a mono frame: (nmatches, frameMatch[j] = map point)."""
    kF = np.ascontiguousarray(kF, KP_DTYPE)
    dF, dMP = np.ascontiguousarray(dF, np.uint8).reshape(-1, 32), np.ascontiguousarray(dMP, np.uint8).reshape(-1, 32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    proj = np.ascontiguousarray(proj, np.float32)
    level = np.ascontiguousarray(level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32)
    ho = np.ascontiguousarray(has_obs, np.uint8)
    out = np.zeros(max(len(kF), 1), np.int32)
    L = lib()
    L.ref_search_by_projection.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p,
                                           _i32p, _f32p, _u8p, _u8p, C.c_int, C.c_float, C.c_float, _i32p]
    n = L.ref_search_by_projection(kF.ctypes.data, _p(dF, _u8p), len(kF), _p(sf, _f32p), len(sf), *[int(b) for b in bounds],
                                   _p(proj, _f32p), _p(level, _i32p), _p(vc, _f32p), _p(dMP, _u8p), _p(ho, _u8p), len(dMP),
                                   float(th), float(nnratio), _p(out, _i32p))
    return n, out[:len(kF)]


def search_by_projection_ex(kL, dF, scale_factors, bounds, proj, level, view_cos, dMP, has_obs, th=3.0, nnratio=0.8, kR=None, u_right=None,
                            occupied=None, l2r=None, r2l=None, in_view=None, in_view_r=None, proj_r=None, level_r=None,
                            view_cos_r=None):
    """The reference's own ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) with occupied features, mvuRight and the
    stereo-fisheye members filled in.  (nmatches, frameMatch[nL + nR])."""
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
    fn = L.ref_search_by_projection_ex
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                   _i32p, _i32p, _u8p, _u8p, _f32p, _f32p, _i32p, _i32p, _f32p, _f32p, _u8p, _u8p, C.c_int, C.c_float,
                   C.c_float, _i32p]
    n = fn(kL.ctypes.data, nL, kR.ctypes.data if nR else None, nR, _p(dF, _u8p), _p(sf, _f32p), len(sf),
           *[int(b) for b in bounds], opt(u_right, np.float32, _f32p), opt(occupied, np.uint8, _u8p), opt(l2r, np.int32, _i32p),
           opt(r2l, np.int32, _i32p), opt(in_view, np.uint8, _u8p), opt(in_view_r, np.uint8, _u8p),
           opt(proj, np.float32, _f32p), opt(proj_r, np.float32, _f32p), opt(level, np.int32, _i32p),
           opt(level_r, np.int32, _i32p), opt(view_cos, np.float32, _f32p), opt(view_cos_r, np.float32, _f32p),
           _p(dMP, _u8p), opt(has_obs, np.uint8, _u8p), len(dMP), float(th), float(nnratio), _p(out, _i32p))
    return n, out[:nL + nR]


def search_by_projection_last(kC, dC, scale_factors, bounds, valid, uv, depth, octave, angle_last, dMP, mp_has_obs, th=15.0,
                              u_right=None, occupied=None, mbf=0.0, forward=False, backward=False, check_ori=True):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) over the stand-in camera
    (x, y, z) -> (x, y): map point i sits at (uv[i], depth[i]).  (nmatches, curMatch[j] = last-frame feature or -1)."""
    kC = np.ascontiguousarray(kC, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dC, dMP = u8(dC).reshape(-1, 32), u8(dMP).reshape(-1, 32)
    sf = f32(scale_factors)
    occ = u8(np.zeros(len(kC)) if occupied is None else occupied)
    ur = None if u_right is None else f32(u_right)
    out = np.zeros(max(len(kC), 1), np.int32)
    L = lib()
    L.ref_search_by_projection_last.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                _f32p, _u8p, C.c_float, _u8p, _f32p, _f32p, _i32p, _f32p, _u8p, _u8p, C.c_int,
                                                C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, _i32p]
    n = L.ref_search_by_projection_last(kC.ctypes.data, _p(dC, _u8p), len(kC), _p(sf, _f32p), len(sf), *[int(b) for b in bounds],
                                        None if ur is None else _p(ur, _f32p), _p(occ, _u8p), float(mbf), _p(u8(valid), _u8p),
                                        _p(f32(uv), _f32p), _p(f32(depth), _f32p), _p(np.ascontiguousarray(octave, np.int32), _i32p),
                                        _p(f32(angle_last), _f32p), _p(dMP, _u8p), _p(u8(mp_has_obs), _u8p), len(dMP), float(th),
                                        int(forward), int(backward), 0.9, int(check_ori), _p(out, _i32p))
    return n, out[:len(kC)]


def search_by_projection_last_fisheye(kC, kR, dC, scale_factors, bounds, valid, uv, shift, depth, octave, angle_last, dMP,
                                      mp_has_obs, th=15.0, occupied=None, forward=False, backward=False, check_ori=True):
    """The reference's SearchByProjection(CurrentFrame, LastFrame, th, false) with a stereo-fisheye current frame (Nleft =
    len(kC)); the right camera sees point i at uv[i] + shift.  (nmatches, curMatch[len(kC) + len(kR)])."""
    kC, kR = np.ascontiguousarray(kC, KP_DTYPE), np.ascontiguousarray(kR, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dC, dMP, sf, uv, depth, al = u8(dC).reshape(-1, 32), u8(dMP).reshape(-1, 32), f32(scale_factors), f32(uv), f32(depth), f32(angle_last)
    N = len(kC) + len(kR)
    occ = u8(np.zeros(N) if occupied is None else occupied)
    va, ho, oc = u8(valid), u8(mp_has_obs), np.ascontiguousarray(octave, np.int32)
    out = np.zeros(max(N, 1), np.int32)
    L = lib()
    fn = L.ref_search_by_projection_last_fisheye
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, _u8p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p,
                   _f32p, C.c_float, C.c_float, _f32p, _i32p, _f32p, _u8p, _u8p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float,
                   C.c_int, _i32p]
    n = fn(kC.ctypes.data, len(kC), kR.ctypes.data, len(kR), _p(dC, _u8p), _p(sf, _f32p), len(sf), *[int(b) for b in bounds],
           _p(occ, _u8p), _p(va, _u8p), _p(uv, _f32p), float(shift[0]), float(shift[1]), _p(depth, _f32p), _p(oc, _i32p),
           _p(al, _f32p), _p(dMP, _u8p), _p(ho, _u8p), len(dMP), float(th), int(forward), int(backward), 0.9, int(check_ori),
           _p(out, _i32p))
    return n, out[:N]


def search_by_projection_kf(kC, dC, scale_factors, bounds, state, uv, depth, level, min_dist, max_dist, angle_kf, dMP, th=10.0,
                            orb_dist=100, occupied=None, check_ori=True):
    """The reference's ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist); state[i]: 0 no map
    point, 1 usable, 2 bad, 3 already found.  Returns (nmatches, curMatch, dist3D the function compared with the window)."""
    kC = np.ascontiguousarray(kC, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dC, dMP = u8(dC).reshape(-1, 32), u8(dMP).reshape(-1, 32)
    sf = f32(scale_factors)
    occ = u8(np.zeros(len(kC)) if occupied is None else occupied)
    out = np.zeros(max(len(kC), 1), np.int32)
    d3 = np.zeros(max(len(dMP), 1), np.float32)
    L = lib()
    L.ref_search_by_projection_kf.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p,
                                              _u8p, _f32p, _f32p, _i32p, _f32p, _f32p, _f32p, _u8p, C.c_int, C.c_float, C.c_int,
                                              C.c_float, C.c_int, _i32p, _f32p]
    n = L.ref_search_by_projection_kf(kC.ctypes.data, _p(dC, _u8p), len(kC), _p(sf, _f32p), len(sf), *[int(b) for b in bounds],
                                      _p(occ, _u8p), _p(u8(state), _u8p), _p(f32(uv), _f32p), _p(f32(depth), _f32p),
                                      _p(np.ascontiguousarray(level, np.int32), _i32p), _p(f32(min_dist), _f32p),
                                      _p(f32(max_dist), _f32p), _p(f32(angle_kf), _f32p), _p(dMP, _u8p), len(dMP), float(th),
                                      int(orb_dist), 0.9, int(check_ori), _p(out, _i32p), _p(d3, _f32p))
    return n, out[:len(kC)], d3[:len(dMP)]


def fuse(kK, dK, scale_factors, inv_level_sigma2, bounds, u_right, kf_has_point, bf, state, uv, depth, min_dist, max_dist, level,
         dMP, n_obs, th=3.0):
    """The reference's ORBmatcher::Fuse(pKF, vpMapPoints, th, false): (nFused, bestIdx[i] = key-frame feature map point i was
    fused with, reconstructed from the logged GetMapPoint / Replace / AddObservation calls; -1 = not fused, or fused with a
    feature whose resident point is bad -- that case leaves no trace)."""
    kK = np.ascontiguousarray(kK, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    dK, dMP = u8(dK).reshape(-1, 32), u8(dMP).reshape(-1, 32)
    sf = f32(scale_factors)
    n = len(dMP)
    out = np.zeros(max(n, 1), np.int32)
    L = lib()
    L.ref_fuse.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                           C.c_float, _u8p, _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, _i32p, C.c_int, C.c_float, _i32p]
    nf = L.ref_fuse(kK.ctypes.data, _p(dK, _u8p), len(kK), _p(sf, _f32p), _p(f32(inv_level_sigma2), _f32p), len(sf),
                    *[int(b) for b in bounds], _p(f32(u_right), _f32p), _p(u8(kf_has_point), _u8p), float(bf), _p(u8(state), _u8p),
                    _p(f32(uv), _f32p), _p(f32(depth), _f32p), _p(f32(min_dist), _f32p), _p(f32(max_dist), _f32p),
                    _p(i32(level), _i32p), _p(dMP, _u8p), _p(i32(n_obs), _i32p), n, float(th), _p(out, _i32p))
    return nf, out[:n]


def fuse_right(n_left, kR, dAll, scale_factors, inv_level_sigma2, bounds, u_right_all, kf_has_point, bf, state, uv, depth, min_dist,
               max_dist, level, dMP, n_obs, th=3.0):
    """The reference's ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = true) on a stereo-fisheye key frame (NLeft = n_left, right
    key points kR, descriptor rows n_left + i): (nFused, bestIdx[i] = n_left + right feature, -1 = none / no trace)."""
    kR = np.ascontiguousarray(kR, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    dAll, dMP = u8(dAll).reshape(-1, 32), u8(dMP).reshape(-1, 32)
    assert len(dAll) == n_left + len(kR)
    sf = f32(scale_factors)
    n = len(dMP)
    out = np.zeros(max(n, 1), np.int32)
    L = lib()
    fn = L.ref_fuse_right
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_void_p, C.c_int, _u8p, _f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p,
                   C.c_float, _u8p, _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, _i32p, C.c_int, C.c_float, _i32p]
    nf = fn(int(n_left), kR.ctypes.data, len(kR), _p(dAll, _u8p), _p(sf, _f32p), _p(f32(inv_level_sigma2), _f32p), len(sf),
            *[int(b) for b in bounds], _p(f32(u_right_all), _f32p), _p(u8(kf_has_point), _u8p), float(bf), _p(u8(state), _u8p),
            _p(f32(uv), _f32p), _p(f32(depth), _f32p), _p(f32(min_dist), _f32p), _p(f32(max_dist), _f32p), _p(i32(level), _i32p),
            _p(dMP, _u8p), _p(i32(n_obs), _i32p), n, float(th), _p(out, _i32p))
    return nf, out[:n]


def search_for_triangulation(k1, d1, has_mp1, u_right1, fv1, k2, d2, has_mp2, u_right2, fv2, scale_factors2, ep, epi_ok,
                             only_stereo=False, coarse=False, check_ori=True):
    """The reference's ORBmatcher::SearchForTriangulation over stand-in poses / camera (epipole = ep, epipolarConstrain = the
    table epi_ok[n1, n2]): (nmatches, match12)."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    d1, d2 = u8(d1).reshape(-1, 32), u8(d2).reshape(-1, 32)
    sf = f32(scale_factors2)
    n1, n2 = _nodes(fv1, len(d1)), _nodes(fv2, len(d2))
    match = np.zeros(max(len(d1), 1), np.int32)
    L = lib()
    L.ref_search_for_triangulation.argtypes = [C.c_void_p, _u8p, _u8p, _f32p, _i32p, C.c_int, C.c_void_p, _u8p, _u8p, _f32p, _i32p,
                                               C.c_int, _f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, _u8p, C.c_float,
                                               C.c_int, _i32p]
    n = L.ref_search_for_triangulation(k1.ctypes.data, _p(d1, _u8p), _p(u8(has_mp1), _u8p), _p(f32(u_right1), _f32p), _p(n1, _i32p),
                                       len(d1), k2.ctypes.data, _p(d2, _u8p), _p(u8(has_mp2), _u8p), _p(f32(u_right2), _f32p),
                                       _p(n2, _i32p), len(d2), _p(sf, _f32p), len(sf), float(ep[0]), float(ep[1]), int(only_stereo),
                                       int(coarse), _p(u8(epi_ok), _u8p), 0.6, int(check_ori), _p(match, _i32p))
    return n, match[:len(d1)]


def search_for_triangulation_rig(k1, d1, has_mp1, fv1, n_left1, k2, d2, has_mp2, fv2, n_left2, scale_factors2, ep, epi_ok,
                                 only_stereo=False, coarse=False, check_ori=True):
    """The reference's ORBmatcher::SearchForTriangulation on two key frames of a stereo-fisheye rig (mpCamera2 set; k* = the
    left key points followed by the right ones, NLeft = n_left*): (nmatches, match12)."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    d1, d2 = u8(d1).reshape(-1, 32), u8(d2).reshape(-1, 32)
    sf = f32(scale_factors2)
    n1, n2 = _nodes(fv1, len(d1)), _nodes(fv2, len(d2))
    match = np.zeros(max(len(d1), 1), np.int32)
    L = lib()
    fn = L.ref_search_for_triangulation_rig
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, _u8p, _u8p, _i32p, C.c_int, C.c_int, C.c_void_p, _u8p, _u8p, _i32p, C.c_int, C.c_int, _f32p, C.c_int,
                   C.c_float, C.c_float, C.c_int, C.c_int, _u8p, C.c_float, C.c_int, _i32p]
    n = fn(k1.ctypes.data, _p(d1, _u8p), _p(u8(has_mp1), _u8p), _p(n1, _i32p), len(d1), int(n_left1), k2.ctypes.data, _p(d2, _u8p),
           _p(u8(has_mp2), _u8p), _p(n2, _i32p), len(d2), int(n_left2), _p(sf, _f32p), len(sf), float(ep[0]), float(ep[1]),
           int(only_stereo), int(coarse), _p(u8(epi_ok), _u8p), 0.6, int(check_ori), _p(match, _i32p))
    return n, match[:len(d1)]


def _sim3_args(kK, dK, scale_factors, bounds, occupied, state, already_at, uv, depth, min_dist, max_dist, level, dMP):
    kK = np.ascontiguousarray(kK, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    dK, dMP = u8(dK).reshape(-1, 32), u8(dMP).reshape(-1, 32)
    sf = f32(scale_factors)
    keep = [kK, dK, dMP, sf, u8(occupied), u8(state), i32(already_at), f32(uv), f32(depth), f32(min_dist), f32(max_dist), i32(level)]
    args = [kK.ctypes.data, _p(dK, _u8p), len(kK), _p(sf, _f32p), len(sf), *[int(b) for b in bounds], _p(keep[4], _u8p),
            _p(keep[5], _u8p), _p(keep[6], _i32p), _p(keep[7], _f32p), _p(keep[8], _f32p), _p(keep[9], _f32p), _p(keep[10], _f32p),
            _p(keep[11], _i32p), _p(dMP, _u8p), len(dMP)]
    types = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u8p, _i32p, _f32p, _f32p, _f32p,
             _f32p, _i32p, _u8p, C.c_int]
    return keep, args, types


def search_by_projection_sim3(variant, kK, dK, scale_factors, bounds, occupied, state, already_at, uv, depth, min_dist, max_dist,
                              level, dMP, th=3, ratio_hamming=1.0):
    """The reference's ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (variant 0) or its overload
    with vpPointsKFs / vpMatchedKF (variant 1): (nmatches, kfMatch, dist3D the function compared with the invariance window)."""
    keep, args, types = _sim3_args(kK, dK, scale_factors, bounds, occupied, state, already_at, uv, depth, min_dist, max_dist, level, dMP)
    out = np.zeros(max(len(keep[0]), 1), np.int32)
    d3 = np.zeros(max(len(keep[2]), 1), np.float32)
    L = lib()
    L.ref_search_by_projection_sim3.argtypes = [C.c_int] + types + [C.c_int, C.c_float, _i32p, _f32p]
    n = L.ref_search_by_projection_sim3(int(variant), *args, int(th), float(ratio_hamming), _p(out, _i32p), _p(d3, _f32p))
    return n, out[:len(keep[0])], d3[:len(keep[2])]


def fuse_sim3(kK, dK, scale_factors, bounds, occupied, state, already_at, uv, depth, min_dist, max_dist, level, dMP, th=3.0):
    """The reference's ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint): (nFused, bestIdx, dist3D)."""
    keep, args, types = _sim3_args(kK, dK, scale_factors, bounds, occupied, state, already_at, uv, depth, min_dist, max_dist, level, dMP)
    out = np.zeros(max(len(keep[2]), 1), np.int32)
    d3 = np.zeros(max(len(keep[2]), 1), np.float32)
    L = lib()
    L.ref_fuse_sim3.argtypes = types + [C.c_float, _i32p, _f32p]
    n = L.ref_fuse_sim3(*args, float(th), _p(out, _i32p), _p(d3, _f32p))
    return n, out[:len(keep[2])], d3[:len(keep[2])]


def search_by_sim3(k1, d1, k2, d2, scale_factors, bounds, state1, pre_match1, uv12, depth1, min1, max1, level12, state2, uv21,
                   depth2, min2, max2, level21, th=7.5):
    """The reference's ORBmatcher::SearchBySim3 over identity poses and a unit pinhole: (nFound, match12, dist12, dist21) --
    the dist arrays are the |p3Dc| norms the function compared with the invariance windows."""
    k1, k2 = np.ascontiguousarray(k1, KP_DTYPE), np.ascontiguousarray(k2, KP_DTYPE)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    d1, d2 = u8(d1).reshape(-1, 32), u8(d2).reshape(-1, 32)
    sf = f32(scale_factors)
    out = np.zeros(max(len(k1), 1), np.int32)
    q12, q21 = np.zeros(max(len(k1), 1), np.float32), np.zeros(max(len(k2), 1), np.float32)
    L = lib()
    L.ref_search_by_sim3.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, _u8p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _u8p, _i32p, _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, _f32p, _f32p, _f32p, _f32p, _i32p,
                                     C.c_float, _i32p, _f32p, _f32p]
    n = L.ref_search_by_sim3(k1.ctypes.data, _p(d1, _u8p), len(k1), k2.ctypes.data, _p(d2, _u8p), len(k2), _p(sf, _f32p), len(sf),
                             *[int(b) for b in bounds], _p(u8(state1), _u8p), _p(i32(pre_match1), _i32p), _p(f32(uv12), _f32p),
                             _p(f32(depth1), _f32p), _p(f32(min1), _f32p), _p(f32(max1), _f32p), _p(i32(level12), _i32p),
                             _p(u8(state2), _u8p), _p(f32(uv21), _f32p), _p(f32(depth2), _f32p), _p(f32(min2), _f32p),
                             _p(f32(max2), _f32p), _p(i32(level21), _i32p), float(th), _p(out, _i32p), _p(q12, _f32p), _p(q21, _f32p))
    return n, out[:len(k1)], q12[:len(k1)], q21[:len(k2)]
