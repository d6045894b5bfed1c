"""ORACLE -- TEST INFRASTRUCTURE ONLY (ctypes binding of oracle/liborb_oracle.so).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package never imports this module.
See oracle/orb_oracle.cpp for what each entry point restates (reference file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    src = os.path.join(_HERE, "orb_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u8p, f32p, i32p, u16p = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_uint16))
        L.orc_tables.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, f32p, i32p, i32p]
        L.orc_level_sizes.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, i32p, i32p]
        L.orc_resize.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_int, C.c_int, C.c_size_t]
        L.orc_resize.restype = None
        L.orc_fast_score_map.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_int16)]
        L.orc_fast_score_map.restype = None
        L.orc_fast.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, f32p, C.c_int]
        L.orc_grid_fast.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, f32p, C.c_int, i32p]
        L.orc_octree.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
        L.orc_octree_stable.argtypes = L.orc_octree.argtypes
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_ic_angle.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int]
        L.orc_ic_angle.restype = C.c_float
        L.orc_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p, C.c_size_t]
        L.orc_blur.restype = None
        L.orc_descriptor.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_float, u8p]
        L.orc_descriptor.restype = None
        L.orc_sincos.argtypes = [C.c_float, f32p, f32p]
        L.orc_sincos.restype = None
        L.orc_extract.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, u8p, C.c_int, i32p, i32p]
        L.orc_level_keypoints.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_describe.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, u8p]
        L.orc_descriptor_distance.argtypes = [u8p, u8p]
        L.orc_hamming_top2.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, u16p, u16p]
        L.orc_hamming_top2.restype = None
        L.orc_stereo_best1.argtypes = [C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int, f32p, C.c_int,
                                       C.c_float, C.c_float, i32p, u16p]
        L.orc_stereo_best1.restype = None
        L.orc_stereo_match.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_size_t, C.c_float, C.c_int, C.c_void_p, u8p, C.c_int,
                                       C.c_void_p, u8p, C.c_int, C.c_float, C.c_float, f32p, f32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def tables(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
    s, i, g, ig = (np.zeros(nlevels, np.float32) for _ in range(4))
    q = np.zeros(nlevels, np.int32)
    um = np.zeros(16, np.int32)
    lib().orc_tables(nfeatures, scale, nlevels, ini, mn, _p(s, C.c_float), _p(i, C.c_float), _p(g, C.c_float),
                     _p(ig, C.c_float), _p(q, C.c_int32), _p(um, C.c_int32))
    return dict(scale=s, inv_scale=i, sigma2=g, inv_sigma2=ig, quota=q, umax=um)


def level_sizes(W, H, scale=1.2, nlevels=8):
    ws, hs = np.zeros(nlevels, np.int32), np.zeros(nlevels, np.int32)
    lib().orc_level_sizes(W, H, scale, nlevels, _p(ws, C.c_int32), _p(hs, C.c_int32))
    return ws, hs


def resize(src, dw, dh):
    src = _img(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize(_p(src, C.c_uint8), src.shape[1], src.shape[0], src.strides[0], _p(dst, C.c_uint8), dw, dh, dw)
    return dst


def pyramid(img, scale=1.2, nlevels=8):
    img = _img(img)
    ws, hs = level_sizes(img.shape[1], img.shape[0], scale, nlevels)
    out = [img.copy()]
    for l in range(1, nlevels):
        out.append(resize(out[-1], int(ws[l]), int(hs[l])))
    return out


def fast_score_map(img):
    img = _img(img)
    sc = np.zeros(img.shape, np.int16)
    lib().orc_fast_score_map(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], _p(sc, C.c_int16))
    return sc


def fast(img, th):
    img = _img(img)
    cap = img.size // 4 + 16
    out = np.zeros((cap, 3), np.float32)
    n = lib().orc_fast(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], th, _p(out, C.c_float), cap)
    return out[:n]


def grid_fast(img, ini=20, mn=7):
    """Candidates (x, y, response) relative to (16,16) in the reference's order + number of fallback cells."""
    img = _img(img)
    cap = img.size // 4 + 16
    out = np.zeros((cap, 3), np.float32)
    nf = C.c_int32(0)
    n = lib().orc_grid_fast(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], ini, mn,
                            _p(out, C.c_float), cap, C.byref(nf))
    return out[:n], nf.value


def octree(xyr, min_x, max_x, min_y, max_y, n_target):
    xyr = np.ascontiguousarray(xyr, np.float32)
    cap = n_target + 16 + len(xyr)
    sel = np.zeros(cap, np.int32)
    n = lib().orc_octree(_p(xyr, C.c_float), len(xyr), min_x, max_x, min_y, max_y, n_target, _p(sel, C.c_int32), cap)
    return sel[:n]


def octree_stable(xyr, min_x, max_x, min_y, max_y, n_target):
    """NOT the reference's behaviour: final phase sorted with std::stable_sort (tie-sensitivity census only)."""
    xyr = np.ascontiguousarray(xyr, np.float32)
    cap = n_target + 16 + len(xyr)
    sel = np.zeros(cap, np.int32)
    n = lib().orc_octree_stable(_p(xyr, C.c_float), len(xyr), min_x, max_x, min_y, max_y, n_target, _p(sel, C.c_int32), cap)
    return sel[:n]


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def ic_angle(img, x, y):
    img = _img(img)
    return lib().orc_ic_angle(_p(img, C.c_uint8), img.strides[0], int(x), int(y))


def blur(img):
    img = _img(img)
    dst = np.zeros_like(img)
    lib().orc_blur(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], _p(dst, C.c_uint8), dst.strides[0])
    return dst


def descriptor(blurred, x, y, angle_deg):
    blurred = _img(blurred)
    d = np.zeros(32, np.uint8)
    lib().orc_descriptor(_p(blurred, C.c_uint8), blurred.strides[0], int(x), int(y), float(angle_deg),
                         _p(d, C.c_uint8))
    return d


def sincos(angle_rad):
    c, s = C.c_float(0), C.c_float(0)
    lib().orc_sincos(float(angle_rad), C.byref(c), C.byref(s))
    return c.value, s.value


def extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lapping=(0, 0)):
    """ORBextractor::operator(): returns (kps[KP_DTYPE], desc[n,32] u8, monoIndex) or None on an empty image."""
    img = _img(img)
    cap = nfeatures + 64 * nlevels
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n, m = C.c_int32(0), C.c_int32(0)
    rc = lib().orc_extract(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], nfeatures, scale, nlevels,
                           ini, mn, int(lapping[0]), int(lapping[1]), kps.ctypes.data, _p(desc, C.c_uint8), cap,
                           C.byref(n), C.byref(m))
    if rc != 0:
        return None
    assert n.value <= cap
    return kps[:n.value].copy(), desc[:n.value].copy(), m.value


def level_keypoints(img, level, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
    img = _img(img)
    cap = nfeatures + 64
    kps = np.zeros(cap, KP_DTYPE)
    n = lib().orc_level_keypoints(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], nfeatures, scale,
                                  nlevels, ini, mn, level, kps.ctypes.data, cap)
    return kps[:n].copy()


def describe(img, kps):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.zeros((len(kps), 32), np.uint8)
    rc = lib().orc_describe(_p(img, C.c_uint8), img.shape[1], img.shape[0], img.strides[0], kps.ctypes.data,
                            len(kps), _p(desc, C.c_uint8))
    return rc, desc


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a, C.c_uint8), _p(b, C.c_uint8))


def hamming_top2(Q, T):
    Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
    T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
    i1 = np.zeros(len(Q), np.int32)
    d1 = np.zeros(len(Q), np.uint16)
    d2 = np.zeros(len(Q), np.uint16)
    lib().orc_hamming_top2(_p(Q, C.c_uint8), len(Q), _p(T, C.c_uint8), len(T), _p(i1, C.c_int32),
                           _p(d1, C.c_uint16), _p(d2, C.c_uint16))
    return i1, d1, d2


def hamming_top2_mt(Q, T, threads=None):
    """hamming_top2 with the query rows split over host threads (ctypes releases the GIL): same result, used by the
    benchmark-scale parity checks and as the all-cores CPU baseline of the matching metric."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
    threads = threads or len(os.sched_getaffinity(0))
    if threads <= 1 or len(Q) < 2 * threads:
        return hamming_top2(Q, T)
    cuts = np.linspace(0, len(Q), threads * 4 + 1).astype(int)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(lambda i: hamming_top2(Q[cuts[i]:cuts[i + 1]], T), range(len(cuts) - 1)))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))


def extract_mt(frames, threads=None, **kw):
    """extract() of many frames on host threads; returns a list of (kps, desc, mono)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or len(os.sched_getaffinity(0))
    build()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda im: extract(im, **kw), frames))


def stereo_best1(Lk, Ld, Rk, Rd, scale_factors, n_rows, min_d, max_d):
    Lk = np.ascontiguousarray(Lk, KP_DTYPE)
    Rk = np.ascontiguousarray(Rk, KP_DTYPE)
    Ld = np.ascontiguousarray(Ld, np.uint8)
    Rd = np.ascontiguousarray(Rd, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    best = np.zeros(len(Lk), np.int32)
    dist = np.zeros(len(Lk), np.uint16)
    lib().orc_stereo_best1(Lk.ctypes.data, _p(Ld, C.c_uint8), len(Lk), Rk.ctypes.data, _p(Rd, C.c_uint8), len(Rk),
                           _p(sf, C.c_float), n_rows, float(min_d), float(max_d), _p(best, C.c_int32),
                           _p(dist, C.c_uint16))
    return best, dist


def stereo_match(imgL, imgR, Lk, Ld, Rk, Rd, mbf, mb, scale=1.2, nlevels=8):
    """Frame::ComputeStereoMatches: returns (mvuRight, mvDepth, number of stereo matches kept)."""
    imgL, imgR = _img(imgL), _img(imgR)
    Lk = np.ascontiguousarray(Lk, KP_DTYPE)
    Rk = np.ascontiguousarray(Rk, KP_DTYPE)
    Ld = np.ascontiguousarray(Ld, np.uint8)
    Rd = np.ascontiguousarray(Rd, np.uint8)
    u = np.zeros(len(Lk), np.float32)
    d = np.zeros(len(Lk), np.float32)
    n = lib().orc_stereo_match(_p(imgL, C.c_uint8), _p(imgR, C.c_uint8), imgL.shape[1], imgL.shape[0], imgL.strides[0],
                               scale, nlevels, Lk.ctypes.data, _p(Ld, C.c_uint8), len(Lk), Rk.ctypes.data,
                               _p(Rd, C.c_uint8), len(Rk), float(mbf), float(mb), _p(u, C.c_float), _p(d, C.c_float))
    return u, d, n
