"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes bindings of oracle/libbow_oracle.so (the C++ restatement of the DBoW2
tree descent, Frame::ComputeBoW and ORBmatcher::SearchByBoW, see bow_oracle.cpp) and of oracle/_ref/librefbow.so (the
UNMODIFIED reference DBoW2 compiled over the stubs; only built where /root/reference exists)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libbow_oracle.so")
_REF = os.path.join(_HERE, "_ref", "librefbow.so")
_REF_SRC = "/root/reference/src/rumi-slam/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

_u8p, _i32p, _f64p, _f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_double, C.c_float))


def _p(a, t):
    return a.ctypes.data_as(t)


def build():
    src = os.path.join(_HERE, "bow_oracle.cpp")
    if not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "libbow_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def build_ref():
    if os.path.exists(_REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "refbow"], stdout=subprocess.DEVNULL)
    return os.path.exists(_REF)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.bow_oracle_create.restype = C.c_void_p
        L.bow_oracle_create.argtypes = [C.c_int] * 5 + [_i32p, _u8p, _u8p, _f64p]
        L.bow_oracle_free.argtypes = [C.c_void_p]
        L.bow_oracle_transform.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _f64p, _i32p]
        L.bow_oracle_vectors.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _f64p, _i32p, _i32p, _i32p,
                                         C.POINTER(C.c_int)]
        L.bow_oracle_search_by_bow.argtypes = [_u8p, _f32p, _u8p, _i32p, _i32p, _i32p, C.c_int, _u8p, _f32p, C.c_int,
                                               _i32p, _i32p, _i32p, C.c_int, C.c_float, C.c_int, C.c_int, _i32p]
        L.bow_oracle_search_by_bow_kf.argtypes = [_u8p, _f32p, _u8p, C.c_int, _i32p, _i32p, _i32p, C.c_int, _u8p, _f32p,
                                                  _u8p, C.c_int, _i32p, _i32p, _i32p, C.c_int, C.c_float, C.c_int,
                                                  C.c_int, _i32p]
        L.bow_oracle_distinctive.argtypes = [_u8p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


class _Common:
    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        word, weight, node = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.int32)
        self._transform(self._h, _p(desc, _u8p), n, levelsup, _p(word, _i32p), _p(weight, _f64p), _p(node, _i32p))
        return word, weight, node

    def vectors(self, desc, levelsup=4):
        """Frame::ComputeBoW: (BowVector {word: value}, FeatureVector {node: [feature indices]})"""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        ids, vals = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        nodes, off, idx = np.zeros(max(n, 1), np.int32), np.zeros(n + 2, np.int32), np.zeros(max(n, 1), np.int32)
        nfv = C.c_int(0)
        k = self._vectors(self._h, _p(desc, _u8p), n, levelsup, _p(ids, _i32p), _p(vals, _f64p), _p(nodes, _i32p),
                          _p(off, _i32p), _p(idx, _i32p), C.byref(nfv))
        bow = {int(ids[i]): float(vals[i]) for i in range(k)}
        fv = {int(nodes[i]): [int(x) for x in idx[off[i]:off[i + 1]]] for i in range(nfv.value)}
        return bow, fv


class Vocabulary(_Common):
    """The restatement.  parent / is_leaf / desc / weight are per node in id order (node 0 = root)."""

    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
        Lb = lib()
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = np.ascontiguousarray(desc, np.uint8)
        self.weight = np.ascontiguousarray(weight, np.float64)
        self._h = Lb.bow_oracle_create(k, L, scoring, weighting, len(self.parent), _p(self.parent, _i32p),
                                       _p(self.is_leaf, _u8p), _p(self.desc, _u8p), _p(self.weight, _f64p))
        self._transform, self._vectors = Lb.bow_oracle_transform, Lb.bow_oracle_vectors


class ReferenceVocabulary(_Common):
    """The reference's own DBoW2, fed through its own loadFromTextFile."""

    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
        if not build_ref():
            raise RuntimeError("oracle/_ref/librefbow.so not built (reference sources absent)")
        R = C.CDLL(_REF)
        R.refbow_load.restype = C.c_void_p
        R.refbow_load.argtypes = [C.c_char_p]
        R.refbow_transform.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _f64p, _i32p]
        R.refbow_vectors.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _f64p, _i32p, _i32p, _i32p,
                                     C.POINTER(C.c_int)]
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write("%d %d %d %d\n" % (k, L, scoring, weighting))
            lines = []
            for nid in range(1, len(parent)):
                lines.append("%d %d %s %r" % (parent[nid], int(is_leaf[nid]), " ".join(str(int(b)) for b in desc[nid]),
                                              float(weight[nid])))
            f.write("\n".join(lines))          # no trailing newline: the loader's eof() loop would add an empty node
            path = f.name
        self._h = R.refbow_load(path.encode())
        os.unlink(path)
        if not self._h:
            raise RuntimeError("reference loadFromTextFile failed")
        self._R = R
        self._transform, self._vectors = R.refbow_transform, R.refbow_vectors


def flatten_fv(fv):
    nodes = np.array(sorted(fv), np.int32)
    off = np.zeros(len(nodes) + 1, np.int32)
    idx = []
    for i, n in enumerate(nodes):
        idx += list(fv[int(n)])
        off[i + 1] = len(idx)
    return nodes, off, np.array(idx if idx else [0], np.int32)[:max(len(idx), 0) or None] if idx else np.zeros(0, np.int32)


def search_by_bow(desc_kf, angle_kf, kf_valid, fv_kf, desc_f, angle_f, fv_f, nnratio=0.7, check_ori=True, th_low=50, n_left=-1):
    dk, df = np.ascontiguousarray(desc_kf, np.uint8), np.ascontiguousarray(desc_f, np.uint8)
    ak, af = np.ascontiguousarray(angle_kf, np.float32), np.ascontiguousarray(angle_f, np.float32)
    vk = np.ascontiguousarray(kf_valid, np.uint8)
    kn, ko, ki = flatten_fv(fv_kf)
    fn, fo, fi = flatten_fv(fv_f)
    ki = np.ascontiguousarray(np.append(ki, 0), np.int32)
    fi = np.ascontiguousarray(np.append(fi, 0), np.int32)
    match = np.zeros(max(len(df), 1), np.int32)
    L = lib()
    L.bow_oracle_search_by_bow_nleft.restype = C.c_int
    L.bow_oracle_search_by_bow_nleft.argtypes = [_u8p, _f32p, _u8p, _i32p, _i32p, _i32p, C.c_int, _u8p, _f32p, C.c_int, _i32p, _i32p,
                                                 _i32p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _i32p]
    n = L.bow_oracle_search_by_bow_nleft(_p(dk, _u8p), _p(ak, _f32p), _p(vk, _u8p), _p(kn, _i32p), _p(ko, _i32p),
                                         _p(ki, _i32p), len(kn), _p(df, _u8p), _p(af, _f32p), len(df), _p(fn, _i32p),
                                         _p(fo, _i32p), _p(fi, _i32p), len(fn), nnratio, 1 if check_ori else 0, th_low,
                                         int(n_left), _p(match, _i32p))
    return n, match[:len(df)]


def distinctive(desc, offsets):
    """MapPoint::ComputeDistinctiveDescriptors per map point (rows offsets[p]:offsets[p+1] of desc)."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(offsets) - 1
    best, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
    L = lib()
    for p in range(n):
        d = np.ascontiguousarray(desc[offsets[p]:offsets[p + 1]])
        m = C.c_int(0)
        best[p] = L.bow_oracle_distinctive(_p(d, _u8p), len(d), C.byref(m))
        med[p] = m.value
    return best, med


def search_by_bow_kf(desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nnratio=0.8, check_ori=True, th_low=50):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12): (nmatches, match12)."""
    d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    a1, a2 = np.ascontiguousarray(angle1, np.float32), np.ascontiguousarray(angle2, np.float32)
    v1, v2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
    n1, o1, i1 = flatten_fv(fv1)
    n2, o2, i2 = flatten_fv(fv2)
    i1 = np.ascontiguousarray(np.append(i1, 0), np.int32)
    i2 = np.ascontiguousarray(np.append(i2, 0), np.int32)
    match = np.zeros(max(len(d1), 1), np.int32)
    n = lib().bow_oracle_search_by_bow_kf(_p(d1, _u8p), _p(a1, _f32p), _p(v1, _u8p), len(d1), _p(n1, _i32p), _p(o1, _i32p),
                                          _p(i1, _i32p), len(n1), _p(d2, _u8p), _p(a2, _f32p), _p(v2, _u8p), len(d2),
                                          _p(n2, _i32p), _p(o2, _i32p), _p(i2, _i32p), len(n2), nnratio,
                                          1 if check_ori else 0, th_low, _p(match, _i32p))
    return n, match[:len(d1)]


def search_for_triangulation(desc1, angle1, has_mp1, stereo1, fv1, desc2, angle2, has_mp2, stereo2, x2, y2, octave2, fv2,
                             scale_factors2, ep, epi_ok, only_stereo=False, coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:806-1013, no second camera): (nmatches, match12).  epi_ok: [n1, n2]
    table of what epipolarConstrain returns for a pair."""
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    d1, d2 = u8(desc1).reshape(-1, 32), u8(desc2).reshape(-1, 32)
    n1, o1, i1 = flatten_fv(fv1)
    n2, o2, i2 = flatten_fv(fv2)
    i1 = np.ascontiguousarray(np.append(i1, 0), np.int32)
    i2 = np.ascontiguousarray(np.append(i2, 0), np.int32)
    match = np.zeros(max(len(d1), 1), np.int32)
    L = lib()
    L.bow_oracle_search_for_triangulation.restype = C.c_int
    L.bow_oracle_search_for_triangulation.argtypes = [_u8p, _f32p, _u8p, _u8p, C.c_int, _i32p, _i32p, _i32p, C.c_int, _u8p, _f32p,
                                                      _u8p, _u8p, _f32p, _f32p, _i32p, C.c_int, _i32p, _i32p, _i32p, C.c_int,
                                                      _f32p, C.c_float, C.c_float, C.c_int, C.c_int, _u8p, C.c_int, _i32p]
    n = L.bow_oracle_search_for_triangulation(
        _p(d1, _u8p), _p(f32(angle1), _f32p), _p(u8(has_mp1), _u8p), _p(u8(stereo1), _u8p), len(d1), _p(n1, _i32p), _p(o1, _i32p),
        _p(i1, _i32p), len(n1), _p(d2, _u8p), _p(f32(angle2), _f32p), _p(u8(has_mp2), _u8p), _p(u8(stereo2), _u8p),
        _p(f32(x2), _f32p), _p(f32(y2), _f32p), _p(np.ascontiguousarray(octave2, np.int32), _i32p), len(d2), _p(n2, _i32p),
        _p(o2, _i32p), _p(i2, _i32p), len(n2), _p(f32(scale_factors2), _f32p), float(ep[0]), float(ep[1]), int(only_stereo),
        int(coarse), _p(u8(epi_ok), _u8p), int(check_ori), _p(match, _i32p))
    return n, match[:len(d1)]
