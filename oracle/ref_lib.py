"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/liborb_ref.so.

liborb_ref.so = the UNMODIFIED reference R/lib_src/ORBextractor.cc compiled over oracle/cvstub (see
oracle/Makefile, oracle/ref_shim.cpp).  Built only where /root/reference exists; travels to the GPU box as a
prebuilt file.  Used to pin oracle/orb_oracle.cpp's control flow (octree / IC_Angle / rBRIEF / assembly) and
as bench.py's `--impl reference` CPU arm.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .orb_oracle import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "liborb_ref.so")
REF_SRC = "/root/reference/src/rumi-slam/lib_src/ORBextractor.cc"


def build():
    if os.path.exists(REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(_LIB)


def available():
    return os.path.exists(_LIB) or build()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/liborb_ref.so not built (reference sources absent)")
        L = C.CDLL(_LIB)
        u8p, f32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.ref_extract.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, u8p, C.c_int, i32p, i32p]
        L.ref_octree.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int]
        L.ref_describe.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, u8p]
        L.ref_tables.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, f32p, f32p, f32p, f32p]
        _lib = L
    return _lib


def extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, lapping=(0, 0)):
    img = np.ascontiguousarray(img, np.uint8)
    cap = nfeatures + 64 * nlevels
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n, m = C.c_int32(0), C.c_int32(0)
    rc = lib().ref_extract(img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[1], img.shape[0], img.strides[0],
                           nfeatures, scale, nlevels, ini, mn, int(lapping[0]), int(lapping[1]), kps.ctypes.data,
                           desc.ctypes.data_as(C.POINTER(C.c_uint8)), cap, C.byref(n), C.byref(m))
    if rc != 0:
        return None
    return kps[:n.value].copy(), desc[:n.value].copy(), m.value


def describe(img, kps):
    """The reference's ORBextractor::CloudFrameComputeDescriptors: (return value, descriptors)."""
    img = np.ascontiguousarray(img, np.uint8)
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    desc = np.zeros((len(kps), 32), np.uint8)
    h, w = (img.shape if img.ndim == 2 and img.size else (0, 0))
    rc = lib().ref_describe(img.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, img.strides[0] if img.size else 0,
                            kps.ctypes.data, len(kps), desc.ctypes.data_as(C.POINTER(C.c_uint8)))
    return rc, desc


def octree(xyr, min_x, max_x, min_y, max_y, n_target):
    xyr = np.ascontiguousarray(xyr, np.float32)
    cap = n_target + 16 + len(xyr)
    out = np.zeros((cap, 3), np.float32)
    n = lib().ref_octree(xyr.ctypes.data_as(C.POINTER(C.c_float)), len(xyr), min_x, max_x, min_y, max_y, n_target,
                         out.ctypes.data_as(C.POINTER(C.c_float)), cap)
    return out[:n]


def tables(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7):
    a = [np.zeros(nlevels, np.float32) for _ in range(4)]
    p = [x.ctypes.data_as(C.POINTER(C.c_float)) for x in a]
    lib().ref_tables(nfeatures, scale, nlevels, ini, mn, *p)
    return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3])
