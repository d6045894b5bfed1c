"""ORACLE -- TEST INFRASTRUCTURE ONLY: ctypes bindings of oracle/libflow_oracle.so, the C++ restatement of the sparse
pyramidal Lucas-Kanade flow behind KFDSample::Step (flow_oracle.cpp; pinned against cv2 in tests/test_flow_oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libflow_oracle.so")
_u8p, _i16p, _f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_int16, C.c_float))


def _p(a, t):
    return a.ctypes.data_as(t)


def build():
    src = os.path.join(_HERE, "flow_oracle.cpp")
    if not os.path.exists(_LIB) or os.path.getmtime(src) > os.path.getmtime(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "libflow_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.flow_pyr_down.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.flow_scharr.argtypes = [_u8p, C.c_int, C.c_int, _i16p]
        L.flow_lk.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                              C.c_float, _f32p, _u8p, _f32p]
        L.flow_mean_magnitude.restype = C.c_float
        L.flow_mean_magnitude.argtypes = [_f32p, _f32p, _u8p, C.c_int, C.POINTER(C.c_int)]
        L.flow_pd_update.restype = C.c_float
        L.flow_pd_update.argtypes = [_f32p] + [C.c_float] * 6 + [C.c_double]
        _lib = L
    return _lib


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().flow_pyr_down(_p(img, _u8p), w, h, _p(out, _u8p))
    return out


def scharr(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().flow_scharr(_p(img, _u8p), w, h, _p(out, _i16p))
    return out


def lk(prev, nxt, pts, win=31, max_level=2, max_count=20, eps=0.03, min_eig=1e-4):
    """calcOpticalFlowPyrLK(prev, next, pts, Size(win, win), max_level, (COUNT+EPS, max_count, eps)) ->
    (next_pts [n,2] f32, status [n] u8, err [n] f32)."""
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    h, w = prev.shape
    n = len(pts)
    out = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    lib().flow_lk(_p(prev, _u8p), _p(nxt, _u8p), w, h, _p(pts, _f32p), n, win, max_level, max_count, eps, min_eig,
                  _p(out, _f32p), _p(st, _u8p), _p(err, _f32p))
    return out, st, err


def mean_magnitude(old, nxt, status):
    old = np.ascontiguousarray(old, np.float32)
    nxt = np.ascontiguousarray(nxt, np.float32)
    status = np.ascontiguousarray(status, np.uint8)
    g = C.c_int(0)
    v = lib().flow_mean_magnitude(_p(old, _f32p), _p(nxt, _f32p), _p(status, _u8p), len(status), C.byref(g))
    return float(v), g.value


class PD:
    """PD controller of the key-frame selector (pd.hpp), float32 state."""

    def __init__(self, kp=0.8, kd=0.005, setpoint=10.0, alpha=1.0, max_output=255.0):
        self.kp, self.kd, self.setpoint, self.alpha, self.max_output = kp, kd, setpoint, alpha, max_output
        self.state = np.zeros(1, np.float32)

    def update(self, value, ts):
        return float(lib().flow_pd_update(_p(self.state, _f32p), self.kp, self.kd, self.alpha, self.setpoint,
                                          self.max_output, value, ts))
