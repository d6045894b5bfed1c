"""Host-side mirror of the Hamming matching core of ORB_SLAM3::ORBmatcher
(R/include/cloud_edge_slam_lib/ORBmatcher.h:36-103, R/lib_src/ORBmatcher.cc:31-33, :1830-1844) over the C ABI.

The GPU returns the raw (best index, best distance, second-best distance) triple of the scan every reference
matcher runs; the call-site specific acceptance tests (TH_LOW / TH_HIGH, '<' vs '<=', which side of the ratio
is cast to float) stay here on the host exactly as written at each call site.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class ORBmatcher:
    TH_HIGH = 100        # ORBmatcher.cc:31
    TH_LOW = 50          # ORBmatcher.cc:32
    HISTO_LENGTH = 30    # ORBmatcher.cc:33

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self._L = _lib.lib()
        self._m = C.c_void_p()
        check(self._L.rumi_match_create(C.byref(self._m), int(device)))
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device

    def close(self):
        if getattr(self, "_m", None) is not None and self._m:
            self._L.rumi_match_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a, b):
        """ORBmatcher::DescriptorDistance (ORBmatcher.cc:1830-1844) for one pair of 32-byte rows."""
        a = np.ascontiguousarray(a, np.uint8).reshape(32)
        b = np.ascontiguousarray(b, np.uint8).reshape(32)
        return _lib.lib().rumi_descriptor_distance(ptr(a), ptr(b))

    # ---- raw top-2 (host buffers) ----
    def top2(self, Q, T):
        Q = np.ascontiguousarray(Q, np.uint8).reshape(-1, 32)
        T = np.ascontiguousarray(T, np.uint8).reshape(-1, 32)
        i1 = np.zeros(len(Q), np.int32)
        d1 = np.zeros(len(Q), np.uint16)
        d2 = np.zeros(len(Q), np.uint16)
        check(self._L.rumi_hamming_top2(self._m, ptr(Q), len(Q), ptr(T), len(T), ptr(i1), ptr(d1), ptr(d2)))
        return i1, d1, d2

    # ---- raw top-2 (CUDA torch tensors; train indices offset by t_base for a train shard) ----
    def top2_device(self, Q, T, t_base=0, out=None, sync=True):
        import torch
        assert Q.is_cuda and T.is_cuda and Q.dtype == torch.uint8 and T.dtype == torch.uint8
        nq, nt = Q.shape[0], T.shape[0]
        if out is None:
            out = (torch.empty(nq, dtype=torch.int32, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device),
                   torch.empty(nq, dtype=torch.int16, device=Q.device))
        i1, d1, d2 = out
        check(self._L.rumi_hamming_top2_device(self._m, ptr(Q), nq, ptr(T), nt, int(t_base), ptr(i1), ptr(d1), ptr(d2),
                                               1 if sync else 0))
        return i1, d1, d2

    def pack_device(self, i1, d1, d2, out=None, sync=True):
        import torch
        nq = i1.shape[0]
        if out is None:
            out = torch.empty(nq, dtype=torch.int64, device=i1.device)
        check(self._L.rumi_top2_pack_device(self._m, ptr(i1), ptr(d1), ptr(d2), nq, ptr(out), 1 if sync else 0))
        return out

    def merge_device(self, packed, nshards, nq, out=None, sync=True):
        import torch
        if out is None:
            out = (torch.empty(nq, dtype=torch.int32, device=packed.device),
                   torch.empty(nq, dtype=torch.int16, device=packed.device),
                   torch.empty(nq, dtype=torch.int16, device=packed.device))
        i1, d1, d2 = out
        check(self._L.rumi_top2_merge_device(self._m, ptr(packed), int(nshards), int(nq), ptr(i1), ptr(d1), ptr(d2),
                                             1 if sync else 0))
        return i1, d1, d2

    def timer_start(self):
        check(self._L.rumi_match_timer_start(self._m))

    def timer_stop(self):
        ms = C.c_float(0)
        check(self._L.rumi_match_timer_stop(self._m, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        return int(self._L.rumi_match_launch_count(self._m, 1 if reset else 0))

    # ---- acceptance rules of the reference call sites, applied to the raw triple ----
    def accept_bow(self, d1, d2, th=None):
        """SearchByBoW KF->F (ORBmatcher.cc:290-291): best1 <= TH_LOW and (float)best1 < ratio*(float)best2."""
        th = self.TH_LOW if th is None else th
        d1 = np.asarray(d1).astype(np.int64)
        d2 = np.asarray(d2).astype(np.int64)
        return (d1 <= th) & (d1.astype(np.float32) < self.mfNNratio * d2.astype(np.float32))

    @staticmethod
    def accept_knn_ratio(d1, d2, ratio=0.7):
        """ComputeStereoFishEyeMatches (Frame.cc:1146): d0 < d1 * 0.7 (float distance, double constant)."""
        d1 = np.asarray(d1).astype(np.float32).astype(np.float64)
        d2 = np.asarray(d2).astype(np.float32).astype(np.float64)
        return d1 < d2 * ratio

    def match_bow(self, Q, T):
        i1, d1, d2 = self.top2(Q, T)
        ok = self.accept_bow(d1, d2) & (i1 >= 0)
        return np.where(ok, i1, -1), d1, d2

    # ---- stereo row-band best-1 (Frame.cc:828-905) ----
    def stereo_best1(self, Lk, Ld, Rk, Rd, scale_factors, n_rows, min_d, max_d):
        Lk = np.ascontiguousarray(Lk, KP_DTYPE)
        Rk = np.ascontiguousarray(Rk, KP_DTYPE)
        Ld = np.ascontiguousarray(Ld, np.uint8)
        Rd = np.ascontiguousarray(Rd, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        best = np.zeros(len(Lk), np.int32)
        dist = np.zeros(len(Lk), np.uint16)
        check(self._L.rumi_stereo_best1(self._m, ptr(Lk), ptr(Ld), len(Lk), ptr(Rk), ptr(Rd), len(Rk), ptr(sf),
                                        len(sf), int(n_rows), float(min_d), float(max_d), ptr(best), ptr(dist)))
        return best, dist

    # ---- Frame::ComputeStereoMatches, complete (Frame.cc:828-985) ----
    def stereo_match(self, ex_left, ex_right, Lk, Ld, Rk, Rd, mbf, mb):
        """ex_left / ex_right: the ORBextractor objects whose LAST call produced (Lk, Ld) / (Rk, Rd); their device
        pyramids are used in place.  Returns (mvuRight, mvDepth, number of matches kept)."""
        Lk = np.ascontiguousarray(Lk, KP_DTYPE)
        Rk = np.ascontiguousarray(Rk, KP_DTYPE)
        Ld = np.ascontiguousarray(Ld, np.uint8)
        Rd = np.ascontiguousarray(Rd, np.uint8)
        u = np.zeros(len(Lk), np.float32)
        d = np.zeros(len(Lk), np.float32)
        n = C.c_int32(0)
        check(self._L.rumi_stereo_match(self._m, ex_left._h, ex_right._h, ptr(Lk), ptr(Ld), len(Lk), ptr(Rk), ptr(Rd),
                                        len(Rk), float(mbf), float(mb), ptr(u), ptr(d), C.byref(n)))
        return u, d, n.value
