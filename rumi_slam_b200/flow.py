"""Host-side mirror of the optical-flow key-frame sampler of RUMI-SLAM over the C ABI.

`SparsePyrLK` = cv::calcOpticalFlowPyrLK as KFDSample::Step calls it (R/lib_src/KFDSample.cc:131-132: winSize 31x31,
maxLevel 2, TermCriteria(COUNT + EPS, 20, 0.03), R/include/cloud_edge_slam_lib/KFDSample.h:47) on the device
(rumi_flow_*).  `KFDSample` keeps the reference class's interface and control flow (Step / Reset / SetThreshold /
SetPDKFselectorParams / GetKF / GetAllKF); its host logic (SelectGoodPts, Calmoptflmag, the PD controller of pd.hpp)
stays on the host in float32, exactly as written in the reference.  No CPU implementation of the flow exists here.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr
from .extractor import ORBextractor


class SparsePyrLK:
    def __init__(self, win=31, max_level=2, max_count=20, epsilon=0.03, min_eig_threshold=1e-4, device=0):
        self._L = _lib.lib()
        self._f = C.c_void_p()
        check(self._L.rumi_flow_create(C.byref(self._f), int(device), int(win), int(max_level), int(max_count),
                                       float(epsilon), float(min_eig_threshold)))
        self.win, self.shape = int(win), None

    def close(self):
        if getattr(self, "_f", None) is not None and self._f:
            self._L.rumi_flow_destroy(self._f)
            self._f = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _img(img):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 2:
            raise ValueError("CV_8UC1 image expected")
        if img.strides[1] != 1:
            img = np.ascontiguousarray(img)
        return img

    def set_prev(self, img):
        img = self._img(img)
        h, w = img.shape
        check(self._L.rumi_flow_set_prev(self._f, ptr(img) if img.size else None, w, h, img.strides[0]))
        self.shape = (h, w)

    def track_next(self, img, pts, advance=False):
        """Tracks `pts` (n x 2 float32) from the held previous frame into `img` -> (next_pts, status, err)."""
        img = self._img(img)
        if self.shape is not None and img.shape != self.shape:
            raise ValueError("frame size differs from the previous frame")
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        nxt = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        check(self._L.rumi_flow_track_next(self._f, ptr(img) if img.size else None, img.strides[0], ptr(pts), n,
                                           ptr(nxt), ptr(st), ptr(err), int(bool(advance))))
        return nxt, st, err

    def calc(self, prev, nxt, pts):
        """cv::calcOpticalFlowPyrLK(prev, next, pts, ...) -> (next_pts, status, err)."""
        self.set_prev(prev)
        return self.track_next(nxt, pts, advance=False)

    def levels(self):
        return int(self._L.rumi_flow_levels(self._f))

    def launches(self):
        return int(self._L.rumi_flow_launches(self._f))

    def timer_start(self):
        check(self._L.rumi_flow_timer_start(self._f))

    def timer_stop(self):
        ms = C.c_float(0)
        check(self._L.rumi_flow_timer_stop(self._f, C.byref(ms)))
        return ms.value

    # ---- parity hooks ----
    def pyramid_level(self, level, which=0):
        w, h = C.c_int32(0), C.c_int32(0)
        check(self._L.rumi_flow_debug_level(self._f, which, level, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(self._L.rumi_flow_debug_level(self._f, which, level, ptr(out), C.byref(w), C.byref(h)))
        return out

    def derivatives(self, level):
        h, w = self.pyramid_level(level).shape
        out = np.zeros((h, w, 2), np.int16)
        check(self._L.rumi_flow_debug_deriv(self._f, level, ptr(out)))
        return out


class PD:
    """R/include/cloud_edge_slam_lib/pd.hpp: PD controller with float32 state (Alpha = 1, maxOutput = 255 as
    constructed by KFDSample::InitPDKFselector, R/lib_src/KFDSample.cc:26-35)."""

    def __init__(self, kp, kd, alpha=1.0, max_output=255.0):
        f = np.float32
        self.kp, self.kd, self.alpha, self.max_output = f(kp), f(kd), f(alpha), f(max_output)
        self.setpoint, self.prev_input = f(0), f(0)

    def update(self, value, ts):                                   # pd.hpp:21-40
        f = np.float32
        value = f(value)
        error = f(self.setpoint - value)
        diff = f(self.alpha * f(self.prev_input - value))
        self.prev_input = f(self.prev_input - diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = f(np.float64(f(self.kp * error)) + np.float64(self.kd) / np.float64(ts) * np.float64(diff))
        return self.max_output if out > self.max_output else out


def mean_flow_magnitude(old, nxt):
    """Calmoptflmag (R/lib_src/KFDSample.cc:181-193): float32 sum of |next - old| in index order / n (NaN for n = 0)."""
    f = np.float32
    d = (nxt.astype(np.float32) - old.astype(np.float32))
    mag = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32)).astype(np.float32)
    s = f(0)
    for m in mag:
        s = f(s + m)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f(s / f(len(mag)))


class KFDSample:
    """Mirror of the reference's optical-flow key-frame sampler (R/include/cloud_edge_slam_lib/KFDSample.h:21-84,
    R/lib_src/KFDSample.cc): Step(image, timestamp) -> True when the frame is selected as a key frame."""

    def __init__(self, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, Kp=0.8, Kd=0.005,
                 th=10.0, device=0):
        self.mpORBextractor = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=device)
        self.flow = SparsePyrLK(31, 2, 20, 0.03, device=device)
        self.Kp, self.Kd, self.th = Kp, Kd, th
        self.mpPDcontroller = PD(Kp, Kd)
        self.mpPDcontroller.setpoint = np.float32(th)
        self.vLapping = (0, 0)                                     # KFDSample.h:53
        self.old = np.zeros((0, 2), np.float32)
        self.next = self.status = self.err = None
        self.mvKeys = self.mDescriptors = None
        self.ltframe = 0.0
        self.moptf = np.float32(0)
        self.KFset = []

    def SetPDKFselectorParams(self, Kp, Kd, th):                   # KFDSample.cc:37-44
        self.Kp, self.Kd, self.th = Kp, Kd, th
        self.mpPDcontroller.kp, self.mpPDcontroller.kd = np.float32(Kp), np.float32(Kd)
        self.mpPDcontroller.setpoint = np.float32(th)

    def SetThreshold(self, th):                                    # KFDSample.cc:56-58
        self.th = th

    def GetKF(self):
        return self.KFset[-1]

    def GetAllKF(self):
        return list(self.KFset)

    def Reset(self):                                               # KFDSample.cc:84-86
        self.old = np.zeros((0, 2), np.float32)

    def _extract(self, im):
        _, self.mvKeys, self.mDescriptors = self.mpORBextractor(im, None, self.vLapping)
        self.old = np.stack([self.mvKeys["x"], self.mvKeys["y"]], 1).astype(np.float32)   # KeyPoint::convert

    def Step(self, inputIm, timeStamp):                            # KFDSample.cc:88-175
        im = np.ascontiguousarray(inputIm)
        if im.ndim == 3:                                           # the reference converts BGR to grey (:101-102)
            raise ValueError("pass a CV_8UC1 image (colour conversion is the caller's cv::cvtColor)")
        if len(self.old) == 0:                                     # initial frame (:109-127)
            self.ltframe = timeStamp
            self.flow.set_prev(im)
            self._extract(im)
            self.KFset.append(im.copy())
            return True
        # next frame becomes the previous one whatever the decision (:169), so advance on the device right away
        self.next, self.status, self.err = self.flow.track_next(im, self.old, advance=True)
        good = self.status == 1                                    # SelectGoodPts (:75-82)
        self.moptf = mean_flow_magnitude(self.old[good], self.next[good])
        vPDKFth = self.mpPDcontroller.update(self.moptf, timeStamp - self.ltframe)
        TH = np.float32(self.moptf + vPDKFth)
        result = False
        if self.moptf > TH:                                        # (:151-165)
            self._extract(im)
            self.KFset.append(im.copy())
            result = True
        else:
            self.old = self.next
        self.ltframe = timeStamp
        return result
