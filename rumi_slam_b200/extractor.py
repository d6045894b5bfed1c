"""Host-side mirror of ORB_SLAM3::ORBextractor (R/include/cloud_edge_slam_lib/ORBextractor.h:42-111) over the C ABI.

Same constructor arguments, same call operator meaning (image, mask, keypoints, descriptors, vLappingArea ->
monoIndex), same getters, same error behaviour (empty image -> -1).  All arithmetic runs in librumi_orb.so on the
GPU; this file only marshals numpy / torch buffers.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, RumiError, check, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0, max_batch=1):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        check(self._L.rumi_orb_create(C.byref(self._h), int(nfeatures), float(scaleFactor), int(nlevels),
                                      int(iniThFAST), int(minThFAST), int(device), int(max_batch)))
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(np.float32(scaleFactor)), int(nlevels)
        self.iniThFAST, self.minThFAST, self.device, self.max_batch = int(iniThFAST), int(minThFAST), device, max_batch
        t = [np.zeros(nlevels, np.float32) for _ in range(4)]
        q = np.zeros(nlevels, np.int32)
        check(self._L.rumi_orb_tables(self._h, *[a.ctypes.data_as(C.POINTER(C.c_float)) for a in t],
                                      q.ctypes.data_as(C.POINTER(C.c_int32))))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = t
        self.mnFeaturesPerLevel = q
        self._last_shape = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.rumi_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (ORBextractor.h:62-84) ----
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self.mvScaleFactor.copy()

    def GetInverseScaleFactors(self):
        return self.mvInvScaleFactor.copy()

    def GetScaleSigmaSquares(self):
        return self.mvLevelSigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self.mvInvLevelSigma2.copy()

    def frame_capacity(self, w, h):
        return check(self._L.rumi_orb_frame_capacity(self._h, int(w), int(h)))

    # ---- operator() (ORBextractor.cc:1014-1091) ----
    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints[KP_DTYPE], descriptors[n,32] u8); monoIndex == -1 for an empty image."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        cap = self.frame_capacity(w, h)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, m = C.c_int32(0), C.c_int32(0)
        rc = self._L.rumi_orb_extract(self._h, ptr(image), w, h, image.strides[0], int(vLappingArea[0]),
                                      int(vLappingArea[1]), ptr(kps), ptr(desc), cap, C.byref(n), C.byref(m))
        if rc == -1:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        check(rc)
        self._last_shape = (h, w)
        return m.value, kps[:n.value], desc[:n.value]

    def begin(self, image, vLappingArea=(0, 0)):
        """First half of __call__: enqueues the frame and returns at once (rumi_orb_extract_begin).  Lets ONE host thread
        keep several extractors busy (left / right image of a stereo frame); finish with end()."""
        assert image.dtype == np.uint8 and image.ndim == 2 and image.size, "CV_8UC1 expected"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        self._pending = (image, h, w)                 # the upload reads the image until end()
        check(self._L.rumi_orb_extract_begin(self._h, ptr(image), w, h, image.strides[0], int(vLappingArea[0]),
                                             int(vLappingArea[1])))

    def end(self):
        """Second half of __call__: waits for the frame of begin() and returns (monoIndex, keypoints, descriptors)."""
        _, h, w = self._pending
        self._pending = None
        cap = self.frame_capacity(w, h)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, m = C.c_int32(0), C.c_int32(0)
        check(self._L.rumi_orb_extract_end(self._h, ptr(kps), ptr(desc), cap, C.byref(n), C.byref(m)))
        self._last_shape = (h, w)
        return m.value, kps[:n.value], desc[:n.value]

    def extract_batch(self, images, vLappingArea=(0, 0), out=None):
        """images: [n,h,w] uint8 HOST array (pinned memory makes the copies asynchronous).
        Returns (kps[n,cap], desc[n,cap,32], n_kp[n], n_mono[n]) host arrays."""
        if hasattr(images, "data_ptr"):          # torch CPU tensor
            n, h, w = images.shape
            stride, pitch = images.stride(1), images.stride(0)
        else:
            assert images.dtype == np.uint8 and images.ndim == 3
            n, h, w = images.shape
            stride, pitch = images.strides[1], images.strides[0]
        cap = self.frame_capacity(w, h)
        if out is None:
            out = (np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32),
                   np.zeros(n, np.int32))
        kps, desc, nkp, nmono = out
        check(self._L.rumi_orb_extract_batch(self._h, ptr(images), n, w, h, stride, pitch, int(vLappingArea[0]),
                                             int(vLappingArea[1]), ptr(kps), ptr(desc), cap, ptr(nkp), ptr(nmono)))
        self._last_shape = (h, w)
        return kps, desc, nkp, nmono

    def extract_batch_device(self, images, vLappingArea=(0, 0), out=None, sync=True):
        """images: [n,h,w] uint8 CUDA torch tensor.  Results stay on the device:
        (kps float32 view [n,cap,7], desc [n,cap,32] u8, n_kp [n] i32, n_mono [n] i32)."""
        import torch
        assert images.is_cuda and images.dtype == torch.uint8 and images.dim() == 3
        n, h, w = images.shape
        cap = self.frame_capacity(w, h)
        if out is None:
            dev = images.device
            out = (torch.zeros((n, cap, 7), dtype=torch.float32, device=dev),
                   torch.zeros((n, cap, 32), dtype=torch.uint8, device=dev),
                   torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev))
        kps, desc, nkp, nmono = out
        # the library runs on its own streams: order them behind torch's current stream (which produced `images` and
        # zero-filled `out`) and, for asynchronous calls, order torch's stream behind the extraction
        st = _lib.torch_stream()
        check(self._L.rumi_orb_wait_stream(self._h, st))
        check(self._L.rumi_orb_extract_batch_device(self._h, ptr(images), n, w, h, images.stride(1), images.stride(0),
                                                    int(vLappingArea[0]), int(vLappingArea[1]), ptr(kps), ptr(desc),
                                                    cap, ptr(nkp), ptr(nmono), 1 if sync else 0))
        if not sync:
            check(self._L.rumi_orb_signal_stream(self._h, st))
        self._last_shape = (h, w)
        return kps, desc, nkp, nmono

    # ---- CloudFrameComputeDescriptors (ORBextractor.cc:989-1011) ----
    def CloudFrameComputeDescriptors(self, image, keypoints):
        if image is None or image.size == 0:
            return -1, np.zeros((0, 32), np.uint8)
        image = np.ascontiguousarray(image, np.uint8)
        kps = np.ascontiguousarray(keypoints, KP_DTYPE)
        desc = np.zeros((len(kps), 32), np.uint8)
        n = check(self._L.rumi_orb_describe(self._h, ptr(image), image.shape[1], image.shape[0], image.strides[0],
                                            ptr(kps), len(kps), ptr(desc)))
        return n, desc

    def CloudFrameComputeDescriptorsBatch(self, images, keypoints, kp_off):
        """CloudFrameComputeDescriptors for many key frames of one shape in ONE call: images [n, h, w] u8, keypoints of
        frame f = keypoints[kp_off[f]:kp_off[f + 1]].  Returns the [total, 32] descriptors."""
        images = np.ascontiguousarray(images, np.uint8)
        kps = np.ascontiguousarray(keypoints, KP_DTYPE)
        off = np.ascontiguousarray(kp_off, np.int32)
        desc = np.zeros((len(kps), 32), np.uint8)
        n, h, w = images.shape
        check(self._L.rumi_orb_describe_batch(self._h, ptr(images), n, w, h, images.strides[1], images.strides[0], ptr(kps),
                                              ptr(off), ptr(desc)))
        return desc

    # ---- mvImagePyramid (ORBextractor.h:86): filled lazily from the device ----
    def _level(self, fn, level):
        w, h = C.c_int32(0), C.c_int32(0)
        check(fn(self._h, level, None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(fn(self._h, level, ptr(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    @property
    def mvImagePyramid(self):
        return [self._level(self._L.rumi_orb_pyramid_level, l) for l in range(self.nlevels)]

    def blurred_pyramid(self):
        return [self._level(self._L.rumi_orb_blurred_level, l) for l in range(self.nlevels)]

    # ---- measurement hooks (bench.py) ----
    STAGES = ("pyramid", "fast", "octree", "slots", "blur", "describe", "h2d", "d2h")

    def timer_start(self):
        check(self._L.rumi_orb_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        check(self._L.rumi_orb_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_streams(self, n):
        """n workspaces / streams for consecutive chunks (0 = default); 1 makes per-stage event times exclusive."""
        check(self._L.rumi_orb_set_streams(self._h, int(n)))

    def profile(self, enable=True):
        check(self._L.rumi_orb_profile(self._h, 1 if enable else 0))

    def profile_read(self, reset=True):
        ms = np.zeros(8, np.float64)
        n = np.zeros(8, np.int64)
        k = check(self._L.rumi_orb_profile_read(self._h, ptr(ms), ptr(n), 1 if reset else 0))
        return {self.STAGES[i]: (float(ms[i]), int(n[i])) for i in range(k)}

    def launch_count(self, reset=False):
        return int(self._L.rumi_orb_launch_count(self._h, 1 if reset else 0))

    def debug_candidates(self, level, selected=False):
        fn = self._L.rumi_orb_debug_selected if selected else self._L.rumi_orb_debug_candidates
        n = check(fn(self._h, level, None, 0))
        out = np.zeros((max(n, 1), 3), np.int32)
        check(fn(self._h, level, ptr(out), n))
        return out[:n]
