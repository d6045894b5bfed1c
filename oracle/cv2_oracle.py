"""ORACLE A -- TEST INFRASTRUCTURE ONLY: the OpenCV primitives the reference calls, via cv2.

The reference's hot path calls cv::resize / cv::FAST / cv::fastAtan2 / cv::GaussianBlur
(R/lib_src/ORBextractor.cc:1103, :767/:784, :96, :1058) and cv::BFMatcher::knnMatch (R/lib_src/Frame.cc:1139).
This module calls the SAME OpenCV algorithms through the container's cv2 wheel so that
oracle/orb_oracle.cpp's integer restatements can be pinned against them (tests/test_oracle_pin.py,
tools/gen_golden.py).  cv2 is optional at test time: everything pinned here is also frozen into
tests/golden/*.npz.
"""
import numpy as np

try:
    import cv2
    cv2.setNumThreads(1)
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False


def pyramid(img, ws, hs):
    out = [img.copy()]
    for l in range(1, len(ws)):
        out.append(cv2.resize(out[-1], (int(ws[l]), int(hs[l])), interpolation=cv2.INTER_LINEAR))
    return out


def fast(img, th):
    det = cv2.FastFeatureDetector_create(int(th), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(np.ascontiguousarray(img))
    return np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.float32).reshape(-1, 3)


def grid_fast(img, ini=20, mn=7):
    """R/lib_src/ORBextractor.cc:726-808 with cv2's FAST per cell."""
    rows, cols = img.shape
    minB = 16
    maxBX, maxBY = cols - 16, rows - 16
    width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
    nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
    wCell = int(np.ceil(width / np.float32(nCols)))
    hCell = int(np.ceil(height / np.float32(nRows)))
    out, nfb = [], 0
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = np.ascontiguousarray(img[iniY:maxY, iniX:maxX])
            k = fast(cell, ini)
            if len(k) == 0:
                k = fast(cell, mn)
                nfb += 1
            if len(k):
                k = k.copy()
                k[:, 0] += j * wCell
                k[:, 1] += i * hCell
                out.append(k)
    return (np.concatenate(out) if out else np.zeros((0, 3), np.float32)), nfb


def fast_atan2(y, x):
    return cv2.fastAtan2(float(y), float(x))


def blur(img):
    return cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)


def knn2(Q, T):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) -> (idx1, d1, d2); distances only are order-independent."""
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(np.ascontiguousarray(Q), np.ascontiguousarray(T), 2)
    i1 = np.array([a[0].trainIdx for a in m], np.int32)
    d1 = np.array([a[0].distance for a in m], np.int32)
    d2 = np.array([a[1].distance if len(a) > 1 else 256 for a in m], np.int32)
    return i1, d1, d2
