// Drop-in for the cv::calcOpticalFlowPyrLK call of KFDSample::Step (R/lib_src/KFDSample.cc:131-132):
//     calcOpticalFlowPyrLK(imprvs, imnext, old, next, status, err, Size(31, 31), 2, criteria);
// with criteria = TermCriteria(COUNT + EPS, 20, 0.03) (R/include/cloud_edge_slam_lib/KFDSample.h:47), on the device
// behind the C ABI (rumi_flow_*, include/rumi_orb.h).  KFDSample keeps its members and control flow; it owns one
// SparsePyrLKAccel and replaces that one line by
//     mFlow.calc(imprvs, imnext, old, next, status, err);
// or, to upload each frame only once (imprvs is always the previous imnext, KFDSample.cc:169), by
//     mFlow.trackNext(imnext, old, next, status, err, /*advance=*/true);      // after mFlow.setPrev(first frame)
// No CPU fallback: without a CUDA device the constructor throws.
#ifndef SPARSEPYRLK_ACCEL_H
#define SPARSEPYRLK_ACCEL_H

#include <vector>
#include <opencv2/core/core.hpp>

struct rumi_flow;

class SparsePyrLKAccel {
public:
    // winSize must be square: 31 (the reference), 21 (OpenCV's default) or 15.  maxCount / epsilon are the two fields of
    // the TermCriteria the reference passes (both active: COUNT + EPS).
    explicit SparsePyrLKAccel(cv::Size winSize = cv::Size(31, 31), int maxLevel = 2, int maxCount = 20,
                              double epsilon = 0.03, double minEigThreshold = 1e-4, int device = 0);
    ~SparsePyrLKAccel();
    SparsePyrLKAccel(const SparsePyrLKAccel&) = delete;
    SparsePyrLKAccel& operator=(const SparsePyrLKAccel&) = delete;

    // == cv::calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status, err, winSize, maxLevel, criteria)
    // for CV_8UC1 images of equal size.  nextPts / status / err are resized to prevPts.size().
    void calc(const cv::Mat& prevImg, const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
              std::vector<cv::Point2f>& nextPts, std::vector<uchar>& status, std::vector<float>& err);

    void setPrev(const cv::Mat& prevImg);
    void trackNext(const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts, std::vector<cv::Point2f>& nextPts,
                   std::vector<uchar>& status, std::vector<float>& err, bool advance);

private:
    rumi_flow* ctx;
};

#endif
