#include "SparsePyrLK_accel.h"

#include <stdexcept>
#include <string>

#include "rumi_orb.h"

static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f layout");

static void flow_check(int rc) {
    if (rc != RUMI_OK) throw std::runtime_error(std::string("SparsePyrLKAccel: ") + rumi_last_error());
}

SparsePyrLKAccel::SparsePyrLKAccel(cv::Size winSize, int maxLevel, int maxCount, double epsilon,
                                   double minEigThreshold, int device) : ctx(nullptr) {
    if (winSize.width != winSize.height) throw std::runtime_error("SparsePyrLKAccel: square windows only");
    flow_check(rumi_flow_create(&ctx, device, winSize.width, maxLevel, maxCount, epsilon, (float)minEigThreshold));
}

SparsePyrLKAccel::~SparsePyrLKAccel() { rumi_flow_destroy(ctx); }

void SparsePyrLKAccel::setPrev(const cv::Mat& prevImg) {
    flow_check(rumi_flow_set_prev(ctx, prevImg.ptr(0), prevImg.cols, prevImg.rows, (size_t)prevImg.step));
}

void SparsePyrLKAccel::trackNext(const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
                                 std::vector<cv::Point2f>& nextPts, std::vector<uchar>& status,
                                 std::vector<float>& err, bool advance) {
    const int n = (int)prevPts.size();
    nextPts.resize(n); status.resize(n); err.resize(n);
    flow_check(rumi_flow_track_next(ctx, nextImg.ptr(0), (size_t)nextImg.step,
                                    n ? reinterpret_cast<const float*>(prevPts.data()) : nullptr, n,
                                    n ? reinterpret_cast<float*>(nextPts.data()) : nullptr, status.data(), err.data(),
                                    advance ? 1 : 0));
}

void SparsePyrLKAccel::calc(const cv::Mat& prevImg, const cv::Mat& nextImg, const std::vector<cv::Point2f>& prevPts,
                            std::vector<cv::Point2f>& nextPts, std::vector<uchar>& status, std::vector<float>& err) {
    if (prevImg.rows != nextImg.rows || prevImg.cols != nextImg.cols)
        throw std::runtime_error("SparsePyrLKAccel: frames differ in size");
    setPrev(prevImg);
    trackNext(nextImg, prevPts, nextPts, status, err, false);
}
