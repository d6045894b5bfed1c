#include "ORBmatcher_accel.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "rumi_orb.h"

namespace ORB_SLAM3 {

static std::vector<uint8_t> rows32(const cv::Mat& m) {
    std::vector<uint8_t> out(32 * (size_t)m.rows);
    for (int i = 0; i < m.rows; ++i) std::memcpy(out.data() + 32 * (size_t)i, m.ptr(i), 32);
    return out;
}

ORBmatcherAccel::ORBmatcherAccel(float nnratio, int device) : ctx(nullptr), mfNNratio(nnratio) {
    if (rumi_match_create(&ctx, device) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}
ORBmatcherAccel::~ORBmatcherAccel() { rumi_match_destroy(ctx); }

int ORBmatcherAccel::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    return rumi_descriptor_distance(a.ptr(0), b.ptr(0));
}

void ORBmatcherAccel::Top2(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& idx1, std::vector<uint16_t>& d1,
                           std::vector<uint16_t>& d2) {
    const std::vector<uint8_t> q = rows32(Q), t = rows32(T);
    idx1.assign(Q.rows, -1); d1.assign(Q.rows, 256); d2.assign(Q.rows, 256);
    if (Q.rows == 0) return;
    if (rumi_hamming_top2(ctx, q.data(), Q.rows, t.data(), T.rows, idx1.data(), d1.data(), d2.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}

int ORBmatcherAccel::MatchRatio(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& matches12, int th) {
    std::vector<int> idx;
    std::vector<uint16_t> d1, d2;
    Top2(Q, T, idx, d1, d2);
    matches12.assign(Q.rows, -1);
    int n = 0;
    for (int q = 0; q < Q.rows; ++q) {
        const int bestDist1 = d1[q], bestDist2 = d2[q];
        if (idx[q] >= 0 && bestDist1 <= th &&
            static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {   // ORBmatcher.cc:290-291
            matches12[q] = idx[q];
            ++n;
        }
    }
    return n;
}

void ORBmatcherAccel::StereoBest1(const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                                  const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR,
                                  const std::vector<float>& scaleFactors, int nRows, float minD, float maxD,
                                  std::vector<int>& bestIdxR, std::vector<uint16_t>& bestDist) {
    static_assert(sizeof(cv::KeyPoint) == sizeof(rumi_kp), "cv::KeyPoint layout");
    const std::vector<uint8_t> l = rows32(descL), r = rows32(descR);
    bestIdxR.assign(keysL.size(), -1); bestDist.assign(keysL.size(), TH_HIGH);
    if (keysL.empty()) return;
    if (rumi_stereo_best1(ctx, reinterpret_cast<const rumi_kp*>(keysL.data()), l.data(), (int)keysL.size(),
                          reinterpret_cast<const rumi_kp*>(keysR.data()), r.data(), (int)keysR.size(),
                          scaleFactors.data(), (int)scaleFactors.size(), nRows, minD, maxD, bestIdxR.data(),
                          bestDist.data()) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
}

int ORBmatcherAccel::ComputeStereoMatches(rumi_orb* extractorLeft, rumi_orb* extractorRight,
                                          const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                                          const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR, float mbf,
                                          float mb, std::vector<float>& mvuRight, std::vector<float>& mvDepth) {
    const std::vector<uint8_t> l = rows32(descL), r = rows32(descR);
    mvuRight.assign(keysL.size(), -1.0f);
    mvDepth.assign(keysL.size(), -1.0f);
    int n = 0;
    if (keysL.empty()) return 0;
    if (rumi_stereo_match(ctx, extractorLeft, extractorRight, reinterpret_cast<const rumi_kp*>(keysL.data()), l.data(),
                          (int)keysL.size(), reinterpret_cast<const rumi_kp*>(keysR.data()), r.data(),
                          (int)keysR.size(), mbf, mb, mvuRight.data(), mvDepth.data(), &n) != RUMI_OK)
        throw std::runtime_error(std::string("ORBmatcherAccel: ") + rumi_last_error());
    return n;
}

}  // namespace ORB_SLAM3
