// Accelerated core of ORB_SLAM3::ORBmatcher (R/include/cloud_edge_slam_lib/ORBmatcher.h:36-103).
// ORBmatcher's Search*/Fuse methods are geometry + bookkeeping around one inner operation: scan candidates,
// keep best / second-best DescriptorDistance, apply a threshold and a ratio (R/lib_src/ORBmatcher.cc:253-291 etc.).
// This header provides (1) DescriptorDistance with the reference's exact signature (host inline popcount, the API
// itself), and (2) batched top-2 / stereo best-1 on the GPU returning the RAW triple so that every call site keeps
// its own acceptance test (TH_LOW / TH_HIGH, '<' vs '<=', float casts) unchanged on the host.
#ifndef ORBMATCHER_ACCEL_H
#define ORBMATCHER_ACCEL_H

#include <cstdint>
#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>

struct rumi_match;
struct rumi_orb;

namespace ORB_SLAM3 {

class ORBmatcherAccel {
public:
    static const int TH_LOW = 50;       // R/lib_src/ORBmatcher.cc:32
    static const int TH_HIGH = 100;     // R/lib_src/ORBmatcher.cc:31
    static const int HISTO_LENGTH = 30; // R/lib_src/ORBmatcher.cc:33

    explicit ORBmatcherAccel(float nnratio = 0.6f, int device = 0);
    ~ORBmatcherAccel();

    // == ORBmatcher::DescriptorDistance (R/lib_src/ORBmatcher.cc:1830-1844); rows must be 32 contiguous bytes.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);

    // Raw top-2 of every row of Q (nq x 32, CV_8U) against every row of T: earliest index among ties, d2 may equal
    // d1, (-1, 256, 256) when T is empty.
    void Top2(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& idx1, std::vector<uint16_t>& d1,
              std::vector<uint16_t>& d2);

    // SearchByBoW-style acceptance (R/lib_src/ORBmatcher.cc:290-291): best1 <= TH_LOW && best1 < ratio * best2.
    // matches12[q] = train index or -1.  Returns the number of matches.
    int MatchRatio(const cv::Mat& Q, const cv::Mat& T, std::vector<int>& matches12, int th = TH_LOW);

    // Descriptor association of MANY matched key-frame pairs in one launch (submap merge: the pairs
    // R/lib_src/CloudMerging.cc:503-551 walks).  Per pair p and feature q of descA[p]: matches12[p][q] = feature of
    // descB[p] accepted by the rule above, else -1.  Same result as MatchRatio pair by pair.  Returns all matches.
    int MatchKeyFramePairs(const std::vector<cv::Mat>& descA, const std::vector<cv::Mat>& descB,
                           std::vector<std::vector<int>>& matches12, int th = TH_LOW);

    // Frame::ComputeStereoMatches row-band best-1 (R/lib_src/Frame.cc:844-905).
    void StereoBest1(const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                     const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR,
                     const std::vector<float>& scaleFactors, int nRows, float minD, float maxD,
                     std::vector<int>& bestIdxR, std::vector<uint16_t>& bestDist);

    // Frame::ComputeStereoMatches, complete (R/lib_src/Frame.cc:828-985): fills mvuRight / mvDepth (-1 = no match)
    // from the keypoints / descriptors the two extractors produced in their last call; the image pyramids stay on
    // the device.  Returns the number of stereo matches kept after the median outlier cut.
    int ComputeStereoMatches(struct rumi_orb* extractorLeft, struct rumi_orb* extractorRight,
                             const std::vector<cv::KeyPoint>& keysL, const cv::Mat& descL,
                             const std::vector<cv::KeyPoint>& keysR, const cv::Mat& descR, float mbf, float mb,
                             std::vector<float>& mvuRight, std::vector<float>& mvDepth);

    // ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (R/lib_src/ORBmatcher.cc:198-370), the
    // F.Nleft == -1 branch, on flattened inputs: featVec* = the two DBoW2::FeatureVector maps as (node id, feature
    // indices) pairs in ascending node order, kfValid[i] = the keyframe feature has a map point that is not bad,
    // angle* = mvKeysUn[i].angle / mvKeys[i].angle.  matchF[j] = keyframe feature whose map point the reference
    // would store in vpMapPointMatches[j], or -1.  All DescriptorDistance calls run on the device; the acceptance is
    // replayed in the reference's order (it depends on earlier acceptances, :249).
    int SearchByBoW(const cv::Mat& descKF, const std::vector<float>& angleKF, const std::vector<uint8_t>& kfValid,
                    const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecKF, const cv::Mat& descF,
                    const std::vector<float>& angleF,
                    const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVecF, bool checkOrientation,
                    std::vector<int>& matchF);

    // ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (R/lib_src/ORBmatcher.cc:682-804), NLeft == -1 for both
    // keyframes.  valid1 / valid2: the feature has a map point that is not bad.  match12[i] = feature of keyframe 2
    // whose map point the reference stores in vpMatches12[i], or -1.
    int SearchByBoWKF(const cv::Mat& desc1, const std::vector<float>& angle1, const std::vector<uint8_t>& valid1,
                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec1, const cv::Mat& desc2,
                      const std::vector<float>& angle2, const std::vector<uint8_t>& valid2,
                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec2, bool checkOrientation,
                      std::vector<int>& match12);

private:
    // distance blocks of the vocabulary nodes common to two feature vectors (rumi_bow_node_distances)
    long long NodeBlocks(const cv::Mat& descA, const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvA,
                         const cv::Mat& descB, const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvB,
                         std::vector<int32_t>& aIdx, std::vector<int32_t>& bIdx, std::vector<int32_t>& segs,
                         std::vector<uint16_t>& dist);
    rumi_match* ctx;
    float mfNNratio;
};

}  // namespace ORB_SLAM3
#endif
