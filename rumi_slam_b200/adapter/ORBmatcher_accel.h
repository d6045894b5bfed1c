// Accelerated core of ORB_SLAM3::ORBmatcher (R/include/cloud_edge_slam_lib/ORBmatcher.h:36-103).
// ORBmatcher's Search*/Fuse methods are geometry + bookkeeping around one inner operation: scan candidates,
// keep best / second-best DescriptorDistance, apply a threshold and a ratio (R/lib_src/ORBmatcher.cc:253-291 etc.).
// This header provides (1) DescriptorDistance with the reference's exact signature (host inline popcount, the API
// itself), and (2) batched top-2 / stereo best-1 on the GPU returning the RAW triple so that every call site keeps
// its own acceptance test (TH_LOW / TH_HIGH, '<' vs '<=', float casts) unchanged on the host.
#ifndef ORBMATCHER_ACCEL_H
#define ORBMATCHER_ACCEL_H

#include <cstdint>
#include <functional>
#include <utility>
#include <vector>
#include <opencv2/core/core.hpp>

struct rumi_match;
struct rumi_orb;

namespace ORB_SLAM3 {

// The 64 x 48 key-point grid a reference Frame / KeyFrame carries (FRAME_GRID_COLS / ROWS,
// R/include/cloud_edge_slam_lib/Frame.h:42-43): AssignFeaturesToGrid + PosInGrid (R/lib_src/Frame.cc:441-466, 752-767)
// and GetFeaturesInArea (:695-750; KeyFrame.cc:887-925 walks the same way).  A caller that already holds a reference
// Frame uses Frame::GetFeaturesInArea itself; this class is for flattened inputs.
class FrameGridAccel {
public:
    FrameGridAccel(const std::vector<cv::KeyPoint>& keysUn, float minX, float minY, float maxX, float maxY);
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                          const int maxLevel = -1) const;
    // Frame::mnMinX / mnMinY / mnMaxX / mnMaxY (the undistorted image bounds the grid was built for)
    float mnMinX, mnMinY, mnMaxX, mnMaxY;
private:
    static const int kCols = 64, kRows = 48;
    const std::vector<cv::KeyPoint>& keys;
    float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[kCols][kRows];
};

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
                    std::vector<int>& matchF, int Nleft = -1);   // Nleft = F.Nleft: the stereo-fisheye branches of :258-340

    // Frame::ComputeStereoFishEyeMatches (R/lib_src/Frame.cc:1120-1161), the matching core: brute-force k = 2 between the
    // lapping-area descriptors (rows monoLeft.. / monoRight.., :1122-1126) + Lowe's ratio d0 < d1 * 0.7 (:1146).  pairs =
    // (left feature, right feature) in the reference's order, indices in the full arrays; the caller triangulates each pair
    // with its camera model (KannalaBrandt8::TriangulateMatches) and keeps those with depth > 0.0001 (:1151-1158).
    void StereoFishEyeMatches(const cv::Mat& descLeft, int monoLeft, const cv::Mat& descRight, int monoRight,
                              std::vector<std::pair<int, int> >& pairs);

    // ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (R/lib_src/ORBmatcher.cc:682-804), NLeft == -1 for both
    // keyframes.  valid1 / valid2: the feature has a map point that is not bad.  match12[i] = feature of keyframe 2
    // whose map point the reference stores in vpMatches12[i], or -1.
    int SearchByBoWKF(const cv::Mat& desc1, const std::vector<float>& angle1, const std::vector<uint8_t>& valid1,
                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec1, const cv::Mat& desc2,
                      const std::vector<float>& angle2, const std::vector<uint8_t>& valid2,
                      const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec2, bool checkOrientation,
                      std::vector<int>& match12);

    // ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (R/lib_src/ORBmatcher.cc:806-1013), the
    // matcher of LocalMapping::CreateNewMapPoints, key frames without a second camera.  hasMapPoint*[i]: GetMapPoint(i) != NULL;
    // stereo*[i]: mvuRight[i] >= 0; keys2 = pKF2->mvKeysUn; scaleFactors2 = pKF2->mvScaleFactors; epipole = the projection of
    // camera centre 1 into image 2 (:815-819); epipolarOk(idx1, idx2) = pCamera1->epipolarConstrain(pCamera2, kp1, kp2, R12, t12,
    // sigma1, sigma2) (:957) -- the caller's geometry, asked only for pairs that survive the distance tests, in the
    // reference's order.  vMatchedPairs as in the reference.  Distances of all common-node blocks run in one launch.
    int SearchForTriangulation(const cv::Mat& desc1, const std::vector<float>& angle1, const std::vector<uint8_t>& hasMapPoint1,
                               const std::vector<uint8_t>& stereo1,
                               const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec1, const cv::Mat& desc2,
                               const std::vector<float>& angle2, const std::vector<uint8_t>& hasMapPoint2,
                               const std::vector<uint8_t>& stereo2, const std::vector<cv::KeyPoint>& keys2,
                               const std::vector<std::pair<unsigned, std::vector<unsigned> > >& featVec2,
                               const std::vector<float>& scaleFactors2, cv::Point2f epipole,
                               const std::function<bool(size_t, size_t)>& epipolarOk, bool bOnlyStereo, bool bCoarse,
                               bool checkOrientation, std::vector<std::pair<size_t, size_t> >& vMatchedPairs);

    // ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (R/lib_src/ORBmatcher.cc:581-680)
    // on flattened frames: keys* = mvKeysUn, desc* = mDescriptors, grid2 = F2's key-point grid.  Every DescriptorDistance
    // of every (level-0 key point of F1, window candidate of F2) pair runs in ONE launch (rumi_hamming_candidates); the
    // acceptance -- vMatchedDistance skip, re-assignment, rotation histogram -- is replayed in the reference's order.
    int SearchForInitialization(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1,
                                const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2, const FrameGridAccel& grid2,
                                std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize,
                                bool checkOrientation);

    // ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, ...) (R/lib_src/ORBmatcher.cc:39-118), mono frame.  Per map
    // point in view: mTrackProjX/Y, mnTrackScaleLevel, mTrackViewCos, GetDescriptor(), Observations() > 0.
    // frameMatch[j] = map point the reference would store in F.mvpMapPoints[j], or -1.
    int SearchByProjection(const std::vector<cv::KeyPoint>& keysF, const cv::Mat& descF, const FrameGridAccel& gridF,
                           const std::vector<float>& scaleFactors, const std::vector<cv::Point2f>& proj,
                           const std::vector<int>& level, const std::vector<float>& viewCos, const cv::Mat& descMP,
                           const std::vector<uint8_t>& hasObservations, float th, std::vector<int>& frameMatch);

    // The rest of that function (R/lib_src/ORBmatcher.cc:39-189) as Tracking::SearchLocalPoints meets it.  All members optional.
    struct LocalPointsExtras {
        const std::vector<uint8_t>* occupied = nullptr;     // F.mvpMapPoints[j] holds a point with observations on entry (:80-82)
        const std::vector<float>* uRight = nullptr;         // F.mvuRight: right-image gate of rectified stereo / RGB-D (:84-88) ...
        const std::vector<cv::Point2f>* projR = nullptr;    // ... against mTrackProjXR (.x); .y = mTrackProjYR (fisheye only)
        // stereo-fisheye rig (F.Nleft = keysF.size(); descF = left rows, then right rows): the second half of the loop (:125-185)
        const std::vector<cv::KeyPoint>* keysRight = nullptr;          // F.mvKeysRight
        const FrameGridAccel* gridRight = nullptr;                     // grid over keysRight (F.mGridRight)
        const std::vector<uint8_t>* inView = nullptr;                  // mbTrackInView (default: all)
        const std::vector<uint8_t>* inViewR = nullptr;                 // mbTrackInViewR
        const std::vector<int>* levelR = nullptr;                      // mnTrackScaleLevelR
        const std::vector<float>* viewCosR = nullptr;                  // mTrackViewCosR
        const std::vector<int>* leftToRight = nullptr;                 // F.mvLeftToRightMatch
        const std::vector<int>* rightToLeft = nullptr;                 // F.mvRightToLeftMatch
    };
    // frameMatch gets keysF.size() (+ keysRight->size()) entries: the map point this call stored in F.mvpMapPoints[j], or -1.
    int SearchByProjection(const std::vector<cv::KeyPoint>& keysF, const cv::Mat& descF, const FrameGridAccel& gridF,
                           const std::vector<float>& scaleFactors, const std::vector<cv::Point2f>& proj,
                           const std::vector<int>& level, const std::vector<float>& viewCos, const cv::Mat& descMP,
                           const std::vector<uint8_t>& hasObservations, float th, const LocalPointsExtras& ex,
                           std::vector<int>& frameMatch);

    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (R/lib_src/ORBmatcher.cc:1498-1684),
    // the matcher of Tracking::TrackWithMotionModel, frames without a second fisheye camera (Nleft == -1).  Last frame, per
    // feature i: valid[i] = has a map point and is not an outlier (:1518-1520); uv[i], invzc[i] = projection of that point
    // into the current frame and 1 / depth (the caller's pose and camera model, :1522-1534); octaveLast[i] / angleLast[i] of
    // its key point; descMP.row(i) = GetDescriptor(); mpHasObservations[i] = Observations() > 0.  Current frame: mvKeysUn,
    // mDescriptors, its grid, mvScaleFactors, mvuRight (empty: mono), occupied[j] = mvpMapPoints[j] already holds a point
    // with observations (empty: none), mbf.  bForward / bBackward as computed at :1513-1514.
    // curMatch[j] = last-frame feature whose map point the reference would store in CurrentFrame.mvpMapPoints[j], or -1.
    int SearchByProjectionLastFrame(const std::vector<cv::KeyPoint>& keysC, const cv::Mat& descC, const FrameGridAccel& gridC,
                                    const std::vector<float>& scaleFactors, const std::vector<float>& uRight,
                                    const std::vector<uint8_t>& occupied, float mbf, const std::vector<uint8_t>& valid,
                                    const std::vector<cv::Point2f>& uv, const std::vector<float>& invzc,
                                    const std::vector<int>& octaveLast, const std::vector<float>& angleLast,
                                    const cv::Mat& descMP, const std::vector<uint8_t>& mpHasObservations, float th,
                                    bool bForward, bool bBackward, bool checkOrientation, std::vector<int>& curMatch);

    // The same with a stereo-fisheye CURRENT frame (CurrentFrame.Nleft = keysC.size(), R/lib_src/ORBmatcher.cc:1602-1656 in
    // addition): keysRight = mvKeysRight with their own grid, descC = left rows followed by right rows, uvR[i] = projection of
    // point i into the right camera (GetRelativePoseTrl() * x3Dc through the caller's camera model).  A point whose left window
    // is not empty is also searched, best-1, in the right camera; both vote in one rotation histogram.  occupied (may be
    // empty) and curMatch cover keysC.size() + keysRight.size() features.
    int SearchByProjectionLastFrameFisheye(const std::vector<cv::KeyPoint>& keysC, const std::vector<cv::KeyPoint>& keysRight,
                                           const cv::Mat& descC, const FrameGridAccel& gridC, const FrameGridAccel& gridRight,
                                           const std::vector<float>& scaleFactors, const std::vector<uint8_t>& occupied,
                                           const std::vector<uint8_t>& valid, const std::vector<cv::Point2f>& uv,
                                           const std::vector<cv::Point2f>& uvR, const std::vector<float>& invzc,
                                           const std::vector<int>& octaveLast, const std::vector<float>& angleLast,
                                           const cv::Mat& descMP, const std::vector<uint8_t>& mpHasObservations, float th,
                                           bool bForward, bool bBackward, bool checkOrientation, std::vector<int>& curMatch);

    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) (R/lib_src/ORBmatcher.cc:1685-1794),
    // the matcher of Tracking::Relocalization.  Key frame, per feature i: valid[i] = map point present, not bad, not in
    // sAlreadyFound; uv[i] = its projection; dist3D[i] = |x3Dw - Ow| against [minDistance[i], maxDistance[i]] (:1717-1725);
    // predictedLevel[i] = PredictScale (:1727); angleKF[i] = pKF->mvKeysUn[i].angle.  A current-frame feature that holds ANY
    // map point (occupied[j], or assigned earlier in this call) is skipped (:1742-1743).
    int SearchByProjectionKeyFrame(const std::vector<cv::KeyPoint>& keysC, const cv::Mat& descC, const FrameGridAccel& gridC,
                                   const std::vector<float>& scaleFactors, const std::vector<uint8_t>& occupied,
                                   const std::vector<uint8_t>& valid, const std::vector<cv::Point2f>& uv,
                                   const std::vector<float>& dist3D, const std::vector<float>& minDistance,
                                   const std::vector<float>& maxDistance, const std::vector<int>& predictedLevel,
                                   const std::vector<float>& angleKF, const cv::Mat& descMP, float th, int ORBdist,
                                   bool checkOrientation, std::vector<int>& curMatch);

    // ORBmatcher::Fuse(KeyFrame* pKF, vpMapPoints, th, bRight = false) (R/lib_src/ORBmatcher.cc:1015-1181), the matcher of
    // LocalMapping::SearchInNeighbors, up to the fuse decision (:1147).  Per map point i: valid[i] = present, not bad, not
    // already in pKF, depth >= 0, viewing-angle test passed (:1045-1100, the caller's pose / camera / normal); uv[i] =
    // projection into pKF, ur[i] = uv.x - bf * invz (:1074); dist3D[i] against [minDistance, maxDistance];
    // predictedLevel[i] = PredictScale.  Key frame: mvKeysUn, mDescriptors, its grid, mvScaleFactors, mvInvLevelSigma2,
    // mvuRight (< 0: mono feature).  bestIdx[i] = key-frame feature to fuse with (bestDist <= TH_LOW) or -1; the caller
    // runs its own Replace / AddObservation / AddMapPoint loop (:1148-1160) on it.  Returns nFused.
    int FuseSearch(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK, const FrameGridAccel& gridK,
                   const std::vector<float>& scaleFactors, const std::vector<float>& invLevelSigma2,
                   const std::vector<float>& uRight, const std::vector<uint8_t>& valid, const std::vector<cv::Point2f>& uv,
                   const std::vector<float>& ur, const std::vector<float>& dist3D, const std::vector<float>& minDistance,
                   const std::vector<float>& maxDistance, const std::vector<int>& predictedLevel, const cv::Mat& descMP,
                   float th, std::vector<int>& bestIdx, std::vector<int>& bestDist, int thDist = TH_LOW);

    // Matching core of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (R/lib_src/ORBmatcher.cc:1182-1292, LoopClosing):
    // FuseSearch without the reprojection gates.  The caller fills vpReplacePoint / AddObservation from bestIdx (:1268-1280).
    int FuseSearchSim3(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK, const FrameGridAccel& gridK,
                       const std::vector<float>& scaleFactors, const std::vector<uint8_t>& valid,
                       const std::vector<cv::Point2f>& uv, const std::vector<float>& dist3D,
                       const std::vector<float>& minDistance, const std::vector<float>& maxDistance,
                       const std::vector<int>& predictedLevel, const cv::Mat& descMP, float th, std::vector<int>& bestIdx,
                       std::vector<int>& bestDist, int thDist = TH_LOW);

    // ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (R/lib_src/ORBmatcher.cc:1293-1497), LoopClosing.  valid1[i1] =
    // feature i1 of pKF1 has a map point that is not bad and not matched yet (vbAlreadyMatched1), depth >= 0 in camera 2;
    // uv12 / dist12 / level12 = its projection into image 2, |p3Dc2| against [min1, max1], PredictScale (:1336-1365); the *2 /
    // *21 arguments likewise for pKF2's points in image 1 (valid2 excludes vbAlreadyMatched2).  Each direction is the gate-free
    // fuse search with TH_HIGH; match12[i1] = feature of pKF2 when both directions agree (:1483-1494), else -1.
    int SearchBySim3(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1, const FrameGridAccel& grid1,
                     const std::vector<cv::KeyPoint>& keys2, const cv::Mat& desc2, const FrameGridAccel& grid2,
                     const std::vector<float>& scaleFactors1, const std::vector<float>& scaleFactors2,
                     const std::vector<uint8_t>& valid1, const std::vector<cv::Point2f>& uv12, const std::vector<float>& dist12,
                     const std::vector<float>& min1, const std::vector<float>& max1, const std::vector<int>& level12,
                     const std::vector<uint8_t>& valid2, const std::vector<cv::Point2f>& uv21, const std::vector<float>& dist21,
                     const std::vector<float>& min2, const std::vector<float>& max2, const std::vector<int>& level21, float th,
                     std::vector<int>& match12);

    // ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (R/lib_src/ORBmatcher.cc:372-471) and its
    // overload with vpPointsKFs / vpMatchedKF (:473-580), the matchers of LoopClosing (same search).  valid[i] = not bad, not in
    // vpMatched on entry, depth >= 0, viewing-angle test passed; occupied[j] = vpMatched[j] != NULL on entry.
    // kfMatch[j] = candidate the reference stores in vpMatched[j] (and vpPointsKFs[kfMatch[j]] in vpMatchedKF[j]), or -1.
    int SearchByProjectionSim3(const std::vector<cv::KeyPoint>& keysK, const cv::Mat& descK, const FrameGridAccel& gridK,
                               const std::vector<float>& scaleFactors, const std::vector<uint8_t>& occupied,
                               const std::vector<uint8_t>& valid, const std::vector<cv::Point2f>& uv,
                               const std::vector<float>& dist3D, const std::vector<float>& minDistance,
                               const std::vector<float>& maxDistance, const std::vector<int>& predictedLevel,
                               const cv::Mat& descMP, int th, float ratioHamming, std::vector<int>& kfMatch);

    // Descriptor-based key-point association of the matched key-frame pairs of a submap merge (the pairs
    // R/lib_src/CloudMerging.cc:503-551 associates by pixel distance): real descriptors for the cloud
    // key frames (ORBextractor::CloudFrameComputeDescriptors, one batched call per side -- they carry zero descriptors in
    // the reference, R/src/cloud_edge_main.cpp:937), the top-2 of all pairs in one launch, SearchByBoW's acceptance.
    // Only key points with a map point (valid*) take part, like the `vpMap1MapPoints[i] && vpMap2MapPoints[j]` test.
    // match12[p][i] = key point of key frame 2 or -1; returns the total number of associations.
    int AssociateSubmap(struct rumi_orb* extractor, const std::vector<cv::Mat>& images1,
                        const std::vector<std::vector<cv::KeyPoint>>& keys1, const std::vector<std::vector<uint8_t>>& valid1,
                        const std::vector<cv::Mat>& images2, const std::vector<std::vector<cv::KeyPoint>>& keys2,
                        const std::vector<std::vector<uint8_t>>& valid2, std::vector<std::vector<int>>& match12,
                        int th = TH_LOW);

private:
    // distance blocks of the vocabulary nodes common to two feature vectors (rumi_bow_node_distances)
    // distances of every (query row, candidate) pair of CSR candidate lists (rumi_hamming_candidates)
    void CandidateDistances(const std::vector<uint8_t>& Q, int nq, const cv::Mat& T, const std::vector<int32_t>& off,
                            const std::vector<int32_t>& idx, std::vector<uint16_t>& dist);
    long long NodeBlocks(const cv::Mat& descA, const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvA,
                         const cv::Mat& descB, const std::vector<std::pair<unsigned, std::vector<unsigned> > >& fvB,
                         std::vector<int32_t>& aIdx, std::vector<int32_t>& bIdx, std::vector<int32_t>& segs,
                         std::vector<uint16_t>& dist);
    rumi_match* ctx;
    float mfNNratio;
};

}  // namespace ORB_SLAM3
#endif
