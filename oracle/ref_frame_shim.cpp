// ORACLE -- TEST INFRASTRUCTURE ONLY.
// oracle/_ref/librefframe.so: reference functions that live in translation units which cannot be compiled here
// (Frame.cc, MapPoint.cc, ORBmatcher.cc include Eigen / Sophus / g2o) are cut out of the reference sources WHERE THEY LIE
// by oracle/extract_ref.py at build time (generated file _ref/gen_frame_functions.inc, deleted after the build, never
// committed) and compiled UNMODIFIED over the class stand-ins below.  The stand-ins only declare the data members and
// trivial accessors those function bodies touch; all arithmetic and control flow that runs is the reference's own text:
//   Frame::ComputeStereoMatches            R/lib_src/Frame.cc:828-985
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea     R/lib_src/Frame.cc:441-466, 695-767
//   MapPoint::ComputeDistinctiveDescriptors   R/lib_src/MapPoint.cc:353-426
//   ORBmatcher::DescriptorDistance / ComputeThreeMaxima / RadiusByViewingCos   R/lib_src/ORBmatcher.cc:1830-1844, 1795-1826, 191-196
//   ORBmatcher::SearchByBoW (KeyFrame, Frame) / (KeyFrame, KeyFrame)           R/lib_src/ORBmatcher.cc:198-370, 682-804
//   ORBmatcher::SearchByProjection (Frame, MapPoints) / SearchForInitialization   R/lib_src/ORBmatcher.cc:39-189, 581-680
//   ORBmatcher::SearchByProjection (CurrentFrame, LastFrame) / (CurrentFrame, KeyFrame, sAlreadyFound)   :1498-1684, 1685-1794
//   ORBmatcher::Fuse (KeyFrame, MapPoints) + KeyFrame::GetFeaturesInArea / IsInImage   ORBmatcher.cc:1015-1181, KeyFrame.cc:887-930
//   ORBmatcher::SearchForTriangulation   ORBmatcher.cc:806-1013 (epipolarConstrain = a table look-up stand-in)
//   ORBmatcher::SearchByProjection (KeyFrame, Sim3, ...) x 2 and ORBmatcher::Fuse (KeyFrame, Sim3, ...)   :372-580, :1182-1292
//   ORBmatcher::SearchBySim3   :1293-1497
//     (these take poses and a camera model: Sophus::SE3f / Eigen::Vector3f / GeometricCamera are minimal stand-ins
//      below -- identity rotation, so that "Tcw * x3Dw" is exact -- and the pin covers everything AFTER the projection,
//      which is what the flattened adapters take over; the projection itself stays with the caller's own Sophus / camera)
// cv::norm(NORM_L1) on 8U is an exact integer sum.
#include <opencv2/opencv.hpp>

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include "DBoW2/FeatureVector.h"          // the reference's own (R/Thirdparty/DBoW2)

namespace cv {
enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int) {
    long long s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) s += std::abs((int)a.ptr(y)[x] - (int)b.ptr(y)[x]);
    return (double)s;
}
}  // namespace cv

namespace Eigen {
struct Vector2f { float v[2]; float operator()(int i) const { return v[i]; } };
struct Vector3f {
    float v[3];
    Vector3f() : v{0, 0, 0} {}
    Vector3f(float a, float b, float c) : v{a, b, c} {}
    float operator()(int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    Vector3f operator/(float d) const { return Vector3f(v[0] / d, v[1] / d, v[2] / d); }
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
};
struct Matrix3f { float m[9]; };
}  // namespace Eigen
namespace Sophus {
struct SE3f {                                   // translation only (rotation = identity): R v + t is exact
    Eigen::Vector3f t;
    SE3f() {}
    SE3f(const Eigen::Matrix3f&, const Eigen::Vector3f& tr) : t(tr) {}
    SE3f inverse() const { SE3f r; r.t = Eigen::Vector3f(-t.v[0], -t.v[1], -t.v[2]); return r; }
    Eigen::Vector3f translation() const { return t; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return Eigen::Vector3f(p.v[0] + t.v[0], p.v[1] + t.v[1], p.v[2] + t.v[2]); }
    SE3f operator*(const SE3f& o) const { SE3f r; r.t = Eigen::Vector3f(t.v[0] + o.t.v[0], t.v[1] + o.t.v[1], t.v[2] + o.t.v[2]); return r; }
    Eigen::Matrix3f rotationMatrix() const { Eigen::Matrix3f r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return r; }
};
template <class T>
struct Sim3 {                                   // identity rotation, translation t, scale s
    Eigen::Vector3f t;
    T s = 1;
    Eigen::Matrix3f rotationMatrix() const { Eigen::Matrix3f r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return r; }
    Eigen::Vector3f translation() const { return t; }
    T scale() const { return s; }
    Sim3 inverse() const { Sim3 r; r.t = Eigen::Vector3f(-t.v[0], -t.v[1], -t.v[2]); return r; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return Eigen::Vector3f(p.v[0] + t.v[0], p.v[1] + t.v[1], p.v[2] + t.v[2]); }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

using namespace std;

namespace ORB_SLAM3 {

#include "_ref/gen_frame_defines.inc"      // FRAME_GRID_ROWS / FRAME_GRID_COLS   R/include/cloud_edge_slam_lib/Frame.h:42-43

class GeometricCamera {
public:
    // stand-in camera model: (x, y, z) -> (x, y); the tests put the wanted pixel position into x, y and the depth into z
    Eigen::Vector2f project(const Eigen::Vector3f& p) const { Eigen::Vector2f r; r.v[0] = p(0); r.v[1] = p(1); return r; }
    // stand-in for the camera model's epipolar test: a table indexed by the feature numbers the test stores in class_id
    static const uint8_t* epiTable;
    static int epiCols;
    bool epipolarConstrain(GeometricCamera*, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f&,
                           const Eigen::Vector3f&, const float, const float) {
        return epiTable[(size_t)kp1.class_id * epiCols + kp2.class_id] != 0;
    }
};
const uint8_t* GeometricCamera::epiTable = nullptr;
int GeometricCamera::epiCols = 0;
class Frame;
class KeyFrame;

class MapPoint {
public:
    // tracking variables (R/include/cloud_edge_slam_lib/MapPoint.h)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    void ComputeDistinctiveDescriptors();
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    float GetMaxDistanceInvariance() { return mfMaxDistance; }
    float GetMinDistanceInvariance() { return mfMinDistance; }
    int PredictScale(const float&, Frame*) { return mnTrackScaleLevel; }       // stand-in: the level the test prescribes
    int PredictScale(const float&, KeyFrame*) { return mnTrackScaleLevel; }
    Eigen::Vector3f GetNormal() { return mNormal; }
    bool IsInKeyFrame(KeyFrame*) { return mbInKF; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame*) { return std::make_tuple(mnIndexInOther, -1); }
    int mnIndexInOther = -1;
    void Replace(MapPoint* other);                                             // Fuse: logged, the map is not edited
    void AddObservation(KeyFrame*, int idx);
    Eigen::Vector3f mNormal;
    bool mbInKF = false;
    int id = -1;
    Eigen::Vector3f mWorldPos;
    float mfMaxDistance = 1e30f, mfMinDistance = 0.0f;
    // state
    bool mbBad = false;
    int nObs = 0;
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    cv::Mat mDescriptor;
    std::mutex mMutexFeatures;
};

class KeyFrame {
public:
    bool isBad() { return false; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx);                                  // Fuse: logged
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> r(mvpMapPoints.begin(), mvpMapPoints.end()); r.erase(nullptr); return r; }
    void AddMapPoint(MapPoint*, const size_t&) {}
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<float> mvuRight, mvScaleFactors, mvInvLevelSigma2;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    int NLeft = -1, N = 0;
    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0;
    Sophus::SE3f mTcw;
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetRightPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Sophus::SE3f GetRightPoseInverse() { return mTcw.inverse(); }
    std::vector<float> mvLevelSigma2;
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Eigen::Vector3f GetRightCameraCenter() { return GetCameraCenter(); }
    // grid (R/include/cloud_edge_slam_lib/KeyFrame.h:341-344, 425-428; copied from the Frame in the KeyFrame constructor)
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
};

struct ExtractorView { std::vector<cv::Mat> mvImagePyramid; };

class Frame {
public:
    int N = 0, Nleft = -1;
    std::vector<float> mvuRight, mvDepth;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mbf = 0, mb = 0;
    ExtractorView *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    GeometricCamera* mpCamera2 = nullptr;
    GeometricCamera* mpCamera = nullptr;
    std::vector<bool> mvbOutlier;
    Sophus::SE3f mTcw, mTrl;
    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    static float mfGridElementWidthInv, mfGridElementHeightInv, mnMinX, mnMinY, mnMaxX, mnMaxY;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void ComputeStereoMatches();
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1, const bool bRight = false) const;
};
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0, Frame::mnMinX = 0, Frame::mnMinY = 0;
float Frame::mnMaxX = 0, Frame::mnMaxY = 0;

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                           const int ORBdist);
    int Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints,
                           std::vector<MapPoint*>& vpMatched, int th, float ratioHamming = 1.0);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3<float>& Scw, const std::vector<MapPoint*>& vpPoints,
                           const std::vector<KeyFrame*>& vpPointsKFs, std::vector<MapPoint*>& vpMatched,
                           std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming = 1.0);
    int Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, float th,
             vector<MapPoint*>& vpReplacePoint);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12, const float th);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<pair<size_t, size_t>>& vMatchedPairs,
                               const bool bOnlyStereo, const bool bCoarse = false);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

// Fuse edits the map through these three calls; the stand-ins only record which map point met which key-frame feature
static std::vector<std::pair<int, int>> g_fuseLog;                                 // (map point id, key-frame feature)
MapPoint* KeyFrame::GetMapPoint(const size_t& idx) { g_fuseLog.push_back(std::make_pair(-1, (int)idx)); return mvpMapPoints[idx]; }
void MapPoint::Replace(MapPoint* other) { g_fuseLog.back().first = id >= 0 ? id : other->id; }
void MapPoint::AddObservation(KeyFrame*, int) { g_fuseLog.back().first = id; }

// ================= the reference's own function bodies, cut from /root/reference at build time =================
#include "_ref/gen_frame_functions.inc"
// ================================================================================================================

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {
struct KpRec { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(KpRec) == 28 && sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");

cv::Mat rows32(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8U);
    if (n > 0) std::memcpy(m.data, d, 32 * (size_t)n);
    m.rows = n;
    return m;
}
std::vector<cv::KeyPoint> keys(const void* k, int n) {
    std::vector<cv::KeyPoint> v(n);
    if (n) std::memcpy((void*)v.data(), k, 28 * (size_t)n);
    return v;
}
void fill_featvec(DBoW2::FeatureVector& fv, const int32_t* node, int n) {
    for (int i = 0; i < n; ++i)
        if (node[i] >= 0) fv.addFeature((DBoW2::NodeId)node[i], (unsigned)i);
}
void set_grid(int minX, int minY, int maxX, int maxY) {      // Frame.cc:98-99 (first frame): static grid constants
    Frame::mnMinX = (float)minX; Frame::mnMinY = (float)minY; Frame::mnMaxX = (float)maxX; Frame::mnMaxY = (float)maxY;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(maxX - minX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(maxY - minY);
}
}  // namespace

extern "C" {

// Frame::ComputeStereoMatches on caller-provided pyramids (levels as separate dense images).
int ref_stereo_match(const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                     const void* Lk, const uint8_t* Ld, int nL, const void* Rk, const uint8_t* Rd, int nR,
                     const float* scale, const float* invScale, float mbf, float mb, float* uRight, float* depth) {
    ExtractorView el, er;
    for (int l = 0; l < nlevels; ++l) {
        cv::Mat a(lh[l], lw[l], CV_8U), b(lh[l], lw[l], CV_8U);
        std::memcpy(a.data, pyrL[l], (size_t)lw[l] * lh[l]);
        std::memcpy(b.data, pyrR[l], (size_t)lw[l] * lh[l]);
        el.mvImagePyramid.push_back(a); er.mvImagePyramid.push_back(b);
    }
    Frame F;
    F.N = nL;
    F.mvKeys = keys(Lk, nL); F.mvKeysRight = keys(Rk, nR);
    F.mDescriptors = rows32(Ld, nL); F.mDescriptorsRight = rows32(Rd, nR);
    F.mvScaleFactors.assign(scale, scale + nlevels); F.mvInvScaleFactors.assign(invScale, invScale + nlevels);
    F.mbf = mbf; F.mb = mb;
    F.mpORBextractorLeft = &el; F.mpORBextractorRight = &er;
    F.ComputeStereoMatches();
    int kept = 0;
    for (int i = 0; i < nL; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; kept += F.mvDepth[i] > 0; }
    return kept;
}

// MapPoint::ComputeDistinctiveDescriptors for one map point observed in n key frames (one left observation each; the
// std::map<KeyFrame*, ...> iterates in pointer order = array order here).  Writes the 32 bytes of mDescriptor.
int ref_distinctive(const uint8_t* desc, int n, uint8_t* out32) {
    std::vector<KeyFrame> kfs(n);
    MapPoint mp;
    for (int i = 0; i < n; ++i) {
        kfs[i].mDescriptors = rows32(desc + 32 * (size_t)i, 1);
        mp.mObservations[&kfs[i]] = std::make_tuple(0, -1);
    }
    mp.ComputeDistinctiveDescriptors();
    if (mp.mDescriptor.empty()) return -1;
    std::memcpy(out32, mp.mDescriptor.data, 32);
    return 0;
}

int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORBmatcher::DescriptorDistance(rows32(a, 1), rows32(b, 1));
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...): node ids per feature (-1 = feature not in the FeatureVector);
// kfValid[i] = the keyframe feature has a map point that is not bad.  matchF[j] = keyframe feature index or -1.
int ref_search_by_bow_nleft(const uint8_t* dKF, const float* angKF, const uint8_t* kfValid, const int32_t* nodeKF, int nKF,
                            const uint8_t* dF, const float* angF, const int32_t* nodeF, int nF, float ratio, int checkOri,
                            int nLeft, int32_t* matchF);
int ref_search_by_bow(const uint8_t* dKF, const float* angKF, const uint8_t* kfValid, const int32_t* nodeKF, int nKF,
                      const uint8_t* dF, const float* angF, const int32_t* nodeF, int nF, float ratio, int checkOri,
                      int32_t* matchF) {
    return ref_search_by_bow_nleft(dKF, angKF, kfValid, nodeKF, nKF, dF, angF, nodeF, nF, ratio, checkOri, -1, matchF);
}
// nLeft = F.Nleft (-1: mono / rectified frame; otherwise the stereo-fisheye branches of :258-340 run; the key frame has no
// second camera, so every key point is read from mvKeysUn / F.mvKeys as :296-303 does)
int ref_search_by_bow_nleft(const uint8_t* dKF, const float* angKF, const uint8_t* kfValid, const int32_t* nodeKF, int nKF,
                            const uint8_t* dF, const float* angF, const int32_t* nodeF, int nF, float ratio, int checkOri,
                            int nLeft, int32_t* matchF) {
    KeyFrame kf; Frame F;
    std::vector<MapPoint> mps(nKF);
    kf.mDescriptors = rows32(dKF, nKF);
    kf.mvKeysUn.resize(nKF); kf.mvpMapPoints.assign(nKF, nullptr);
    for (int i = 0; i < nKF; ++i) { kf.mvKeysUn[i].angle = angKF[i]; if (kfValid[i]) kf.mvpMapPoints[i] = &mps[i]; }
    fill_featvec(kf.mFeatVec, nodeKF, nKF);
    F.N = nF; F.Nleft = nLeft; F.mDescriptors = rows32(dF, nF); F.mvKeys.resize(nF);
    for (int j = 0; j < nF; ++j) F.mvKeys[j].angle = angF[j];
    fill_featvec(F.mFeatVec, nodeF, nF);
    ORBmatcher m(ratio, checkOri != 0);
    std::vector<MapPoint*> out;
    const int n = m.SearchByBoW(&kf, F, out);
    for (int j = 0; j < nF; ++j) matchF[j] = out[j] ? (int)(out[j] - mps.data()) : -1;
    return n;
}

int ref_search_by_bow_kf(const uint8_t* d1, const float* ang1, const uint8_t* valid1, const int32_t* node1, int n1,
                         const uint8_t* d2, const float* ang2, const uint8_t* valid2, const int32_t* node2, int n2,
                         float ratio, int checkOri, int32_t* match12) {
    KeyFrame a, b;
    std::vector<MapPoint> mp1(n1), mp2(n2);
    a.mDescriptors = rows32(d1, n1); b.mDescriptors = rows32(d2, n2);
    a.mvKeysUn.resize(n1); b.mvKeysUn.resize(n2);
    a.mvpMapPoints.assign(n1, nullptr); b.mvpMapPoints.assign(n2, nullptr);
    for (int i = 0; i < n1; ++i) { a.mvKeysUn[i].angle = ang1[i]; if (valid1[i]) a.mvpMapPoints[i] = &mp1[i]; }
    for (int i = 0; i < n2; ++i) { b.mvKeysUn[i].angle = ang2[i]; if (valid2[i]) b.mvpMapPoints[i] = &mp2[i]; }
    fill_featvec(a.mFeatVec, node1, n1); fill_featvec(b.mFeatVec, node2, n2);
    ORBmatcher m(ratio, checkOri != 0);
    std::vector<MapPoint*> out;
    const int n = m.SearchByBoW(&a, &b, out);
    for (int i = 0; i < n1; ++i) match12[i] = out[i] ? (int)(out[i] - mp2.data()) : -1;
    return n;
}

// Frame::GetFeaturesInArea on the grid Frame::AssignFeaturesToGrid builds (mono frame: Nleft == -1).
int ref_features_in_area(const void* kps, int n, int minX, int minY, int maxX, int maxY, float x, float y, float r,
                         int minLevel, int maxLevel, int32_t* out, int cap) {
    set_grid(minX, minY, maxX, maxY);
    Frame F;
    F.N = n; F.mvKeysUn = keys(kps, n); F.mvKeys = F.mvKeysUn;
    F.AssignFeaturesToGrid();
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize); prev = [n1][2] in / out.
int ref_search_for_initialization(const void* k1, const uint8_t* d1, int n1, const void* k2, const uint8_t* d2, int n2,
                                  int minX, int minY, int maxX, int maxY, float* prev, int windowSize, float ratio, int checkOri,
                                  int32_t* matches12) {
    set_grid(minX, minY, maxX, maxY);
    Frame F1, F2;
    F1.N = n1; F1.mvKeysUn = keys(k1, n1); F1.mvKeys = F1.mvKeysUn; F1.mDescriptors = rows32(d1, n1);
    F2.N = n2; F2.mvKeysUn = keys(k2, n2); F2.mvKeys = F2.mvKeysUn; F2.mDescriptors = rows32(d2, n2);
    F2.AssignFeaturesToGrid();
    std::vector<cv::Point2f> pm(n1);
    for (int i = 0; i < n1; ++i) pm[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchForInitialization(F1, F2, pm, m12, windowSize);
    for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev[2 * i] = pm[i].x; prev[2 * i + 1] = pm[i].y; }
    return n;
}

// ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, false, 50) on a mono frame (Nleft == -1, mvuRight = -1):
// per map point: projection (x, y), predicted level, view cosine, descriptor, Observations() > 0 flag.
// frameMatch[j] = map point index assigned to frame feature j (or -1).
int ref_search_by_projection(const void* kF, const uint8_t* dF, int nF, const float* scaleFactors, int nlevels,
                             int minX, int minY, int maxX, int maxY, const float* proj /* [nMP][2] */, const int32_t* level,
                             const float* viewCos, const uint8_t* dMP, const uint8_t* hasObs, int nMP, float th, float ratio,
                             int32_t* frameMatch) {
    set_grid(minX, minY, maxX, maxY);
    Frame F;
    F.N = nF; F.mvKeysUn = keys(kF, nF); F.mvKeys = F.mvKeysUn; F.mDescriptors = rows32(dF, nF);
    F.mvuRight.assign(nF, -1.0f);
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mvpMapPoints.assign(nF, nullptr);
    F.AssignFeaturesToGrid();
    std::vector<MapPoint> mps(nMP);
    std::vector<MapPoint*> ptrs(nMP);
    for (int i = 0; i < nMP; ++i) {
        mps[i].mbTrackInView = true;
        mps[i].mTrackProjX = proj[2 * i]; mps[i].mTrackProjY = proj[2 * i + 1];
        mps[i].mnTrackScaleLevel = level[i]; mps[i].mTrackViewCos = viewCos[i];
        mps[i].mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        mps[i].nObs = hasObs[i] ? 1 : 0;
        ptrs[i] = &mps[i];
    }
    ORBmatcher m(ratio, true);
    const int n = m.SearchByProjection(F, ptrs, th, false, 50.0f);
    for (int j = 0; j < nF; ++j) frameMatch[j] = F.mvpMapPoints[j] ? (int)(F.mvpMapPoints[j] - mps.data()) : -1;
    return n;
}

// The same reference function with everything Tracking::SearchLocalPoints can hand it: a frame that already holds matches
// (occupied[j]: mvpMapPoints[j] = a point with observations), mvuRight (rectified stereo / RGB-D), and a stereo-fisheye
// frame (nR > 0: Nleft = nL, mvKeysRight, descriptor rows nL + i, mvLeftToRightMatch / mvRightToLeftMatch).  Map points carry
// both halves of their tracking state.  frameMatch[j] = map point stored by the call (-1: untouched).
int ref_search_by_projection_ex(const void* kL, int nL, const void* kR, int nR, const uint8_t* dF, const float* scaleFactors,
                                int nlevels, int minX, int minY, int maxX, int maxY, const float* uRight, const uint8_t* occupied,
                                const int32_t* l2r, const int32_t* r2l, const uint8_t* inView, const uint8_t* inViewR,
                                const float* proj, const float* projR, const int32_t* level, const int32_t* levelR,
                                const float* viewCos, const float* viewCosR, const uint8_t* dMP, const uint8_t* hasObs, int nMP,
                                float th, float ratio, int32_t* frameMatch) {
    set_grid(minX, minY, maxX, maxY);
    const int N = nL + nR;
    Frame F;
    F.N = N; F.mvKeys = keys(kL, nL); F.mvKeysUn = F.mvKeys; F.mDescriptors = rows32(dF, N);
    if (nR > 0) { F.Nleft = nL; F.mvKeysRight = keys(kR, nR); }
    F.mvuRight.assign(N, -1.0f);
    if (uRight) for (int j = 0; j < nL; ++j) F.mvuRight[j] = uRight[j];
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint taken; taken.nObs = 1;
    F.mvpMapPoints.assign(N, nullptr);
    if (occupied) for (int j = 0; j < N; ++j) if (occupied[j]) F.mvpMapPoints[j] = &taken;
    F.mvLeftToRightMatch.assign(nL, -1); F.mvRightToLeftMatch.assign(nR, -1);
    if (l2r) F.mvLeftToRightMatch.assign(l2r, l2r + nL);
    if (r2l) F.mvRightToLeftMatch.assign(r2l, r2l + nR);
    F.AssignFeaturesToGrid();
    std::vector<MapPoint> mps(nMP);
    std::vector<MapPoint*> ptrs(nMP);
    for (int i = 0; i < nMP; ++i) {
        MapPoint& m = mps[i];
        m.mbTrackInView = inView ? inView[i] != 0 : true;
        m.mbTrackInViewR = inViewR ? inViewR[i] != 0 : false;
        m.mTrackProjX = proj[2 * i]; m.mTrackProjY = proj[2 * i + 1];
        m.mnTrackScaleLevel = level[i]; m.mTrackViewCos = viewCos[i];
        if (projR) { m.mTrackProjXR = projR[2 * i]; m.mTrackProjYR = projR[2 * i + 1]; }
        if (levelR) m.mnTrackScaleLevelR = levelR[i];
        if (viewCosR) m.mTrackViewCosR = viewCosR[i];
        m.mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        m.nObs = hasObs[i] ? 1 : 0;
        ptrs[i] = &m;
    }
    ORBmatcher m(ratio, true);
    const int n = m.SearchByProjection(F, ptrs, th, false, 50.0f);
    for (int j = 0; j < N; ++j)
        frameMatch[j] = (F.mvpMapPoints[j] && F.mvpMapPoints[j] != &taken) ? (int)(F.mvpMapPoints[j] - mps.data()) : -1;
    return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1498-1684), Nleft == -1.  Last frame:
// valid[i] = has a map point and is not an outlier, (uv, depth) = where the stand-in camera puts it in the current frame,
// octave / angle of its key point, descriptor and Observations() > 0 of its map point.  Current frame: key points,
// descriptors, uRight (may be null: all -1), occupied[j] = mvpMapPoints[j] holds a point with observations.
// forward / backward steer bForward / bBackward through the relative pose (tlc(2) vs mb).
int ref_search_by_projection_last(const void* kC, const uint8_t* dC, int nC, const float* scaleFactors, int nlevels, int minX,
                                  int minY, int maxX, int maxY, const float* uRight, const uint8_t* occupied, float mbf,
                                  const uint8_t* valid, const float* uv, const float* depth, const int32_t* octave,
                                  const float* angleLast, const uint8_t* dMP, const uint8_t* mpHasObs, int nL, float th,
                                  int forward, int backward, float ratio, int checkOri, int32_t* curMatch) {
    set_grid(minX, minY, maxX, maxY);
    GeometricCamera cam;
    Frame C, Lf;
    C.N = nC; C.mvKeysUn = keys(kC, nC); C.mvKeys = C.mvKeysUn; C.mDescriptors = rows32(dC, nC);
    C.mvuRight.assign(nC, -1.0f);
    if (uRight) C.mvuRight.assign(uRight, uRight + nC);
    C.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    C.mbf = mbf; C.mb = 1.0f; C.mpCamera = &cam;
    MapPoint taken; taken.nObs = 1;
    C.mvpMapPoints.assign(nC, nullptr);
    for (int j = 0; j < nC; ++j) if (occupied && occupied[j]) C.mvpMapPoints[j] = &taken;
    C.AssignFeaturesToGrid();
    std::vector<MapPoint> mps(nL);
    Lf.N = nL; Lf.mvKeys.resize(nL); Lf.mvKeysUn.resize(nL); Lf.mvpMapPoints.assign(nL, nullptr); Lf.mvbOutlier.assign(nL, false);
    for (int i = 0; i < nL; ++i) {
        Lf.mvKeys[i].octave = octave[i]; Lf.mvKeysUn[i].octave = octave[i];
        Lf.mvKeys[i].angle = angleLast[i]; Lf.mvKeysUn[i].angle = angleLast[i];
        mps[i].mWorldPos = Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], depth[i]);
        mps[i].mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        mps[i].nObs = mpHasObs[i] ? 1 : 0;
        // "not valid" alternates between the two ways the reference skips a feature: no map point / outlier
        if (valid[i]) Lf.mvpMapPoints[i] = &mps[i];
        else if (i & 1) { Lf.mvpMapPoints[i] = &mps[i]; Lf.mvbOutlier[i] = true; }
    }
    // Tcw = identity -> twc = 0, tlc = Tlw.translation(): bForward <=> tlc(2) > mb, bBackward <=> -tlc(2) > mb
    Lf.mTcw.t = Eigen::Vector3f(0, 0, forward ? 2.0f : backward ? -2.0f : 0.0f);
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchByProjection(C, Lf, th, false);
    for (int j = 0; j < nC; ++j) {
        MapPoint* p = C.mvpMapPoints[j];
        curMatch[j] = (p && p != &taken) ? (int)(p - mps.data()) : -1;
    }
    return n;
}

// The same reference function with a stereo-fisheye current frame: Nleft = nC, mvKeysRight = kR, descriptor rows nC + i;
// the right camera sees last-frame point i at uv[i] + (shiftX, shiftY) (GetRelativePoseTrl() = that translation under the
// pass-through stand-in camera).  occupied / curMatch have nC + nR entries.
int ref_search_by_projection_last_fisheye(const void* kC, int nC, const void* kR, int nR, const uint8_t* dC,
                                          const float* scaleFactors, int nlevels, int minX, int minY, int maxX, int maxY,
                                          const uint8_t* occupied, const uint8_t* valid, const float* uv, float shiftX,
                                          float shiftY, const float* depth, const int32_t* octave, const float* angleLast,
                                          const uint8_t* dMP, const uint8_t* mpHasObs, int nL, float th, int forward,
                                          int backward, float ratio, int checkOri, int32_t* curMatch) {
    set_grid(minX, minY, maxX, maxY);
    GeometricCamera cam;
    Frame C, Lf;
    const int N = nC + nR;
    C.N = N; C.Nleft = nC; C.mvKeys = keys(kC, nC); C.mvKeysUn = C.mvKeys; C.mvKeysRight = keys(kR, nR);
    C.mDescriptors = rows32(dC, N);
    C.mvuRight.assign(N, -1.0f);
    C.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    C.mbf = 40.0f; C.mb = 1.0f; C.mpCamera = &cam;
    C.mTrl.t = Eigen::Vector3f(shiftX, shiftY, 0.0f);
    MapPoint taken; taken.nObs = 1;
    C.mvpMapPoints.assign(N, nullptr);
    for (int j = 0; j < N; ++j) if (occupied && occupied[j]) C.mvpMapPoints[j] = &taken;
    C.AssignFeaturesToGrid();
    std::vector<MapPoint> mps(nL);
    Lf.N = nL; Lf.mvKeys.resize(nL); Lf.mvKeysUn.resize(nL); Lf.mvpMapPoints.assign(nL, nullptr); Lf.mvbOutlier.assign(nL, false);
    for (int i = 0; i < nL; ++i) {
        Lf.mvKeys[i].octave = octave[i]; Lf.mvKeysUn[i].octave = octave[i];
        Lf.mvKeys[i].angle = angleLast[i]; Lf.mvKeysUn[i].angle = angleLast[i];
        mps[i].mWorldPos = Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], depth[i]);
        mps[i].mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        mps[i].nObs = mpHasObs[i] ? 1 : 0;
        if (valid[i]) Lf.mvpMapPoints[i] = &mps[i];
        else if (i & 1) { Lf.mvpMapPoints[i] = &mps[i]; Lf.mvbOutlier[i] = true; }
    }
    Lf.mTcw.t = Eigen::Vector3f(0, 0, forward ? 2.0f : backward ? -2.0f : 0.0f);
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchByProjection(C, Lf, th, false);
    for (int j = 0; j < N; ++j) {
        MapPoint* p = C.mvpMapPoints[j];
        curMatch[j] = (p && p != &taken) ? (int)(p - mps.data()) : -1;
    }
    return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1685-1794).  Key frame:
// state[i] = 0 no map point, 1 usable, 2 bad, 3 in sAlreadyFound; (uv, depth), predicted level (stand-in PredictScale),
// distance-invariance window [minDist, maxDist] against |x3Dw - Ow| (Ow = 0 here), key point angle, descriptor.
int ref_search_by_projection_kf(const void* kC, const uint8_t* dC, int nC, const float* scaleFactors, int nlevels, int minX,
                                int minY, int maxX, int maxY, const uint8_t* occupied, const uint8_t* state, const float* uv,
                                const float* depth, const int32_t* level, const float* minDist, const float* maxDist,
                                const float* angleKF, const uint8_t* dMP, int nK, float th, int orbDist, float ratio,
                                int checkOri, int32_t* curMatch, float* dist3D) {
    set_grid(minX, minY, maxX, maxY);
    GeometricCamera cam;
    Frame C;
    C.N = nC; C.mvKeysUn = keys(kC, nC); C.mvKeys = C.mvKeysUn; C.mDescriptors = rows32(dC, nC);
    C.mvuRight.assign(nC, -1.0f);
    C.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    C.mpCamera = &cam;
    MapPoint taken; taken.nObs = 1;
    C.mvpMapPoints.assign(nC, nullptr);
    for (int j = 0; j < nC; ++j) if (occupied && occupied[j]) C.mvpMapPoints[j] = &taken;
    C.AssignFeaturesToGrid();
    KeyFrame kf;
    std::vector<MapPoint> mps(nK);
    std::set<MapPoint*> found;
    kf.mvpMapPoints.assign(nK, nullptr); kf.mvKeysUn.resize(nK);
    for (int i = 0; i < nK; ++i) {
        kf.mvKeysUn[i].angle = angleKF[i];
        mps[i].mWorldPos = Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], depth[i]);
        mps[i].mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        mps[i].mnTrackScaleLevel = level[i];
        mps[i].mfMinDistance = minDist[i]; mps[i].mfMaxDistance = maxDist[i];
        mps[i].nObs = 1;
        if (state[i]) kf.mvpMapPoints[i] = &mps[i];
        if (state[i] == 2) mps[i].mbBad = true;
        if (state[i] == 3) found.insert(&mps[i]);
        if (dist3D) dist3D[i] = mps[i].mWorldPos.norm();            // what the function compares with the window (Ow = 0)
    }
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchByProjection(C, &kf, found, th, orbDist);
    for (int j = 0; j < nC; ++j) {
        MapPoint* p = C.mvpMapPoints[j];
        curMatch[j] = (p && p != &taken) ? (int)(p - mps.data()) : -1;
    }
    return n;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, false) (ORBmatcher.cc:1015-1181) on a key frame without a second camera.  Map
// point i: state[i] = 0 null, 1 usable, 2 bad, 3 already in pKF, 4 viewing angle > 60 deg (normal points away); position
// (uv, depth) through the stand-in camera, invariance window, prescribed PredictScale level, descriptor.  Key frame: key
// points, descriptors, mvuRight, mvInvLevelSigma2, kfHasPoint[j] = 0 none, 1 a point with 5 observations, 2 a bad point.
// bestIdx[i] = the key-frame feature the reference fused map point i with (from the logged GetMapPoint / Replace /
// AddObservation calls), else -1.  Returns nFused.
int ref_fuse(const void* kK, const uint8_t* dK, int nK, const float* scaleFactors, const float* invLevelSigma2, int nlevels, int minX,
             int minY, int maxX, int maxY, const float* uRight, const uint8_t* kfHasPoint, float bf, const uint8_t* state,
             const float* uv, const float* depth, const float* minDist, const float* maxDist, const int32_t* level,
             const uint8_t* dMP, const int32_t* nObs, int nMP, float th, int32_t* bestIdx) {
    set_grid(minX, minY, maxX, maxY);
    GeometricCamera cam;
    Frame F;                                       // the frame the key frame was made from: builds the grid
    F.N = nK; F.mvKeysUn = keys(kK, nK); F.mvKeys = F.mvKeysUn;
    F.AssignFeaturesToGrid();
    KeyFrame kf;
    kf.N = nK; kf.mvKeysUn = F.mvKeysUn; kf.mvKeys = F.mvKeysUn; kf.mDescriptors = rows32(dK, nK);
    kf.mvuRight.assign(uRight, uRight + nK);
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + nlevels);
    kf.mpCamera = &cam; kf.mbf = bf;
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mnMinX = minX; kf.mnMinY = minY; kf.mnMaxX = maxX; kf.mnMaxY = maxY;
    kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) kf.mGrid[i][j] = F.mGrid[i][j];
    MapPoint inKF, badInKF;
    inKF.nObs = 5; inKF.id = -1000; badInKF.mbBad = true; badInKF.id = -1000;
    kf.mvpMapPoints.assign(nK, nullptr);
    for (int j = 0; j < nK; ++j) kf.mvpMapPoints[j] = kfHasPoint[j] == 1 ? &inKF : kfHasPoint[j] == 2 ? &badInKF : nullptr;
    std::vector<MapPoint> mps(nMP);
    std::vector<MapPoint*> ptrs(nMP, nullptr);
    for (int i = 0; i < nMP; ++i) {
        MapPoint& m = mps[i];
        m.id = i;
        m.mWorldPos = Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], depth[i]);
        const float n = m.mWorldPos.norm();
        const float sgn = state[i] == 4 ? -1.0f : 1.0f;                        // normal along / against the viewing ray
        m.mNormal = Eigen::Vector3f(sgn * m.mWorldPos(0) / n, sgn * m.mWorldPos(1) / n, sgn * m.mWorldPos(2) / n);
        m.mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        m.mnTrackScaleLevel = level[i];
        m.mfMinDistance = minDist[i]; m.mfMaxDistance = maxDist[i];
        m.nObs = nObs[i];
        m.mbBad = state[i] == 2; m.mbInKF = state[i] == 3;
        if (state[i]) ptrs[i] = &m;
        bestIdx[i] = -1;
    }
    g_fuseLog.clear();
    ORBmatcher matcher(0.6f, true);
    const int nFused = matcher.Fuse(&kf, ptrs, th, false);
    for (const auto& e : g_fuseLog)
        if (e.first >= 0) bestIdx[e.first] = e.second;
        else if (kfHasPoint[e.second] != 2) return -1000 - e.second;           // only a bad resident point leaves no trace
    return nFused;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = true): the key frame of a stereo-fisheye rig (NLeft = nLeft; mvKeysRight = kR
// with mGridRight; descriptor rows nLeft + i; mpCamera2, GetRightPose / GetRightCameraCenter = the stand-ins of the left
// side).  dAll = all N = nLeft + nR descriptor rows, uRightAll = mvuRight (the reference indexes it with the RIGHT-relative
// index, :1131), kfHasPoint over all N features.  bestIdx[i] = feature (nLeft + right index) point i was fused with.
int ref_fuse_right(int nLeft, const void* kR, int nR, const uint8_t* dAll, const float* scaleFactors, const float* invLevelSigma2,
                   int nlevels, int minX, int minY, int maxX, int maxY, const float* uRightAll, const uint8_t* kfHasPoint, float bf,
                   const uint8_t* state, const float* uv, const float* depth, const float* minDist, const float* maxDist,
                   const int32_t* level, const uint8_t* dMP, const int32_t* nObs, int nMP, float th, int32_t* bestIdx) {
    set_grid(minX, minY, maxX, maxY);
    GeometricCamera cam;
    const int N = nLeft + nR;
    Frame F;                                       // builds the right camera's grid exactly as Frame::AssignFeaturesToGrid does
    F.N = N; F.Nleft = nLeft; F.mvKeys.resize(nLeft); F.mvKeysUn = F.mvKeys; F.mvKeysRight = keys(kR, nR);
    for (auto& k : F.mvKeys) { k.pt.x = -1e6f; k.pt.y = -1e6f; }                       // left features: outside every grid cell
    F.AssignFeaturesToGrid();
    KeyFrame kf;
    kf.N = N; kf.NLeft = nLeft; kf.mvKeys = F.mvKeys; kf.mvKeysUn = F.mvKeys; kf.mvKeysRight = F.mvKeysRight;
    kf.mDescriptors = rows32(dAll, N);
    kf.mvuRight.assign(uRightAll, uRightAll + N);
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvInvLevelSigma2.assign(invLevelSigma2, invLevelSigma2 + nlevels);
    kf.mpCamera = &cam; kf.mpCamera2 = &cam; kf.mbf = bf;
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mnMinX = minX; kf.mnMinY = minY; kf.mnMaxX = maxX; kf.mnMaxY = maxY;
    kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    kf.mGridRight.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) { kf.mGrid[i][j] = F.mGrid[i][j]; kf.mGridRight[i][j] = F.mGridRight[i][j]; }
    MapPoint inKF, badInKF;
    inKF.nObs = 5; inKF.id = -1000; badInKF.mbBad = true; badInKF.id = -1000;
    kf.mvpMapPoints.assign(N, nullptr);
    for (int j = 0; j < N; ++j) kf.mvpMapPoints[j] = kfHasPoint[j] == 1 ? &inKF : kfHasPoint[j] == 2 ? &badInKF : nullptr;
    std::vector<MapPoint> mps(nMP);
    std::vector<MapPoint*> ptrs(nMP, nullptr);
    for (int i = 0; i < nMP; ++i) {
        MapPoint& m = mps[i];
        m.id = i;
        m.mWorldPos = Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], depth[i]);
        const float n = m.mWorldPos.norm();
        const float sgn = state[i] == 4 ? -1.0f : 1.0f;
        m.mNormal = Eigen::Vector3f(sgn * m.mWorldPos(0) / n, sgn * m.mWorldPos(1) / n, sgn * m.mWorldPos(2) / n);
        m.mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
        m.mnTrackScaleLevel = level[i];
        m.mfMinDistance = minDist[i]; m.mfMaxDistance = maxDist[i];
        m.nObs = nObs[i];
        m.mbBad = state[i] == 2; m.mbInKF = state[i] == 3;
        if (state[i]) ptrs[i] = &m;
        bestIdx[i] = -1;
    }
    g_fuseLog.clear();
    ORBmatcher matcher(0.6f, true);
    const int nFused = matcher.Fuse(&kf, ptrs, th, true);
    for (const auto& e : g_fuseLog)
        if (e.first >= 0) bestIdx[e.first] = e.second;
        else if (kfHasPoint[e.second] != 2) return -1000 - e.second;
    return nFused;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (ORBmatcher.cc:806-1013), key frames
// without a second camera.  node*[i] = vocabulary node of feature i (-1: not in the feature vector); hasMP*[i]; uRight*[i];
// the epipole (epx, epy) is produced through the stand-in poses (T1w = identity, T2w = translation (epx, epy, 1), camera
// (x, y, z) -> (x, y)); epiOk[i1 * n2 + i2] = result of epipolarConstrain for the pair.  match12[i1] = i2 or -1.
int ref_search_for_triangulation(const void* k1, const uint8_t* d1, const uint8_t* hasMP1, const float* uRight1, const int32_t* node1,
                                 int n1, const void* k2, const uint8_t* d2, const uint8_t* hasMP2, const float* uRight2,
                                 const int32_t* node2, int n2, const float* scaleFactors2, int nlevels, float epx, float epy,
                                 int onlyStereo, int coarse, const uint8_t* epiOk, float ratio, int checkOri, int32_t* match12) {
    GeometricCamera cam;
    GeometricCamera::epiTable = epiOk; GeometricCamera::epiCols = n2;
    MapPoint some;
    KeyFrame kf1, kf2;
    auto fill = [&](KeyFrame& kf, const void* k, const uint8_t* d, const uint8_t* hasMP, const float* uRight, const int32_t* node, int n) {
        kf.N = n; kf.mvKeysUn = keys(k, n);
        for (int i = 0; i < n; ++i) kf.mvKeysUn[i].class_id = i;
        kf.mvKeys = kf.mvKeysUn; kf.mDescriptors = rows32(d, n);
        kf.mvuRight.assign(uRight, uRight + n);
        kf.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; ++i) if (hasMP[i]) kf.mvpMapPoints[i] = &some;
        fill_featvec(kf.mFeatVec, node, n);
        kf.mpCamera = &cam;
        kf.mvScaleFactors.assign(scaleFactors2, scaleFactors2 + nlevels);
        kf.mvLevelSigma2.assign(nlevels, 1.0f);
    };
    fill(kf1, k1, d1, hasMP1, uRight1, node1, n1);
    fill(kf2, k2, d2, hasMP2, uRight2, node2, n2);
    kf2.mTcw.t = Eigen::Vector3f(epx, epy, 1.0f);              // C2 = T2w * Cw = (0,0,0) + t2 -> ep = (epx, epy)
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchForTriangulation(&kf1, &kf2, pairs, onlyStereo != 0, coarse != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (const auto& p : pairs) match12[p.first] = (int)p.second;
    return n;
}

// The same reference function on two key frames of a stereo-fisheye rig (mpCamera2 set, NLeft = nLeft*, mvKeys = the first
// nLeft* key points, mvKeysRight = the rest; class_id = flattened feature number, so the epipolar table is indexed as before).
int ref_search_for_triangulation_rig(const void* k1, const uint8_t* d1, const uint8_t* hasMP1, const int32_t* node1, int n1,
                                     int nLeft1, const void* k2, const uint8_t* d2, const uint8_t* hasMP2, const int32_t* node2,
                                     int n2, int nLeft2, const float* scaleFactors2, int nlevels, float epx, float epy,
                                     int onlyStereo, int coarse, const uint8_t* epiOk, float ratio, int checkOri, int32_t* match12) {
    GeometricCamera cam, cam2;
    GeometricCamera::epiTable = epiOk; GeometricCamera::epiCols = n2;
    MapPoint some;
    KeyFrame kf1, kf2;
    auto fill = [&](KeyFrame& kf, const void* k, const uint8_t* d, const uint8_t* hasMP, const int32_t* node, int n, int nLeft) {
        std::vector<cv::KeyPoint> all = keys(k, n);
        for (int i = 0; i < n; ++i) all[i].class_id = i;
        kf.N = n; kf.NLeft = nLeft;
        kf.mvKeys.assign(all.begin(), all.begin() + nLeft); kf.mvKeysUn = kf.mvKeys;
        kf.mvKeysRight.assign(all.begin() + nLeft, all.end());
        kf.mDescriptors = rows32(d, n);
        kf.mvuRight.assign(n, 5.0f);                   // would make every feature "stereo" if the second camera did not rule it out
        kf.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; ++i) if (hasMP[i]) kf.mvpMapPoints[i] = &some;
        fill_featvec(kf.mFeatVec, node, n);
        kf.mpCamera = &cam; kf.mpCamera2 = &cam2;
        kf.mvScaleFactors.assign(scaleFactors2, scaleFactors2 + nlevels);
        kf.mvLevelSigma2.assign(nlevels, 1.0f);
    };
    fill(kf1, k1, d1, hasMP1, node1, n1, nLeft1);
    fill(kf2, k2, d2, hasMP2, node2, n2, nLeft2);
    kf2.mTcw.t = Eigen::Vector3f(epx, epy, 1.0f);
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher m(ratio, checkOri != 0);
    const int n = m.SearchForTriangulation(&kf1, &kf2, pairs, onlyStereo != 0, coarse != 0);
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (const auto& p : pairs) match12[p.first] = (int)p.second;
    return n;
}

}  // extern "C"

namespace {
// key frame + candidate map points of the three Sim3 functions (stand-in camera (x, y, z) -> (x, y) for the overloads that
// call mpCamera->project; the overload that projects with fx, fy, cx, cy gets fx = fy = z-compensating values instead:
// u = fx * x / z + cx with fx = 1, cx = 0 needs world point (u * z, v * z, z))
struct Sim3Case {
    GeometricCamera cam;
    Frame F;
    KeyFrame kf;
    MapPoint resident;
    std::vector<MapPoint> mps;
    std::vector<MapPoint*> ptrs;
    Sim3Case(const void* kK, const uint8_t* dK, int nK, const float* scaleFactors, int nlevels, int minX, int minY, int maxX, int maxY,
             const uint8_t* occupied, const uint8_t* state, const float* uv, const float* depth, const float* minDist,
             const float* maxDist, const int32_t* level, const uint8_t* dMP, int nMP, bool pinholeInline)
        : mps(nMP), ptrs(nMP) {
        set_grid(minX, minY, maxX, maxY);
        F.N = nK; F.mvKeysUn = keys(kK, nK); F.mvKeys = F.mvKeysUn;
        F.AssignFeaturesToGrid();
        kf.N = nK; kf.mvKeysUn = F.mvKeysUn; kf.mvKeys = F.mvKeysUn; kf.mDescriptors = rows32(dK, nK);
        kf.mvuRight.assign(nK, -1.0f);
        kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
        kf.mpCamera = &cam;
        kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
        kf.mnMinX = minX; kf.mnMinY = minY; kf.mnMaxX = maxX; kf.mnMaxY = maxY;
        kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
        for (int i = 0; i < FRAME_GRID_COLS; ++i)
            for (int j = 0; j < FRAME_GRID_ROWS; ++j) kf.mGrid[i][j] = F.mGrid[i][j];
        resident.id = -1000; resident.nObs = 3;
        kf.mvpMapPoints.assign(nK, nullptr);
        for (int j = 0; j < nK; ++j) if (occupied[j]) kf.mvpMapPoints[j] = &resident;
        for (int i = 0; i < nMP; ++i) {
            MapPoint& m = mps[i];
            m.id = i;
            const float z = depth[i];
            m.mWorldPos = pinholeInline ? Eigen::Vector3f(uv[2 * i] * z, uv[2 * i + 1] * z, z) : Eigen::Vector3f(uv[2 * i], uv[2 * i + 1], z);
            const float n = m.mWorldPos.norm();
            const float sgn = state[i] == 4 ? -1.0f : 1.0f;
            m.mNormal = Eigen::Vector3f(sgn * m.mWorldPos(0) / n, sgn * m.mWorldPos(1) / n, sgn * m.mWorldPos(2) / n);
            m.mDescriptor = rows32(dMP + 32 * (size_t)i, 1);
            m.mnTrackScaleLevel = level[i];
            m.mfMinDistance = minDist[i]; m.mfMaxDistance = maxDist[i];
            m.mbBad = state[i] == 2;
            ptrs[i] = &m;
        }
    }
};
}  // namespace

extern "C" {

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (variant 0, ORBmatcher.cc:372-471) or the
// overload with vpPointsKFs / vpMatchedKF (variant 1, :473-580; projects with fx, fy, cx, cy).  state[i]: 1 usable, 2 bad,
// 3 already in vpMatched (it then occupies feature alreadyAt[i]), 4 oblique.  dist3D[i] = the norm the function compares.
// kfMatch[j] = candidate stored in vpMatched[j] by the call, or -1.
int ref_search_by_projection_sim3(int variant, const void* kK, const uint8_t* dK, int nK, const float* scaleFactors, int nlevels,
                                  int minX, int minY, int maxX, int maxY, const uint8_t* occupied, const uint8_t* state,
                                  const int32_t* alreadyAt, const float* uv, const float* depth, const float* minDist,
                                  const float* maxDist, const int32_t* level, const uint8_t* dMP, int nMP, int th,
                                  float ratioHamming, int32_t* kfMatch, float* dist3D) {
    Sim3Case c(kK, dK, nK, scaleFactors, nlevels, minX, minY, maxX, maxY, occupied, state, uv, depth, minDist, maxDist, level, dMP,
               nMP, variant == 1);
    std::vector<MapPoint*> vpMatched(nK, nullptr);
    for (int j = 0; j < nK; ++j) if (occupied[j]) vpMatched[j] = &c.resident;
    for (int i = 0; i < nMP; ++i) {
        if (state[i] == 3) vpMatched[alreadyAt[i]] = &c.mps[i];
        dist3D[i] = c.mps[i].mWorldPos.norm();
    }
    const std::vector<MapPoint*> before = vpMatched;
    Sophus::Sim3f Scw;
    ORBmatcher m(0.6f, true);
    int n;
    if (variant == 0) n = m.SearchByProjection(&c.kf, Scw, c.ptrs, vpMatched, th, ratioHamming);
    else {
        KeyFrame other;
        std::vector<KeyFrame*> kfs(nMP, &other), matchedKF(nK, nullptr);
        n = m.SearchByProjection(&c.kf, Scw, c.ptrs, kfs, vpMatched, matchedKF, th, ratioHamming);
        for (int j = 0; j < nK; ++j)
            if ((vpMatched[j] != before[j]) != (matchedKF[j] == &other)) return -1000;      // both arrays are written together
    }
    for (int j = 0; j < nK; ++j) kfMatch[j] = vpMatched[j] != before[j] ? (int)(vpMatched[j] - c.mps.data()) : -1;
    return n;
}

// ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1182-1292).  state[i]: 1 usable, 2 bad, 3 already
// a map point of pKF (at feature alreadyAt[i]), 4 oblique.  bestIdx[i] from the logged GetMapPoint calls.
int ref_fuse_sim3(const void* kK, const uint8_t* dK, int nK, const float* scaleFactors, int nlevels, int minX, int minY, int maxX,
                  int maxY, const uint8_t* occupied, const uint8_t* state, const int32_t* alreadyAt, const float* uv,
                  const float* depth, const float* minDist, const float* maxDist, const int32_t* level, const uint8_t* dMP, int nMP,
                  float th, int32_t* bestIdx, float* dist3D) {
    Sim3Case c(kK, dK, nK, scaleFactors, nlevels, minX, minY, maxX, maxY, occupied, state, uv, depth, minDist, maxDist, level, dMP,
               nMP, false);
    for (int i = 0; i < nMP; ++i) {
        if (state[i] == 3) c.kf.mvpMapPoints[alreadyAt[i]] = &c.mps[i];
        dist3D[i] = c.mps[i].mWorldPos.norm();
        bestIdx[i] = -1;
    }
    std::vector<MapPoint*> replace(nMP, nullptr);
    Sophus::Sim3f Scw;
    g_fuseLog.clear();
    ORBmatcher m(0.6f, true);
    const int nFused = m.Fuse(&c.kf, Scw, c.ptrs, th, replace);
    // GetMapPoint(bestIdx) is called once per fused point, in candidate order; AddObservation names the point when the
    // feature was free, vpReplacePoint[i] marks it when the feature was taken
    size_t e = 0;
    for (int i = 0; i < nMP && e < g_fuseLog.size(); ++i) {
        const bool added = g_fuseLog[e].first == i, replaced = replace[i] != nullptr;
        if (added || replaced) bestIdx[i] = g_fuseLog[e++].second;
    }
    return e == g_fuseLog.size() ? nFused : -1000;
}

}  // extern "C"

namespace {
void fill_sim3_kf(KeyFrame& kf, Frame& F, const void* k, const uint8_t* d, int n, const float* scaleFactors, int nlevels, int minX,
                  int minY, int maxX, int maxY) {
    F.N = n; F.mvKeysUn = keys(k, n); F.mvKeys = F.mvKeysUn;
    F.AssignFeaturesToGrid();
    kf.N = n; kf.mvKeysUn = F.mvKeysUn; kf.mvKeys = F.mvKeysUn; kf.mDescriptors = rows32(d, n);
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mfGridElementWidthInv = Frame::mfGridElementWidthInv; kf.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    kf.mnMinX = minX; kf.mnMinY = minY; kf.mnMaxX = maxX; kf.mnMaxY = maxY;
    kf.mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<size_t>>(FRAME_GRID_ROWS));
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) kf.mGrid[i][j] = F.mGrid[i][j];
}
}  // namespace

extern "C" {

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) (ORBmatcher.cc:1293-1497) with identity poses / Sim3 and
// fx = fy = 1, cx = cy = 0: the map point of feature i of key frame A sits at (u * z, v * z, z) so that it projects to the
// prescribed pixel (u, v) of the other image.  state*[i]: 0 no map point, 1 usable, 2 bad; preMatch1[i1] = -2 none, else
// vpMatches12[i1] holds a point whose index in pKF2 is preMatch1[i1] (-1: not observed there).
// match12[i1] = feature of key frame 2 whose map point the call stored in vpMatches12[i1], or -1.
int ref_search_by_sim3(const void* k1, const uint8_t* d1, int n1, const void* k2, const uint8_t* d2, int n2, const float* scaleFactors,
                       int nlevels, int minX, int minY, int maxX, int maxY, const uint8_t* state1, const int32_t* preMatch1,
                       const float* uv12, const float* depth1, const float* min1, const float* max1, const int32_t* level12,
                       const uint8_t* state2, const float* uv21, const float* depth2, const float* min2, const float* max2,
                       const int32_t* level21, float th, int32_t* match12, float* dist12, float* dist21) {
    set_grid(minX, minY, maxX, maxY);
    Frame F1, F2;
    KeyFrame kf1, kf2;
    fill_sim3_kf(kf1, F1, k1, d1, n1, scaleFactors, nlevels, minX, minY, maxX, maxY);
    fill_sim3_kf(kf2, F2, k2, d2, n2, scaleFactors, nlevels, minX, minY, maxX, maxY);
    std::vector<MapPoint> mp1(n1), mp2(n2), pre(n1);
    auto points = [](KeyFrame& kf, std::vector<MapPoint>& mp, const uint8_t* desc, int n, const uint8_t* state, const float* uv,
                     const float* depth, const float* mn, const float* mx, const int32_t* level, float* dist) {
        kf.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; ++i) {
            const float z = depth[i];
            mp[i].mWorldPos = Eigen::Vector3f(uv[2 * i] * z, uv[2 * i + 1] * z, z);
            mp[i].mDescriptor = rows32(desc + 32 * (size_t)i, 1);
            mp[i].mnTrackScaleLevel = level[i];
            mp[i].mfMinDistance = mn[i]; mp[i].mfMaxDistance = mx[i];
            mp[i].mbBad = state[i] == 2;
            if (state[i]) kf.mvpMapPoints[i] = &mp[i];
            dist[i] = mp[i].mWorldPos.norm();
        }
    };
    points(kf1, mp1, d1, n1, state1, uv12, depth1, min1, max1, level12, dist12);
    points(kf2, mp2, d2, n2, state2, uv21, depth2, min2, max2, level21, dist21);
    std::vector<MapPoint*> vpMatches12(n1, nullptr);
    for (int i = 0; i < n1; ++i)
        if (preMatch1[i] > -2) { pre[i].mnIndexInOther = preMatch1[i]; vpMatches12[i] = &pre[i]; }
    const std::vector<MapPoint*> before = vpMatches12;
    Sophus::Sim3f S12;
    ORBmatcher m(0.6f, true);
    const int n = m.SearchBySim3(&kf1, &kf2, vpMatches12, S12, th);
    for (int i = 0; i < n1; ++i) match12[i] = vpMatches12[i] != before[i] ? (int)(vpMatches12[i] - mp2.data()) : -1;
    return n;
}

}  // extern "C"
