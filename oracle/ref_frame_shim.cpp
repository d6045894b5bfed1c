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

using namespace std;

namespace ORB_SLAM3 {

#include "_ref/gen_frame_defines.inc"      // FRAME_GRID_ROWS / FRAME_GRID_COLS   R/include/cloud_edge_slam_lib/Frame.h:42-43

class GeometricCamera {};
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
    std::vector<MapPoint*> mvpMapPoints;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    GeometricCamera* mpCamera2 = nullptr;
    int NLeft = -1;
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
    static float mfGridElementWidthInv, mfGridElementHeightInv, mnMinX, mnMinY;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void ComputeStereoMatches();
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1, const bool bRight = false) const;
};
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0, Frame::mnMinX = 0, Frame::mnMinY = 0;

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
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
protected:
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

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
    Frame::mnMinX = (float)minX; Frame::mnMinY = (float)minY;
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
int ref_search_by_bow(const uint8_t* dKF, const float* angKF, const uint8_t* kfValid, const int32_t* nodeKF, int nKF,
                      const uint8_t* dF, const float* angF, const int32_t* nodeF, int nF, float ratio, int checkOri,
                      int32_t* matchF) {
    KeyFrame kf; Frame F;
    std::vector<MapPoint> mps(nKF);
    kf.mDescriptors = rows32(dKF, nKF);
    kf.mvKeysUn.resize(nKF); kf.mvpMapPoints.assign(nKF, nullptr);
    for (int i = 0; i < nKF; ++i) { kf.mvKeysUn[i].angle = angKF[i]; if (kfValid[i]) kf.mvpMapPoints[i] = &mps[i]; }
    fill_featvec(kf.mFeatVec, nodeKF, nKF);
    F.N = nF; F.mDescriptors = rows32(dF, nF); F.mvKeys.resize(nF);
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

}  // extern "C"
