// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Glue for oracle/_ref/liborb_ref.so: defines the cvstub image primitives with the cv2-pinned restatements
// of orb_oracle.cpp and exports a C entry point that drives the UNMODIFIED reference class
// ORB_SLAM3::ORBextractor (R/lib_src/ORBextractor.cc, compiled from /root/reference by oracle/Makefile).
#include "orb_oracle.cpp"          // anonymous-namespace restatements (resize, FAST, blur, atan2)
#include <opencv2/opencv.hpp>
#include "ORBextractor.h"          // the reference's own header

namespace cv {

void resize(InputArray src, OutputArray dst, Size dsize, double, double, int) {
    Mat s = src.getMat();
    dst.create(dsize, s.type());
    Mat d = dst.getMat();
    resize_linear_u8(s.data, s.cols, s.rows, s.step, d.data, d.cols, d.rows, d.step);
}

void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int) {
    Mat s = src.getMat();
    dst.create(s.rows + top + bottom, s.cols + left + right, s.type());
    Mat d = dst.getMat();
    std::vector<uchar> rowbuf(d.cols);
    std::vector<std::vector<uchar>> rows(s.rows);
    for (int y = 0; y < s.rows; ++y) rows[y].assign(s.ptr(y), s.ptr(y) + s.cols);   // src may alias dst's interior
    for (int y = 0; y < d.rows; ++y) {
        const std::vector<uchar>& r = rows[reflect101(y - top, s.rows)];
        for (int x = 0; x < d.cols; ++x) d.ptr(y)[x] = r[reflect101(x - left, s.cols)];
    }
}

void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY, int) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2);
    Mat s = src.getMat().clone();
    dst.create(s.rows, s.cols, s.type());
    Mat d = dst.getMat();
    gaussian7_u8(s.data, s.cols, s.rows, s.step, d.data, d.step);
}

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax) {
    assert(nonmax);
    Mat m = image.getMat();
    std::vector<Cand> c;
    fast_subimage(m.data, m.step, 0, 0, m.cols, m.rows, threshold, c);
    keypoints.clear();
    for (const Cand& k : c) keypoints.push_back(KeyPoint(k.x, k.y, 7.f, -1, k.response));
}

float fastAtan2(float y, float x) { return fast_atan2_deg(y, x); }

}  // namespace cv

namespace {
struct Exposed : ORB_SLAM3::ORBextractor {
    using ORB_SLAM3::ORBextractor::ORBextractor;
    using ORB_SLAM3::ORBextractor::DistributeOctTree;
};
}

extern "C" {

int ref_extract(const uint8_t* img, int W, int H, size_t stride, int nfeatures, float scaleFactor, int nlevels,
                int iniTh, int minTh, int lap0, int lap1, void* kps, uint8_t* desc, int cap, int* nkp, int* nmono) {
    cv::Mat image;
    if (img && W > 0 && H > 0) {
        image = cv::Mat(H, W, CV_8UC1);
        for (int y = 0; y < H; ++y) std::memcpy(image.ptr(y), img + (size_t)y * stride, W);
    }
    ORB_SLAM3::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    std::vector<int> lap = {lap0, lap1};
    int mono = ex(image, cv::Mat(), k, d, lap);
    if (image.empty()) return mono;     // -1
    *nkp = (int)k.size(); *nmono = mono;
    static_assert(sizeof(cv::KeyPoint) == 28, "KeyPoint layout");
    const int m = std::min((int)k.size(), cap);
    std::memcpy(kps, k.data(), (size_t)m * 28);
    for (int i = 0; i < m; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    return 0;
}

// ORBextractor::CloudFrameComputeDescriptors(image, keypoints, descriptors)  (R/lib_src/ORBextractor.cc:989-1011)
int ref_describe(const uint8_t* img, int W, int H, size_t stride, const void* kps, int n, uint8_t* desc) {
    cv::Mat image;
    if (img && W > 0 && H > 0) {
        image = cv::Mat(H, W, CV_8UC1);
        for (int y = 0; y < H; ++y) std::memcpy(image.ptr(y), img + (size_t)y * stride, W);
    }
    ORB_SLAM3::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> k(n);
    if (n) std::memcpy((void*)k.data(), kps, (size_t)n * 28);
    cv::Mat d;
    const int rc = ex.CloudFrameComputeDescriptors(image, k, d);
    for (int i = 0; i < n && rc > 0; ++i) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    return rc;
}

int ref_octree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out_xyr, int cap) {
    Exposed ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> in;
    for (int i = 0; i < n; ++i) in.push_back(cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]));
    std::vector<cv::KeyPoint> r = ex.DistributeOctTree(in, minX, maxX, minY, maxY, N, 0);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) {
        out_xyr[3 * i] = r[i].pt.x; out_xyr[3 * i + 1] = r[i].pt.y; out_xyr[3 * i + 2] = r[i].response;
    }
    return (int)r.size();
}

int ref_tables(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh,
               float* scale, float* invScale, float* sigma2, float* invSigma2) {
    ORB_SLAM3::ORBextractor ex(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(),
                       d = ex.GetInverseScaleSigmaSquares();
    for (int l = 0; l < nlevels; ++l) { scale[l] = a[l]; invScale[l] = b[l]; sigma2[l] = c[l]; invSigma2[l] = d[l]; }
    return ex.GetLevels();
}

}  // extern "C"
