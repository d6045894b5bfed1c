// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for the OpenCV C++ API surface that the reference's lib_src/ORBextractor.cc touches, so
// that the UNMODIFIED reference source can be compiled in a container without OpenCV headers
// (oracle/Makefile target `ref`).  Containers and geometry types are re-implemented here; the image
// primitives (resize, FAST, GaussianBlur, fastAtan2) are declared here and defined in oracle/ref_shim.cpp
// by the cv2-pinned integer restatements of oracle/orb_oracle.cpp.
#pragma once
#include "core/persistence_stub.hpp"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

inline int cvRound(double v) { return (int)std::nearbyint(v); }
inline int cvRound(float v) { return (int)std::nearbyintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {} };

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

struct ZerosExpr { int rows, cols, type; };

class _OutputArray;
class Mat {
public:
    int rows, cols;
    size_t step;
    uchar* data;
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int t) { alloc(r, c, t); }
    Mat(Size s, int t) { alloc(s.height, s.width, t); }
    // `step` is in bytes; CV_32F (only used by the DBoW2 FORB helpers) has 4-byte elements, everything else is 8U
    void alloc(int r, int c, int t = CV_8U) {
        rows = r; cols = c; step = (size_t)c * (t == CV_32F ? 4 : 1);
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 64);
        data = buf->data();
    }
    void create(int r, int c, int t) { if (data && r == rows && c == cols && step == (size_t)c * (t == CV_32F ? 4 : 1)) return; alloc(r, c, t); }
    void release() { rows = cols = 0; step = 0; data = nullptr; buf.reset(); }
    int type() const { return CV_8UC1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    Mat roi(int x, int y, int w, int h) const {
        Mat m; m.rows = h; m.cols = w; m.step = step; m.data = data + (size_t)y * step + x; m.buf = buf; return m;
    }
    Mat operator()(const Rect& r) const { return roi(r.x, r.y, r.width, r.height); }
    Mat rowRange(int a, int b) const { return roi(0, a, cols, b - a); }
    Mat colRange(int a, int b) const { return roi(a, 0, b - a, rows); }
    Mat row(int i) const { return roi(0, i, cols, 1); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    Mat clone() const {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), ptr(y), cols);
        return m;
    }
    static ZerosExpr zeros(int r, int c, int t) { return ZerosExpr{r, c, t}; }
    // Mat = MatExpr: create() is a no-op for a same-shape buffer, then the buffer is filled in place.
    Mat& operator=(const ZerosExpr& z) {
        create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; ++y) std::memset(ptr(y), 0, step);
        return *this;
    }
    inline void copyTo(const _OutputArray& o) const;
};

class _InputArray {
public:
    const Mat* m;
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return m->empty(); }
    Mat getMat() const { return *m; }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    _OutputArray(const Mat& mm) : m(const_cast<Mat*>(&mm)) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void create(Size s, int t) const { m->create(s.height, s.width, t); }
    void release() const { *m = Mat(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline void Mat::copyTo(const _OutputArray& o) const {
    o.create(rows, cols, 0);
    for (int y = 0; y < rows; ++y) std::memmove(o.m->ptr(y), ptr(y), cols);
}

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);

struct KeyPointsFilter {   // only referenced from dead code (ComputeKeyPointsOld, ORBextractor.cc:833-979)
    static void retainBest(std::vector<KeyPoint>&, int) {}
};

}  // namespace cv
