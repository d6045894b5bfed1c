// TEST INFRASTRUCTURE (oracle) -- never linked or imported by the product path.
//
// CPU restatement of the sparse pyramidal Lucas-Kanade flow that KFDSample::Step runs on every untracked frame
// (SURVEY.md 8f rank 4):
//   R/lib_src/KFDSample.cc:131-132   calcOpticalFlowPyrLK(imprvs, imnext, old, next, status, err, Size(31,31), 2, criteria)
//   R/include/cloud_edge_slam_lib/KFDSample.h:47   criteria = TermCriteria(COUNT + EPS, 20, 0.03)
//   R/lib_src/KFDSample.cc:75-82     SelectGoodPts (status == 1)
//   R/lib_src/KFDSample.cc:181-193   Calmoptflmag (mean flow magnitude, float accumulation in index order)
//   R/include/cloud_edge_slam_lib/pd.hpp:21-40   PD::update
//
// The arithmetic is OpenCV's (third party, not vendored under /root/reference: find_package(OpenCV 3.4),
// R/CMakeLists.txt:35).  Restated from OpenCV's published algorithm (modules/video/src/lkpyramid.cpp,
// modules/imgproc/src/pyramids.cpp):
//   * buildOpticalFlowPyramid: level 0 = the image, level l = pyrDown(level l-1) (5-tap [1 4 6 4 1] separable,
//     (sum + 128) >> 8, BORDER_REFLECT_101), size (w+1)/2 x (h+1)/2; the pyramid stops at the last level whose
//     NEXT size would be <= the window; every level is read with a REFLECT_101 border of one window;
//   * calcScharrDeriv: int16 (dx, dy) with the 3/10/3 Scharr kernels, REFLECT_101 at the image edge, ZERO outside;
//   * LKTrackerInvoker: 14-bit fixed-point bilinear weights, integer patch / derivative samples, float32
//     normal equations, at most maxCount Newton steps per level, the "oscillation" half-step exit.
// Integer parts (pyramid, derivatives, patch samples) are exact and compared bit for bit against cv2.  The sums of
// the normal equations (A11, A12, A22, b1, b2) are sums of INTEGER products; OpenCV accumulates them in a
// build-dependent type and order (`acctype`: float in raster order in the scalar build, 4 float lanes over pair sums
// with SSE2, exact int64 with NEON).  This restatement uses the exact integer sum (the NEON variant): it is the
// value every build approximates and it is independent of the summation order, so the device can match it bit for
// bit.  Positions are pinned against this container's cv2 4.13.0 (x86 SIMD build) within a tolerance, status flags
// exactly (tests/test_flow_oracle.py, tests/golden/flow_kats.npz).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

inline int cv_floor(float v) { return (int)std::floor(v); }
inline int cv_round(float v) { return (int)std::nearbyintf(v); }   // round-half-even (default rounding mode)
inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    int at(int x, int y) const { return px[(size_t)reflect101(y, h) * w + reflect101(x, w)]; }
};

struct Deriv {
    int w = 0, h = 0;
    std::vector<int16_t> d;   // interleaved dx, dy
    int at(int x, int y, int c) const {
        if (x < 0 || y < 0 || x >= w || y >= h) return 0;
        return d[((size_t)y * w + x) * 2 + c];
    }
};

void pyr_down(const Image& s, Image& d) {
    d.w = (s.w + 1) / 2;
    d.h = (s.h + 1) / 2;
    d.px.assign((size_t)d.w * d.h, 0);
    std::vector<int> rows((size_t)5 * d.w);
    for (int y = 0; y < d.h; ++y) {
        for (int k = 0; k < 5; ++k) {
            const int sy = reflect101(2 * y - 2 + k, s.h);
            const uint8_t* r = &s.px[(size_t)sy * s.w];
            int* o = &rows[(size_t)k * d.w];
            for (int x = 0; x < d.w; ++x) {
                const int c = 2 * x;
                o[x] = r[reflect101(c, s.w)] * 6 + (r[reflect101(c - 1, s.w)] + r[reflect101(c + 1, s.w)]) * 4 +
                       r[reflect101(c - 2, s.w)] + r[reflect101(c + 2, s.w)];
            }
        }
        for (int x = 0; x < d.w; ++x) {
            const int v = rows[2 * d.w + x] * 6 + (rows[d.w + x] + rows[3 * d.w + x]) * 4 + rows[x] + rows[4 * d.w + x];
            d.px[(size_t)y * d.w + x] = (uint8_t)((v + 128) >> 8);
        }
    }
}

void scharr(const Image& s, Deriv& d) {
    d.w = s.w;
    d.h = s.h;
    d.d.assign((size_t)s.w * s.h * 2, 0);
    std::vector<int> t0(s.w + 2), t1(s.w + 2);
    for (int y = 0; y < s.h; ++y) {
        const uint8_t* r0 = &s.px[(size_t)(y > 0 ? y - 1 : s.h > 1 ? 1 : 0) * s.w];
        const uint8_t* r1 = &s.px[(size_t)y * s.w];
        const uint8_t* r2 = &s.px[(size_t)(y < s.h - 1 ? y + 1 : s.h > 1 ? s.h - 2 : 0) * s.w];
        for (int x = 0; x < s.w; ++x) {
            t0[x + 1] = (r0[x] + r2[x]) * 3 + r1[x] * 10;
            t1[x + 1] = r2[x] - r0[x];
        }
        const int x0 = s.w > 1 ? 1 : 0, x1 = s.w > 1 ? s.w - 2 : 0;
        t0[0] = t0[x0 + 1]; t0[s.w + 1] = t0[x1 + 1];
        t1[0] = t1[x0 + 1]; t1[s.w + 1] = t1[x1 + 1];
        for (int x = 0; x < s.w; ++x) {
            d.d[((size_t)y * s.w + x) * 2] = (int16_t)(t0[x + 2] - t0[x]);
            d.d[((size_t)y * s.w + x) * 2 + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}

int build_pyramid(const uint8_t* img, int w, int h, int win, int maxLevel, std::vector<Image>& pyr) {
    pyr.clear();
    pyr.resize(maxLevel + 1);
    pyr[0].w = w; pyr[0].h = h;
    pyr[0].px.assign(img, img + (size_t)w * h);
    int sw = w, sh = h;
    for (int level = 0; level <= maxLevel; ++level) {
        if (level) pyr_down(pyr[level - 1], pyr[level]);
        sw = (sw + 1) / 2; sh = (sh + 1) / 2;
        if (sw <= win || sh <= win) { pyr.resize(level + 1); return level; }
    }
    return maxLevel;
}

struct Weights { int w00, w01, w10, w11; };
inline Weights weights(float a, float b) {
    Weights k;
    k.w00 = cv_round((1.f - a) * (1.f - b) * (float)(1 << 14));
    k.w01 = cv_round(a * (1.f - b) * (float)(1 << 14));
    k.w10 = cv_round((1.f - a) * b * (float)(1 << 14));
    k.w11 = (1 << 14) - k.w00 - k.w01 - k.w10;
    return k;
}

}  // namespace

extern "C" {

// cv::pyrDown on a packed u8 image; dst is ((w+1)/2) x ((h+1)/2).
void flow_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) {
    Image s, d;
    s.w = w; s.h = h; s.px.assign(src, src + (size_t)w * h);
    pyr_down(s, d);
    memcpy(dst, d.px.data(), d.px.size());
}

// calcScharrDeriv: dst[h][w][2] int16.
void flow_scharr(const uint8_t* src, int w, int h, int16_t* dst) {
    Image s; Deriv d;
    s.w = w; s.h = h; s.px.assign(src, src + (size_t)w * h);
    scharr(s, d);
    memcpy(dst, d.d.data(), d.d.size() * 2);
}

// calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), maxLevel,
//                      TermCriteria(COUNT+EPS, maxCount, eps), flags = 0, minEigThreshold).
// Returns the number of pyramid levels - 1 actually used.
int flow_lk(const uint8_t* prevImg, const uint8_t* nextImg, int w, int h, const float* prevPts, int n, int win,
            int maxLevel, int maxCount, double eps, float minEigThreshold, float* nextPts, uint8_t* status,
            float* err) {
    std::vector<Image> P, N;
    int lv = build_pyramid(prevImg, w, h, win, maxLevel, P);
    int lv2 = build_pyramid(nextImg, w, h, win, lv, N);
    maxLevel = lv2 < lv ? lv2 : lv;
    maxCount = maxCount < 0 ? 0 : maxCount > 100 ? 100 : maxCount;
    double e = eps < 0 ? 0 : eps > 10 ? 10 : eps;
    e *= e;
    for (int i = 0; i < n; ++i) { status[i] = 1; if (err) err[i] = 0; }
    const float half = (win - 1) * 0.5f;
    std::vector<int16_t> Iw((size_t)win * win), dIw((size_t)win * win * 2);
    for (int level = maxLevel; level >= 0; --level) {
        const Image& I = P[level];
        const Image& J = N[level];
        Deriv D;
        scharr(I, D);
        for (int p = 0; p < n; ++p) {
            const float sc = (float)(1. / (1 << level));
            float px = prevPts[2 * p] * sc, py = prevPts[2 * p + 1] * sc;
            float nx, ny;
            if (level == maxLevel) { nx = px; ny = py; }
            else { nx = nextPts[2 * p] * 2.f; ny = nextPts[2 * p + 1] * 2.f; }
            nextPts[2 * p] = nx; nextPts[2 * p + 1] = ny;
            px -= half; py -= half;
            const int ipx = cv_floor(px), ipy = cv_floor(py);
            if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
                if (level == 0) { status[p] = 0; if (err) err[p] = 0; }
                continue;
            }
            Weights k = weights(px - ipx, py - ipy);
            int64_t iA11 = 0, iA12 = 0, iA22 = 0;
            for (int y = 0; y < win; ++y)
                for (int x = 0; x < win; ++x) {
                    const int X = ipx + x, Y = ipy + y;
                    const int iv = descale(I.at(X, Y) * k.w00 + I.at(X + 1, Y) * k.w01 + I.at(X, Y + 1) * k.w10 +
                                           I.at(X + 1, Y + 1) * k.w11, 14 - 5);
                    const int ix = descale(D.at(X, Y, 0) * k.w00 + D.at(X + 1, Y, 0) * k.w01 + D.at(X, Y + 1, 0) * k.w10 +
                                           D.at(X + 1, Y + 1, 0) * k.w11, 14);
                    const int iy = descale(D.at(X, Y, 1) * k.w00 + D.at(X + 1, Y, 1) * k.w01 + D.at(X, Y + 1, 1) * k.w10 +
                                           D.at(X + 1, Y + 1, 1) * k.w11, 14);
                    Iw[(size_t)y * win + x] = (int16_t)iv;
                    dIw[((size_t)y * win + x) * 2] = (int16_t)ix;
                    dIw[((size_t)y * win + x) * 2 + 1] = (int16_t)iy;
                    iA11 += ix * ix;
                    iA12 += ix * iy;
                    iA22 += iy * iy;
                }
            const float FLT_SCALE = 1.f / (1 << 20);
            const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float Dt = A11 * A22 - A12 * A12;
            const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < minEigThreshold || Dt < FLT_EPSILON) {
                if (level == 0) status[p] = 0;
                continue;
            }
            Dt = 1.f / Dt;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < maxCount; ++j) {
                const int inx = cv_floor(nx), iny = cv_floor(ny);
                if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                k = weights(nx - inx, ny - iny);
                int64_t ib1 = 0, ib2 = 0;
                for (int y = 0; y < win; ++y)
                    for (int x = 0; x < win; ++x) {
                        const int X = inx + x, Y = iny + y;
                        const int diff = descale(J.at(X, Y) * k.w00 + J.at(X + 1, Y) * k.w01 + J.at(X, Y + 1) * k.w10 +
                                                 J.at(X + 1, Y + 1) * k.w11, 14 - 5) - Iw[(size_t)y * win + x];
                        ib1 += diff * dIw[((size_t)y * win + x) * 2];
                        ib2 += diff * dIw[((size_t)y * win + x) * 2 + 1];
                    }
                const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                const float dx = (float)((A12 * b2 - A22 * b1) * Dt), dy = (float)((A12 * b1 - A11 * b2) * Dt);
                nx += dx; ny += dy;
                nextPts[2 * p] = nx + half; nextPts[2 * p + 1] = ny + half;
                if ((double)dx * dx + (double)dy * dy <= e) break;
                if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
                    nextPts[2 * p] -= dx * 0.5f; nextPts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[p] && err && level == 0) {
                const float ex = nextPts[2 * p] - half, ey = nextPts[2 * p + 1] - half;
                const int iex = cv_floor(ex), iey = cv_floor(ey);
                if (iex < -win || iex >= J.w || iey < -win || iey >= J.h) { status[p] = 0; continue; }
                k = weights(ex - iex, ey - iey);
                float ev = 0;
                for (int y = 0; y < win; ++y)
                    for (int x = 0; x < win; ++x) {
                        const int X = iex + x, Y = iey + y;
                        const int diff = descale(J.at(X, Y) * k.w00 + J.at(X + 1, Y) * k.w01 + J.at(X, Y + 1) * k.w10 +
                                                 J.at(X + 1, Y + 1) * k.w11, 14 - 5) - Iw[(size_t)y * win + x];
                        ev += std::abs((float)diff);
                    }
                err[p] = ev * 1.f / (32 * win * win);
            }
        }
    }
    return maxLevel;
}

// KFDSample::SelectGoodPts + Calmoptflmag (R/lib_src/KFDSample.cc:75-82,181-193): mean |next - old| over status == 1,
// float accumulation in index order; 0/0 = NaN when nothing was tracked, exactly like the reference.
float flow_mean_magnitude(const float* oldPts, const float* nextPts, const uint8_t* status, int n, int* ngood) {
    float sum = 0;
    int g = 0;
    for (int i = 0; i < n; ++i)
        if (status[i] == 1) {
            const float dx = nextPts[2 * i] - oldPts[2 * i], dy = nextPts[2 * i + 1] - oldPts[2 * i + 1];
            sum += std::sqrt(dx * dx + dy * dy);
            ++g;
        }
    if (ngood) *ngood = g;
    return sum / (float)g;
}

// PD::update (R/include/cloud_edge_slam_lib/pd.hpp:21-40). state[0] = prevInput.
float flow_pd_update(float* state, float kp, float kd, float alpha, float setpoint, float maxOutput, float input,
                     double Ts) {
    const float error = setpoint - input;
    const float diff = alpha * (state[0] - input);
    state[0] -= diff;
    float output = (float)(kp * error + kd / Ts * diff);
    if (output > maxOutput) output = maxOutput;
    return output;
}

}  // extern "C"
