// Host-side tables of the ORB front-end: scale pyramid, per-level feature quotas, FAST grid, octree roots,
// bilinear-resize coefficients.  Pure float32/float64 host arithmetic that restates the reference's constructor
// and per-level set-up so the device only ever sees integers.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "orb_common.h"

namespace rumi {

inline int rint_f(float v) { return (int)std::nearbyintf(v); }      // cvRound(float): round-half-even
inline int rint_d(double v) { return (int)std::nearbyint(v); }

struct ScaleTables {
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;
};

// R/lib_src/ORBextractor.cc:405-438.  `scaleFactor` is stored as double in the class (ORBextractor.h:98).
inline ScaleTables make_scale_tables(int nfeatures, float scaleFactorArg, int nlevels) {
    ScaleTables t;
    const double sf = scaleFactorArg;
    t.scale.assign(nlevels, 1.0f); t.sigma2.assign(nlevels, 1.0f);
    t.invScale.assign(nlevels, 1.0f); t.invSigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; ++i) {
        t.scale[i] = (float)(t.scale[i - 1] * sf);
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        t.invScale[i] = 1.0f / t.scale[i];
        t.invSigma2[i] = 1.0f / t.sigma2[i];
    }
    t.quota.assign(nlevels, 0);
    float factor = (float)(1.0f / sf);
    float nd = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        t.quota[l] = rint_f(nd);
        sum += t.quota[l];
        nd *= factor;
    }
    t.quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    return t;
}

// R/lib_src/ORBextractor.cc:446-460.
inline void make_umax(int umax[16]) {
    const int HP = kHalfPatch;
    const int vmax = (int)std::floor(HP * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(HP * std::sqrt(2.f) / 2);
    const double hp2 = HP * HP;
    for (int v = 0; v <= vmax; ++v) umax[v] = rint_d(std::sqrt(hp2 - v * v));
    for (int v = HP, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// cv::resize INTER_LINEAR 8U coefficient tables (OpenCV resize.cpp): per destination
// index the source offset and the two 11-bit weights.
struct AxisCoef { std::vector<uint16_t> ofs; std::vector<int16_t> a0, a1; };
inline AxisCoef make_axis_coef(int sn, int dn) {
    AxisCoef c;
    c.ofs.resize(dn); c.a0.resize(dn); c.a1.resize(dn);
    const double scale = (double)sn / dn;
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        c.ofs[d] = (uint16_t)s;
        c.a0[d] = (int16_t)rint_f((1.f - f) * 2048.f);
        c.a1[d] = (int16_t)rint_f(f * 2048.f);
    }
    return c;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// Fills OrbConst for one image shape.  Returns 0, or a negative code when the shape cannot be processed the way
// the reference would (level too small for one 35-px cell, aspect ratio that makes nIni == 0, > 16 roots).
inline int build_orb_const(OrbConst& oc, int W, int H, int nfeatures, float scaleFactor, int nlevels, int iniTh,
                           int minTh) {
    if (nlevels < 1 || nlevels > kMaxLevels || W <= 0 || H <= 0 || nfeatures < 0) return -2;
    if (W > kMaxLevelDim - 32 || H > kMaxLevelDim - 32) return -3;
    std::memset(&oc, 0, sizeof(oc));
    ScaleTables st = make_scale_tables(nfeatures, scaleFactor, nlevels);
    oc.nlevels = nlevels; oc.nfeatures = nfeatures; oc.iniTh = iniTh; oc.minTh = minTh; oc.W = W; oc.H = H;
    make_umax(oc.umax);
    int cellBase = 0, kpBase = 0;
    long long pyrOff = 0, candOff = 0;
    for (int l = 0; l < nlevels; ++l) {
        LevelGeom& g = oc.lv[l];
        g.w = rint_f((float)W * st.invScale[l]);                       // :1095-1096
        g.h = rint_f((float)H * st.invScale[l]);
        g.stride = align_up(g.w, 16);
        g.scale = st.scale[l];
        g.patchSize = (float)(int)(31 * st.scale[l]);                  // :816 (PATCH_SIZE * float -> int)
        g.quota = st.quota[l];
        const int maxBX = g.w - kMinBorder, maxBY = g.h - kMinBorder;  // :734-735
        const float width = (float)(maxBX - kMinBorder), height = (float)(maxBY - kMinBorder);
        g.nCols = (int)(width / 35.f);                                 // :743-746
        g.nRows = (int)(height / 35.f);
        if (g.nCols <= 0 || g.nRows <= 0) return -4;                   // reference divides by zero here
        g.wCell = (int)std::ceil(width / g.nCols);
        g.hCell = (int)std::ceil(height / g.nRows);
        g.cellBase = cellBase; cellBase += g.nCols * g.nRows;
        g.nIni = (int)std::round((float)(maxBX - kMinBorder) / (maxBY - kMinBorder));   // :541
        if (g.nIni <= 0 || g.nIni > (1 << kRootBits)) return -5;
        // the first split pass runs unconditionally, so a tiny quota can still yield 4 leaves per root
        g.kpBase = kpBase; kpBase += (g.quota > 4 * g.nIni ? g.quota : 4 * g.nIni) + kOverQuota;
        g.hX = (float)(maxBX - kMinBorder) / g.nIni;                   // :543
        int span = (maxBX - kMinBorder) > (maxBY - kMinBorder) ? (maxBX - kMinBorder) : (maxBY - kMinBorder);
        int d = 1;
        while ((1 << d) < span + 2) ++d;
        g.treeDepth = d + 1 < kMaxTreeDepth ? d + 1 : kMaxTreeDepth;
        // worst case after strict 3x3 NMS: one survivor per 2x2 block of each cell's detection area
        g.candCap = ((g.wCell + 1) / 2) * ((g.hCell + 1) / 2) * g.nCols * g.nRows;
        g.pyrOff = pyrOff; pyrOff += (long long)g.stride * g.h;
        g.candOff = candOff; candOff += g.candCap;
    }
    oc.totalCells = cellBase;
    oc.kpCap = kpBase;
    return 0;
}

}  // namespace rumi
