// Host/device scalar arithmetic of the ORB front-end.  Everything here is integer or explicitly-rounded float32 /
// float64 so the GPU reproduces the reference's x86-64 (no FMA) results bit for bit.
#pragma once
#include "orb_common.h"

#if defined(__CUDA_ARCH__)
#define RUMI_FMUL(a, b) __fmul_rn((a), (b))
#define RUMI_FADD(a, b) __fadd_rn((a), (b))
#define RUMI_FSUB(a, b) __fsub_rn((a), (b))
#define RUMI_FDIV(a, b) __fdiv_rn((a), (b))
#define RUMI_DMUL(a, b) __dmul_rn((a), (b))
#define RUMI_DADD(a, b) __dadd_rn((a), (b))
#define RUMI_RINT(a) __float2int_rn(a)
#else
#include <cmath>
#define RUMI_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define RUMI_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define RUMI_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define RUMI_FDIV(a, b) ((float)((float)(a) / (float)(b)))
#define RUMI_DMUL(a, b) ((double)((double)(a) * (double)(b)))
#define RUMI_DADD(a, b) ((double)((double)(a) + (double)(b)))
#define RUMI_RINT(a) ((int)nearbyintf(a))
#endif

namespace rumi {

// ---- cv::resize INTER_LINEAR, 8U: one output pixel from the two horizontally-filtered rows ----
RUMI_HD int resize_hrow(int p0, int p1, int a0, int a1) { return p0 * a0 + p1 * a1; }
RUMI_HD int resize_vcomb(int r0, int r1, int b0, int b1) {
    return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

// ---- cv::FAST 9/16 ----
// d[k] = center - ring[k].  Returns max over 16 arcs of max(min9(d), min9(-d)) - 1  (== cornerScore<16>).
RUMI_HD int fast_score16(const int d[16]) {
    // sliding-window min / max of 9 over the circular array via doubling: w2, w4, w8, then +1
    int mn[16], mx[16], t0[16], t1[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a = d[k], b = d[(k + 1) & 15];
        mn[k] = a < b ? a : b; mx[k] = a > b ? a : b;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a = mn[k], b = mn[(k + 2) & 15], c = mx[k], e = mx[(k + 2) & 15];
        t0[k] = a < b ? a : b; t1[k] = c > e ? c : e;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int a = t0[k], b = t0[(k + 4) & 15], c = t1[k], e = t1[(k + 4) & 15];
        mn[k] = a < b ? a : b; mx[k] = c > e ? c : e;
    }
    // NOTE (measured on B200, CUDA 12.9 ptxas for sm_100a): folding max(max(lo, -hi), best) into one expression is
    // compiled to VIMNMX3 with the operand negation dropped (device returned max(d)-1).  Keep the two chains separate
    // and combine them once with a compare + select instead of max(a, -b).
    int bestLo = -256, bestHi = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int x = d[(k + 8) & 15];
        const int lo = mn[k] < x ? mn[k] : x;          // min of the 9 ring differences starting at k
        const int hi = mx[k] > x ? mx[k] : x;          // max of the same 9
        bestLo = bestLo > lo ? bestLo : lo;
        bestHi = bestHi < hi ? bestHi : hi;
    }
    // max(bestLo, -bestHi) without a negated max operand:  bestLo > -bestHi  <=>  bestLo + bestHi > 0
    const int best = (bestLo + bestHi > 0) ? bestLo : (0 - bestHi);
    return best - 1;
}

// 16-bit ring masks -> is there a run of >= 9 consecutive set bits (circular)?
RUMI_HD bool ring_has_run9(uint32_t m) {
    m |= m << 16;
    uint32_t r = m & (m >> 1);
    r &= r >> 2;
    r &= r >> 4;
    r &= m >> 8;
    return (r & 0xFFFFu) != 0;
}

// ---- cv::fastAtan2 in degrees, float32, no FMA ----
RUMI_HD float fast_atan2_deg(float y, float x) {
    // p_k = (float)c_k * (float)(180/pi), product rounded to float32 (exact values as hex floats)
    const float P1 = 0x1.ca44dep+5f, P3 = -0x1.2aaddcp+4f, P5 = 0x1.1d3f7ep+3f, P7 = -0x1.4515b2p+1f;
    const float ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
    const float eps = 0x1p-52f;                 // (float)DBL_EPSILON
    float a, c, c2;
    if (ax >= ay) {
        c = RUMI_FDIV(ay, RUMI_FADD(ax, eps));
        c2 = RUMI_FMUL(c, c);
        a = RUMI_FMUL(RUMI_FADD(RUMI_FMUL(RUMI_FADD(RUMI_FMUL(RUMI_FADD(RUMI_FMUL(P7, c2), P5), c2), P3), c2), P1), c);
    } else {
        c = RUMI_FDIV(ax, RUMI_FADD(ay, eps));
        c2 = RUMI_FMUL(c, c);
        a = RUMI_FSUB(90.f,
                      RUMI_FMUL(RUMI_FADD(RUMI_FMUL(RUMI_FADD(RUMI_FMUL(RUMI_FADD(RUMI_FMUL(P7, c2), P5), c2), P3), c2), P1), c));
    }
    if (x < 0) a = RUMI_FSUB(180.f, a);
    if (y < 0) a = RUMI_FSUB(360.f, a);
    return a;
}

// ---- glibc 2.39 sinf / cosf for |x| < 120 (sysdeps/ieee754/flt-32 s_sincosf.h: double-precision reduction by
// pi/2 and the c0..c4 / s1..s3 minimax polynomials, one final rounding to float).  Constants verified against
// the container's libm.so.6 .rodata; 0 mismatches vs glibc over 3.6e8 floats in [0, 2*pi]. ----
RUMI_HD void glibc_sincosf(float y, float* sinp, float* cosp) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    double x = (double)y;
    const float ay = y < 0 ? -y : y;
    int n = 0;
    double sgn = 1.0;        // p->sign[n&3]
    double tbl = 1.0;        // -1 selects __sincosf_table[1] (negated cosine coefficients)
    if (ay < 0x1.921FB6p-1f && ay >= 0x1p-12f) {
        // |y| < pi/4: n = 0
    } else if (ay < 0x1p-12f) {
        *sinp = y; *cosp = 1.0f;
        return;
    } else {
        const double r = RUMI_DMUL(x, HPI_INV);
        n = ((int32_t)r + 0x800000) >> 24;
        x = RUMI_DADD(x, -RUMI_DMUL((double)n, HPI));
        sgn = (n & 3) == 1 || (n & 3) == 2 ? -1.0 : 1.0;
        tbl = (n & 2) ? -1.0 : 1.0;
    }
    const double xs = x * sgn;                  // exact
    const double x2 = RUMI_DMUL(x, x);
    // sine polynomial on xs
    const double x3 = RUMI_DMUL(xs, x2);
    const double s1 = RUMI_DADD(S2, RUMI_DMUL(x2, S3));
    const double x7 = RUMI_DMUL(x3, x2);
    const double s = RUMI_DADD(xs, RUMI_DMUL(x3, S1));
    const float sin_v = (float)RUMI_DADD(s, RUMI_DMUL(x7, s1));
    // cosine polynomial (table sign folded in)
    const double x4 = RUMI_DMUL(x2, x2);
    const double c2 = RUMI_DADD(tbl * C3, RUMI_DMUL(x2, tbl * C4));
    const double c1 = RUMI_DADD(tbl * C0, RUMI_DMUL(x2, tbl * C1));
    const double x6 = RUMI_DMUL(x4, x2);
    const double c = RUMI_DADD(c1, RUMI_DMUL(x4, tbl * C2));
    const float cos_v = (float)RUMI_DADD(c, RUMI_DMUL(x6, c2));
    // sinf uses sinf_poly(x*s, x2, p, n); cosf uses sinf_poly(x*s, x2, p, n^1): odd n swaps the polynomials.
    if (n & 1) { *sinp = cos_v; *cosp = sin_v; } else { *sinp = sin_v; *cosp = cos_v; }
}

// ---- quad-tree path code of a candidate (R/lib_src/ORBextractor.cc:471-522, :541-566) ----
// Root = int(x / hX); then per depth: halfX = ceil((UR.x-UL.x)/2), halfY = ceil((BR.y-UL.y)/2);
// digit = (y >= UL.y+halfY)*2 + (x >= UL.x+halfX)  (n1=0, n2=1, n3=2, n4=3).
// code = root << (2*depth) | digits, most significant digit first.
RUMI_HD uint32_t tree_code(int x, int y, float hX, int nIni, int height, int depth) {
    int root = (int)RUMI_FDIV((float)x, hX);
    if (root > nIni - 1) root = nIni - 1;     // unreachable for in-range x; keeps the table lookup safe
    int ulx = (int)RUMI_FMUL(hX, (float)root), brx = (int)RUMI_FMUL(hX, (float)(root + 1));
    int uly = 0, bry = height;
    uint32_t code = (uint32_t)root;
    for (int d = 0; d < depth; ++d) {
        const int mx = ulx + ((brx - ulx + 1) >> 1);        // ceil(w/2) for w >= 0
        const int my = uly + ((bry - uly + 1) >> 1);
        const int right = x >= mx, low = y >= my;
        code = (code << 2) | (uint32_t)(low * 2 + right);
        if (right) ulx = mx; else brx = mx;
        if (low) uly = my; else bry = my;
    }
    return code;
}

// Geometry (UL.x only is needed by compareNodes) of the node reached by following `digits` of a code prefix.
RUMI_HD int tree_node_ulx(uint32_t prefix, int ndigits, float hX) {
    const int root = (int)(prefix >> (2 * ndigits));
    int ulx = (int)RUMI_FMUL(hX, (float)root), brx = (int)RUMI_FMUL(hX, (float)(root + 1));
    for (int d = ndigits - 1; d >= 0; --d) {
        const int mx = ulx + ((brx - ulx + 1) >> 1);
        if ((prefix >> (2 * d)) & 1u) ulx = mx; else brx = mx;
    }
    return ulx;
}

}  // namespace rumi
