// K1 (default): scale pyramid (ORBextractor::ComputePyramid, R/lib_src/ORBextractor.cc:1093-1112) -- ALL levels of ALL
// frames of a chunk in ONE launch, every level after the first read from SHARED MEMORY.
//
// cv::resize(INTER_LINEAR) on 8UC1 is the 11-bit fixed-point bilinear of SURVEY.md A.1, and level l is resized from
// level l-1 (a chain).  A CTA owns one horizontal strip of one frame and carries it through every level: it produces its
// rows of level l from its rows of level l-1, writes them to global memory (FAST / blur / describe read them later) AND
// keeps them in one of two ping-pong shared-memory buffers as the source of level l+1.  The only global reads are the
// rows of level 0; no level is ever re-read from L2, there are no inter-CTA dependencies (neighbouring strips recompute
// the 1-2 halo rows per level they both need -- identical values, a benign double store) and no launch gaps between the
// levels: seven dependent launches (~15 us each of fill, drain and L2 round trips) become seven __syncthreads().
//
// Inside a level the strip is cut into items of 4 destination columns x R destination rows, one thread per item,
// items dealt round-robin (adjacent lanes = adjacent column groups: coalesced 4-byte stores).  Per SOURCE row a thread
// loads the three aligned words covering the <= 8 source bytes of its 4 outputs, aligns them with two funnel shifts,
// picks the (p0, p1) byte pairs with two byte permutes (selectors are per column group) and gets the four horizontal
// sums from four DP2A (weights (a0, a1) as 16-bit pairs); only the last two horizontally filtered rows are kept (in
// registers): the row table gives, per destination row, how many source rows to advance (0, 1 or 2) before combining
// them with two multiply-high per pixel.  The strip's row records of the current level sit in shared memory as well.
#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

struct HRow { uint32_t x, y, z, w; };         // (p0 * a0 + p1 * a1) >> 4 of a thread's 4 outputs

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ HRow hfilter(const PyrColGroup& cg, uint32_t A, uint32_t B, uint32_t C) {
    const uint32_t lo = __funnelshift_r(A, B, cg.shift), hi = __funnelshift_r(B, C, cg.shift);   // source bytes s0 .. s0+7
    const uint32_t p01 = __byte_perm(lo, hi, cg.sel01), p23 = __byte_perm(lo, hi, cg.sel23);
    HRow h;
    h.x = __dp2a_lo(cg.coef[0], p01, 0u) >> 4;
    h.y = __dp2a_hi(cg.coef[1], p01, 0u) >> 4;
    h.z = __dp2a_lo(cg.coef[2], p23, 0u) >> 4;
    h.w = __dp2a_hi(cg.coef[3], p23, 0u) >> 4;
    return h;
}

// ((b * hx) >> 16) | (((b * hy) >> 16) << 16): the vertical term of two pixels in one register (b <= 2048, h < 2^15: the
// products fit 32 bits, each term fits 16).  Two IMAD + one PRMT.
__device__ __forceinline__ uint32_t vpair(uint32_t b, uint32_t hx, uint32_t hy) {
    return __byte_perm(b * hx, b * hy, 0x7632);
}

// resize_vcomb of a thread's four pixels: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2, two pixels per register
__device__ __forceinline__ uint32_t vfilter(uint32_t b0s, uint32_t b1s, const HRow& h0, const HRow& h1) {
    const uint32_t b0 = b0s >> 16, b1 = b1s >> 16;
    const uint32_t s01 = vpair(b0, h0.x, h0.y) + vpair(b1, h1.x, h1.y) + 0x00020002u;
    const uint32_t s23 = vpair(b0, h0.z, h0.w) + vpair(b1, h1.z, h1.w) + 0x00020002u;
    return __byte_perm(s01 >> 2, s23 >> 2, 0x6420);             // each half: (t0 + t1 + 2) >> 2 <= 255
}

__device__ __forceinline__ PyrRow lds_row(uint32_t addr) {      // one 16-byte record of the strip's row table
    PyrRow r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.b0s), "=r"(r.b1s), "=r"(r.sy1), "=r"(r.adv) : "r"(addr));
    return r;
}

// One item of level 1: destination rows [y0, y1) of one column group, source = level 0 in GLOBAL memory.  Driven by the
// source rows, kAhead of them in flight (statically indexed registers): HBM / L2 latency is covered by loads issued four
// rows before their use.  o0..o2: byte offsets of the three words, clamped to the row (caller memory has no slack).
constexpr int kAhead = 4;
__device__ __forceinline__ void march_item_global(const uint8_t* __restrict__ src, int srcStride, const PyrColGroup& cg,
                                                  int o0, int o1, int o2, uint32_t rowS, int y0, int y1,
                                                  uint8_t* __restrict__ dG, int dStride, uint32_t dS, int dSstride, bool keep) {
    PyrRow pr = lds_row(rowS);
    const int r0 = pr.sy1 - 1, r1 = lds_row(rowS + 16u * (uint32_t)(y1 - 1 - y0)).sy1;     // source rows r0 .. r1
    const uint8_t* rp = src + (long long)r0 * srcStride;
    uint32_t A[kAhead], B[kAhead], C[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d)
        if (r0 + d <= r1) {
            const uint8_t* p = rp + (long long)d * srcStride;
            A[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o0));
            B[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o1));
            C[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o2));
        }
    HRow hPrev, hCur;
    hCur.x = hCur.y = hCur.z = hCur.w = 0u;
    int y = y0;
    for (int rb = r0; rb <= r1; rb += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int r = rb + d;
            if (r <= r1) {
                const uint32_t ra = A[d], rbw = B[d], rc = C[d];
                if (r + kAhead <= r1) {
                    const uint8_t* p = rp + (long long)(r + kAhead - r0) * srcStride;
                    A[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o0));
                    B[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o1));
                    C[d] = __ldg(reinterpret_cast<const uint32_t*>(p + o2));
                }
                hPrev = hCur;
                hCur = hfilter(cg, ra, rbw, rc);
                while (y < y1 && pr.sy1 == r) {                   // every row whose second tap is this source row
                    const uint32_t o = vfilter(pr.b0s, pr.b1s, hPrev, hCur);
                    *reinterpret_cast<uint32_t*>(dG) = o;
                    if (keep) sts32(dS, o);
                    dG += dStride; dS += dSstride; rowS += 16u;
                    if (++y < y1) pr = lds_row(rowS);
                }
            }
        }
    }
}

// One item of a level >= 2: source = the shared-memory copy of the previous level (16 spare bytes behind every row and one
// spare row behind the last: the three words are [addr], [addr + 4], [addr + 8], the row after the current one is always
// in flight, the advance is one add).  Driven by the destination rows.  Instead of the previous horizontally filtered row
// the thread keeps the FIRST-tap terms of the coming destination row (two 16-bit terms per register), computed from the
// current row just before it is replaced: no register moves between iterations.  The record of row y carries the advance
// count of row y + 1, so one 16-byte shared load per row, issued most of an iteration before its use, feeds both.
// (adv == 0 only happens for rows clamped at the bottom of the image, whose first-tap weight is 0.)
__device__ __forceinline__ void march_item_shared(uint32_t sAddr, uint32_t sStride, const PyrColGroup& cg, uint32_t rowS,
                                                  int nrows, uint8_t* __restrict__ dG, int dStride, uint32_t dS, int dSstride,
                                                  bool keep) {
    uint32_t A, B, C;
    auto load = [&]() {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(A) : "r"(sAddr));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(B) : "r"(sAddr));
        asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(C) : "r"(sAddr));
        sAddr += sStride;
    };
    load();
    PyrRow pr = lds_row(rowS);
    HRow h = hfilter(cg, A, B, C);                // the first tap of the first row
    load();
    int adv = 1;                                  // the first row: exactly one advance after priming with its first tap
    for (int j = 0;;) {
        if (adv == 2) {                           // a source row that is only this row's FIRST tap
            h = hfilter(cg, A, B, C);
            load();
        }
        const uint32_t b0 = pr.b0s >> 16, b1 = pr.b1s >> 16;
        uint32_t t01 = vpair(b0, h.x, h.y), t23 = vpair(b0, h.z, h.w);      // (b0 == 0 when adv == 0)
        if (adv != 0) {
            h = hfilter(cg, A, B, C);
            load();
        }
        adv = pr.adv;                             // record of row y: weights of y, advance count of y + 1
        rowS += 16u;
        if (j + 1 < nrows) pr = lds_row(rowS);
        const uint32_t s01 = t01 + vpair(b1, h.x, h.y) + 0x00020002u, s23 = t23 + vpair(b1, h.z, h.w) + 0x00020002u;
        const uint32_t o = __byte_perm(s01 >> 2, s23 >> 2, 0x6420);          // each half: (t0 + t1 + 2) >> 2 <= 255
        *reinterpret_cast<uint32_t*>(dG) = o;
        if (keep) sts32(dS, o);
        if (++j >= nrows) break;
        dG += dStride; dS += dSstride;
    }
}

__global__ void __launch_bounds__(kPyrStripThreads, 1) pyramid_strip_kernel(const __grid_constant__ PyrStripArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int strip = blockIdx.x % a.nstrips, f = blockIdx.x / a.nstrips;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t rowTab = sbase + a.rowTabOffset;               // the strip's row records of the current level
    const int2* ranges = a.ranges + strip * kMaxLevels;
    int srcRow0 = 0;
    for (int l = 1; l < a.nlevels; ++l) {
        const PyrStripLevel& L = a.lv[l];
        const int2 rg = ranges[l];                                // destination rows [ya, yb) this strip computes
        const int ya = rg.x, yb = rg.y;
        const LevelView sv = a.cv.src[l - 1], dv = a.cv.src[l];
        for (int i = threadIdx.x; i < yb - ya; i += kPyrStripThreads)
            reinterpret_cast<uint4*>(smem + a.rowTabOffset)[i] = __ldg(reinterpret_cast<const uint4*>(L.rows + ya) + i);
        __syncthreads();
        uint8_t* dG = const_cast<uint8_t*>(dv.ptr) + (long long)f * dv.pitch;
        const bool keep = l + 1 < a.nlevels;                      // the last level is nobody's source
        const uint32_t dSbase = sbase + ((l & 1) ? a.buf1Offset : 0), sSbase = sbase + (((l - 1) & 1) ? a.buf1Offset : 0);
        const int dSstride = dv.stride + 16, sSstride = sv.stride + 16;   // shared copies: 16 spare bytes behind every row
        const int R = L.rowsPerItem;
        const int items = ((yb - ya + R - 1) / R) * L.groups;
        for (int it = threadIdx.x; it < items; it += kPyrStripThreads) {
            const int blk = it / L.groups, g = it - blk * L.groups;
            const PyrColGroup cg = L.cols[g];
            const int y0 = ya + blk * R, y1 = min(y0 + R, yb);
            uint8_t* d = dG + (long long)y0 * dv.stride + 4 * g;
            const uint32_t ds = dSbase + (uint32_t)((y0 - ya) * dSstride + 4 * g);
            const uint32_t rowS = rowTab + 16u * (uint32_t)(y0 - ya);
            if (l == 1) {
                march_item_global(sv.ptr + (long long)f * sv.pitch, sv.stride, cg, 4 * min((int)cg.word0, L.srcLastWord),
                                  4 * min(cg.word0 + 1, L.srcLastWord), 4 * min(cg.word0 + 2, L.srcLastWord), rowS, y0, y1, d,
                                  dv.stride, ds, dSstride, keep);
            } else {
                const int rFirst = lds_row(rowS).sy1 - 1;         // first tap of the item's first row
                march_item_shared(sSbase + (uint32_t)((rFirst - srcRow0) * sSstride) + 4u * cg.word0, (uint32_t)sSstride, cg,
                                  rowS, y1 - y0, d, dv.stride, ds, dSstride, keep);
            }
        }
        __syncthreads();
        srcRow0 = ya;
    }
}

int pyramid_strip_prepare(size_t smemBytes) {
    return (int)cudaFuncSetAttribute(pyramid_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes);
}

void launch_pyramid_strip(const PyrStripArgs& a, size_t smemBytes, cudaStream_t s) {
    pyramid_strip_kernel<<<a.cv.nframes * a.nstrips, kPyrStripThreads, smemBytes, s>>>(a);
}

}  // namespace rumi
