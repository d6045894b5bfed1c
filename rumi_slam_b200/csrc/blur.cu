// K5: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of every pyramid level (R/lib_src/ORBextractor.cc:1057-1058).
//
// On 8UC1 OpenCV evaluates the separable kernel [18,34,48,56,48,34,18]/256 in fixed point: an exact 16-bit horizontal
// pass, an exact 32-bit vertical pass and ONE rounding, (acc + 32768) >> 16.  All sums are exact
// integers, so only that final rounding has to be reproduced.
//
// Register-marching design, no shared memory and no block barrier: a WARP owns a strip of 30 x 4 columns and walks
// down the rows.  Per row every lane loads ONE aligned 32-bit word (4 pixels), gets its left / right neighbour words
// with two shuffles (lanes 0 and 31 only feed their neighbours), forms the four 7-tap horizontal sums with 6 funnel
// shifts + 8 DP4A, and keeps the last 7 rows of those sums in a statically indexed register ring; the vertical pass
// is 3 adds + 4 multiply-adds per pixel on that ring (symmetric taps).  BORDER_REFLECT_101 costs nothing in the
// common case: the row index is reflected once per row (warp uniform) and the left / right image edges are patched
// with byte permutes whose selectors depend only on (width mod 4).  All levels of all frames in one launch.
#include <cstdlib>

#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

constexpr int kBlurWarps = 4;            // warps per CTA; one work item (level, strip, column block) per warp
constexpr int kBlurOutLanes = 30;        // lanes 1..30 produce output words
constexpr int kBlurStripRows = 48;       // target rows per strip (6 halo rows are recomputed per strip)
constexpr unsigned kBlurFull = 0xFFFFFFFFu;

struct BlurTable {
    int itemBase[kMaxLevels + 1];        // first work item of each level
    int nColBlocks[kMaxLevels];
    int stripRows[kMaxLevels];
    int lastWord[kMaxLevels];            // index of the 32-bit word holding pixel w-1
    uint32_t selFix[kMaxLevels];         // PRMT(prev word, last word): last word with its bytes >= w reflected
    uint32_t selRight[kMaxLevels];       // PRMT(prev word, last word): the (virtual) word right of the last word
    int wordLoads[kMaxLevels];           // 1: rows can be read as aligned 32-bit words
    int nlevels;
};

template <bool kWords>
__device__ __forceinline__ uint32_t load_row_word(const uint8_t* p, int x0, int w, bool words) {
    if (kWords || words) {
        uint32_t v;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
    }
    uint32_t v = 0;                                  // caller memory that is not 4-byte aligned: byte gather
#pragma unroll
    for (int b = 0; b < 4; ++b) v |= (uint32_t)p[min(x0 + b, w - 1) - x0] << (8 * b);
    return v;
}

// kWords: every level can be read with aligned 32-bit loads (always true for the internal pyramid buffers)
template <bool kWords>
__global__ void __launch_bounds__(kBlurWarps * 32) blur_kernel(const __grid_constant__ ChunkView cv,
                                                               const __grid_constant__ BlurTable bt) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * kBlurWarps + (threadIdx.x >> 5);
    if (item >= bt.itemBase[bt.nlevels]) return;
    int l = 0;
    while (l + 1 < bt.nlevels && item >= bt.itemBase[l + 1]) ++l;
    const int t = item - bt.itemBase[l];
    const int cb = t % bt.nColBlocks[l], strip = t / bt.nColBlocks[l];
    const LevelView sv = cv.src[l], dv = cv.blur[l];
    const int w = sv.w, h = sv.h;
    const int lastWord = bt.lastWord[l];
    const int wi = cb * kBlurOutLanes + lane - 1;                    // word column of this lane (may be a halo word)
    const int wic = min(max(wi, 0), lastWord);
    const bool isFirst = wi == 0, isLast = wi == lastWord;
    const bool words = bt.wordLoads[l] != 0;
    const uint32_t selFix = bt.selFix[l], selRight = bt.selRight[l];
    const int y0 = strip * bt.stripRows[l], y1 = min(y0 + bt.stripRows[l], h);
    const int nIn = y1 - y0 + 6;                                     // source rows y0-3 .. y1+2
    // column base pointers, made opaque so that every row address is ONE 64-bit multiply-add (row * stride + base)
    unsigned long long sBase = (unsigned long long)(sv.ptr + (long long)blockIdx.y * sv.pitch + 4 * wic);
    unsigned long long dBase = (unsigned long long)(dv.ptr + (long long)blockIdx.y * dv.pitch + 4 * wic);
    asm volatile("" : "+l"(sBase), "+l"(dBase));
    const int sStride = sv.stride, dStride = dv.stride;
    const bool store = lane >= 1 && lane <= kBlurOutLanes && wi <= lastWord;
    const int hm2 = 2 * h - 2, lastIn = nIn - 1;

    auto src_row = [&](int rr) {              // rr-th source row of the strip (clamped to the strip), reflect-101
        const int t = abs(y0 - 3 + min(rr, lastIn));
        const int gy = min(t, hm2 - t);
        return load_row_word<kWords>(reinterpret_cast<const uint8_t*>(sBase + (long long)gy * sStride), 4 * wic, w, words);
    };

    uint32_t H[7][4];
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) H[j][k] = 0u;

    // one source row: neighbours by shuffle, image-edge reflection by byte permutes, horizontal sums into ring slot j,
    // then (kOut) one output row from the ring
    auto process = [&](const uint32_t raw, const int j, const int rr, const bool kOut) {
        const uint32_t rawL = __shfl_up_sync(kBlurFull, raw, 1);
        const uint32_t cur = isLast ? __byte_perm(rawL, raw, selFix) : raw;
        uint32_t right = __shfl_down_sync(kBlurFull, cur, 1);
        if (isLast) right = __byte_perm(rawL, raw, selRight);
        uint32_t left = __shfl_up_sync(kBlurFull, cur, 1);
        if (isFirst) left = __byte_perm(cur, right, 0x1234);        // x = -4..-1 -> src[4], src[3], src[2], src[1]
        // horizontal 7-tap sums of the 4 pixels of `cur`: bytes k+1 .. k+7 of (left, cur, right)
        const uint32_t kLo = 0x38302212u, kHi = 0x00122230u;         // (18,34,48,56) and (48,34,18,0)
        H[j][0] = __dp4a(__funnelshift_r(left, cur, 8), kLo, __dp4a(__funnelshift_r(cur, right, 8), kHi, 0u));
        H[j][1] = __dp4a(__funnelshift_r(left, cur, 16), kLo, __dp4a(__funnelshift_r(cur, right, 16), kHi, 0u));
        H[j][2] = __dp4a(__funnelshift_r(left, cur, 24), kLo, __dp4a(__funnelshift_r(cur, right, 24), kHi, 0u));
        H[j][3] = __dp4a(cur, kLo, __dp4a(right, kHi, 0u));
        // vertical pass: ring slot j holds row y+3, slot (j+1)%7 row y-3
        if (kOut) {
            uint32_t acc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t s06 = H[(j + 1) % 7][k] + H[j][k];
                const uint32_t s15 = H[(j + 2) % 7][k] + H[(j + 6) % 7][k];
                const uint32_t s24 = H[(j + 3) % 7][k] + H[(j + 5) % 7][k];
                acc[k] = 18u * s06 + 34u * s15 + 48u * s24 + 56u * H[(j + 4) % 7][k] + 32768u;
            }
            const uint32_t o = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062),
                                           0x5410);                     // byte 2 of each accumulator = acc >> 16
            const unsigned long long dp = dBase + (long long)(y0 + rr - 6) * dStride;
            if (store) asm volatile("st.global.u32 [%0], %1;" ::"l"(dp), "r"(o) : "memory");
        }
    };

    // Prologue: rows 0..5 of the strip only fill the ring.  Main loop: two register buffers of 7 rows in ping-pong,
    // the 7 loads of the next group are in flight while a group is processed (7 independent 128-byte requests per
    // warp keep enough bytes in flight to cover the DRAM latency).  Row rr always lands in ring slot rr % 7.
    uint32_t bufA[7], bufB[7];
#pragma unroll
    for (int j = 0; j < 6; ++j) bufB[j] = src_row(j);
#pragma unroll
    for (int j = 0; j < 7; ++j) bufA[j] = src_row(6 + j);
#pragma unroll
    for (int j = 0; j < 6; ++j) process(bufB[j], j, j, false);
    for (int g = 6; g < nIn; g += 14) {
#pragma unroll
        for (int j = 0; j < 7; ++j) bufB[j] = src_row(g + 7 + j);
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (g + j < nIn) process(bufA[j], (6 + j) % 7, g + j, true);       // warp uniform
#pragma unroll
        for (int j = 0; j < 7; ++j) bufA[j] = src_row(g + 14 + j);
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (g + 7 + j < nIn) process(bufB[j], (6 + j) % 7, g + 7 + j, true);
    }
}

void launch_blur(const ChunkView& cv, const OrbConst& oc, cudaStream_t s) {
    BlurTable bt;
    bt.nlevels = oc.nlevels;
    int base = 0;
    for (int l = 0; l < oc.nlevels; ++l) {
        const LevelView& sv = cv.src[l];
        const int w = sv.w, h = sv.h;
        const int lastWord = (w - 1) >> 2, r = ((w - 1) & 3) + 1;     // r = valid bytes of the last word
        const int nwords = lastWord + 1;
        // chunks: taller strips (6 recomputed halo rows per strip: 12.5 % of the work at 48 rows, 6 % at 96).  Alone the kernel
        // gets slower that way (fewer, longer items: 76 -> 78 us), the pipelined batch faster (171.8 -> 173.4 k frames/s)
        // because it is bound by total issue slots; calls of a few frames keep the short strips (latency)
        static const int targetRows = getenv("RUMI_BLUR_ROWS") ? atoi(getenv("RUMI_BLUR_ROWS")) : 2 * kBlurStripRows;
        const int rowsPerStrip = cv.nframes >= 8 ? targetRows : kBlurStripRows;
        const int nstrips = (h + rowsPerStrip - 1) / rowsPerStrip;
        bt.itemBase[l] = base;
        bt.nColBlocks[l] = (nwords + kBlurOutLanes - 1) / kBlurOutLanes;
        bt.stripRows[l] = (h + nstrips - 1) / nstrips;
        bt.lastWord[l] = lastWord;
        // PRMT source = (prev word: indices 0-3, last word: indices 4-7); pixel w-1 sits at index 4 + r-1 and
        // BORDER_REFLECT_101 maps index i > 4+r-1 to 2*(4+r-1) - i
        uint32_t fix = 0, right = 0;
        for (int p = 0; p < 4; ++p) {
            const int iFix = 4 + p, iRight = 8 + p, edge = 4 + r - 1;
            const int a = iFix > edge ? 2 * edge - iFix : iFix;
            int b = 2 * edge - iRight;
            if (b < 0) b = 0;                                         // byte not used by any valid output
            fix |= (uint32_t)a << (4 * p);
            right |= (uint32_t)b << (4 * p);
        }
        bt.selFix[l] = fix;
        bt.selRight[l] = right;
        bt.wordLoads[l] = ((((uintptr_t)sv.ptr | (uintptr_t)sv.pitch | (uintptr_t)sv.stride) & 3) == 0 &&
                           sv.stride >= 4 * nwords) ? 1 : 0;
        base += bt.nColBlocks[l] * nstrips;
    }
    bt.itemBase[oc.nlevels] = base;
    bool allWords = true;
    for (int l = 0; l < oc.nlevels; ++l) allWords = allWords && bt.wordLoads[l] != 0;
    const dim3 grid((base + kBlurWarps - 1) / kBlurWarps, cv.nframes);
    if (allWords) blur_kernel<true><<<grid, kBlurWarps * 32, 0, s>>>(cv, bt);
    else blur_kernel<false><<<grid, kBlurWarps * 32, 0, s>>>(cv, bt);
}

}  // namespace rumi
