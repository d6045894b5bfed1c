// K1' (default): scale pyramid (ORBextractor::ComputePyramid, R/lib_src/ORBextractor.cc:1093-1112) as register-marching
// warps, ALL levels of ALL frames of a chunk in one launch.
//
// cv::resize(INTER_LINEAR) on 8UC1 is the 11-bit fixed-point bilinear of SURVEY.md A.1.  A warp owns 128 destination
// columns x 32 destination rows.  Per SOURCE row a lane loads the three aligned words that cover the <= 8 source bytes
// of its 4 outputs, aligns them with two funnel shifts, picks the (p0, p1) byte pairs with two byte permutes
// (selectors are per column group, constant down the rows) and gets the four horizontal sums from four DP2A
// (weights (a0, a1) as 16-bit pairs).  Each horizontal row is computed ONCE and kept (>> 4, as the reference does
// before the vertical pass) in a per-lane shared-memory ring; a destination row costs two 16-byte ring reads, eight
// multiply-high and one coalesced 4-pixel store.  The tile kernel of pyramid.cu (TMA-staged source boxes) needs ~46
// thread instructions per pixel for the same arithmetic, this one ~13, and the 7 dependent launches become one.
//
// Level l+1 needs level l: work items are ordered level by level inside a frame (grid.x) and an item publishes a
// completion epoch (threadfence + store) that its consumers poll (relaxed loads, then one fence).  Blocks are dispatched in index order,
// so a waiting item only ever waits for items that are already resident or finished; a bounded spin turns a lost
// dependency into an error flag instead of a hang.
#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int kAhead = 4;                     // source rows of loads in flight per lane
constexpr int kRing = 8;                       // horizontal rows kept per lane (a destination row needs 2 adjacent)

__global__ void __launch_bounds__(kPyrMarchWarps * 32) pyramid_march_kernel(const __grid_constant__ PyrMarchArgs a) {
    __shared__ uint4 ring[kPyrMarchWarps][kRing][32];
    __shared__ PyrRow rowTab[kPyrMarchWarps][kPyrStripRows];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.x enumerates (level, frame, 4 items): level-major, so that when the blocks of level l+1 are dispatched
    // the blocks of level l of EVERY frame have been dispatched before them and are mostly finished
    int l = a.levelFirst;
    while (l < a.levelLast && (int)blockIdx.x >= a.lv[l + 1].blockBase) ++l;
    const int bl = blockIdx.x - a.lv[l].blockBase;
    const int f = bl / a.lv[l].blocksPerFrame;
    const int t0 = (bl - f * a.lv[l].blocksPerFrame) * kPyrMarchWarps + warp;      // item inside (level, frame)
    if (t0 >= a.lv[l].nColBlocks * a.lv[l].nStrips) return;
    const int item = a.lv[l].itemBase + t0;
    const PyrMarchLevel& L = a.lv[l];
    const int t = t0;
    const int cb = t % L.nColBlocks, strip = t / L.nColBlocks;
    const LevelView sv = a.cv.src[l - 1], dv = a.cv.src[l];
    const int y0 = strip * L.stripRows, y1 = min(y0 + L.stripRows, dv.h);
    const int g = cb * 32 + lane;                                  // destination column group of this lane
    const bool live = g < L.groups;
    const PyrColGroup cg = L.cols[live ? g : L.groups - 1];
    int* flags = a.flags + (long long)f * a.itemsPerFrame;

    // the strip's row table (source rows + vertical weights) goes to shared memory once: no dependent global loads
    // inside the row loop
    const int nrows = y1 - y0;
    if (lane < nrows) rowTab[warp][lane] = L.rows[y0 + lane];
    __syncwarp();
    const int r0 = rowTab[warp][0].sy0, r1 = rowTab[warp][nrows - 1].sy1;      // source rows r0 .. r1, all needed

    // ---- wait for the producers of the source rows / columns this item reads (level 1 reads the input) ----
    // Relaxed polling (no cache maintenance per poll) + ONE fence; the data itself is then read through L2 (ld.cg).
    if (l > a.levelFirst) {
        const PyrMarchLevel& P = a.lv[l - 1];
        const int gFirst = cb * 32, gLast = min(cb * 32 + 31, L.groups - 1);
        const int sc0 = 4 * L.cols[gFirst].word0, sc1 = 4 * (int)L.cols[gLast].word0 + 11;
        const int ps0 = r0 / P.stripRows, ps1 = min(r1 / P.stripRows, P.nStrips - 1);
        const int pc0 = sc0 / kPyrBlockCols, pc1 = min(sc1 / kPyrBlockCols, P.nColBlocks - 1);
        const int nc = pc1 - pc0 + 1, nflags = (ps1 - ps0 + 1) * nc;
        for (int i = lane; i < nflags; i += 32) {
            const int* fl = flags + P.itemBase + (ps0 + i / nc) * P.nColBlocks + pc0 + i % nc;
            int spins = 0;
            while (ld_relaxed(fl) != a.epoch) {
                __nanosleep(400);
                if (++spins > (1 << 22)) { *reinterpret_cast<volatile int*>(a.err) = 1; break; }
            }
        }
        __threadfence();
        __syncwarp();
    }

    const uint8_t* sBase = sv.ptr + (long long)f * sv.pitch;
    uint8_t* dBase = const_cast<uint8_t*>(dv.ptr) + (long long)f * dv.pitch + 4 * (live ? g : 0);
    const int w0 = min((int)cg.word0, L.srcLastWord), w1 = min(cg.word0 + 1, L.srcLastWord),
              w2 = min(cg.word0 + 2, L.srcLastWord);
    const uint32_t sh = cg.shift;
    const bool viaL1 = l == 1;         // the input is read-only for the whole launch; produced levels are read through L2

    auto load_row = [&](int r, uint32_t& A, uint32_t& B, uint32_t& C) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(sBase + (long long)r * sv.stride);
        if (viaL1) { A = __ldg(row + w0); B = __ldg(row + w1); C = __ldg(row + w2); }
        else { A = __ldcg(row + w0); B = __ldcg(row + w1); C = __ldcg(row + w2); }
    };
    // horizontal pass of source row r -> ring slot r % kRing:  (p0 * a0 + p1 * a1) >> 4 for the lane's 4 outputs
    auto hrow = [&](int r, uint32_t A, uint32_t B, uint32_t C) {
        const uint32_t lo = __funnelshift_r(A, B, sh), hi = __funnelshift_r(B, C, sh);     // source bytes s0 .. s0+7
        const uint32_t p01 = __byte_perm(lo, hi, cg.sel01), p23 = __byte_perm(lo, hi, cg.sel23);
        uint4 h;
        h.x = __dp2a_lo(cg.coef[0], p01, 0u) >> 4;
        h.y = __dp2a_hi(cg.coef[1], p01, 0u) >> 4;
        h.z = __dp2a_lo(cg.coef[2], p23, 0u) >> 4;
        h.w = __dp2a_hi(cg.coef[3], p23, 0u) >> 4;
        ring[warp][r & (kRing - 1)][lane] = h;
    };

    // Source rows are consumed in order with kAhead rows of loads in flight; a destination row is emitted as soon as
    // its second source row has been filtered.
    uint32_t A[kAhead], B[kAhead], C[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d)
        if (r0 + d <= r1) load_row(r0 + d, A[d], B[d], C[d]);
    int y = y0;
    PyrRow pr = rowTab[warp][0];
    for (int rb = r0; rb <= r1; rb += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int r = rb + d;
            if (r <= r1) {                                             // warp uniform
                const uint32_t ra = A[d], rbw = B[d], rc = C[d];
                if (r + kAhead <= r1) load_row(r + kAhead, A[d], B[d], C[d]);
                hrow(r, ra, rbw, rc);
                while (y < y1 && (int)pr.sy1 == r) {
                    const uint4 h0 = ring[warp][pr.sy0 & (kRing - 1)][lane], h1 = ring[warp][pr.sy1 & (kRing - 1)][lane];
                    // resize_vcomb: (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2; b << 16 makes >> 16 a multiply-high
                    const uint32_t o0 = (__umulhi(pr.b0s, h0.x) + __umulhi(pr.b1s, h1.x) + 2u) >> 2;
                    const uint32_t o1 = (__umulhi(pr.b0s, h0.y) + __umulhi(pr.b1s, h1.y) + 2u) >> 2;
                    const uint32_t o2 = (__umulhi(pr.b0s, h0.z) + __umulhi(pr.b1s, h1.z) + 2u) >> 2;
                    const uint32_t o3 = (__umulhi(pr.b0s, h0.w) + __umulhi(pr.b1s, h1.w) + 2u) >> 2;
                    const uint32_t o = __byte_perm(__byte_perm(o0, o1, 0x0040), __byte_perm(o2, o3, 0x0040), 0x5410);
                    if (live) *reinterpret_cast<uint32_t*>(dBase + (long long)y * dv.stride) = o;
                    ++y;
                    if (y < y1) pr = rowTab[warp][y - y0];
                }
            }
        }
    }

    // ---- publish ----
    if (l < a.levelLast) {
        __threadfence();
        __syncwarp();
        if (lane == 0) st_relaxed(flags + item, a.epoch);
    }
}

void launch_pyramid_march(const PyrMarchArgs& a, cudaStream_t s) {
    if (a.itemsPerFrame <= 0) return;
    // a.levelFirst .. a.levelLast chained inside one launch (levels before levelFirst were produced by earlier launches)
    PyrMarchArgs b = a;
    int blocks = 0;
    for (int l = a.levelFirst; l <= a.levelLast; ++l) {
        b.lv[l].blocksPerFrame = (a.lv[l].nColBlocks * a.lv[l].nStrips + kPyrMarchWarps - 1) / kPyrMarchWarps;
        b.lv[l].blockBase = blocks;
        blocks += b.lv[l].blocksPerFrame * a.cv.nframes;
    }
    pyramid_march_kernel<<<blocks, kPyrMarchWarps * 32, 0, s>>>(b);
}

}  // namespace rumi
