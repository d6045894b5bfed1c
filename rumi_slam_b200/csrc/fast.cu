// K2: grid FAST-9/16 with cell-local non-maximum suppression and the ini/min threshold fallback
// (ORBextractor::ComputeKeyPointsOctTree, R/lib_src/ORBextractor.cc:726-808; cv::FAST).
//
// One WARP per 35-px grid cell, all levels of all frames in one launch.  The reference runs cv::FAST twice per cell
// (threshold iniThFAST, then minThFAST if nothing survived NMS).  So does the warp: pass 0 at iniThFAST, and only a cell
// that ends up without a keypoint repeats the scan at minThFAST (a few per cent of the cells).  Working at the high
// threshold first matters: its pretest rejects ~3x more pixels, and corners weaker than iniThFAST never have to be
// scored -- they cannot suppress a stronger neighbour (strict '>') and are not emitted when the cell has a strong one.
// Phases of a pass (all warp-synchronous, no block barrier):
//   0. (once) the cell's sub-image (detection area + 3-px ring halo) is staged in shared memory with 16-byte vector
//      loads of the aligned superset of every row (coalesced uint4; byte loads only for unaligned caller memory);
//   A. lanes scan the detection area 4 pixels at a time; a SWAR pretest on the 8 even ring pixels (a 9-arc always
//      covers 4 consecutive of them) rejects most pixels; survivors are compacted into a small warp queue;
//   B. the queue is drained 32 at a time so the exact score always runs with full lanes; corners (score >= threshold)
//      write their score into a zero-framed score tile and append themselves to a corner list;
//   C. NMS runs over the corner list only: strict '>' against the 8 neighbours INSIDE the cell's detection area
//      (outside = 0, exactly like cv::FAST on the sub-image); survivors are compacted in place;
//   D. survivors are sorted by raster offset (warp bitonic sort) and written into a block reserved with one atomicAdd
//      per cell.  The octree kernel later walks the cells in the reference's row-major order, so the result does not
//      depend on the order of those reservations.
#include <cstdlib>

#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

constexpr int kFastWarps = 4;
constexpr int kCornerListCap = 512;        // corner list entries per cell; beyond that NMS scans the score tile
constexpr int kQueueCap = 320;             // pretest survivors waiting for the exact score (< 32 + 8 * 32)
constexpr unsigned kFull = 0xFFFFFFFFu;

// sign bit of a -> LSB of the running mask (one SHF per ring pixel)
__device__ __forceinline__ uint32_t push_sign(uint32_t mask, int a) { return __funnelshift_l((uint32_t)a, mask, 1); }

// cornerScore<16> with 16-bit packed SIMD (sm_100a has native VIMNMX3.S16x2), BOTH arc polarities in one pass.
// For ring pixel p_k and centre v, ONE multiply-add builds the packed pair
//     P[k] = p_k * 0xFFFF + K,  K = ((-v) << 16) | (v + 256)   ->   low lane = (v - p_k) + 256,  high lane = p_k - v
// (p * 0xFFFF = (p-1) << 16 | (65536 - p); the low lane sum always carries exactly once because v + 256 - p >= 1).
// The sliding minimum over 9 consecutive ring positions is min3 over k, k+1, k+2 followed by min3 over k, k+3, k+6,
// evaluated on both lanes at once: the low lane yields max_w min_w (v - p) + 256 (dark arcs), the high lane
// max_w min_w (p - v) (bright arcs); score = max(low - 256, high) - 1, exactly cv::cornerScore<16>.
__device__ __forceinline__ int fast_score16_packed(const uint32_t P[16]) {
    uint32_t w3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w3[k] = __vimin3_s16x2(P[k], P[(k + 1) & 15], P[(k + 2) & 15]);
    uint32_t best = 0x80008000u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a = __vimin3_s16x2(w3[k], w3[(k + 3) & 15], w3[(k + 6) & 15]);
        const uint32_t b = __vimin3_s16x2(w3[k + 1], w3[(k + 4) & 15], w3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a, b);
    }
    const int lo = (int)(best & 0xFFFFu) - 256, hi = (int)(short)(best >> 16);
    return (lo > hi ? lo : hi) - 1;
}

// exact score of the pixel whose centre is p (shared tile, row pitch tp); the pixel is a corner at threshold th
// iff score >= th (cv::FAST: 9 contiguous ring pixels all brighter than centre+th or all darker than centre-th).
template <int TP>
__device__ __forceinline__ int corner_score(const uint8_t* p, int tpRuntime) {
    const int tp = TP ? TP : tpRuntime;
    const uint32_t v = p[0];
    const uint32_t K = ((0u - v) << 16) | (v + 256u);
    uint32_t P[16];
#define RUMI_RING(k, off) P[k] = (uint32_t)p[off] * 0xFFFFu + K
    RUMI_RING(0, 3 * tp);       RUMI_RING(1, 3 * tp + 1);   RUMI_RING(2, 2 * tp + 2);   RUMI_RING(3, tp + 3);
    RUMI_RING(4, 3);            RUMI_RING(5, -tp + 3);      RUMI_RING(6, -2 * tp + 2);  RUMI_RING(7, -3 * tp + 1);
    RUMI_RING(8, -3 * tp);      RUMI_RING(9, -3 * tp - 1);  RUMI_RING(10, -2 * tp - 2); RUMI_RING(11, -tp - 3);
    RUMI_RING(12, -3);          RUMI_RING(13, tp - 3);      RUMI_RING(14, 2 * tp - 2);  RUMI_RING(15, 3 * tp - 1);
#undef RUMI_RING
    return fast_score16_packed(P);
}

// magic-number division for small operands: q = n / d for n < 65536, d < 65536, magic = ceil(2^32 / d) (d > 1)
__device__ __forceinline__ uint32_t magic_of(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d); }
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

// Shared-memory layout of one warp (offsets computed on the host, FastLayout).  The score tile uses the SAME pitch as
// the image tile, so one 16-bit offset o = y * tp + x (detection-area coordinates) addresses both and no phase after
// the pretest needs a division.
constexpr int kValidCap = 32;              // column groups of 4 pixels per cell row (cells are < 128 px wide)
FastLayout fast_layout(int tp, int tileRows, int scoreRows) {
    FastLayout L;
    L.score = tp * tileRows;
    L.valid = L.score + tp * scoreRows + 32;
    L.clist = L.valid + 4 * kValidCap;
    L.queue = L.clist + 2 * kCornerListCap;
    L.warpBytes = (L.queue + 2 * kQueueCap + 15) & ~15;
    return L;
}

template <int TP>
__global__ void __launch_bounds__(kFastWarps * 32, 8) fast_kernel(const __grid_constant__ FastArgs a,
                                                                  const __grid_constant__ OrbConst oc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cellId = blockIdx.x * (blockDim.x >> 5) + warp;
    const int f = blockIdx.y;
    if (cellId >= oc.totalCells) return;

    // cell geometry (:748-763), precomputed on the host (api.cu: build_fast_cells)
    const uint2 cellBits = *reinterpret_cast<const uint2*>(a.cells + cellId);     // FastCell is 8 bytes
    const int l = (int)((cellBits.y >> 16) & 0xFFu);
    int* cellCount = a.cellCount + (long long)f * oc.totalCells + cellId;
    int* cellOff = a.cellOff + (long long)f * oc.totalCells + cellId;
    if ((cellBits.y >> 24) == 0u) {                     // cell outside the level border (:752, :759)
        if (lane == 0) { *cellCount = 0; *cellOff = 0; }
        return;
    }
    const int iniX = (int)(cellBits.x & 0xFFFFu), iniY = (int)(cellBits.x >> 16);
    const int cw = (int)(cellBits.y & 0xFFu), ch = (int)((cellBits.y >> 8) & 0xFFu);       // sub-image
    const int dw = cw - 6, dh = ch - 6;                 // detection area, origin (iniX+3, iniY+3)

    // per-warp shared memory: image tile | score tile | valid-byte masks | corner list | queue
    const int tp = TP ? TP : a.tilePitch;               // TP > 0: compile-time pitch -> immediate offsets
    uint8_t* tile = smem + warp * a.lay.warpBytes;
    uint8_t* score = tile + a.lay.score;
    uint32_t* validTab = reinterpret_cast<uint32_t*>(tile + a.lay.valid);
    uint16_t* clist = reinterpret_cast<uint16_t*>(tile + a.lay.clist);
    uint16_t* queue = reinterpret_cast<uint16_t*>(tile + a.lay.queue);

    // ---- 0. stage the sub-image: tile column ox <-> image column iniX ----
    const LevelView lv = a.cv.src[l];
    const uint8_t* img = lv.ptr + (long long)f * lv.pitch + (long long)iniY * lv.stride + iniX;
    int ox;
    if ((((uintptr_t)lv.ptr | (uintptr_t)lv.pitch | (uintptr_t)lv.stride) & 15) == 0) {
        ox = (int)((uintptr_t)img & 15);
        const int nvec = (ox + cw + 15) >> 4;           // <= 6 for cells up to 76+15 columns
        const int vc = lane & 7, rr = lane >> 3;
        if (vc < nvec) {
            const uint8_t* src = img - ox + 16 * vc;
            for (int r = rr; r < ch; r += 4)
                *reinterpret_cast<uint4*>(tile + r * tp + 16 * vc) =
                    *reinterpret_cast<const uint4*>(src + (long long)r * lv.stride);
        }
    } else {
        ox = 0;
        for (int r = 0; r < ch; ++r) {
            const uint8_t* row = img + (long long)r * lv.stride;
            for (int x = lane; x < cw; x += 32) tile[r * tp + x] = row[x];
        }
    }
    // zero the score tile (everything outside the detection area stays 0 = "neighbour outside the sub-image"); per
    // column group of 4 pixels, the bytes that lie inside the detection area (bit 7 of each valid byte)
    const int c0 = ox + 3;                               // tile column of detection x = 0
    const int g0 = c0 >> 2, ng = ((c0 + dw - 1) >> 2) - g0 + 1;
    {
        uint4* z = reinterpret_cast<uint4*>(score);
        const int n16 = (tp * (dh + 2) + 32 + 15) >> 4;
        for (int i = lane; i < n16; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        if (lane < ng) {
            const int g = g0 + lane;
            const int xlo = c0 - 4 * g, xhi = c0 + dw - 4 * g;              // valid bytes: xlo <= j < xhi
            uint32_t valid = 0x80808080u;
            if (xlo > 0) valid &= 0xFFFFFFFFu << (8 * xlo);
            if (xhi < 4) valid &= 0xFFFFFFFFu >> (8 * (4 - xhi));
            validTab[lane] = valid;
        }
    }
    __syncwarp();

    // The reference runs cv::FAST at iniThFAST and, only if the cell yields NO keypoint, again at minThFAST (:771-783).
    // Same here: pass 0 at iniThFAST -- its pretest rejects far more pixels and only corners with score >= iniThFAST
    // are scored and suppressed (weaker neighbours can never block them: strict '>' against a smaller score) -- and the
    // few low-texture cells that end up empty repeat the scan at minThFAST.  Scores written by pass 0 are the same
    // values pass 1 would write, so the score tile is not cleared in between.
    int th = oc.iniTh;
    const uint8_t* t0 = tile + 3 * tp + 3 + ox;         // detection-area origin inside the image tile
    uint8_t* s0 = score + tp + 16;                      // same origin inside the score tile
    int qn = 0, ncorner = 0;                            // queue fill, corner-list fill (warp uniform)
    bool overflow = false;
    int nMin = 0;                                       // NMS survivors of the last pass
    const int npx = dw * dh;
    const uint32_t magicW = magic_of((uint32_t)dw);

    auto drain = [&](int o, bool valid) {
        int sc = 0;
        if (valid) {
            sc = corner_score<TP>(t0 + o, tp);
            if (sc < th) sc = 0;                         // not a corner at the low threshold
            if (sc > 0) s0[o] = (uint8_t)sc;
        }
        const unsigned m = __ballot_sync(kFull, sc > 0);
        if (ncorner + __popc(m) <= kCornerListCap) {
            if (sc > 0) clist[ncorner + __popc(m & ((1u << lane) - 1u))] = (uint16_t)o;
        } else {
            overflow = true;                             // too many corners for the list: NMS will scan all pixels
        }
        ncorner += __popc(m);
    };

    // ---- A. SWAR pretest, 4 horizontally adjacent pixels per lane ----
    // A 9-arc of the 16-ring always covers 4 consecutive EVEN ring positions, all differing from the centre by more
    // than th.  Per lane: the 4-byte windows of the 8 even ring positions are cut out of aligned shared-memory words
    // with funnel shifts, |centre - ring| comes from one VABSDIFF4 per position, "> th" sets bit 7 of each byte,
    // and 16 ANDs + 4 ORs give "4 consecutive positions" for the 4 pixels at once.  (Sign-agnostic, so slightly
    // weaker than the exact test: survivors go to phase B, which is exact.)
    for (int pass = 0; pass < 2; ++pass) {
        qn = 0; ncorner = 0; overflow = false; nMin = 0;
        const uint32_t K = (uint32_t)(th <= 126 ? 127 - th : 0) * 0x01010101u;
        const uint32_t forceAll = th <= 126 ? 0u : 0x80808080u;
        // a lane takes TWO horizontally adjacent column groups (8 pixels) per iteration: the aligned words they share are
        // loaded once (16 instead of 22 word loads), the index arithmetic and the loop are paid once per 8 pixels
        const int npair = (ng + 1) >> 1;
        const uint32_t magicP = magic_of((uint32_t)npair);
        const int nitems = npair * dh;
        const uint8_t* trow = tile + 3 * tp;            // tile row of detection y = 0
        // ring positions 0, 2, ..., 14 of the 4 pixels of centre word `ctr`: (l0, ctr, r0) = the words of its row, up / dn =
        // the words 3 rows below / above, (lP, mP, rP) / (lM, mM, rM) = the words 2 rows below / above
        auto test4 = [&](uint32_t ctr, uint32_t l0, uint32_t r0w, uint32_t up, uint32_t dn, uint32_t lP, uint32_t mP, uint32_t rP,
                         uint32_t lM, uint32_t mM, uint32_t rM) -> uint32_t {
            uint32_t fl[8];
            fl[2] = __vabsdiffu4(ctr, __funnelshift_r(ctr, r0w, 24));      // ring 4  (+3, 0)
            fl[6] = __vabsdiffu4(ctr, __funnelshift_r(l0, ctr, 8));        // ring 12 (-3, 0)
            fl[0] = __vabsdiffu4(ctr, up);                                 // ring 0  (0, +3)
            fl[4] = __vabsdiffu4(ctr, dn);                                 // ring 8  (0, -3)
            fl[1] = __vabsdiffu4(ctr, __funnelshift_r(mP, rP, 16));        // ring 2  (+2, +2)
            fl[7] = __vabsdiffu4(ctr, __funnelshift_r(lP, mP, 16));        // ring 14 (-2, +2)
            fl[3] = __vabsdiffu4(ctr, __funnelshift_r(mM, rM, 16));        // ring 6  (+2, -2)
            fl[5] = __vabsdiffu4(ctr, __funnelshift_r(lM, mM, 16));        // ring 10 (-2, -2)
#pragma unroll
            for (int k = 0; k < 8; ++k) fl[k] = (((fl[k] & 0x7F7F7F7Fu) + K) | fl[k]) | forceAll;   // bit 7: > th
            uint32_t p2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) p2[k] = fl[k] & fl[(k + 1) & 7];
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= p2[k] & p2[(k + 2) & 7];
            return any;
        };
        for (int b = 0; b < nitems; b += 32) {
            const int i = b + lane;
            uint32_t hitsA = 0, hitsB = 0;
            int y = 0, ga = 0;
            if (i < nitems) {
                y = (int)div_magic((uint32_t)i, magicP);
                const int pl = i - y * npair;                                   // pair inside the row
                const int gla = 2 * pl;                                         // first group of the pair (always valid)
                ga = g0 + gla;
                const uint8_t* r0 = trow + y * tp;
                const int wl = 4 * max(ga - 1, 0), wa = 4 * ga, wb = wa + 4, wr = wa + 8;
#define LDW(row, off) (*reinterpret_cast<const uint32_t*>((row) + (off)))
                const uint32_t L0 = LDW(r0, wl), A0 = LDW(r0, wa), B0 = LDW(r0, wb), R0 = LDW(r0, wr);
                const uint32_t upA = LDW(r0 + 3 * tp, wa), upB = LDW(r0 + 3 * tp, wb);
                const uint32_t dnA = LDW(r0 - 3 * tp, wa), dnB = LDW(r0 - 3 * tp, wb);
                const uint8_t* rp = r0 + 2 * tp;
                const uint32_t LP = LDW(rp, wl), AP = LDW(rp, wa), BP = LDW(rp, wb), RP = LDW(rp, wr);
                const uint8_t* rm = r0 - 2 * tp;
                const uint32_t LM = LDW(rm, wl), AM = LDW(rm, wa), BM = LDW(rm, wb), RM = LDW(rm, wr);
#undef LDW
                hitsA = test4(A0, L0, B0, upA, dnA, LP, AP, BP, LM, AM, BM) & validTab[gla];
                if (gla + 1 < ng) hitsB = test4(B0, A0, R0, upB, dnB, AP, BP, RP, AM, BM, RM) & validTab[gla + 1];
                hitsA &= 0x80808080u; hitsB &= 0x80808080u;
            }
            const int o0 = y * tp + 4 * ga - c0;                                 // offset of byte 0 of the first group
            {   // append the surviving pixels (the order inside the queue is irrelevant).  A lane has 0..8 of them: the warp
                // prefix sum of the counts comes from four ballots on the bits of the count, then every lane stores its
                // own hits back to back.
                const unsigned lt = (1u << lane) - 1u;
                const int cnt = __popc(hitsA) + __popc(hitsB);
                const unsigned b0 = __ballot_sync(kFull, cnt & 1), b1 = __ballot_sync(kFull, cnt & 2),
                               b2 = __ballot_sync(kFull, cnt & 4), b3 = __ballot_sync(kFull, cnt & 8);
                if (b0 | b1 | b2 | b3) {                                             // warp uniform
                    int pos = qn + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt) + 8 * __popc(b3 & lt);
                    qn += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (hitsA & (0x80u << (8 * j))) queue[pos++] = (uint16_t)(o0 + j);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (hitsB & (0x80u << (8 * j))) queue[pos++] = (uint16_t)(o0 + 4 + j);
                }
            }
            __syncwarp();
            // ---- B. exact score on the compacted survivors, 32 at a time ----
            while (qn >= 32) {
                qn -= 32;
                drain(queue[qn + lane], true);
                __syncwarp();
            }
        }
        if (qn > 0) drain(lane < qn ? queue[lane] : 0, lane < qn);
        __syncwarp();

        // ---- C. NMS over the corner list: strict '>' against the 8 neighbours, branch-free.  Survivors are compacted
        //         IN PLACE into the front of the list (entry i of iteration b is read before anything is written, and
        //         at most b entries precede it) ----
        const int nmsCount = overflow ? npx : ncorner;
        for (int b = 0; b < nmsCount; b += 32) {
            bool k = false;
            int s = 0, o = 0;
            if (b + lane < nmsCount) {
                if (overflow) {
                    const int idq = b + lane, y = (int)div_magic((uint32_t)idq, magicW);
                    o = y * tp + (idq - y * dw);
                } else {
                    o = (int)clist[b + lane];
                }
                const uint8_t* q = s0 + o;
                s = q[0];
                const int m0 = max(max((int)q[-tp - 1], (int)q[-tp]), (int)q[-tp + 1]);
                const int m1 = max(max((int)q[-1], (int)q[1]), (int)q[tp - 1]);
                const int m2 = max(max((int)q[tp], (int)q[tp + 1]), m0);
                k = s > max(m1, m2);                     // s == 0 (overflow scan of a non-corner) can never pass
            }
            __syncwarp();
            const unsigned mk = __ballot_sync(kFull, k);
            if (k && !overflow) clist[nMin + __popc(mk & ((1u << lane) - 1u))] = (uint16_t)o;
            nMin += __popc(mk);
        }
        __syncwarp();
        if (nMin > 0 || oc.minTh >= oc.iniTh) break;      // :783: the low threshold only for a cell without keypoints
        th = oc.minTh;
    }

    if (a.dbg && f == 0 && cellId == a.dbgCell) {
        const int nb = tp * a.tileRows + tp * a.scoreRows;
        for (int i = lane; i < nb; i += 32) a.dbg[i] = tile[i];
    }
    const int total = nMin;
    int off = 0;
    if (lane == 0) {
        off = total ? atomicAdd(a.levelCount + (long long)f * oc.nlevels + l, total) : 0;
        *cellCount = total;
        *cellOff = off;
    }
    off = __shfl_sync(kFull, off, 0);
    if (total == 0) return;

    // ---- D. raster-order emission: rank of a survivor = number of emitted survivors with a smaller offset
    //         (offset = y * tp + x is row major).  A cell keeps a handful of survivors, so the all-pairs rank is
    //         cheaper than scanning a per-pixel bitmap ----
    uint32_t* out = a.cand + a.candLevelOff[l] + (long long)f * oc.lv[l].candCap + off;
    const int relX = iniX + 3 - kMinBorder, relY = iniY + 3 - kMinBorder;   // candidate coords are relative to (16,16)
    const uint32_t magicT = magic_of((uint32_t)tp);
    if (overflow) {
        // more corners than the list holds (noise): the survivors were only counted; emit them from a second raster
        // scan of the detection area that repeats the NMS test
        int w = 0;
        for (int b = 0; b < npx; b += 32) {
            bool k = false;
            int s = 0, x = 0, y = 0;
            if (b + lane < npx) {
                y = (int)div_magic((uint32_t)(b + lane), magicW);
                x = b + lane - y * dw;
                const uint8_t* q = s0 + y * tp + x;
                s = q[0];
                const int m0 = max(max((int)q[-tp - 1], (int)q[-tp]), (int)q[-tp + 1]);
                const int m1 = max(max((int)q[-1], (int)q[1]), (int)q[tp - 1]);
                const int m2 = max(max((int)q[tp], (int)q[tp + 1]), m0);
                k = s > max(m1, m2);
            }
            const unsigned mk = __ballot_sync(kFull, k);
            if (k) out[w + __popc(mk & ((1u << lane) - 1u))] = pack_cand(relX + x, relY + y, s);
            w += __popc(mk);
        }
        return;
    }
    const int nEmit = nMin;                              // every survivor of the last pass is a keypoint candidate
    auto emit = [&](int o, int sc, int rank) {
        const int y = TP == 64 ? (o >> 6) : (int)div_magic((uint32_t)o, magicT), x = o - y * tp;
        out[rank] = pack_cand(relX + x, relY + y, sc);
    };
    if (nEmit <= 32) {
        // one key per lane, bitonic sort across the warp (15 shuffle stages): lane i ends up with the i-th offset
        uint32_t key = 0xFFFFFFFFu;
        if (lane < nEmit) { const int o = (int)clist[lane]; key = ((uint32_t)o << 8) | (uint32_t)s0[o]; }
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const uint32_t other = __shfl_xor_sync(kFull, key, j);
                const bool keepMin = ((lane & j) == 0) == ((lane & k) == 0);
                key = keepMin ? min(key, other) : max(key, other);
            }
        }
        if (lane < nEmit) emit((int)(key >> 8), (int)(key & 0xFFu), lane);
        return;
    }
    for (int i = lane; i < nEmit; i += 32) {             // many weak survivors: all-pairs rank
        const int o = (int)clist[i];
        int rank = 0;
        for (int j = 0; j < nEmit; ++j) rank += (int)clist[j] < o;
        emit(o, s0[o], rank);
    }
}

void launch_fast(const FastArgs& a, const OrbConst& oc, cudaStream_t s) {
    static int warps = 0;                      // warps (= cells) per CTA; RUMI_FAST_WARPS overrides for experiments
    if (!warps) {
        const char* e = getenv("RUMI_FAST_WARPS");
        warps = e && e[0] >= '1' && e[0] <= '0' + kFastWarps ? e[0] - '0' : 2;   // 2: less tail imbalance than 4 (measured)
    }
    const size_t smem = (size_t)a.lay.warpBytes * warps;
    dim3 grid((oc.totalCells + warps - 1) / warps, a.cv.nframes);
    auto go = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, warps * 32, smem, s>>>(a, oc);
    };
    switch (a.tilePitch) {
        case 64: go(fast_kernel<64>); break;
        case 80: go(fast_kernel<80>); break;
        case 96: go(fast_kernel<96>); break;
        default: go(fast_kernel<0>); break;
    }
}

}  // namespace rumi
