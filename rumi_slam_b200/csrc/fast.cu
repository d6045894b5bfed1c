// K2: grid FAST-9/16 with cell-local non-maximum suppression and the ini/min threshold fallback
// (ORBextractor::ComputeKeyPointsOctTree, R/lib_src/ORBextractor.cc:726-808; cv::FAST).
//
// One WARP per 35-px grid cell, all levels of all frames in one launch.  The reference runs cv::FAST twice per
// cell (threshold 20, then 7 if nothing survived NMS); because the corner score is threshold independent and a
// pixel whose score is below the threshold can never beat a corner, ONE score tile per cell at the low threshold
// plus a per-cell vote reproduces both runs exactly.  Phases (all warp-synchronous, no block barrier):
//   0. the cell's sub-image (detection area + 3-px ring halo) is staged in shared memory with 16-byte vector
//      loads of the aligned superset of every row (coalesced uint4; byte loads only for unaligned caller memory);
//   A. lanes scan the detection area; a pretest on the 8 even ring pixels (a 9-arc always covers 4 consecutive of
//      them) rejects ~80 % of the pixels; survivors are compacted with __ballot_sync into a small warp queue;
//   B. the queue is drained 32 at a time so the exact 16-pixel arc test and the exact score always run with full
//      lanes; corners write their score into a zero-framed score tile and append themselves to a corner list;
//   C. NMS runs over the corner list only: strict '>' against the 8 neighbours INSIDE the cell's detection area
//      (outside = 0, exactly like cv::FAST on the sub-image); survivors set a bit in a per-pixel bitmap and the
//      warp votes "any survivor with score >= iniThFAST" (fallback rule, :783);
//   D. survivors are emitted in raster order from the bitmap (popc prefix over bitmap words) into a block reserved
//      with one atomicAdd per cell.  The octree kernel later walks the cells in the reference's row-major order, so
//      the result does not depend on the order of those reservations.
#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

constexpr int kFastWarps = 4;
constexpr int kCornerListCap = 512;        // corner list entries per cell; beyond that NMS scans the score tile
constexpr unsigned kFull = 0xFFFFFFFFu;

// sign bit of a -> LSB of the running mask (one SHF per ring pixel)
__device__ __forceinline__ uint32_t push_sign(uint32_t mask, int a) { return __funnelshift_l((uint32_t)a, mask, 1); }

// cornerScore<16> with 16-bit packed SIMD (sm_100a has native VIMNMX.S16x2 / VIMNMX3.S16x2).
// Ring differences d[k] = centre - ring[k] are held as 8 registers of adjacent pairs R[i] = (d[2i], d[2i+1]); the
// sliding min (and max) of 9 consecutive differences is min3 over k, k+1, k+2 followed by min3 over k, k+3, k+6:
//   w3[k] = min3(d[k], d[k+1], d[k+2])          (pairs: R[i], S[i] = (d[2i+1], d[2i+2]), R[i+1])
//   w9[k] = min3(w3[k], w3[k+3], w3[k+6])       (pairs: W[i], T[i+1] = (w3[2i+3], w3[2i+4]), W[i+3])
// score = max(max_k w9min[k], -min_k w9max[k]) - 1.  The combination avoids max(a, -b) (see orb_math.cuh).
__device__ __forceinline__ int fast_score16_packed(const int d[16]) {
    uint32_t R[8], S[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) R[i] = __byte_perm((uint32_t)d[2 * i], (uint32_t)d[2 * i + 1], 0x5410);   // (lo16, lo16)
#pragma unroll
    for (int i = 0; i < 8; ++i) S[i] = __byte_perm(R[i], R[(i + 1) & 7], 0x5432);       // (hi of R[i], lo of R[i+1])
    uint32_t Wn[8], Wx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        Wn[i] = __vimin3_s16x2(R[i], S[i], R[(i + 1) & 7]);
        Wx[i] = __vimax3_s16x2(R[i], S[i], R[(i + 1) & 7]);
    }
    uint32_t bestLo = 0x80008000u, bestHi = 0x7FFF7FFFu;      // packed running max of min9 / min of max9
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t Tn = __byte_perm(Wn[(i + 1) & 7], Wn[(i + 2) & 7], 0x5432);     // (w3[2i+3], w3[2i+4])
        const uint32_t Tx = __byte_perm(Wx[(i + 1) & 7], Wx[(i + 2) & 7], 0x5432);
        const uint32_t n9 = __vimin3_s16x2(Wn[i], Tn, Wn[(i + 3) & 7]);
        const uint32_t x9 = __vimax3_s16x2(Wx[i], Tx, Wx[(i + 3) & 7]);
        bestLo = __vmaxs2(bestLo, n9);
        bestHi = __vmins2(bestHi, x9);
    }
    const int lo0 = (int)(short)(bestLo & 0xFFFFu), lo1 = (int)(short)(bestLo >> 16);
    const int hi0 = (int)(short)(bestHi & 0xFFFFu), hi1 = (int)(short)(bestHi >> 16);
    const int lo = lo0 > lo1 ? lo0 : lo1, hi = hi0 < hi1 ? hi0 : hi1;
    const int best = (lo + hi > 0) ? lo : (0 - hi);
    return best - 1;
}

// exact score of the pixel whose centre is p (shared tile, row pitch tp); the pixel is a corner at threshold th
// iff score >= th (cv::FAST: 9 contiguous ring pixels all brighter than centre+th or all darker than centre-th).
template <int TP>
__device__ __forceinline__ int corner_score(const uint8_t* p, int tpRuntime) {
    const int tp = TP ? TP : tpRuntime;
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * tp];       d[1] = v - p[3 * tp + 1];   d[2] = v - p[2 * tp + 2];   d[3] = v - p[tp + 3];
    d[4] = v - p[3];            d[5] = v - p[-tp + 3];      d[6] = v - p[-2 * tp + 2];  d[7] = v - p[-3 * tp + 1];
    d[8] = v - p[-3 * tp];      d[9] = v - p[-3 * tp - 1];  d[10] = v - p[-2 * tp - 2]; d[11] = v - p[-tp - 3];
    d[12] = v - p[-3];          d[13] = v - p[tp - 3];      d[14] = v - p[2 * tp - 2];  d[15] = v - p[3 * tp - 1];
    return fast_score16_packed(d);
}

// magic-number division for small operands: q = n / d for n < 65536, d < 65536, magic = ceil(2^32 / d) (d > 1)
__device__ __forceinline__ uint32_t magic_of(uint32_t d) { return d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d); }
__device__ __forceinline__ uint32_t div_magic(uint32_t n, uint32_t magic) { return magic ? __umulhi(n, magic) : n; }

template <int TP>
__global__ void __launch_bounds__(kFastWarps * 32, 8) fast_kernel(const __grid_constant__ FastArgs a,
                                                                  const __grid_constant__ OrbConst oc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cellId = blockIdx.x * kFastWarps + warp;
    const int f = blockIdx.y;
    if (cellId >= oc.totalCells) return;

    int l = 0;
    while (l + 1 < oc.nlevels && cellId >= oc.lv[l + 1].cellBase) ++l;
    const LevelGeom& g = oc.lv[l];
    const int c = cellId - g.cellBase;
    const int ci = c / g.nCols, cj = c - ci * g.nCols;

    // per-warp shared memory: image tile | score tile | survivor bitmap | corner list | queue
    const int tp = TP ? TP : a.tilePitch, sp = a.scorePitch;     // TP > 0: compile-time pitch -> immediate offsets
    const size_t perWarp = (size_t)tp * a.tileRows + (size_t)sp * a.scoreRows + 4u * a.maskWords +
                           2u * kCornerListCap + 2u * 160;
    uint8_t* base = smem + (size_t)warp * ((perWarp + 15) & ~(size_t)15);
    uint8_t* tile = base;
    uint8_t* score = tile + (size_t)tp * a.tileRows;
    uint32_t* kept = reinterpret_cast<uint32_t*>(score + (size_t)sp * a.scoreRows);
    uint16_t* clist = reinterpret_cast<uint16_t*>(kept + a.maskWords);
    uint16_t* queue = clist + kCornerListCap;

    // cell geometry (:748-763)
    const int maxBX = g.w - kMinBorder, maxBY = g.h - kMinBorder;
    const int iniY = kMinBorder + ci * g.hCell, iniX = kMinBorder + cj * g.wCell;
    int maxY = iniY + g.hCell + 6, maxX = iniX + g.wCell + 6;
    int* cellCount = a.cellCount + (long long)f * oc.totalCells + cellId;
    int* cellOff = a.cellOff + (long long)f * oc.totalCells + cellId;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6) {
        if (lane == 0) { *cellCount = 0; *cellOff = 0; }
        return;
    }
    if (maxY > maxBY) maxY = maxBY;
    if (maxX > maxBX) maxX = maxBX;
    const int cw = maxX - iniX, ch = maxY - iniY;       // sub-image
    const int dw = cw - 6, dh = ch - 6;                 // detection area, origin (iniX+3, iniY+3)
    if (dw <= 0 || dh <= 0) {
        if (lane == 0) { *cellCount = 0; *cellOff = 0; }
        return;
    }

    // ---- 0. stage the sub-image: tile column ox <-> image column iniX ----
    const LevelView lv = a.cv.src[l];
    const uint8_t* img = lv.ptr + (long long)f * lv.pitch + (long long)iniY * lv.stride + iniX;
    int ox;
    if ((((uintptr_t)lv.ptr | (uintptr_t)lv.pitch | (uintptr_t)lv.stride) & 15) == 0) {
        ox = (int)((uintptr_t)img & 15);
        const int nvec = (ox + cw + 15) >> 4;           // <= 8 for cells up to 76+15 columns
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
    // zero the score tile (1-px zero frame = "neighbour outside the detection area") and the survivor bitmap
    const int npx = dw * dh, nwords = (npx + 31) >> 5;
    for (int i = lane; i < (sp * (dh + 2) + 3) / 4; i += 32) reinterpret_cast<uint32_t*>(score)[i] = 0u;
    for (int i = lane; i < nwords; i += 32) kept[i] = 0u;
    __syncwarp();

    const int th = oc.minTh;
    const uint8_t* t0 = tile + 3 * tp + 3 + ox;         // detection-area origin inside the tile
    uint8_t* s0 = score + sp + 1;
    const uint32_t magicW = magic_of((uint32_t)dw);
    int qn = 0, ncorner = 0;                            // queue fill, corner-list fill (warp uniform)
    bool overflow = false;

    auto drain = [&](int idq, bool valid) {
        int sc = 0;
        if (valid) {
            const int y = (int)div_magic((uint32_t)idq, magicW), x = idq - y * dw;
            sc = corner_score<TP>(t0 + y * tp + x, tp);
            if (sc < th) sc = 0;                         // not a corner at the low threshold
            if (sc > 0) s0[y * sp + x] = (uint8_t)sc;
        }
        const unsigned m = __ballot_sync(kFull, sc > 0);
        if (ncorner + __popc(m) <= kCornerListCap) {
            if (sc > 0) clist[ncorner + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idq;
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
    {
        const int c0 = ox + 3;                           // tile column of detection x = 0
        const int g0 = c0 >> 2, ng = ((c0 + dw - 1) >> 2) - g0 + 1;
        const uint32_t magicG = magic_of((uint32_t)ng);
        const uint32_t K = (uint32_t)(th <= 126 ? 127 - th : 0) * 0x01010101u;
        const uint32_t forceAll = th <= 126 ? 0u : 0x80808080u;
        const int ngroups = ng * dh;
        const uint8_t* trow = tile + 3 * tp;            // tile row of detection y = 0
        for (int b = 0; b < ngroups; b += 32) {
            const int i = b + lane;
            uint32_t pass4 = 0;
            int y = 0, g = 0;
            if (i < ngroups) {
                y = (int)div_magic((uint32_t)i, magicG);
                g = g0 + (i - y * ng);
                const uint8_t* r0 = trow + y * tp;
                const int wl = 4 * max(g - 1, 0), wc = 4 * g, wr = 4 * g + 4;
#define LDW(row, off) (*reinterpret_cast<const uint32_t*>((row) + (off)))
                const uint32_t ctr = LDW(r0, wc);
                uint32_t fl[8];
                {
                    const uint32_t l = LDW(r0, wl), r = LDW(r0, wr);
                    fl[2] = __vabsdiffu4(ctr, __funnelshift_r(ctr, r, 24));      // ring 4  (+3, 0)
                    fl[6] = __vabsdiffu4(ctr, __funnelshift_r(l, ctr, 8));       // ring 12 (-3, 0)
                }
                fl[0] = __vabsdiffu4(ctr, LDW(r0 + 3 * tp, wc));                 // ring 0  (0, +3)
                fl[4] = __vabsdiffu4(ctr, LDW(r0 - 3 * tp, wc));                 // ring 8  (0, -3)
                {
                    const uint8_t* rr = r0 + 2 * tp;
                    const uint32_t l = LDW(rr, wl), m = LDW(rr, wc), r = LDW(rr, wr);
                    fl[1] = __vabsdiffu4(ctr, __funnelshift_r(m, r, 16));        // ring 2  (+2, +2)
                    fl[7] = __vabsdiffu4(ctr, __funnelshift_r(l, m, 16));        // ring 14 (-2, +2)
                }
                {
                    const uint8_t* rr = r0 - 2 * tp;
                    const uint32_t l = LDW(rr, wl), m = LDW(rr, wc), r = LDW(rr, wr);
                    fl[3] = __vabsdiffu4(ctr, __funnelshift_r(m, r, 16));        // ring 6  (+2, -2)
                    fl[5] = __vabsdiffu4(ctr, __funnelshift_r(l, m, 16));        // ring 10 (-2, -2)
                }
#undef LDW
#pragma unroll
                for (int k = 0; k < 8; ++k) fl[k] = (((fl[k] & 0x7F7F7F7Fu) + K) | fl[k]) | forceAll;   // bit 7: > th
                uint32_t p2[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) p2[k] = fl[k] & fl[(k + 1) & 7];
                uint32_t any = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) any |= p2[k] & p2[(k + 2) & 7];
                // bytes of this group that are inside the detection area
                const int xlo = c0 - 4 * g, xhi = c0 + dw - 4 * g;              // valid bytes: xlo <= j < xhi
                uint32_t valid = 0x80808080u;
                if (xlo > 0) valid &= 0xFFFFFFFFu << (8 * xlo);
                if (xhi < 4) valid &= 0xFFFFFFFFu >> (8 * (4 - xhi));
                pass4 = any & valid;
            }
            const int idx0 = y * dw + 4 * g - c0;                                // detection index of byte 0
            {   // append the (up to 4) surviving pixels of every lane: popc + shuffle prefix instead of 4 ballots
                const int c = __popc(pass4);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int n = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += n;
                }
                int pos = qn + incl - c;
                uint32_t t = pass4;
                while (t) {
                    const int bit = __ffs(t) - 1;
                    t &= t - 1;
                    queue[pos++] = (uint16_t)(idx0 + (bit >> 3));
                }
                qn += __shfl_sync(kFull, incl, 31);
            }
            __syncwarp();
            // ---- B. exact score on the compacted survivors, 32 at a time ----
            while (qn >= 32) {
                qn -= 32;
                drain(queue[qn + lane], true);
                __syncwarp();
            }
        }
    }
    drain(lane < qn ? queue[lane] : 0, lane < qn);
    __syncwarp();

    if (a.dbg && f == 0 && cellId == a.dbgCell) {
        const int nb = tp * a.tileRows + sp * a.scoreRows;
        for (int i = lane; i < nb; i += 32) a.dbg[i] = tile[i];
    }

    // ---- C. NMS over the corner list ----
    int nIni = 0, nMin = 0;
    const int nmsCount = overflow ? npx : ncorner;
    for (int b = 0; b < nmsCount; b += 32) {
        bool k = false;
        int s = 0, idq = 0;
        if (b + lane < nmsCount) {
            idq = overflow ? b + lane : (int)clist[b + lane];
            const int y = (int)div_magic((uint32_t)idq, magicW), x = idq - y * dw;
            const uint8_t* q = s0 + y * sp + x;
            s = q[0];
            k = s > 0 && s > q[-1] && s > q[1] && s > q[-sp - 1] && s > q[-sp] && s > q[-sp + 1] && s > q[sp - 1] &&
                s > q[sp] && s > q[sp + 1];
        }
        if (k) atomicOr(&kept[idq >> 5], 1u << (idq & 31));
        nMin += __popc(__ballot_sync(kFull, k));
        nIni += __popc(__ballot_sync(kFull, k && s >= oc.iniTh));
    }
    __syncwarp();
    const int thEmit = nIni > 0 ? oc.iniTh : oc.minTh;         // :783 fallback on an empty cell
    const int total = nIni > 0 ? nIni : nMin;
    int off = 0;
    if (lane == 0) {
        off = total ? atomicAdd(a.levelCount + (long long)f * oc.nlevels + l, total) : 0;
        *cellCount = total;
        *cellOff = off;
    }
    off = __shfl_sync(kFull, off, 0);
    if (total == 0) return;

    // ---- D. raster-order emission from the survivor bitmap ----
    uint32_t* out = a.cand + a.candLevelOff[l] + (long long)f * g.candCap + off;
    const int relX = iniX + 3 - kMinBorder, relY = iniY + 3 - kMinBorder;   // candidate coords are relative to (16,16)
    int w = 0;
    for (int wb = 0; wb < nwords; wb += 32) {
        uint32_t bits = wb + lane < nwords ? kept[wb + lane] : 0u;
        if (nIni > 0 && bits) {                        // keep only survivors with score >= iniThFAST
            uint32_t keep = 0, t = bits;
            while (t) {
                const int bit = __ffs(t) - 1;
                t &= t - 1;
                const int idx = (wb + lane) * 32 + bit;
                const int y = (int)div_magic((uint32_t)idx, magicW), x = idx - y * dw;
                if (s0[y * sp + x] >= thEmit) keep |= 1u << bit;
            }
            bits = keep;
        }
        int cnt = __popc(bits), pre = cnt;             // inclusive prefix over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(kFull, pre, o);
            if (lane >= o) pre += n;
        }
        int pos = w + pre - cnt;
        while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            const int idx = (wb + lane) * 32 + bit;
            const int y = (int)div_magic((uint32_t)idx, magicW), x = idx - y * dw;
            out[pos++] = pack_cand(relX + x, relY + y, s0[y * sp + x]);
        }
        w += __shfl_sync(kFull, pre, 31);
    }
}

void launch_fast(const FastArgs& a, const OrbConst& oc, cudaStream_t s) {
    const size_t perWarp = ((size_t)a.tilePitch * a.tileRows + (size_t)a.scorePitch * a.scoreRows +
                            4u * a.maskWords + 2u * kCornerListCap + 2u * 160 + 15) & ~(size_t)15;
    const size_t smem = perWarp * kFastWarps;
    dim3 grid((oc.totalCells + kFastWarps - 1) / kFastWarps, a.cv.nframes);
    auto go = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<grid, kFastWarps * 32, smem, s>>>(a, oc);
    };
    switch (a.tilePitch) {
        case 64: go(fast_kernel<64>); break;
        case 80: go(fast_kernel<80>); break;
        case 96: go(fast_kernel<96>); break;
        default: go(fast_kernel<0>); break;
    }
}

}  // namespace rumi
