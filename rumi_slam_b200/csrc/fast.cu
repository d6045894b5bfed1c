// K2: grid FAST-9/16 with cell-local non-maximum suppression and the ini/min threshold fallback
// (ORBextractor::ComputeKeyPointsOctTree, R/lib_src/ORBextractor.cc:726-808; cv::FAST).
//
// One WARP per 35-px grid cell, all levels of all frames in one launch.  The reference runs cv::FAST twice per
// cell (threshold 20, then 7 if nothing survived NMS); because the corner score is threshold independent and a
// pixel whose score is below the threshold can never beat a corner, ONE score tile per cell at the low threshold
// plus a per-cell vote reproduces both runs exactly:
//   - the cell's sub-image (detection area + 3-px ring halo) is staged in shared memory with coalesced loads;
//   - lanes scan the detection area; an antipodal-pair pretest rejects most pixels after 5 loads; survivors are
//     compacted with __ballot_sync into a small warp queue so that the expensive exact score (16 ring differences,
//     sliding min/max of 9) always runs with full lanes;
//   - NMS is strict '>' against the 8 neighbours INSIDE the cell's detection area only (neighbours outside = 0),
//     exactly like cv::FAST on the sub-image;
//   - "any survivor with score >= iniThFAST" is a warp vote; survivors are emitted in raster order with ballot /
//     popc ranks into a block reserved with one atomicAdd per cell.  The octree kernel later walks the cells in
//     the reference's row-major order, so the result does not depend on the order of those reservations.
#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

constexpr int kFastWarps = 8;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ring offsets (dx, dy), OpenCV order
__device__ __constant__ int8_t c_ringDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__device__ __constant__ int8_t c_ringDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// exact test + score of one pixel whose centre is p (shared-memory tile, row pitch tp); returns 0 if not a corner
__device__ __forceinline__ int corner_score(const uint8_t* p, int tp, int th) {
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * tp];       d[1] = v - p[3 * tp + 1];   d[2] = v - p[2 * tp + 2];   d[3] = v - p[tp + 3];
    d[4] = v - p[3];            d[5] = v - p[-tp + 3];      d[6] = v - p[-2 * tp + 2];  d[7] = v - p[-3 * tp + 1];
    d[8] = v - p[-3 * tp];      d[9] = v - p[-3 * tp - 1];  d[10] = v - p[-2 * tp - 2]; d[11] = v - p[-tp - 3];
    d[12] = v - p[-3];          d[13] = v - p[tp - 3];      d[14] = v - p[2 * tp - 2];  d[15] = v - p[3 * tp - 1];
    uint32_t dark = 0, bright = 0;      // ring pixel darker / brighter than centre by more than th
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        dark |= (uint32_t)(d[k] > th) << k;
        bright |= (uint32_t)(d[k] < -th) << k;
    }
    if (!ring_has_run9(dark) && !ring_has_run9(bright)) return 0;
    return fast_score16(d);
}

__global__ void __launch_bounds__(kFastWarps * 32) fast_kernel(const __grid_constant__ FastArgs a,
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

    // per-warp shared memory: image tile, score tile, kept mask, queue
    const int tp = a.tilePitch, sp = a.scorePitch;
    const size_t perWarp = (size_t)tp * a.tileRows + (size_t)sp * a.scoreRows + 4u * a.maskWords + 2u * 64;
    uint8_t* base = smem + (size_t)warp * ((perWarp + 15) & ~(size_t)15);
    uint8_t* tile = base;
    uint8_t* score = tile + (size_t)tp * a.tileRows;
    uint32_t* kept = reinterpret_cast<uint32_t*>(score + (size_t)sp * a.scoreRows);
    uint16_t* queue = reinterpret_cast<uint16_t*>(kept + a.maskWords);

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

    // stage the sub-image (tile origin = (iniX, iniY))
    const LevelView lv = a.cv.src[l];
    const uint8_t* img = lv.ptr + (long long)f * lv.pitch + (long long)iniY * lv.stride + iniX;
    for (int r = 0; r < ch; ++r) {
        const uint8_t* row = img + (long long)r * lv.stride;
        for (int x = lane; x < cw; x += 32) tile[r * tp + x] = row[x];
    }
    // zero the score tile ((dw+2) x (dh+2), 1-px zero frame = "neighbour outside the detection area")
    for (int i = lane; i < (sp * (dh + 2) + 3) / 4; i += 32) reinterpret_cast<uint32_t*>(score)[i] = 0u;
    __syncwarp();

    const int th = oc.minTh, npx = dw * dh;
    const uint8_t* t0 = tile + 3 * tp + 3;              // detection-area origin inside the tile
    uint8_t* s0 = score + sp + 1;
    int qn = 0;                                         // queue fill (warp uniform)
    for (int b = 0; b < npx; b += 32) {
        const int idx = b + lane;
        bool pass = false;
        if (idx < npx) {
            const int y = idx / dw, x = idx - y * dw;
            const uint8_t* p = t0 + y * tp + x;
            const int v = p[0], hi = v + th, lo = v - th;
            int q0 = p[3 * tp], q8 = p[-3 * tp];
            int cls = ((q0 > hi) | (q8 > hi)) | (((q0 < lo) | (q8 < lo)) << 1);
            if (cls) {
                q0 = p[3]; q8 = p[-3];
                cls &= ((q0 > hi) | (q8 > hi)) | (((q0 < lo) | (q8 < lo)) << 1);
                pass = cls != 0;
            }
        }
        const unsigned m = __ballot_sync(kFull, pass);
        if (pass) queue[qn + __popc(m & ((1u << lane) - 1u))] = (uint16_t)idx;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            qn -= 32;
            const int idq = queue[qn + lane];
            const int y = idq / dw, x = idq - y * dw;
            const int sc = corner_score(t0 + y * tp + x, tp, th);
            if (sc > 0) s0[y * sp + x] = (uint8_t)sc;
            __syncwarp();
        }
    }
    if (lane < qn) {
        const int idq = queue[lane];
        const int y = idq / dw, x = idq - y * dw;
        const int sc = corner_score(t0 + y * tp + x, tp, th);
        if (sc > 0) s0[y * sp + x] = (uint8_t)sc;
    }
    __syncwarp();

    if (a.dbg && f == 0 && cellId == a.dbgCell) {
        const int nb = tp * a.tileRows + sp * a.scoreRows;
        for (int i = lane; i < nb; i += 32) a.dbg[i] = tile[i];
    }
    // NMS + counts at both thresholds
    int nIni = 0, nMin = 0;
    for (int b = 0, wi = 0; b < npx; b += 32, ++wi) {
        const int idx = b + lane;
        bool k = false;
        int s = 0;
        if (idx < npx) {
            const int y = idx / dw, x = idx - y * dw;
            const uint8_t* q = s0 + y * sp + x;
            s = q[0];
            if (s) {
                k = s > q[-1] && s > q[1] && s > q[-sp - 1] && s > q[-sp] && s > q[-sp + 1] && s > q[sp - 1] &&
                    s > q[sp] && s > q[sp + 1];
            }
        }
        const unsigned m = __ballot_sync(kFull, k);
        const unsigned mi = __ballot_sync(kFull, k && s >= oc.iniTh);
        if (lane == 0) kept[wi] = m;
        nMin += __popc(m);
        nIni += __popc(mi);
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
    uint32_t* out = a.cand + a.candLevelOff[l] + (long long)f * g.candCap + off;
    const int relX = iniX + 3 - kMinBorder, relY = iniY + 3 - kMinBorder;   // candidate coords are relative to (16,16)
    int w = 0;
    for (int b = 0, wi = 0; b < npx; b += 32, ++wi) {
        const int idx = b + lane;
        bool e = false;
        int s = 0, x = 0, y = 0;
        if ((kept[wi] >> lane) & 1u) {
            y = idx / dw; x = idx - y * dw;
            s = s0[y * sp + x];
            e = s >= thEmit;
        }
        const unsigned m = __ballot_sync(kFull, e);
        if (e) out[w + __popc(m & ((1u << lane) - 1u))] = pack_cand(relX + x, relY + y, s);
        w += __popc(m);
    }
}

void launch_fast(const FastArgs& a, const OrbConst& oc, cudaStream_t s) {
    const size_t perWarp = ((size_t)a.tilePitch * a.tileRows + (size_t)a.scorePitch * a.scoreRows +
                            4u * a.maskWords + 2u * 64 + 15) & ~(size_t)15;
    const size_t smem = perWarp * kFastWarps;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid((oc.totalCells + kFastWarps - 1) / kFastWarps, a.cv.nframes);
    fast_kernel<<<grid, kFastWarps * 32, smem, s>>>(a, oc);
}

}  // namespace rumi
