// K1: scale pyramid (ORBextractor::ComputePyramid, R/lib_src/ORBextractor.cc:1093-1112) and
// K5: 7x7 fixed-point Gaussian of every level (R/lib_src/ORBextractor.cc:1057-1058).
//
// K1 restates cv::resize(INTER_LINEAR) on 8UC1 as the 11-bit fixed-point bilinear OpenCV uses.
// One CTA produces a 64x32 tile of level l from a source box of level l-1 staged in shared memory by ONE TMA
// bulk-tensor load (3-D tensor map: x, y, frame); a plain-load variant covers caller memory that does not meet
// TMA's 16-byte alignment rules.  The per-column / per-row (offset, a0, a1) tables are frame independent and are
// precomputed on the host (orb_geom.h).  The EDGE_THRESHOLD reflect border of the reference's mvImagePyramid is
// never read by extraction (cells start at x=16, keypoints sit >= 19 px inside) and is not materialised.
#include "kernels.cuh"
#include "orb_math.cuh"

namespace rumi {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Set when a TMA transaction did not complete in time; checked by the host after synchronisation so that a lost
// transaction becomes an error code instead of a hung GPU.
__device__ int g_tma_timeout = 0;

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1 << 20)) { atomicExch(&g_tma_timeout, 1); break; }
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

template <bool kTMA>
__global__ void __launch_bounds__(kPyrThreads)
pyramid_level_kernel(const __grid_constant__ CUtensorMap tmap, const PyramidLevelArgs a) {
    extern __shared__ __align__(128) uint8_t box[];          // [boxH][boxW] source pixels
    __shared__ __align__(8) uint64_t bar;

    const int ox = blockIdx.x * kPyrTileW, oy = blockIdx.y * kPyrTileH, f = blockIdx.z;
    const int sw = a.src.w, sh = a.src.h, dw = a.dst.w, dh = a.dst.h;
    // TMA needs the box to start on a 16-byte boundary of the row (u8: x multiple of 16, measured on B200: an
    // unaligned inner coordinate raises an illegal-instruction fault), so the box starts at the aligned column at or
    // before the first source pixel; boxW includes the 15 spare columns.
    const int sx0 = a.xc[ox].ofs & ~15, sy0 = a.yc[oy].ofs;
    const int boxW = a.boxW, boxH = a.boxH;

    if (kTMA) {
        if (threadIdx.x == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, (uint32_t)(boxW * boxH));
            tma_load_3d(box, &tmap, &bar, sx0, sy0, f);      // out-of-image part of the box is zero filled
        }
        mbar_wait(&bar, 0);
    } else {
        const uint8_t* s = a.src.ptr + (long long)f * a.src.pitch;
        for (int i = threadIdx.x; i < boxW * boxH; i += kPyrThreads) {
            const int r = i / boxW, c = i - r * boxW;
            const int gy = min(sy0 + r, sh - 1), gx = min(sx0 + c, sw - 1);
            box[i] = s[(long long)gy * a.src.stride + gx];
        }
        __syncthreads();
    }

    const int cx = (threadIdx.x & 15) * 4, ry = threadIdx.x >> 4;
    if (ox + cx >= dw) return;
    int c0[4], c1[4], xa0[4], xa1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const ResizeCoef xc = a.xc[min(ox + cx + k, dw - 1)];
        c0[k] = xc.ofs - sx0;
        c1[k] = min(xc.ofs + 1, sw - 1) - sx0;
        xa0[k] = xc.a0; xa1[k] = xc.a1;
    }
    uint8_t* d = const_cast<uint8_t*>(a.dst.ptr) + (long long)f * a.dst.pitch;
#pragma unroll
    for (int pass = 0; pass < kPyrTileH / 16; ++pass) {
        const int dy = oy + ry + 16 * pass;
        if (dy >= dh) break;
        const ResizeCoef yc = a.yc[dy];
        const uint8_t* r0 = box + (yc.ofs - sy0) * boxW;
        const uint8_t* r1 = box + (min(yc.ofs + 1, sh - 1) - sy0) * boxW;
        uchar4 o;
        uint8_t* ob = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h0 = resize_hrow(r0[c0[k]], r0[c1[k]], xa0[k], xa1[k]);
            const int h1 = resize_hrow(r1[c0[k]], r1[c1[k]], xa0[k], xa1[k]);
            ob[k] = (uint8_t)resize_vcomb(h0, h1, yc.a0, yc.a1);
        }
        *reinterpret_cast<uchar4*>(d + (long long)dy * a.dst.stride + ox + cx) = o;   // dst rows are 16-B padded
    }
}

int read_tma_timeout_flag() {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_tma_timeout, sizeof(int)) != cudaSuccess) return -1;
    return v;
}

void launch_pyramid_level(const PyramidLevelArgs& a, const CUtensorMap* tmap, cudaStream_t s) {
    dim3 grid((a.dst.w + kPyrTileW - 1) / kPyrTileW, (a.dst.h + kPyrTileH - 1) / kPyrTileH, a.nframes);
    const size_t smem = (size_t)a.boxW * a.boxH;
    if (tmap) {
        pyramid_level_kernel<true><<<grid, kPyrThreads, smem, s>>>(*tmap, a);
    } else {
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof(dummy));
        pyramid_level_kernel<false><<<grid, kPyrThreads, smem, s>>>(dummy, a);
    }
}

// ------------------------------------------------------------------------------------------------------------
// K5: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on 8UC1 == separable [18,34,48,56,48,34,18]/256 in
// fixed point, exact 16-bit horizontal pass, (acc + 32768) >> 16 after the vertical pass.
// All levels of all frames in one launch: blockIdx.x enumerates (level, tile), blockIdx.y the frame.
constexpr int kBlurTW = 64, kBlurTH = 32, kBlurThreads = 256;

struct BlurTable {
    int tileBase[kMaxLevels + 1];
    int tilesX[kMaxLevels];
    int nlevels;
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// Work per CTA: 64x32 outputs.  Source tile = 38 rows x 72 columns starting 4 columns left of the tile so that
// every row is a whole number of aligned 32-bit words.  Horizontal pass: a thread produces 4 adjacent outputs from
// three words with byte-aligning funnel shifts and two DP4A each (weights 18,34,48,56 | 48,34,18,0).  The 16-bit
// results are stored TRANSPOSED, so the vertical pass reads 10 consecutive u16 of one column as five words and
// produces 4 vertically adjacent outputs with DP2A (weights (18,34) (48,56) (48,34) (18,0)).
constexpr int kBlurSrcRows = kBlurTH + 6, kBlurSrcWords = 18, kBlurSrcPitch = 76, kBlurColPitch = 42;

__global__ void __launch_bounds__(kBlurThreads) blur_kernel(const __grid_constant__ ChunkView cv,
                                                            const __grid_constant__ BlurTable bt) {
    __shared__ __align__(16) uint8_t src[kBlurSrcRows * kBlurSrcPitch];
    __shared__ __align__(16) uint16_t hbT[kBlurTW * kBlurColPitch];
    int l = 0;
    while (l + 1 < bt.nlevels && (int)blockIdx.x >= bt.tileBase[l + 1]) ++l;
    const int t = blockIdx.x - bt.tileBase[l];
    const int ox = (t % bt.tilesX[l]) * kBlurTW, oy = (t / bt.tilesX[l]) * kBlurTH;
    const LevelView sv = cv.src[l], dv = cv.blur[l];
    const int w = sv.w, h = sv.h;
    const uint8_t* s = sv.ptr + (long long)blockIdx.y * sv.pitch;
    const bool aligned = ((((uintptr_t)sv.ptr | (uintptr_t)sv.pitch | (uintptr_t)sv.stride) & 3) == 0);

    // ---- stage the source tile (reflect-101 at the image border) ----
    for (int i = threadIdx.x; i < kBlurSrcRows * kBlurSrcWords; i += kBlurThreads) {
        const int r = i / kBlurSrcWords, c = i - r * kBlurSrcWords;
        int gy = min(oy - 3 + r, h + 2);                 // rows past h+2 belong to unused outputs of a partial tile
        gy = gy < 0 ? -gy : gy;                           // BORDER_REFLECT_101, branch-free (|overshoot| <= 3 < h)
        gy = gy >= h ? 2 * h - 2 - gy : gy;
        const int gx = ox - 4 + 4 * c;
        const uint8_t* row = s + (long long)gy * sv.stride;
        uint32_t v;
        if (aligned && gx >= 0 && gx + 3 < w) {
            v = *reinterpret_cast<const uint32_t*>(row + gx);
        } else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)row[reflect101(min(gx + b, w + 2), w)] << (8 * b);
        }
        *reinterpret_cast<uint32_t*>(src + r * kBlurSrcPitch + 4 * c) = v;
    }
    __syncthreads();

    // ---- horizontal pass: item = (row r, group g of 4 outputs); consecutive threads take consecutive rows ----
    for (int i = threadIdx.x; i < kBlurSrcRows * (kBlurTW / 4); i += kBlurThreads) {
        const int g = i / kBlurSrcRows, r = i - g * kBlurSrcRows;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(src + r * kBlurSrcPitch + 4 * g);
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
        // output j (tile column 4g+4+j) uses bytes 1+j .. 7+j of the 12-byte window
        const uint32_t a0 = __funnelshift_r(w0, w1, 8), b0 = __funnelshift_r(w1, w2, 8);
        const uint32_t a1 = __funnelshift_r(w0, w1, 16), b1 = __funnelshift_r(w1, w2, 16);
        const uint32_t a2 = __funnelshift_r(w0, w1, 24), b2 = __funnelshift_r(w1, w2, 24);
        const uint32_t kLo = 0x38302212u, kHi = 0x00122230u;      // (18,34,48,56) and (48,34,18,0)
        uint16_t* o = hbT + (4 * g) * kBlurColPitch + r;
        o[0] = (uint16_t)__dp4a(a0, kLo, __dp4a(b0, kHi, 0u));
        o[kBlurColPitch] = (uint16_t)__dp4a(a1, kLo, __dp4a(b1, kHi, 0u));
        o[2 * kBlurColPitch] = (uint16_t)__dp4a(a2, kLo, __dp4a(b2, kHi, 0u));
        o[3 * kBlurColPitch] = (uint16_t)__dp4a(w1, kLo, __dp4a(w2, kHi, 0u));
    }
    __syncthreads();

    // ---- vertical pass: item = (column x, group of 4 rows) ----
    uint8_t* d = const_cast<uint8_t*>(dv.ptr) + (long long)blockIdx.y * dv.pitch;
    for (int i = threadIdx.x; i < kBlurTW * (kBlurTH / 4); i += kBlurThreads) {
        const int yg = i / kBlurTW, x = i - yg * kBlurTW;
        if (ox + x >= dv.stride) continue;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(hbT + x * kBlurColPitch + 4 * yg);
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = p[4];
        const uint32_t s0 = __funnelshift_r(w0, w1, 16), s1 = __funnelshift_r(w1, w2, 16);
        const uint32_t s2 = __funnelshift_r(w2, w3, 16), s3 = __funnelshift_r(w3, w4, 16), s4 = w4 >> 16;
        const uint32_t kA = 0x2212u, kB = 0x3830u, kC = 0x2230u, kD = 0x0012u;   // (18,34) (48,56) (48,34) (18,0)
        uint32_t acc[4];
        acc[0] = __dp2a_lo(w0, kA, __dp2a_lo(w1, kB, __dp2a_lo(w2, kC, __dp2a_lo(w3, kD, 32768u))));
        acc[1] = __dp2a_lo(s0, kA, __dp2a_lo(s1, kB, __dp2a_lo(s2, kC, __dp2a_lo(s3, kD, 32768u))));
        acc[2] = __dp2a_lo(w1, kA, __dp2a_lo(w2, kB, __dp2a_lo(w3, kC, __dp2a_lo(w4, kD, 32768u))));
        acc[3] = __dp2a_lo(s1, kA, __dp2a_lo(s2, kB, __dp2a_lo(s3, kC, __dp2a_lo(s4, kD, 32768u))));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = oy + 4 * yg + j;
            if (y < h) d[(long long)y * dv.stride + ox + x] = (uint8_t)(acc[j] >> 16);
        }
    }
}

void launch_blur(const ChunkView& cv, const OrbConst& oc, cudaStream_t s) {
    BlurTable bt;
    bt.nlevels = oc.nlevels;
    int base = 0;
    for (int l = 0; l < oc.nlevels; ++l) {
        bt.tileBase[l] = base;
        bt.tilesX[l] = (oc.lv[l].w + kBlurTW - 1) / kBlurTW;
        base += bt.tilesX[l] * ((oc.lv[l].h + kBlurTH - 1) / kBlurTH);
    }
    bt.tileBase[oc.nlevels] = base;
    blur_kernel<<<dim3(base, cv.nframes), kBlurThreads, 0, s>>>(cv, bt);
}

}  // namespace rumi
