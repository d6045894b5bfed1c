// K1: scale pyramid (ORBextractor::ComputePyramid, R/lib_src/ORBextractor.cc:1093-1112).  (K5, the blur, is blur.cu.)
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
// A TMA transaction that does not complete in time sets the handle's mapped host flag `err`; the host reads it after
// synchronisation, so a lost transaction becomes an error code instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
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
        if (!done && ++spins > (1 << 20)) { *reinterpret_cast<volatile int*>(err) = 1; break; }
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
        mbar_wait(&bar, 0, a.err);
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

}  // namespace rumi
