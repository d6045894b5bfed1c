// Standalone probe: does a 2-D / 3-D u8 TMA tile load work on this GPU the way pyramid.cu issues it?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DIMS>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int boxW, int boxH, int x, int y, int z, int* flag) {
    extern __shared__ __align__(128) uint8_t box[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(boxW * boxH) : "memory");
        if (DIMS == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(box)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(box)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    }
    uint32_t done = 0; int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (!done && ++spins > (1 << 18)) { *flag = 1; break; }
    }
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[i] = box[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int dims = argc > 1 ? atoi(argv[1]) : 3;
    const int W = 640, H = 480, N = 2, boxW = argc > 2 ? atoi(argv[2]) : 96, boxH = argc > 3 ? atoi(argv[3]) : 41;
    const int x = argc > 4 ? atoi(argv[4]) : 75, y = argc > 5 ? atoi(argv[5]) : 37, z = 1;
    std::vector<uint8_t> h((size_t)W * H * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 24);
    uint8_t *d, *dout; int* dflag;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, boxW * boxH); cudaMalloc(&dflag, 4); cudaMemset(dflag, 0, 4);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p sizeof(map)=%zu align=%zu\n", cudaGetErrorString(e), (int)q, p, sizeof(CUtensorMap), alignof(CUtensorMap));
    CUtensorMap m;
    cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gs[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t bx[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, dims, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d dims=%d box=%dx%d at (%d,%d,%d)\n", (int)r, dims, boxW, boxH, x, y, z);
    if (r != CUDA_SUCCESS) return 2;
    if (dims == 3) probe<3><<<1, 128, boxW * boxH>>>(m, dout, boxW, boxH, x, y, z, dflag);
    else probe<2><<<1, 128, boxW * boxH>>>(m, dout, boxW, boxH, x, y, z, dflag);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 3;
    std::vector<uint8_t> o(boxW * boxH); int flag = 0;
    cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost); cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    const int zz = dims == 3 ? z : 0;
    for (int r2 = 0; r2 < boxH; ++r2) for (int c = 0; c < boxW; ++c) {
        const int gx = x + c, gy = y + r2;
        const uint8_t want = (gx < W && gy < H) ? h[(size_t)zz * W * H + (size_t)gy * W + gx] : 0;
        bad += o[r2 * boxW + c] != want;
    }
    printf("timeout flag=%d mismatches=%d\n", flag, bad);
    return 0;
}
