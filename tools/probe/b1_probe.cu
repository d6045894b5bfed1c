// Micro-benchmark: throughput of mma.sync m16n8k256 b1 and.popc as ptxas lowers it for sm_100a (B200 has no b1
// tensor path; ptxas expands it to IMMA.16832 on unpacked operands).  128 (query, train) pairs per instruction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) b1_kernel(int* out, int iters, const uint32_t* bsrc) {
    uint32_t a0 = threadIdx.x * 2654435761u, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7;
    int c[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t b0 = bsrc[(i * 4 + u) & 1023] + threadIdx.x, b1 = b0 * 13;      // B changes every MMA (train tile)
            asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
    for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) s += c[u][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    int* out; cudaMalloc(&out, 4ull * p.multiProcessorCount * 8 * 256);
    uint32_t* b; cudaMalloc(&b, 4096); cudaMemset(b, 0x5A, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 5000, blocks = p.multiProcessorCount * 8;
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); b1_kernel<<<blocks, 256>>>(out, iters, b); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * 8 * iters * 4;
    printf("b1 and.popc m16n8k256: %.3f ms, %.4f warp-MMAs/clk/SM -> %.2f pairs/clk/SM, %.3e pairs/s on %d SMs (%s)\n", ms,
           mmas / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), 128 * mmas / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3),
           128 * mmas / (ms * 1e-3), p.multiProcessorCount, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
