// Micro-benchmark (B200): issue rate of POPC, LOP3, IADD3 and legacy mma.sync int8 (IMMA.16832) per SM per clock.
// Used to state the roofline of the Hamming top-2 kernel from measurement instead of folklore.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) pipe_kernel(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = a ^ 0x9E3779B9u, c = a + 7, d = b + 11;
    uint32_t e = a * 3, f = b * 5, g = c * 7, h = d * 9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) { a = __popc(a) + b; b = __popc(b) + c; c = __popc(c) + d; d = __popc(d) + a;
                           e = __popc(e) + f; f = __popc(f) + g; g = __popc(g) + h; h = __popc(h) + e; }
            if (OP == 1) { a = (a & b) ^ c; b = (b | c) ^ d; c = (c & d) ^ a; d = (d | a) ^ b;
                           e = (e & f) ^ g; f = (f | g) ^ h; g = (g & h) ^ e; h = (h | e) ^ f; }
            if (OP == 2) { a = a + b + c; b = b + c + d; c = c + d + a; d = d + a + b;
                           e = e + f + g; f = f + g + h; g = g + h + e; h = h + e + f; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

__global__ void __launch_bounds__(1024) imma_kernel(int* out, int iters) {
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[u][0]), "+r"(c[u][1]), "+r"(c[u][2]), "+r"(c[u][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int u = 0; u < 4; ++u) for (int j = 0; j < 4; ++j) s += c[u][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out; cudaMalloc(&out, 4ull * p.multiProcessorCount * 2 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = p.multiProcessorCount * 2, threads = 1024;
    const char* names[3] = {"POPC(+IADD)", "LOP3(2 per)", "IADD3"};
    for (int op = 0; op < 3; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (op == 0) pipe_kernel<0><<<blocks, threads>>>(out, iters, 1);
            if (op == 1) pipe_kernel<1><<<blocks, threads>>>(out, iters, 1);
            if (op == 2) pipe_kernel<2><<<blocks, threads>>>(out, iters, 1);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * iters * 16 * 8;      // statements per thread
        printf("%-12s %.3f ms  %.1f thread-stmts/clk/SM at %d MHz nominal (stmt = 1 op + 1 dependent add for POPC)\n",
               names[op], ms, ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
    }
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        imma_kernel<<<blocks, threads>>>((int*)out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * iters * 4;
    printf("IMMA.16832   %.3f ms  %.3f warp-MMAs/clk/SM  = %.1f int8 MAC/clk/SM  (%.1f TOPS dense)\n", ms,
           mmas / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3),
           mmas * 16 * 8 * 32 / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3),
           2 * mmas * 16 * 8 * 32 / (ms * 1e-3) / 1e12);
    printf("SMs %d clock %d kHz err %s\n", p.multiProcessorCount, clk, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
