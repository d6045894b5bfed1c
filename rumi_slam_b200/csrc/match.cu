// K8: all-pairs 256-bit Hamming top-2 (ORBmatcher::DescriptorDistance, R/lib_src/ORBmatcher.cc:1830-1844, inside
// the best / second-best scan every matcher shares, e.g. :253-261, and cv::BFMatcher::knnMatch(k=2),
// R/lib_src/Frame.cc:1139).  K9: merge of per-slice / per-GPU candidates.
//
// Integer / POPC-pipe bound, not HBM bound: operands are 32 B per descriptor and every (query, train) pair costs
// 8 XOR + 8 POPC.  Each thread keeps kQPT query descriptors in registers; the CTA streams train descriptors through
// shared memory in tiles that every thread reads with broadcast 128-bit loads.  A carry-save adder tree (LOP3)
// folds the eight 32-bit XOR words before counting so that only 5 POPC are issued per pair instead of 8.
// Train rows are visited in ascending index with a strict '<' update, so ties keep the earliest index exactly like
// the reference's sequential scan; the train set is additionally cut into slices (grid.y) to fill the GPU, and the
// slice results are merged in slice order with the same rule (K9) -- the merge used after the NCCL all-gather of
// per-GPU candidates is the same kernel.
#include "kernels.cuh"

namespace rumi {

constexpr int kMatchThreads = 128;
constexpr int kQPT = 2;                       // queries per thread
constexpr int kTileRows = 256;                // train descriptors per shared-memory tile (8 KB)

// popcount of the 256-bit XOR of two descriptors held as 8 words each
__device__ __forceinline__ int hamming256(const uint32_t q[8], const uint4 t0, const uint4 t1) {
    const uint32_t x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const uint32_t x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    // carry-save adders: (a,b,c) -> sum = a^b^c, carry = maj(a,b,c); both are single LOP3s
    const uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
    const uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
    const uint32_t s3 = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
    // ones: s3, x7 ; twos: c1, c2, c3
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

struct Top2 { int b1, b2, i1; };

__global__ void __launch_bounds__(kMatchThreads)
hamming_top2_kernel(const uint8_t* __restrict__ Q, int nq, const uint8_t* __restrict__ T, int nt, int sliceRows,
                    int tBase, uint64_t* __restrict__ partial /* [gridDim.y][nq] */) {
    __shared__ uint4 tile[kTileRows * 2];
    const int q0 = blockIdx.x * (kMatchThreads * kQPT) + threadIdx.x;
    uint32_t q[kQPT][8];
    Top2 best[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int qi = q0 + k * kMatchThreads;
        const uint4* src = reinterpret_cast<const uint4*>(Q) + (size_t)min(qi, nq - 1) * 2;
        const uint4 a = src[0], b = src[1];
        q[k][0] = a.x; q[k][1] = a.y; q[k][2] = a.z; q[k][3] = a.w;
        q[k][4] = b.x; q[k][5] = b.y; q[k][6] = b.z; q[k][7] = b.w;
        best[k].b1 = 256; best[k].b2 = 256; best[k].i1 = -1;
    }
    const int t0 = blockIdx.y * sliceRows, t1 = min(t0 + sliceRows, nt);
    for (int base = t0; base < t1; base += kTileRows) {
        const int rows = min(kTileRows, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 2; i += kMatchThreads)
            tile[i] = reinterpret_cast<const uint4*>(T)[(size_t)base * 2 + i];
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const uint4 a = tile[2 * r], b = tile[2 * r + 1];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) {
                const int d = hamming256(q[k], a, b);
                if (d < best[k].b1) { best[k].b2 = best[k].b1; best[k].b1 = d; best[k].i1 = base + r; }
                else if (d < best[k].b2) best[k].b2 = d;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int qi = q0 + k * kMatchThreads;
        if (qi < nq) {
            const uint32_t idx = best[k].i1 < 0 ? 0xFFFFFFFFu : (uint32_t)(best[k].i1 + tBase);
            partial[(size_t)blockIdx.y * nq + qi] =
                ((uint64_t)best[k].b1 << 48) | ((uint64_t)best[k].b2 << 32) | (uint64_t)idx;
        }
    }
}

// K9: merge candidate triples {d1:16, d2:16, idx:32} of `nshards` shards, gathered in ascending train-index order.
__global__ void top2_merge_kernel(const uint64_t* __restrict__ packed, int nshards, int nq, int32_t* idx1,
                                  uint16_t* d1, uint16_t* d2) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int b1 = 256, b2 = 256;
    uint32_t i1 = 0xFFFFFFFFu;
    for (int s = 0; s < nshards; ++s) {
        const uint64_t p = packed[(size_t)s * nq + qi];
        const int e1 = (int)(p >> 48), e2 = (int)((p >> 32) & 0xFFFF);
        if (e1 < b1) { b2 = min(b1, e2); b1 = e1; i1 = (uint32_t)p; }
        else b2 = min(b2, e1);
    }
    idx1[qi] = (int32_t)i1;
    d1[qi] = (uint16_t)b1;
    d2[qi] = (uint16_t)b2;
}

__global__ void pack_top2_kernel(const int32_t* idx1, const uint16_t* d1, const uint16_t* d2, int nq,
                                 uint64_t* packed) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    packed[qi] = ((uint64_t)d1[qi] << 48) | ((uint64_t)d2[qi] << 32) | (uint64_t)(uint32_t)idx1[qi];
}

// Chooses the number of train slices so that the grid is a whole number of waves of the 148 SMs.
int match_slices(int nq, int nt) {
    const int qBlocks = (nq + kMatchThreads * kQPT - 1) / (kMatchThreads * kQPT);
    const int maxSlices = (nt + kTileRows - 1) / kTileRows;
    if (maxSlices <= 1) return 1;
    int want = (148 * 8 + qBlocks - 1) / qBlocks;          // ~8 CTAs per SM
    if (want < 1) want = 1;
    if (want > maxSlices) want = maxSlices;
    if (want > 64) want = 64;
    return want;
}

void launch_hamming_top2_partial(const uint8_t* Q, int nq, const uint8_t* T, int nt, int tBase, int slices,
                                 uint64_t* partial, cudaStream_t s) {
    int sliceRows = (nt + slices - 1) / slices;
    sliceRows = (sliceRows + kTileRows - 1) / kTileRows * kTileRows;
    if (sliceRows < kTileRows) sliceRows = kTileRows;
    dim3 grid((nq + kMatchThreads * kQPT - 1) / (kMatchThreads * kQPT), slices);
    hamming_top2_kernel<<<grid, kMatchThreads, 0, s>>>(Q, nq, T, nt, sliceRows, tBase, partial);
}

void launch_top2_merge(const uint64_t* packed, int nshards, int nq, int32_t* idx1, uint16_t* d1, uint16_t* d2,
                       cudaStream_t s) {
    if (nq <= 0) return;
    top2_merge_kernel<<<(nq + 255) / 256, 256, 0, s>>>(packed, nshards, nq, idx1, d1, d2);
}

void launch_pack_top2(const int32_t* idx1, const uint16_t* d1, const uint16_t* d2, int nq, uint64_t* packed,
                      cudaStream_t s) {
    if (nq <= 0) return;
    pack_top2_kernel<<<(nq + 255) / 256, 256, 0, s>>>(idx1, d1, d2, nq, packed);
}

}  // namespace rumi
