// K8-U: all-pairs 256-bit Hamming top-2 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// TMEM).  Same contract as K8 / K8-T (match.cu, match_imma.cu): per query the best and second-best
// ORBmatcher::DescriptorDistance (R/lib_src/ORBmatcher.cc:1830-1844) over the train set, earliest index among ties.
//
// Hamming(q, t) = pop(q) + pop(t) - 2 <q, t> on descriptors expanded to 256 bytes of 0/1, so the pair loop is an
// int8 GEMM.  What is B200-specific here:
//   * the MMA is ONE tcgen05.mma.cta_group::1.kind::i8 (M = 128 queries, N = 128 train rows, K = 32) issued by one
//     thread, 8 per tile; the 128 x 128 int32 accumulator tile lives in TMEM (two of them, ping-pong: 256 of the 512
//     columns, so two CTAs share an SM) and never touches registers until the top-2 epilogue reads it with tcgen05.ld;
//   * operands are expanded from the packed 32-byte descriptors INSIDE the kernel, straight into the canonical
//     no-swizzle K-major shared-memory layout ([16-byte K chunk][row][16 B]: core matrices of 8 rows x 16 B, SBO = 128 B,
//     LBO = 2048 B) -- no expanded copy in HBM / L2 (the mma.sync kernel streams 256 B per descriptor from L2, this one
//     32 B), one IMAD + one LOP per 4 output bytes;
//   * completion is tracked with tcgen05.commit -> mbarrier; the tensor pipe works on tile i+1 while the CUDA cores
//     run the top-2 update of tile i and expand tile i+2.
// Key per pair = ((pop(t) - 2 <q,t> + 256) << 22 | train index) built by one IMAD from the accumulator, then
// k2 = min(k2, max(key, k1)), k1 = min(k1, key); pop(q) is added at the end (it does not change the order).
#include "kernels.cuh"

namespace rumi {

namespace {

constexpr int kUmBM = 128, kUmBN = 128;            // queries per CTA, train rows per tile
constexpr int kUmIdxBits = 22;
constexpr uint32_t kUmIdxMask = (1u << kUmIdxBits) - 1u;
constexpr int kUmTileBytes = kUmBN * 256;          // one expanded operand tile: 32 KB
constexpr int kUmChunkStride = kUmBN * 16;         // bytes between consecutive 16-byte K chunks (LBO)
constexpr int kUmTmemCols = 256;                   // two 128-column accumulators

struct UmmaSmem {
    alignas(1024) uint8_t A[kUmTileBytes];
    alignas(1024) uint8_t B[2][kUmTileBytes];
    uint32_t cst[3][kUmBN];                        // per train row of a tile: ((pop + 256) << 22) | index
    alignas(8) uint64_t bar[2];                    // MMA-complete barriers of the two accumulators
    uint32_t tmemBase;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle K-major shared-memory matrix descriptor (start address, LBO, SBO in 16-byte units; version 1 = sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(kUmChunkStride >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}

// instruction descriptor: D = S32 (bits 4-5 = 2), A / B = unsigned 8 bit (0), both K-major, N >> 3 at bit 17, M >> 4 at 24
constexpr uint32_t kUmIdesc = (2u << 4) | ((uint32_t)(kUmBN >> 3) << 17) | ((uint32_t)(kUmBM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(kUmIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(bar), "r"(parity) : "memory");
}

// Expands row `row` of the packed descriptor array D (n rows) into the operand tile at `tile` (this thread's row r)
// and returns its popcount.  Output byte k of the row = bit k of the descriptor; rows beyond n are zero.
__device__ __forceinline__ int expand_row(const uint8_t* __restrict__ D, int n, int row, uint8_t* tile, int r) {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row < n) {
        const uint4* src = reinterpret_cast<const uint4*>(D) + (size_t)row * 2;
        const uint4 a = src[0], b = src[1];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
    int pop = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pop += __popc(w[i]);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                       // 16 bits -> one 16-byte chunk
            uint4 o;
            const uint32_t v = w[i] >> (16 * hf);
            o.x = (((v >> 0) & 0xFu) * 0x00204081u) & 0x01010101u;
            o.y = (((v >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
            o.z = (((v >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
            o.w = (((v >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint4*>(tile + (size_t)(2 * i + hf) * kUmChunkStride + (size_t)r * 16) = o;
        }
    }
    return pop;
}

__global__ void __launch_bounds__(128, 2)
hamming_top2_umma_kernel(const uint8_t* __restrict__ Q, int nq, const uint8_t* __restrict__ T, int nt, int tilesPerSlice,
                         int tBase, uint64_t* __restrict__ partial /* [gridDim.y][nq] */) {
    extern __shared__ __align__(1024) uint8_t smemRaw[];
    UmmaSmem& sm = *reinterpret_cast<UmmaSmem*>(smemRaw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int q0 = blockIdx.x * kUmBM;
    const int nTilesAll = (nt + kUmBN - 1) / kUmBN;
    const int tile0 = blockIdx.y * tilesPerSlice, tile1 = min(tile0 + tilesPerSlice, nTilesAll);
    const int ntiles = tile1 - tile0;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&sm.tmemBase)), "r"((uint32_t)kUmTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int popq = expand_row(Q, nq, q0 + tid, sm.A, tid);
    if (ntiles > 0) {
        const int row = tile0 * kUmBN + tid;
        const int pop = expand_row(T, nt, row, sm.B[0], tid);
        sm.cst[0][tid] = row < nt ? ((uint32_t)(pop + 256) << kUmIdxBits) | (uint32_t)row : 0xFFFFFFFFu;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmemBase;
    const uint32_t aBase = smem_u32(sm.A), bBase0 = smem_u32(sm.B[0]), bBase1 = smem_u32(sm.B[1]);
    const uint32_t bar0 = smem_u32(&sm.bar[0]), bar1 = smem_u32(&sm.bar[1]);

    auto issue = [&](int i) {                                           // tile i of this slice -> accumulator i & 1
        const uint32_t bBase = (i & 1) ? bBase1 : bBase0;
        const uint32_t d = tmem + (uint32_t)(i & 1) * kUmBN;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            umma_i8(d, umma_desc(aBase + k * 2 * kUmChunkStride), umma_desc(bBase + k * 2 * kUmChunkStride), k > 0);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     :: "r"((i & 1) ? bar1 : bar0) : "memory");
    };
    if (ntiles > 0 && tid == 0) issue(0);
    __syncwarp();

    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    for (int i = 0; i < ntiles; ++i) {
        if (i + 1 < ntiles) {                                           // expand the next tile while MMA i runs
            const int row = (tile0 + i + 1) * kUmBN + tid;
            const int pop = expand_row(T, nt, row, sm.B[(i + 1) & 1], tid);
            sm.cst[(i + 1) % 3][tid] = row < nt ? ((uint32_t)(pop + 256) << kUmIdxBits) | (uint32_t)row : 0xFFFFFFFFu;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        mbar_wait((i & 1) ? bar1 : bar0, (uint32_t)(i >> 1) & 1u);       // accumulator i & 1 complete
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                // next tile expanded, epilogue i-1 finished
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (i + 1 < ntiles && tid == 0) issue(i + 1);
        __syncwarp();
        // ---- top-2 update of tile i: this thread = query row (TMEM lane), 128 columns = train rows of the tile ----
        const uint32_t* cst = sm.cst[i % 3];
        const uint32_t taddr = tmem + (uint32_t)(i & 1) * kUmBN + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < kUmBN; c0 += 32) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + (uint32_t)c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const uint4 c = *reinterpret_cast<const uint4*>(cst + c0 + j);       // broadcast
                const uint32_t cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t key = cc[e] - (v[j + e] << (kUmIdxBits + 1));     // - 2 <q,t> in the distance field
                    k2 = min(k2, max(key, k1));
                    k1 = min(k1, key);
                }
            }
        }
    }
    // every accumulator read is complete (wait::ld above); release TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kUmTmemCols) : "memory");
    }
    const int qi = q0 + tid;
    if (qi < nq) {
        // key >> 22 = pop(t) - 2 dot + 256; distance = that - 256 + pop(q).  A distance of 256 is "no match"
        // (the reference's scan starts from bestDist = 256 with strict '<').
        int d1 = k1 == 0xFFFFFFFFu ? 256 : (int)(k1 >> kUmIdxBits) - 256 + popq;
        int d2 = k2 == 0xFFFFFFFFu ? 256 : (int)(k2 >> kUmIdxBits) - 256 + popq;
        d1 = min(d1, 256); d2 = min(d2, 256);
        const uint32_t idx = d1 >= 256 ? 0xFFFFFFFFu : (uint32_t)tBase + (k1 & kUmIdxMask);
        partial[(size_t)blockIdx.y * nq + qi] = ((uint64_t)d1 << 48) | ((uint64_t)d2 << 32) | (uint64_t)idx;
    }
}

}  // namespace

int umma_slices(int nq, int nt) {
    const int qBlocks = (nq + kUmBM - 1) / kUmBM;
    const int nTiles = (nt + kUmBN - 1) / kUmBN;
    const int maxSlices = std::max(1, nTiles / 8);                      // at least 8 tiles per slice
    int want = (148 * 2 * 4 + qBlocks - 1) / qBlocks;                    // >= 4 waves of two CTAs per SM
    want = std::min(std::min(want, maxSlices), 64);
    return std::max(want, 1);
}

void launch_hamming_top2_umma(const uint8_t* Q, int nq, const uint8_t* T, int nt, int tBase, int slices,
                              uint64_t* partial, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(hamming_top2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UmmaSmem) + 1024);
        configured = true;
    }
    const int nTiles = (nt + kUmBN - 1) / kUmBN;
    const int tilesPerSlice = (nTiles + slices - 1) / slices;
    dim3 grid((nq + kUmBM - 1) / kUmBM, slices);
    hamming_top2_umma_kernel<<<grid, 128, sizeof(UmmaSmem) + 1024, s>>>(Q, nq, T, nt, tilesPerSlice, tBase, partial);
}

}  // namespace rumi
