// K8-U: all-pairs 256-bit Hamming top-2 on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// TMEM).  Same contract as K8 (match.cu): per query the best and second-best
// ORBmatcher::DescriptorDistance (R/lib_src/ORBmatcher.cc:1830-1844) over the train set, earliest index among ties.
//
// Hamming(q, t) = pop(q) + pop(t) - 2 <q, t> on descriptors expanded to one byte per bit, so the pair loop is an int8
// GEMM -- and the pop(t) term rides in the GEMM too: a ninth K step (32 more bytes per row: query side all 1, train side
// 32 signed bytes that sum to -pop(t)) makes the accumulator  acc = 2 <q,t> - pop(t) = pop(q) - Hamming(q, t)  directly
// comparable ACROSS the columns of a tile.  What is B200-specific here:
//   * one CTA per SM owns 256 queries at a time and keeps their expanded operand IN TMEM (2 x 72 columns, written once per
//     segment with tcgen05.st): every tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 64, K = 32; A unsigned from TMEM, B
//     signed from shared memory) reads only the train operand from shared memory.  With both operands in shared memory the
//     MMAs read 144 KB + the ring refill wrote 36 KB per tile = 115 of the 128 B/clk the shared memory delivers, and the
//     tensor pipe waited for operands (76 % busy); with A in TMEM it is 72 + 36 KB.  A 128-row train tile is issued as two
//     64-column sub-tiles, each into its own accumulator set (2 query halves x 64 columns of int32), so the epilogue of
//     one sub-tile overlaps the MMAs of the other; accumulators never touch registers until tcgen05.ld reads them;
//   * the train set is packed ONCE per call (umma_pack_train_kernel) into ready-to-load operand tiles: per 128 rows 36 KB
//     in the canonical no-swizzle K-major shared-memory layout ([16-byte K chunk][row][16 B]: core matrices of 8 rows x
//     16 B, SBO = 128 B, LBO = 2048 B).  8 bits become 8 operand bytes with ONE 64-bit multiply (see expand_row).  The
//     query block is expanded the same way inside the kernel, once per CTA;
//   * warp roles: a LOADER warp (one thread streams the tiles into a 4-stage ring with cp.async.bulk + mbarrier
//     complete_tx), an ISSUER warp (converged, one ELECTED lane issues the MMAs with every operand in a uniform register:
//     18 UTCIMMA + 2 UTCBAR in ~150 straight-line uniform-datapath instructions per tile.  tcgen05.mma blocks its issuing
//     thread while the tensor queue is full, so the issuer must not be a thread anybody else waits for; and it has only
//     ~85 clk per MMA, so instruction count matters -- see the note at the issuer loop) and 16 EPILOGUE warps.  Roles meet only at
//     mbarriers: full[4] (stage landed), bar[2] (tcgen05.commit: accumulator set ready), accFree[2] (accumulator set
//     drained), slotFree[4] (tcgen05.commit: the MMAs that read the stage are done).
// Epilogue: the raw accumulators of a 32-column tcgen05.ld group go through ONE max tree (VIMNMX3, ~0.5 instructions per
// pair) and one warp vote against the accumulator value that would tie the current second best; only the 8-column
// sub-groups that pass build keys ((256 - acc) << 22 | train index, one IMAD each) and run the exact update
// k2 = min(k2, max(key, k1)), k1 = min(k1, key).  (Round 1 built the key of EVERY pair first -- 1 IMAD + 1 constant word per
// pair, ALU pipe 58 %, tensor pipe 54 % busy -- because pop(t) differed per column.)  pop(q) is added at the end (it does
// not change the order).  A query row is scanned by two threads (column halves of every tile); keys are unique, so their
// merge is an exact min / max.
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>

namespace rumi {

namespace {

constexpr int kUmBM = 256, kUmBN = 128;            // queries per CTA (two M = 128 MMAs share every train tile), train rows per tile
constexpr int kUmWorkers = 2 * kUmBM;              // 16 worker warps: (query half) x (TMEM lane quarter) x (column half of the tile)
constexpr int kUmThreads = kUmWorkers + 64;        // + one warp that issues the MMAs + one warp that issues the bulk copies
constexpr int kUmStages = 4;                       // train-tile ring
constexpr int kUmK = 256 + 32;                     // operand bytes per row: 256 descriptor bits + the 32 bytes of the pop(t) K step
constexpr int kUmIdxBits = 22;
constexpr uint32_t kUmIdxMask = (1u << kUmIdxBits) - 1u;
constexpr int kUmTileBytes = 128 * kUmK;           // one expanded 128-row operand tile: 36 KB
constexpr int kUmStageBytes = kUmTileBytes;        // one bulk copy per tile
constexpr int kUmChunkStride = 128 * 16;           // bytes between consecutive 16-byte K chunks (LBO)
constexpr int kUmSubN = 64;                        // train rows per MMA: a tile is issued as two 64-column halves
constexpr int kUmTmemCols = 512;
// TMEM map (columns): the QUERY operand of both halves lives in TMEM for the whole segment -- [0, 72) rows 0-127,
// [72, 144) rows 128-255 (288 operand bytes = 72 columns, byte k of a row = K index k) -- so an MMA reads only the train
// operand from shared memory; accumulators at 256 + 128 * (sub-tile) + 64 * (query half), 64 columns each.
constexpr uint32_t kUmTmemA = 0, kUmTmemAcc = 256;
constexpr uint32_t kUmACols = kUmK / 4;

struct UmmaSmem {
    alignas(1024) uint8_t B[kUmStages][kUmStageBytes];   // ring of pre-expanded train tiles, filled by cp.async.bulk
    uint32_t mergeK[2][kUmBM];                     // (k1, k2) of the threads that scanned the upper columns of every sub-tile
    alignas(8) uint64_t bar[2];                    // MMAs of sub-tile j of the current tile complete: accumulator set j full
    alignas(8) uint64_t full[kUmStages];           // stage landed (cp.async.bulk complete_tx)
    alignas(8) uint64_t slotFree[kUmStages];       // stage consumed: the MMAs that read it are done (tcgen05.commit)
    alignas(8) uint64_t accFree[2];                // accumulator set drained by all worker warps
    uint32_t tmemBase;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle K-major shared-memory matrix descriptor (start address, LBO, SBO in 16-byte units; version 1 = sm_100),
// as two words so that the address arithmetic of the issuer stays 32 bit: lo = start | LBO << 16, hi = SBO | 1 << 14
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((uint32_t)(kUmChunkStride >> 4) << 16); }
constexpr uint32_t kUmDescHi = (uint32_t)(128 >> 4) | (1u << 14);

// instruction descriptor: D = S32 (bits 4-5 = 2), A = unsigned 8 bit (bits 7-9 = 0), B = SIGNED 8 bit (bits 10-12 = 1),
// both K-major, N >> 3 at bit 17, M >> 4 at 24
constexpr uint32_t kUmIdesc = (2u << 4) | (1u << 10) | ((uint32_t)(kUmSubN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// D[tmem] (+)= A[tmem] * B[smem]: M = 128 query rows (TMEM lanes) x N = 64 train rows x K = 32 bytes
__device__ __forceinline__ void umma_i8_ts(uint32_t dTmem, uint32_t aTmem, uint32_t bDescLo, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], bd, %4, p;\n\t}\n"
        :: "r"(dTmem), "r"(aTmem), "r"(bDescLo), "r"(kUmDescHi), "r"(kUmIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// A warp-uniform value the compiler cannot prove uniform (loaded from shared memory, result of a division) -> REDUX
// writes a UNIFORM register.  Every lane of a converged warp must pass the same x.
__device__ __forceinline__ uint32_t uniform(uint32_t x) { return __reduce_or_sync(0xFFFFFFFFu, x); }
__device__ __forceinline__ int uniform(int x) { return (int)__reduce_or_sync(0xFFFFFFFFu, (uint32_t)x); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t umad(uint32_t a, uint32_t b, uint32_t c) {      // stays an IMAD (FMA pipe)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Expands words [w0, w0 + NW) of row `row` of the packed descriptor array D (n rows) into the operand tile at `tile`
// (row r of the tile) and returns the popcount of the WHOLE row.  Output byte k = bit k of the descriptor; rows beyond
// n are zero.  4 bits -> 4 bytes by one multiply: (nibble * 0x00204081) & 0x01010101.
struct PackedRow { uint4 a, b; };
__device__ __forceinline__ PackedRow load_packed(const uint8_t* __restrict__ D, int n, int row) {
    PackedRow p;
    p.a = make_uint4(0, 0, 0, 0); p.b = p.a;
    if (row < n) {
        const uint4* src = reinterpret_cast<const uint4*>(D) + (size_t)row * 2;
        p.a = src[0]; p.b = src[1];
    }
    return p;
}

// Expands words [w0, w0 + NW) of a packed row into the operand tile (row r).  One byte of 8 bits becomes 8 output
// bytes with ONE 64-bit multiply: x * 0x8040201008040201 puts bit i of x at positions i + 9 j (all distinct, so no
// carries); position 8 j + 7 holds bit 7 - j.  The order of the 256 products inside a dot product is free, so the
// bit-reversed byte order is used as is on both sides.  Train side: bytes 0 / 1; query side: bytes 0 / 2, so that the
// accumulator holds 2 <q,t>.
template <int NW, bool kQuery>
__device__ __forceinline__ int expand_row(const PackedRow& p, uint8_t* tile, int r, int w0) {
    const uint32_t w[8] = {p.a.x, p.a.y, p.a.z, p.a.w, p.b.x, p.b.y, p.b.z, p.b.w};
    int pop = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) pop += __popc(w[i]);
    constexpr int kShift = kQuery ? 6 : 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < w0 || i >= w0 + NW) continue;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                       // 16 bits -> one 16-byte chunk
            const unsigned long long lo = (unsigned long long)__byte_perm(w[i], 0, 0x4440 + 2 * hf) * 0x8040201008040201ull;
            const unsigned long long hi = (unsigned long long)__byte_perm(w[i], 0, 0x4441 + 2 * hf) * 0x8040201008040201ull;
            uint4 o;
            o.x = ((uint32_t)lo & 0x80808080u) >> kShift; o.y = ((uint32_t)(lo >> 32) & 0x80808080u) >> kShift;
            o.z = ((uint32_t)hi & 0x80808080u) >> kShift; o.w = ((uint32_t)(hi >> 32) & 0x80808080u) >> kShift;
            *reinterpret_cast<uint4*>(tile + (size_t)(2 * i + hf) * kUmChunkStride + (size_t)r * 16) = o;
        }
    }
    return pop;
}

// The two K chunks (32 bytes) of the pop(t) step.  Query side: all ones.  Train side: 32 signed bytes that sum to -pop
// (byte e = -floor((pop + e) / 32), each in [-8, 0]); a padding row (beyond the train set) gets pop = 256 with all
// descriptor bits zero: accumulator -256 = "distance 256 + pop(q)", never a match.
__device__ __forceinline__ void write_pop_chunks(uint8_t* tile, int r, int pop, bool query) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t wds[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0x01010101u;
            if (!query) {
                v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) v |= (uint32_t)((0 - ((pop + 16 * c + 4 * k + b) >> 5)) & 0xFF) << (8 * b);
            }
            wds[k] = v;
        }
        *reinterpret_cast<uint4*>(tile + (size_t)(16 + c) * kUmChunkStride + (size_t)r * 16) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
}

// Query side: words [w0, w0 + 4) of a packed row expanded (bytes 0 / 2, same byte order as expand_row: the K index of a
// byte is its offset in the row on both sides) and stored to 32 consecutive TMEM columns of this thread's lane.
// tcgen05.st.32x32b: lane i of the warp writes TMEM lane 32 * (warp % 4) + i.
__device__ __forceinline__ void store_query_tmem(const PackedRow& p, int w0, uint32_t taddr) {
    const uint32_t w[8] = {p.a.x, p.a.y, p.a.z, p.a.w, p.b.x, p.b.y, p.b.z, p.b.w};
    uint32_t o[32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t x = w0 == 0 ? w[i] : w[4 + i];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const unsigned long long e = (unsigned long long)__byte_perm(x, 0, 0x4440 + b) * 0x8040201008040201ull;
            o[8 * i + 2 * b] = ((uint32_t)e & 0x80808080u) >> 6;
            o[8 * i + 2 * b + 1] = ((uint32_t)(e >> 32) & 0x80808080u) >> 6;
        }
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]),
           "r"(o[8]), "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15]),
           "r"(o[16]), "r"(o[17]), "r"(o[18]), "r"(o[19]), "r"(o[20]), "r"(o[21]), "r"(o[22]), "r"(o[23]),
           "r"(o[24]), "r"(o[25]), "r"(o[26]), "r"(o[27]), "r"(o[28]), "r"(o[29]), "r"(o[30]), "r"(o[31]) : "memory");
}

// Pre-pass (once per call): the train set as ready-to-load tiles -- per 128 rows 36 KB in the shared-memory operand layout.
__global__ void __launch_bounds__(128)
umma_pack_train_kernel(const uint8_t* __restrict__ T, int nt, uint8_t* __restrict__ tiles) {
    const int tile = blockIdx.x, r = threadIdx.x, row = tile * kUmBN + r;
    uint8_t* dst = tiles + (size_t)tile * kUmStageBytes;
    const PackedRow p = load_packed(T, nt, row);
    const int pop = expand_row<8, false>(p, dst, r, 0);
    write_pop_chunks(dst, r, row < nt ? pop : 256, false);
}

// Work decomposition: the (query block, train tile) grid is flattened query-block-major and cut into gridDim.x equal
// ranges of `unitsPerCta` tiles -- ONE persistent CTA per SM, no wave quantisation and one pipeline fill per CTA instead of
// one per (query block, slice) (round 1: 942 CTAs in 6.4 waves for 40 000 x 40 000, ~12 us of fixed cost per wave).
// A range that crosses a query-block boundary is processed as two (rarely three) SEGMENTS: the query operand is
// re-expanded, the running (k1, k2) are flushed.  Segment results go to partial[slot][query] with slot = CTA index minus
// the first CTA that touches the query block: slots of a block are ascending train ranges, the merge folds them in order.
__global__ void __launch_bounds__(kUmThreads, 1)
hamming_top2_umma_kernel(const uint8_t* __restrict__ Q, int nq, const uint8_t* __restrict__ tiles, int nt, int unitsPerCta,
                         int tBase, uint64_t* __restrict__ partial /* [slots][nq] */) {
    extern __shared__ __align__(1024) uint8_t smemRaw[];
    UmmaSmem& sm = *reinterpret_cast<UmmaSmem*>(smemRaw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r128 = tid & 127, quarter = (tid >> 7) & 3;               // row inside a 128-row tile / quarter of the work
    const int half = quarter & 1, colHalf = quarter >> 1;               // epilogue: query half, columns [64 * colHalf, +64)
    const int qrowLocal = half * 128 + r128;                            // query row of this thread inside its query block
    const int nTilesAll = (nt + kUmBN - 1) / kUmBN;
    const int nQB = (nq + kUmBM - 1) / kUmBM;
    const long long total = (long long)nQB * nTilesAll;
    const long long u0 = (long long)blockIdx.x * unitsPerCta, u1 = min(u0 + unitsPerCta, total);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&sm.tmemBase)), "r"((uint32_t)kUmTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.bar[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&sm.accFree[b])), "r"(kUmWorkers / 32) : "memory");
        }
        for (int b = 0; b < kUmStages; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.full[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.slotFree[b])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                               // TMEM base and barriers visible to every role
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmemBase;
    const uint32_t bar0 = smem_u32(&sm.bar[0]), bar1 = smem_u32(&sm.bar[1]);
    const uint32_t fullBase = smem_u32(&sm.full[0]), slotBase = smem_u32(&sm.slotFree[0]);
    const uint32_t free0 = smem_u32(&sm.accFree[0]), free1 = smem_u32(&sm.accFree[1]);
    const uint32_t bBase0 = smem_u32(sm.B[0]);

    int done = 0;                                  // tiles this CTA has pushed through the ring so far (all roles count alike)
    for (long long u = u0; u < u1;) {
        const int qb = (int)(u / nTilesAll), tile0 = (int)(u - (long long)qb * nTilesAll);
        const int ntiles = (int)min((long long)(nTilesAll - tile0), u1 - u);
        const int q0 = qb * kUmBM;
        // ---- the query block of this segment, expanded by the worker threads (quarter rows) into the operand layout.
        //      (Every MMA of the previous segment is complete: its accumulators were all read.)
        int popq = 0;
        if (tid < kUmWorkers) {
            const PackedRow qrow = load_packed(Q, nq, q0 + qrowLocal);
            popq = __popc(qrow.a.x) + __popc(qrow.a.y) + __popc(qrow.a.z) + __popc(qrow.a.w) +
                   __popc(qrow.b.x) + __popc(qrow.b.y) + __popc(qrow.b.z) + __popc(qrow.b.w);
            const uint32_t arow = tmem + kUmTmemA + (uint32_t)half * kUmACols + ((uint32_t)((warp & 3) * 32) << 16);
            store_query_tmem(qrow, 4 * colHalf, arow + 32u * (uint32_t)colHalf);
            if (colHalf == 0) {                    // the pop(t) K step: query side all ones
                const uint32_t one = 0x01010101u;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};"
                             :: "r"(arow + 64u), "r"(one) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;

        if (warp == kUmWorkers / 32 + 1) {
            // =========================== loader: one thread feeds the train-tile ring with bulk copies ===========================
            if (lane == 0) {
                const uint8_t* src = tiles + (size_t)tile0 * kUmStageBytes;
                for (int i = 0; i < ntiles; ++i) {
                    const int it = done + i;
                    const uint32_t s = (uint32_t)it & (kUmStages - 1);
                    if (it >= kUmStages) mbar_wait(slotBase + 8u * s, (uint32_t)((it / kUmStages) - 1) & 1u);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                 :: "r"(fullBase + 8u * s), "r"((uint32_t)kUmStageBytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(bBase0 + s * kUmStageBytes), "l"(src + (size_t)i * kUmStageBytes), "r"((uint32_t)kUmStageBytes),
                                    "r"(fullBase + 8u * s) : "memory");
                }
            }
        } else if (warp == kUmWorkers / 32) {
            // =========================== MMA issuer: one thread, never touches data ===========================
            // tcgen05.mma blocks its issuing thread while the tensor pipe's queue is full (measured: ~1800 clk per tile when
            // the issuer was also an epilogue thread -- the whole CTA then waited for it at the next barrier), so the
            // issuer is a warp of its own and talks to the other roles through mbarriers only.
            // The WHOLE warp walks the loop and one elected lane issues (the elect.sync form ptxas recognises): with the
            // loop inside `if (lane == 0)` every UTCIMMA was wrapped in a 6-instruction elect/branch loop and -- once ring
            // position and trip count came out of a 64-bit division, i.e. vector registers -- in R2UR.BROADCAST waterfalls
            // too: ~17 instructions per MMA on a thread that has ~85 clk per MMA, which made the ISSUER the limiter of long
            // scans.  uniform() moves the three run-time values into uniform registers; everything else is constants.
            const int itBase = uniform(done), nIssue = uniform(ntiles);
            const uint32_t tmemU = uniform(tmem);
            for (int i = 0; i < nIssue; ++i) {
                const int it = itBase + i;
                const uint32_t s = (uint32_t)it & (kUmStages - 1);
                mbar_wait(fullBase + 8u * s, (uint32_t)(it / kUmStages) & 1u);                    // stage landed
#pragma unroll
                for (int j = 0; j < 2; ++j) {                      // sub-tile j (train rows 64 j .. 64 j + 63) -> accumulator set j
                    if (it >= 1) mbar_wait(j ? free1 : free0, (uint32_t)(it - 1) & 1u);           // set j drained (tile it - 1)
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        // rows of a K chunk are 16 B apart: sub-tile j starts 64 * 16 B into every chunk (descriptor units: 16 B)
                        const uint32_t bd0 = umma_desc_lo(bBase0 + s * kUmStageBytes) + (uint32_t)j * (kUmSubN * 16 >> 4);
                        const uint32_t d = tmemU + kUmTmemAcc + (uint32_t)j * 128u;
                        const uint32_t a = tmemU + kUmTmemA;
#pragma unroll
                        for (int k = 0; k < kUmK / 32; ++k) {      // 8 K steps of descriptor bits + the pop(t) step
                            const uint32_t bd = bd0 + (uint32_t)k * (2 * kUmChunkStride >> 4);
                            umma_i8_ts(d, a + 8u * k, bd, k > 0);
                            umma_i8_ts(d + 64u, a + kUmACols + 8u * k, bd, k > 0);
                        }
                        // arrives when the MMAs above are complete: accumulator set j is full
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     :: "r"(j ? bar1 : bar0) : "memory");
                        if (j == 1)                                 // ... and the stage may be refilled
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                         :: "r"(slotBase + 8u * s) : "memory");
                    }
                    __syncwarp();
                }
            }
        } else {
            // =========================== workers: the top-2 epilogue ===========================
            const uint32_t negOne = 0u - (1u << kUmIdxBits);               // key = (256 - acc) << 22 | index = acc * negOne + (256 << 22 | index)
            int thr = 256 - (int)(k2 >> kUmIdxBits);                        // an accumulator >= thr can change (k1, k2); ties included
            // 32 raw accumulators at a time (one tcgen05.ld group): one max tree, one warp vote; keys only for the 8-column
            // sub-groups that can matter.  rowBase = train row (inside this call) of column 0 of the group.
            auto scan32 = [&](const uint32_t* v, uint32_t rowBase) {
                int g[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int* a8 = reinterpret_cast<const int*>(v) + 8 * q;
                    g[q] = max(max(max(max(a8[0], a8[1]), a8[2]), max(max(a8[3], a8[4]), a8[5])), max(a8[6], a8[7]));
                }
                const int m = max(max(g[0], g[1]), max(g[2], g[3]));
                if (__any_sync(0xFFFFFFFFu, m >= thr)) {           // warp-uniform: no divergence bookkeeping
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (__any_sync(0xFFFFFFFFu, g[q] >= thr)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {           // idempotent for lanes whose keys do not qualify
                                const uint32_t key = umad(v[8 * q + e], negOne, (256u << kUmIdxBits) | (rowBase + 8u * q + e));
                                k2 = min(k2, max(key, k1));
                                k1 = min(k1, key);
                            }
                        }
                    }
                    thr = 256 - (int)(k2 >> kUmIdxBits);
                }
            };
#define RUMI_LDTM32(v, addr)                                                                                          \
    asm volatile(                                                                                                     \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                     \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),             \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),       \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),     \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])      \
        : "r"(addr) : "memory")

            // this thread = query row (TMEM lane 32 * (warp % 4) + lane of its half's accumulators) x 32 of the 64 columns
            // (train rows) of every sub-tile; the scan of sub-tile j runs while the MMAs of the next one do
            const uint32_t tacc = tmem + kUmTmemAcc + (uint32_t)half * 64u + (uint32_t)colHalf * 32u + ((uint32_t)((warp & 3) * 32) << 16);
            for (int i = 0; i < ntiles; ++i) {
                const int it = done + i;
                const uint32_t rowBase = (uint32_t)(tile0 + i) * kUmBN + 32u * (uint32_t)colHalf;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mbar_wait(j ? bar1 : bar0, (uint32_t)it & 1u);             // accumulator set j complete
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t v[32];
                    RUMI_LDTM32(v, tacc + (uint32_t)j * 128u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(j ? free1 : free0);            // the issuer may overwrite this accumulator set
                    scan32(v, rowBase + 64u * (uint32_t)j);
                }
            }
#undef RUMI_LDTM32
            // the two threads of a query row scanned disjoint train rows (unique keys): park one half for the exact merge
            if (colHalf == 1) { sm.mergeK[0][qrowLocal] = k1; sm.mergeK[1][qrowLocal] = k2; }
        }
        done += ntiles;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                           // every role is through this segment; the parked halves are visible
        const int qi = q0 + qrowLocal;
        if (tid < kUmBM) {                         // colHalf == 0 workers: min / max merge with the parked half
            const uint32_t o1 = sm.mergeK[0][qrowLocal], o2 = sm.mergeK[1][qrowLocal];
            k2 = min(min(k2, o2), max(k1, o1));
            k1 = min(k1, o1);
            if (qi < nq) {
                // key >> 22 = 256 - acc = pop(t) - 2 dot + 256; distance = that - 256 + pop(q).  A distance of 256 is "no match"
                // (the reference's scan starts from bestDist = 256 with strict '<').
                int d1 = k1 == 0xFFFFFFFFu ? 256 : (int)(k1 >> kUmIdxBits) - 256 + popq;
                int d2 = k2 == 0xFFFFFFFFu ? 256 : (int)(k2 >> kUmIdxBits) - 256 + popq;
                d1 = min(d1, 256); d2 = min(d2, 256);
                const uint32_t idx = d1 >= 256 ? 0xFFFFFFFFu : (uint32_t)tBase + (k1 & kUmIdxMask);
                const int slot = (int)blockIdx.x - (int)(((long long)qb * nTilesAll) / unitsPerCta);
                partial[(size_t)slot * nq + qi] = ((uint64_t)d1 << 48) | ((uint64_t)d2 << 32) | (uint64_t)idx;
            }
        }
        u += ntiles;
        // (the next segment's expansion overwrites A and mergeK only after the barrier at its start; k1 / k2 restart)
        __syncthreads();
    }
    // every accumulator read is complete (wait::ld above); release TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kUmTmemCols) : "memory");
    }
}

// Slots of partial[] that no CTA writes (a query block touched by fewer CTAs than the maximum) must read as "no candidate".
__global__ void umma_fill_partial_kernel(uint64_t* partial, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) partial[i] = (256ull << 48) | (256ull << 32) | 0xFFFFFFFFull;
}

}  // namespace

namespace {
// ranges of the flattened (query block, tile) grid: one persistent CTA per SM, at least 8 tiles each
void umma_plan(int nq, int nt, int* nCta, int* unitsPerCta, int* slots) {
    const long long nQB = (nq + kUmBM - 1) / kUmBM, nTiles = (nt + kUmBN - 1) / kUmBN, total = nQB * nTiles;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (const char* e = getenv("RUMI_UMMA_CTAS")) sms = std::max(1, atoi(e));          // A/B runs
    long long ctas = std::min<long long>(sms, std::max<long long>(1, total / 8));
    long long U = (total + ctas - 1) / ctas;
    // A train set that does not fit the L2 (36 KB per tile) must be streamed by all CTAs IN STEP, so that a tile is fetched
    // from HBM once and the other CTAs hit it in L2 (measured: staggered ranges made 10^6 train rows 16 % slower, HBM
    // bound): every CTA then takes whole query blocks and starts at tile 0.
    if (nTiles * kUmStageBytes > (48ll << 20) && nQB >= sms) U = ((nQB + sms - 1) / sms) * nTiles;
    ctas = (total + U - 1) / U;
    *nCta = (int)ctas; *unitsPerCta = (int)U;
    // most CTAs that touch one query block: its nTiles units can straddle floor((nTiles - 1) / U) + 2 ranges
    *slots = (int)std::min<long long>(ctas, (nTiles - 1) / U + 2);
}
}  // namespace

int umma_slices(int nq, int nt) {
    int nCta, U, slots;
    umma_plan(nq, nt, &nCta, &U, &slots);
    return slots;
}

size_t umma_train_bytes(int nt) { return (size_t)((nt + kUmBN - 1) / kUmBN) * kUmStageBytes; }

void launch_hamming_top2_umma(const uint8_t* Q, int nq, const uint8_t* T, int nt, uint8_t* trainTiles, int tBase, int slices,
                              uint64_t* partial, cudaStream_t s) {
    // (per device and cheap: set on every launch rather than cached in a process-wide flag)
    cudaFuncSetAttribute(hamming_top2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UmmaSmem) + 1024);
    const int nTiles = (nt + kUmBN - 1) / kUmBN;
    int nCta, U, slots;
    umma_plan(nq, nt, &nCta, &U, &slots);
    umma_pack_train_kernel<<<nTiles, 128, 0, s>>>(T, nt, trainTiles);
    const size_t nPartial = (size_t)slices * nq;
    umma_fill_partial_kernel<<<(unsigned)((nPartial + 255) / 256), 256, 0, s>>>(partial, nPartial);
    hamming_top2_umma_kernel<<<nCta, kUmThreads, sizeof(UmmaSmem) + 1024, s>>>(Q, nq, trainTiles, nt, U, tBase, partial);
}

}  // namespace rumi
