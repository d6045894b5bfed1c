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
constexpr int kQPT = 4;                       // queries per thread
constexpr int kTileRows = 256;                // train descriptors per shared-memory tile (8 KB)
constexpr int kIdxBits = 21;                  // local train index inside one slice (slice <= 2^21 rows)

// Multiply-add that stays an IMAD (FMA pipe): the multiplier is a run-time value (blockDim.z == 1, which ptxas
// cannot fold), so these additions do not compete with XOR / LOP3 / min-max for the ALU pipe, which ncu showed
// at 88 % utilisation when they were plain IADD3s.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// popcount of the 256-bit XOR of two descriptors held as 8 words each
__device__ __forceinline__ uint32_t hamming256(const uint32_t q[8], const uint4 t0, const uint4 t1, uint32_t one,
                                               uint32_t two) {
    // The XOR words are materialised through opaque asm so that ptxas keeps the 8 XOR + 6 LOP3 form (14 ALU ops per
    // pair); left to itself it re-associates the tree over (q, t) directly and spends 20 LOP3 per pair.
    uint32_t x0, x1, x2, x3, x4, x5, x6, x7;
    asm("xor.b32 %0, %1, %2;" : "=r"(x0) : "r"(q[0]), "r"(t0.x));
    asm("xor.b32 %0, %1, %2;" : "=r"(x1) : "r"(q[1]), "r"(t0.y));
    asm("xor.b32 %0, %1, %2;" : "=r"(x2) : "r"(q[2]), "r"(t0.z));
    asm("xor.b32 %0, %1, %2;" : "=r"(x3) : "r"(q[3]), "r"(t0.w));
    asm("xor.b32 %0, %1, %2;" : "=r"(x4) : "r"(q[4]), "r"(t1.x));
    asm("xor.b32 %0, %1, %2;" : "=r"(x5) : "r"(q[5]), "r"(t1.y));
    asm("xor.b32 %0, %1, %2;" : "=r"(x6) : "r"(q[6]), "r"(t1.z));
    asm("xor.b32 %0, %1, %2;" : "=r"(x7) : "r"(q[7]), "r"(t1.w));
    // carry-save adders: (a,b,c) -> sum = a^b^c (LOP3 0x96), carry = majority(a,b,c) (LOP3 0xE8)
    uint32_t s1, c1, s2, c2, s3, c3;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s1) : "r"(x0), "r"(x1), "r"(x2));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c1) : "r"(x0), "r"(x1), "r"(x2));
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s2) : "r"(x3), "r"(x4), "r"(x5));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c2) : "r"(x3), "r"(x4), "r"(x5));
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s3) : "r"(s1), "r"(s2), "r"(x6));
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c3) : "r"(s1), "r"(s2), "r"(x6));
    // ones: s3, x7 ; twos: c1, c2, c3.  The additions run as IMADs (FMA pipe).
    const uint32_t ones = imad((uint32_t)__popc(s3), one, (uint32_t)__popc(x7));
    const uint32_t twos = imad((uint32_t)__popc(c1), one, imad((uint32_t)__popc(c2), one, (uint32_t)__popc(c3)));
    return imad(twos, two, ones);
}

// Per query the running best is kept as ONE key (distance << 21 | local train index): a single min keeps the
// smallest distance and, among equal distances, the earliest index; the second-best distance is
// min(b2, max(key, previous best)) in the same key domain.  Keys start at 256 << 21 (the reference's initial 256).
__global__ void __launch_bounds__(kMatchThreads)
hamming_top2_kernel(const uint8_t* __restrict__ Q, int nq, const uint8_t* __restrict__ T, int nt, int sliceRows,
                    int tBase, uint64_t* __restrict__ partial /* [gridDim.y][nq] */) {
    __shared__ uint4 tile[kTileRows * 2];
    const int q0 = blockIdx.x * (kMatchThreads * kQPT) + threadIdx.x;
    uint32_t q[kQPT][8];
    uint32_t k1[kQPT], k2[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int qi = q0 + k * kMatchThreads;
        const uint4* src = reinterpret_cast<const uint4*>(Q) + (size_t)min(qi, nq - 1) * 2;
        const uint4 a = src[0], b = src[1];
        q[k][0] = a.x; q[k][1] = a.y; q[k][2] = a.z; q[k][3] = a.w;
        q[k][4] = b.x; q[k][5] = b.y; q[k][6] = b.z; q[k][7] = b.w;
        k1[k] = k2[k] = 256u << kIdxBits;
    }
    const uint32_t one = blockDim.z, two = one + one, keyMul = one << kIdxBits;     // run-time constants (see imad)
    const int t0 = blockIdx.y * sliceRows, t1 = min(t0 + sliceRows, nt);
    for (int base = t0; base < t1; base += kTileRows) {
        const int rows = min(kTileRows, t1 - base);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 2; i += kMatchThreads)
            tile[i] = reinterpret_cast<const uint4*>(T)[(size_t)base * 2 + i];
        __syncthreads();
        const uint32_t local = (uint32_t)(base - t0);
#pragma unroll 2
        for (int r = 0; r < rows; ++r) {
            const uint4 a = tile[2 * r], b = tile[2 * r + 1];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) {
                const uint32_t key = imad(hamming256(q[k], a, b, one, two), keyMul, local + (uint32_t)r);
                k2[k] = min(k2[k], max(key, k1[k]));
                k1[k] = min(k1[k], key);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int qi = q0 + k * kMatchThreads;
        if (qi < nq) {
            const uint32_t d1 = k1[k] >> kIdxBits, d2 = k2[k] >> kIdxBits;
            const uint32_t idx = d1 >= 256u ? 0xFFFFFFFFu
                                            : (uint32_t)(t0 + tBase) + (k1[k] & ((1u << kIdxBits) - 1u));
            partial[(size_t)blockIdx.y * nq + qi] = ((uint64_t)d1 << 48) | ((uint64_t)d2 << 32) | (uint64_t)idx;
        }
    }
}

// K8-S: the same top-2 scan for MANY independent (query set, train set) pairs in one launch -- the descriptor
// association of matched key-frame pairs when two submaps are merged (one pair of key frames = one segment of
// ~1000 x ~1000 descriptors; R/lib_src/CloudMerging.cc:503-551 walks the same pairs).  grid = (query blocks, segment);
// one query per thread (segments are small: more CTAs matter more than register reuse), the segment's train rows
// stream through shared memory in ascending index, strict '<' => earliest index among ties; the train index written
// is relative to the segment (= feature index inside the second key frame).
constexpr int kSegThreads = 128;
__global__ void __launch_bounds__(kSegThreads)
hamming_top2_segments_kernel(const uint8_t* __restrict__ Q, const uint8_t* __restrict__ T, const PairSegment* __restrict__ segs,
                             int32_t* __restrict__ idx1, uint16_t* __restrict__ d1, uint16_t* __restrict__ d2) {
    __shared__ uint4 tile[kTileRows * 2];
    const PairSegment sg = segs[blockIdx.y];
    if ((int)(blockIdx.x * kSegThreads) >= sg.qCount) return;
    const int ql = blockIdx.x * kSegThreads + threadIdx.x;
    const bool live = ql < sg.qCount;
    uint32_t q[8];
    {
        const uint4* src = reinterpret_cast<const uint4*>(Q) + (size_t)(sg.qStart + min(ql, sg.qCount - 1)) * 2;
        const uint4 a = src[0], b = src[1];
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    uint32_t k1 = 256u << kIdxBits, k2 = k1;
    const uint32_t one = blockDim.z, two = one + one, keyMul = one << kIdxBits;
    for (int base = 0; base < sg.tCount; base += kTileRows) {
        const int rows = min(kTileRows, sg.tCount - base);
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 2; i += kSegThreads)
            tile[i] = reinterpret_cast<const uint4*>(T)[(size_t)(sg.tStart + base) * 2 + i];
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const uint32_t key = imad(hamming256(q, tile[2 * r], tile[2 * r + 1], one, two), keyMul, (uint32_t)(base + r));
            k2 = min(k2, max(key, k1));
            k1 = min(k1, key);
        }
    }
    if (live) {
        const uint32_t e1 = k1 >> kIdxBits, e2 = k2 >> kIdxBits;
        const int qi = sg.qStart + ql;
        idx1[qi] = e1 >= 256u ? -1 : (int32_t)(k1 & ((1u << kIdxBits) - 1u));
        d1[qi] = (uint16_t)e1;
        d2[qi] = (uint16_t)e2;
    }
}

// K8-C: candidate-list matching -- the form every projection / window search of ORBmatcher uses: query q is compared only
// with the train rows listed for it (Frame::GetFeaturesInArea, R/lib_src/Frame.cc:695-750).  One warp per query, lanes
// stride over its list: every entry's distance is written out (the adapters replay the call sites' order-dependent
// acceptance rules on them), plus the raw top-2 in list order (strict '<': the earliest entry wins ties).
__global__ void __launch_bounds__(256)
hamming_candidates_kernel(const uint8_t* __restrict__ Q, int nq, const uint8_t* __restrict__ T, const int32_t* __restrict__ off,
                          const int32_t* __restrict__ idx, uint16_t* __restrict__ dist, int32_t* __restrict__ idx1,
                          uint16_t* __restrict__ d1, int32_t* __restrict__ idx2, uint16_t* __restrict__ d2) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const uint4* qs = reinterpret_cast<const uint4*>(Q) + (size_t)q * 2;
    const uint4 qa = __ldg(qs), qb = __ldg(qs + 1);
    const int b = off[q], e = off[q + 1];
    uint32_t k1 = (256u << 16) | 0xFFFFu, k2 = k1;                   // key = distance << 16 | position in the list
    for (int p = b + lane; p < e; p += 32) {
        const uint4* ts = reinterpret_cast<const uint4*>(T) + (size_t)idx[p] * 2;
        const uint4 ta = __ldg(ts), tb = __ldg(ts + 1);
        const uint32_t d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                           __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
        dist[p] = (uint16_t)d;
        const uint32_t key = (d << 16) | (uint32_t)min(p - b, 0xFFFE);
        k2 = min(k2, max(key, k1));
        k1 = min(k1, key);
    }
    if (!idx1) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, k1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2, o);
        k2 = min(max(k1, o1), min(k2, o2));
        k1 = min(k1, o1);
    }
    if (lane == 0) {
        const uint32_t e1 = k1 >> 16, e2 = k2 >> 16;
        idx1[q] = e1 >= 256u ? -1 : idx[b + (int)(k1 & 0xFFFFu)];
        d1[q] = (uint16_t)e1;
        idx2[q] = e2 >= 256u ? -1 : idx[b + (int)(k2 & 0xFFFFu)];
        d2[q] = (uint16_t)e2;
    }
}

void launch_hamming_candidates(const uint8_t* Q, int nq, const uint8_t* T, const int32_t* off, const int32_t* idx,
                               uint16_t* dist, int32_t* idx1, uint16_t* d1, int32_t* idx2, uint16_t* d2, cudaStream_t s) {
    if (nq <= 0) return;
    hamming_candidates_kernel<<<(nq + 7) / 8, 256, 0, s>>>(Q, nq, T, off, idx, dist, idx1, d1, idx2, d2);
}

void launch_hamming_top2_segments(const uint8_t* Q, const uint8_t* T, const PairSegment* segs, int nseg, int maxQ,
                                  int32_t* idx1, uint16_t* d1, uint16_t* d2, cudaStream_t s) {
    if (nseg <= 0 || maxQ <= 0) return;
    dim3 grid((maxQ + kSegThreads - 1) / kSegThreads, nseg);
    hamming_top2_segments_kernel<<<grid, kSegThreads, 0, s>>>(Q, T, segs, idx1, d1, d2);
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

// Same fold, result kept in the packed candidate form: the local stage of a train-sharded match writes what the
// all-gather sends (no separate pack launch).
__global__ void top2_merge_packed_kernel(const uint64_t* __restrict__ packed, int nshards, int nq, uint64_t* __restrict__ out) {
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
    out[qi] = ((uint64_t)b1 << 48) | ((uint64_t)b2 << 32) | (uint64_t)i1;
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
    const int need = (int)(((long long)nt + (1 << kIdxBits) - 1) >> kIdxBits);   // local index must fit kIdxBits
    if (want < need) want = need;
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

void launch_top2_merge_packed(const uint64_t* packed, int nshards, int nq, uint64_t* out, cudaStream_t s) {
    if (nq <= 0) return;
    top2_merge_packed_kernel<<<(nq + 255) / 256, 256, 0, s>>>(packed, nshards, nq, out);
}

void launch_pack_top2(const int32_t* idx1, const uint16_t* d1, const uint16_t* d2, int nq, uint64_t* packed,
                      cudaStream_t s) {
    if (nq <= 0) return;
    pack_top2_kernel<<<(nq + 255) / 256, 256, 0, s>>>(idx1, d1, d2, nq, packed);
}

}  // namespace rumi
