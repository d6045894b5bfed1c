// Bag-of-words side of the matching path:
//   K10  DBoW2 vocabulary tree descent -- TemplatedVocabulary::transform,
//        R/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1258 (distance = FORB::distance, FORB.cpp:81-101);
//   K11  per-node all-pairs Hamming blocks for ORBmatcher::SearchByBoW (R/lib_src/ORBmatcher.cc:198-370): the
//        reference's scan has a sequential "already matched" dependency (:249), so the device returns every distance
//        of every common vocabulary node and the host replays the acceptance over those small matrices.
#include "kernels.cuh"

namespace rumi {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

__device__ __forceinline__ int hamming256(const uint32_t a[8], const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// Half a warp per feature: lane c of the half takes child c of the current node (k <= 16 in one step, larger fan-outs
// in steps of 16), a 4-step shuffle minimum over (distance << 8 | child position) picks the first child with the
// smallest distance -- the reference keeps the first because it only replaces on '<' (:1243-1247).
__global__ void __launch_bounds__(256) bow_transform_kernel(const BowTreeView t, const uint8_t* __restrict__ desc, int n,
                                                            int levelsup, int32_t* __restrict__ word,
                                                            double* __restrict__ weight, int32_t* __restrict__ node) {
    const int half = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, sub = threadIdx.x & 15;
    const bool live = half < n;
    const int f = live ? half : n - 1;
    uint32_t a[8];
    {
        const uint4* p = reinterpret_cast<const uint4*>(desc + 32 * (size_t)f);
        const uint4 lo = __ldg(p), hi = __ldg(p + 1);
        a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    }
    const int nidLevel = t.L - levelsup;
    // the two half-warps descend independent features and may leave the loop at different depths (unbalanced trees):
    // the shuffles name only the lanes of this half
    const unsigned halfMask = 0xFFFFu << (threadIdx.x & 16);
    int cur = 0, level = 0, nid = 0;
    int nchild = t.childCount[0];
    while (nchild > 0 && level < 64) {
        ++level;
        const int base = t.childStart[cur];
        uint32_t best = 0xFFFFFFFFu;
        for (int c0 = 0; c0 < nchild; c0 += 16) {
            const int c = c0 + sub;
            uint32_t key = 0xFFFFFFFFu;
            if (c < nchild) {
                const int id = t.childIds[base + c];
                const uint4* p = reinterpret_cast<const uint4*>(t.desc + 32 * (size_t)id);
                key = ((uint32_t)hamming256(a, __ldg(p), __ldg(p + 1)) << 16) | (uint32_t)c;
            }
            best = min(best, key);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(halfMask, best, o));
        cur = t.childIds[base + (int)(best & 0xFFFFu)];
        if (level == nidLevel) nid = cur;
        nchild = t.childCount[cur];
    }
    if (live && sub == 0) {
        word[f] = t.wordId[cur];
        weight[f] = t.weight[cur];
        node[f] = nid;
    }
}

void launch_bow_transform(const BowTreeView& t, const uint8_t* desc, int n, int levelsup, int32_t* word, double* weight,
                          int32_t* node, cudaStream_t s) {
    if (n <= 0) return;
    const int threads = 256, perBlock = threads / 16;
    bow_transform_kernel<<<(n + perBlock - 1) / perBlock, threads, 0, s>>>(t, desc, n, levelsup, word, weight, node);
}

// One warp per common vocabulary node: dist[out + i * bc + j] = Hamming(A[aIdx[as + i]], B[bIdx[bs + j]]).
__global__ void __launch_bounds__(128) bow_node_distances_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                                                 const int32_t* __restrict__ aIdx,
                                                                 const int32_t* __restrict__ bIdx,
                                                                 const BowSegment* __restrict__ segs, int nseg,
                                                                 uint16_t* __restrict__ dist) {
    const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (seg >= nseg) return;
    const BowSegment sg = segs[seg];
    const int npair = sg.aCount * sg.bCount;
    for (int p = lane; p < npair; p += 32) {
        const int i = p / sg.bCount, j = p - i * sg.bCount;
        const uint4* pa = reinterpret_cast<const uint4*>(A + 32 * (size_t)aIdx[sg.aStart + i]);
        const uint4* pb = reinterpret_cast<const uint4*>(B + 32 * (size_t)bIdx[sg.bStart + j]);
        const uint4 a0 = __ldg(pa), a1 = __ldg(pa + 1);
        const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        dist[sg.outOff + p] = (uint16_t)hamming256(a, __ldg(pb), __ldg(pb + 1));
    }
}

void launch_bow_node_distances(const uint8_t* A, const uint8_t* B, const int32_t* aIdx, const int32_t* bIdx,
                               const BowSegment* segs, int nseg, uint16_t* dist, cudaStream_t s) {
    if (nseg <= 0) return;
    const int threads = 128, perBlock = threads / 32;
    bow_node_distances_kernel<<<(nseg + perBlock - 1) / perBlock, threads, 0, s>>>(A, B, aIdx, bIdx, segs, nseg, dist);
}


// K12: MapPoint::ComputeDistinctiveDescriptors (R/lib_src/MapPoint.cc:355-426), batched: one warp per map point.
// For every observed descriptor i the lanes hold the distances to the other descriptors (lane j <-> descriptor j, in
// chunks of 32); the median vDists[0.5 * (N - 1)] of the sorted row is found without sorting, by a 9-step binary search
// on the value (distances are 0..256): the smallest t with #(d <= t) >= m + 1.  The first row with the smallest median
// wins (strict '<', :412).
__global__ void __launch_bounds__(128) distinctive_kernel(const uint8_t* __restrict__ desc,
                                                          const int32_t* __restrict__ offsets, int npoints,
                                                          int32_t* __restrict__ bestIdx, int32_t* __restrict__ bestMedian) {
    const int pt = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (pt >= npoints) return;
    const int base = offsets[pt], N = offsets[pt + 1] - base;
    if (N <= 0) {
        if (lane == 0) { bestIdx[pt] = -1; bestMedian[pt] = -1; }
        return;
    }
    const int m = (int)(0.5 * (double)(N - 1));                   // index of the median in the sorted row (:407)
    int best = 0x7FFFFFFF, bestI = 0;
    for (int i = 0; i < N; ++i) {
        const uint4* pi = reinterpret_cast<const uint4*>(desc + 32 * (size_t)(base + i));
        const uint4 a0 = __ldg(pi), a1 = __ldg(pi + 1);
        const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        int lo = 0, hi = 256;                                     // smallest t in [0, 256] with count(d <= t) >= m + 1
        while (lo < hi) {
            const int t = (lo + hi) >> 1;
            int cnt = 0;
            for (int j0 = 0; j0 < N; j0 += 32) {
                const int j = j0 + lane;
                bool le = false;
                if (j < N) {
                    const uint4* pj = reinterpret_cast<const uint4*>(desc + 32 * (size_t)(base + j));
                    le = hamming256(a, __ldg(pj), __ldg(pj + 1)) <= t;     // Distances[i][i] = 0 falls out of the formula
                }
                cnt += __popc(__ballot_sync(kFullMask, le));
            }
            if (cnt >= m + 1) hi = t; else lo = t + 1;
        }
        if (lo < best) { best = lo; bestI = i; }
    }
    if (lane == 0) { bestIdx[pt] = bestI; bestMedian[pt] = best; }
}

void launch_distinctive(const uint8_t* desc, const int32_t* offsets, int npoints, int32_t* bestIdx, int32_t* bestMedian,
                        cudaStream_t s) {
    if (npoints <= 0) return;
    const int threads = 128, perBlock = threads / 32;
    distinctive_kernel<<<(npoints + perBlock - 1) / perBlock, threads, 0, s>>>(desc, offsets, npoints, bestIdx, bestMedian);
}

}  // namespace rumi
