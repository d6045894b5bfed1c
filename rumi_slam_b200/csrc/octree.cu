// K3: DistributeOctTree (R/lib_src/ORBextractor.cc:538-724) -- one thread block per (frame, level).
// The algorithm itself lives in octree_core.cuh (shared with the host test harness); this file gathers the
// FAST candidates in the reference's insertion order (cell-row-major, raster inside a cell, :800-806), carves the
// work buffers out of shared memory (or a global scratch for pathological candidate counts) and runs it.
#include "kernels.cuh"
#include "octree_core.cuh"

namespace rumi {

constexpr int kOctThreads = 256;

struct BlockCtx {
    int tid, nthr;
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ int atomic_add(int* p, int v) { return atomicAdd(p, v); }
};

struct OctreeSmemLayout {
    size_t keys, lkeys, glo, crlo, crcnt, pend, next, next2, meta, qbase, hist, part, scal, cellScan, total;
    int nodeCap, createCap, pendCap;
};

__host__ __device__ inline OctreeSmemLayout octree_layout(int smemKeys, int maxNodeCap, int nthreads) {
    OctreeSmemLayout L;
    L.nodeCap = maxNodeCap;
    L.createCap = 3 * maxNodeCap + 16;
    L.pendCap = 2 * maxNodeCap + 16;
    int gpad = 1;
    while (gpad < maxNodeCap) gpad <<= 1;
    size_t o = 0;
    L.keys = o; o += 8ull * smemKeys;
    L.lkeys = o; o += 8ull * gpad;
    L.pend = o; o += 8ull * L.pendCap;
    L.glo = o; o += 4ull * (maxNodeCap + 1);
    L.crlo = o; o += 4ull * L.createCap;
    L.crcnt = o; o += 4ull * L.createCap;
    L.next = o; o += 4ull * L.pendCap;
    L.next2 = o; o += 4ull * L.pendCap;
    L.meta = o; o += 4ull * L.pendCap;
    L.qbase = o; o += 4ull * L.pendCap;
    L.hist = o; o += 4ull * 2 * (kMaxTreeDepth + 2);
    L.part = o; o += 4ull * (nthreads + 1);
    L.scal = o; o += 4ull * SC_COUNT;
    L.cellScan = o; o += 4ull * (nthreads + 1);
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

size_t octree_smem_bytes(int smemKeys, int maxNodeCap, int nthreads) {
    return octree_layout(smemKeys, maxNodeCap, nthreads).total;
}

__global__ void __launch_bounds__(kOctThreads) octree_kernel(const __grid_constant__ OctreeArgs a,
                                                             const __grid_constant__ OrbConst oc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int f = blockIdx.x, l = blockIdx.y;
    const LevelGeom& g = oc.lv[l];
    const OctreeSmemLayout L = octree_layout(a.smemKeys, a.maxNodeCap, kOctThreads);
    BlockCtx ctx{(int)threadIdx.x, kOctThreads};

    const int M = a.levelCount[(long long)f * oc.nlevels + l];
    const uint32_t* cand = a.cand + a.candLevelOff[l] + (long long)f * g.candCap;
    uint32_t* ordered = a.candOrdered + a.candLevelOff[l] + (long long)f * g.candCap;

    // ---- gather candidates in the reference's cell order: exclusive scan of the per-cell counts ----
    {
        int* scan = reinterpret_cast<int*>(smem + L.cellScan);
        const int ncell = g.nCols * g.nRows;
        const int* cc = a.cellCount + (long long)f * oc.totalCells + g.cellBase;
        const int* co = a.cellOff + (long long)f * oc.totalCells + g.cellBase;
        const int chunk = (ncell + kOctThreads - 1) / kOctThreads;
        const int c0 = threadIdx.x * chunk, c1 = min(c0 + chunk, ncell);
        int s = 0;
        for (int c = c0; c < c1; ++c) s += cc[c];
        scan[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int t = 0; t < kOctThreads; ++t) { const int v = scan[t]; scan[t] = run; run += v; }
        }
        __syncthreads();
        int dst = scan[threadIdx.x];
        for (int c = c0; c < c1; ++c) {
            const int n = cc[c], o = co[c];
            for (int k = 0; k < n; ++k) ordered[dst + k] = cand[o + k];
            dst += n;
        }
        __syncthreads();
    }

    OctreeWork w;
    const int Mpad = next_pow2(M > 1 ? M : 2);
    w.keys = Mpad <= a.smemKeys ? reinterpret_cast<uint64_t*>(smem + L.keys)
                                : a.bigKeys + a.bigKeysLevelOff[l] + (long long)f * a.bigKeysCap[l];
    w.lkeys = reinterpret_cast<uint64_t*>(smem + L.lkeys);
    w.pend = reinterpret_cast<uint64_t*>(smem + L.pend);
    w.glo = reinterpret_cast<uint32_t*>(smem + L.glo);
    w.cr_lo = reinterpret_cast<uint32_t*>(smem + L.crlo);
    w.cr_cnt = reinterpret_cast<uint32_t*>(smem + L.crcnt);
    w.next = reinterpret_cast<uint32_t*>(smem + L.next);
    w.next2 = reinterpret_cast<uint32_t*>(smem + L.next2);
    w.meta = reinterpret_cast<uint32_t*>(smem + L.meta);
    w.qbase = reinterpret_cast<uint32_t*>(smem + L.qbase);
    w.hist = reinterpret_cast<int*>(smem + L.hist);
    w.part = reinterpret_cast<int*>(smem + L.part);
    w.scal = reinterpret_cast<int*>(smem + L.scal);
    w.nodeCap = L.nodeCap; w.createCap = L.createCap; w.pendCap = L.pendCap;

    const int slots = (l + 1 < oc.nlevels ? oc.lv[l + 1].kpBase : oc.kpCap) - g.kpBase;
    uint32_t* out = a.sel + (long long)f * oc.kpCap + g.kpBase;
    if (threadIdx.x == 0) w.scal[SC_NOUT] = 0;
    __syncthreads();
    distribute_quadtree(ctx, ordered, M, g.quota, g, w, out, slots);
    __syncthreads();
    if (threadIdx.x == 0) a.selCount[(long long)f * oc.nlevels + l] = min(w.scal[SC_NOUT], slots);
}

void launch_octree(const OctreeArgs& a, const OrbConst& oc, cudaStream_t s) {
    const size_t smem = octree_smem_bytes(a.smemKeys, a.maxNodeCap, kOctThreads);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    octree_kernel<<<dim3(a.nframes, oc.nlevels), kOctThreads, smem, s>>>(a, oc);
}

}  // namespace rumi
