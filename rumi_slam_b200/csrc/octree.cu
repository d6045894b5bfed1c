// K3: DistributeOctTree (R/lib_src/ORBextractor.cc:538-724) -- one thread block per (frame, level).
// The algorithm itself lives in octree_core.cuh (shared with the host test harness); this file gathers the
// FAST candidates in the reference's insertion order (cell-row-major, raster inside a cell, :800-806), carves the
// work buffers out of shared memory (or a global scratch for pathological candidate counts) and runs it.
#include <cstdlib>

#include <algorithm>

#include "kernels.cuh"
#include "slots.cuh"
#include "octree_core.cuh"

namespace rumi {

constexpr int kOctThreads = 256;        // pass 0 of a chunk: many CTAs per SM hide each other's latencies
constexpr int kOctThreadsBig = 1024;    // calls of a few frames and pass 1: one CTA per SM, so the CTA brings its own warps

// Block-wide ascending sort of n u64 keys (n a power of two, E <= n <= E * blockDim.x) held E per thread in REGISTERS:
// a bitonic network whose stages with partner distance j < E are pure register compare-exchanges, E <= j < 32 E are
// warp shuffles, and only j >= 32 E go through shared memory with block barriers (6 of the 66 stages for 2048 keys).
// A lone warp walking a shared-memory network pays a dependent ~500-cycle round trip per stage; here the E
// exchanges of a stage are independent instructions.
template <int E>
__device__ void block_sort_regs(uint64_t* a, int n, int tid) {
    const int T = n / E;                                   // threads that hold data
    const bool active = tid < T;
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = active ? a[tid * E + r] : ~0ull;
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        for (; j >= 32 * E; j >>= 1) {                     // partner in another warp: through shared memory
            __syncthreads();
            if (active) {
#pragma unroll
                for (int r = 0; r < E; ++r) a[tid * E + r] = v[r];
            }
            __syncthreads();
            if (active) {
                const int pt = tid ^ (j / E);
                const bool lower = (tid & (j / E)) == 0;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const uint64_t o = a[pt * E + r];
                    const bool up = (((tid * E + r) & k) == 0);
                    const bool keepMin = lower == up;
                    v[r] = ((v[r] < o) == keepMin) ? v[r] : o;
                }
            }
        }
        for (; j >= E; j >>= 1) {                          // partner in another lane: shuffles (all lanes take part)
            const int lm = j / E;
            const bool lower = (tid & lm) == 0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, v[r], lm);
                const bool up = (((tid * E + r) & k) == 0);
                const bool keepMin = lower == up;
                v[r] = ((v[r] < o) == keepMin) ? v[r] : o;
            }
        }
#pragma unroll
        for (int js = E / 2; js > 0; js >>= 1) {           // partner in this thread: registers
            if (js <= (k >> 1)) {
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    if ((r & js) == 0) {
                        const bool up = (((tid * E + r) & k) == 0);
                        const uint64_t x = v[r], y = v[r | js];
                        const bool sw = (x > y) == up;
                        v[r] = sw ? y : x;
                        v[r | js] = sw ? x : y;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int r = 0; r < E; ++r) a[tid * E + r] = v[r];
    }
    __syncthreads();
}

// Exclusive scan of the bucket counts in start[0 .. nb) (nb <= 4 * nthr): per-thread chunk + warp scans + one pass over the
// warp sums; start[nb] = n.  Sets fill[nb] when a bucket holds more than 192 keys.  Ends with a barrier.
__device__ __forceinline__ void bucket_scan(int* start, int* fill, int nb, int n, int tid, int nthr) {
    const int per = (nb + nthr - 1) / nthr;
    const int b0 = min(tid * per, nb), b1 = min(b0 + per, nb);
    int s = 0, mx = 0;
    for (int b = b0; b < b1; ++b) { s += start[b]; mx = max(mx, start[b]); }
    if (mx > 192) fill[nb] = 1;                                      // crowded bucket: quadratic ranking would be slow
    const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ int wsum[33];
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int t = lane < nw ? wsum[lane] : 0;
        int ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xFFFFFFFFu, ti, o);
            if (lane >= o) ti += u;
        }
        if (lane < nw) wsum[lane] = ti - t;
    }
    __syncthreads();
    int run = wsum[wid] + incl - s;
    for (int b = b0; b < b1; ++b) { const int c = start[b]; start[b] = run; run += c; }
    if (tid == nthr - 1) start[nb] = n;
    __syncthreads();
}

// Block-wide ascending sort of the first n (<= 16 * blockDim.x) u64 keys by BUCKETS of their leading bits: bucket(key) =
// key >> bshift < nb.  Tree keys are (path code | insertion index | response): the code of a candidate is unique (every
// leaf of the full-depth tree is one pixel), the leading 2p digits of the code spread the keys over nb = roots * 4^p
// buckets of a handful of keys each.  Histogram (shared atomics) -> scan -> index lists per bucket -> every key counts
// the smaller keys of its own bucket -> final position.  ~M * (bucket size) shared reads instead of the ~M log^2 M
// compare-exchanges of a bitonic network (29 k -> 4 k cycles for 1 500 keys).  Returns false (nothing moved) when a
// bucket is too crowded (clustered candidates) -- the caller falls back to the bitonic network.
// scratch: ints [2 * nb + 2] then u16 [n].
template <int E>
__device__ bool block_bucket_sort(uint64_t* a, int n, int bshift, int nb, int* scratch, int tid, int nthr) {
    int* start = scratch;                 // [nb + 1]
    int* fill = scratch + nb + 1;         // [nb] (+1: crowded flag)
    uint16_t* idx = reinterpret_cast<uint16_t*>(fill + nb + 1);
    for (int b = tid; b <= nb; b += nthr) { start[b] = 0; fill[b] = 0; }
    __syncthreads();
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid + r * nthr;
        v[r] = i < n ? a[i] : ~0ull;
        if (i < n) atomicAdd(&start[(int)(v[r] >> bshift)], 1);
    }
    __syncthreads();
    bucket_scan(start, fill, nb, n, tid, nthr);
    if (fill[nb]) return false;                                        // uniform: read after the barrier
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid + r * nthr;
        if (i < n) {
            const int b = (int)(v[r] >> bshift);
            idx[start[b] + atomicAdd(&fill[b], 1)] = (uint16_t)i;
        }
    }
    __syncthreads();
    int rank[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid + r * nthr;
        rank[r] = 0;
        if (i < n) {
            const int b = (int)(v[r] >> bshift);
            const int p0 = start[b], p1 = start[b + 1];
            int c = p0;
            for (int p = p0; p < p1; ++p) {
                const int j = idx[p];
                const uint64_t o = a[j];
                c += (o < v[r]) || (o == v[r] && j < i);
            }
            rank[r] = c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const int i = tid + r * nthr;
        if (i < n) a[rank[r]] = v[r];
    }
    __syncthreads();
    return true;
}

// Loop form of the bucket sort for more keys than a thread holds in registers (dense frames: thousands of candidates on
// a level).  a[] is the shared-memory key buffer; the keys are scattered bucket by bucket into `bounce` (global memory,
// >= n u64, L2-resident), copied back, ranked inside their now contiguous bucket and bounced once more into their final
// places: four coalesced-or-scattered passes over 8 n bytes of L2 plus ~n * (bucket size) shared reads, against the
// ~n log^2 n shared compare-exchanges of the bitonic network (442 k -> 30 k cycles for 10 700 keys).  scratch: ints [2 * nb + 2].
__device__ bool block_bucket_sort_loop(uint64_t* a, int n, int bshift, int nb, int* scratch, uint64_t* bounce, int tid,
                                       int nthr) {
    int* start = scratch;                 // [nb + 1]
    int* fill = scratch + nb + 1;         // [nb] (+1: crowded flag)
    for (int b = tid; b <= nb; b += nthr) { start[b] = 0; fill[b] = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) atomicAdd(&start[(int)(a[i] >> bshift)], 1);
    __syncthreads();
    bucket_scan(start, fill, nb, n, tid, nthr);
    if (fill[nb]) return false;                                        // uniform: read after the barrier
    for (int i = tid; i < n; i += nthr) {
        const uint64_t k = a[i];
        const int b = (int)(k >> bshift);
        bounce[start[b] + atomicAdd(&fill[b], 1)] = k;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) a[i] = __ldcg(bounce + i);
    __syncthreads();
    for (int i = tid; i < n; i += nthr) {
        const uint64_t k = a[i];
        const int b = (int)(k >> bshift);
        const int p0 = start[b], p1 = start[b + 1];
        int c = p0;
        for (int p = p0; p < p1; ++p) c += a[p] < k;                   // distinct keys
        bounce[c] = k;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) a[i] = __ldcg(bounce + i);
    __syncthreads();
    return true;
}

// Ascending sort of n <= 2 * blockDim.x u64 keys by counting, per key, the smaller ones (distinct keys): two passes
// over shared memory instead of a sorting network (the leaf-list keys: a few hundred).
__device__ void block_rank_sort(uint64_t* a, int n, int tid, int nthr) {
    uint64_t v[2];
    int rank[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = tid + r * nthr;
        v[r] = i < n ? a[i] : ~0ull;
        rank[r] = 0;
    }
    if (tid < n) {                                        // threads without a key only join the barriers
#pragma unroll 8
        for (int j = 0; j < n; ++j) {                     // n is a power of two >= 8: the loads of 8 keys are in flight
            const uint64_t o = a[j];
#pragma unroll
            for (int r = 0; r < 2; ++r) rank[r] += (o < v[r]) || (o == v[r] && j < tid + r * nthr);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r)
        if (tid + r * nthr < n) a[rank[r]] = v[r];
    __syncthreads();
}

struct BlockCtx {
    int tid, nthr;
    long long* clk;          // optional phase clock accumulators (test / profiling hook), else nullptr
    long long last;
    int* sortScratch;        // shared memory that is idle while the tree keys are sorted
    int sortScratchBytes;
    uint64_t* bounce;        // global scratch of the loop-form bucket sort (>= the level's candidate capacity), or nullptr
    // adds the cycles since the previous mark to accumulator `id` (thread 0 only)
    __device__ __forceinline__ void mark(int id) {
        if (clk && tid == 0) { const long long now = clock64(); clk[id] += now - last; last = now; }
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void sync_warp() { __syncwarp(); }
    // ascending sort of n (power of two >= 8) u64 keys; everything written before the call is visible (callers sync)
    __device__ __forceinline__ void sort_u64(uint64_t* a, int n) {
        if (n <= 256) block_rank_sort(a, n, tid, nthr);   // leaf lists: one pass over n keys per thread
        else if (n <= 8 * nthr) block_sort_regs<8>(a, n, tid);
        else if (n <= 16 * nthr) block_sort_regs<16>(a, n, tid);
        else bitonic_sort_u64(*this, a, n);               // global-memory scratch for pathological candidate counts
    }
    // ascending sort of the tree keys: a[0 .. m) are real keys (bucket = key >> bshift < nb), a[m .. npad) are ~0
    __device__ __forceinline__ void sort_tree_keys(uint64_t* a, int m, int npad, int bshift, int nb) {
        const bool fits = nb <= 4 * nthr && (size_t)sortScratchBytes >= 4 * (size_t)(2 * nb + 2) + 2 * (size_t)m;
        if (fits && m > 64 && m <= nthr) { if (block_bucket_sort<1>(a, m, bshift, nb, sortScratch, tid, nthr)) return; }
        else if (fits && m > nthr && m <= 2 * nthr) { if (block_bucket_sort<2>(a, m, bshift, nb, sortScratch, tid, nthr)) return; }
        else if (fits && m > 2 * nthr && m <= 4 * nthr) { if (block_bucket_sort<4>(a, m, bshift, nb, sortScratch, tid, nthr)) return; }
        else if (fits && m > 4 * nthr && m <= 8 * nthr) { if (block_bucket_sort<8>(a, m, bshift, nb, sortScratch, tid, nthr)) return; }
        else if (m > 8 * nthr && bounce && nb <= 4 * nthr && (size_t)sortScratchBytes >= 4 * (size_t)(2 * nb + 2)) {
            if (block_bucket_sort_loop(a, m, bshift, nb, sortScratch, bounce, tid, nthr)) return;
        }
        sort_u64(a, npad);                                 // crowded buckets, keys in global memory or a tiny problem
    }
    __device__ __forceinline__ int warp_id() const { return tid >> 5; }
    __device__ __forceinline__ int num_warps() const { return nthr >> 5; }
    __device__ __forceinline__ int lane() const { return tid & 31; }
    __device__ __forceinline__ int warp_size() const { return 32; }
    __device__ __forceinline__ int atomic_add(int* p, int v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_min(int* p, int v) { atomicMin(p, v); }
    // exclusive prefix sum of v over the block in thread order (+ block total); tmp holds >= nthr / 32 + 1 ints.
    // Ends with a barrier, so everything written before the call is visible after it.
    __device__ __forceinline__ int exclusive_scan(int v, int* tmp, int* total) {
        const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += n;
        }
        __syncthreads();                                   // tmp may still be read from a previous scan
        if (lane == 31) tmp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < nw ? tmp[lane] : 0;
            int ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xFFFFFFFFu, ti, o);
                if (lane >= o) ti += n;
            }
            if (lane < nw) tmp[lane] = ti - t;
            if (lane == nw - 1) tmp[nw] = ti;
        }
        __syncthreads();
        *total = tmp[nw];
        return tmp[wid] + incl - v;
    }
};

struct OctreeSmemLayout {
    size_t keys, lkeys, glo, crlo, crcnt, pend, lsort, next, next2, meta, qbase, hist, part, scal, cellScan, total;
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
    L.lsort = o; o += 8ull * L.pendCap;
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

// One (frame, level) problem, by a whole CTA.  smemKeys = keys that fit the shared-memory key buffer of THIS launch.
__device__ void octree_problem(const OctreeArgs& a, const OrbConst& oc, uint8_t* smem, int smemKeys, int f, int l) {
    const LevelGeom& g = oc.lv[l];
    const int nthr = blockDim.x;
    const OctreeSmemLayout L = octree_layout(smemKeys, a.maxNodeCap, nthr);
    // while the tree keys are sorted everything between the key buffer and the histograms is idle
    BlockCtx ctx{(int)threadIdx.x, nthr, (a.dbgClk && f == 0) ? a.dbgClk + 16 * l : nullptr, clock64(),
                 reinterpret_cast<int*>(smem + L.lkeys), (int)(L.hist - L.lkeys), nullptr};

    const int M = a.levelCount[(long long)f * oc.nlevels + l];
    const uint32_t* cand = a.cand + a.candLevelOff[l] + (long long)f * g.candCap;
    uint32_t* ordered = a.candOrdered + a.candLevelOff[l] + (long long)f * g.candCap;

    // ---- gather candidates in the reference's cell order (cell-row-major, raster inside a cell, :800-806) ----
    // exclusive scan of the per-cell counts -> first output index of every cell (kept in the not-yet-used key buffer),
    // then one thread per OUTPUT element: binary search for its cell, independent loads batched four at a time.
    {
        int* scan = reinterpret_cast<int*>(smem + L.cellScan);
        const int ncell = g.nCols * g.nRows;
        const int* cc = a.cellCount + (long long)f * oc.totalCells + g.cellBase;
        const int* co = a.cellOff + (long long)f * oc.totalCells + g.cellBase;
        int* cellDst = ncell + 1 <= 2 * smemKeys
                           ? reinterpret_cast<int*>(smem + L.keys)
                           : reinterpret_cast<int*>(a.bigKeys + a.bigKeysLevelOff[l] + (long long)f * a.bigKeysCap[l]);
        const int chunk = (ncell + nthr - 1) / nthr;
        const int c0 = min((int)threadIdx.x * chunk, ncell), c1 = min(c0 + chunk, ncell);
        int s = 0;
        for (int c = c0; c < c1; ++c) s += cc[c];
        int total;
        int dst = ctx.exclusive_scan(s, scan, &total);
        for (int c = c0; c < c1; ++c) { cellDst[c] = dst; dst += cc[c]; }
        if (threadIdx.x == 0) cellDst[ncell] = M;
        __syncthreads();
        constexpr int U = 4;
        for (int i0 = threadIdx.x; i0 < M; i0 += U * nthr) {
            uint32_t v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * nthr;
                if (i < M) {
                    int lo = 0, hi = ncell;                 // largest c with cellDst[c] <= i (empty cells share a start)
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (cellDst[mid] <= i) lo = mid; else hi = mid;
                    }
                    v[u] = cand[co[lo] + (i - cellDst[lo])];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * nthr;
                if (i < M) ordered[i] = v[u];
            }
        }
        __syncthreads();
    }

    OctreeWork w;
    const int Mpad = next_pow2(M > 8 ? M : 8);
    uint64_t* const gkeys = a.bigKeys + a.bigKeysLevelOff[l] + (long long)f * a.bigKeysCap[l];
    w.keys = Mpad <= smemKeys ? reinterpret_cast<uint64_t*>(smem + L.keys) : gkeys;
    if (Mpad <= smemKeys) ctx.bounce = gkeys;          // keys in shared memory: the global buffer is free as sort scratch
    w.lkeys = reinterpret_cast<uint64_t*>(smem + L.lkeys);
    w.pend = reinterpret_cast<uint64_t*>(smem + L.pend);
    w.lsort = reinterpret_cast<uint64_t*>(smem + L.lsort);
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
    ctx.mark(10);
    distribute_quadtree(ctx, ordered, M, g.quota, g, w, out, slots);
    ctx.mark(11);
    if (ctx.clk && threadIdx.x == 0) { ctx.clk[12] = M; ctx.clk[13] = w.scal[SC_NOUT]; ctx.clk[14] += w.scal[SC_NREPLAY]; }
    __syncthreads();
    if (threadIdx.x == 0) a.selCount[(long long)f * oc.nlevels + l] = min(w.scal[SC_NOUT], slots);
    if (!a.fuseSlots) return;
    // ---- the last level CTA of this frame assigns the output slots (every CTA publishes its selection first) ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.frameDone + f, 1) == oc.nlevels - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    assign_slots_frame(a.slots, oc, f, threadIdx.x);
}


// Pass 0: one CTA per (frame, level).  A level with more candidates than the shared-memory key buffer holds (2048: denser
// than the benchmark frames, common on real textured images) is not sorted in global memory here -- that path costs 5x
// per key -- but appended to a work list for pass 1.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) octree_kernel(const __grid_constant__ OctreeArgs a,
                                                          const __grid_constant__ OrbConst oc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int f = blockIdx.x, l = blockIdx.y + a.levelFirst;
    const int M = a.levelCount[(long long)f * oc.nlevels + l];
    if (next_pow2(M > 8 ? M : 8) > a.smemKeys) {
        if (a.bigList) {
            if (threadIdx.x == 0) a.bigList[atomicAdd(a.bigCount, 1)] = (f << 8) | l;
            return;
        }
        if (threadIdx.x == 0) *(volatile int*)a.denseFlag = 1;       // tells the host to run two passes from now on
    }
    octree_problem(a, oc, smem, a.smemKeys, f, l);
}

// Pass 1: a few persistent CTAs with a large key buffer (16384 keys = 128 KB of shared memory) walk the work list; on sparse
// frames the list is empty and the launch costs a couple of microseconds.
__global__ void __launch_bounds__(kOctThreadsBig) octree_big_kernel(const __grid_constant__ OctreeArgs a,
                                                                    const __grid_constant__ OrbConst oc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = *a.bigCount;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int item = a.bigList[i];
        octree_problem(a, oc, smem, a.smemKeysBig, item >> 8, item & 0xFF);
        __syncthreads();                                   // the next problem reuses the shared memory
    }
}

void launch_octree(const OctreeArgs& a, const OrbConst& oc, cudaStream_t s) {
    const bool wide = a.threads > kOctThreads;
    const int nthr = wide ? a.threads : kOctThreads;          // the 1024-thread build also runs with 512
    auto kernel = wide ? octree_kernel<kOctThreadsBig> : octree_kernel<kOctThreads>;
    const size_t smem = octree_smem_bytes(a.smemKeys, a.maxNodeCap, nthr);
    // per device and cheap: set on every launch rather than cached in a process-wide flag (a process may drive several GPUs)
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // RUMI_OCTREE_SPLIT=1: one launch per level (per-level timing under ncu)
    static const bool split = [] { const char* e = getenv("RUMI_OCTREE_SPLIT"); return e && e[0] == '1'; }();
    if (split) {
        for (int l = 0; l < oc.nlevels; ++l) {
            OctreeArgs b = a;
            b.levelFirst = l;
            b.bigList = nullptr;                // everything in one pass
            kernel<<<dim3(a.nframes, 1), nthr, smem, s>>>(b, oc);
        }
        return;
    }
    kernel<<<dim3(a.nframes, oc.nlevels), nthr, smem, s>>>(a, oc);
    if (a.bigList) {
        const size_t smemBig = octree_smem_bytes(a.smemKeysBig, a.maxNodeCap, kOctThreadsBig);
        cudaFuncSetAttribute(octree_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBig);
        // RUMI_OCT_BIG_CTAS: persistent CTAs of pass 1 (A/B runs)
        static const int maxCtas = [] { const char* e = getenv("RUMI_OCT_BIG_CTAS"); return e ? std::max(1, atoi(e)) : 148; }();
        const int ctas = std::min(a.nframes * oc.nlevels, maxCtas);
        octree_big_kernel<<<ctas, kOctThreadsBig, smemBig, s>>>(a, oc);
    }
}

}  // namespace rumi
