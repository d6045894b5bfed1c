// Quad-tree keypoint distribution (ORBextractor::DistributeOctTree, R/lib_src/ORBextractor.cc:538-724) re-designed
// for one thread block per (frame, level) problem.  NOT a translation of the list-based reference:
//
//   1. every candidate gets its full root+path code (the tree geometry is data independent, :471-522), and the
//      candidates are sorted by (code, insertion order) -> every tree node is a contiguous range;
//   2. the level-synchronous phase (:587-650) is evaluated in closed form from the depth at which neighbouring codes
//      first differ: |list| after pass t == number of distinct depth-t prefixes, nToExpand == groups with > 1 key;
//   3. the node list after that phase is produced in the reference's std::list order by sorting leaves on
//      (creation pass desc, prefix with alternating digit directions) -- push_front reverses the visiting order
//      once per pass;
//   4. only the final "largest nodes first" phase (:651-702) is serial (thread 0): it replays libstdc++'s
//      std::sort (introsort, non-total comparator compareNodes :524-536) on (count, UL.x) pairs so that ties break
//      exactly as in the reference, then splits from the back until the quota is reached;
//   5. per-leaf arg-max of the response (:706-721; strict '>' => first in insertion order) in parallel.
//
// The same source compiles for the device (blockDim.x threads, __syncthreads) and for the host test harness
// (1 thread), which is how it is checked against the oracle without a GPU.
#pragma once
#include "orb_common.h"
#include "orb_math.cuh"

namespace rumi {

constexpr int kKeyCodeShift = kOrderBits + 8;        // key = code << 30 | order << 8 | response  (30 + 22 + 8 <= 64 bits)
constexpr int kRankBits = 20;

RUMI_HD uint64_t make_tree_key(uint32_t code, uint32_t order, uint32_t resp) {
    return ((uint64_t)code << kKeyCodeShift) | ((uint64_t)order << 8) | (uint64_t)resp;
}
RUMI_HD uint32_t key_code(uint64_t k) { return (uint32_t)(k >> kKeyCodeShift); }
RUMI_HD uint32_t key_order(uint64_t k) { return (uint32_t)(k >> 8) & ((1u << kOrderBits) - 1u); }
RUMI_HD uint32_t key_resp(uint64_t k) { return (uint32_t)k & 0xFFu; }

RUMI_HD int highest_bit(uint32_t v) {     // v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

// Depth at which two codes first differ: 0 = different roots, j = digit j (1-based from the root), D+1 = identical.
RUMI_HD int first_diff_depth(uint32_t a, uint32_t b, int D) {
    const uint32_t x = a ^ b;
    if (x == 0) return D + 1;
    const int hb = highest_bit(x);
    if (hb >= 2 * D) return 0;
    return D - (hb >> 1);
}

// Bits to flip so that ascending order of (code ^ mask) on the first `cd` digits equals the reference's list order
// of the nodes created in pass `cd` (children are pushed to the FRONT of the list while their parents are visited
// front to back, so each pass reverses the visiting order of the pass before).
RUMI_HD uint32_t list_order_mask(int cd, int D) {
    uint32_t m = 0;
    for (int i = 1; i <= cd; ++i)
        if (((cd - i) & 1) == 0) m |= 3u << (2 * (D - i));
    if (cd & 1) m |= ((1u << kRootBits) - 1u) << (2 * D);
    return m;
}

struct OctreeWork {
    uint64_t* keys;        // [Mpad]  sorted tree keys (padding = ~0)
    uint64_t* lkeys;       // [Gpad]  leaf list keys, sorted = reference list order
    uint32_t* glo;         // [nodeCap + 1] first key of each depth-t* group, in code order (+ sentinel M)
    uint32_t* cr_lo;       // [createCap] nodes created by the sorted phase (range start)
    uint32_t* cr_cnt;      // [createCap] (count; 0 = erased)
    uint64_t* pend;        // [pendCap] sort elements: (count << 13 | UL.x) << 32 | node ref
    uint64_t* lsort;       // [pendCap] scratch of the parallel stable rank sort
    uint32_t* next;        // [pendCap] node refs of the current round
    uint32_t* next2;       // [pendCap] node refs of the next round
    uint32_t* meta;        // [pendCap] per pending node: non-empty children | multi-key children << 4
    uint32_t* qbase;       // [pendCap] per split node (processing order): created base | next base << 16
    int* hist;             // [2 * (kMaxTreeDepth + 2)]
    int* part;             // [nthreads + 1] per-thread partial counts
    int* scal;             // [SC_COUNT] scalars shared between threads
    int nodeCap, createCap, pendCap;
};

enum { SC_TSTAR = 0, SC_PHASEB = 1, SC_G = 2, SC_NOUT = 3, SC_NCREATED = 4, SC_NPEND = 5, SC_SIZE = 6, SC_DEPTH = 7,
       SC_DONE = 8, SC_J = 9, SC_NEWC = 10, SC_NEWM = 11, SC_TIE = 12, SC_NREPLAY = 13, SC_COUNT = 14 };
constexpr uint32_t kRefListBit = 0x80000000u;        // node ref: bit 31 set -> position in the leaf list

// ---- bitonic sort of n (power of two) u64 ascending; all threads of the block participate ----
// Pair p of stage (k, j) touches elements i = ((p & ~(j-1)) << 1) | (p & (j-1)) and i | j.  Pairs are assigned to
// warps in fixed ranges of kSortSegPairs, so every stage with j <= kSortSegPairs stays inside one warp's 2 *
// kSortSegPairs elements and only needs a warp barrier; block barriers are left for the few stages with larger j
// (6 instead of 66 for 2048 keys).
constexpr int kSortSegPairs = 128;

RUMI_HD void bitonic_cx(uint64_t* a, int p, int k, int j) {
    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
    const int q = i | j;
    const uint64_t x = a[i], y = a[q];
    const bool up = (i & k) == 0;
    if ((x > y) == up) { a[i] = y; a[q] = x; }
}

template <class Ctx>
RUMI_HD void bitonic_sort_u64(Ctx& ctx, uint64_t* a, int n) {
    const int half = n >> 1;
    const int nseg = half > kSortSegPairs ? half / kSortSegPairs : 1;
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        for (; j > kSortSegPairs; j >>= 1) {                 // block-wide stages
            ctx.sync();
            for (int p = ctx.tid; p < half; p += ctx.nthr) bitonic_cx(a, p, k, j);
        }
        if (k > 2 * kSortSegPairs) ctx.sync();               // the warp-local stages read what other warps wrote
        for (int seg = ctx.warp_id(); seg < nseg; seg += ctx.num_warps()) {
            const int pEnd = (seg + 1) * kSortSegPairs < half ? (seg + 1) * kSortSegPairs : half;
            for (int jj = j; jj > 0; jj >>= 1) {
                for (int p = seg * kSortSegPairs + ctx.lane(); p < pEnd; p += ctx.warp_size()) bitonic_cx(a, p, k, jj);
                ctx.sync_warp();
            }
        }
    }
    ctx.sync();
}

RUMI_HD int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---- libstdc++ std::sort replay (bits/stl_algo.h: __introsort_loop / __unguarded_partition_pivot /
// __move_median_to_first / __final_insertion_sort, threshold 16, depth limit 2*lg(n), heapsort fallback) ----
// Elements are u64; the comparator looks at the high 32 bits only (count, UL.x), exactly like compareNodes.
namespace stdsort {
RUMI_HD bool lt(uint64_t a, uint64_t b) { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
RUMI_HD void swp(uint64_t* v, int i, int j) { const uint64_t t = v[i]; v[i] = v[j]; v[j] = t; }

RUMI_HD void push_heap(uint64_t* v, int first, int hole, int top, uint64_t val) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt(v[first + parent], val)) {
        v[first + hole] = v[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[first + hole] = val;
}
RUMI_HD void adjust_heap(uint64_t* v, int first, int hole, int len, uint64_t val) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (lt(v[first + child], v[first + child - 1])) --child;
        v[first + hole] = v[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[first + hole] = v[first + child - 1];
        hole = child - 1;
    }
    push_heap(v, first, hole, top, val);
}
RUMI_HD void make_heap(uint64_t* v, int first, int last) {
    const int len = last - first;
    if (len < 2) return;
    int parent = (len - 2) / 2;
    while (true) {
        const uint64_t val = v[first + parent];
        adjust_heap(v, first, parent, len, val);
        if (parent == 0) return;
        --parent;
    }
}
RUMI_HD void pop_heap(uint64_t* v, int first, int last, int result) {
    const uint64_t val = v[result];
    v[result] = v[first];
    adjust_heap(v, first, 0, last - first, val);
}
// __partial_sort(first, last, last) == heap_select + sort_heap over the whole range
RUMI_HD void heap_sort(uint64_t* v, int first, int last) {
    make_heap(v, first, last);
    // __heap_select with middle == last: the loop over [middle, last) is empty
    while (last - first > 1) {
        --last;
        pop_heap(v, first, last, last);
    }
}
RUMI_HD void move_median_to_first(uint64_t* v, int result, int a, int b, int c) {
    if (lt(v[a], v[b])) {
        if (lt(v[b], v[c])) swp(v, result, b);
        else if (lt(v[a], v[c])) swp(v, result, c);
        else swp(v, result, a);
    } else if (lt(v[a], v[c])) swp(v, result, a);
    else if (lt(v[b], v[c])) swp(v, result, c);
    else swp(v, result, b);
}
RUMI_HD int unguarded_partition(uint64_t* v, int first, int last, int pivot) {
    // the pivot sits in front of [first, last) and is never swapped: its key is read once (this loop is the serial
    // critical path of the kernel); the two elements of a swap are already in registers when it happens
    const uint32_t pk = (uint32_t)(v[pivot] >> 32);
    while (true) {
        uint64_t a = v[first];
        while ((uint32_t)(a >> 32) < pk) a = v[++first];
        --last;
        uint64_t b = v[last];
        while (pk < (uint32_t)(b >> 32)) b = v[--last];
        if (!(first < last)) return first;
        v[first] = b; v[last] = a;
        ++first;
    }
}
RUMI_HD void unguarded_linear_insert(uint64_t* v, int last) {
    const uint64_t val = v[last];
    int next = last - 1;
    while (lt(val, v[next])) {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}
RUMI_HD void insertion_sort(uint64_t* v, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (lt(v[i], v[first])) {
            const uint64_t val = v[i];
            for (int k = i; k > first; --k) v[k] = v[k - 1];     // move_backward(first, i, i + 1)
            v[first] = val;
        } else {
            unguarded_linear_insert(v, i);
        }
    }
}
// __introsort_loop only (the partitioning phase of std::sort): leaves segments of <= 16 elements unsorted
RUMI_HD void introsort_loop(uint64_t* v, int n) {
    if (n < 2) return;
    // tail recursion on the left part turned into an explicit stack of left parts
    int depth_limit = 2 * highest_bit((uint32_t)n);
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int first = 0, last = n;
    while (true) {
        while (last - first > 16) {
            if (depth_limit == 0) {
                heap_sort(v, first, last);
                break;                                   // this range is done
            }
            --depth_limit;
            const int mid = first + (last - first) / 2;
            move_median_to_first(v, first, first + 1, mid, last - 1);
            const int cut = unguarded_partition(v, first + 1, last, first);
            // reference recurses on [cut, last) first, then loops on [first, cut)
            stack_first[sp] = first; stack_last[sp] = cut; stack_depth[sp] = depth_limit; ++sp;
            first = cut;
        }
        if (sp == 0) break;
        --sp;
        first = stack_first[sp]; last = stack_last[sp]; depth_limit = stack_depth[sp];
    }
}
RUMI_HD void final_insertion_sort(uint64_t* v, int n) {
    if (n > 16) {
        insertion_sort(v, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(v, i);
    } else {
        insertion_sort(v, 0, n);
    }
}
RUMI_HD void sort(uint64_t* v, int n) {
    if (n < 2) return;
    introsort_loop(v, n);
    final_insertion_sort(v, n);
}
}  // namespace stdsort

// Number of keys of [lo, lo+cnt) whose code digit at `depth` (1-based) is < q, i.e. start of child q.
RUMI_HD uint32_t child_start(const uint64_t* keys, uint32_t lo, uint32_t cnt, int shift, uint32_t q) {
    uint32_t a = lo, b = lo + cnt;
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        if (((key_code(keys[m]) >> shift) & 3u) < q) a = m + 1; else b = m;
    }
    return a;
}

// The whole distribution for one (frame, level).  cand[0..M) are packed candidates in the reference's insertion
// order.  Selected candidates are written to out[] (packed, level coordinates relative to (16,16)) in the
// reference's output order; returns their number through scal[SC_NOUT].
template <class Ctx>
RUMI_HD void distribute_quadtree(Ctx& ctx, const uint32_t* cand, int M, int N, const LevelGeom& g, OctreeWork& w,
                                 uint32_t* out, int outCap) {
    const int D = g.treeDepth;
    const int height = g.h - 2 * kMinBorder;
    const int Mpad = next_pow2(M > 8 ? M : 8);

    // 1. keys
    // four candidates per thread and trip: their (L2-latency) loads are in flight together
    for (int i0 = ctx.tid; i0 < Mpad; i0 += 4 * ctx.nthr) {
        uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ctx.nthr;
            c[u] = i < M ? cand[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * ctx.nthr;
            if (i >= Mpad) break;
            w.keys[i] = i < M ? make_tree_key(tree_code(cand_x(c[u]), cand_y(c[u]), g.hX, g.nIni, height, D), (uint32_t)i,
                                              cand_resp(c[u]))
                              : ~0ull;
        }
    }
    for (int i = ctx.tid; i < 2 * (kMaxTreeDepth + 2); i += ctx.nthr) w.hist[i] = 0;
    ctx.sync();
    ctx.mark(0);
    {
        // buckets of the key sort: root index and the first p digits of the path code (p <= 4, at most 1024 buckets)
        int p = D < 4 ? D : 4;
        while (p > 1 && (g.nIni << (2 * p)) > 1024) --p;
        ctx.sort_tree_keys(w.keys, M, Mpad, kKeyCodeShift + 2 * (D - p), g.nIni << (2 * p));
    }
    ctx.mark(1);

    // 2. closed-form level-synchronous phase
    int* histB = w.hist;                          // boundaries first visible at depth u
    int* histS = w.hist + (kMaxTreeDepth + 2);    // keys that become singletons at depth u
    for (int i = ctx.tid; i < M; i += ctx.nthr) {
        const uint32_t c = key_code(w.keys[i]);
        const int a = i == 0 ? 0 : first_diff_depth(key_code(w.keys[i - 1]), c, D);
        const int b = i == M - 1 ? 0 : first_diff_depth(c, key_code(w.keys[i + 1]), D);
        ctx.atomic_add(&histB[a], 1);
        ctx.atomic_add(&histS[a > b ? a : b], 1);
    }
    ctx.sync();
    if (ctx.tid == 0) {
        int size = histB[0], singles = histS[0];
        int tstar = 0, phaseB = 0;
        if (M > 0) {
            int prev = size;
            for (int t = 1;; ++t) {
                if (t <= D) { size += histB[t]; singles += histS[t]; }
                tstar = t;
                const int nexp = size - singles;
                if (size >= N || size == prev) break;
                if (size + 3 * nexp > N) { phaseB = 1; break; }
                prev = size;
            }
            if (tstar > D) tstar = D;       // only reached when nothing can be split any more
            w.scal[SC_G] = 0;
            int s = 0;
            for (int t = 0; t <= tstar; ++t) s += histB[t];
            w.scal[SC_G] = s;
        } else {
            w.scal[SC_G] = 0;
        }
        w.scal[SC_TSTAR] = tstar;
        w.scal[SC_PHASEB] = phaseB;
        w.scal[SC_NOUT] = 0;
        w.scal[SC_NCREATED] = 0;
        w.scal[SC_NREPLAY] = 0;
    }
    ctx.sync();
    ctx.mark(2);
    const int tstar = w.scal[SC_TSTAR];
    const int G = w.scal[SC_G];
    if (M == 0 || G == 0) return;

    // 3. groups at depth t* in code order (ordered compaction of the boundaries), then the list order
    {
        const int chunk = (M + ctx.nthr - 1) / ctx.nthr;
        const int i0 = ctx.tid * chunk, i1 = (i0 + chunk < M) ? i0 + chunk : M;
        int c = 0;
        for (int i = i0; i < i1; ++i) {
            const int a = i == 0 ? 0 : first_diff_depth(key_code(w.keys[i - 1]), key_code(w.keys[i]), D);
            c += a <= tstar;
        }
        int tot;
        int r = ctx.exclusive_scan(c, w.part, &tot);
        if (ctx.tid == 0) w.glo[G] = (uint32_t)M;
        for (int i = i0; i < i1; ++i) {
            const int a = i == 0 ? 0 : first_diff_depth(key_code(w.keys[i - 1]), key_code(w.keys[i]), D);
            if (a <= tstar) w.glo[r++] = (uint32_t)i;
        }
        ctx.sync();
    }
    ctx.mark(3);
    const int Gpad = next_pow2(G > 8 ? G : 8);
    for (int r = ctx.tid; r < Gpad; r += ctx.nthr) {
        uint64_t lk = ~0ull;
        if (r < G) {
            const uint32_t lo = w.glo[r], hi = w.glo[r + 1];
            const uint32_t c = key_code(w.keys[lo]);
            int cd = tstar;                                   // pass in which this leaf was created
            if (hi - lo == 1) {
                const int a = lo == 0 ? 0 : first_diff_depth(key_code(w.keys[lo - 1]), c, D);
                const int b = hi == (uint32_t)M ? 0 : first_diff_depth(c, key_code(w.keys[hi]), D);
                cd = a > b ? a : b;
                if (cd > tstar) cd = tstar;
            }
            const int shift = 2 * (D - cd);
            const uint32_t pref = ((c ^ list_order_mask(cd, D)) >> shift) << shift;
            lk = ((uint64_t)(D - cd) << (2 * kMaxTreeDepth + kRootBits + kRankBits)) |
                 ((uint64_t)pref << kRankBits) | (uint64_t)r;
        }
        w.lkeys[r] = lk;
    }
    ctx.sync();
    ctx.sort_u64(w.lkeys, Gpad);
    ctx.mark(4);

    // 4. sorted final phase.  Only the std::sort replay and the "split from the back until the quota is reached"
    //    scan are sequential (thread 0); everything around them -- sort keys, child ranges, node creation, list
    //    assembly -- is data parallel.
    const uint32_t rankMask = (1u << kRankBits) - 1u;
    auto node_range = [&](uint32_t ref, uint32_t& lo, uint32_t& cnt) {
        if (ref & kRefListBit) {
            const uint32_t r = (uint32_t)w.lkeys[ref & ~kRefListBit] & rankMask;
            lo = w.glo[r]; cnt = w.glo[r + 1] - lo;
        } else {
            lo = w.cr_lo[ref]; cnt = w.cr_cnt[ref];
        }
    };
    if (w.scal[SC_PHASEB]) {
        // pending = multi-key nodes created in pass t*, in creation order == reverse list order (ordered compaction)
        {
            const int chunk = (G + ctx.nthr - 1) / ctx.nthr;
            const int i0 = ctx.tid * chunk < G ? ctx.tid * chunk : G, i1 = (i0 + chunk < G) ? i0 + chunk : G;
            int c = 0;
            for (int i = i0; i < i1; ++i) {                     // i counts from the BACK of the list
                const uint32_t r = (uint32_t)w.lkeys[G - 1 - i] & rankMask;
                c += (w.glo[r + 1] - w.glo[r]) > 1;
            }
            int sum;
            int pos = ctx.exclusive_scan(c, w.part, &sum);
            if (ctx.tid == 0) {
                w.scal[SC_NPEND] = sum;
                w.scal[SC_SIZE] = G;
                w.scal[SC_DEPTH] = tstar;
                w.scal[SC_DONE] = 0;
            }
            for (int i = i0; i < i1; ++i) {
                const int k = G - 1 - i;
                const uint32_t r = (uint32_t)w.lkeys[k] & rankMask;
                if ((w.glo[r + 1] - w.glo[r]) > 1) w.next[pos++] = kRefListBit | (uint32_t)k;
            }
            ctx.sync();
        }
        ctx.mark(5);
        uint32_t* cur = w.next;
        uint32_t* nxt = w.next2;
        while (true) {
            const int npend = w.scal[SC_NPEND];
            const int depth = w.scal[SC_DEPTH];
            const int shift = 2 * (D - depth - 1);
            // (R1) sort key (count, UL.x) and number of non-empty / multi-key children of every pending node
            for (int p = ctx.tid; p < npend; p += ctx.nthr) {
                uint32_t lo, cnt;
                node_range(cur[p], lo, cnt);
                const uint32_t prefix = key_code(w.keys[lo]) >> (2 * (D - depth));
                const uint32_t ulx = (uint32_t)tree_node_ulx(prefix, depth, g.hX);
                w.pend[p] = ((uint64_t)((cnt << 13) | ulx) << 32) | (uint64_t)p;
                uint32_t b[5];
                b[0] = lo; b[4] = lo + cnt;
                if (depth + 1 <= D) {
                    b[1] = child_start(w.keys, lo, cnt, shift, 1);
                    b[2] = child_start(w.keys, lo, cnt, shift, 2);
                    b[3] = child_start(w.keys, lo, cnt, shift, 3);
                } else {
                    b[1] = b[2] = b[3] = lo + cnt;
                }
                uint32_t ne = 0, nm = 0;
                for (int q = 0; q < 4; ++q) { const uint32_t cc = b[q + 1] - b[q]; ne += cc > 0; nm += cc > 1; }
                w.meta[p] = ne | (nm << 4);
            }
            ctx.sync();
            ctx.mark(6);
            // (R2) libstdc++ std::sort replay.  Only __introsort_loop (the partitioning) is order dependent and serial
            // (thread 0).  __final_insertion_sort is a STABLE sort of whatever the partitioning left (strict '<' in
            // __unguarded_linear_insert, and an element smaller than the current minimum goes to the front), so its
            // result is reproduced in parallel: position = #(smaller keys) + #(equal keys that sit earlier).
            // Elements whose key is unique end up at the same position whatever the partitioning did; only the order
            // INSIDE a block of equal keys depends on it.  So the stable rank of the ORIGINAL order is tried first: if
            // no two equal keys meet inside the part of the array that (R3) actually consumes (the J largest elements,
            // and the element just below them), that is the reference's result and the serial replay is skipped.
            const int size0 = w.scal[SC_SIZE];
            auto rank_sort = [&]() {                               // w.pend -> w.lsort
                for (int i = ctx.tid; i < npend; i += ctx.nthr) {
                    const uint64_t e = w.pend[i];
                    const uint32_t ke = (uint32_t)(e >> 32);
                    int rank = 0;
                    for (int j = 0; j < npend; ++j) {
                        const uint32_t kj = (uint32_t)(w.pend[j] >> 32);
                        rank += (kj < ke) || (kj == ke && j < i);
                    }
                    w.lsort[rank] = e;
                }
                if (ctx.tid == 0) { w.scal[SC_J] = npend; w.scal[SC_TIE] = 0; }
                ctx.sync();
            };
            // (R3) split from the back of the sorted array until size >= N: prefix sums over the processing order q
            // (q-th processed = sorted position npend-1-q) of the children each split creates.
            auto split_scan = [&]() {                              // reads w.lsort
                const uint64_t* arr = w.lsort;
                const int chunk = (npend + ctx.nthr - 1) / ctx.nthr;
                const int q0 = ctx.tid * chunk < npend ? ctx.tid * chunk : npend;
                const int q1 = q0 + chunk < npend ? q0 + chunk : npend;
                uint32_t loc = 0;
                for (int q = q0; q < q1; ++q) {
                    const uint32_t m = w.meta[(uint32_t)arr[npend - 1 - q]];
                    loc += (m & 15u) | ((m >> 4) << 16);
                }
                int tot;
                uint32_t run = (uint32_t)ctx.exclusive_scan((int)loc, w.part, &tot);
                for (int q = q0; q < q1; ++q) {
                    const uint32_t m = w.meta[(uint32_t)arr[npend - 1 - q]];
                    w.qbase[q] = run;                                  // created before q | multi-key before q << 16
                    run += (m & 15u) | ((m >> 4) << 16);
                    // size after processing q = size0 + (children created so far) - (nodes split so far)
                    if (size0 + (int)(run & 0xFFFFu) - (q + 1) >= N) ctx.atomic_min(&w.scal[SC_J], q + 1);
                }
                ctx.sync();
                const int J = w.scal[SC_J];
                if (J > 0 && J - 1 >= q0 && J - 1 < q1) {                  // owner of the last processed node
                    const uint32_t m = w.meta[(uint32_t)arr[npend - J]];
                    const uint32_t endRun = w.qbase[J - 1] + ((m & 15u) | ((m >> 4) << 16));
                    const int size = size0 + (int)(endRun & 0xFFFFu) - J;
                    w.scal[SC_SIZE] = size;
                    w.scal[SC_NEWC] = (int)(endRun & 0xFFFFu);
                    w.scal[SC_NEWM] = (int)(endRun >> 16);
                    w.scal[SC_DONE] = (size >= N || size == size0) ? 1 : 0;
                }
                if (npend == 0 && ctx.tid == 0) {
                    w.scal[SC_NEWC] = 0; w.scal[SC_NEWM] = 0; w.scal[SC_DONE] = 1;
                }
                // equal keys inside the consumed suffix [npend - J, npend) (or across its lower edge)?
                for (int sIdx = npend - J + ctx.tid; sIdx < npend; sIdx += ctx.nthr)
                    if (sIdx > 0 && (uint32_t)(arr[sIdx] >> 32) == (uint32_t)(arr[sIdx - 1] >> 32)) w.scal[SC_TIE] = 1;
                ctx.sync();
            };
            rank_sort();
            split_scan();
            if (w.scal[SC_TIE]) {
                ctx.sync();                                        // everybody has read SC_TIE before it is reset
                if (ctx.tid == 0) { stdsort::introsort_loop(w.pend, npend); w.scal[SC_NREPLAY] += 1; }
                ctx.sync();
                rank_sort();
                split_scan();
            }
            for (int i = ctx.tid; i < npend; i += ctx.nthr) w.pend[i] = w.lsort[i];
            ctx.sync();
            ctx.mark(7);
            // (R4) create the children of the J split nodes (processing order = from the back of the sorted array)
            const int J = w.scal[SC_J];
            const int cbase0 = w.scal[SC_NCREATED];
            for (int q = ctx.tid; q < J; q += ctx.nthr) {
                const uint32_t p = (uint32_t)w.pend[npend - 1 - q];
                const uint32_t ref = cur[p];
                uint32_t lo, cnt;
                node_range(ref, lo, cnt);
                if (ref & kRefListBit) w.lkeys[ref & ~kRefListBit] |= 1ull << 63;     // erased from the list
                else w.cr_cnt[ref] = 0;
                uint32_t b[5];
                b[0] = lo; b[4] = lo + cnt;
                if (depth + 1 <= D) {
                    b[1] = child_start(w.keys, lo, cnt, shift, 1);
                    b[2] = child_start(w.keys, lo, cnt, shift, 2);
                    b[3] = child_start(w.keys, lo, cnt, shift, 3);
                } else {
                    b[1] = b[2] = b[3] = lo + cnt;
                }
                uint32_t ci = (uint32_t)cbase0 + (w.qbase[q] & 0xFFFFu), mi = w.qbase[q] >> 16;
                for (int qd = 0; qd < 4; ++qd) {
                    const uint32_t c2 = b[qd + 1] - b[qd];
                    if (c2 == 0) continue;
                    w.cr_lo[ci] = b[qd];
                    w.cr_cnt[ci] = c2;
                    if (c2 > 1) nxt[mi++] = ci;
                    ++ci;
                }
            }
            ctx.sync();
            if (ctx.tid == 0) {
                int ncreated = cbase0 + w.scal[SC_NEWC];
                int nnext = w.scal[SC_NEWM];
                if (!w.scal[SC_DONE] && ncreated + 2 * (w.nodeCap + 4) > w.createCap) {
                    // rare: drop erased entries so one more round of children fits (keeps creation order)
                    int wpos = 0, pn = 0;
                    for (int c = 0; c < ncreated; ++c) {
                        if (w.cr_cnt[c] == 0) continue;
                        if (pn < nnext && nxt[pn] == (uint32_t)c) nxt[pn++] = (uint32_t)wpos;
                        w.cr_lo[wpos] = w.cr_lo[c];
                        w.cr_cnt[wpos] = w.cr_cnt[c];
                        ++wpos;
                    }
                    ncreated = wpos;
                }
                w.scal[SC_NCREATED] = ncreated;
                w.scal[SC_NPEND] = nnext;
                w.scal[SC_DEPTH] = depth + 1;
            }
            ctx.sync();
            ctx.mark(8);
            if (w.scal[SC_DONE]) break;
            uint32_t* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
    // final list: nodes created by the sorted phase, newest first, then the surviving pass-t* list (two ordered
    // compactions into pend[] as lo << 32 | cnt)
    {
        const int ncreated = w.scal[SC_NCREATED];
        const int total = ncreated + G;
        const int chunk = (total + ctx.nthr - 1) / ctx.nthr;
        const int i0 = ctx.tid * chunk < total ? ctx.tid * chunk : total, i1 = (i0 + chunk < total) ? i0 + chunk : total;
        auto leaf = [&](int i, uint64_t& out) {          // i-th element of [created reversed | list]; false = erased
            if (i < ncreated) {
                const int c = ncreated - 1 - i;
                if (w.cr_cnt[c] == 0) return false;
                out = ((uint64_t)w.cr_lo[c] << 32) | (uint64_t)w.cr_cnt[c];
                return true;
            }
            const uint64_t lk = w.lkeys[i - ncreated];
            if (lk >> 63) return false;
            const uint32_t r = (uint32_t)lk & rankMask;
            out = ((uint64_t)w.glo[r] << 32) | (uint64_t)(w.glo[r + 1] - w.glo[r]);
            return true;
        };
        int c = 0;
        uint64_t v;
        for (int i = i0; i < i1; ++i) c += leaf(i, v) ? 1 : 0;
        int sum;
        int pos = ctx.exclusive_scan(c, w.part, &sum);
        if (ctx.tid == 0) w.scal[SC_NOUT] = sum;
        for (int i = i0; i < i1; ++i)
            if (leaf(i, v)) w.pend[pos++] = v;
    }
    ctx.sync();

    ctx.mark(9);
    // 5. best response per leaf; ties keep the first candidate in insertion order
    const int nout = w.scal[SC_NOUT] < outCap ? w.scal[SC_NOUT] : outCap;
    for (int k = ctx.tid; k < nout; k += ctx.nthr) {
        const uint32_t lo = (uint32_t)(w.pend[k] >> 32), cnt = (uint32_t)w.pend[k];
        uint32_t bestResp = 0, bestOrd = 0xFFFFFFFFu;
        for (uint32_t i = lo; i < lo + cnt; ++i) {
            const uint64_t key = w.keys[i];
            const uint32_t r = key_resp(key), o = key_order(key);
            if (r > bestResp || (r == bestResp && o < bestOrd)) { bestResp = r; bestOrd = o; }
        }
        out[k] = cand[bestOrd];
    }
}

}  // namespace rumi
