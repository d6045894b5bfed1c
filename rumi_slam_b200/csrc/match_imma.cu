// K8-T: all-pairs Hamming top-2 on the integer tensor cores, for LARGE query x train problems (cfg 5a / 5b).
//
// north_star keeps "a b1 AND-popc MMA variant only if ncu shows it wins".  On sm_100a ptxas lowers
// mma.sync ... b1.and.popc to IMMA.16832.U8.U8 on operands it unpacks PER INSTRUCTION (tools/probe/b1_probe.cu: 2.7
// pairs/clk/SM in a bare register loop, no better than the LOP3+POPC kernel's 2.6 with the whole top-2 logic); the plain
// int8 IMMA itself runs at 0.48 warp-MMA/clk/SM = 7.6 pairs/clk/SM (tools/probe/pipe_probe.cu).  So the unpacking is
// done ONCE: descriptors are expanded to 256 bytes of 0/1 (K8-X), and
//     Hamming(q, t) = pop(q) + pop(t) - 2 * <q, t>          (<q, t> = 256-long u8 dot product = 8 x m16n8k32 per 16x8 pairs)
// A CTA of 8 warps (2 x 4) holds 128 queries and streams 128-train tiles through a 3-stage cp.async pipeline (one
// barrier per tile);
// a warp owns 64 x 32 pairs = 16 accumulator fragments.  Operands are read with 16-byte LDS: the dot product does not
// care about the order of k, so thread t of a quad simply owns bytes [16t, 16t+16) of every 64-byte group of BOTH
// operands (row pitch 320 B keeps those loads bank-conflict free; the query tile is stored pair-interleaved so that an
// A operand quad is one 16-byte load).  The top-2 of a query row lives in keys
//     (pop(t) - 2 <q,t> + 256) << 22 | train index         (pop(q) is constant per row and added at the very end)
// so that one shift-subtract builds the key and 2.5 min / max per pair update (best, second best) with the reference's
// tie rule (smaller index first).  Two accumulator sets alternate: the top-2 update of tile i-1 is interleaved with
// the IMMAs of tile i, so the integer pipe and the tensor pipe work at the same time inside every warp.  Output = the same per-slice {d1:16, d2:16, idx:32} candidates as K8, merged by K9.
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"

namespace rumi {

constexpr int kImmaBN = 128, kImmaPitch = 320;   // query rows / threads per CTA depend on the variant: 64 * WM / 128 * WM
constexpr int kImmaQPitch = 528;          // one interleaved pair of query rows (g, g + 8): 512 B + 16 B bank skew
constexpr int kImmaIdxBits = 22;                                  // train index inside one call (nt <= 2^22 per launch)
constexpr uint32_t kImmaIdxMask = (1u << kImmaIdxBits) - 1u;

// ---- K8-X: bits -> 0/1 bytes (+ popcount per descriptor).  Thread = one 32-bit word of one descriptor. ----
__global__ void __launch_bounds__(256) expand_bits_kernel(const uint8_t* __restrict__ D, int n, uint8_t* __restrict__ X,
                                                          uint16_t* __restrict__ pop) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // word index
    const bool live = i < (long long)n * 8;
    uint32_t w = live ? reinterpret_cast<const uint32_t*>(D)[i] : 0u;
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (((w >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;    // 4 bits -> 4 bytes
    if (live) {
        uint4* dst = reinterpret_cast<uint4*>(X + 32 * i);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    if (pop) {
        int c = __popc(w);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, 1);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, 2);
        c += __shfl_xor_sync(0xFFFFFFFFu, c, 4);
        if (live && (threadIdx.x & 7) == 0) pop[i >> 3] = (uint16_t)c;
    }
}

void launch_expand_bits(const uint8_t* D, int n, uint8_t* X, uint16_t* pop, cudaStream_t s) {
    if (n <= 0) return;
    const long long words = (long long)n * 8;
    expand_bits_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(D, n, X, pop);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;                                   // src-size 0 -> 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int WM>
struct ImmaSmem {
    uint8_t q[(64 * WM / 2) * kImmaQPitch];        // query rows pair-interleaved in A-fragment order (see load_q)
    uint8_t t[3][kImmaBN * kImmaPitch];
    uint32_t base[3][kImmaBN];                     // ((pop(t) + 256) << 22) | index, 0xFFFFFFFF past the slice end
    uint32_t merge[4][64 * WM][2];                 // per N-warp top-2 of every query row
};

// WM = warps along the query dimension (CTA = WM x 4 warps, 64 * WM queries); kPingPong = two accumulator sets (248
// registers, 8 warps per SM) instead of one (147 registers, 12 warps per SM with WM = 3)
template <int WM, bool kPingPong>
__global__ void __launch_bounds__(128 * WM, 1)
hamming_top2_imma_kernel(const uint8_t* __restrict__ Q, const uint8_t* __restrict__ Qx, int nq,
                         const uint8_t* __restrict__ Tx, const uint16_t* __restrict__ popT, int nt, int sliceRows,
                         int tBase, uint64_t* __restrict__ partial /* [gridDim.y][nq] */) {
    extern __shared__ __align__(16) uint8_t smemRaw[];
    constexpr int kImmaBM = 64 * WM, kImmaThreads = 128 * WM;
    ImmaSmem<WM>& S = *reinterpret_cast<ImmaSmem<WM>*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                         // WM x 4 warps: 64 query rows x 32 train columns each
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * kImmaBM;
    const int t0 = blockIdx.y * sliceRows, t1 = min(t0 + sliceRows, nt);
    const int ntiles = (t1 - t0 + kImmaBN - 1) / kImmaBN;

    auto load_tile = [&](uint8_t* dst, const uint8_t* src, int row0, int rowEnd) {       // 128 rows x 256 B, 16 B chunks
        for (int c = tid; c < kImmaBN * 16; c += kImmaThreads) {
            const int r = c >> 4, k = c & 15;
            const int row = row0 + r;
            cp_async16(dst + r * kImmaPitch + 16 * k, src + ((size_t)min(row, rowEnd - 1) * 256 + 16 * k), row < rowEnd);
        }
    };
    auto load_base = [&](int st, int row0) {
        if (tid < kImmaBN) {
            const int row = row0 + tid;
            S.base[st][tid] = row < t1 ? ((((uint32_t)popT[row] + 256u) << kImmaIdxBits) | (uint32_t)row) : 0xFFFFFFFFu;
        }
    };

    // The A operand of m16n8k32 is 4 CONSECUTIVE registers (row g | row g+8 | row g, k+16 | row g+8, k+16).  The query
    // tile is written once per CTA, so it is stored in that order: rows g and g+8 of a 16-row block share one 512-byte
    // "pair row" in which the 4-byte words of the two rows alternate; a thread then reads its operand quads of two
    // k-steps with two 16-byte loads and no register shuffling.
    auto load_q = [&]() {
        for (int c = tid; c < kImmaBM * 64; c += kImmaThreads) {       // 4-byte words of the 128 x 256 B tile
            const int r = c >> 6, kw = c & 63;                          // row, word inside the row
            const int prow = (r >> 4) * 8 + (r & 7), half = (r >> 3) & 1;
            const int p = kw >> 4, tt = (kw >> 2) & 3, w = kw & 3;      // 64-byte group, owning thread of the quad, word
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(S.q + prow * kImmaQPitch + 128 * p + 32 * tt + 8 * w + 4 * half);
            const int row = q0 + r;
            const int sz = row < nq ? 4 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa),
                         "l"(Qx + ((size_t)min(row, nq - 1) * 256 + 4 * kw)), "r"(sz) : "memory");
        }
    };
    load_q();
    if (ntiles > 0) { load_tile(S.t[0], Tx, t0, t1); load_base(0, t0); }
    cp_async_commit();

    uint32_t k1[8], k2[8];                                           // rows: mb * 16 + g (+ 8), mb = 0..3
#pragma unroll
    for (int i = 0; i < 8; ++i) k1[i] = k2[i] = 0xFFFFFFFFu;

    // top-2 update of the rows of m-block mb from one tile's accumulators: key = base[col] - (dot << 23)
    auto epilogue_mb = [&](const int (&acc)[4][4][4], const uint2 (&bs)[4], int mb) {
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = mb * 2 + h;
                // out-of-slice columns keep 0xFFFFFFFF: their rows were zero-filled, dot = 0
                const uint32_t x = bs[nb].x - ((uint32_t)acc[mb][nb][h * 2] << 23);
                const uint32_t y = bs[nb].y - ((uint32_t)acc[mb][nb][h * 2 + 1] << 23);
                const uint32_t lo = min(x, y), hi = max(x, y);
                const uint32_t a = max(k1[r], lo);
                k1[r] = min(k1[r], lo);
                k2[r] = min(min(k2[r], hi), a);
            }
        }
    };
    // One tile: 128 IMMAs into `cur`, with the top-2 update of the PREVIOUS tile's accumulators (`prev`) spread between
    // the k-steps, so that the integer pipe works while the tensor pipe does -- two accumulator sets in ping-pong.
    // (hasPrev is a compile-time tag: a run-time test would split the IMMAs and the update into separate basic blocks
    // and ptxas would not interleave them)
    auto tile_step = [&](auto hasPrev, int (&cur)[4][4][4], const int (&prev)[4][4][4], const uint2 (&bsPrev)[4], int st) {
#pragma unroll
        for (int mb = 0; mb < 4; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                for (int e = 0; e < 4; ++e) cur[mb][nb][e] = 0;
        const uint8_t* qa = S.q + (wm * 32 + g) * kImmaQPitch + 32 * t;
        const uint8_t* tb = S.t[st] + (wn * 32 + g) * kImmaPitch + 16 * t;
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {                             // 64 bytes of k per step = 2 MMA k-steps
            uint4 b[4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const uint4*>(tb + nb * 8 * kImmaPitch + 64 * ss);
            // all 16 accumulators take k-step 2 ss, then all 16 take k-step 2 ss + 1: two IMMAs on the same accumulator
            // are 16 instructions apart, far more than the IMMA latency
            uint4 a0[4], a1[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                a0[mb] = *reinterpret_cast<const uint4*>(qa + mb * 8 * kImmaQPitch + 128 * ss);
                a1[mb] = *reinterpret_cast<const uint4*>(qa + mb * 8 * kImmaQPitch + 128 * ss + 16);
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) imma16832(cur[mb][nb], a0[mb].x, a0[mb].y, a0[mb].z, a0[mb].w, b[nb].x, b[nb].y);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) imma16832(cur[mb][nb], a1[mb].x, a1[mb].y, a1[mb].z, a1[mb].w, b[nb].z, b[nb].w);
            if constexpr (decltype(hasPrev)::value) epilogue_mb(prev, bsPrev, ss);   // a quarter of the previous tile's update
        }
    };
    auto load_bases = [&](uint2 (&bs)[4], int st) {
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) bs[nb] = *reinterpret_cast<const uint2*>(&S.base[st][wn * 32 + nb * 8 + 2 * t]);
    };

    int acc0[4][4][4], acc1[kPingPong ? 4 : 1][4][4];
    uint2 bs0[4], bs1[4];
    // tile tl lives in stage tl % 3; the stage filled by the prefetch was last read two tiles ago, so ONE barrier per
    // tile (after the wait) is enough
    auto advance = [&](int tl) {
        if (tl + 1 < ntiles) {
            load_tile(S.t[(tl + 1) % 3], Tx, t0 + (tl + 1) * kImmaBN, t1);
            load_base((tl + 1) % 3, t0 + (tl + 1) * kImmaBN);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
    };
    if constexpr (kPingPong) {
        if (ntiles > 0) {
            advance(0);
            tile_step(std::false_type{}, acc0, acc1, bs1, 0);
            load_bases(bs0, 0);
        }
        for (int tl = 1; tl < ntiles; tl += 2) {
            advance(tl);
            tile_step(std::true_type{}, acc1, acc0, bs0, tl % 3);
            load_bases(bs1, tl % 3);
            if (tl + 1 < ntiles) {                                   // CTA uniform
                advance(tl + 1);
                tile_step(std::true_type{}, acc0, acc1, bs1, (tl + 1) % 3);
                load_bases(bs0, (tl + 1) % 3);
            }
        }
        if (ntiles > 0) {                                            // update from the last tile
            if ((ntiles - 1) & 1) {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) epilogue_mb(acc1, bs1, mb);
            } else {
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) epilogue_mb(acc0, bs0, mb);
            }
        }
    } else {
        for (int tl = 0; tl < ntiles; ++tl) {                        // one accumulator set: MMAs, then the update; the
            advance(tl);                                             // overlap comes from the other warps of the SM
            tile_step(std::false_type{}, acc0, acc0, bs0, tl % 3);
            load_bases(bs0, tl % 3);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) epilogue_mb(acc0, bs0, mb);
        }
    }

    // ---- merge the 4 lanes of a quad, then the 4 N-warps, then write the slice candidate ----
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, k1[r], o), o2 = __shfl_xor_sync(0xFFFFFFFFu, k2[r], o);
            const uint32_t hi = max(k1[r], o1);
            k1[r] = min(k1[r], o1);
            k2[r] = min(min(k2[r], o2), hi);
        }
        if (t == 0) {
            const int row = wm * 64 + (r >> 1) * 16 + (r & 1) * 8 + g;
            S.merge[wn][row][0] = k1[r];
            S.merge[wn][row][1] = k2[r];
        }
    }
    __syncthreads();
    if (tid < kImmaBM) {
        const int qi = q0 + tid;
        uint32_t a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint32_t b1 = S.merge[w][tid][0], b2 = S.merge[w][tid][1];
            const uint32_t hi = max(a1, b1);
            a1 = min(a1, b1);
            a2 = min(min(a2, b2), hi);
        }
        if (qi < nq) {
            const uint4* qp = reinterpret_cast<const uint4*>(Q + 32 * (size_t)qi);
            const uint4 x = qp[0], y = qp[1];
            const int pq = __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w) + __popc(y.x) + __popc(y.y) +
                           __popc(y.z) + __popc(y.w);
            // key >> 22 = pop(t) - 2 dot + 256 ; distance = that - 256 + pop(q).  The reference's scan starts from
            // bestDist = 256 with strict '<': a distance of 256 is "no match".
            int d1 = a1 == 0xFFFFFFFFu ? 256 : (int)(a1 >> kImmaIdxBits) - 256 + pq;
            int d2 = a2 == 0xFFFFFFFFu ? 256 : (int)(a2 >> kImmaIdxBits) - 256 + pq;
            d1 = min(d1, 256); d2 = min(d2, 256);
            const uint32_t idx = d1 >= 256 ? 0xFFFFFFFFu : (uint32_t)tBase + (a1 & kImmaIdxMask);
            partial[(size_t)blockIdx.y * nq + qi] = ((uint64_t)d1 << 48) | ((uint64_t)d2 << 32) | (uint64_t)idx;
        }
    }
}

// Variant: RUMI_IMMA_VARIANT = "3n" (default: 3 x 4 warps, one accumulator set, 12 warps per SM), "2n" (2 x 4 warps),
// "2p" (2 x 4 warps, ping-pong accumulators).  Measured on B200, 40000 x 40000: 1.16 / 1.07-1.12 / 1.12-1.13 e12 pairs/s
// -- the three land within 8 %: the tensor pipe sits at ~50 % in all of them.
static int imma_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RUMI_IMMA_VARIANT");
        v = 2;
        if (e && e[0] == '2') v = e[1] == 'n' ? 1 : 0;
    }
    return v;
}
static int imma_bm() { return imma_variant() == 2 ? 192 : 128; }

int imma_slices(int nq, int nt) {
    const int qBlocks = (nq + imma_bm() - 1) / imma_bm();
    const int maxSlices = (nt + 8 * kImmaBN - 1) / (8 * kImmaBN);     // at least 8 tiles per slice
    int want = (148 * 8 + qBlocks - 1) / qBlocks;                      // >= 8 waves of one CTA per SM: tail < 10 %
    if (want > maxSlices) want = maxSlices;
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    return want;
}

template <int WM, bool PP>
static void launch_imma_variant(const uint8_t* Q, const uint8_t* Qx, int nq, const uint8_t* Tx, const uint16_t* popT, int nt,
                                int sliceRows, int tBase, int slices, uint64_t* partial, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(hamming_top2_imma_kernel<WM, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ImmaSmem<WM>));
        configured = true;
    }
    dim3 grid((nq + 64 * WM - 1) / (64 * WM), slices);
    hamming_top2_imma_kernel<WM, PP><<<grid, 128 * WM, sizeof(ImmaSmem<WM>), s>>>(Q, Qx, nq, Tx, popT, nt, sliceRows, tBase,
                                                                                  partial);
}

void launch_hamming_top2_imma(const uint8_t* Q, const uint8_t* Qx, int nq, const uint8_t* Tx, const uint16_t* popT,
                              int nt, int tBase, int slices, uint64_t* partial, cudaStream_t s) {
    int sliceRows = (nt + slices - 1) / slices;
    sliceRows = (sliceRows + kImmaBN - 1) / kImmaBN * kImmaBN;
    switch (imma_variant()) {
        case 1: launch_imma_variant<2, false>(Q, Qx, nq, Tx, popT, nt, sliceRows, tBase, slices, partial, s); break;
        case 2: launch_imma_variant<3, false>(Q, Qx, nq, Tx, popT, nt, sliceRows, tBase, slices, partial, s); break;
        default: launch_imma_variant<2, true>(Q, Qx, nq, Tx, popT, nt, sliceRows, tBase, slices, partial, s); break;
    }
}

}  // namespace rumi
