// Output slots of one frame's key points (ORBextractor::operator(), R/lib_src/ORBextractor.cc:1077-1085): key points inside the
// lapping area fill the output from the back, the others from the front, both in level order.  Two prefix sums replace the
// sequential front / back fill.  Called by a whole CTA of >= 256 threads (the first 256 work, the others only join the barriers): either the stand-alone kernel (describe.cu) or the last
// quad-tree CTA of the frame (octree.cu); the inputs were written by other CTAs, so they are read past L1 (__ldcg).
#pragma once
#include "kernels.cuh"
#include "orb_common.h"

namespace rumi {

__device__ __forceinline__ void assign_slots_frame(const SlotArgs& a, const OrbConst& oc, int f, int tid) {
    __shared__ int s_part[8];
    __shared__ int s_lvlBase[kMaxLevels + 1];
    const int* selCount = a.selCount + (long long)f * oc.nlevels;
    if (tid == 0) {
        int n = 0;
        for (int l = 0; l < oc.nlevels; ++l) { s_lvlBase[l] = n; n += __ldcg(selCount + l); }
        s_lvlBase[oc.nlevels] = n;
    }
    __syncthreads();
    const int nkp = s_lvlBase[oc.nlevels];
    // thread t owns a contiguous run of the concatenated level-ordered keypoints
    const int chunk = (nkp + 255) / 256;
    const int i0 = min(tid * chunk, nkp), i1 = min(i0 + chunk, nkp);
    const uint32_t* sel = a.sel + (long long)f * oc.kpCap;
    auto lapping = [&](int i, int& slotIdx) {
        int l = 0;
        while (i >= s_lvlBase[l + 1]) ++l;
        slotIdx = oc.lv[l].kpBase + (i - s_lvlBase[l]);
        const uint32_t c = __ldcg(sel + slotIdx);
        float x = (float)(cand_x(c) + kMinBorder);
        if (l != 0) x = __fmul_rn(x, oc.lv[l].scale);                 // keypoint->pt *= scale  (:1073-1075)
        return x >= (float)a.lap0 && x <= (float)a.lap1;              // (:1077)
    };
    int nlap = 0;
    for (int i = i0; i < i1; ++i) { int s; nlap += lapping(i, s) ? 1 : 0; }
    // exclusive block scan of the per-thread counts: warp shuffles + the 8 warp totals
    int incl = nlap;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31 && tid < 256) s_part[tid >> 5] = incl;
    __syncthreads();
    int warpBase = 0, total = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) {
        const int v = s_part[wv];
        if (wv < (tid >> 5)) warpBase += v;
        total += v;
    }
    if (tid == 0) {
        a.nkp[f] = nkp;
        a.nmono[f] = nkp - total;
    }
    int lapBefore = warpBase + incl - nlap;
    int* slot = a.slot + (long long)f * oc.kpCap;
    for (int i = i0; i < i1; ++i) {
        int s;
        const bool lp = lapping(i, s);
        // lapping keypoints fill from the back, the others from the front, both in level order
        slot[s] = lp ? (nkp - 1 - lapBefore) : (i - lapBefore);
        lapBefore += lp ? 1 : 0;
    }
}

}  // namespace rumi
