// K8': stereo row-band best-1 Hamming (Frame::ComputeStereoMatches, R/lib_src/Frame.cc:828-905).
// The reference buckets right keypoints by the image rows their scale-dependent band [y-r, y+r] (r = 2*scale)
// covers and scans the bucket of the left keypoint's row in ascending right index.  Here one WARP owns a left
// keypoint and strides over ALL right keypoints testing band membership, octave +-1 and the disparity window
// directly (a few thousand integer tests), so no bucket table is built; a (distance, index) min-reduction keeps
// the lowest right index among equal distances, which is what the reference's strict '<' scan keeps.
#include "kernels.cuh"

namespace rumi {

__device__ __forceinline__ int hamming256_rows(const uint4* a, const uint4* b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256)
stereo_best1_kernel(const KeyPointRec* __restrict__ Lk, const uint8_t* __restrict__ Ld, int nL,
                    const KeyPointRec* __restrict__ Rk, const uint8_t* __restrict__ Rd, int nR,
                    const float* __restrict__ scaleFactors, int nRows, float minD, float maxD,
                    int32_t* __restrict__ bestR, uint16_t* __restrict__ bestDist) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + warp;
    if (iL >= nL) return;
    const KeyPointRec kl = Lk[iL];
    const int row = (int)kl.y;                                       // vRowIndices[vL]   (:871)
    const float minU = __fsub_rn(kl.x, maxD), maxU = __fsub_rn(kl.x, minD);
    unsigned best = (100u << 20) | 0xFFFFFu;                         // TH_HIGH (:885), no index
    if (row >= 0 && row < nRows && !(maxU < 0)) {
        const uint4* dl = reinterpret_cast<const uint4*>(Ld) + (size_t)iL * 2;
        for (int iR = lane; iR < nR; iR += 32) {
            const KeyPointRec kr = Rk[iR];
            const float r = __fmul_rn(2.0f, scaleFactors[kr.octave]);          // (:847)
            const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));
            if (row < minr || row > maxr) continue;
            if (kr.octave < kl.octave - 1 || kr.octave > kl.octave + 1) continue;   // (:893)
            if (!(kr.x >= minU && kr.x <= maxU)) continue;                          // (:898)
            const unsigned d = (unsigned)hamming256_rows(dl, reinterpret_cast<const uint4*>(Rd) + (size_t)iR * 2);
            const unsigned key = (d << 20) | (unsigned)iR;
            if (d < 100u && key < best) best = key;                  // strict '<' on distance, then lowest index
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    if (lane == 0) {
        const unsigned idx = best & 0xFFFFFu;
        bestR[iL] = idx == 0xFFFFFu ? -1 : (int32_t)idx;
        bestDist[iL] = (uint16_t)(best >> 20);
    }
}

void launch_stereo_best1(const KeyPointRec* Lk, const uint8_t* Ld, int nL, const KeyPointRec* Rk, const uint8_t* Rd,
                         int nR, const float* scaleFactors, int nRows, float minD, float maxD, int32_t* bestR,
                         uint16_t* bestDist, cudaStream_t s) {
    if (nL <= 0) return;
    stereo_best1_kernel<<<(nL + 7) / 8, 256, 0, s>>>(Lk, Ld, nL, Rk, Rd, nR, scaleFactors, nRows, minD, maxD, bestR,
                                                    bestDist);
}

}  // namespace rumi
