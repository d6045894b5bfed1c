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

// ---------------------------------------------------------------------------------------------------------------
// Remainder of Frame::ComputeStereoMatches (R/lib_src/Frame.cc:907-971): for every left keypoint whose best Hamming
// distance beat (TH_HIGH + TH_LOW) / 2, slide an 11x11 window +-5 px along the row of the RIGHT pyramid level of the
// left keypoint's octave (both pyramids are still resident on the device from the two extractor calls, so
// mvImagePyramid never travels to the host), take the L1 distance (exact integer), fit a parabola through the three
// distances around the minimum and turn the sub-pixel column into disparity / depth.  One warp per left keypoint:
// lanes 0..10 own one window shift each.  The median-based outlier cut (:973-984) is a sequential sort and stays on
// the host.
namespace rumi {

__global__ void __launch_bounds__(256)
stereo_refine_kernel(const __grid_constant__ StereoRefineArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + warp;
    if (iL >= a.nL) return;
    float outU = -1.0f, outD = -1.0f;
    int outSad = -1;
    const int bestIdxR = a.bestR[iL];
    const int bestHam = a.bestDist[iL];
    if (bestIdxR >= 0 && bestHam < 75) {                                        // thOrbDist (:832, :908)
        const KeyPointRec kl = a.Lk[iL];
        const int oct = kl.octave;
        const float uR0 = a.Rk[bestIdxR].x;
        const float sf = a.invScale[oct];
        const float scaleduL = roundf(__fmul_rn(kl.x, sf)), scaledvL = roundf(__fmul_rn(kl.y, sf));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, L = 5;
        const LevelView lv = a.left[oct], rv = a.right[oct];
        const float iniu = scaleduR0 + (float)(L - w), endu = scaleduR0 + (float)(L + w + 1);
        if (!(iniu < 0.f || endu >= (float)rv.w)) {
            const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
            unsigned key = 0xFFFFFFFFu;
            int sad = 0;
            if (lane < 2 * L + 1) {
                const int inc = lane - L;
                for (int dy = -w; dy <= w; ++dy) {
                    const uint8_t* pl = lv.ptr + (long long)(cy + dy) * lv.stride + cxl;
                    const uint8_t* pr = rv.ptr + (long long)(cy + dy) * rv.stride + cxr + inc;
#pragma unroll
                    for (int dx = -w; dx <= w; ++dx) sad += abs((int)pl[dx] - (int)pr[dx]);
                }
                key = ((unsigned)sad << 4) | (unsigned)lane;                    // first minimum in ascending incR
            }
            unsigned best = key;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
            const int bl = (int)(best & 15u), bestinc = bl - L;
            const float d1 = (float)__shfl_sync(0xFFFFFFFFu, sad, max(bl - 1, 0));
            const float d2 = (float)__shfl_sync(0xFFFFFFFFu, sad, bl);
            const float d3 = (float)__shfl_sync(0xFFFFFFFFu, sad, min(bl + 1, 2 * L));
            if (bestinc != -L && bestinc != L) {
                // deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2))   (:949)
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
                if (!(deltaR < -1.f || deltaR > 1.f)) {
                    float bestuR = __fmul_rn(a.scale[oct], __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(kl.x, bestuR);
                    if (disparity >= a.minD && disparity < a.maxD) {
                        if (disparity <= 0.f) {
                            disparity = 0.01f;                                   // (float)0.01
                            bestuR = (float)((double)kl.x - 0.01);               // uL - 0.01 evaluated in double
                        }
                        outD = __fdiv_rn(a.mbf, disparity);
                        outU = bestuR;
                        outSad = (int)(best >> 4);
                    }
                }
            }
        }
    }
    if (lane == 0) {
        a.uRight[iL] = outU;
        a.depth[iL] = outD;
        a.sad[iL] = outSad;
    }
}

void launch_stereo_refine(const StereoRefineArgs& a, cudaStream_t s) {
    if (a.nL <= 0) return;
    stereo_refine_kernel<<<(a.nL + 7) / 8, 256, 0, s>>>(a);
}

}  // namespace rumi
