// K4 IC_Angle orientation (R/lib_src/ORBextractor.cc:73-97, :463-469), K6 rotated-BRIEF descriptors (:99-143),
// K7 output assembly (level post-processing :816-825, coordinate scaling + lapping-area ordering :1069-1087).
//
// assign_slots: one CTA per frame turns the reference's sequential front/back fill (monoIndex++ / stereoIndex--)
// into two prefix sums over the level-ordered keypoints.
// describe: one WARP per keypoint.  Lanes first cooperate on the 749-pixel circular moment (lane = column u,
// int32 accumulation, shuffle reduction), then lane i produces descriptor byte i from 16 rotated samples of the
// blurred level.  Float32 steps use explicit round-to-nearest mul/add (no FMA) and glibc's sinf/cosf algorithm so
// that sample coordinates round exactly as on the reference's x86-64 build.
#include "kernels.cuh"
#include "slots.cuh"
#include "orb_math.cuh"

namespace rumi {

constexpr unsigned kFull = 0xFFFFFFFFu;

// rBRIEF sampling pattern (bit_pattern_31_, R/lib_src/ORBextractor.cc:145-403) as int8 (x0,y0,x1,y1) per bit.
// Lanes index it with 32 different addresses per step, which would serialise on the constant cache, so it lives
// in global memory behind the read-only L1 path instead of __constant__.
__device__ const int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};
// (A float32 copy of the table -- no int -> float conversions -- was measured SLOWER, 110 vs 70 us per chunk: 32 lanes x
// 16 bytes are four L1 wavefronts per load.)

// cvRound(v) for |v| < 2^22: adding 1.5 * 2^23 leaves the integer, rounded half to even like cvRound / rintf, in the low
// mantissa bits -- a full-rate FADD + IADD instead of a quarter-rate F2I
__device__ __forceinline__ int round_half_even(float v) {
    return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

__global__ void __launch_bounds__(256) assign_slots_kernel(const __grid_constant__ SlotArgs a,
                                                           const __grid_constant__ OrbConst oc) {
    assign_slots_frame(a, oc, blockIdx.x, threadIdx.x);
}

void launch_assign_slots(const DescribeArgs& a, const OrbConst& oc, cudaStream_t s) {
    SlotArgs sa;
    sa.sel = a.sel; sa.selCount = a.selCount; sa.lap0 = a.lap0; sa.lap1 = a.lap1; sa.slot = a.slot; sa.nkp = a.nkp; sa.nmono = a.nmono;
    assign_slots_kernel<<<a.cv.nframes, 256, 0, s>>>(sa, oc);
}

// 512 rotated samples -> 32 bytes; lane = byte index.  `center` points at the keypoint in the (blurred) image.
__device__ __forceinline__ uint8_t brief_byte(const uint8_t* center, int step, float ca, float sb, int lane) {
    const int8_t* pt = d_pattern + lane * 32;
    int val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const char4 q = *reinterpret_cast<const char4*>(pt + 4 * j);
        const float x0 = (float)q.x, y0 = (float)q.y, x1 = (float)q.z, y1 = (float)q.w;
        // center[cvRound(x*b + y*a) * step + cvRound(x*a - y*b)]   (:109-110)
        const int r0 = round_half_even(__fadd_rn(__fmul_rn(x0, sb), __fmul_rn(y0, ca)));
        const int c0 = round_half_even(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sb)));
        const int r1 = round_half_even(__fadd_rn(__fmul_rn(x1, sb), __fmul_rn(y1, ca)));
        const int c1 = round_half_even(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sb)));
        const int t0 = center[r0 * step + c0], t1 = center[r1 * step + c1];
        val |= (t0 < t1) << j;
    }
    return (uint8_t)val;
}

// The 512 samples of a keypoint fall into a 37 x 37 window (|pattern| <= 13, rotated: radius <= 18.4).  Gathering them
// straight from global memory costs one L1 tag look-up per distinct line and lane; instead the warp copies the window
// with 16-byte loads (37 rows x 4 aligned chunks) into shared memory and samples it there (+5 % on the pipelined batch).
// (The moment patch of IC_Angle was also tried from shared memory with DP4A row sums, lane = row: 31 rows x 3 chunks staged,
// 16 DP4A per lane instead of 30 predicated byte loads -- measured slower, 86-91 vs 76 us: the extra registers and shared
// memory halve the resident warps of a latency-bound kernel.  The simple form stays.)
constexpr int kPatchR = 18, kPatchRows = 2 * kPatchR + 1, kPatchPitch = 64;

__global__ void __launch_bounds__(256) describe_kernel(const __grid_constant__ DescribeArgs a,
                                                       const __grid_constant__ OrbConst oc) {
    __shared__ __align__(16) uint8_t s_patch[8][kPatchRows * kPatchPitch];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;                 // level-ordered keypoint slot
    if (i >= oc.kpCap) return;
    int l = 0;
    while (l + 1 < oc.nlevels && i >= oc.lv[l + 1].kpBase) ++l;
    const LevelGeom& g = oc.lv[l];
    if (i - g.kpBase >= a.selCount[(long long)f * oc.nlevels + l]) return;

    const uint32_t c = a.sel[(long long)f * oc.kpCap + i];
    const int x = cand_x(c) + kMinBorder, y = cand_y(c) + kMinBorder;      // pt += minBorder  (:821-822)

    // ---- IC_Angle on the un-blurred level: lane = u + 15 ----
    const LevelView sv = a.cv.src[l];
    const uint8_t* ctr = sv.ptr + (long long)f * sv.pitch + (long long)y * sv.stride + x;
    const int u = lane - kHalfPatch;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        m10 = u * ctr[u];
        const int au = u < 0 ? -u : u;
#pragma unroll
        for (int v = 1; v <= kHalfPatch; ++v) {
            if (au <= oc.umax[v]) {
                const int p = ctr[u + v * sv.stride], m = ctr[u - v * sv.stride];
                m01 += v * (p - m);
                m10 += u * (p + m);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(kFull, m10, o);
        m01 += __shfl_xor_sync(kFull, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- rBRIEF on the blurred level ----
    const float rad = __fmul_rn(angle, 0.017453292519943295f);           // factorPI = (float)(CV_PI/180.f)  (:99)
    float sb, ca;
    glibc_sincosf(rad, &sb, &ca);
    const LevelView bv = a.cv.blur[l];
    const int xa = (x - kPatchR) & ~15;                                   // keypoints sit >= 19 px inside the level
    {
        const uint8_t* src = bv.ptr + (long long)f * bv.pitch + (long long)(y - kPatchR) * bv.stride + xa;
        uint8_t* dst = s_patch[warp];
        for (int i = lane; i < kPatchRows * 4; i += 32) {
            const int r = i >> 2, c = i & 3;
            *reinterpret_cast<uint4*>(dst + r * kPatchPitch + 16 * c) =
                __ldg(reinterpret_cast<const uint4*>(src + (long long)r * bv.stride + 16 * c));
        }
    }
    __syncwarp();
    const uint8_t byte = brief_byte(s_patch[warp] + kPatchR * kPatchPitch + (x - xa), kPatchPitch, ca, sb, lane);

    const int slot = a.slot[(long long)f * oc.kpCap + i];
    if (slot >= a.outCap) return;
    a.desc[((long long)f * a.outCap + slot) * 32 + lane] = byte;
    if (lane == 0) {
        KeyPointRec k;
        k.x = (float)x; k.y = (float)y;
        if (l != 0) { k.x = __fmul_rn(k.x, g.scale); k.y = __fmul_rn(k.y, g.scale); }
        k.size = g.patchSize;
        k.angle = angle;
        k.response = (float)cand_resp(c);
        k.octave = l;
        k.class_id = -1;
        a.kps[(long long)f * a.outCap + slot] = k;
    }
}

void launch_describe(const DescribeArgs& a, const OrbConst& oc, cudaStream_t s) {
    describe_kernel<<<dim3((oc.kpCap + 7) / 8, a.cv.nframes), 256, 0, s>>>(a, oc);
}

// CloudFrameComputeDescriptors (R/lib_src/ORBextractor.cc:989-1011): caller's keypoints (pt, angle as given),
// caller's image as is -- no pyramid, no blur, no orientation.  Batched over images of one shape: keypoint i belongs
// to the image f with kpOff[f] <= i < kpOff[f + 1] (nimg == 1: kpOff may be null).
__global__ void __launch_bounds__(256) describe_given_kernel(const uint8_t* imgs, int w, int h, int stride, long long pitch,
                                                             int nimg, const int* __restrict__ kpOff,
                                                             const KeyPointRec* kps, int n, uint8_t* desc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    if (i >= n) return;
    int f = 0;
    if (nimg > 1) {                                          // largest f with kpOff[f] <= i
        int lo = 0, hi = nimg;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (kpOff[mid] <= i) lo = mid; else hi = mid;
        }
        f = lo;
    }
    const KeyPointRec k = kps[i];
    const int x = __float2int_rn(k.x), y = __float2int_rn(k.y);
    const float rad = __fmul_rn(k.angle, 0.017453292519943295f);
    float sb, ca;
    glibc_sincosf(rad, &sb, &ca);
    desc[(long long)i * 32 + lane] = brief_byte(imgs + f * pitch + (long long)y * stride + x, stride, ca, sb, lane);
}

void launch_describe_given(const uint8_t* imgs, int w, int h, int stride, long long pitch, int nimg, const int* kpOff,
                           const KeyPointRec* kps, int n, uint8_t* desc, cudaStream_t s) {
    if (n <= 0) return;
    describe_given_kernel<<<(n + 7) / 8, 256, 0, s>>>(imgs, w, h, stride, pitch, nimg, kpOff, kps, n, desc);
}

}  // namespace rumi
