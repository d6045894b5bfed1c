// K13-K15: sparse pyramidal Lucas-Kanade flow of KFDSample::Step (R/lib_src/KFDSample.cc:131-132:
// calcOpticalFlowPyrLK(imprvs, imnext, old, next, status, err, Size(31,31), 2, TermCriteria(COUNT+EPS, 20, 0.03))).
// OpenCV's algorithm (buildOpticalFlowPyramid / calcScharrDeriv / LKTrackerInvoker) re-designed for the device:
//   * K13 pyrDown: thread per destination pixel, 5x5 [1 4 6 4 1] taps with REFLECT_101 indices, (sum + 128) >> 8;
//   * K14 Scharr: thread per pixel, all levels in one launch, int16 (dx, dy) pairs;
//   * K15 tracker: ONE CTA OF 4 WARPS PER POINT walks ALL pyramid levels inside one launch (points are independent,
//     so the per-level launches of the CPU code collapse into one).  The kernel is a chain of dependent Newton steps,
//     i.e. latency bound: the window rows are split between the 4 warps (one per SM sub-partition) to shorten each
//     step (one warp per point: 95 us for 1000 points; 4 warps: see profiles/).  Lane = window column (win + 1 <= 32
//     columns: the extra lane supplies the right neighbour of the bilinear sample through a shuffle); the template
//     patch and its two derivative patches stay in REGISTERS (3 ints per row and lane, rows fully unrolled); every
//     Newton step is one byte load per row and lane, 4 IMAD for the fixed-point bilinear sample, 2 IMAD for the
//     mismatch vector, then one shuffle + shared-memory reduction with a single barrier.  The sums of the normal
//     equations are sums of integer products: accumulated exactly (int32 per lane, int64 across the CTA), i.e.
//     OpenCV's integer-accumulator variant, so the result does not depend on the reduction order and equals the
//     oracle bit for bit; the float32 tail uses explicit IEEE operations (--fmad=false, IEEE sqrt / division).
#include "kernels.cuh"

namespace rumi {

namespace {

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
flow_pyrdown_kernel(const uint8_t* __restrict__ src, int sw, int sh, int sstride, uint8_t* __restrict__ dst, int dw,
                    int dh, int dstride) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * x - 2 + k, sw);
    int rows[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint8_t* r = src + (size_t)reflect101(2 * y - 2 + k, sh) * sstride;
        rows[k] = r[xs[2]] * 6 + (r[xs[1]] + r[xs[3]]) * 4 + r[xs[0]] + r[xs[4]];
    }
    const int v = rows[2] * 6 + (rows[1] + rows[3]) * 4 + rows[0] + rows[4];
    dst[(size_t)y * dstride + x] = (uint8_t)((v + 128) >> 8);
}

__global__ void __launch_bounds__(256) flow_scharr_kernel(FlowPyramidView img, FlowDerivView der) {
    const int level = blockIdx.z;
    const int w = img.w[level], h = img.h[level], stride = img.stride[level];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const uint8_t* base = img.ptr[level];
    const uint8_t* r0 = base + (size_t)(y > 0 ? y - 1 : h > 1 ? 1 : 0) * stride;
    const uint8_t* r1 = base + (size_t)y * stride;
    const uint8_t* r2 = base + (size_t)(y < h - 1 ? y + 1 : h > 1 ? h - 2 : 0) * stride;
    const int xl = x > 0 ? x - 1 : w > 1 ? 1 : 0, xr = x < w - 1 ? x + 1 : w > 1 ? w - 2 : 0;
    const int t0l = (r0[xl] + r2[xl]) * 3 + r1[xl] * 10, t0r = (r0[xr] + r2[xr]) * 3 + r1[xr] * 10;
    const int t1l = r2[xl] - r0[xl], t1c = r2[x] - r0[x], t1r = r2[xr] - r0[xr];
    der.ptr[level][(size_t)y * w + x] = make_short2((short)(t0r - t0l), (short)((t1r + t1l) * 3 + t1c * 10));
}

struct Weights { int w00, w01, w10, w11; };
__device__ __forceinline__ Weights lk_weights(float a, float b) {
    const float oa = __fsub_rn(1.f, a), ob = __fsub_rn(1.f, b);
    Weights k;
    k.w00 = __float2int_rn(__fmul_rn(__fmul_rn(oa, ob), 16384.f));
    k.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ob), 16384.f));
    k.w10 = __float2int_rn(__fmul_rn(__fmul_rn(oa, b), 16384.f));
    k.w11 = 16384 - k.w00 - k.w01 - k.w10;
    return k;
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

constexpr int kLkWarps = 4;                       // warps per point: the window rows are split between them

// Exact sum of up to three per-lane integers over the CTA (4 warps).  Two shared-memory buffers used alternately
// (`phase`), so one barrier per reduction is enough.
template <int N>
__device__ __forceinline__ void block_sum(long long (&v)[N], long long (*red)[kLkWarps][3], int& phase, int warp,
                                          int lane) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    long long(*buf)[3] = red[phase & 1];
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) buf[warp][i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = buf[0][i] + buf[1][i] + buf[2][i] + buf[3][i];
    ++phase;
}

// Mismatch of this warp's rows [row0, row0 + RPW) of the WIN x WIN window of J at (ix, iy) + fractional weights
// against the template rows in registers.  MODE 0: b-vector (sum diff * Ix, sum diff * Iy); MODE 1: sum |diff|.
template <int WIN, int RPW, int MODE>
__device__ __forceinline__ void lk_rows(const uint8_t* __restrict__ J, int w, int h, int stride, int ix, int iy,
                                        const Weights k, int lane, int row0, const int (&Iw)[RPW],
                                        const int (&Ix)[RPW], const int (&Iy)[RPW], int& a1, int& a2) {
    const int xr = reflect101(ix + lane, w);
    int prevH = 0;
    a1 = 0; a2 = 0;
#pragma unroll
    for (int r = 0; r <= RPW; ++r) {
        const int c = J[(size_t)reflect101(iy + row0 + r, h) * stride + xr];
        const int cr = __shfl_down_sync(0xFFFFFFFFu, c, 1);
        if (r > 0 && row0 + r - 1 < WIN) {
            const int diff = ((prevH + c * k.w10 + cr * k.w11 + (1 << 8)) >> 9) - Iw[r - 1];
            if (MODE == 0) { a1 += diff * Ix[r - 1]; a2 += diff * Iy[r - 1]; }
            else a1 += abs(diff);
        }
        prevH = c * k.w00 + cr * k.w01;
    }
    if (lane >= WIN) { a1 = 0; a2 = 0; }
}

template <int WIN>
__global__ void __launch_bounds__(32 * kLkWarps) flow_lk_kernel(FlowTrackArgs a) {
    constexpr int RPW = (WIN + kLkWarps - 1) / kLkWarps;
    __shared__ long long red[2][kLkWarps][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x;
    const int row0 = warp * RPW;
    int phase = 0;
    const float half = (WIN - 1) * 0.5f;
    const float2 pp = a.prevPts[p];
    float outx = 0.f, outy = 0.f, errv = 0.f;
    int st = 1;
    int Iw[RPW], Ix[RPW], Iy[RPW];
    for (int level = a.maxLevel; level >= 0; --level) {
        const int w = a.I.w[level], h = a.I.h[level];
        const float sc = (float)(1. / (1 << level));
        float px = __fmul_rn(pp.x, sc), py = __fmul_rn(pp.y, sc);
        float nx, ny;
        if (level == a.maxLevel) { nx = px; ny = py; }
        else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;
        px = __fsub_rn(px, half); py = __fsub_rn(py, half);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) {
            if (level == 0) { st = 0; errv = 0.f; }
            continue;
        }
        Weights k = lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy));
        long long sA[3];
        {
            int s11 = 0, s12 = 0, s22 = 0;
            const uint8_t* I = a.I.ptr[level];
            const short2* D = a.D.ptr[level];
            const int stride = a.I.stride[level];
            const int X = ipx + lane, xr = reflect101(X, w);
            const bool xin = X >= 0 && X < w;
            int pH = 0, pHx = 0, pHy = 0;
#pragma unroll
            for (int r = 0; r <= RPW; ++r) {
                const int Y = ipy + row0 + r;
                const int c = I[(size_t)reflect101(Y, h) * stride + xr];
                short2 d = make_short2(0, 0);
                if (xin && Y >= 0 && Y < h) d = D[(size_t)Y * w + X];
                const int cr = __shfl_down_sync(0xFFFFFFFFu, c, 1);
                const int dpk = __shfl_down_sync(0xFFFFFFFFu, ((int)d.y << 16) | ((int)d.x & 0xFFFF), 1);
                const int dxr = (int)(short)(dpk & 0xFFFF), dyr = dpk >> 16;
                if (r > 0) {
                    const int iv = (pH + c * k.w10 + cr * k.w11 + (1 << 8)) >> 9;
                    const int gx = (pHx + d.x * k.w10 + dxr * k.w11 + (1 << 13)) >> 14;
                    const int gy = (pHy + d.y * k.w10 + dyr * k.w11 + (1 << 13)) >> 14;
                    Iw[r - 1] = iv; Ix[r - 1] = gx; Iy[r - 1] = gy;
                    if (row0 + r - 1 < WIN) { s11 += gx * gx; s12 += gx * gy; s22 += gy * gy; }
                }
                pH = c * k.w00 + cr * k.w01;
                pHx = d.x * k.w00 + dxr * k.w01;
                pHy = d.y * k.w00 + dyr * k.w01;
            }
            if (lane >= WIN) { s11 = 0; s12 = 0; s22 = 0; }
            sA[0] = s11; sA[1] = s12; sA[2] = s22;
        }
        block_sum<3>(sA, red, phase, warp, lane);
        const float FLT_SCALE = 1.f / (1 << 20);
        const float A11 = __fmul_rn(__ll2float_rn(sA[0]), FLT_SCALE);
        const float A12 = __fmul_rn(__ll2float_rn(sA[1]), FLT_SCALE);
        const float A22 = __fmul_rn(__ll2float_rn(sA[2]), FLT_SCALE);
        float Dt = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WIN * WIN));
        if (minEig < a.minEig || Dt < 1.1920928955078125e-7f) {
            if (level == 0) st = 0;
            continue;
        }
        Dt = __fdiv_rn(1.f, Dt);
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        const uint8_t* J = a.J.ptr[level];
        const int jstride = a.J.stride[level];
        for (int j = 0; j < a.maxCount; ++j) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
                if (level == 0) st = 0;
                break;
            }
            k = lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny));
            int a1, a2;
            lk_rows<WIN, RPW, 0>(J, w, h, jstride, inx, iny, k, lane, row0, Iw, Ix, Iy, a1, a2);
            long long sb[2] = {a1, a2};
            block_sum<2>(sb, red, phase, warp, lane);
            const float b1 = __fmul_rn(__ll2float_rn(sb[0]), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(sb[1]), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), Dt);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), Dt);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= a.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (st && level == 0) {
            const float ex = __fsub_rn(outx, half), ey = __fsub_rn(outy, half);
            const int iex = (int)floorf(ex), iey = (int)floorf(ey);
            if (iex < -WIN || iex >= w || iey < -WIN || iey >= h) { st = 0; continue; }
            k = lk_weights(__fsub_rn(ex, (float)iex), __fsub_rn(ey, (float)iey));
            int a1, a2;
            lk_rows<WIN, RPW, 1>(J, w, h, jstride, iex, iey, k, lane, row0, Iw, Ix, Iy, a1, a2);
            long long se[1] = {a1};
            block_sum<1>(se, red, phase, warp, lane);
            errv = __fdiv_rn(__fmul_rn(__ll2float_rn(se[0]), 1.f), (float)(32 * WIN * WIN));
        }
    }
    if (threadIdx.x == 0) {
        a.nextPts[p] = make_float2(outx, outy);
        a.status[p] = (uint8_t)st;
        a.err[p] = errv;
    }
}

}  // namespace

void launch_flow_pyrdown(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dstride, cudaStream_t s) {
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    flow_pyrdown_kernel<<<dim3((dw + 31) / 32, (dh + 7) / 8), 256, 0, s>>>(src, sw, sh, sstride, dst, dw, dh, dstride);
}

void launch_flow_scharr(const FlowPyramidView& img, const FlowDerivView& der, int levels, cudaStream_t s) {
    flow_scharr_kernel<<<dim3((img.w[0] + 31) / 32, (img.h[0] + 7) / 8, levels), 256, 0, s>>>(img, der);
}

bool launch_flow_track(const FlowTrackArgs& a, int win, cudaStream_t s) {
    if (a.n <= 0) return true;
    const int grid = a.n, threads = 32 * kLkWarps;          // one CTA (4 warps) per point
    switch (win) {
        case 31: flow_lk_kernel<31><<<grid, threads, 0, s>>>(a); return true;
        case 21: flow_lk_kernel<21><<<grid, threads, 0, s>>>(a); return true;
        case 15: flow_lk_kernel<15><<<grid, threads, 0, s>>>(a); return true;
        default: return false;
    }
}

}  // namespace rumi
