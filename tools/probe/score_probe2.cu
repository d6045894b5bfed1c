#include <cstdio>
#include <vector>
#include <cstdlib>
#include <algorithm>
#include "../../rumi_slam_b200/csrc/orb_math.cuh"
using namespace rumi;
__host__ __device__ inline int naive(const int* d) {
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { int v = d[(k + j) & 15]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
        int s = mn > -mx ? mn : -mx;
        best = s > best ? s : best;
    }
    return best - 1;
}
__global__ void k1(const int* d, int n, int* out, int mode) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int dd[16];
    for (int j = 0; j < 16; ++j) dd[j] = d[i * 16 + j];
    if (mode == 0) out[i] = fast_score16(dd);
    else if (mode == 1) out[i] = naive(dd);
    else if (mode == 2) out[i] = min(dd[0], dd[1]) * 1000 + max(dd[0], dd[1]);
    else {
        int mn2 = dd[0] < dd[1] ? dd[0] : dd[1];
        int mx2 = dd[0] > dd[1] ? dd[0] : dd[1];
        out[i] = mn2 * 1000 + mx2;
    }
}
int main() {
    const int n = 1024;
    std::vector<int> h(n * 16), got(n);
    srand(1);
    for (auto& v : h) v = rand() % 101 - 50;
    int first[16] = {14, 48, 46, 43, 47, 48, 47, 50, -17, -14, -6, -4, 1, 7, 8, 13};
    for (int j = 0; j < 16; ++j) h[j] = first[j];
    int *dd, *dout;
    cudaMalloc(&dd, h.size() * 4); cudaMalloc(&dout, n * 4);
    cudaMemcpy(dd, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 4; ++mode) {
        k1<<<(n + 127) / 128, 128>>>(dd, n, dout, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(got.data(), dout, n * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < n; ++i) {
            int ref = mode == 0 ? fast_score16(&h[i * 16]) : mode == 1 ? naive(&h[i * 16])
                      : (std::min(h[i*16], h[i*16+1]) * 1000 + std::max(h[i*16], h[i*16+1]));
            bad += got[i] != ref;
        }
        printf("mode %d: %s bad=%d first dev=%d host fast=%d naive=%d\n", mode, cudaGetErrorString(e), bad, got[0],
               fast_score16(&h[0]), naive(&h[0]));
    }
    return 0;
}
