#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../../rumi_slam_b200/csrc/orb_math.cuh"
using namespace rumi;
__global__ void k(const int* d, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int dd[16];
    for (int j = 0; j < 16; ++j) dd[j] = d[i * 16 + j];
    out[i] = fast_score16(dd);
}
int main() {
    const int n = 4096;
    std::vector<int> h(n * 16), ref(n), got(n);
    srand(1);
    for (auto& v : h) v = rand() % 101 - 50;
    int first[16] = {14, 48, 46, 43, 47, 48, 47, 50, -17, -14, -6, -4, 1, 7, 8, 13};
    for (int j = 0; j < 16; ++j) h[j] = first[j];
    for (int i = 0; i < n; ++i) ref[i] = fast_score16(&h[i * 16]);
    int *dd, *dout;
    cudaMalloc(&dd, h.size() * 4); cudaMalloc(&dout, n * 4);
    cudaMemcpy(dd, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    k<<<(n + 127) / 128, 128>>>(dd, n, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got.data(), dout, n * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += got[i] != ref[i];
    printf("score probe: %s bad=%d first host=%d dev=%d\n", cudaGetErrorString(e), bad, ref[0], got[0]);
    return 0;
}
