// Latency of one dependent chain of Fr inversions on a single warp (nvcc -arch=sm_100a -I../../shielded_pool_pinocchio_solana_b200/csrc).
#include <cstdio>
#include "ff.cuh"
using namespace g16;
template <int KIND>
__global__ void k_chain(Fr* x, int n) {
    Fr a = x[threadIdx.x + blockIdx.x * blockDim.x];
    Fr one = Fr::one();
    for (int i = 0; i < n; i++) {
        a = (KIND == 0 ? a.inverse() : KIND == 1 ? a.inverse_euclid() : KIND == 2 ? a.inverse_fermat() : a * a) + one;
    }
    x[threadIdx.x + blockIdx.x * blockDim.x] = a;
}
template <int KIND>
float timeit(Fr* d, int n, int blocks, int threads) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<KIND><<<blocks, threads>>>(d, 2);
    cudaEventRecord(e0);
    k_chain<KIND><<<blocks, threads>>>(d, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    Fr h[1024];
    for (int i = 0; i < 1024; i++) { h[i] = Fr::zero(); for (int l = 0; l < 7; l++) h[i].v[l] = 0x9e3779b9u * (i + 3) * (l + 1) + l; h[i].v[7] = 0x0fffffff & (i * 2654435761u); }
    Fr* d; cudaMalloc(&d, sizeof(h));
    const int n = 200;
    for (int threads : {1, 32}) {
        cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
        printf("threads=%d  new %.2f us  euclid %.2f us  fermat %.2f us  modmul %.3f us per op\n", threads,
               timeit<0>(d, n, 1, threads) * 1e3 / n, timeit<1>(d, n, 1, threads) * 1e3 / n, timeit<2>(d, n, 1, threads) * 1e3 / n,
               timeit<3>(d, n * 100, 1, threads) * 1e3 / (n * 100));
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
