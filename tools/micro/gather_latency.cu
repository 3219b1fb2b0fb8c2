// Microbenchmark (not product code): latency of dependent random byte gathers on B200 as a function of
// the footprint they are spread over -- does the per-hop latency of the front-proportional kernels
// (2-5 us per dependent load, profiles/r02*) come from address translation?
//   nvcc -arch=sm_100a -O3 -o gather_latency gather_latency.cu && ./gather_latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_chase(const uint8_t* buf, unsigned long long mask, int hops, int ilp, unsigned long long* out) {
    unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long a[4];
    for (int k = 0; k < 4; ++k) a[k] = (tid * 0x9E3779B97F4A7C15ull + k * 0xD1B54A32D192ED03ull) & mask;
    unsigned long long acc = 0;
    for (int h = 0; h < hops; ++h) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < ilp) {
                const unsigned v = buf[a[k]];
                a[k] = (a[k] * 6364136223846793005ull + 1442695040888963407ull + v) & mask;  // depends on the loaded byte
                acc += v;
            }
        }
    }
    out[tid] = acc + a[0] + a[1] + a[2] + a[3];
}

int main() {
    const size_t maxb = 48ull << 30;
    uint8_t* buf;
    if (cudaMalloc(&buf, maxb) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, maxb);
    unsigned long long* out;
    cudaMalloc(&out, 148 * 2048 * 8 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int hops = 64;
    printf("footprint_MB threads_per_SM ilp us_per_hop Greq_per_s\n");
    for (int tps : {256, 1024, 2048})
        for (int ilp : {1, 4})
            for (size_t fp : {64ull << 20, 256ull << 20, 1ull << 30, 4ull << 30, 16ull << 30, 32ull << 30}) {
                const int blocks = 148 * tps / 256;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    k_chase<<<blocks, 256>>>(buf, fp - 1, hops, ilp, out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                printf("%8zu %5d %d %8.3f %8.2f\n", fp >> 20, tps, ilp, ms * 1e3 / hops, (double)blocks * 256 * ilp * hops / (ms * 1e-3) / 1e9);
            }
    return 0;
}
