// Cost of an LDS.64 whose 32 lanes read random entries of a small table (b doubles) vs consecutive entries.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void __launch_bounds__(1024) k(int iters, int b, int random, double *sink) {
    __shared__ double tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i * 1e-6;
    __syncthreads();
    uint32_t st[CHAINS];
    double acc[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { st[c] = (threadIdx.x * 2654435761u + c * 40503u) | 1u; acc[c] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            st[c] = st[c] * 1664525u + 1013904223u;                 // cheap LCG (1 IMAD)
            uint32_t idx = random ? (uint32_t)(((uint64_t)(st[c] >> 8) * (uint32_t)b) >> 24) : (threadIdx.x & 31) % b;
            acc[c] += tab[c * 128 + idx];
        }
    }
    double s = 0; for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 1234.5) sink[0] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    double *sink; cudaMalloc(&sink, 8);
    printf("%s: SM-cycles per warp-level LDS.64 (+IMAD, IMAD.HI-ish, DADD), 8 independent chains per thread\n", p.name);
    for (int threads : {256, 1024})
    for (int random = 0; random < 2; ++random)
    for (int b : {3, 7, 16, 31, 71, 128}) {
        int iters = 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<8><<<sms, threads>>>(iters, b, random, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
        double lds = (double)threads / 32 * iters * 8;          // warp-level LDS per SM
        printf("threads/SM=%4d %s b=%3d : %7.3f ms  %5.2f SM-cycles per LDS.64\n", threads, random ? "random     " : "consecutive", b, best, best * 1e-3 * 1.965e9 / lds);
    }
    return 0;
}
