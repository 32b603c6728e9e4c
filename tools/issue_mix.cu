// Microbenchmark: do independent integer / LDS instructions issue "for free" in the shadow of DFMA on sm_100a?
// Each loop iteration issues 8 independent DFMA plus NI integer ops (and NL shared loads); if FP64 issue leaves
// every other slot free, time stays flat until NI+NL ~ 8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/issue_mix tools/issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int NL>
__global__ void __launch_bounds__(256) mix(int iters, double *sink, const unsigned *seed) {
    __shared__ double tab[256];
    tab[threadIdx.x] = threadIdx.x * 0.5;
    __syncthreads();
    double v[NF > 0 ? NF : 1];
    unsigned w[NI > 0 ? NI : 1];
    double l[NL > 0 ? NL : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) v[i] = 1.0 + threadIdx.x + i;
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) w[i] = seed[i] + threadIdx.x;
    for (int i = 0; i < (NL > 0 ? NL : 1); ++i) l[i] = 0.0;
    const double m = 1.0000000001, ad = 1e-9;
    unsigned idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NF; ++i) v[i] = fma(v[i], m, ad);
#pragma unroll
            for (int i = 0; i < NI; ++i) w[i] = w[i] * 0x9E3779B1u + 0x7F4A7C15u;      // IMAD
#pragma unroll
            for (int i = 0; i < NL; ++i) { l[i] += tab[(idx + i * 8) & 255]; }           // LDS + DADD (counted in NF-equivalent below)
            idx += 33;
        }
    }
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += v[i];
    for (int i = 0; i < (NI > 0 ? NI : 1); ++i) s += w[i];
    for (int i = 0; i < (NL > 0 ? NL : 1); ++i) s += l[i];
    if (s == 1234.5) sink[0] = s;
}

template <int NF, int NI, int NL>
void run(int sms, const unsigned *seed) {
    double *sink; cudaMalloc(&sink, 8);
    int iters = 2048, blocks = sms * 4, threads = 256;   // 8 warps per SMSP... 4 blocks x 8 warps = 32 warps/SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        mix<NF, NI, NL><<<blocks, threads>>>(iters, sink, seed);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    double warp_iters = (double)blocks * threads / 32 * iters * 4;
    double cyc = best * 1e-3 * 1.965e9 * sms * 4 / warp_iters;    // SMSP-cycles per unrolled step
    printf("DFMA=%d IMAD=%d LDS=%d : %8.3f ms  %6.2f SMSP-cycles per step (%d instr) -> IPC %.2f  (%s)\n", NF, NI, NL, best, cyc,
           NF + NI + 2 * NL + 0, (NF + NI + 2 * NL) / cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    unsigned *seed; cudaMalloc(&seed, 64 * 4); cudaMemset(seed, 1, 64 * 4);
    int sms = p.multiProcessorCount;
    printf("%s\n", p.name);
    run<8, 0, 0>(sms, seed);
    run<8, 2, 0>(sms, seed);
    run<8, 4, 0>(sms, seed);
    run<8, 8, 0>(sms, seed);
    run<8, 12, 0>(sms, seed);
    run<8, 16, 0>(sms, seed);
    run<0, 8, 0>(sms, seed);
    run<0, 16, 0>(sms, seed);
    run<8, 0, 2>(sms, seed);
    run<8, 0, 4>(sms, seed);
    run<8, 4, 2>(sms, seed);
    return 0;
}
