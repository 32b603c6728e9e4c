// Microbenchmark: issue cost of DADD / DMUL / DFMA (and of a mix) on sm_100a, in SMSP-cycles per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(int iters, double *sink) {
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 1.0000000001, ad = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) v[i] = fma(v[i], m, ad);
                else if (OP == 1) v[i] = __dadd_rn(v[i], ad);
                else if (OP == 2) v[i] = __dmul_rn(v[i], m);
                else if (OP == 3) v[i] = (i & 1) ? __dadd_rn(v[i], ad) : fma(v[i], m, ad);
                else if (OP == 4) v[i] = __dadd_rn(fabs(v[i]), ad);
                else v[i] = (i % 3 == 0) ? __dadd_rn(fabs(v[i]), ad) : ((i % 3 == 1) ? fma(v[i], m, ad) : __dmul_rn(v[i], m));
            }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 1234.5) sink[0] = s;
}
template <int OP> void run(const char *name, int sms) {
    double *sink; cudaMalloc(&sink, 8);
    int iters = 1024, blocks = sms * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 2 && ms < best) best = ms;
    }
    double winstr = (double)blocks * threads / 32 * iters * 64;
    printf("%-22s %8.3f ms  %5.2f SMSP-cycles per warp instruction\n", name, best, best * 1e-3 * 1.965e9 * sms * 4 / winstr);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    run<0>("DFMA", sms); run<1>("DADD", sms); run<2>("DMUL", sms); run<3>("DADD/DFMA alternating", sms); run<4>("DADD |x|", sms); run<5>("DADD/DFMA/DMUL mix", sms);
    return 0;
}
