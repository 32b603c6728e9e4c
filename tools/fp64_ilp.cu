// FP64 issue rate of ONE / TWO / FOUR warps per SM sub-partition as a function of the independent chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(int iters, double *sink) {
    double v[ILP];
    for (int i = 0; i < ILP; ++i) v[i] = 1.0 + threadIdx.x * 1e-3 + i;
    const double m = 1.0000000001, ad = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], m, ad);
    }
    double s = 0; for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 1234.5) sink[0] = s;
}
template <int ILP> void run(int sms, int warps_per_sm) {
    double *sink; cudaMalloc(&sink, 8);
    int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<ILP><<<sms, warps_per_sm * 32>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    double instr_per_smsp = (double)warps_per_sm / 4 * iters * 4 * ILP;
    printf("warps/SMSP=%d ILP=%2d : %6.2f SMSP-cycles per DFMA  (dependent-issue distance %5.1f cycles)\n", warps_per_sm / 4, ILP,
           best * 1e-3 * 1.965e9 / instr_per_smsp, best * 1e-3 * 1.965e9 / (iters * 4.0));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16}) { run<1>(sms, w); run<2>(sms, w); run<4>(sms, w); run<8>(sms, w); run<16>(sms, w); run<24>(sms, w); }
    return 0;
}
