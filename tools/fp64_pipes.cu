// Microbenchmark: FP64 vector pipe (DFMA) vs FP64 tensor path (DMMA m8n8k4) on sm_100a, alone and mixed.
// Decides how the fused kernel's Gram update should be issued (SURVEY.md §7 "Gram on tensor cores").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_pipes tools/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void __launch_bounds__(256) mix_kernel(int iters, double *sink) {
    double v[NF > 0 ? NF : 1], c[NM > 0 ? 2 * NM : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) v[i] = 1.0 + threadIdx.x + i;
    for (int i = 0; i < (NM > 0 ? 2 * NM : 1); ++i) c[i] = 0.5 * i;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    const double m = 1.0000000001, ad = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NF; ++i) v[i] = fma(v[i], m, ad);
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += v[i];
    for (int i = 0; i < (NM > 0 ? 2 * NM : 1); ++i) s += c[i];
    if (s == 1234.5) sink[0] = s;
}

template <int NF, int NM>
void run(const char *name, int sms) {
    double *sink;
    cudaMalloc(&sink, 8);
    int iters = 2048, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        mix_kernel<NF, NM><<<blocks, threads>>>(iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best) best = ms;
    }
    double thr = (double)blocks * threads, it = (double)iters * 4;
    double f_dfma = 2.0 * NF * it * thr;                       // FMA = 2 flops per lane
    double f_dmma = 2.0 * 8 * 8 * 4 * NM * it * (thr / 32);    // m8n8k4 per warp
    printf("%-28s %8.3f ms  DFMA %7.2f TF  DMMA %7.2f TF  total %7.2f TF  (%s)\n", name, best, f_dfma / best / 1e9,
           f_dmma / best / 1e9, (f_dfma + f_dmma) / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, sm_%d%d, clock %d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate);
    run<8, 0>("DFMA x8 chains", p.multiProcessorCount);
    run<16, 0>("DFMA x16 chains", p.multiProcessorCount);
    run<0, 4>("DMMA x4 accumulators", p.multiProcessorCount);
    run<0, 8>("DMMA x8 accumulators", p.multiProcessorCount);
    run<0, 16>("DMMA x16 accumulators", p.multiProcessorCount);
    run<8, 4>("DFMA x8 + DMMA x4", p.multiProcessorCount);
    run<8, 8>("DFMA x8 + DMMA x8", p.multiProcessorCount);
    run<16, 2>("DFMA x16 + DMMA x2", p.multiProcessorCount);
    return 0;
}
