// Which instruction classes issue for free in the shadow of a saturated DFMA stream on sm_100a?
// 8 warps/SMSP, each loop step = 8 independent DFMA + N ops of one class.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int CLS, int N>
__global__ void __launch_bounds__(256) mix(int iters, double *sink, const unsigned *seed) {
    __shared__ unsigned tab[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) tab[i] = i * 7;
    __syncthreads();
    double v[NF > 0 ? NF : 1];
    unsigned w[N > 0 ? N : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) v[i] = 1.0 + threadIdx.x + i;
    for (int i = 0; i < (N > 0 ? N : 1); ++i) w[i] = seed[i] + threadIdx.x * 3 + i;
    const double m = 1.0000000001, ad = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < NF; ++i) v[i] = fma(v[i], m, ad);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (CLS == 0) w[i] = w[i] * 0x9E3779B1u + 0x7F4A7C15u;                 // IMAD (fma pipe)
                else if (CLS == 1) w[i] = __funnelshift_r(w[i], w[i], 7) ^ 0x5bd1e995u; // SHF + LOP3 (alu pipe) -> 2 instr
                else if (CLS == 2) w[i] = __umulhi(w[i], 0x51eb851fu) + 12345u;         // IMAD.HI
                else if (CLS == 3) w[i] = tab[w[i] & 1023];                              // dependent LDS.32 chain (+LOP)
                else if (CLS == 4) w[i] = min(w[i] + 17u, 0x7fffffffu) + (w[i] >> 31);   // IADD + VIMNMX...
            }
        }
    }
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += v[i];
    for (int i = 0; i < (N > 0 ? N : 1); ++i) s += w[i];
    if (s == 1234.5) sink[0] = s;
}
template <int NF, int CLS, int N> void run(const char *name, int sms, const unsigned *seed) {
    double *sink; cudaMalloc(&sink, 8);
    int iters = 1024, blocks = sms * 4, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); mix<NF, CLS, N><<<blocks, threads>>>(iters, sink, seed); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 2 && ms < best) best = ms;
    }
    double steps = (double)blocks * threads / 32 * iters * 4;
    printf("DFMA=%d + %2d x %-22s %8.3f ms  %6.2f SMSP-cycles per step\n", NF, N, name, best, best * 1e-3 * 1.965e9 * sms * 4 / steps);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    unsigned *seed; cudaMalloc(&seed, 256); cudaMemset(seed, 3, 256);
    run<8, 0, 0>("(nothing)", sms, seed);
    run<8, 0, 8>("IMAD", sms, seed);
    run<0, 0, 8>("IMAD", sms, seed);
    run<8, 1, 4>("SHF+LOP3", sms, seed);
    run<8, 1, 8>("SHF+LOP3", sms, seed);
    run<0, 1, 8>("SHF+LOP3", sms, seed);
    run<8, 2, 8>("IMAD.HI+IADD", sms, seed);
    run<0, 2, 8>("IMAD.HI+IADD", sms, seed);
    run<8, 3, 4>("LOP+LDS.32 (dependent)", sms, seed);
    run<8, 3, 8>("LOP+LDS.32 (dependent)", sms, seed);
    run<0, 3, 8>("LOP+LDS.32 (dependent)", sms, seed);
    run<8, 4, 8>("IADD+IMNMX", sms, seed);
    run<0, 4, 8>("IADD+IMNMX", sms, seed);
    return 0;
}
