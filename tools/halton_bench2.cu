// Latency anatomy of the fused kernels' Halton generation (new digit_step code), k=20, 8 warps/SM like the fused kernel.
// MODE 0 real | 1 table value replaced by a register constant (no LDS) | 2 LDS kept, result not chained (x = t) |
//      3 no DADD and no LDS (index chain only) | 4 real but perm[] read replaced by arithmetic
#include <cstdio>
#include <cstdint>
#include <vector>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>
constexpr int K = 20, HG = 4;
__host__ __device__ constexpr uint32_t prime_at(int d) {
    constexpr uint32_t P[32] = {2,3,5,7,11,13,17,19,23,29,31,37,41,43,47,53,59,61,67,71,73,79,83,89,97,101,103,107,109,113,127,131};
    return P[d];
}
template <int N, class Fn, int... I> __device__ __forceinline__ void sfi(Fn &&fn, std::integer_sequence<int, I...>) { (fn(std::integral_constant<int, I>{}), ...); }
template <int N, class Fn> __device__ __forceinline__ void static_for(Fn &&fn) { sfi<N>(fn, std::make_integer_sequence<int, N>{}); }
__host__ __device__ constexpr int ndmax32(uint32_t b) { int c = 0; for (uint64_t m = 0xFFFFFFFFull; m > 0; m /= b) ++c; return c; }
__host__ __device__ constexpr uint32_t foff(int d) { uint32_t s = 0; for (int e = 1; e < d; ++e) s += prime_at(e) * (uint32_t)ndmax32(prime_at(e)); return s; }
__host__ __device__ constexpr int jfast(uint32_t b) { int j = 1; for (uint64_t p = 1; p < 8; p *= b) ++j; return j; }

template <uint32_t B, bool FAST, int MODE>
__device__ __forceinline__ void digit_step(uint32_t &m, double &x, const char *__restrict__ row, double cst) {
    uint32_t off8;
    if constexpr (FAST) {
        constexpr uint32_t C = (uint32_t)((0x100000000ull + B - 1) / B);
        const uint64_t w = (uint64_t)m * C;
        off8 = __umulhi((uint32_t)w, 8u * B);
        m = (uint32_t)(w >> 32);
    } else {
        const uint32_t q = m / B;
        off8 = (m - q * B) * 8u;
        m = q;
    }
    if (MODE == 0 || MODE == 4) x = __dadd_rn(x, *reinterpret_cast<const double *>(row + off8));
    else if (MODE == 1) x = __dadd_rn(x, cst + (double)0) , m ^= (off8 & 0);        // keep off8 alive cheaply
    else if (MODE == 2) x = *reinterpret_cast<const double *>(row + off8);
    else m += (off8 & 0);
}

template <int MODE, int JCAP>
__global__ void __launch_bounds__(256) hb(const double *__restrict__ g, const uint32_t *__restrict__ perm, uint32_t n, uint32_t start, double *out, int small) {
    extern __shared__ double terms[];
    for (uint32_t e = threadIdx.x; e < foff(K); e += blockDim.x) terms[e] = g[e];
    __syncthreads();
    double acc = 0.0;
    const char *tb = reinterpret_cast<const char *>(terms);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t ia = start + i, ib = start + n + (MODE == 4 ? (i * 2654435761u) % n : perm[i]);
        static_for<(K - 1 + HG - 1) / HG>([&](auto Gc) {
            constexpr int D0 = 1 + decltype(Gc)::value * HG;
            constexpr int N = (K - D0) < HG ? (K - D0) : HG;
            constexpr int JM0 = ndmax32(prime_at(D0));
            constexpr int JMAX = JM0 < JCAP ? JM0 : JCAP;
            uint32_t ma[N], mb[N]; double xa[N], xb[N];
#pragma unroll
            for (int u = 0; u < N; ++u) { ma[u] = ia; mb[u] = ib; xa[u] = 0.0; xb[u] = 0.0; }
            static_for<JMAX>([&](auto Jc) {
                constexpr int J = decltype(Jc)::value;
                static_for<N>([&](auto Uc) {
                    constexpr int U = decltype(Uc)::value;
                    constexpr uint32_t B = prime_at(D0 + U);
                    if constexpr (J < ndmax32(B)) {
                        const char *row = tb + (size_t)(foff(D0 + U) + J * B) * 8;
                        if constexpr (J == 0) { digit_step<B, false, MODE>(ma[U], xa[U], row, 0.5); digit_step<B, false, MODE>(mb[U], xb[U], row, 0.5); }
                        else if constexpr (J >= jfast(B)) { digit_step<B, true, MODE>(ma[U], xa[U], row, 0.5); digit_step<B, true, MODE>(mb[U], xb[U], row, 0.5); }
                        else if (small) { digit_step<B, true, MODE>(ma[U], xa[U], row, 0.5); digit_step<B, true, MODE>(mb[U], xb[U], row, 0.5); }
                        else { digit_step<B, false, MODE>(ma[U], xa[U], row, 0.5); digit_step<B, false, MODE>(mb[U], xb[U], row, 0.5); }
                    }
                });
            });
#pragma unroll
            for (int u = 0; u < N; ++u) acc += xa[u] + xb[u] + (double)(ma[u] + mb[u]);
        });
    }
    if (acc == 1234.5678) out[0] = acc;
}

template <int MODE, int JCAP>
void run(const char *name, int threads, const double *dterms, const uint32_t *dperm, uint32_t n, double *dout, int sms) {
    size_t smem = foff(K) * 8;
    cudaFuncSetAttribute(hb<MODE, JCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); hb<MODE, JCAP><<<sms, threads, smem>>>(dterms, dperm, n, 401, dout, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 1 && ms < best) best = ms;
    }
    double per_batch = best * 1e-3 * 1.965e9 / ((double)n / 32 / (sms * threads / 32));
    printf("%-44s JCAP=%2d thr/SM=%4d %8.3f ms  %7.0f cycles per 32-row batch per warp (%s)\n", name, JCAP, threads, best, per_batch, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const uint32_t n = 1u << 24;
    std::vector<double> terms(foff(K), 0.0);
    for (int d = 1; d < K; ++d) { uint32_t b = prime_at(d); double bp = b; for (int j = 0; j < ndmax32(b); ++j) { for (uint32_t g = 0; g < b; ++g) terms[foff(d) + j * b + g] = (double)g / bp; bp *= b; } }
    std::vector<uint32_t> perm(n); uint64_t s = 88172645463325252ull;
    for (uint32_t i = 0; i < n; ++i) perm[i] = i;
    for (uint32_t i = n - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint32_t j = s % (i + 1); std::swap(perm[i], perm[j]); }
    double *dterms, *dout; uint32_t *dperm;
    cudaMalloc(&dterms, terms.size() * 8); cudaMemcpy(dterms, terms.data(), terms.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&dperm, n * 4); cudaMemcpy(dperm, perm.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 64);
    int sms = p.multiProcessorCount;
    printf("%s  k=%d n=2^24, table %u doubles\n", p.name, K, foff(K));
    for (int th : {128, 256, 512, 1024}) run<0, 32>("real", th, dterms, dperm, n, dout, sms);
    run<0, 17>("real, digits capped at 17 (enough for 2^26)", 256, dterms, dperm, n, dout, sms);
    run<1, 32>("no LDS (constant term)", 256, dterms, dperm, n, dout, sms);
    run<2, 32>("LDS, no add chain", 256, dterms, dperm, n, dout, sms);
    run<3, 32>("index chain only", 256, dterms, dperm, n, dout, sms);
    run<4, 32>("real, no perm[] load", 256, dterms, dperm, n, dout, sms);
    run<1, 32>("no LDS (constant term)", 1024, dterms, dperm, n, dout, sms);
    run<3, 32>("index chain only", 1024, dterms, dperm, n, dout, sms);
    return 0;
}
