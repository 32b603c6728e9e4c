// Microbenchmark: throughput of the in-order Halton digit-sum generation alone (k=20, two points per row),
// as a function of warps per SM and of the number of dimensions interleaved per digit loop (HG).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/halton_bench tools/halton_bench.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>

constexpr int K = 20;
__host__ __device__ constexpr uint32_t prime_at(int d) {
    constexpr uint32_t P[32] = {2,3,5,7,11,13,17,19,23,29,31,37,41,43,47,53,59,61,67,71,73,79,83,89,97,101,103,107,109,113,127,131};
    return P[d];
}
template <int N, class Fn, int... I> __device__ __forceinline__ void sfi(Fn &&fn, std::integer_sequence<int, I...>) { (fn(std::integral_constant<int, I>{}), ...); }
template <int N, class Fn> __device__ __forceinline__ void static_for(Fn &&fn) { sfi<N>(fn, std::make_integer_sequence<int, N>{}); }

struct FC { uint32_t toff[K], nd[K]; };

template <int HG, int MODE>
__global__ void hbench(const double *__restrict__ gterms, uint32_t nterms, FC fc, const uint32_t *__restrict__ perm, uint32_t n,
                       uint32_t start, double *out) {
    extern __shared__ double terms[];
    for (uint32_t e = threadIdx.x; e < nterms; e += blockDim.x) terms[e] = gterms[e];
    __syncthreads();
    double acc = 0.0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t ia = start + i, ib = start + n + perm[i];
        acc += (double)__brev(ia) * 2.3283064365386962890625e-10 + (double)__brev(ib) * 2.3283064365386962890625e-10;
        static_for<(K - 1 + HG - 1) / HG>([&](auto Gc) {
            constexpr int D0 = 1 + decltype(Gc)::value * HG;
            constexpr int N = (K - D0) < HG ? (K - D0) : HG;
            uint32_t ma[N], mb[N], off[N], offend[N];
            double xa[N], xb[N];
            int ndmax = 0;
#pragma unroll
            for (int u = 0; u < N; ++u) {
                ma[u] = ia; mb[u] = ib; xa[u] = 0.0; xb[u] = 0.0;
                off[u] = fc.toff[D0 + u];
                offend[u] = fc.toff[D0 + u] + (fc.nd[D0 + u] - 1) * prime_at(D0 + u);
                ndmax = ndmax > (int)fc.nd[D0 + u] ? ndmax : (int)fc.nd[D0 + u];
            }
            for (int j = 0; j < ndmax; ++j) {
                static_for<N>([&](auto Uc) {
                    constexpr int U = decltype(Uc)::value;
                    constexpr uint32_t base = prime_at(D0 + U);
                    uint32_t qa = ma[U] / base, qb = mb[U] / base;
                    if (MODE == 0) {           // real thing: table lookup + in-order add
                        const double *Tt = terms + off[U];
                        xa[U] = __dadd_rn(xa[U], Tt[ma[U] - qa * base]);
                        xb[U] = __dadd_rn(xb[U], Tt[mb[U] - qb * base]);
                    } else if (MODE == 1) {    // no table: integer part only (digit converted, added)
                        xa[U] = __dadd_rn(xa[U], (double)(int)(ma[U] - qa * base));
                        xb[U] = __dadd_rn(xb[U], (double)(int)(mb[U] - qb * base));
                    } else {                   // lookup at a conflict-free address (lane), same instruction count
                        const double *Tt = terms + (off[U] & 0) + (threadIdx.x & 31);
                        xa[U] = __dadd_rn(xa[U], Tt[(ma[U] - qa * base) & 0]);
                        xb[U] = __dadd_rn(xb[U], Tt[(mb[U] - qb * base) & 0]);
                    }
                    ma[U] = qa; mb[U] = qb;
                    off[U] = min(off[U] + base, offend[U]);
                });
            }
#pragma unroll
            for (int u = 0; u < N; ++u) acc += xa[u] + xb[u];
        });
    }
    if (acc == 1234.5678) out[0] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = acc;
}

template <int HG, int MODE>
void run(const char *name, int threads, const double *dterms, uint32_t nterms, const FC &fc, const uint32_t *dperm, uint32_t n, double *dout, int sms) {
    size_t smem = nterms * 8;
    cudaFuncSetAttribute(hbench<HG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        hbench<HG, MODE><<<sms, threads, smem>>>(dterms, nterms, fc, dperm, n, 401, dout);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 1 && ms < best) best = ms;
    }
    printf("%-34s HG=%d threads/SM=%4d  %8.3f ms  %7.2f Grows/s  (%s)\n", name, HG, threads, best, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const uint32_t n = 1u << 24;
    FC fc; std::vector<double> terms;
    uint64_t maxidx = 401ull + 2ull * n;
    for (int d = 0; d < K; ++d) {
        uint32_t b = prime_at(d), nd = 0;
        for (uint64_t m = maxidx; m > 0; m /= b) ++nd;
        fc.toff[d] = (uint32_t)terms.size(); fc.nd[d] = nd;
        double bp = b;
        for (uint32_t j = 0; j < nd; ++j) { for (uint32_t dg = 0; dg < b; ++dg) terms.push_back((double)dg / bp); bp *= b; }
    }
    std::vector<uint32_t> perm(n);
    uint64_t s = 88172645463325252ull;
    for (uint32_t i = 0; i < n; ++i) perm[i] = i;
    for (uint32_t i = n - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint32_t j = s % (i + 1); std::swap(perm[i], perm[j]); }
    double *dterms, *dout; uint32_t *dperm;
    cudaMalloc(&dterms, terms.size() * 8); cudaMemcpy(dterms, terms.data(), terms.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&dperm, n * 4); cudaMemcpy(dperm, perm.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 64);
    printf("%s, %d SMs; k=%d, n=2^24 rows, 2 points per row, %zu table terms\n", p.name, p.multiProcessorCount, K, terms.size());
    int sms = p.multiProcessorCount;
    for (int th : {128, 256, 512, 1024}) run<4, 0>("table lookup (real)", th, dterms, (uint32_t)terms.size(), fc, dperm, n, dout, sms);
    for (int th : {128, 256, 512, 1024}) run<2, 0>("table lookup (real)", th, dterms, (uint32_t)terms.size(), fc, dperm, n, dout, sms);
    for (int th : {128, 256, 512}) run<8, 0>("table lookup (real)", th, dterms, (uint32_t)terms.size(), fc, dperm, n, dout, sms);
    for (int th : {128, 512, 1024}) run<4, 1>("no table (int + I2F + DADD)", th, dterms, (uint32_t)terms.size(), fc, dperm, n, dout, sms);
    for (int th : {128, 512, 1024}) run<4, 2>("conflict-free lookup", th, dterms, (uint32_t)terms.size(), fc, dperm, n, dout, sms);
    return 0;
}
