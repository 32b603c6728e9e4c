// Generators and export-mode sample assembly.
#include "device.cuh"

namespace vs {

// ---------------------------------------------------------------------------------------------
// K1: Halton points, row-major (count, k).  One element per thread; consecutive threads write
// consecutive doubles.  Replaces varsens/saltelli.py:82-84 (+ :92/:95 scaling).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) halton_kernel(int k, uint64_t first, uint64_t total, HaltonDev h, ScaleDev s,
                                                     double *__restrict__ out) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        uint64_t row = e / (uint32_t)k;
        int d = (int)(e - row * (uint32_t)k);
        double p = halton_coord(h, d, (uint32_t)(first + row));
        out[e] = apply_scale(s, d, p);
    }
}

int launch_halton(vs_ctx *c, int k, uint64_t first, uint64_t count, const HaltonDev &h, const ScaleDev &s, double *out) {
    uint64_t total = count * (uint64_t)k;
    if (total == 0) return VS_OK;
    uint64_t want = (total + 255) / 256;
    int grid = (int)(want < (uint64_t)c->sm_count * 32 ? want : (uint64_t)c->sm_count * 32);
    halton_kernel<<<grid, 256, 0, c->stream>>>(k, first, total, h, s, out);
    c->launches++;
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// ---------------------------------------------------------------------------------------------
// K2: Gray-code Sobol by direct indexing (skip-ahead): point m = XOR of V[d][b] over the set bits
// b of m ^ (m >> 1); value = x * 2^-32.  Replaces quantlib/sobolGen.cpp:47-63.
// quantize6: the reference pipes the points through `cout << double` (6 significant digits) and
// numpy.loadtxt; that round trip is reproduced exactly in integer arithmetic:
//   x = X / 2^32 -> decimal D * 10^-p with D = round_half_even(X * 10^p / 2^32) in [10^5, 10^6]
//   -> strtod = correctly rounded D / 10^p (one IEEE division, both operands exact).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double quantize_6sig(uint32_t X) {
    if (X == 0u) return 0.0;
    // p = number of decimals so that X*10^p/2^32 has 6 integer digits: x in [10^-e, 10^-(e-1)) -> p = 5 + e
    // 10^p * X < 10^(p) * 2^32; p <= 15 keeps the product below 2^96 -> use 128-bit via two 64-bit halves.
    const double x = (double)X * 2.3283064365386962890625e-10;
    int e = 1;                       // x in [0.1, 1) -> e = 1
    double lim = 0.1;
    while (x < lim && e < 11) { lim *= 0.1; ++e; }   // coarse; corrected below with exact integers
    unsigned __int128 num, den = ((unsigned __int128)1) << 32;
    uint64_t D;
    int p = 5 + e;
    for (;;) {
        unsigned __int128 pw = 1;
        for (int i = 0; i < p; ++i) pw *= 10;
        num = (unsigned __int128)X * pw;
        unsigned __int128 q = num / den, r = num - q * den;
        D = (uint64_t)q;
        unsigned __int128 half = den >> 1;
        if (r > half || (r == half && (D & 1ull))) ++D;
        if (D < 100000ull) { ++p; continue; }        // estimate of e was one too small
        if (D > 1000000ull) { --p; continue; }
        break;                                       // D == 10^6 is fine: prints as 1 followed by zeros
    }
    double pw10 = 1.0;
    for (int i = 0; i < p; ++i) pw10 *= 10.0;        // exact for p <= 22
    return __ddiv_rn((double)D, pw10);
}

__global__ void __launch_bounds__(256) sobol_kernel(int k, uint64_t first, uint64_t total, const uint32_t *__restrict__ V,
                                                    int quantize6, ScaleDev s, double *__restrict__ out) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        uint64_t row = e / (uint32_t)k;
        int d = (int)(e - row * (uint32_t)k);
        uint32_t m = (uint32_t)(first + row);
        uint32_t g = m ^ (m >> 1), x = 0u;
        const uint32_t *v = V + (size_t)d * 32;
        while (g) {
            int b = __ffs(g) - 1;
            x ^= v[b];
            g &= g - 1;
        }
        double p = quantize6 ? quantize_6sig(x) : (double)x * 2.3283064365386962890625e-10;
        out[e] = apply_scale(s, d, p);
    }
}

int launch_sobol(vs_ctx *c, int k, uint64_t first, uint64_t count, const uint32_t *dir_dev, int quantize6, const ScaleDev &s,
                 double *out) {
    uint64_t total = count * (uint64_t)k;
    if (total == 0) return VS_OK;
    uint64_t want = (total + 255) / 256;
    int grid = (int)(want < (uint64_t)c->sm_count * 32 ? want : (uint64_t)c->sm_count * 32);
    sobol_kernel<<<grid, 256, 0, c->stream>>>(k, first, total, dir_dev, quantize6, s, out);
    c->launches++;
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: export mode.  A CTA owns TI consecutive base rows: it generates A (M_1 rows) and B (shuffled
// M_2 rows) ONCE into shared memory (2 * TI * k radical inverses) and then streams the 2+2k flat
// blocks that contain those rows -- each a contiguous TI*k-double run in HBM -- with the single
// substituted column patched on the fly.  HBM sees only the perm read and fully coalesced writes.
// Replaces varsens/saltelli.py:92-125 and :127-160.
//   flat row R = t*n + i:  t = 0: A_i | t = 1: B_i | t = 2+j: B_i with col j <- A_i[j]
//                          t = 2+k+j: A_i with col j <- B_i[j]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_flat_kernel(int k, int TI, SourceDev src, ScaleDev s, uint64_t i_lo, uint64_t i_hi, uint64_t row_begin,
                   uint64_t row_end, double *__restrict__ out) {
    extern __shared__ double smem[];
    double *A = smem;                       // [TI][k]
    double *B = smem + (size_t)TI * k;      // [TI][k]
    const uint64_t n = src.n;
    const uint64_t i0 = i_lo + (uint64_t)blockIdx.x * TI;
    const int rows = (int)((i0 + TI <= i_hi) ? TI : (i_hi - i0));
    const int cells = rows * k;
    for (int e = threadIdx.x; e < cells; e += blockDim.x) {
        int r = e / k, d = e - r * k;
        uint64_t i = i0 + r;
        A[e] = apply_scale(s, d, source_a(src, k, i, d));
        B[e] = apply_scale(s, d, source_b(src, k, src.perm[i], d));
    }
    __syncthreads();
    const int nblk = 2 + 2 * k;
    for (int t = 0; t < nblk; ++t) {
        uint64_t R0 = (uint64_t)t * n + i0;                 // first flat row of this run
        if (R0 + rows <= row_begin || R0 >= row_end) continue;
        const bool base_is_A = (t == 0) || (t >= 2 + k);
        const int j = (t < 2) ? -1 : (t < 2 + k ? t - 2 : t - 2 - k);
        const double *base = base_is_A ? A : B;
        const double *other = base_is_A ? B : A;
        int e_lo = (R0 < row_begin) ? (int)((row_begin - R0) * k) : 0;
        int e_hi = (R0 + rows > row_end) ? (int)((row_end - R0) * k) : cells;
        double *dst = out + (R0 - row_begin) * (uint64_t)k;  // may point before `out` when e_lo > 0; only [e_lo,e_hi) is touched
        for (int e = e_lo + threadIdx.x; e < e_hi; e += blockDim.x) {
            int d = e % k;
            double v = (d == j) ? other[e] : base[e];
            __stcs(dst + e, v);
        }
    }
}

int launch_sample_flat(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, uint64_t row_begin, uint64_t row_end,
                       double *out) {
    if (row_end <= row_begin) return VS_OK;
    const uint64_t n = src.n;
    // base rows touched by the window: if it spans a whole block, all of them; else the union of <= 2 arcs.
    uint64_t i_lo = 0, i_hi = n;
    uint64_t t_first = row_begin / n, t_last = (row_end - 1) / n;
    if (t_first == t_last) {
        i_lo = row_begin - t_first * n;
        i_hi = row_end - t_first * n;
    }
    // TI rows per CTA: as many as fit ~48 KB of shared memory, at most 32.
    size_t per_row = 2 * (size_t)k * sizeof(double);
    int TI = (int)(48 * 1024 / per_row);
    if (TI > 32) TI = 32;
    if (TI < 1) {
        TI = 1;
        VS_REQUIRE(per_row <= c->smem_optin, VS_ERR_UNSUPPORTED, "k=%d needs %zu bytes of shared memory per row", k, per_row);
    }
    size_t smem = per_row * TI;
    if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(sample_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t tiles = (i_hi - i_lo + TI - 1) / TI;
    VS_REQUIRE(tiles < (1ull << 31), VS_ERR_RANGE, "too many tiles");
    time_begin(c);
    sample_flat_kernel<<<(unsigned)tiles, 256, smem, c->stream>>>(k, TI, src, s, i_lo, i_hi, row_begin, row_end, out);
    time_end(c);
    c->launches++;
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vs
