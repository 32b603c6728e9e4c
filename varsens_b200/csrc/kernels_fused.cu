// Fused pipeline: generation + scaling + assembly + objective + estimator reductions in one kernel.
//
// Replaces varsens/saltelli.py:82-125 (Sample), :308-353 (Objective loops) and the reductions of
// :577-622 (Varsens.compute_varsens) for registered functors.  The sample matrices M_1, M_2, N_j,
// N_nj (2*k*n*k*8 bytes in the reference, saltelli.py:119) never exist: a base row i is two
// register vectors A_i, B_i and each of its 2+2k design points is a compile-time selection of
// those registers.
//
// Mapping (one warp = 32 consecutive base rows per batch, persistent grid):
//   phase 1  lane = base row.  A_i, B_i by in-order Halton digit sums (term table in shared
//            memory), scaled; the functor is evaluated on each of the 2+2k points and the value
//            is parked in the warp's shared tile Y[32][M].
//   phase 2  lane = T x T register tile of the symmetric M x M Gram  G += Y^T Y  (M = 2+2k): for
//            each of the 32 rows 2T broadcast shared loads feed T^2 DFMA.  G holds every sum the
//            estimators need.  Shifted sums for var_y stay lane-local in phase 1.
// Warps, then CTAs, are combined in a fixed order -> results are bit-reproducible run to run.
#include <type_traits>
#include <utility>

#include "device.cuh"

namespace vs {

constexpr int FUSED_WARPS = 8;

__host__ __device__ constexpr uint32_t prime_at(int d) {
    constexpr uint32_t P[32] = {2,  3,  5,  7,  11, 13, 17, 19, 23, 29, 31, 37,  41,  43,  47,  53,
                                59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103, 107, 109, 113, 127, 131};
    return P[d];
}

__host__ __device__ constexpr int gram_tile_for(int M) {
    int T = 1;
    while (((M + T - 1) / T) * ((M + T - 1) / T + 1) / 2 > 32) ++T;
    return T;
}

template <int K>
struct FusedConst {            // kernel parameter -> constant bank; indexed with compile-time subscripts
    double lb[K], wr[K];
    uint32_t toff[K];          // offset of dimension d's terms inside the shared copy of the table
    uint32_t nd[K];            // digits to sum for the largest index of the run
    int scale_kind;
};

// compile-time loop: fn(std::integral_constant<int, I>{}) for I in [0, N)
template <int N, class Fn, int... I>
__device__ __forceinline__ void static_for_impl(Fn &&fn, std::integer_sequence<int, I...>) {
    (fn(std::integral_constant<int, I>{}), ...);
}
template <int N, class Fn>
__device__ __forceinline__ void static_for(Fn &&fn) {
    static_for_impl<N>(fn, std::make_integer_sequence<int, N>{});
}

template <int N>
__device__ __forceinline__ double tree_product(double (&s)[N]) {
    if constexpr (N == 1) return s[0];
    else {
        constexpr int H = (N + 1) / 2;
        double t[H];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) t[i] = s[2 * i] * s[2 * i + 1];
        if constexpr (N % 2) t[H - 1] = s[N - 1];
        return tree_product<H>(t);
    }
}

// ---- register-resident functors --------------------------------------------------------------
// Every design point is handed to the functor as its own k-vector of registers plus an opaque
// scalar `tok` (value F::token, re-read from shared memory through a volatile load for every
// point).  The functor must fold tok into its first operation on each coordinate.  Without it the
// compiler notices that neighbouring points share k-1 coordinates and hoists the common
// sub-expressions -- i.e. silently applies the separable shortcut -- and the generic path would
// no longer evaluate each point (SURVEY.md §7 "honest flop accounting", §8d).  The token costs one
// LDS per point and no arithmetic.
//
// g-function with the division hoisted: prod_c (|4x_c-2| + a_c)/(1+a_c) = C * prod_c (|4x_c-2| + a_c),
// C = prod_c 1/(1+a_c).  Per factor: DFMA (tok*x - 2, tok = 4), DADD (|.| + a_c, a_c from the
// constant bank), DMUL (tree product).
template <int K>
struct GFunctionReg {
    double a[K];
    double C;
    static constexpr bool separable = true;
    static constexpr double token = 4.0;
    __device__ __forceinline__ double factor(int c, double x, double tok) const { return fabs(fma(tok, x, -2.0)) + a[c]; }
    __device__ __forceinline__ double finish(double p) const { return C * p; }
    __device__ __forceinline__ double operator()(const double (&x)[K], double tok) const {
        double s[K];
#pragma unroll
        for (int c = 0; c < K; ++c) s[c] = factor(c, x[c], tok);
        return C * tree_product<K>(s);
    }
};

template <int K>
struct IshigamiReg {
    double A, B;
    static constexpr bool separable = false;
    static constexpr double token = 1.0;
    __device__ __forceinline__ double operator()(const double (&x)[K], double tok) const {
        double s0 = sin(x[0] * tok), s1 = sin(x[1] * tok), x2 = x[2] * tok;
        double x22 = x2 * x2;
        return s0 + A * s1 * s1 + B * (x22 * x22) * s0;
    }
};

template <int K, class F, bool SECOND, bool SEPARABLE>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 1)
fused_kernel(SourceDev src, FusedConst<K> fc, F f, uint64_t i_begin, uint64_t i_end, const double *__restrict__ shift_ptr,
             double *__restrict__ blockpart) {
    constexpr int M = 2 + 2 * K;
    constexpr int T = SECOND ? gram_tile_for(M) : 2;
    constexpr int NT = (M + T - 1) / T;
    constexpr int MP = (NT * T) % 2 ? NT * T : NT * T + 1;     // odd row pitch
    constexpr int NTILES = SECOND ? NT * (NT + 1) / 2 : NT;    // first-order only: tile row 0 (fM_1, fM_2) x all columns
    static_assert(SECOND ? NTILES <= 32 : true, "Gram does not fit one tile per lane");
    constexpr int TPL = SECOND ? 1 : (NTILES + 31) / 32;

    extern __shared__ double smem[];
    double *terms = smem;                                           // shared copy of the Halton term table
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nterms = src.raw ? 0u : src.h.total_terms;
    volatile double *tokp = smem + nterms;                          // opaque functor token (see functors above)
    double *Y = smem + nterms + 1 + (size_t)warp * 32 * MP;         // this warp's [32][MP] value tile
    for (uint32_t e = threadIdx.x; e < nterms; e += blockDim.x) terms[e] = src.h.terms[e];
    if (threadIdx.x == 0) *tokp = F::token;
    // zero the padding columns once (never written again)
    for (int e = lane; e < 32 * (MP - M); e += 32) Y[(e / (MP - M)) * MP + M + e % (MP - M)] = 0.0;
    __syncthreads();

    const double shift = *shift_ptr;
    double sA = 0.0, qA = 0.0, sB = 0.0, qB = 0.0;
    double acc[TPL][T][T];
#pragma unroll
    for (int q = 0; q < TPL; ++q)
#pragma unroll
        for (int x = 0; x < T; ++x)
#pragma unroll
            for (int y = 0; y < T; ++y) acc[q][x][y] = 0.0;
    int tr[TPL], tc[TPL];
#pragma unroll
    for (int q = 0; q < TPL; ++q) {
        int id = lane + 32 * q;
        if (SECOND) tile_coords(id < NTILES ? id : 0, NT, tr[q], tc[q]);
        else { tr[q] = 0; tc[q] = id < NTILES ? id : 0; }
    }

    const uint64_t rows = i_end - i_begin;
    const uint64_t nbatch = (rows + 31) / 32;
    const uint64_t wstride = (uint64_t)gridDim.x * FUSED_WARPS;
    for (uint64_t bt = (uint64_t)blockIdx.x * FUSED_WARPS + warp; bt < nbatch; bt += wstride) {
        // ------------------------------ phase 1: lane = base row ------------------------------
        uint64_t r = bt * 32 + lane;
        const bool valid = r < rows;
        const uint64_t i = i_begin + (valid ? r : rows - 1);
        const uint64_t pi = src.perm[i];
        double a[K], b[K];
        if (src.raw) {
            const double *ra = src.raw + i * (uint64_t)K, *rb = src.raw + (src.n + pi) * (uint64_t)K;
#pragma unroll
            for (int d = 0; d < K; ++d) { a[d] = ra[d]; b[d] = rb[d]; }
        } else {
            const uint32_t ia = (uint32_t)(src.start + i), ib = (uint32_t)(src.start + src.n + pi);
            static_for<K>([&](auto Dc) {
                constexpr int D = decltype(Dc)::value;
                constexpr uint32_t base = prime_at(D);
                if constexpr (base == 2u) {
                    a[D] = (double)__brev(ia) * 2.3283064365386962890625e-10;
                    b[D] = (double)__brev(ib) * 2.3283064365386962890625e-10;
                } else {
                    const double *Tt = terms + fc.toff[D];
                    uint32_t ma = ia, mb = ib;
                    double xa = 0.0, xb = 0.0;
                    const int nd = (int)fc.nd[D];
                    for (int j = 0; j < nd; ++j) {            // least-significant digit first; 0-digits add +0.0 (exact)
                        uint32_t qa = ma / base, qb = mb / base;
                        xa = __dadd_rn(xa, Tt[ma - qa * base]);
                        xb = __dadd_rn(xb, Tt[mb - qb * base]);
                        ma = qa;
                        mb = qb;
                        Tt += base;
                    }
                    a[D] = xa;
                    b[D] = xb;
                }
            });
        }
        if (fc.scale_kind == VS_SCALE_LINEAR) {                     // scale.py:33, two roundings
#pragma unroll
            for (int d = 0; d < K; ++d) {
                a[d] = __dadd_rn(__dmul_rn(a[d], fc.wr[d]), fc.lb[d]);
                b[d] = __dadd_rn(__dmul_rn(b[d], fc.wr[d]), fc.lb[d]);
            }
        } else if (fc.scale_kind == VS_SCALE_POWER) {               // scale.py:62
#pragma unroll
            for (int d = 0; d < K; ++d) {
                a[d] = __dmul_rn(fc.lb[d], pow(fc.wr[d], a[d]));
                b[d] = __dmul_rn(fc.lb[d], pow(fc.wr[d], b[d]));
            }
        }
        double *Yrow = Y + lane * MP;
        double fA, fB;
        if constexpr (SEPARABLE && F::separable) {
            // product-form shortcut: all 2+2k values from prefix/suffix products of the 2k factors
            // (O(k) per row instead of O(k^2)).  Prefix pass parks prefix*factor in the tile, suffix
            // pass completes it in place -- no k-long register arrays besides the factors.
            double ga[K], gb[K];
#pragma unroll
            for (int c = 0; c < K; ++c) { ga[c] = f.factor(c, a[c], F::token); gb[c] = f.factor(c, b[c], F::token); }
            double pa = 1.0, pb = 1.0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                Yrow[2 + j] = pb * ga[j];                            // N_j[j]  = B with column j from A
                Yrow[2 + K + j] = pa * gb[j];                        // N_nj[j] = A with column j from B
                pa *= ga[j];
                pb *= gb[j];
            }
            fA = f.finish(pa);
            fB = f.finish(pb);
            double sa = f.finish(1.0), sb = f.finish(1.0);           // suffix products carry the constant C
#pragma unroll
            for (int j = K - 1; j >= 0; --j) {
                double vj = Yrow[2 + j] * sb, vn = Yrow[2 + K + j] * sa;
                Yrow[2 + j] = valid ? vj : 0.0;
                Yrow[2 + K + j] = valid ? vn : 0.0;
                sa *= ga[j];
                sb *= gb[j];
            }
        } else {
            fA = f(a, *tokp);
            fB = f(b, *tokp);
            static_for<K>([&](auto Jc) {
                constexpr int J = decltype(Jc)::value;
                double xj[K], xn[K];
#pragma unroll
                for (int c = 0; c < K; ++c) {
                    xj[c] = (c == J) ? a[c] : b[c];                  // N_j[J]   (saltelli.py:119-123)
                    xn[c] = (c == J) ? b[c] : a[c];                  // N_nj[J]
                }
                double vj = f(xj, *tokp), vn = f(xn, *tokp);
                Yrow[2 + J] = valid ? vj : 0.0;
                Yrow[2 + K + J] = valid ? vn : 0.0;
            });
        }
        Yrow[0] = valid ? fA : 0.0;
        Yrow[1] = valid ? fB : 0.0;
        if (valid) {
            double dA = fA - shift, dB = fB - shift;
            sA += dA;
            qA = fma(dA, dA, qA);
            sB += dB;
            qB = fma(dB, dB, qB);
        }
        __syncwarp();
        // ------------------------------ phase 2: lane = Gram tile ------------------------------
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const double *row = Y + rr * MP;
#pragma unroll
            for (int q = 0; q < TPL; ++q) {
                double ta[T], tb[T];
#pragma unroll
                for (int x = 0; x < T; ++x) { ta[x] = row[tr[q] * T + x]; tb[x] = row[tc[q] * T + x]; }
                if constexpr (SECOND) {
#pragma unroll
                    for (int x = 0; x < T; ++x)
#pragma unroll
                        for (int y = 0; y < T; ++y) acc[q][x][y] = fma(ta[x], tb[y], acc[q][x][y]);
                } else {
#pragma unroll
                    for (int x = 0; x < 2; ++x)                       // rows fM_1, fM_2 only (T == 2)
#pragma unroll
                        for (int y = 0; y < T; ++y) acc[q][x][y] = fma(ta[x], tb[y], acc[q][x][y]);
                }
            }
        }
        __syncwarp();
    }

    // ---- combine warps in warp order through shared memory, then write this CTA's partial ----
    __syncthreads();
    constexpr int TT = T * T;
    double *red = smem;                                   // [TPL*32][TT] + 4
    sA = warp_sum(sA); qA = warp_sum(qA); sB = warp_sum(sB); qB = warp_sum(qB);
    for (int w = 0; w < FUSED_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int q = 0; q < TPL; ++q)
#pragma unroll
                for (int x = 0; x < T; ++x)
#pragma unroll
                    for (int y = 0; y < T; ++y) {
                        double *p = red + ((size_t)(q * 32 + lane)) * TT + x * T + y;
                        *p = (w == 0) ? acc[q][x][y] : *p + acc[q][x][y];
                    }
            if (lane == 0) {
                double *s4 = red + (size_t)TPL * 32 * TT;
                if (w == 0) { s4[0] = sA; s4[1] = sB; s4[2] = qA; s4[3] = qB; }
                else { s4[0] += sA; s4[1] += sB; s4[2] += qA; s4[3] += qB; }
            }
        }
        __syncthreads();
    }
    constexpr int PER_BLOCK = TPL * 32 * TT + 4;
    for (int e = threadIdx.x; e < PER_BLOCK; e += blockDim.x) blockpart[(size_t)blockIdx.x * PER_BLOCK + e] = red[e];
}

// f(M_1[0]): the common shift for the variance sums (identical on every rank).
template <int K, class F>
__global__ void shift_kernel(SourceDev src, FusedConst<K> fc, F f, double *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double x[K];
    for (int d = 0; d < K; ++d) {
        double p = src.raw ? src.raw[d] : halton_coord(src.h, d, (uint32_t)src.start);
        if (fc.scale_kind == VS_SCALE_LINEAR) p = __dadd_rn(__dmul_rn(p, fc.wr[d]), fc.lb[d]);
        else if (fc.scale_kind == VS_SCALE_POWER) p = __dmul_rn(fc.lb[d], pow(fc.wr[d], p));
        x[d] = p;
    }
    *out = f(x, F::token);
}

template <int K, class F, bool SECOND, bool SEPARABLE>
static int launch_fused_t(vs_ctx *c, const SourceDev &src, const FusedConst<K> &fc, const F &f, uint64_t i_begin, uint64_t i_end,
                          double *partials) {
    constexpr int M = 2 + 2 * K;
    constexpr int T = SECOND ? gram_tile_for(M) : 2;
    constexpr int NT = (M + T - 1) / T;
    constexpr int MP = (NT * T) % 2 ? NT * T : NT * T + 1;
    constexpr int NTILES = SECOND ? NT * (NT + 1) / 2 : NT;
    constexpr int TPL = SECOND ? 1 : (NTILES + 31) / 32;
    const uint32_t nterms = src.raw ? 0u : src.h.total_terms;
    size_t smem_run = ((size_t)nterms + 1 + (size_t)FUSED_WARPS * 32 * MP) * sizeof(double);
    size_t smem_red = ((size_t)TPL * 32 * T * T + 4) * sizeof(double);
    size_t smem = smem_run > smem_red ? smem_run : smem_red;
    VS_REQUIRE(smem <= c->smem_optin, VS_ERR_UNSUPPORTED, "fused kernel needs %zu bytes of shared memory", smem);
    auto kern = fused_kernel<K, F, SECOND, SEPARABLE>;
    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint64_t rows = i_end - i_begin;
    uint64_t nbatch = (rows + 31) / 32;
    uint64_t want = (nbatch + FUSED_WARPS - 1) / FUSED_WARPS;
    int grid = (int)(want < (uint64_t)c->sm_count ? want : (uint64_t)c->sm_count);
    if (grid < 1) grid = 1;
    constexpr int PER_BLOCK = TPL * 32 * T * T + 4;
    VS_TRY(ensure(c, c->block_buf, (size_t)grid * PER_BLOCK * sizeof(double)));
    VS_TRY(ensure(c, c->misc_buf, 64));
    shift_kernel<K, F><<<1, 32, 0, c->stream>>>(src, fc, f, (double *)c->misc_buf.p);
    c->launches++;
    time_begin(c);
    kern<<<grid, FUSED_WARPS * 32, smem, c->stream>>>(src, fc, f, i_begin, i_end, (const double *)c->misc_buf.p,
                                                      (double *)c->block_buf.p);
    time_end(c);
    c->launches++;
    VS_CUDA(cudaGetLastError());
    GramGeom g{};
    g.m = M; g.l = 1; g.T = T; g.nt = NT; g.mp = MP;
    g.tr_max = SECOND ? NT : 1;
    g.ntiles = NTILES; g.LG = 32; g.RG = 1; g.R = 32; g.passes = TPL;
    return launch_gram_scatter(c, g, grid, (const double *)c->block_buf.p, partials, (int)vs_partials_len(K, 1));
}

template <int K>
static int fill_const(vs_ctx *c, const SourceDev &src, const ScaleDev &s, FusedConst<K> &fc) {
    fc.scale_kind = s.kind;
    double h[2 * K];
    if (s.kind != VS_SCALE_IDENTITY) {
        VS_CUDA(cudaMemcpyAsync(h, s.lb, sizeof(double) * 2 * K, cudaMemcpyDeviceToHost, c->stream));   // lb | wr are contiguous
        VS_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (int d = 0; d < K; ++d) {
        fc.lb[d] = s.kind != VS_SCALE_IDENTITY ? h[d] : 0.0;
        fc.wr[d] = s.kind != VS_SCALE_IDENTITY ? h[K + d] : 1.0;
        fc.toff[d] = 0;
        fc.nd[d] = 0;
    }
    if (!src.raw) {
        uint32_t off = 0;
        for (int d = 0; d < K; ++d) {
            fc.toff[d] = off;
            uint32_t need = 0;                                   // digits of this run's largest index
            for (uint64_t m = src.start + 2 * src.n - 1; m > 0; m /= prime_at(d)) ++need;
            fc.nd[d] = need;
            off += c->halton.ndigits[d] * prime_at(d);         // layout of the (possibly longer) cached table
        }
        // the cached table may have more digits than this run needs: its layout is what matters
        VS_REQUIRE(off == src.h.total_terms, VS_ERR_ARG, "Halton table layout mismatch (%u vs %u)", off, src.h.total_terms);
    }
    return VS_OK;
}

template <int K>
static int dispatch_k(vs_ctx *c, const SourceDev &src, const ScaleDev &s, const ObjectiveDev &o, uint64_t i_begin, uint64_t i_end,
                      int flags, double *partials) {
    FusedConst<K> fc;
    VS_TRY(fill_const<K>(c, src, s, fc));
    const bool second = flags & VS_FLAG_SECOND_ORDER, sep = flags & VS_FLAG_SEPARABLE;
    double hp[3 * K + 2];
    VS_CUDA(cudaMemcpyAsync(hp, o.params, sizeof(double) * o.n_params, cudaMemcpyDeviceToHost, c->stream));
    VS_CUDA(cudaStreamSynchronize(c->stream));
    if (o.id == VS_OBJ_GFUNCTION) {
        GFunctionReg<K> f;
        f.C = 1.0;
        for (int d = 0; d < K; ++d) { f.a[d] = hp[d]; f.C *= hp[K + d]; }
        if (second) {
            if (sep) return launch_fused_t<K, GFunctionReg<K>, true, true>(c, src, fc, f, i_begin, i_end, partials);
            return launch_fused_t<K, GFunctionReg<K>, true, false>(c, src, fc, f, i_begin, i_end, partials);
        }
        if (sep) return launch_fused_t<K, GFunctionReg<K>, false, true>(c, src, fc, f, i_begin, i_end, partials);
        return launch_fused_t<K, GFunctionReg<K>, false, false>(c, src, fc, f, i_begin, i_end, partials);
    }
    if constexpr (K == 3) {
        if (o.id == VS_OBJ_ISHIGAMI) {
            IshigamiReg<K> f{hp[0], hp[1]};
            if (second) return launch_fused_t<K, IshigamiReg<K>, true, false>(c, src, fc, f, i_begin, i_end, partials);
            return launch_fused_t<K, IshigamiReg<K>, false, false>(c, src, fc, f, i_begin, i_end, partials);
        }
    }
    set_error("objective %d has no fused kernel for k=%d", o.id, K);
    return VS_ERR_UNSUPPORTED;
}

bool fused_supported(int k, int objective, int flags) {
    (void)flags;
    if (objective == VS_OBJ_GFUNCTION) {
        switch (k) {
        case 2: case 3: case 4: case 5: case 6: case 8: case 10: case 12: case 16: case 20: return true;
        default: return false;
        }
    }
    if (objective == VS_OBJ_ISHIGAMI) return k == 3;
    return false;
}

int launch_fused(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, const ObjectiveDev &o, uint64_t i_begin,
                 uint64_t i_end, int flags, double *partials) {
    switch (k) {
#define VS_FUSED_CASE(KK) case KK: return dispatch_k<KK>(c, src, s, o, i_begin, i_end, flags, partials);
        VS_FUSED_CASE(2) VS_FUSED_CASE(3) VS_FUSED_CASE(4) VS_FUSED_CASE(5) VS_FUSED_CASE(6) VS_FUSED_CASE(8)
        VS_FUSED_CASE(10) VS_FUSED_CASE(12) VS_FUSED_CASE(16) VS_FUSED_CASE(20)
#undef VS_FUSED_CASE
    }
    set_error("no fused kernel for k=%d", k);
    return VS_ERR_UNSUPPORTED;
}

}  // namespace vs
