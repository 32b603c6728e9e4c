// Device building blocks: Halton radical inverse, scaling, objective functors.
#pragma once

#include "vs_internal.cuh"

namespace vs {

// ---------------------------------------------------------------------------------------------
// Halton: ghalton's in-order digit sum (call sites varsens/saltelli.py:82-84).
//   x = 0; for j = 0.. (least-significant digit first): x += digit_j / b^(j+1)
// digit_j / b^(j+1) is read from the host-built table, so the device performs exactly the same
// sequence of IEEE additions as the CPU statement and no fp64 division.  __dadd_rn forbids any
// re-association / contraction.  Base 2 is exact in every partial sum, hence a bit reversal.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double radical_inverse(const double *__restrict__ T, uint32_t b, uint64_t magic, uint32_t m) {
    if (b == 2u) return (double)__brev(m) * 2.3283064365386962890625e-10;  // 2^-32, exact
    double x = 0.0;
    while (m != 0u) {
        uint32_t q = (uint32_t)__umul64hi((uint64_t)m, magic);
        uint32_t d = m - q * b;
        x = __dadd_rn(x, T[d]);
        T += b;
        m = q;
    }
    return x;
}

__device__ __forceinline__ double halton_coord(const HaltonDev &h, int d, uint32_t m) {
    return radical_inverse(h.terms + h.off[d], h.base[d], h.magic[d], m);
}

// ---------------------------------------------------------------------------------------------
// scale.py:33 (two roundings: multiply, then add -- never an FMA) and :62.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double apply_scale(const ScaleDev &s, int d, double p) {
    if (s.kind == VS_SCALE_LINEAR) return __dadd_rn(__dmul_rn(p, s.wr[d]), s.lb[d]);
    if (s.kind == VS_SCALE_POWER) return __dmul_rn(s.lb[d], pow(s.wr[d], p));
    return p;
}

// Unscaled coordinate d of A_i / B_i:
//   A_i = h(start + i),  B_i = h(start + n + perm[i])      (or the rows of `raw`)
__device__ __forceinline__ double source_a(const SourceDev &src, int k, uint64_t i, int d) {
    if (src.raw) return src.raw[i * (uint64_t)k + d];
    return halton_coord(src.h, d, (uint32_t)(src.start + i));
}
__device__ __forceinline__ double source_b(const SourceDev &src, int k, uint64_t pi, int d) {
    if (src.raw) return src.raw[(src.n + pi) * (uint64_t)k + d];
    return halton_coord(src.h, d, (uint32_t)(src.start + src.n + pi));
}

// ---------------------------------------------------------------------------------------------
// Objective functors.  X is any type with `double operator[](int) const` (registers, shared
// memory view, ...).  params layout on the device (built by get_objective):
//   GFUNCTION : [0,k) a_c        [k,2k) 1/(1+a_c)      [2k,3k) a_c/(1+a_c)
//   ISHIGAMI  : A, B
//   RK4_CHAIN : dt, nsteps
// ---------------------------------------------------------------------------------------------
struct GFunction {
    const double *inv;   // 1/(1+a_c)
    const double *off;   // a_c/(1+a_c)
    static constexpr bool product_form = true;           // f(x) = prod_c term(c, x_c), evaluated in ascending c from 1.0
    __device__ __forceinline__ double term(int c, double x) const {
        double t = fma(4.0, x, -2.0);                     // 4x - 2         (2 flops)
        return fma(fabs(t), inv[c], off[c]);              // (|t| + a)/(1+a) (2)
    }
    template <class X>
    __device__ __forceinline__ double operator()(const X &x, int k) const {
        double p = 1.0;
        for (int c = 0; c < k; ++c) p *= term(c, x[c]);   // running product (1)
        return p;
    }
};

struct Ishigami {
    double A, B;
    template <class X>
    __device__ __forceinline__ double operator()(const X &x, int) const {
        double s0 = sin(x[0]), s1 = sin(x[1]), x2 = x[2];
        double x22 = x2 * x2;
        return s0 + A * s1 * s1 + B * (x22 * x22) * s0;
    }
};

// Reversible chain, S links, S+1 species.  Register-resident for compile-time S.
template <int S>
struct RK4Chain {
    double dt;
    int nsteps;
    __device__ __forceinline__ static void rhs(const double (&X)[S + 1], const double (&kf)[S], const double (&kr)[S],
                                               double (&d)[S + 1]) {
        double prev = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            double flux = kf[s] * X[s] - kr[s] * X[s + 1];
            d[s] = prev - flux;
            prev = flux;
        }
        d[S] = prev;
    }
    template <class XV>
    __device__ __forceinline__ double operator()(const XV &x, int) const {
        double kf[S], kr[S], X[S + 1], T[S + 1], a[S + 1], d[S + 1];
#pragma unroll
        for (int s = 0; s < S; ++s) { kf[s] = x[s]; kr[s] = x[S + s]; }
#pragma unroll
        for (int s = 0; s <= S; ++s) X[s] = 0.0;
        X[0] = 1.0;
        const double h2 = 0.5 * dt, h6 = dt / 6.0;
        for (int it = 0; it < nsteps; ++it) {
            rhs(X, kf, kr, d);                                   // k1
#pragma unroll
            for (int s = 0; s <= S; ++s) { a[s] = d[s]; T[s] = X[s] + h2 * d[s]; }
            rhs(T, kf, kr, d);                                   // k2
#pragma unroll
            for (int s = 0; s <= S; ++s) { a[s] += 2.0 * d[s]; T[s] = X[s] + h2 * d[s]; }
            rhs(T, kf, kr, d);                                   // k3
#pragma unroll
            for (int s = 0; s <= S; ++s) { a[s] += 2.0 * d[s]; T[s] = X[s] + dt * d[s]; }
            rhs(T, kf, kr, d);                                   // k4
#pragma unroll
            for (int s = 0; s <= S; ++s) X[s] += h6 * (a[s] + d[s]);
        }
        return X[S];
    }
};

// Run-time S (any k): state in local memory.
struct RK4ChainDyn {
    double dt;
    int nsteps;
    static constexpr int MAXS = 64;
    template <class XV>
    __device__ double operator()(const XV &x, int k) const {
        const int S = k / 2;
        double X[MAXS + 1], T[MAXS + 1], a[MAXS + 1], d[MAXS + 1];
        for (int s = 0; s <= S; ++s) X[s] = 0.0;
        X[0] = 1.0;
        const double h2 = 0.5 * dt, h6 = dt / 6.0;
        auto rhs = [&](const double *Y) {
            double prev = 0.0;
            for (int s = 0; s < S; ++s) {
                double flux = x[s] * Y[s] - x[S + s] * Y[s + 1];
                d[s] = prev - flux;
                prev = flux;
            }
            d[S] = prev;
        };
        for (int it = 0; it < nsteps; ++it) {
            rhs(X);
            for (int s = 0; s <= S; ++s) { a[s] = d[s]; T[s] = X[s] + h2 * d[s]; }
            rhs(T);
            for (int s = 0; s <= S; ++s) { a[s] += 2.0 * d[s]; T[s] = X[s] + h2 * d[s]; }
            rhs(T);
            for (int s = 0; s <= S; ++s) { a[s] += 2.0 * d[s]; T[s] = X[s] + dt * d[s]; }
            rhs(T);
            for (int s = 0; s <= S; ++s) X[s] += h6 * (a[s] + d[s]);
        }
        return X[S];
    }
};

__host__ __device__ __forceinline__ void tile_coords(int id, int nt, int &tr, int &tc) {
    tr = 0;
    int rowlen = nt;
    while (id >= rowlen) { id -= rowlen; --rowlen; ++tr; }
    tc = tr + id;
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace vs
