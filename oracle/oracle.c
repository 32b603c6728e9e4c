/* CPU restatement (plain C + OpenMP) of the varsens Saltelli hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Never linked into libvarsens_b200.so.
 * Used (through oracle/cport.py) by tests/ as a full-size checker and by bench.py's
 * cpu_baseline / --impl reference legs as the timed CPU arm ("port").
 *
 * PARITY: the Halton generator restates ghalton (third party, unpinned, not under
 * /root/reference; call sites varsens/saltelli.py:82-84) -> "parity unpinned" at bit level;
 * it is cross-checked bit-for-bit against oracle/halton.py in tests/test_oracle_cport.py.
 *
 * Row i of the base design (varsens/saltelli.py:83-84,92-101):
 *   A_i = scale(h(s+1+i)),  B_i = scale(h(s+1+n+perm[i])),  s = 20k + discard.
 * Flat layout (varsens/saltelli.py:127-160): rows M_1 | M_2 | N_j[0..k) | N_nj[0..k).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_K 512

static void first_primes(int k, uint32_t *p) {
    int c = 0;
    for (uint32_t v = 2; c < k; ++v) {
        int ok = 1;
        for (int i = 0; i < c && p[i] * p[i] <= v; ++i)
            if (v % p[i] == 0) { ok = 0; break; }
        if (ok) p[c++] = v;
    }
}

/* ghalton Halton::get restated: least-significant digit first, divide then add. */
static inline double radical_inverse(uint64_t m, uint32_t b) {
    double x = 0.0, bp = (double)b;
    while (m > 0) {
        x += (double)(m % b) / bp;
        m /= b;
        bp *= (double)b;
    }
    return x;
}

/* varsens/scale.py:33 (linear; w = ub - lb computed once, as numpy does) and :62 (power; r = ub/lb). */
static inline double apply_scale(int kind, double p, double lb, double w_or_r) {
    if (kind == 1) { double t = p * w_or_r; return t + lb; }
    if (kind == 2) return lb * pow(w_or_r, p);
    return p;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

void orc_halton(int k, uint64_t first_index, uint64_t count, double *out) {
    uint32_t b[ORC_MAX_K];
    first_primes(k, b);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)count; ++r)
        for (int d = 0; d < k; ++d) out[(size_t)r * k + d] = radical_inverse(first_index + (uint64_t)r, b[d]);
}

static void base_row(int k, uint64_t n, uint64_t s, const uint32_t *b, const uint32_t *perm, const double *raw,
                     int kind, const double *lb, const double *wr, uint64_t i, double *A, double *B) {
    uint64_t ia = s + 1 + i, ib = s + 1 + n + perm[i];
    for (int d = 0; d < k; ++d) {
        double pa = raw ? raw[(size_t)i * k + d] : radical_inverse(ia, b[d]);
        double pb = raw ? raw[((size_t)n + perm[i]) * k + d] : radical_inverse(ib, b[d]);
        A[d] = apply_scale(kind, pa, lb ? lb[d] : 0.0, wr ? wr[d] : 0.0);
        B[d] = apply_scale(kind, pb, lb ? lb[d] : 0.0, wr ? wr[d] : 0.0);
    }
}

/* Sample.flat() rows [row_begin,row_end) -> out (row_end-row_begin, k).  varsens/saltelli.py:127-160. */
void orc_sample_flat(int k, uint64_t n, uint64_t discard, const uint32_t *perm, const double *raw, int kind,
                     const double *lb, const double *wr, uint64_t row_begin, uint64_t row_end, double *out) {
    uint32_t b[ORC_MAX_K];
    first_primes(k, b);
    uint64_t s = 20ull * k + discard;
#pragma omp parallel for schedule(static)
    for (int64_t R = (int64_t)row_begin; R < (int64_t)row_end; ++R) {
        double A[ORC_MAX_K], B[ORC_MAX_K];
        uint64_t blk = (uint64_t)R / n, i = (uint64_t)R % n;
        base_row(k, n, s, b, perm, raw, kind, lb, wr, i, A, B);
        double *o = out + (size_t)(R - (int64_t)row_begin) * k;
        if (blk == 0) memcpy(o, A, sizeof(double) * k);
        else if (blk == 1) memcpy(o, B, sizeof(double) * k);
        else if (blk < 2 + (uint64_t)k) { memcpy(o, B, sizeof(double) * k); o[blk - 2] = A[blk - 2]; }
        else { memcpy(o, A, sizeof(double) * k); o[blk - 2 - k] = B[blk - 2 - k]; }
    }
}

/* ---- objectives (oracle/objectives.py is the numpy twin) ---- */
static double f_gfunction(const double *x, int k, const double *a) {
    double p = 1.0;
    for (int c = 0; c < k; ++c) p *= (fabs(4.0 * x[c] - 2.0) + a[c]) / (1.0 + a[c]);
    return p;
}

static double f_ishigami(const double *x, int k, const double *par) {
    (void)k;
    double s0 = sin(x[0]), s1 = sin(x[1]), x2 = x[2];
    return s0 + par[0] * s1 * s1 + par[1] * (x2 * x2) * (x2 * x2) * s0;
}

/* RK4 mass-action chain, FROZEN ARITHMETIC shared with the device functor (varsens_b200/csrc/device.cuh: RK4Chain):
 *   flux_s = fma(kf_s, X_s, -(kr_s * X_{s+1}))     d_s = flux_{s-1} - flux_s
 *   stage  : T = fma(h, d, X)                       sum: a = k1, a = fma(2, k2, a), a = fma(2, k3, a), a = a + k4
 *   update : X = fma(dt/6, a, X)
 * fma() is the C99 correctly rounded fused multiply-add (hardware vfmadd through glibc's ifunc, or exact software);
 * everything else is compiled with -ffp-contract=off, so these are the only fused operations.  With identical rate
 * constants the trajectories are bit-identical to the device's. */
static void chain_rhs(const double *X, const double *kf, const double *kr, int S, double *d) {
    double prev = 0.0;
    for (int s = 0; s < S; ++s) {
        double back = kr[s] * X[s + 1];
        double flux = fma(kf[s], X[s], -back);
        d[s] = prev - flux;
        prev = flux;
    }
    d[S] = prev;
}

static double f_rk4_chain(const double *x, int k, const double *par) {
    int S = k / 2, nsteps = (int)par[1];
    double dt = par[0];
    double h2 = 0.5 * dt, h6 = dt / 6.0;
    double X[ORC_MAX_K / 2 + 1], T[ORC_MAX_K / 2 + 1], a[ORC_MAX_K / 2 + 1], d[ORC_MAX_K / 2 + 1];
    for (int s = 0; s <= S; ++s) X[s] = 0.0;
    X[0] = 1.0;
    for (int it = 0; it < nsteps; ++it) {
        chain_rhs(X, x, x + S, S, d);
        for (int s = 0; s <= S; ++s) { a[s] = d[s]; T[s] = fma(h2, d[s], X[s]); }
        chain_rhs(T, x, x + S, S, d);
        for (int s = 0; s <= S; ++s) { a[s] = fma(2.0, d[s], a[s]); T[s] = fma(h2, d[s], X[s]); }
        chain_rhs(T, x, x + S, S, d);
        for (int s = 0; s <= S; ++s) { a[s] = fma(2.0, d[s], a[s]); T[s] = fma(dt, d[s], X[s]); }
        chain_rhs(T, x, x + S, S, d);
        for (int s = 0; s <= S; ++s) { double t = a[s] + d[s]; X[s] = fma(h6, t, X[s]); }
    }
    return X[S];
}

static double eval_objective(int id, const double *x, int k, const double *par) {
    switch (id) {
    case 0: return f_gfunction(x, k, par);
    case 1: return f_ishigami(x, k, par);
    default: return f_rk4_chain(x, k, par);
    }
}

/* The 2+2k objective values of base row i, in the order fM_1, fM_2, fN_j[0..k), fN_nj[0..k)
 * (varsens/saltelli.py:329-353). */
static void row_values(int k, const double *A, const double *B, int id, const double *par, double *v) {
    double X[ORC_MAX_K];
    v[0] = eval_objective(id, A, k, par);
    v[1] = eval_objective(id, B, k, par);
    memcpy(X, B, sizeof(double) * k);
    for (int j = 0; j < k; ++j) { X[j] = A[j]; v[2 + j] = eval_objective(id, X, k, par); X[j] = B[j]; }
    memcpy(X, A, sizeof(double) * k);
    for (int j = 0; j < k; ++j) { X[j] = B[j]; v[2 + k + j] = eval_objective(id, X, k, par); X[j] = A[j]; }
}

/* Flat objective values for base rows [i0,i1): out[t*(i1-i0) + (i-i0)], t in [0, 2+2k). */
void orc_values(int k, uint64_t n, uint64_t discard, const uint32_t *perm, const double *raw, int kind,
                const double *lb, const double *wr, int id, const double *par, uint64_t i0, uint64_t i1,
                double *out) {
    uint32_t b[ORC_MAX_K];
    first_primes(k, b);
    uint64_t s = 20ull * k + discard, rows = i1 - i0;
#pragma omp parallel for schedule(static)
    for (int64_t i = (int64_t)i0; i < (int64_t)i1; ++i) {
        double A[ORC_MAX_K], B[ORC_MAX_K], v[2 * ORC_MAX_K + 2];
        base_row(k, n, s, b, perm, raw, kind, lb, wr, (uint64_t)i, A, B);
        row_values(k, A, B, id, par, v);
        for (int t = 0; t < 2 + 2 * k; ++t) out[(size_t)t * rows + (size_t)(i - (int64_t)i0)] = v[t];
    }
}

/* Sufficient statistics of base rows [i0,i1) in long double: sums[0..m) = sum v_t,
 * gram[t*m+u] = sum v_t v_u (full m x m), m = 2+2k.  The estimators of
 * varsens/saltelli.py:577-622 are functions of these (oracle/pipeline.py:indices_from_sums). */
void orc_sums(int k, uint64_t n, uint64_t discard, const uint32_t *perm, const double *raw, int kind,
              const double *lb, const double *wr, int id, const double *par, uint64_t i0, uint64_t i1,
              int second_order, long double *sums, long double *gram) {
    uint32_t b[ORC_MAX_K];
    first_primes(k, b);
    uint64_t s = 20ull * k + discard;
    int m = 2 + 2 * k;
    for (int t = 0; t < m; ++t) sums[t] = 0.0L;
    for (int t = 0; t < m * m; ++t) gram[t] = 0.0L;
#pragma omp parallel
    {
        long double *ls = (long double *)calloc((size_t)m + (size_t)m * m, sizeof(long double));
        long double *lg = ls + m;
#pragma omp for schedule(static)
        for (int64_t i = (int64_t)i0; i < (int64_t)i1; ++i) {
            double A[ORC_MAX_K], B[ORC_MAX_K], v[2 * ORC_MAX_K + 2];
            base_row(k, n, s, b, perm, raw, kind, lb, wr, (uint64_t)i, A, B);
            row_values(k, A, B, id, par, v);
            for (int t = 0; t < m; ++t) ls[t] += (long double)v[t];
            int tmax = second_order ? m : 2;
            for (int t = 0; t < tmax; ++t)
                for (int u = t; u < m; ++u) lg[t * m + u] += (long double)v[t] * (long double)v[u];
        }
#pragma omp critical
        {
            for (int t = 0; t < m; ++t) sums[t] += ls[t];
            for (int t = 0; t < m; ++t)
                for (int u = t; u < m; ++u) gram[t * m + u] += lg[t * m + u];
        }
        free(ls);
    }
    for (int t = 0; t < m; ++t)
        for (int u = 0; u < t; ++u) gram[t * m + u] = gram[u * m + t];
}
