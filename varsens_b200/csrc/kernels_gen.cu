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

// ---------------------------------------------------------------------------------------------
// K3, bulk form (even k, 16-byte aligned output): persistent CTAs, warp-specialised.
//   warps 1..7  GENERATE the tile of TI base rows: A (M_1 rows) and B (shuffled M_2 rows), 2*TI*k radical inverses with the
//               term table, the bases and the division magics in shared memory (the old kernel went to global memory for every
//               digit), into one of two tile buffers -- the next tile is generated while the current one is being stored;
//   warp 0      STORES the 2+2k flat blocks of the tile with the TMA: a block of the flat layout is the tile with ONE column
//               replaced, so the warp patches that column in place (lane = row: TI shared stores), issues one
//               cp.async.bulk.global.shared::cta of the whole TI*k-double tile (12.8 KB at k = 50), and restores the column
//               after the bulk read has drained; the A tile and the B tile alternate (N_j[j] is B with column j from A,
//               N_nj[j] is A with column j from B), so one bulk store is always in flight while the other tile is patched.
// No thread touches the 171 GB going out: per flat block the SM issues ~70 instructions instead of ~40 per 8 bytes.
// Two addressings of the output: a window [row_begin,row_end) of Sample.flat() (mode 0: export batches), or the BASE-ROW
// shard [i_lo,i_hi) of every block (mode 1: out[(t*rows + i - i_lo)*k + c]; multi-GPU export: a rank generates only its rows).
// ---------------------------------------------------------------------------------------------
int launch_sample_flat(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, uint64_t row_begin, uint64_t row_end, double *out);

struct ExportGeom {
    int k, TI, mode;
    uint64_t i_lo, i_hi;             // base rows covered by the launch
    uint64_t row_begin, row_end;     // mode 0: flat-row window
    uint32_t table_len;              // doubles of the term table copied to shared memory (0: read it from global memory)
    uint64_t ntiles;
};

#ifndef EX_WAIT_HINT_NS
#define EX_WAIT_HINT_NS 500u
#endif
__device__ __forceinline__ uint32_t ex_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ex_bar_init(uint64_t *b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ex_smem(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void ex_bar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ex_smem(b)) : "memory");
}
__device__ __forceinline__ void ex_bar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(ex_smem(b)), "r"(parity), "r"(EX_WAIT_HINT_NS)      // short suspend hint: the hand-over latency is on the tile's critical path
        : "memory");
}
__device__ __forceinline__ void ex_bulk_store(double *dst, const double *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ex_smem(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ex_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ex_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ex_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int EX_THREADS = 512, EX_GEN = EX_THREADS - 32;

__global__ void __launch_bounds__(EX_THREADS, 1)
sample_flat_bulk_kernel(ExportGeom g, SourceDev src, ScaleDev s, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int k = g.k, TI = g.TI;
    // layout: bars[4] | base[k] off[k] (u32) | magic[k] (u64) | lb[k] wr[k] | table[table_len] | A0 B0 A1 B1 (TI*k each)
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);                 // full[2], empty[2]
    uint32_t *sbase = reinterpret_cast<uint32_t *>(smem + 4);
    uint32_t *soff = sbase + k;
    uint64_t *smagic = reinterpret_cast<uint64_t *>(smem + 4 + ((2 * k + 1) / 2));
    double *slb = reinterpret_cast<double *>(smagic + k);
    double *swr = slb + k;
    double *table = swr + k;
    double *tiles = table + ((g.table_len + 1) & ~1u);
    const size_t tile = (size_t)TI * k;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        ex_bar_init(bars + 0, EX_GEN / 32);
        ex_bar_init(bars + 1, EX_GEN / 32);
        ex_bar_init(bars + 2, 1);
        ex_bar_init(bars + 3, 1);
    }
    if (!src.raw) {
        for (int d = tid; d < k; d += EX_THREADS) { sbase[d] = src.h.base[d]; soff[d] = src.h.off[d]; smagic[d] = src.h.magic[d]; }
        for (uint32_t e = tid; e < g.table_len; e += EX_THREADS) table[e] = src.h.terms[e];
    }
    for (int d = tid; d < k; d += EX_THREADS) {
        slb[d] = s.kind != VS_SCALE_IDENTITY ? s.lb[d] : 0.0;
        swr[d] = s.kind != VS_SCALE_IDENTITY ? s.wr[d] : 1.0;
    }
    __syncthreads();
    const double *T = g.table_len ? table : src.h.terms;
    const uint64_t n = src.n;
    const uint64_t my_tiles = blockIdx.x < g.ntiles ? (g.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp != 0) {
        // ------------------------------------------ generators ------------------------------------------
        // lane = tile row, warp = a set of dimensions: all lanes of a warp walk the SAME base, so the digit loop has one trip
        // count per warp (no divergence), the base / magic / table row are warp-uniform, and the A chain (consecutive indices:
        // conflict-free table reads) and the B chain (permuted indices) of a row advance together as two independent chains.
        const int gw = warp - 1;
        for (uint64_t it = 0; it < my_tiles; ++it) {
            const int pb = (int)(it & 1);
            const uint64_t i0 = g.i_lo + (blockIdx.x + it * gridDim.x) * (uint64_t)TI;
            const int rows = (int)((i0 + TI <= g.i_hi) ? TI : (g.i_hi - i0));
            double *A = tiles + (size_t)(2 * pb) * tile, *B = A + tile;
            const bool live = lane < rows;
            const uint64_t i = i0 + (live ? lane : 0);
            const uint64_t pi = src.perm[i];
            ex_bar_wait(bars + 2 + pb, (uint32_t)(((it >> 1) & 1) ^ 1));                 // the store warp is done with this buffer pair
            for (int d = gw; d < k; d += EX_GEN / 32) {
                double pa, pbv;
                if (src.raw) {
                    pa = src.raw[i * (uint64_t)k + d];
                    pbv = src.raw[(n + pi) * (uint64_t)k + d];
                } else {
                    const uint32_t b = sbase[d];
                    const uint64_t magic = smagic[d];
                    uint32_t ma = (uint32_t)(src.start + i), mb = (uint32_t)(src.start + n + pi);
                    if (src.h.mode == VS_HALTON_HORNER) {
                        pa = radical_inverse_horner(b, magic, ma);
                        pbv = radical_inverse_horner(b, magic, mb);
                    } else if (b == 2u) {
                        pa = (double)__brev(ma) * 2.3283064365386962890625e-10;
                        pbv = (double)__brev(mb) * 2.3283064365386962890625e-10;
                    } else {
                        const double *row = T + soff[d];
                        pa = 0.0;
                        pbv = 0.0;
                        while ((ma | mb) != 0u) {                       // an exhausted index keeps adding row[0] == 0.0: exact
                            const uint32_t qa = (uint32_t)__umul64hi((uint64_t)ma, magic), qb = (uint32_t)__umul64hi((uint64_t)mb, magic);
                            pa = __dadd_rn(pa, row[ma - qa * b]);
                            pbv = __dadd_rn(pbv, row[mb - qb * b]);
                            row += b;
                            ma = qa;
                            mb = qb;
                        }
                    }
                }
                if (s.kind == VS_SCALE_LINEAR) { pa = __dadd_rn(__dmul_rn(pa, swr[d]), slb[d]); pbv = __dadd_rn(__dmul_rn(pbv, swr[d]), slb[d]); }
                else if (s.kind == VS_SCALE_POWER) { pa = __dmul_rn(slb[d], pow(swr[d], pa)); pbv = __dmul_rn(slb[d], pow(swr[d], pbv)); }
                if (live) {
                    A[(size_t)lane * k + d] = pa;
                    B[(size_t)lane * k + d] = pbv;
                }
            }
            ex_fence_async();                                   // my tile entries must be visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0) ex_bar_arrive(bars + pb);
        }
        return;
    }
    // ---------------------------------------------- store warp ----------------------------------------------
    const uint64_t shard = g.i_hi - g.i_lo;
    for (uint64_t it = 0; it < my_tiles; ++it) {
        const int pb = (int)(it & 1);
        const uint64_t i0 = g.i_lo + (blockIdx.x + it * gridDim.x) * (uint64_t)TI;
        const int rows = (int)((i0 + TI <= g.i_hi) ? TI : (g.i_hi - i0));
        double *A = tiles + (size_t)(2 * pb) * tile, *B = A + tile;
        ex_bar_wait(bars + pb, (uint32_t)((it >> 1) & 1));
        // one block of the flat layout: tile rows [r0, r1) that fall into the window, contiguous in HBM
        auto put = [&](int t, const double *buf) {
            int r0 = 0, r1 = rows;
            uint64_t orow;
            if (g.mode == 0) {
                const uint64_t R0 = (uint64_t)t * n + i0;
                if (R0 + rows <= g.row_begin || R0 >= g.row_end) return;
                if (R0 < g.row_begin) r0 = (int)(g.row_begin - R0);
                if (R0 + rows > g.row_end) r1 = (int)(g.row_end - R0);
                orow = R0 + r0 - g.row_begin;
            } else {
                orow = (uint64_t)t * shard + (i0 - g.i_lo);
            }
            if (lane == 0) ex_bulk_store(out + orow * (uint64_t)k, buf + (size_t)r0 * k, (uint32_t)((r1 - r0) * k * 8));
        };
        // blocks [t_min, t_max] of this tile intersect the window (R0(t) = t n + i0 is monotonic in t): two divisions per tile,
        // then blocks outside cost one integer compare -- no patch, no fence, no wait
        int t_min = 0, t_max = 2 * k + 1;
        if (g.mode == 0) {
            const uint64_t lo_num = g.row_begin > i0 + (uint64_t)rows - 1 ? g.row_begin - i0 - (uint64_t)rows + 1 : 0;   // t n > row_begin - i0 - rows
            const uint64_t tm = (lo_num + n - 1) / n;
            t_min = tm > (uint64_t)(2 * k + 2) ? 2 * k + 2 : (int)tm;
            if (g.row_end <= i0) t_max = -1;
            else {
                const uint64_t tx = (g.row_end - 1 - i0) / n;
                t_max = tx > (uint64_t)(2 * k + 1) ? 2 * k + 1 : (int)tx;
            }
        }
        auto wanted = [&](int t) { return t >= t_min && t <= t_max; };
        int last = -1;                                          // tile of the most recent bulk group: 0 = A, 1 = B
        int pa_col = -1, pb_col = -1;                           // column currently patched in A / B
        double pa_val = 0.0, pb_val = 0.0;                      // ... and its original value (this lane's row)
        // Before a tile changes, its own last bulk read must have drained: if the newest group belongs to the OTHER tile,
        // "all but one group done" is enough (that one keeps streaming while we patch), else everything has to be done.
        auto quiesce = [&](int tilesel) {
            if (lane == 0) {
                if (last == tilesel) ex_wait_read<0>();
                else ex_wait_read<1>();
            }
            __syncwarp();
        };
        if (wanted(1)) { put(1, B); if (lane == 0) ex_commit(); last = 1; }          // M_2
        if (wanted(0)) { put(0, A); if (lane == 0) ex_commit(); last = 0; }          // M_1
        for (int j = 0; j < k; ++j) {
            const bool nb = wanted(2 + j), na = wanted(2 + k + j);
            if (!nb && !na) continue;
            const double a_j = lane < rows ? A[(size_t)lane * k + j] : 0.0;     // column j is never the patched one (pa_col, pb_col < j)
            const double b_j = lane < rows ? B[(size_t)lane * k + j] : 0.0;
            if (nb) {                                           // N_j[j] = B with column j from A
                quiesce(1);
                if (lane < rows) {
                    if (pb_col >= 0) B[(size_t)lane * k + pb_col] = pb_val;
                    B[(size_t)lane * k + j] = a_j;
                }
                pb_col = j;
                pb_val = b_j;
                ex_fence_async();
                __syncwarp();
                put(2 + j, B);
                if (lane == 0) ex_commit();
                last = 1;
            }
            if (na) {                                           // N_nj[j] = A with column j from B
                quiesce(0);
                if (lane < rows) {
                    if (pa_col >= 0) A[(size_t)lane * k + pa_col] = pa_val;
                    A[(size_t)lane * k + j] = b_j;
                }
                pa_col = j;
                pa_val = a_j;
                ex_fence_async();
                __syncwarp();
                put(2 + k + j, A);
                if (lane == 0) ex_commit();
                last = 0;
            }
        }
        if (lane == 0) ex_wait_read<0>();                       // both tiles may be overwritten by the generators now
        __syncwarp();
        if (lane == 0) ex_bar_arrive(bars + 2 + pb);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores complete before the CTA exits
}

// base rows touched by a flat-row window: if it lies inside one block, that block's row range; else all of them
static void window_rows(uint64_t n, uint64_t row_begin, uint64_t row_end, uint64_t *i_lo, uint64_t *i_hi) {
    *i_lo = 0;
    *i_hi = n;
    const uint64_t t_first = row_begin / n, t_last = (row_end - 1) / n;
    if (t_first == t_last) {
        *i_lo = row_begin - t_first * n;
        *i_hi = row_end - t_first * n;
    } else if (t_last == t_first + 1 && row_end - t_last * n <= row_begin - t_first * n) {
        // two arcs that do not overlap: [row_begin - t_first n, n) of the first block and [0, row_end - t_last n) of the next.
        // (kept simple: generate everything; the arcs case only matters for windows shorter than one block)
    }
}

static int launch_sample_flat_bulk(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, int mode, uint64_t i_lo, uint64_t i_hi,
                                   uint64_t row_begin, uint64_t row_end, double *out, bool *done) {
    *done = false;
    if ((k & 1) || (reinterpret_cast<uintptr_t>(out) & 15) || i_hi <= i_lo) return VS_OK;
    ExportGeom g{};
    g.k = k;
    g.mode = mode;
    g.i_lo = i_lo;
    g.i_hi = i_hi;
    g.row_begin = row_begin;
    g.row_end = row_end;
    const size_t avail = c->smem_optin;
    const size_t fixed = (4 + (2 * (size_t)k + 1) / 2 + 3 * (size_t)k + 2) * sizeof(double);
    const size_t tab = src.raw ? 0 : (((size_t)src.h.total_terms + 1) & ~(size_t)1) * sizeof(double);
    // tile height: 32 rows (one lane per row in the store warp) if four tiles fit beside the term table; else without the table
    int TI = 32;
    bool with_table = !src.raw && src.h.mode != VS_HALTON_HORNER && fixed + tab + 4 * (size_t)TI * k * 8 <= avail;
    if (!with_table)
        while (TI > 1 && fixed + 4 * (size_t)TI * k * 8 > avail) TI >>= 1;
    if (fixed + (with_table ? tab : 0) + 4 * (size_t)TI * k * 8 > avail) return VS_OK;
    if (((size_t)TI * k * 8) >= (1u << 20)) return VS_OK;                       // bulk copy size field
    g.TI = TI;
    g.table_len = with_table ? src.h.total_terms : 0;
    g.ntiles = (i_hi - i_lo + TI - 1) / TI;
    const size_t smem = fixed + (with_table ? tab : 0) + 4 * (size_t)TI * k * 8;
    static size_t smem_set[64] = {};
    if (c->device >= 64 || smem_set[c->device] < smem) {
        VS_CUDA(cudaFuncSetAttribute(sample_flat_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->device < 64) smem_set[c->device] = smem;
    }
    const unsigned grid = (unsigned)(g.ntiles < (uint64_t)c->sm_count ? g.ntiles : (uint64_t)c->sm_count);
    time_begin(c);
    sample_flat_bulk_kernel<<<grid, EX_THREADS, smem, c->stream>>>(g, src, s, out);
    time_end(c);
    c->launches++;
    VS_CUDA(cudaGetLastError());
    *done = true;
    return VS_OK;
}

// Base-row shard of every block (vs_sample_flat_shard): out[(t*rows + i - i_begin)*k + c], rows = i_end - i_begin.
int launch_sample_shard(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, uint64_t i_begin, uint64_t i_end, double *out) {
    if (i_end <= i_begin) return VS_OK;
    bool done = false;
    VS_TRY(launch_sample_flat_bulk(c, k, src, s, 1, i_begin, i_end, 0, 0, out, &done));
    if (done) return VS_OK;
    // odd k / unaligned output: block by block through the scalar kernel (each block's shard rows are one flat-row window)
    const uint64_t rows = i_end - i_begin;
    for (int t = 0; t < 2 + 2 * k; ++t)
        VS_TRY(launch_sample_flat(c, k, src, s, (uint64_t)t * src.n + i_begin, (uint64_t)t * src.n + i_end, out + (uint64_t)t * rows * k));
    return VS_OK;
}

int launch_sample_flat(vs_ctx *c, int k, const SourceDev &src, const ScaleDev &s, uint64_t row_begin, uint64_t row_end,
                       double *out) {
    if (row_end <= row_begin) return VS_OK;
    const uint64_t n = src.n;
    if (!c->opt.no_bulk_export) {
        uint64_t bl, bh;
        window_rows(n, row_begin, row_end, &bl, &bh);
        bool done = false;
        VS_TRY(launch_sample_flat_bulk(c, k, src, s, 0, bl, bh, row_begin, row_end, out, &done));
        if (done) return VS_OK;
    }
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
