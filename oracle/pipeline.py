"""Chunked, vectorised numpy form of the reference pipeline.  TEST INFRASTRUCTURE.

Same arithmetic as oracle.saltelli (= varsens/saltelli.py:82-125, :308-355, :572-622) but the
2n(1+k) rows are produced and consumed ``chunk`` base rows at a time so that n = 2^20..2^24
fits in memory (the literal reference needs 2*k*n*k*8 bytes for N_j/N_nj, saltelli.py:119).
This is BASELINE.md §3.2's "CPU numpy" path and the checker for large-n parity tests.

Row i of the base design:
    A_i = scale(h(s + 1 + i)),  B_i = scale(h(s + 1 + n + perm[i])),  s = 20k + discard,
    perm = numpy.random.RandomState(1).permutation(n)   (== seed(1); shuffle(M_2), :100-101).
"""
import numpy

from . import halton as _halton


def permutation(n):
    """The row permutation the reference applies to M_2 (saltelli.py:100-101)."""
    return numpy.random.RandomState(1).permutation(int(n))


def base_rows(k, n, discard, scaling, i0, i1, perm, raw=None):
    s = 20 * k + int(discard)
    if raw is None:
        A = _halton.halton_points(k, s + 1 + i0, i1 - i0)
        b_idx = (s + 1 + n) + perm[i0:i1].astype(numpy.int64)
        B = numpy.empty((i1 - i0, k))
        # gather: evaluate each needed index once, in sorted order, then un-sort
        order = numpy.argsort(b_idx, kind='stable')
        B[order] = _halton_at(k, b_idx[order])
    else:
        A = raw[i0:i1]
        B = raw[n + perm[i0:i1]]
    return scaling(A), scaling(B)


def _halton_at(k, idx):
    idx = numpy.asarray(idx, dtype=numpy.uint64)
    out = numpy.zeros((idx.size, k))
    for d, b in enumerate(_halton.first_primes(k)):
        m = idx.copy()
        x = numpy.zeros(idx.size)
        bp = float(b)
        ub = numpy.uint64(b)
        while m.size and m.max() > 0:
            x += (m % ub).astype(numpy.float64) / bp
            m //= ub
            bp *= b
        out[:, d] = x
    return out


def evaluate_chunk(A, B, frows):
    """f over the 2+2k points of every base row.  Returns fA, fB (rows,), fNj, fNnj (k, rows)."""
    rows, k = A.shape
    fA, fB = frows(A), frows(B)
    fNj = numpy.empty((k, rows))
    fNnj = numpy.empty((k, rows))
    X = B.copy()
    Y = A.copy()
    for j in range(k):
        X[:, j] = A[:, j]           # N_j[j]  = M_2 with column j from M_1   (saltelli.py:119-123)
        Y[:, j] = B[:, j]           # N_nj[j] = M_1 with column j from M_2
        fNj[j] = frows(X)
        fNnj[j] = frows(Y)
        X[:, j] = B[:, j]
        Y[:, j] = A[:, j]
    return fA, fB, fNj, fNnj


def chunk_sums(fA, fB, fNj, fNnj, acc_dtype=numpy.longdouble):
    """Partial sums of one chunk in extended precision (the 'exact' checker)."""
    t = acc_dtype
    a, b, J, N = fA.astype(t), fB.astype(t), fNj.astype(t), fNnj.astype(t)
    return dict(
        s_ab=(a * b).sum(), s_a=a.sum(), s_b=b.sum(), q_a=(a * a).sum(), q_b=(b * b).sum(),
        aJ=(J * a).sum(axis=1), bN=(N * b).sum(axis=1), aN=(N * a).sum(axis=1), bJ=(J * b).sum(axis=1),
        NJ=N @ J.T, NN=N @ N.T, JJ=J @ J.T)


def add_sums(x, y):
    return y if x is None else {key: x[key] + y[key] for key in x}


def indices_from_sums(S, k, n):
    """The estimators of saltelli.py:577-622 written on the sufficient statistics."""
    n = int(n)
    ld = numpy.longdouble
    E_2 = S['s_ab'] / ld(n)                                             # :577
    tot = S['s_a'] + S['s_b']
    var_y = (S['q_a'] + S['q_b'] - tot * tot / ld(2 * n)) / ld(2 * n - 1)   # :583 (ddof=1 over 2n values)
    U_j = (S['aJ'] / ld(n - 1) + S['bN'] / ld(n - 1)) / ld(2)           # :591-593
    U_nj = (S['aN'] / ld(n - 1) + S['bJ'] / ld(n - 1)) / ld(2)          # :594-596
    sens = (U_j - E_2) / var_y                                          # :608
    sens_t = 1.0 - (U_nj - E_2) / var_y                                 # :609
    sens_2 = ((S['NJ'] + S['NJ'].T) / ld(2 * (n - 1)) - E_2) / var_y    # :612-616
    sens_2n = ((S['NN'] + S['JJ']) / ld(2 * (n - 1)) - E_2) / var_y     # :618-622
    f = lambda v: numpy.asarray(v, dtype=numpy.float64)
    return dict(E_2=f(E_2).reshape(1), var_y=f(var_y).reshape(1), U_j=f(U_j).reshape(k, 1),
                U_nj=f(U_nj).reshape(k, 1), sens=f(sens).reshape(k, 1), sens_t=f(sens_t).reshape(k, 1),
                sens_2=f(sens_2).reshape(k, 1, k, 1), sens_2n=f(sens_2n).reshape(k, 1, k, 1))


def run(k, n, scaling, frows, discard=0, chunk=1 << 15, i0=0, i1=None, perm=None, raw=None,
        acc_dtype=numpy.longdouble, finalize=True):
    """Whole pipeline (or the slice [i0,i1) of it) -> indices dict (or raw sums if finalize=False)."""
    k, n = int(k), int(n)
    i1 = n if i1 is None else i1
    perm = permutation(n) if perm is None else perm
    S = None
    for c0 in range(i0, i1, chunk):
        c1 = min(c0 + chunk, i1)
        A, B = base_rows(k, n, discard, scaling, c0, c1, perm, raw)
        S = add_sums(S, chunk_sums(*evaluate_chunk(A, B, frows), acc_dtype=acc_dtype))
    return indices_from_sums(S, k, n) if finalize else S
