"""Estimators of varsens/saltelli.py:572-622 evaluated in x87 extended precision (numpy.longdouble: 64-bit significand).

TEST INFRASTRUCTURE (oracle/__init__.py).  The reference computes them in fp64 with numpy reductions whose summation
order -- hence last bits -- depends on the memory layout (3.5e-15 drift between two layouts of the same
values at n = 4096, and more on ill-conditioned data where U_j - E_2 cancels).  This restatement performs the SAME
formulas, with the same mixed n / n-1 normalisation, on longdouble copies of the fp64 objective values: products of two
doubles are rounded once to 64 bits and the sums carry 11 extra bits, so the result is the correctly rounded fp64 answer
for every practical input.  It is the yardstick the 1e-10 relative / 1e-12 absolute contract is measured against where
the fp64 reference's own rounding noise would otherwise be part of the difference."""
import numpy

LD = numpy.longdouble


def indices(vals, k, n, rows=None):
    """vals: (2*rows*(1+k), l) fp64 objective values in Objective.flat() order (rows == n unless NaN rows were trimmed,
    saltelli.py:474-495: the divisors stay n and n-1).  Returns a dict of float64 arrays shaped like the reference's."""
    vals = numpy.asarray(vals, dtype=numpy.float64)
    if vals.ndim == 1:
        vals = vals.reshape(-1, 1)
    rows = n if rows is None else rows
    l = vals.shape[1]
    v = vals.astype(LD).reshape(2 + 2 * k, rows, l)
    fA, fB, fJ, fN = v[0], v[1], v[2:2 + k], v[2 + k:]
    nn = LD(n)
    E2 = (fA * fB).sum(axis=0) / nn                                             # :577
    both = numpy.concatenate((fA, fB), axis=0)
    mean = both.sum(axis=0) / LD(2 * rows)
    var = ((both - mean) ** 2).sum(axis=0) / LD(2 * rows - 1)                   # :583 ddof=1
    Uj = ((fA[None] * fJ).sum(axis=1) / (nn - 1) + (fB[None] * fN).sum(axis=1) / (nn - 1)) / LD(2)     # :591-593
    Unj = ((fA[None] * fN).sum(axis=1) / (nn - 1) + (fB[None] * fJ).sum(axis=1) / (nn - 1)) / LD(2)    # :594-596
    sens = (Uj - E2) / var                                                       # :608
    sens_t = LD(1) - (Unj - E2) / var                                            # :609
    s2 = numpy.tensordot(fN, fJ, axes=([1], [1])) + numpy.tensordot(fJ, fN, axes=([1], [1]))          # :612-613
    s2 = (s2 / (LD(2) * (nn - 1)) - E2) / var                                    # :614-616 (broadcast over the LAST output index)
    s2n = numpy.tensordot(fN, fN, axes=([1], [1])) + numpy.tensordot(fJ, fJ, axes=([1], [1]))         # :618-619
    s2n = (s2n / (LD(2) * (nn - 1)) - E2) / var                                  # :620-622
    f = lambda a: numpy.asarray(a, dtype=numpy.float64)
    return dict(E_2=f(E2), var_y=f(var), U_j=f(Uj), U_nj=f(Unj), sens=f(sens), sens_t=f(sens_t), sens_2=f(s2), sens_2n=f(s2n))
