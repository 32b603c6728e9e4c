"""Restatement of the 32-bit Gray-code Sobol generator behind quantlib/sobolGen.cpp.

TEST INFRASTRUCTURE (see oracle/__init__.py).  **parity unpinned**: QuantLib (version unpinned,
sobolGen.cpp:1 links bare -lQuantLib) and its Levitan-Lemieux direction integers are not under
/root/reference.  Call sites: quantlib/sobolGen.cpp:47 (SobolRsg(dims, seed,
SobolLevitanLemieux)), :50 (skipTo(4096)), :54 (nextSequence).  What is restated is the
published, table-independent algorithm:

    state x[d] (u32);  point number m (m = 0 is the origin) has
    x[d] = XOR over set bits b of gray(m) = m ^ (m >> 1) of V[d][b];   value = x[d] * 2^-32
    sequential form:  x[d] ^= V[d][ctz(~c)] with c the count of points already drawn.

sobolGen row r (0-based) is point number 4097 + r in that numbering, printed with 6
significant digits (default ostream precision, sobolGen.cpp:59).

Direction integers: Joe-Kuo (new-joe-kuo-6.21201) built from scipy's bundled table, the only
table available offline; the logic is pinned bit-exactly against
scipy.stats.qmc.Sobol(scramble=False, bits=32) in tests/test_oracle_reference_parity.py
(test_sobol_against_scipy_fixture) and the golden fixture tests/golden/golden.npz.
"""
import os

import numpy


def joe_kuo_direction_numbers(k, bits=32):
    """(k, bits) uint32 direction integers V[d][j], MSB-aligned (V[0][j] = 1 << (31-j))."""
    import scipy
    path = os.path.join(os.path.dirname(scipy.__file__), 'stats', '_sobol_direction_numbers.npz')
    z = numpy.load(path)
    poly, vinit = z['poly'], z['vinit']
    if k > len(poly):
        raise ValueError("Joe-Kuo table has %d dimensions" % len(poly))
    V = numpy.zeros((k, bits), dtype=numpy.uint64)
    for j in range(bits):
        V[0, j] = 1 << (bits - 1 - j)
    for d in range(1, k):
        p = int(poly[d])
        s = p.bit_length() - 1                      # degree
        m = [int(vinit[d][i]) for i in range(s)]    # initial odd integers m_1..m_s
        for i in range(s, bits):
            new = m[i - s] ^ (m[i - s] << s)
            for j in range(1, s):
                if (p >> (s - j)) & 1:              # coefficient a_j of x^(s-j)
                    new ^= m[i - j] << j
            m.append(new)
        for j in range(bits):
            V[d, j] = m[j] << (bits - 1 - j)
    return V.astype(numpy.uint32)


def sobol_points(V, first_point, count):
    """(count, k) float64 by direct Gray-code indexing: points first_point .. first_point+count-1."""
    V = numpy.asarray(V, dtype=numpy.uint32)
    m = numpy.arange(int(first_point), int(first_point) + int(count), dtype=numpy.uint64)
    g = (m ^ (m >> numpy.uint64(1))).astype(numpy.uint32)
    x = numpy.zeros((int(count), V.shape[0]), dtype=numpy.uint32)
    for b in range(32):
        sel = ((g >> numpy.uint32(b)) & numpy.uint32(1)).astype(bool)
        x[sel] ^= V[:, b]
    return x.astype(numpy.float64) * (2.0 ** -32)


def sobol_points_sequential(V, first_point, count):
    """Same points by the stateful next() recurrence (cross-check of the skip-ahead)."""
    V = numpy.asarray(V, dtype=numpy.uint32)
    k = V.shape[0]
    m = int(first_point)
    g = m ^ (m >> 1)
    x = numpy.zeros(k, dtype=numpy.uint32)
    for b in range(32):
        if (g >> b) & 1:
            x ^= V[:, b]
    out = numpy.zeros((int(count), k))
    for r in range(int(count)):
        out[r] = x.astype(numpy.float64) * (2.0 ** -32)
        c = m + r                       # points drawn so far; flip the lowest zero bit's direction
        j = 0
        while (c >> j) & 1:
            j += 1
        x = x ^ V[:, j]
    return out


def quantize_6sig(x):
    """What ``cout << double`` (6 significant digits) followed by numpy.loadtxt returns."""
    flat = numpy.asarray(x, dtype=numpy.float64).ravel()
    return numpy.array([float("%.6g" % v) for v in flat]).reshape(numpy.shape(x))
