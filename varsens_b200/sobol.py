"""Sobol points for ``Sample(raw=...)`` -- the device replacement of quantlib/sobolGen.cpp.

The reference generates Sobol samples with a stand-alone QuantLib program (``SobolGen dims samples [seed]``,
quantlib/sobolGen.cpp:47-63), writes them to a CSV file with 6 significant digits and loads the (2n, k) file
through ``Sample(..., loadFile=..., delimiter=',')`` (varsens/saltelli.py:74-77,225-227).  Here the same Gray-code
sequence is produced by direct indexing on the GPU (``vs_sobol``): row r is point number ``4097 + r`` (``skipTo(4096)``
leaves the generator on point 4097, sobolGen.cpp:50-54) and ``quantize6=True`` reproduces the decimal round trip of the
file exactly.

Direction integers: QuantLib's Levitan-Lemieux table is not available offline ("parity unpinned"), so
the default is the Joe-Kuo table bundled with scipy; any (k, 32) uint32 table, MSB-aligned, can be passed instead.
"""
import os

import numpy

from . import _cabi

SOBOLGEN_FIRST_POINT = 4097


def joe_kuo_direction_numbers(k, bits=32):
    """(k, 32) uint32 direction integers from scipy's new-joe-kuo-6.21201 table (V[0][j] = 1 << (31 - j))."""
    import scipy
    z = numpy.load(os.path.join(os.path.dirname(scipy.__file__), "stats", "_sobol_direction_numbers.npz"))
    poly, vinit = z["poly"], z["vinit"]
    if k > len(poly):
        raise _cabi.VarsensError("the Joe-Kuo table has %d dimensions, k = %d" % (len(poly), k))
    V = numpy.zeros((k, bits), dtype=numpy.uint64)
    V[0] = [1 << (bits - 1 - j) for j in range(bits)]
    for d in range(1, k):
        p = int(poly[d])
        s = p.bit_length() - 1
        m = [int(v) for v in vinit[d][:s]]
        for i in range(s, bits):
            new = m[i - s] ^ (m[i - s] << s)
            for j in range(1, s):
                if (p >> (s - j)) & 1:
                    new ^= m[i - j] << j
            m.append(new)
        V[d] = [m[j] << (bits - 1 - j) for j in range(bits)]
    return V.astype(numpy.uint32)


def sobol_raw(k, n, dirnums=None, first_point=SOBOLGEN_FIRST_POINT, quantize6=True, device=None, out=None):
    """The unscaled (2n, k) sample ``SobolGen k 2n`` would have written, ready for ``Sample(k, n, scaling, raw=...)``.
    ``out`` may be a CUDA torch tensor (the points then stay on the GPU)."""
    V = joe_kuo_direction_numbers(k) if dirnums is None else numpy.ascontiguousarray(dirnums, dtype=numpy.uint32)
    ctx = _cabi.Context.get(device)
    return ctx.sobol(k, first_point, 2 * int(n), V, quantize6=quantize6, out=out)
