"""Sobol points for ``Sample(raw=...)`` -- the device replacement of quantlib/sobolGen.cpp.

The reference generates Sobol samples with a stand-alone QuantLib program (``SobolGen dims samples [seed]``,
quantlib/sobolGen.cpp:47-63), writes them to a CSV file with 6 significant digits and loads the (2n, k) file
through ``Sample(..., loadFile=..., delimiter=',')`` (varsens/saltelli.py:74-77,225-227).  Here the same Gray-code
sequence is produced by direct indexing on the GPU (``vs_sobol``): row r is point number ``4097 + r`` (``skipTo(4096)``
leaves the generator on point 4097, sobolGen.cpp:50-54) and ``quantize6=True`` reproduces the decimal round trip of the
file exactly.

Direction integers: QuantLib's Levitan-Lemieux table is not available offline ("parity unpinned"), so
the default is the Joe-Kuo table bundled with scipy; any (k, 32) uint32 table, MSB-aligned, can be passed instead --
in particular one read from QuantLib itself: ``quantlib_direction_numbers(k, "ql/math/randomnumbers/sobolrsg.cpp")``
parses the initialiser arrays of a QuantLib source tree (or a plain text file, one dimension per line) and runs
QuantLib's recurrence on them.
"""
import os
import re

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


def _poly_table():
    import scipy
    z = numpy.load(os.path.join(os.path.dirname(scipy.__file__), "stats", "_sobol_direction_numbers.npz"))
    return [int(p) for p in z["poly"]]


_C_ARRAY = re.compile(r"(?:unsigned\s+long|unsigned\s+int|unsigned|std::uint_least32_t|std::uint32_t|uint32_t|long)\s+(\w+)\s*"
                      r"\[\s*\d*\s*\]\s*=\s*\{([^{}]*)\}", re.S)
_C_PTR_TABLE = re.compile(r"\*\s*(?:const\s+)?(\w+)\s*\[\s*\d*\s*\]\s*=\s*\{([^{}]*)\}", re.S)


def read_sobol_initializers(source, kind="LevitanLemieux"):
    """Initialisers m_1..m_s of dimensions 2, 3, ... (dimension 1 has none) as a list of int lists.

    ``source`` is a path or the text of either
      * a QuantLib ``sobolrsg.cpp``: per-dimension arrays ``static const unsigned long dim<NN><tag>initializers[] =
        { m_1, ..., m_s, 0UL };`` and a pointer table ``<tag...>initializers[N] = { dim02..., dim03..., ... }`` that
        fixes their order.  ``kind`` selects the table by a case-insensitive substring of its name
        (SobolRsg::DirectionIntegers: "LevitanLemieux" -> the table sobolGen.cpp:34,47 asks for; "SobolLevitan", "Kuo",
        "JoeKuoD7", ...).  Without a pointer table the arrays whose names contain ``kind`` are taken in the order of the
        number in their name.
      * plain text: one dimension per line (starting with dimension 2), integers separated by blanks or commas,
        ``#`` comments."""
    text = source
    if "\n" not in source and os.path.exists(source):
        with open(source) as fh:
            text = fh.read()
    arrays = {}
    for name, body in _C_ARRAY.findall(text):
        vals = [int(re.sub(r"[uUlL]+$", "", t), 0) for t in re.split(r"[\s,]+", body.strip()) if t]
        if vals and vals[-1] == 0:
            vals = vals[:-1]                                            # QuantLib terminates each array with 0UL
        arrays[name] = vals
    if arrays:
        for tname, body in _C_PTR_TABLE.findall(text):
            names = [t for t in re.split(r"[\s,]+", body.strip()) if t]
            if kind.lower() in tname.lower() and names and all(nm in arrays for nm in names):
                return [arrays[nm] for nm in names]
        picked = [(int(re.search(r"(\d+)", nm).group(1)), nm) for nm in arrays if kind.lower() in nm.lower() and re.search(r"\d", nm)]
        if not picked:
            raise _cabi.VarsensError("no initialiser arrays matching %r in the given source" % (kind,))
        return [arrays[nm] for _, nm in sorted(picked)]
    out = []
    for line in text.splitlines():
        line = line.split("#")[0].strip()
        if line:
            out.append([int(t, 0) for t in re.split(r"[\s,]+", line) if t])
    if not out:
        raise _cabi.VarsensError("no initialisers found")
    return out


def direction_numbers_from_initializers(k, initializers, polynomials=None, bits=32):
    """(k, 32) uint32 direction integers, MSB-aligned, from per-dimension initialisers by the recurrence QuantLib's SobolRsg
    constructor runs (V_l = V_{l-s} ^ (V_{l-s} >> s) ^ XOR_{j<s, a_j=1} V_{l-j}; the first s are m_l << (bits - l)); the same
    Bratley-Fox recurrence as Joe-Kuo's.  ``polynomials``: primitive polynomials of dimensions 2, 3, ... as integers with all
    coefficient bits (x^3 + x + 1 -> 11).  Default: the degree-then-value listing of scipy's table, which is the order
    QuantLib's PrimitivePolynomials table (Jaeckel's PPMT) enumerates them in as far as could be established offline --
    pass QuantLib's own list if it differs."""
    polys = _poly_table()[1:] if polynomials is None else [int(p) for p in polynomials]
    if k - 1 > len(initializers) or k - 1 > len(polys):
        raise _cabi.VarsensError("initialisers for %d and polynomials for %d dimensions given, k = %d"
                                 % (len(initializers) + 1, len(polys) + 1, k))
    V = numpy.zeros((k, bits), dtype=numpy.uint64)
    V[0] = [1 << (bits - 1 - j) for j in range(bits)]
    for d in range(1, k):
        p = polys[d - 1]
        s = p.bit_length() - 1
        m = [int(v) for v in initializers[d - 1]]
        if len(m) < s:
            raise _cabi.VarsensError("dimension %d: polynomial of degree %d needs %d initialisers, got %d" % (d + 1, s, s, len(m)))
        m = m[:s]
        for i, v in enumerate(m):
            if v % 2 == 0 or v >= (1 << (i + 1)):
                raise _cabi.VarsensError("dimension %d: initialiser m_%d = %d is not an odd integer below 2^%d" % (d + 1, i + 1, v, i + 1))
        for i in range(s, bits):
            new = m[i - s] ^ (m[i - s] << s)
            for j in range(1, s):
                if (p >> (s - j)) & 1:
                    new ^= m[i - j] << j
            m.append(new)
        V[d] = [m[j] << (bits - 1 - j) for j in range(bits)]
    return V.astype(numpy.uint32)


def quantlib_direction_numbers(k, source, kind="LevitanLemieux", polynomials=None):
    """Direction integers for ``sobol_raw(dirnums=...)`` from a QuantLib source file / initialiser text (see
    read_sobol_initializers): with QuantLib's Levitan-Lemieux table this makes ``sobol_raw`` the same numbers
    ``SobolGen`` prints (quantlib/sobolGen.cpp:34,47-63) for every dimension the table covers."""
    return direction_numbers_from_initializers(k, read_sobol_initializers(source, kind), polynomials)


def sobol_raw(k, n, dirnums=None, first_point=SOBOLGEN_FIRST_POINT, quantize6=True, device=None, out=None):
    """The unscaled (2n, k) sample ``SobolGen k 2n`` would have written, ready for ``Sample(k, n, scaling, raw=...)``.
    ``out`` may be a CUDA torch tensor (the points then stay on the GPU)."""
    V = joe_kuo_direction_numbers(k) if dirnums is None else numpy.ascontiguousarray(dirnums, dtype=numpy.uint32)
    ctx = _cabi.Context.get(device)
    return ctx.sobol(k, first_point, 2 * int(n), V, quantize6=quantize6, out=out)
