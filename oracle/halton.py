"""Restatement of ghalton's ``Halton(k).get(n)`` (identity digit permutation).

TEST INFRASTRUCTURE (see oracle/__init__.py).  **parity unpinned**: ghalton (PyPI, F.-M. De
Rainville, github fmder/ghalton; version unpinned by the reference: setup.py has no
install_requires, doc/requirements.txt:3 says bare ``ghalton``) is not vendored under
/root/reference and not installable here.  Call sites: varsens/saltelli.py:1,82,83,84.

Published algorithm restated: per dimension ``d`` the base is the d-th
prime; the point of 1-based index ``m`` is

    x = 0.0; bp = (double) b
    for j in 0 .. ndigits(m)-1:          # j = 0 is the LEAST significant base-b digit of m
        x += digit_j(m) / bp             # one IEEE fp64 divide, one IEEE fp64 add, no FMA
        bp *= b

The first point ``get`` returns is index 1 (= [1/2, 1/3, 1/5, ...]).
"""
import numpy


def first_primes(k):
    """First k primes (ghalton's bases)."""
    primes = []
    c = 2
    while len(primes) < k:
        if all(c % p for p in primes if p * p <= c):
            primes.append(c)
        c += 1
    return primes


def radical_inverse_scalar(m, b):
    """Pure-Python scalar form of the loop above (slow; small cases and the term table)."""
    x = 0.0
    bp = float(b)
    m = int(m)
    while m > 0:
        x += float(m % b) / bp
        m //= b
        bp *= b
    return x


def halton_points(k, first_index, count):
    """(count, k) float64: Halton points of indices first_index .. first_index+count-1.

    Vectorised numpy form of the same in-order digit sum; identical bits to
    radical_inverse_scalar (adding 0/bp for exhausted indices is an exact no-op).
    """
    idx = numpy.arange(int(first_index), int(first_index) + int(count), dtype=numpy.uint64)
    out = numpy.zeros((int(count), int(k)), dtype=numpy.float64)
    for d, b in enumerate(first_primes(k)):
        m = idx.copy()
        x = numpy.zeros(int(count), dtype=numpy.float64)
        bp = float(b)
        ub = numpy.uint64(b)
        while m.size and m.max() > 0:
            x += (m % ub).astype(numpy.float64) / bp
            m //= ub
            bp *= b
        out[:, d] = x
    return out


def halton_term_table(k, max_index):
    """Host term table T[dim][j][digit] = digit / b^(j+1) and its digit counts.

    Returns (bases, ndigits, offsets, terms): ``terms[offsets[d] + j*bases[d] + digit]``.
    ndigits[d] = number of base-b digits of max_index.  This is the table the CUDA path sums
    in order (varsens_b200/csrc); the library builds its own copy in C++ and the CPU tests
    compare the two.
    """
    bases = first_primes(k)
    ndigits, offsets, terms = [], [], []
    for b in bases:
        nd, m = 0, int(max_index)
        while m > 0:
            nd += 1
            m //= b
        nd = max(nd, 1)
        offsets.append(len(terms))
        ndigits.append(nd)
        bp = float(b)
        for _ in range(nd):
            for digit in range(b):
                terms.append(float(digit) / bp)
            bp *= b
    return (numpy.array(bases, dtype=numpy.uint32), numpy.array(ndigits, dtype=numpy.uint32),
            numpy.array(offsets, dtype=numpy.uint32), numpy.array(terms, dtype=numpy.float64))


MODES = ("divide", "reciprocal", "running_reciprocal", "horner")


def term_table(k, max_index, mode="divide"):
    """The term table in one of the alternative fp64 arithmetics the library can be switched to (enum vs_halton_mode,
    include/varsens_b200.h) -- candidates for what an actual ghalton build does, should it ever disagree with the restatement
    above (tests/golden/make_ghalton_golden.py turns a real install into goldens):
      divide              term = digit / b^(j+1)                   (the restatement above)
      reciprocal          term = digit * (1.0 / b^(j+1))
      running_reciprocal  term = digit * f_j, f_0 = 1.0 / b, f_{j+1} = f_j * (1.0 / b)
    Returns dict(bases, ndigits, offsets, terms)."""
    if mode not in MODES[:3]:
        raise ValueError("mode %r has no term table" % (mode,))
    bases, ndigits, offsets, _ = halton_term_table(k, max_index)
    terms = []
    for b, nd in zip(bases.tolist(), ndigits.tolist()):
        bp = float(b)
        ib = 1.0 / float(b)
        f = ib
        for _ in range(nd):
            for digit in range(b):
                if mode == "divide":
                    terms.append(float(digit) / bp)
                elif mode == "reciprocal":
                    terms.append(float(digit) * (1.0 / bp))
                else:
                    terms.append(float(digit) * f)
            bp *= b
            f = f * ib
    return dict(bases=bases, ndigits=ndigits, offsets=offsets, terms=numpy.array(terms, dtype=numpy.float64))


def halton_points_mode(k, first_index, count, mode="divide"):
    """(count, k) points in any of MODES.  The three term-table modes sum table entries least significant digit first;
    ``horner`` evaluates x = (x + digit_j) / b from the most significant digit down (one division per digit)."""
    first_index, count = int(first_index), int(count)
    out = numpy.zeros((count, int(k)), dtype=numpy.float64)
    if mode == "horner":
        for d, b in enumerate(first_primes(k)):
            for r in range(count):
                m, digits = first_index + r, []
                while m > 0:
                    digits.append(m % b)
                    m //= b
                x = 0.0
                for dg in reversed(digits):
                    x = (x + float(dg)) / float(b)
                out[r, d] = x
        return out
    t = term_table(k, first_index + count - 1, mode)
    for d, b in enumerate(t["bases"].tolist()):
        rows = t["terms"][t["offsets"][d]:].reshape(-1)[: t["ndigits"][d] * b].reshape(t["ndigits"][d], b)
        m = numpy.arange(first_index, first_index + count, dtype=numpy.uint64)
        x = numpy.zeros(count)
        for j in range(int(t["ndigits"][d])):
            x = x + rows[j][(m % numpy.uint64(b)).astype(numpy.int64)]
            m //= numpy.uint64(b)
        out[:, d] = x
    return out


class Halton(object):
    """Stateful stand-in with ghalton's interface: ``Halton(k).get(n)`` -> list of lists."""

    def __init__(self, k):
        self.k = int(k)
        self._next = 1

    def get(self, n):
        pts = halton_points(self.k, self._next, n)
        self._next += int(n)
        return pts.tolist()
