#!/usr/bin/env python
"""Pins the Halton arithmetic on a REAL ghalton install (the reference's generator, varsens/saltelli.py:1,82-84).

ghalton (PyPI, F.-M. De Rainville; unpinned by the reference) is not vendored under the reference checkout and cannot be
installed in the build container, so the repo's Halton arithmetic is a restatement ("parity unpinned").
Run this script once on any machine where ``pip install ghalton`` works:

    python tests/golden/make_ghalton_golden.py            # writes tests/golden/ghalton_golden.npz

and commit the file.  tests/test_oracle_reference_parity.py::test_ghalton_golden then activates: it compares the vectors
with the oracle in each of the library's selectable arithmetics (enum vs_halton_mode) and fails unless the DEFAULT one is
bit-identical -- if another mode is, the failure message names it: set VS_HALTON_MODE (or vs_ctx_set_halton_mode) and the
CUDA kernels follow, because they only add host-built table entries in digit order (vs_halton_terms_mode).

What is dumped mirrors the reference's call pattern: Halton(k); get(20k + discard) thrown away; get(2n) kept
(saltelli.py:82-84), for a few (k, n, discard), plus long single-dimension runs that reach indices > 2^25.
"""
import os
import sys

import numpy


def main():
    try:
        import ghalton
    except ImportError:
        sys.exit("ghalton is not importable here; run this where `pip install ghalton` works")
    out = {"ghalton_version": numpy.array(getattr(ghalton, "__version__", "unknown"))}
    cases = [(6, 1024, 0), (20, 333, 11), (3, 5, 0), (50, 64, 7)]
    for k, n, discard in cases:
        seq = ghalton.Halton(k)
        seq.get(20 * k + discard)
        out["k%d_n%d_d%d" % (k, n, discard)] = numpy.array(seq.get(2 * n), dtype=numpy.float64)
    # far into the sequence: skip 2^25 points of a 4-dimensional generator, keep 4096
    seq = ghalton.Halton(4)
    left = 1 << 25
    while left:
        step = min(left, 1 << 20)
        seq.get(step)
        left -= step
    out["k4_after_2p25"] = numpy.array(seq.get(4096), dtype=numpy.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ghalton_golden.npz")
    numpy.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
