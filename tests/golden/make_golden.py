"""Regenerates tests/golden/*.npz.  Run from the repo root: python tests/golden/make_golden.py

What the vectors are anchored on (see oracle/__init__.py for the parity status):
* halton_*     -- oracle.halton (ghalton restatement, parity unpinned) -- the reference cannot be
                  imported here (Python 2, ghalton absent); these freeze the restatement so that any
                  later change of the term arithmetic is caught.
* sobol_joekuo -- scipy.stats.qmc.Sobol(scramble=False, bits=32), an independent implementation.
* c1_*         -- oracle.saltelli (restatement of varsens/saltelli.py) on BASELINE config 1
                  (README.md:33-37: g-function k=6 n=1024 a=[0,.5,3,9,99,99]).
* perm_*       -- numpy.random.RandomState(1).permutation(n) (the reference's own RNG call,
                  varsens/saltelli.py:100-101).
"""
import os
import sys
import warnings

import numpy

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import halton, saltelli, scale, sobol, objectives  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    out["halton_k20_first401"] = halton.halton_points(20, 401, 64)
    out["halton_k50_first1001"] = halton.halton_points(50, 1001, 16)
    out["halton_k20_first33554000"] = halton.halton_points(20, 33554000, 16)   # > 2^25: 26 binary digits
    out["perm_1024_head"] = numpy.random.RandomState(1).permutation(1024)[:16]
    out["perm_2p24_head"] = numpy.random.RandomState(1).permutation(1 << 24)[:16]

    a = [0, 0.5, 3, 9, 99, 99]
    v = saltelli.Varsens(lambda x: objectives.g_function_row(x, a), lambda x: x, 6, 1024, verbose=False)
    out["c1_M_1_head"] = v.sample.M_1[:4]
    out["c1_M_2_head"] = v.sample.M_2[:4]
    for name in ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n"):
        out["c1_" + name] = getattr(v, name)

    lb = numpy.array([-100.0, -10.0, 1000.0, 0.5, 3.0])
    ub = numpy.array([100.0, 20.0, 2000.0, 0.75, 3.5])
    s = saltelli.Sample(5, 13, lambda x: scale.linear(x, lb, ub), verbose=False)
    out["flat_k5_n13_linear_lb"] = lb
    out["flat_k5_n13_linear_ub"] = ub
    out["flat_k5_n13_linear"] = s.flat()
    s = saltelli.Sample(5, 13, lambda x: x, discard=7, verbose=False)
    out["flat_k5_n13_identity_discard7"] = s.flat()

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from scipy.stats import qmc
        q = qmc.Sobol(8, scramble=False, bits=32)
        q.fast_forward(4097)                      # quantlib/sobolGen.cpp:50 skipTo(4096), first draw = point 4097
        out["sobol_joekuo_k8_from4097"] = q.random(32)
    out["sobol_joekuo_k8_dirnums"] = sobol.joe_kuo_direction_numbers(8)
    numpy.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), sorted(out))


if __name__ == "__main__":
    main()
