"""Runs the REAL reference code (/root/reference/varsens/{saltelli,scale}.py) in this container and
freezes its outputs as tests/golden/reference_golden.npz.

Run from the repo root (only where /root/reference exists -- the build container):
    python tests/golden/make_reference_golden.py

The reference is Python 2 and imports the absent third-party ``ghalton``.  Nothing is copied into
the repo: the two source files are read where they lie, their ``print`` statements are rewritten
to calls in memory (the only py2->py3 change they need), a stand-in ``ghalton`` module backed by
oracle.halton is put in sys.modules, and the result is exec'd.  So these vectors pin the oracle's
restatement of saltelli.py/scale.py against the reference's own code, *conditional on* the
Halton restatement (which stays "parity unpinned", oracle/halton.py).

Reference HEAD cannot run ``Varsens(callable, ...)`` (saltelli.py:567 passes ``verbose`` into
``objective_vals``), so the vectors use the 3-step route its examples use
(varsens/examples/varsens_earm_scipy.py:163-165): Sample -> Objective -> Varsens(objective).
"""
import os
import re
import sys
import types

import numpy

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import halton as _oh  # noqa: E402

REF = "/root/reference/varsens"


def _py3(src):
    out = []
    for line in src.splitlines():
        m = re.match(r"^(\s*)(.*?)\bprint (.+?)(,?)\s*(#.*)?$", line)
        if m and not line.lstrip().startswith("#") and "print(" not in line:
            indent, head, body, trailing, _ = m.groups()
            end = ", end=' '" if trailing else ""
            line = "%s%sprint(%s%s)" % (indent, head, body, end)
        out.append(line)
    return "\n".join(out) + "\n"


def load_reference():
    gh = types.ModuleType("ghalton")
    gh.Halton = _oh.Halton
    sys.modules["ghalton"] = gh
    pkg = types.ModuleType("varsens")
    pkg.__path__ = []
    sys.modules["varsens"] = pkg
    mods = {}
    for name in ("scale", "saltelli"):
        mod = types.ModuleType("varsens." + name)
        src = _py3(open(os.path.join(REF, name + ".py")).read())
        exec(compile(src, os.path.join(REF, name + ".py"), "exec"), mod.__dict__)
        sys.modules["varsens." + name] = mod
        setattr(pkg, name, mod)
        mods[name] = mod
    return mods["saltelli"], mods["scale"]


def main():
    saltelli, scale = load_reference()
    out = {}

    def gi(xi, ai):
        return (numpy.abs(4.0 * xi - 2.0) + ai) / (1.0 + ai)

    def g(x, a):
        return numpy.prod([gi(xi, a[i]) for i, xi in enumerate(x)])

    a = [0, 0.5, 3, 9, 99, 99]
    # --- BASELINE config 1 through the reference's own classes
    s = saltelli.Sample(6, 1024, lambda x: x, 0, False)
    o = saltelli.Objective(6, 1024, s, lambda x: g(x, a), verbose=False)
    v = saltelli.Varsens(o, verbose=False)
    out["c1_flat_head"] = s.flat()[:8]
    out["c1_flat_tail"] = s.flat()[-8:]
    out["c1_flat_rows_probe"] = numpy.array([0, 1023, 1024, 2047, 2048, 3071, 8191, 8192, 14335])
    out["c1_flat_probe"] = s.flat()[out["c1_flat_rows_probe"]]
    out["c1_obj_flat"] = o.flat()
    for name in ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n"):
        out["c1_" + name] = numpy.asarray(getattr(v, name))

    # --- two-output objective (test_g_function.py:77-88), k=6 n=256
    s2 = saltelli.Sample(6, 256, lambda x: x, 0, False)
    o2 = saltelli.Objective(6, 256, s2, lambda x: [g(x, a), g(x, a[::-1])], verbose=False)
    v2 = saltelli.Varsens(o2, verbose=False)
    out["two_obj_flat"] = o2.flat()
    for name in ("E_2", "var_y", "U_j", "U_nj", "sens", "sens_t", "sens_2", "sens_2n"):
        out["two_" + name] = numpy.asarray(getattr(v2, name))

    # --- scaled samples: linear and magnitude, with discard
    lb = numpy.array([-100.0, -10.0, 1000.0, 0.5, 3.0])
    ub = numpy.array([100.0, 20.0, 2000.0, 0.75, 3.5])
    s3 = saltelli.Sample(5, 13, lambda x: scale.linear(x, lb, ub), 7, False)
    out["lin_lb"], out["lin_ub"], out["lin_flat_k5_n13_discard7"] = lb, ub, s3.flat()
    ref = numpy.array([1.0, 10.0, 1000.0, 0.5])
    s4 = saltelli.Sample(4, 9, lambda x: scale.magnitude(x, ref, orders=1.0), 0, False)
    out["mag_ref"], out["mag_flat_k4_n9"] = ref, s4.flat()
    s5 = saltelli.Sample(3, 11, lambda x: scale.percentage(x, numpy.array([1.0, 10.0, 1000.0]), 33.0), 0, False)
    out["pct_flat_k3_n11"] = s5.flat()

    # --- raw= entry (saltelli.py:69-73): in-memory (2n,k) array, e.g. a Sobol file
    rng = numpy.random.RandomState(7)
    raw = rng.rand(2 * 10, 4)
    out["raw_in"] = raw.copy()
    s6 = saltelli.Sample(4, 10, lambda x: scale.linear(x, -1.0, 2.0), 0, False, raw=raw.copy())
    out["raw_flat_k4_n10"] = s6.flat()

    # --- scale helpers on a fixed grid
    p = numpy.linspace(0.0, 1.0, 17)
    out["scale_p"] = p
    out["scale_linear"] = scale.linear(p, -3.5, 12.25)
    out["scale_power"] = scale.power(p, 0.01, 250.0)
    out["scale_percentage"] = scale.percentage(p, 40.0, 33.0)
    out["scale_magnitude"] = scale.magnitude(p, 2.5, 2.0, 10.0)

    # --- NaN-row trimming on load (saltelli.py:474-495)
    vals = o2.flat().copy()
    vals[5, 0] = numpy.nan            # fM_1 row 5
    vals[256 * 3 + 17, 0] = numpy.nan  # fN_j[1] row 17
    o3 = saltelli.Objective(6, 256, objective_vals=vals, verbose=False)
    v3 = saltelli.Varsens(o3, verbose=False)
    out["nan_obj_in"] = vals
    out["nan_fM_1_shape"] = numpy.array(o3.fM_1.shape)
    for name in ("E_2", "var_y", "sens", "sens_t", "sens_2"):
        out["nan_" + name] = numpy.asarray(getattr(v3, name))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
    numpy.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
