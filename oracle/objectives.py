"""Objective functions of the benchmark workloads, numpy form.  TEST INFRASTRUCTURE.

Each has a per-row form (what the reference's Objective loop calls, saltelli.py:308-353) and a
whole-matrix form (rows, k) -> (rows,) used by the vectorised CPU baseline.

* g-function: varsens/tests/test_g_function.py:9-13 and README.md:33-36.
* Ishigami: not in the reference (BASELINE.json config 2); A=7, B=0.1.
* RK4 mass-action chain: not in the reference (BASELINE.json config 5); spec frozen here and in
  varsens_b200/csrc/functors.cuh: species X_0..X_S with S = k/2 reversible links
  X_s <-> X_{s+1}, forward rate x[s], reverse rate x[S+s]; X(0) = e_0; classic RK4 with fixed
  dt for nsteps steps; objective = X_S(T).
"""
import numpy


# ---- Sobol g-function ------------------------------------------------------------------
def gi_function(xi, ai):
    return (numpy.abs(4.0 * xi - 2.0) + ai) / (1.0 + ai)


def g_function_row(x, a):
    # README.md:34 / test_g_function.py:12-13 (list comprehension + numpy.prod)
    return numpy.prod([gi_function(xi, a[i]) for i, xi in enumerate(x)])


def g_function_rows(X, a):
    a = numpy.asarray(a, dtype=numpy.float64)
    return numpy.prod((numpy.abs(4.0 * X - 2.0) + a) / (1.0 + a), axis=1)


def g_truth(a):
    # test_g_function.py:20-21
    return 1.0 / (3.0 * ((numpy.asarray(a, dtype=numpy.float64) + 1.0) ** 2.0))


def g_var(a):
    # closed form of test_g_function.py:40-49 (sum over all non-empty subsets of prod V_i)
    return float(numpy.prod(1.0 + g_truth(a)) - 1.0)


def g_truth_2(a, i, j):
    # test_g_function.py:23-25 -- closed index of the pair {i,j}, times Var
    v = g_truth(a)
    return v[i] + v[j] + v[i] * v[j]


def g_truth_vnc(a, drop):
    # test_g_function.py:27-36 -- closed variance of all factors except those in `drop`
    v = numpy.delete(g_truth(a), list(drop))
    return float(numpy.prod(1.0 + v) - 1.0)


def g_truth_t(a, i):
    # test_g_function.py:38-39
    return g_truth(a)[i] * (1.0 + g_truth_vnc(a, [i]))


# ---- Ishigami ---------------------------------------------------------------------------
def ishigami_row(x, A=7.0, B=0.1):
    return numpy.sin(x[0]) + A * numpy.sin(x[1]) ** 2 + B * x[2] ** 4 * numpy.sin(x[0])


def ishigami_rows(X, A=7.0, B=0.1):
    s0 = numpy.sin(X[:, 0])
    s1 = numpy.sin(X[:, 1])
    x2 = X[:, 2]
    return s0 + A * s1 * s1 + B * (x2 * x2) * (x2 * x2) * s0


# ---- RK4 mass-action reversible chain ----------------------------------------------------
def _chain_rhs(X, kf, kr):
    # X: (rows, S+1); kf, kr: (rows, S)
    flux = kf * X[:, :-1] - kr * X[:, 1:]
    d = numpy.zeros_like(X)
    d[:, :-1] -= flux
    d[:, 1:] += flux
    return d


def rk4_chain_rows(P, dt=0.01, nsteps=1000):
    P = numpy.asarray(P, dtype=numpy.float64)
    S = P.shape[1] // 2
    kf, kr = P[:, :S], P[:, S:2 * S]
    X = numpy.zeros((P.shape[0], S + 1))
    X[:, 0] = 1.0
    for _ in range(int(nsteps)):
        k1 = _chain_rhs(X, kf, kr)
        k2 = _chain_rhs(X + (0.5 * dt) * k1, kf, kr)
        k3 = _chain_rhs(X + (0.5 * dt) * k2, kf, kr)
        k4 = _chain_rhs(X + dt * k3, kf, kr)
        X = X + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
    return X[:, S]


def rk4_chain_row(x, dt=0.01, nsteps=1000):
    return float(rk4_chain_rows(numpy.asarray(x)[None, :], dt, nsteps)[0])
