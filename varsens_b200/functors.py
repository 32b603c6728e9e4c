"""Registered device objective functors.

The reference calls a Python callable once per sample row (varsens/saltelli.py:308-353).  An
instance of one of these classes can be passed wherever the reference takes ``objective_func`` /
``objective``; the evaluation then happens inside the CUDA kernels (fused with generation and
the estimator reductions when a fused kernel exists for k, otherwise through the two-phase
eval-to-HBM + Gram path).  They carry no host implementation: there is no CPU fallback.
"""
import numpy

from . import _cabi


class Functor(object):
    objective_id = None
    separable = False

    def params(self, k):
        raise NotImplementedError

    def __call__(self, x):
        raise _cabi.VarsensError(
            "%s is a device functor: pass it to Objective/Varsens instead of calling it on the host"
            % type(self).__name__)


class GFunction(Functor):
    """Sobol g-function  prod_c (|4 x_c - 2| + a_c) / (1 + a_c)
    (varsens/tests/test_g_function.py:9-13, README.md:33-36)."""
    objective_id = _cabi.OBJ_GFUNCTION
    separable = True

    def __init__(self, a):
        self.a = numpy.asarray(a, dtype=numpy.float64).copy()

    def params(self, k):
        if self.a.shape != (k,):
            raise _cabi.VarsensError("GFunction has %d coefficients, k = %d" % (self.a.size, k))
        return self.a


class Ishigami(Functor):
    """sin x0 + A sin^2 x1 + B x2^4 sin x0  (BASELINE.json config 2)."""
    objective_id = _cabi.OBJ_ISHIGAMI

    def __init__(self, A=7.0, B=0.1):
        self.A, self.B = float(A), float(B)

    def params(self, k):
        return numpy.array([self.A, self.B])


class RK4Chain(Functor):
    """Fixed-step RK4 of the reversible mass-action chain X_0 <-> ... <-> X_{k/2}; the k parameters are the
    forward (first half) and reverse (second half) rate constants; objective = X_{k/2}(nsteps*dt)
    (BASELINE.json config 5; spec frozen in include/varsens_b200.h)."""
    objective_id = _cabi.OBJ_RK4_CHAIN

    def __init__(self, dt=0.01, nsteps=1000):
        self.dt, self.nsteps = float(dt), int(nsteps)

    def params(self, k):
        return numpy.array([self.dt, float(self.nsteps)])


def vectorized(fn):
    """Mark a Python objective as vectorised: it receives a (rows, k) CUDA torch.float64 tensor of
    sample rows and returns a (rows,) or (rows, l) tensor.  Sample rows are materialised on the GPU
    in batches (export-mode kernel) and never leave it; only the callable itself is user code."""
    fn.varsens_vectorized = True
    return fn
