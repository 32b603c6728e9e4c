"""Scaling helpers with the reference's names and argument meaning (varsens/scale.py).

Called with arrays they behave exactly like the reference (same expressions, numpy broadcasting).
Called -- through the user's own scaling callable -- on the tracing probe that Sample passes in,
they return a descriptor instead, which lets Sample fuse the scaling into the CUDA generation
kernels (`scale_desc` of the C ABI).  A callable that does anything else to its argument is not
traceable; Sample then applies it to the unscaled block on the host, as the reference does.
"""
import numpy

from . import _cabi

__all__ = ["linear", "power", "percentage", "magnitude"]


class _Probe(object):
    """Stands for 'the (n,k) block of unit-cube points' while a scaling callable is traced."""
    shape = None

    def __init__(self, k):
        self.k = k


class _Traced(object):
    """Result of tracing: a lowered scale descriptor.  Deliberately supports no arithmetic."""

    def __init__(self, desc):
        self.desc = desc


def _bounds(k, lower, upper):
    lo = numpy.array(numpy.broadcast_to(numpy.asarray(lower, dtype=numpy.float64), (k,)))
    up = numpy.array(numpy.broadcast_to(numpy.asarray(upper, dtype=numpy.float64), (k,)))
    return lo, up


def linear(points, lower_bound, upper_bound):
    """[0,1] -> [lower, upper], ``points*(upper-lower)+lower`` (varsens/scale.py:6-33)."""
    if isinstance(points, _Probe):
        lo, up = _bounds(points.k, lower_bound, upper_bound)
        return _Traced(_cabi.Scale(_cabi.SCALE_LINEAR, lo, up))
    return points * (upper_bound - lower_bound) + lower_bound


def power(points, lower_bound, upper_bound):
    """[0,1] -> [lower, upper] geometrically, ``lower*((upper/lower)**points)`` (varsens/scale.py:35-62)."""
    if isinstance(points, _Probe):
        lo, up = _bounds(points.k, lower_bound, upper_bound)
        return _Traced(_cabi.Scale(_cabi.SCALE_POWER, lo, up))
    return lower_bound * ((upper_bound / lower_bound) ** points)


def percentage(points, reference, percentage=50.0):
    """reference +/- percentage, linear (varsens/scale.py:64-91)."""
    diff = percentage * reference / 100.0
    return linear(points, reference - diff, reference + diff)


def magnitude(points, reference, orders=3.0, base=10.0):
    """reference * base**(+/- orders), geometric (varsens/scale.py:93-122)."""
    factor = base ** orders
    return power(points, reference / factor, reference * factor)


def trace(scaling, k):
    """Lower a user scaling callable to a _cabi.Scale, or None if it is not a pure scale.* call.

    The lowering is verified against the callable itself on a small block of host points, so a
    callable that merely *looks* traceable cannot silently change results.
    """
    if scaling is None:
        return _cabi.IDENTITY
    if isinstance(scaling, _cabi.Scale):
        return scaling
    probe = _Probe(k)
    try:
        out = scaling(probe)
    except Exception:
        return None
    if out is probe:
        desc = _cabi.IDENTITY
    elif isinstance(out, _Traced):
        desc = out.desc
    else:
        return None
    rng = numpy.random.RandomState(12345)
    pts = rng.rand(4, k)
    pts[0, :] = 0.0
    pts[1, :] = 1.0
    try:
        want = numpy.asarray(scaling(pts.copy()), dtype=numpy.float64)
    except Exception:
        return None
    got = desc.apply_numpy(pts)
    if want.shape != got.shape or not numpy.array_equal(want, got):
        return None
    return desc
