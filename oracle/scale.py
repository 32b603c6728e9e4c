"""Restatement of varsens/scale.py.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Pinned by the reference's own known answers: scale.py doctests (:28-29, :57-58, :85-86,
:116-117) and varsens/tests/test_scaling.py (:9,16,23,31,38,45,53,61) -> tests/test_oracle_scale.py.
"""


def linear(points, lower_bound, upper_bound):
    # varsens/scale.py:33 -- width first, then multiply, then add (two roundings on the point)
    return points * (upper_bound - lower_bound) + lower_bound


def power(points, lower_bound, upper_bound):
    # varsens/scale.py:62 -- ratio first, libm pow, then multiply
    return lower_bound * ((upper_bound / lower_bound) ** points)


def percentage(points, reference, percentage=50.0):
    # varsens/scale.py:90-91
    diff = percentage * reference / 100.0
    return linear(points, reference - diff, reference + diff)


def magnitude(points, reference, orders=3.0, base=10.0):
    # varsens/scale.py:121-122
    factor = base ** orders
    return power(points, reference / factor, reference * factor)
