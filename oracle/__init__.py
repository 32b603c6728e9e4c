"""CPU oracle for the varsens Saltelli hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``varsens_b200/`` may import this package.  The only legal callers are
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker or the timed CPU arm.

PARITY STATUS: **parity unpinned** at bit level for the two third-party generators
(ghalton's Halton, QuantLib's SobolRsg): neither is vendored in /root/reference, neither is
installable here, and no reference test pins a single generated value.  Their
published algorithms are restated in ``halton.py`` / ``sobol.py``.  Everything that *is* in
/root/reference (saltelli.py, scale.py) is restated line-for-line in semantics and pinned by
the reference's own known-answer tests (scale doctests, test_scaling.py, g-function closed
forms), see tests/test_oracle_*.py.
"""
