"""TEST INFRASTRUCTURE ONLY — CPU restatement of kllloyd/GPSurvival's GP hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and R is not installed in this image (F1), so the oracle cannot be
pinned against the reference's own outputs.  It is pinned instead by
(1) an independent 50-digit mpmath restatement (``oracle/mp_oracle.py``),
(2) central finite differences of the NLML against the analytic gradient,
(3) closed forms and algebraic invariants (tests/test_oracle_*.py).
"""
