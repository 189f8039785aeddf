"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Diamond PPO hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline.  The product path (``diamond-ppo_b200/``) never imports it.

Parity pin: the reference ships no tests/golden vectors, so the
oracle is pinned against outputs of the UNMODIFIED reference executed in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``),
torch 2.11.0 / numpy 2.3.5.  ``tests/test_oracle.py`` checks every restatement
here against those fixtures.
"""
