"""TEST INFRASTRUCTURE ONLY — CPU restatement of ColdRec's scoring / propagation / generator hot path.

Nothing under ``coldrec_b200/`` may import this package.  The only allowed importers are
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` (there only as the checker or the timed CPU baseline, never as the product path).

Parity status: the reference (YuanchenBei/ColdRec) ships no tests, golden vectors or fixtures, so
parity is pinned instead on outputs of the reference's *own code* imported in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks this
restatement against those vectors.
"""
