"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the molann hot path.

Nothing under ``molann_b200/`` (the product) may import this package.  Allowed
importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- always as the checker or as the timed
CPU baseline, never as the thing shipped.

Parity status: PINNED.  The reference's own tests hold no golden values
(test/test_molann.py has no assert), so the restatement in ``restatement.py`` is
pinned against outputs of the unmodified reference executed in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle.py`` re-runs
the live comparison whenever ``/root/reference`` is importable).
"""
