"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the TauFactor steady-state diffusion solve.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as the
checker or the timed CPU baseline.  ``taufactor_b200`` never imports this package.

Parity is PINNED: ``tests/test_oracle_golden.py`` checks this restatement against golden vectors
generated in the build container by importing the unmodified reference (``/root/reference``,
taufactor v1.2.1) -- see ``tests/golden/make_golden.py`` -- and against the known answers of the
reference's own test-suite (``/root/reference/tests/test_taufactor.py``).

``oracle/_ref``: the reference is a pure-Python package (no C/C++/CUDA sources),
so there is nothing to compile; the real reference is used only in this container to generate the
committed fixtures.
"""
