"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy / scipy.sparse) of the compas_tno hot path, used as the
checker for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``compas_tno_b200`` never imports anything from here.

Parity status: PINNED.  ``oracle/callbacks_np.py`` is checked (tests/test_oracle.py)
against golden vectors produced in the build container by importing the
reference's own callbacks from ``/root/reference/src`` (``oracle/refstub.py`` +
``tests/golden/make_golden.py``), and against the known answers stored in the
reference's ``data/*.json`` fixtures (z(q), q = B q_ind + d, reactions, dome and
cross-vault envelope closed forms).

Parity UNPINNED (stated, not hidden): the arithmetic that lives in the
un-vendored ``compas_tna`` package (``Envelope.compute_bounds`` and
``compute_bounds_derivatives``).  ``oracle/envelopes_np.py`` restates the closed
forms recovered from the fixtures; ub/lb are pinned at the fixtures' thickness,
their thickness-derivatives are pinned only by finite differences.
"""
