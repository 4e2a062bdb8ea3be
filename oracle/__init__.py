"""CPU oracle for the punchclock per-step hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of the reference algorithm
(`punchclock` v0.8.5) for propagate / visibility / UKF.  It exists so that the
CUDA path in ``clockpunch_b200`` can be checked against something that runs
without a GPU and without ``/root/reference``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``clockpunch_b200/`` imports it; the product path raises if the CUDA library
is missing instead of falling back to this code.

Pinning status (see DESIGN.md "Oracle"):
  * dynamics / transforms / UKF  -- pinned against outputs of the live
    reference (imported from /root/reference in the build container by
    ``tests/golden/make_golden.py``; vectors committed under tests/golden/).
  * visibility (``satvis==0.2.4`` arithmetic, not vendored in the reference,
    pyproject.toml:21) -- PARITY UNPINNED except for the one geometric known
    answer in the reference's tests/environment/test_env.py:31-41,132-136.
"""
