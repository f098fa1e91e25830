"""CPU oracle: a restatement of the reference's arithmetic for the primal-dual iLQR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mpx_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and
there only as the checker or the timed CPU baseline, never as the product path.

PARITY UNPINNED: the reference (iit-DLSLab/mpx @ 09c275e5) is JAX + MuJoCo-MJX code that cannot
be imported in this environment (no jax / mujoco / trajax wheels, no network) and whose own tests
hold no golden vectors for this path.  The oracle is therefore pinned only by
(i) the closed-form LQ problem of the reference's ``tests/simple_tests.py:82-111`` solved through a
dense KKT system, (ii) KKT-residual identities, (iii) physical invariants of the MJCF models and
(iv) float64 finite differences of every Jacobian.  See ``tests/test_oracle_*.py``.

Everything is written in PyTorch (float64 by default) in a functional style so that
``torch.func.jacrev`` / ``jacfwd`` reproduce the reference's ``jax.jacobian`` / ``jax.grad`` calls,
including the derivative quirks of ``where`` / ``clip`` / safe-norm patterns.
"""
