"""The FP32 breakdown case of tests/golden/go2_illcond.npz on the oracle (CPU): an iterate from bench.py's 2048-environment
closed loop (environment 1828 of seed 0, third tick; dumped by scripts/dump_nonfinite.py).  In float32 the Riccati
recursion's R + B'PB stops being positive definite, so the Cholesky of `lqr_step` (primal_tvlqr.py:38-44 through
linalg_helpers.solve_symmetric_positive_definite_system) fails -- under JAX that is NaNs in the outputs, which the
wrappers detect (mpc_wrapper.py:96,280).  In float64 the same inputs solve.  The GPU side of the pair is
tests/test_gpu_wb.py::test_wb_fp32_breakdown_is_reported_in_band."""
import os

import numpy as np
import pytest
import torch

import mpx_b200.config.config_go2 as cfg
from oracle.problems import Problem

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "go2_illcond.npz"))
ARGS = [G[k] for k in ("reference", "parameter", "x0", "X0", "U0", "V0")]


def test_float64_oracle_solves_it():
    det = {}
    X, U, V = Problem(cfg).mpc(True, *ARGS, detail=det)
    assert bool(torch.isfinite(X).all()) and det["alpha_index"] == 1
    assert float(np.abs(np.asarray(det["dX"])).max()) > 50.0  # a far-from-feasible iterate: a huge Newton step


def test_float32_oracle_loses_positive_definiteness():
    try:
        X, U, V = Problem(cfg, dtype=torch.float32).mpc(True, *ARGS)
    except torch.linalg.LinAlgError as ex:  # torch raises where jnp.linalg.cholesky returns NaN
        assert "positive-definite" in str(ex)
        return
    assert not bool(torch.isfinite(X).all())
