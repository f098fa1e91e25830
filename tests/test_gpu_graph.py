"""CUDA-graph capture of the solve path (the C-ABI calls are stream-ordered, allocation-free and
sync-free, so a whole tick can be captured once and replayed): replayed results must equal the eager launches bitwise."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_srbd_graph_run_equals_run():
    """mpc_wrapper_srbd.graph_run == run over 6 closed-loop ticks (states, multipliers, GRFs, gait timer)."""
    import mpx_b200.config.config_srbd as cfg
    from mpx_b200.utils.mpc_wrapper_srbd import BatchedMPCControllerWrapper
    from oracle.problems import Problem, synthetic_inputs
    E = 4
    inp = synthetic_inputs(Problem(cfg), E, seed=5)
    outs = []
    for graphed in (False, True):
        mpc = BatchedMPCControllerWrapper(cfg, E)
        x0 = torch.as_tensor(inp["x0"], dtype=torch.float32).cuda()
        feet = torch.as_tensor(inp["parameter"][:, 0, 4:], dtype=torch.float32).cuda()
        cmd = torch.tensor([[0.3, 0.05, 0, 0, 0, 0.1, cfg.robot_height]] * E, dtype=torch.float32).cuda()
        contact = torch.ones(E, 4).cuda()
        mpc.liftoff.copy_(feet)
        trace = []
        for _ in range(6):
            (mpc.graph_run if graphed else mpc.run)(x0, cmd, feet, contact)
            x0 = mpc.X[:, 1, :].clone()
            feet = mpc.parameter[:, 0, 4:].clone()
            trace.append([t.clone() for t in (mpc.X, mpc.U, mpc.V, mpc.grf, mpc.contact_time, mpc.liftoff, mpc.batch_X0)])
        torch.cuda.synchronize()
        outs.append(trace)
    for a, b in zip(*outs):
        for ta, tb in zip(a, b):
            assert torch.isfinite(ta).all()
            assert torch.equal(ta, tb)


def test_wb_mpc_step_captured_in_graph():
    """mpxb_mpc_step + mpxb_shift_extract (Go2 whole-body, both LQ kernels) captured in a CUDA graph and replayed."""
    cfg = importlib.import_module("mpx_b200.config.config_go2")
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    path = os.path.join(GOLDEN, "go2_warm.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture missing")
    g = np.load(path)
    E = g["x0"].shape[0]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    for limited_memory in (True, False):
        s = MpcSolver(cfg, E, limited_memory=limited_memory)
        ref, par, W, x0, u_ref = T(g["reference"]), T(g["parameter"]), T(cfg.W), T(g["x0"]), T(cfg.u_ref)
        X0, U0, V0 = T(g["X0"]), T(g["U0"]), T(g["V0"])
        X, U, V = torch.empty_like(X0), torch.empty_like(U0), torch.empty_like(V0)

        def tick():
            s.solve(ref, par, W, x0, X0, U0, V0, out=(X, U, V))
            s.shift_extract(X, U, V, x0, u_ref, X0, U0, V0, 1, want_cmd=False)
        init = [t.clone() for t in (X0, U0, V0)]
        eager = []
        for _ in range(3):
            tick()
            eager.append([t.clone() for t in (X, U, V, X0, s.alpha_index)])
        for t, t0 in zip((X0, U0, V0), init):
            t.copy_(t0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                tick()
        torch.cuda.current_stream().wait_stream(side)
        for k in range(3):
            graph.replay()
            for ta, tb in zip(eager[k], (X, U, V, X0, s.alpha_index)):
                assert torch.equal(ta, tb)
        torch.cuda.synchronize()


def test_lq_tail_overlap_is_bitwise_neutral_and_capturable():
    """Batches that leave the last wave of LQ CTAs partly empty run the full waves and the remainder as two launches and
    put the alpha = 1 trial points of the first slice on the handle's side stream (csrc/mpxb.cu, `EA`).  Same kernels on
    the same data: every output must equal the single-launch order (MPXB_OPT_EAGER_LINE_SEARCH never splits and is
    bitwise equal to the lazy search), and the fork / join must survive CUDA-graph capture."""
    cfg = importlib.import_module("mpx_b200.config.config_go2")
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    path = os.path.join(GOLDEN, "go2_warm.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture missing")
    g = np.load(path)
    E0 = g["x0"].shape[0]
    E = 148 * 4 + 100                      # one full wave of the Go2 LQ kernel + a remainder: splits
    reps = -(-E // E0)
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    R = lambda a: T(a).repeat((reps,) + (1,) * (np.asarray(a).ndim - 1))[:E].contiguous()
    ref, par, x0, X0, U0, V0 = (R(g[k]) for k in ("reference", "parameter", "x0", "X0", "U0", "V0"))
    # make the environments differ (and some reject alpha = 1): perturb the iterate
    gen = torch.Generator(device="cuda").manual_seed(1)
    X0 = X0 + 0.02 * torch.randn(X0.shape, generator=gen, device="cuda") * (torch.arange(E, device="cuda") % 5)[:, None, None]
    W = T(cfg.W)
    outs = {}
    for name, eager_ls in (("split", False), ("eager", True)):
        s = MpcSolver(cfg, E, limited_memory=True, eager_line_search=eager_ls)
        X, U, V = torch.empty_like(X0), torch.empty_like(U0), torch.empty_like(V0)
        s.solve(ref, par, W, x0, X0, U0, V0, out=(X, U, V))
        torch.cuda.synchronize()
        outs[name] = [t.clone() for t in (X, U, V, s.alpha_index)]
        if name == "split":
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            X.zero_(); U.zero_(); V.zero_()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    s.solve(ref, par, W, x0, X0, U0, V0, out=(X, U, V))
            torch.cuda.current_stream().wait_stream(side)
            graph.replay()
            torch.cuda.synchronize()
            outs["graph"] = [t.clone() for t in (X, U, V, s.alpha_index)]
    assert int((outs["split"][3] > 0).sum()) > 0, "the case should exercise the second pass as well"
    for a, b, c in zip(outs["split"], outs["eager"], outs["graph"]):
        assert torch.equal(a, b)
        assert torch.equal(a, c)
