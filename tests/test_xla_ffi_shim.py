"""The XLA-FFI adapter (integration/xla_ffi_shim.cc): north_star asks for "a thin XLA-FFI (jax.ffi custom-call) C-ABI
layer".  JAX is not installable in this image, so the shim is compiled against a minimal stand-in for
xla/ffi/api/ffi.h (integration/stub/, the binding API names the shim uses over a plain-C call frame) -- the shim
SOURCE is what a maintainer compiles against jax.ffi.include_dir().

CPU: it compiles, links against libmpxb.so, exports the handler symbols, rejects a malformed call frame and reports
libmpxb's argument errors through ffi::Error.  GPU: a solve driven through the handler equals MpcSolver.solve bitwise."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))
import build_shim  # noqa: E402

U8, S32, F32 = 0, 1, 2


class StubBuffer(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("rank", C.c_int32), ("dims", C.POINTER(C.c_int64))]


class StubAttr(C.Structure):
    _fields_ = [("name", C.c_char_p), ("value", C.c_int64)]


class StubCallFrame(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("nargs", C.c_int32), ("nrets", C.c_int32), ("nattrs", C.c_int32),
                ("args", C.POINTER(StubBuffer)), ("rets", C.POINTER(StubBuffer)), ("attrs", C.POINTER(StubAttr)),
                ("error", C.c_char * 256)]


def _frame(stream, args, rets, attrs):
    """args / rets: lists of (data_ptr, dtype, shape); attrs: dict name -> int.  Returns (frame, keepalive)."""
    keep = []

    def bufs(lst):
        arr = (StubBuffer * max(1, len(lst)))()
        for i, (ptr, dt, shape) in enumerate(lst):
            dims = (C.c_int64 * max(1, len(shape)))(*shape)
            keep.append(dims)
            arr[i] = StubBuffer(ptr, dt, len(shape), dims)
        keep.append(arr)
        return arr
    a, r = bufs(args), bufs(rets)
    at = (StubAttr * max(1, len(attrs)))()
    for i, (k, v) in enumerate(attrs.items()):
        at[i] = StubAttr(k.encode(), int(v))
    keep.append(at)
    f = StubCallFrame(stream, len(args), len(rets), len(attrs), a, r, at, b"")
    return f, keep


@pytest.fixture(scope="module")
def shim(lib_built):
    lib = C.CDLL(build_shim.build())
    for nm in ("mpxb_mpc_step_ffi", "mpxb_shift_extract_ffi"):
        fn = getattr(lib, nm)
        fn.argtypes = [C.POINTER(StubCallFrame)]
        fn.restype = C.c_void_p
    return lib


def test_shim_exports_handlers_and_rejects_bad_frames(shim):
    assert shim.mpxb_mpc_step_ffi(None)                       # null frame -> error
    f, keep = _frame(None, [], [], {"model": 0, "lq_mode": 0, "max_iter": 1})
    assert shim.mpxb_mpc_step_ffi(C.byref(f))                 # no buffers -> binding mismatch
    assert b"does not match" in f.error
    # a well-formed frame with a null model handle reaches the handler and comes back as an ffi::Error
    z = (0, F32, (1, 2, 3))
    rets = [z, z, z, (0, S32, (1,)), (0, S32, (1,)), (0, S32, (1,)), (0, U8, (16,))]
    f, keep = _frame(None, [z, z, (0, F32, (4, 4)), (0, F32, (1, 2)), z, z, z], rets, {"model": 0, "lq_mode": 0, "max_iter": 1})
    assert shim.mpxb_mpc_step_ffi(C.byref(f))
    assert b"null model handle" in f.error
    # wrong dtype in a result slot is caught by the binding, not passed on
    rets[3] = (0, F32, (1,))
    f, keep = _frame(None, [z, z, (0, F32, (4, 4)), (0, F32, (1, 2)), z, z, z], rets, {"model": 1, "lq_mode": 0, "max_iter": 1})
    assert shim.mpxb_mpc_step_ffi(C.byref(f))
    assert b"does not match" in f.error


@pytest.mark.gpu
def test_solve_through_the_ffi_handler_equals_direct_call(shim):
    import torch
    import mpx_b200.config.config_go2 as cfg
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    g = np.load(os.path.join(ROOT, "tests", "golden", "go2_warm.npz"))
    E = g["x0"].shape[0]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous().cuda()
    solver = MpcSolver(cfg, E, limited_memory=True, cost_structure="diagonal")
    ins = [T(g[k]) for k in ("reference", "parameter")] + [T(cfg.W)] + [T(g[k]) for k in ("x0", "X0", "U0", "V0")]
    Xd, Ud, Vd = (t.clone() for t in solver.solve(*ins))
    alpha_d = solver.alpha_index.clone()
    X, U, V = torch.empty_like(Xd), torch.empty_like(Ud), torch.empty_like(Vd)
    ai, it, stt = (torch.full((E,), -7, dtype=torch.int32, device="cuda") for _ in range(3))
    ws = torch.empty(solver.ws_bytes, dtype=torch.uint8, device="cuda")
    buf = lambda t, dt: (t.data_ptr(), dt, tuple(t.shape))
    f, keep = _frame(torch.cuda.current_stream().cuda_stream, [buf(t, F32) for t in ins],
                     [buf(X, F32), buf(U, F32), buf(V, F32), buf(ai, S32), buf(it, S32), buf(stt, S32), buf(ws, U8)],
                     {"model": solver.handle.value, "lq_mode": solver.lq_mode, "max_iter": 1})
    err = shim.mpxb_mpc_step_ffi(C.byref(f))
    torch.cuda.synchronize()
    assert not err, f.error
    assert torch.equal(X, Xd) and torch.equal(U, Ud) and torch.equal(V, Vd)
    assert torch.equal(ai, alpha_d) and (it == 1).all() and (stt == 0).all()
    # a workspace that is too small comes back as an error, not a crash
    f, keep = _frame(torch.cuda.current_stream().cuda_stream, [buf(t, F32) for t in ins],
                     [buf(X, F32), buf(U, F32), buf(V, F32), buf(ai, S32), buf(it, S32), buf(stt, S32), buf(ws[:1024], U8)],
                     {"model": solver.handle.value, "lq_mode": solver.lq_mode, "max_iter": 1})
    assert shim.mpxb_mpc_step_ffi(C.byref(f)) and b"MPXB error -3" in f.error
