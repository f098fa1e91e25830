"""CPU-side checks of the C ABI: the library loads and exports every symbol include/mpxb.h declares."""
import ctypes
import os
import re

from mpx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpxb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpxb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib_built):
    L = ctypes.CDLL(lib_built)
    syms = _declared_symbols()
    assert "mpxb_mpc_step" in syms and "mpxb_model_create" in syms and len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in mpxb.h but not exported"


def test_desc_struct_matches_header(lib_built):
    """sizeof(mpxb_model_desc) as laid out by ctypes must equal the C compiler's."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "mpxb.h"\nint main(){printf("%zu\\n", sizeof(mpxb_model_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert int(out) == ctypes.sizeof(_lib.ModelDesc)


def test_argument_errors_without_gpu(lib_built):
    L = _lib.lib()
    assert L.mpxb_version_string().startswith(b"mpxb")
    h = ctypes.c_void_p()
    assert L.mpxb_model_create(None, ctypes.byref(h)) == -1
    d = _lib.ModelDesc()
    d.version = 99
    assert L.mpxb_model_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert L.mpxb_workspace_bytes(None, 4, 0) == 0


def test_make_desc_from_configs():
    import mpx_b200.config.config_srbd as cs
    import mpx_b200.config.config_go2 as cg
    import mpx_b200.config.config_h1 as ch
    import mpx_b200.config.config_talos as ct
    d = _lib.make_desc(cs)
    assert (d.kind, d.n, d.m, d.N, d.nref, d.npar, d.nw) == (0, 13, 12, 25, 17, 16, 24)
    d = _lib.make_desc(cg)
    assert (d.kind, d.n, d.m, d.nv, d.nbody, d.nref, d.npar, d.nw) == (1, 61, 12, 18, 14, 53, 4, 72)
    assert d.swing_tracking == 1 and abs(d.baumgarte_alpha - 25) < 1e-6
    d = _lib.make_desc(ch)
    assert (d.kind, d.n, d.m, d.nv, d.nref, d.nw) == (2, 75, 19, 25, 60, 93)
    assert list(d.contact_body[:4]) == [d.contact_body[0]] * 2 + [d.contact_body[2]] * 2
    d = _lib.make_desc(ct)
    assert (d.kind, d.n, d.m, d.nv, d.nref, d.npar, d.nw) == (3, 81, 46, 28, 91, 8, 126)


def test_no_cpu_fallback():
    import pytest, torch
    from mpx_b200.primal_dual_ilqr.optimizers import MpcSolver
    import mpx_b200.config.config_srbd as cs
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        MpcSolver(cs, 1)
