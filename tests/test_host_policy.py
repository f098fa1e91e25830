"""Host-side logic that needs no GPU: which LQ kernel the wrappers bind, and that the Python constants match the header."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lq_mode_constants_match_header():
    from mpx_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mpxb.h")).read()
    for name, val in [("MPXB_LQ_SEQUENTIAL", _lib.LQ_SEQUENTIAL), ("MPXB_LQ_SCAN", _lib.LQ_SCAN),
                      ("MPXB_LQ_SCAN_SERIAL", _lib.LQ_SCAN_SERIAL), ("MPXB_LQ_SCAN_PARALLEL", _lib.LQ_SCAN_PARALLEL)]:
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, hdr)
        assert m and int(m.group(1)) == val, name


def test_wrapper_lq_mode_policy(monkeypatch):
    from mpx_b200 import _lib
    from mpx_b200.primal_dual_ilqr.optimizers import wrapper_lq_mode
    monkeypatch.delenv("MPXB_LQ", raising=False)
    # limited_memory=True is literal (examples/multi_env.py:44)
    assert wrapper_lq_mode(True, 1, 81, 200) == _lib.LQ_SEQUENTIAL
    assert wrapper_lq_mode(True, 1024, 61, 25) == _lib.LQ_SEQUENTIAL
    # limited_memory=False: the parallel scan where it was measured faster ...
    assert wrapper_lq_mode(False, 1, 13, 100) == _lib.LQ_SCAN_PARALLEL     # Aliengo SRBD, horizon beyond the small kernel
    assert wrapper_lq_mode(False, 1, 81, 25) == _lib.LQ_SCAN_PARALLEL      # Talos
    assert wrapper_lq_mode(False, 1, 81, 200) == _lib.LQ_SCAN_PARALLEL     # Talos, long horizon
    assert wrapper_lq_mode(False, 1, 61, 100) == _lib.LQ_SCAN_PARALLEL     # barrel-roll horizon
    # ... and the sequential kernel elsewhere
    assert wrapper_lq_mode(False, 1, 13, 25) == _lib.LQ_SEQUENTIAL         # Aliengo SRBD single instance (lq_small kernel)
    assert wrapper_lq_mode(False, 1, 61, 25) == _lib.LQ_SEQUENTIAL         # Aliengo whole-body single instance
    assert wrapper_lq_mode(False, 1, 75, 25) == _lib.LQ_SEQUENTIAL         # H1
    assert wrapper_lq_mode(False, 1024, 61, 25) == _lib.LQ_SEQUENTIAL      # large batches
    assert wrapper_lq_mode(False, 16, 81, 25) == _lib.LQ_SEQUENTIAL        # more than one wave of CTAs per level
    # environment overrides
    monkeypatch.setenv("MPXB_LQ", "scan")
    assert wrapper_lq_mode(False, 1024, 61, 25) == _lib.LQ_SCAN
    assert wrapper_lq_mode(True, 1, 13, 25) == _lib.LQ_SEQUENTIAL
    monkeypatch.setenv("MPXB_LQ", "sequential")
    assert wrapper_lq_mode(False, 1, 13, 25) == _lib.LQ_SEQUENTIAL
