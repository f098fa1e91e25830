"""Per-knot, per-block relative error -- the metric of the parity tests (TEST INFRASTRUCTURE).

north_star: "trajectories, gains and multipliers within 1e-4 relative in FP32".  A single max-norm over a whole array
lets the largest entry (a ground-reaction force of 50 N, a value-function entry of 5e6) hide errors in every smaller
block, so every comparison here is made block by block:

* the leading axis of every array is the knot; each knot is judged on its own;
* a state axis is split into the physical segments of the reference's state layout (objectives.py:79-86,
  models.py:61,115): position | quaternion | joints | linear velocity | angular velocity | joint speeds | feet | GRFs;
  a control axis into joint torques | GRFs (Talos);
* for block b of knot t:  err = max|a_b - ref_b| / max(max|ref_b|, FLOOR * max|ref_t|, ABS_FLOOR)
  -- the block's own magnitude, floored at FLOOR (1e-3) of that KNOT's largest entry (an FP32 sum cannot be more
  accurate than 1e-7 of its largest term, so a block that is structurally ~0 next to a large one is judged
  against the knot) and at an absolute 1e-3 for identically-zero knots (M == 0 for every shipped config).

* the dynamics defect ``c_{t+1} = f(X_t, U_t) - X_{t+1}`` (``c_0 = x0 - X_0``) is a difference of two state vectors, so
  near a feasible iterate it is pure cancellation (a 1e-3 N defect between two 200 N ground-reaction forces); its
  blocks are judged against the magnitude of the state block they are a defect OF (``magnitude=`` the iterate X).

``block_rel_err`` returns the worst block; the golden fixtures store the float32 oracle's own error under the SAME
metric as the FP32 budget (tests/golden/make_golden.py).
"""
import numpy as np

FLOOR = 1e-3
ABS_FLOOR = 1e-3

# axes after the knot axis: 'x' = state, 'u' = control
AXES = {"A": "xx", "B": "xu", "Q": "xx", "R": "uu", "M": "xu", "q": "x", "r": "u", "c": "x", "K": "ux", "k": "u",
        "P": "xx", "p": "x", "dX": "x", "dU": "u", "dV": "x", "X": "x", "U": "u", "V": "x"}


def segments(cfg):
    """(state segments, control segments) as lists of slices, from a config module."""
    if getattr(cfg, "n", None) == 13 and not hasattr(cfg, "grf_as_state"):
        xs = [(0, 3), (3, 7), (7, 10), (10, 13)]
        us = [(0, cfg.m)]
    else:
        nj, nc, n, m = cfg.n_joints, cfg.n_contact, cfg.n, cfg.m
        xs = [(0, 3), (3, 7), (7, 7 + nj), (7 + nj, 10 + nj), (10 + nj, 13 + nj), (13 + nj, 13 + 2 * nj),
              (13 + 2 * nj, 13 + 2 * nj + 3 * nc)]
        if n > 13 + 2 * nj + 3 * nc:
            xs.append((13 + 2 * nj + 3 * nc, n))
        us = [(0, nj)] + ([(nj, m)] if m > nj else [])
    return [slice(a, b) for a, b in xs], [slice(a, b) for a, b in us]


def segments_generic(n, m):
    """One block per axis (generic LQ problems without a physical layout)."""
    return [slice(0, n)], [slice(0, m)]


def block_rel_err(name, a, ref, segs, axes=None, magnitude=None):
    """Worst per-knot, per-block relative error of ``a`` against ``ref`` ([knot, ...]); see the module docstring.
    ``magnitude`` (same shape as ``ref``, optional): the quantity whose blocks set the scale instead of ``ref``'s own
    (used for the defect ``c``)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (name, a.shape, ref.shape)
    axes = AXES[name] if axes is None else axes
    xs, us = segs
    seg = {"x": xs, "u": us}
    mag = ref if magnitude is None else np.maximum(np.abs(ref), np.abs(np.asarray(magnitude, dtype=np.float64)))
    worst = 0.0
    for t in range(ref.shape[0]):
        at, rt, mt = a[t], ref[t], mag[t]
        knot = float(np.abs(mt).max()) if mt.size else 0.0
        floor = max(FLOOR * knot, ABS_FLOOR)
        if len(axes) == 1:
            blocks = [(s,) for s in seg[axes[0]]]
        else:
            blocks = [(s0, s1) for s0 in seg[axes[0]] for s1 in seg[axes[1]]]
        for b in blocks:
            rb, ab = rt[b], at[b]
            d = np.abs(ab - rb)
            if not np.all(np.isfinite(d)):
                return float("inf")
            e = float(d.max()) / max(float(np.abs(mt[b]).max()), floor)
            worst = max(worst, e)
    return worst


def scalar_rel_err(a, ref):
    a, ref = float(a), float(ref)
    return abs(a - ref) / max(abs(ref), ABS_FLOOR)
