"""Objective descriptors: the names of reference ``mpx/utils/objectives.py`` bound to CUDA cost kinds.

``config.cost`` / ``config.hessian_approx`` keep their reference names (objectives.py:15-450); the kernels
implement the matching cost, gradient and Gauss-Newton Hessian.  ``functools.partial(quadruped_wb_obj, True)``
(config_go2.py:67-68) still works: the bound ``swing_tracking`` flag is read back by the wrapper.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class ObjectiveKind:
    name: str
    kind: int

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self.name} is a kernel descriptor; the cost runs inside libmpxb.so (mpxb_mpc_step)")


quadruped_srbd_obj = ObjectiveKind("quadruped_srbd_obj", 0)
quadruped_srbd_hessian_gn = ObjectiveKind("quadruped_srbd_hessian_gn", 0)
quadruped_wb_obj = ObjectiveKind("quadruped_wb_obj", 1)
quadruped_wb_hessian_gn = ObjectiveKind("quadruped_wb_hessian_gn", 1)
h1_wb_obj = ObjectiveKind("h1_wb_obj", 2)
h1_wb_hessian_gn = ObjectiveKind("h1_wb_hessian_gn", 2)
talos_wb_obj = ObjectiveKind("talos_wb_obj", 3)
talos_wb_hessian_gn = ObjectiveKind("talos_wb_hessian_gn", 3)
