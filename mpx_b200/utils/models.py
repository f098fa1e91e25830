"""Dynamics descriptors: the names of reference ``mpx/utils/models.py`` bound to CUDA model kinds.

In the reference these names are JAX functions that ``optimizers.mpc`` traces (models.py:6-217).  Here
they are descriptors: the config modules keep ``config.dynamics = mpc_dyn_model.quadruped_wb_dynamics``
unchanged, and the wrappers translate the descriptor to the ``kind`` field of ``mpxb_model_desc``
(include/mpxb.h), which selects the hand-written kernel family.
"""
from dataclasses import dataclass

KIND_SRBD, KIND_QUAD_WB, KIND_H1_WB, KIND_TALOS_WB = 0, 1, 2, 3


@dataclass(frozen=True)
class DynamicsKind:
    name: str
    kind: int
    baumgarte_alpha: float = 0.0   # models.py:93 (25), :146 (5)
    contact_reg: float = 0.0       # models.py:155 (1e-3 on J^T M^-1 J for H1)

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self.name} is a kernel descriptor; the dynamics run inside libmpxb.so (mpxb_mpc_step)")


quadruped_srbd_dynamics = DynamicsKind("quadruped_srbd_dynamics", KIND_SRBD)
quadruped_wb_dynamics = DynamicsKind("quadruped_wb_dynamics", KIND_QUAD_WB, 25.0, 0.0)
h1_wb_dynamics = DynamicsKind("h1_wb_dynamics", KIND_H1_WB, 5.0, 1e-3)
talos_wb_dynamics = DynamicsKind("talos_wb_dynamics", KIND_TALOS_WB)
