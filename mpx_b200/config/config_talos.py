"""PAL Talos whole-body MPC -- same attribute names/values as reference mpx/config/config_talos.py."""
import numpy as np

from ._common import block_diag, table
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives

model_path = table("talos")
contact_frame = [f"foot_left_{i}" for i in range(1, 5)] + [f"foot_right_{i}" for i in range(1, 5)]
body_name = ["leg_left_6_link", "leg_right_6_link"]

dt = 0.02
N = 25
mpc_frequency = 50
timer_t = np.array([0.5, 0.5, 0.5, 0.5, 0.0, 0.0, 0.0, 0.0])
duty_factor = 0.7
step_freq = 1.2
step_height = 0.08
initial_height = 1.0
robot_height = 1.0

p0 = np.array([0, 0, 1.01])
quat0 = np.array([1.0, 0, 0, 0])
q0 = np.array([0.0, 0.006761,
               0.25847, 0.173046, 0.0002, -0.525366,
               -0.25847, -0.173046, -0.0002, -0.525366,
               0, 0, -0.411354, 0.859395, -0.448041, -0.001708,
               0, 0, -0.411354, 0.859395, -0.448041, -0.001708])
p_legs0 = np.array([0.08592681, 0.145, 0.01690434,
                    0.08592681, 0.025, 0.01690434,
                    -0.11407319, 0.145, 0.01690434,
                    -0.11407319, 0.025, 0.01690434,
                    0.08592681, -0.025, 0.01690434,
                    0.08592681, -0.145, 0.01690434,
                    -0.11407319, -0.025, 0.01690434,
                    -0.11407319, -0.145, 0.01690434])

n_joints = 22
n_contact = len(contact_frame)
n = 13 + 2 * n_joints + 3 * n_contact
m = n_joints + 3 * n_contact
grf_as_state = False
u_ref = np.zeros(m)

Qp = np.diag([0, 0, 1e4])
Qrot = np.diag([1.0, 1, 0]) * 1e3
Qq = np.diag(np.array([1e3] * 2 + [1e1] * 8 + [1e0] * 12))
Qdp = np.diag([1.0, 1, 1]) * 1e3
Qomega = np.diag([1.0, 1, 1]) * 1e2
Qdq = np.diag(np.ones(n_joints)) * 1e0
Qtau = np.diag(np.ones(n_joints)) * 1e-2
Qleg = np.diag(np.tile(np.array([1e5, 1e5, 1e5]), n_contact))
Qgrf = np.diag(np.ones(3 * n_contact)) * 1e-3
W = block_diag(Qp, Qrot, Qq, Qdp, Qomega, Qdq, Qleg, Qtau, Qgrf)

use_terrain_estimation = False

cost = mpc_objectives.talos_wb_obj
hessian_approx = mpc_objectives.talos_wb_hessian_gn
dynamics = mpc_dyn_model.talos_wb_dynamics

max_torque = 1000
min_torque = -1000
