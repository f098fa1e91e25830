"""Aliengo whole-body MPC -- same attribute names/values as reference mpx/config/config_aliengo.py."""
from functools import partial

import numpy as np

from ._common import block_diag, table
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives

model_path = table("aliengo")
contact_frame = ["FL", "FR", "RL", "RR"]
body_name = ["FL_calf", "FR_calf", "RL_calf", "RR_calf"]

dt = 0.02
N = 25
mpc_frequency = 50
timer_t = np.array([0.5, 0.0, 0.0, 0.5])
duty_factor = 0.65
step_freq = 1.35
step_height = 0.14
initial_height = 0.1
robot_height = 0.36

p0 = np.array([0, 0, robot_height])
quat0 = np.array([1.0, 0, 0, 0])
q0 = np.array([0.2, 0.8, -1.8, -0.2, 0.8, -1.8, 0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
q0_init = np.array([-0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
p_legs0 = np.array([0.27092872, 0.193, 0.0, 0.27092872, -0.193, 0.0,
                    -0.20887128, 0.193, 0.0, -0.20887128, -0.193, 0.0])

n_joints = 12
n_contact = len(contact_frame)
n = 13 + 2 * n_joints + 6 * n_contact
m = n_joints
grf_as_state = True
u_ref = np.zeros(m)

Qp = np.diag([0, 0, 1e4])
Qrot = np.diag([1000.0, 1000, 0])
Qq = np.diag(np.ones(n_joints)) * 1e-1
Qdp = np.diag([1.0, 1, 1]) * 5e3
Qomega = np.diag([1.0, 1, 1]) * 1e2
Qdq = np.diag(np.ones(n_joints)) * 1e-1
Qtau = np.diag(np.ones(n_joints)) * 1e-1
Q_grf = np.diag(np.ones(3 * n_contact)) * 1e-2
Qleg = np.diag(np.tile(np.array([1e4, 1e4, 1e5]), n_contact))
W = block_diag(Qp, Qrot, Qq, Qdp, Qomega, Qdq, Qleg, Qtau, Q_grf)

use_terrain_estimation = True

cost = partial(mpc_objectives.quadruped_wb_obj, True)
hessian_approx = partial(mpc_objectives.quadruped_wb_hessian_gn, True)
dynamics = mpc_dyn_model.quadruped_wb_dynamics
max_torque = 35
min_torque = -35
