"""Aliengo barrel roll, solved offline (MPCControllerWrapper.runOffline) -- same attribute names/values as reference
mpx/config/config_barrel_roll.py: horizon N = 100 at dt = 0.01, no swing tracking (the foot-position residuals are
masked by the contact flags), `reference` = the pre-planned roll trajectory instead of the gait reference generator."""
from functools import partial

import numpy as np

from ._common import block_diag, table
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives
import mpx_b200.utils.mpc_utils as mpc_utils

model_path = table("aliengo")
contact_frame = ["FL", "FR", "RL", "RR"]
body_name = ["FL_calf", "FR_calf", "RL_calf", "RR_calf"]

dt = 0.01
N = 100
mpc_frequency = 50

timer_t = np.array([0.5, 0.0, 0.0, 0.5])
duty_factor = 0.65
step_freq = 1.35
step_height = 0.12
initial_height = 0.1
robot_height = 0.33
grf_as_state = True

p0 = np.array([0, 0, 0.33])
quat0 = np.array([1.0, 0, 0, 0])
q0 = np.array([0.2, 0.8, -1.8, -0.2, 0.8, -1.8, 0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
q0_init = np.array([-0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
p_legs0 = np.array([0.27092872, 0.174, -0.31, 0.27092872, -0.174, -0.31,
                    -0.20887128, 0.174, -0.31, -0.20887128, -0.174, -0.31])

n_joints = 12
n_contact = len(contact_frame)
n = 13 + 2 * n_joints + 6 * n_contact
m = n_joints
tau_ref = np.zeros(n_joints)
u_ref = np.concatenate([tau_ref])

Qp = np.diag([0, 0, 5e4])
Qrot = np.diag([100.0, 100, 100])
Qq = np.diag(np.ones(n_joints)) * 1e2
Qdp = np.diag([100.0, 100, 100])
Qomega = np.diag([1.0, 1, 1]) * 1e2
Qdq = np.diag(np.ones(n_joints)) * 1e-1
Qtau = np.diag(np.ones(n_joints)) * 1e-1
Q_grf = np.diag(np.ones(3 * n_contact)) * 0
Qleg = np.diag(np.tile(np.array([1e3, 1e3, 5e4]), n_contact))
W = block_diag(Qp, Qrot, Qq, Qdp, Qomega, Qdq, Qleg, Qtau, Q_grf)

use_terrain_estimation = True

cost = partial(mpc_objectives.quadruped_wb_obj, False)
hessian_approx = partial(mpc_objectives.quadruped_wb_hessian_gn, False)
dynamics = mpc_dyn_model.quadruped_wb_dynamics
reference = mpc_utils.reference_barell_roll
max_torque = 40
min_torque = -40
