"""Unitree H1 whole-body MPC -- same attribute names/values as reference mpx/config/config_h1.py."""
import numpy as np

from ._common import block_diag, table
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives

model_path = table("h1")
contact_frame = ["FL", "RL", "FR", "RR"]
body_name = ["left_ankle_link", "right_ankle_link"]

dt = 0.02
N = 25
mpc_frequency = 50
timer_t = np.array([0.5, 0.5, 0.0, 0.0])
duty_factor = 0.7
step_freq = 1.2
step_height = 0.08
initial_height = 0.9
robot_height = 0.9

p0 = np.array([0, 0, 0.9])
quat0 = np.array([1.0, 0, 0, 0])
q0 = np.array([0, 0, -0.54, 1.2, -0.68,
               0, 0, -0.54, 1.2, -0.68,
               0,
               0.5, 0.25, 0.0, 0.5,
               0.5, -0.25, 0.0, 0.5])
p_legs0 = np.array([0.14738185, 0.20541158, 0.01398883,
                    -0.00253908, 0.2102815, 0.01398485,
                    0.14787466, -0.20581408, 0.01399987,
                    -0.00203967, -0.21088305, 0.0139761])

n_joints = 19
n_contact = len(contact_frame)
n = 13 + 2 * n_joints + 3 * n_contact + 3 * n_contact
m = n_joints
grf_as_state = True
u_ref = np.zeros(m)

Qp = np.diag([0, 0, 1e4])
Qrot = np.diag([1.0, 1, 0]) * 1e3
Qq = np.diag(np.array([4e0] * 10 + [4e1] * 9))
Qdp = np.diag([1.0, 1, 1]) * 1e3
Qomega = np.diag([1.0, 1, 1]) * 1e2
Qdq = np.diag(np.ones(n_joints)) * 1e0
Qtau = np.diag(np.ones(n_joints)) * 1e-3
Qleg = np.diag(np.tile(np.array([1e3, 1e3, 1e5]), n_contact))
Qgrf = np.diag(np.ones(3 * n_contact)) * 1e-3
W = block_diag(Qp, Qrot, Qq, Qdp, Qomega, Qdq, Qleg, Qtau, Qgrf)

use_terrain_estimation = False

cost = mpc_objectives.h1_wb_obj
hessian_approx = mpc_objectives.h1_wb_hessian_gn
dynamics = mpc_dyn_model.h1_wb_dynamics

max_torque = 1000
min_torque = -1000
