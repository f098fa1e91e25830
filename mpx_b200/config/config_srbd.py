"""Aliengo single-rigid-body MPC -- same attribute names/values as reference mpx/config/config_srbd.py."""
import numpy as np

from ._common import block_diag, table
import mpx_b200.utils.models as mpc_dyn_model
import mpx_b200.utils.objectives as mpc_objectives

model_path = table("aliengo")
joints_name = [f"{leg}_{j}_joint" for leg in ("FL", "FR", "RL", "RR") for j in ("hip", "thigh", "calf")]
contact_frame = ["FL", "FR", "RL", "RR"]
body_name = ["FL_calf", "FR_calf", "RL_calf", "RR_calf"]

dt = 0.02
N = 25
mpc_frequency = 50
whole_body_frequency = 200
timer_t = np.array([0.5, 0.0, 0.0, 0.5])
duty_factor = 0.65
step_freq = 1.35
step_height = 0.065
robot_height = 0.35
clearence_speed = 0.4
use_terrain_estimator = False

p0 = np.array([0, 0, 0.33])
quat0 = np.array([1.0, 0, 0, 0])
q0 = np.array([0.2, 0.8, -1.8, -0.2, 0.8, -1.8, 0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
q0_init = np.array([-0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8, -0.2, 0.8, -1.8])
p_legs0 = np.array([0.27092872, 0.193, 0.0, 0.27092872, -0.193, 0.0,
                    -0.20887128, 0.193, 0.0, -0.20887128, -0.193, 0.0])

n_joints = len(joints_name)
n_contact = len(contact_frame)
n = 13
m = 3 * n_contact
mass = 24.637
inertia = np.array([[0.2310941359705289, -0.0014987128245817424, -0.021400468992761768],
                    [-0.0014987128245817424, 1.4485084687476608, 0.0004641447134275615],
                    [-0.021400468992761768, 0.0004641447134275615, 1.503217877350808]])
grf_ref = np.zeros(3 * n_contact)
u_ref = np.concatenate([grf_ref])

Kp = np.diag(np.tile(np.array([100.0, 100, 100]), n_contact))
Kd = np.diag(np.tile(np.array([5.0, 5, 5]), n_contact))

Qp = np.diag([0, 0, 1e4])
Qrot = np.diag([1e3, 1e3, 0])
Qdp = np.diag([1.0, 1, 1]) * 1e3
Qomega = np.diag([1.0, 1, 1]) * 1e1
Qgrf = np.diag(np.ones(3 * n_contact)) * 1e-2
W = block_diag(Qp, Qrot, Qdp, Qomega, Qgrf)

cost = mpc_objectives.quadruped_srbd_obj
hessian_approx = mpc_objectives.quadruped_srbd_hessian_gn
dynamics = mpc_dyn_model.quadruped_srbd_dynamics
