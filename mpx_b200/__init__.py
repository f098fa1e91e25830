"""mpx_b200: B200-native (sm_100a) primal-dual iLQR solve behind the MPX controller API.

Only the hot path of iit-DLSLab/mpx -- ``optimizers.mpc`` behind ``MPCControllerWrapper`` /
``BatchedMPCControllerWrapper`` / ``mpc_wrapper_srbd.BatchedMPCControllerWrapper`` -- is implemented, as
hand-written CUDA kernels reached through the C ABI in ``include/mpxb.h`` (``libmpxb.so``).
There is no CPU fallback: importing the solver without the built library, or calling it with
tensors that are not on a CUDA device, raises.
"""
__version__ = "0.1.0"
