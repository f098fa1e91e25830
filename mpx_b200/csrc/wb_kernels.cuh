// wb_kernels.cuh -- whole-body (MuJoCo-style tree) model kernels.  STUB: filled in after the SRBD slice.
#pragma once
#include "../../include/mpxb.h"
#include "srbd_kernels.cuh"

namespace mpxb {
struct WbModel { int nbody; };
inline int wb_build_model(const mpxb_model_desc&, WbModel*) { return MPXB_EUNSUPPORTED; }
inline int wb_launch_linearize(const WbModel&, const WbModel*, const ModelIO&, const LinOut&, cudaStream_t) { return MPXB_EUNSUPPORTED; }
inline int wb_launch_trial(const WbModel&, const WbModel*, const ModelIO&, const TrialOut&, cudaStream_t) { return MPXB_EUNSUPPORTED; }
}  // namespace mpxb
