// wb_inst_quad.cu -- whole-body kernels instantiated for the "quad" tree class (nbody <= 14, nv = 18, nc = 4).
#include <cuda_runtime.h>

#include "wb_kernels.cuh"
#include "wb_launch.h"

namespace mpxb {

int wb_quad_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2) {
  return wb_class_linearize<14, 18, 4, 4, 192, 64>(h, dev, io, lo, st, force_v2);
}
int wb_quad_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st) {
  return wb_class_trial<14, 18, 4>(h, dev, io, to, st);
}

int wb_quad_interface(const WbModel& h, const WbModel* dev, const WbiArgs& a, cudaStream_t st) {
  return wb_class_interface<14, 18, 4>(h, dev, a, st);
}

int wb_quad_configure() { return wb_class_configure<14, 18, 4, 4, 192, 64>(true); }

}  // namespace mpxb
