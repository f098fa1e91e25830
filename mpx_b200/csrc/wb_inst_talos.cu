// wb_inst_talos.cu -- whole-body kernels instantiated for the "talos" tree class (nbody <= 32, nv = 28, nc = 8).
#include <cuda_runtime.h>

#include "wb_kernels.cuh"
#include "wb_launch.h"

namespace mpxb {

int wb_talos_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2) {
  return wb_class_linearize<32, 28, 8, 10, 640, 128>(h, dev, io, lo, st, force_v2);
}
int wb_talos_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st) {
  return wb_class_trial<32, 28, 8>(h, dev, io, to, st);
}

int wb_talos_configure() { return wb_class_configure<32, 28, 8, 10, 640, 128>(false); }

}  // namespace mpxb
