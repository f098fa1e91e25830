// wb_inst_h1.cu -- whole-body kernels instantiated for the "h1" tree class (nbody <= 21, nv = 25, nc = 4).
#include <cuda_runtime.h>

#include "wb_kernels.cuh"
#include "wb_launch.h"

#ifndef MPXB_H1_NT
#define MPXB_H1_NT 128   // threads per knot of the linearise kernel (96 measured slower, see MPXB_LIN3_MINB)
#endif

namespace mpxb {

int wb_h1_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2) {
  return wb_class_linearize<21, 25, 4, 6, 320, MPXB_H1_NT>(h, dev, io, lo, st, force_v2);
}
int wb_h1_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st) {
  return wb_class_trial<21, 25, 4>(h, dev, io, to, st);
}

int wb_h1_configure() { return wb_class_configure<21, 25, 4, 6, 320, MPXB_H1_NT>(false); }

}  // namespace mpxb
