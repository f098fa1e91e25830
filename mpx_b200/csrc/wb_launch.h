// wb_launch.h -- host-side entry points of the whole-body kernels; each robot class is instantiated in its own
// translation unit (wb_inst_*.cu) so that the classes compile in parallel.
#pragma once
#include <stdlib.h>

#include "model_io.cuh"
#include "wb.cuh"

namespace mpxb {

int wb_quad_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2);
int wb_quad_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st);
int wb_quad_interface(const WbModel& h, const WbModel* dev, const WbiArgs& a, cudaStream_t st);
int wb_h1_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2);
int wb_h1_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st);
int wb_talos_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st, int force_v2);
int wb_talos_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st);

int wb_quad_configure();
int wb_h1_configure();
int wb_talos_configure();

// shared-memory function attributes of the model's robot class on the current device (mpxb_model_create)
inline int wb_launch_configure(const WbModel& h) {
  if (h.nv == 18 && h.nc == 4 && h.nbody <= 14) return wb_quad_configure();
  if (h.nv == 25 && h.nc == 4 && h.nbody <= 21) return wb_h1_configure();
  if (h.nv == 28 && h.nc == 8 && h.nbody <= 32) return wb_talos_configure();
  return MPXB_EUNSUPPORTED;
}

// tree classes: quadrupeds (aliengo, go2), H1, Talos.  MPXB_WB_LINEARIZE=2 forces the forward-mode v2 kernel.
inline int wb_launch_linearize(const WbModel& h, const WbModel* dev, const ModelIO& io, const LinOut& lo, cudaStream_t st) {
  static const int force_v2 = [] { const char* v = getenv("MPXB_WB_LINEARIZE"); return (v && v[0] == '2') ? 1 : 0; }();
  if (h.nv == 18 && h.nc == 4 && h.nbody <= 14) return wb_quad_linearize(h, dev, io, lo, st, force_v2);
  if (h.nv == 25 && h.nc == 4 && h.nbody <= 21) return wb_h1_linearize(h, dev, io, lo, st, force_v2);
  if (h.nv == 28 && h.nc == 8 && h.nbody <= 32) return wb_talos_linearize(h, dev, io, lo, st, force_v2);
  return MPXB_EUNSUPPORTED;
}

inline int wb_launch_trial(const WbModel& h, const WbModel* dev, const ModelIO& io, const TrialOut& to, cudaStream_t st) {
  if (h.nv == 18 && h.nc == 4 && h.nbody <= 14) return wb_quad_trial(h, dev, io, to, st);
  if (h.nv == 25 && h.nc == 4 && h.nbody <= 21) return wb_h1_trial(h, dev, io, to, st);
  if (h.nv == 28 && h.nc == 8 && h.nbody <= 32) return wb_talos_trial(h, dev, io, to, st);
  return MPXB_EUNSUPPORTED;
}

}  // namespace mpxb
