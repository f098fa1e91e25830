// model_io.cuh -- argument blocks shared by the model kernels (SRBD and whole-body): the iterate and problem data
// of one mpc call (optimizers.py:609-616), the LQ data a linearise kernel writes, and the trial-point outputs of
// the filter line search (optimizers.py:518-540).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace mpxb {

constexpr int NALPHA = 11;  // alpha_values = exp2(-arange(11))  (optimizers.py:518)

struct ModelIO {
  int E, N;
  const float *reference, *parameter, *W;
  size_t w_stride;  // floats between consecutive envs' W (0 = shared)
  const float *x0, *X, *U, *V;
  const int32_t* done;
};

struct LinOut {
  float *A, *B, *Q, *R, *M, *q, *r, *c, *cost;
};

struct TrialOut {
  const float *dX, *dU;
  float *ls_theta, *ls_phi;
};

}  // namespace mpxb
