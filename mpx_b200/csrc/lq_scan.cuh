// lq_scan.cuh -- associative-scan LQ solver (tvlqr_gpu / rollout_gpu).  STUB: filled in after the SRBD slice.
#pragma once
#include "lq.cuh"
namespace mpxb {
inline size_t lq_scan_floats(int, int, int, int) { return 0; }
inline int dispatch_lq_scan(int, int, const LqArgs&, float*, cudaStream_t) { return -2; }
}  // namespace mpxb
