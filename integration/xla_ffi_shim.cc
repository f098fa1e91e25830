// xla_ffi_shim.cc -- XLA-FFI (jax.ffi custom call) adapter over the C ABI of libmpxb.so (include/mpxb.h).
//
// What a maintainer of the reference (iit-DLSLab/mpx) builds once against the real header and links with -lmpxb:
//     g++ -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//         integration/xla_ffi_shim.cc -Lmpx_b200 -lmpxb -o libmpxb_xla.so
// and binds with jax.ffi.register_ffi_target / jax.ffi.ffi_call.  It replaces the jitted
// callable  work = partial(optimizers.mpc, cost, dynamics, hessian_approx, limited_memory)  of the reference
// (mpx/utils/mpc_wrapper.py:74,244,249; mpx/utils/mpc_wrapper_srbd.py:68,77).
//
// JAX and the XLA headers are not installable in this image, so the repository also carries a minimal stand-in for
// xla/ffi/api/ffi.h (integration/stub/, the binding API names this file uses and nothing else);
// tests/test_xla_ffi_shim.py compiles THIS file against it, checks the exported handler symbols and (on a GPU) drives a
// solve through the handler.  The file compiles unchanged against either header.
#include <cstdint>
#include <string>

#include "xla/ffi/api/ffi.h"

#include "mpxb.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Failure(const char* what, int rc) {
  return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(what) + " failed with MPXB error " + std::to_string(rc));
}

// optimizers.mpc: (reference, parameter, W, x0, X_in, U_in, V_in) -> (X, U, V) [+ alpha_index, iters, status, workspace]
// Every buffer carries the leading environment axis (jax.vmap of the call: vmap_method="expand_dims"); W is [nw, nw]
// (shared) or [E, nw, nw].  `model` is the mpxb_model* created once per controller, passed as an int64 attribute.
ffi::Error MpcStep(cudaStream_t stream, int64_t model, int32_t lq_mode, int32_t max_iter,
                   ffi::Buffer<ffi::F32> reference, ffi::Buffer<ffi::F32> parameter, ffi::Buffer<ffi::F32> W,
                   ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> X_in, ffi::Buffer<ffi::F32> U_in,
                   ffi::Buffer<ffi::F32> V_in, ffi::ResultBuffer<ffi::F32> X, ffi::ResultBuffer<ffi::F32> U,
                   ffi::ResultBuffer<ffi::F32> V, ffi::ResultBuffer<ffi::S32> alpha_index,
                   ffi::ResultBuffer<ffi::S32> iters, ffi::ResultBuffer<ffi::S32> status,
                   ffi::ResultBuffer<ffi::U8> workspace) {
  const mpxb_model* mdl = reinterpret_cast<const mpxb_model*>(static_cast<intptr_t>(model));
  if (mdl == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mpxb_mpc_step: null model handle");
  if (x0.dimensions().size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mpxb_mpc_step: x0 must be [E, n]");
  const int E = static_cast<int>(x0.dimensions()[0]);
  const auto wd = W.dimensions();
  const size_t w_stride = wd.size() == 3 ? static_cast<size_t>(wd[1]) * static_cast<size_t>(wd[2]) : 0;
  const int rc = mpxb_mpc_step(mdl, stream, E, lq_mode, max_iter, reference.typed_data(), parameter.typed_data(),
                               W.typed_data(), w_stride, x0.typed_data(), X_in.typed_data(), U_in.typed_data(),
                               V_in.typed_data(), X->typed_data(), U->typed_data(), V->typed_data(),
                               alpha_index->typed_data(), iters->typed_data(), status->typed_data(),
                               workspace->typed_data(), workspace->size_bytes());
  return rc == MPXB_OK ? ffi::Error::Success() : Failure("mpxb_mpc_step", rc);
}

// update_and_extract_helper (mpc_wrapper.py:79-98): warm-start shift, NaN guard, command extraction.  X0 / U0 / V0 are
// updated in place (input_output_aliases on the JAX side); tau / q / dq are results.
ffi::Error ShiftExtract(cudaStream_t stream, int64_t model, int32_t shift, ffi::Buffer<ffi::F32> X, ffi::Buffer<ffi::F32> U,
                        ffi::Buffer<ffi::F32> V, ffi::Buffer<ffi::F32> x0, ffi::Buffer<ffi::F32> u_ref,
                        ffi::ResultBuffer<ffi::F32> X0, ffi::ResultBuffer<ffi::F32> U0, ffi::ResultBuffer<ffi::F32> V0,
                        ffi::ResultBuffer<ffi::F32> tau, ffi::ResultBuffer<ffi::F32> q, ffi::ResultBuffer<ffi::F32> dq) {
  const mpxb_model* mdl = reinterpret_cast<const mpxb_model*>(static_cast<intptr_t>(model));
  if (mdl == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "mpxb_shift_extract: null model handle");
  const int E = static_cast<int>(x0.dimensions()[0]);
  const int rc = mpxb_shift_extract(mdl, stream, E, shift, X.typed_data(), U.typed_data(), V.typed_data(), x0.typed_data(),
                                    u_ref.typed_data(), X0->typed_data(), U0->typed_data(), V0->typed_data(),
                                    tau->typed_data(), q->typed_data(), dq->typed_data());
  return rc == MPXB_OK ? ffi::Error::Success() : Failure("mpxb_shift_extract", rc);
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(mpxb_mpc_step_ffi, MpcStep,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Attr<int32_t>("lq_mode")
                                  .Attr<int32_t>("max_iter")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // reference [E, N+1, nref]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // parameter [E, N+1, npar]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // W
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x0 [E, n]
                                  .Arg<ffi::Buffer<ffi::F32>>()   // X_in
                                  .Arg<ffi::Buffer<ffi::F32>>()   // U_in
                                  .Arg<ffi::Buffer<ffi::F32>>()   // V_in
                                  .Ret<ffi::Buffer<ffi::F32>>()   // X
                                  .Ret<ffi::Buffer<ffi::F32>>()   // U
                                  .Ret<ffi::Buffer<ffi::F32>>()   // V
                                  .Ret<ffi::Buffer<ffi::S32>>()   // alpha_index [E]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // iters [E]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // status [E]
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace [mpxb_workspace_bytes]

XLA_FFI_DEFINE_HANDLER_SYMBOL(mpxb_shift_extract_ffi, ShiftExtract,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model")
                                  .Attr<int32_t>("shift")
                                  .Arg<ffi::Buffer<ffi::F32>>()   // X
                                  .Arg<ffi::Buffer<ffi::F32>>()   // U
                                  .Arg<ffi::Buffer<ffi::F32>>()   // V
                                  .Arg<ffi::Buffer<ffi::F32>>()   // x0
                                  .Arg<ffi::Buffer<ffi::F32>>()   // u_ref
                                  .Ret<ffi::Buffer<ffi::F32>>()   // X0 (aliased in / out)
                                  .Ret<ffi::Buffer<ffi::F32>>()   // U0
                                  .Ret<ffi::Buffer<ffi::F32>>()   // V0
                                  .Ret<ffi::Buffer<ffi::F32>>()   // tau [E, nj]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // q
                                  .Ret<ffi::Buffer<ffi::F32>>()); // dq
