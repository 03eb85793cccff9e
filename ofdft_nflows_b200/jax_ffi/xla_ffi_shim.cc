// xla_ffi_shim.cc -- XLA typed-FFI handlers over the C ABI of include/ofdft_b200.h.
//
// NOT BUILT IN THIS IMAGE: JAX is not installed here (no `xla/ffi/api/ffi.h` on disk), so this file is
// compiled only where `python -c "import jax; print(jax.ffi.include_dir())"` works:
//
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax;print(jax.ffi.include_dir())") -I../../include \
//       xla_ffi_shim.cc -L.. -lofdft_b200 -Wl,-rpath,'$ORIGIN/..' -o ../libofdft_b200_xla.so
//
// All logic lives under the C ABI; a handler only unpacks buffers/attributes and forwards the stream.
// Python side: ofdft_nflows_b200/jax_ffi/bindings.py (jax.ffi.register_ffi_target + jax.custom_vjp).
#include <cstdint>

#include "ofdft_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error as_error(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + ofdft_b200_strerror(rc));
}

// y1, ckpt, act <- gnn_fwd(theta, nuclei, onehot, y0)           replaces jax_ode.py:9-49 / 51-110 on Gen_EqvFlow
static ffi::Error GnnFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> theta, ffi::Buffer<ffi::F32> nuclei,
                             ffi::Buffer<ffi::F32> onehot, ffi::Buffer<ffi::F32> y0, ffi::ResultBuffer<ffi::F32> y1,
                             ffi::ResultBuffer<ffi::F32> ckpt, ffi::ResultBuffer<ffi::F32> act,
                             ffi::ResultBuffer<ffi::U8> scratch, int64_t n_a, int32_t jets_a, int32_t jets_b,
                             int32_t n_steps, float t0, float t1, float y0sign) {
  const auto yd = y0.dimensions();
  const auto od = onehot.dimensions();
  float* act_ptr = act->element_count() > 1 ? act->typed_data() : nullptr;   // size-1 dummy = recompute mode
  return as_error(ofdft_cnf_gnn_fwd(stream, theta.typed_data(), nuclei.typed_data(), onehot.typed_data(), y0.typed_data(),
                                    y1->typed_data(), ckpt->typed_data(), act_ptr, scratch->typed_data(),
                                    scratch->element_count(), yd[0], n_a, jets_a, jets_b, (int32_t)yd[1], (int32_t)yd[1],
                                    (int32_t)od[0], (int32_t)od[1], n_steps, t0, t1, y0sign),
                  "ofdft_cnf_gnn_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftGnnFwd, GnnFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("n_a").Attr<int32_t>("jets_a").Attr<int32_t>("jets_b")
                                  .Attr<int32_t>("n_steps").Attr<float>("t0").Attr<float>("t1").Attr<float>("y0sign"));

// theta_bar, ybar0 <- gnn_bwd(theta, nuclei, onehot, ckpt, act, ybar1)      the custom_vjp backward
static ffi::Error GnnBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> theta, ffi::Buffer<ffi::F32> nuclei,
                             ffi::Buffer<ffi::F32> onehot, ffi::Buffer<ffi::F32> ckpt, ffi::Buffer<ffi::F32> act,
                             ffi::Buffer<ffi::F32> ybar1, ffi::ResultBuffer<ffi::F32> theta_bar,
                             ffi::ResultBuffer<ffi::F32> ybar0, ffi::ResultBuffer<ffi::U8> scratch, int64_t n_a,
                             int32_t jets_a, int32_t jets_b, int32_t n_steps, float t0, float t1, float y0sign) {
  const auto yd = ybar1.dimensions();
  const auto od = onehot.dimensions();
  const float* act_ptr = act.element_count() > 1 ? act.typed_data() : nullptr;
  return as_error(ofdft_cnf_gnn_bwd(stream, theta.typed_data(), nuclei.typed_data(), onehot.typed_data(), ckpt.typed_data(),
                                    act_ptr, ybar1.typed_data(), theta_bar->typed_data(), ybar0->typed_data(),
                                    scratch->typed_data(), scratch->element_count(), yd[0], n_a, jets_a, jets_b,
                                    (int32_t)yd[1], (int32_t)od[0], (int32_t)od[1], n_steps, t0, t1, y0sign),
                  "ofdft_cnf_gnn_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftGnnBwd, GnnBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("n_a").Attr<int32_t>("jets_a").Attr<int32_t>("jets_b")
                                  .Attr<int32_t>("n_steps").Attr<float>("t0").Attr<float>("t1").Attr<float>("y0sign"));

// sums, ybar1 <- functionals(y1)      replaces the functional calls + jnp.mean of `loss` (H20_...py:158-173)
static ffi::Error FunctionalsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> y1, ffi::ResultBuffer<ffi::F64> sums,
                                  ffi::ResultBuffer<ffi::F32> ybar1, ffi::ResultBuffer<ffi::U8> scratch, int32_t d,
                                  int32_t kin, int32_t hart, int32_t nuc, int32_t x, int32_t c, float Ne, float R, float Z_alpha,
                                  float Z_beta, ffi::Span<const float> coords, ffi::Span<const float> z) {
  const auto yd = y1.dimensions();
  ofdft_energy_spec spec{d, kin, hart, nuc, x, c, (int32_t)z.size(), Ne, R, Z_alpha, Z_beta};
  return as_error(ofdft_functionals_fwd_bwd(stream, &spec, coords.begin(), z.begin(), y1.typed_data(), (int32_t)yd[1],
                                            yd[0] / 2, sums->typed_data(), ybar1->typed_data(), scratch->typed_data(),
                                            scratch->element_count()),
                  "ofdft_functionals_fwd_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftFunctionals, FunctionalsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("d").Attr<int32_t>("kin").Attr<int32_t>("hart").Attr<int32_t>("nuc")
                                  .Attr<int32_t>("x").Attr<int32_t>("c").Attr<float>("Ne").Attr<float>("R").Attr<float>("Z_alpha")
                                  .Attr<float>("Z_beta").Attr<ffi::Span<const float>>("coords")
                                  .Attr<ffi::Span<const float>>("z"));
