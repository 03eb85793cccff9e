// xla_ffi_shim.cc -- XLA typed-FFI handlers over the C ABI of include/ofdft_b200.h (ABI version 3).
//
// One handler per entry point of the hot path; a handler only unpacks buffers / attributes and forwards the stream --
// all logic lives under the C ABI.  XLA owns every buffer, including scratch and checkpoints (declared as extra results,
// sized on the Python side by the *_bytes helpers of the C ABI).  Python side: jax_ffi/bindings.py
// (jax.ffi.register_ffi_target + jax.custom_vjp).
//
// Build where JAX is installed (`python -c "import jax; print(jax.ffi.include_dir())"`):
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax;print(jax.ffi.include_dir())") -I/usr/local/cuda/include \
//       -I../../include xla_ffi_shim.cc -L.. -lofdft_b200 -Wl,-rpath,'$ORIGIN/..' -o ../libofdft_b200_xla.so
// JAX is NOT installed in the build image of this repository: there the file is compiled (syntax + signature check of
// every handler against its binding and against ofdft_b200.h) with the stand-in header tests/xla_ffi_stub/ by
// tests/test_capi.py::test_xla_ffi_shim_compiles_against_the_c_abi; it has not been executed under XLA.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "ofdft_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

static ffi::Error as_error(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + ofdft_b200_strerror(rc));
}
// a size-1 dummy operand / result stands for "not requested" (XLA has no optional buffers)
template <class B> static auto* opt(B& b) { return b.element_count() > 1 ? b.typed_data() : nullptr; }

// ---- 3-D equivariant GNN flow ------------------------------------------------------------------------------------------
// y1, ckpt, act <- gnn_fwd(theta, nuclei, onehot, y0)          replaces jax_ode.py:9-49 / 51-110 on Gen_EqvFlow
static ffi::Error GnnFwdImpl(cudaStream_t stream, F32 theta, F32 nuclei, F32 onehot, F32 y0, RF32 y1, RF32 ckpt, RF32 act,
                             RU8 scratch, int64_t n_a, int32_t jets_a, int32_t jets_b, int32_t n_layers, int32_t n_steps,
                             float t0, float t1, float y0sign, int32_t path) {
  const auto yd = y0.dimensions();
  const auto od = onehot.dimensions();
  return as_error(ofdft_cnf_gnn_fwd(stream, theta.typed_data(), nuclei.typed_data(), onehot.typed_data(), y0.typed_data(),
                                    y1->typed_data(), opt(*ckpt), opt(*act), scratch->typed_data(), scratch->element_count(),
                                    yd[0], n_a, jets_a, jets_b, (int32_t)yd[1], (int32_t)yd[1], (int32_t)od[0], (int32_t)od[1],
                                    n_layers, n_steps, t0, t1, y0sign, path),
                  "ofdft_cnf_gnn_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftGnnFwd, GnnFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("n_a").Attr<int32_t>("jets_a").Attr<int32_t>("jets_b")
                                  .Attr<int32_t>("n_layers").Attr<int32_t>("n_steps").Attr<float>("t0").Attr<float>("t1")
                                  .Attr<float>("y0sign").Attr<int32_t>("path"));

// theta_bar, ybar0 <- gnn_bwd(theta, nuclei, onehot, ckpt, act, ybar1)     the custom_vjp backward (discrete adjoint)
static ffi::Error GnnBwdImpl(cudaStream_t stream, F32 theta, F32 nuclei, F32 onehot, F32 ckpt, F32 act, F32 ybar1,
                             RF32 theta_bar, RF32 ybar0, RU8 scratch, int64_t n_a, int32_t jets_a, int32_t jets_b,
                             int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign, int32_t path) {
  const auto yd = ybar1.dimensions();
  const auto od = onehot.dimensions();
  return as_error(ofdft_cnf_gnn_bwd(stream, theta.typed_data(), nuclei.typed_data(), onehot.typed_data(), ckpt.typed_data(),
                                    opt(act), ybar1.typed_data(), theta_bar->typed_data(), opt(*ybar0), scratch->typed_data(),
                                    scratch->element_count(), yd[0], n_a, jets_a, jets_b, (int32_t)yd[1], (int32_t)od[0],
                                    (int32_t)od[1], n_layers, n_steps, t0, t1, y0sign, path),
                  "ofdft_cnf_gnn_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftGnnBwd, GnnBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int64_t>("n_a").Attr<int32_t>("jets_a").Attr<int32_t>("jets_b")
                                  .Attr<int32_t>("n_layers").Attr<int32_t>("n_steps").Attr<float>("t0").Attr<float>("t1")
                                  .Attr<float>("y0sign").Attr<int32_t>("path"));

// dy <- f.apply(params, t, states)  (equiv_flows.py:124-127) or the score-augmented right-hand side (jax_ode.py:80-94)
static ffi::Error GnnRhsImpl(cudaStream_t stream, F32 theta, F32 nuclei, F32 onehot, F32 y, RF32 dy, RU8 scratch,
                             int32_t jets, int32_t n_layers, float t, float y0sign, int32_t path) {
  const auto yd = y.dimensions();
  const auto od = onehot.dimensions();
  return as_error(ofdft_cnf_gnn_rhs(stream, theta.typed_data(), nuclei.typed_data(), onehot.typed_data(), y.typed_data(),
                                    dy->typed_data(), scratch->typed_data(), scratch->element_count(), yd[0], jets,
                                    (int32_t)yd[1], (int32_t)od[0], (int32_t)od[1], n_layers, t, y0sign, path),
                  "ofdft_cnf_gnn_rhs");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftGnnRhs, GnnRhsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("jets").Attr<int32_t>("n_layers").Attr<float>("t").Attr<float>("y0sign")
                                  .Attr<int32_t>("path"));

// ---- 1-D tanh-MLP flow (Gen_CNFSimpleMLP, cn_flows.py:166-182) ----------------------------------------------------------
static ffi::Error Mlp1dFwdImpl(cudaStream_t stream, F32 theta, F32 y0, RF32 y1, RF32 ckpt, RU8 scratch, int32_t jets,
                               int32_t H, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign, int32_t path) {
  const auto yd = y0.dimensions();
  return as_error(ofdft_cnf_mlp1d_fwd(stream, theta.typed_data(), y0.typed_data(), y1->typed_data(), opt(*ckpt),
                                      scratch->typed_data(), scratch->element_count(), yd[0], jets, (int32_t)yd[1], H, n_layers,
                                      n_steps, t0, t1, y0sign, path),
                  "ofdft_cnf_mlp1d_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftMlp1dFwd, Mlp1dFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("jets").Attr<int32_t>("H").Attr<int32_t>("n_layers").Attr<int32_t>("n_steps")
                                  .Attr<float>("t0").Attr<float>("t1").Attr<float>("y0sign").Attr<int32_t>("path"));

static ffi::Error Mlp1dBwdImpl(cudaStream_t stream, F32 theta, F32 ckpt, F32 ybar1, RF32 theta_bar, RF32 ybar0, RU8 scratch,
                               int32_t jets, int32_t H, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign,
                               int32_t path) {
  const auto yd = ybar1.dimensions();
  return as_error(ofdft_cnf_mlp1d_bwd(stream, theta.typed_data(), ckpt.typed_data(), ybar1.typed_data(), theta_bar->typed_data(),
                                      opt(*ybar0), scratch->typed_data(), scratch->element_count(), yd[0], jets, (int32_t)yd[1],
                                      H, n_layers, n_steps, t0, t1, y0sign, path),
                  "ofdft_cnf_mlp1d_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftMlp1dBwd, Mlp1dBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("jets").Attr<int32_t>("H").Attr<int32_t>("n_layers").Attr<int32_t>("n_steps")
                                  .Attr<float>("t0").Attr<float>("t1").Attr<float>("y0sign").Attr<int32_t>("path"));

static ffi::Error Mlp1dRhsImpl(cudaStream_t stream, F32 theta, F32 y, RF32 dy, RU8 scratch, int32_t jets, int32_t H,
                               int32_t n_layers, float t, float y0sign) {
  const auto yd = y.dimensions();
  return as_error(ofdft_cnf_mlp1d_rhs(stream, theta.typed_data(), y.typed_data(), dy->typed_data(), scratch->typed_data(),
                                      scratch->element_count(), yd[0], jets, (int32_t)yd[1], H, n_layers, t, y0sign),
                  "ofdft_cnf_mlp1d_rhs");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftMlp1dRhs, Mlp1dRhsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("jets").Attr<int32_t>("H").Attr<int32_t>("n_layers").Attr<float>("t")
                                  .Attr<float>("y0sign"));

// ---- functionals -----------------------------------------------------------------------------------------------------------
// sums, ybar1 <- functionals(y1)      replaces the functional calls + jnp.mean of `loss` (H20_...py:158-173, LiH.py:118-138)
static ffi::Error FunctionalsImpl(cudaStream_t stream, F32 y1, RF64 sums, RF32 ybar1, RU8 scratch, int32_t d, int32_t kin,
                                  int32_t hart, int32_t nuc, int32_t x, int32_t c, float Ne, float R, float Z_alpha,
                                  float Z_beta, ffi::Span<const float> coords, ffi::Span<const float> z) {
  const auto yd = y1.dimensions();
  ofdft_energy_spec spec{d, kin, hart, nuc, x, c, (int32_t)z.size(), Ne, R, Z_alpha, Z_beta};
  return as_error(ofdft_functionals_fwd_bwd(stream, &spec, coords.begin(), z.begin(), y1.typed_data(), (int32_t)yd[1],
                                            yd[0] / 2, sums->typed_data(), opt(*ybar1), scratch->typed_data(),
                                            scratch->element_count()),
                  "ofdft_functionals_fwd_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftFunctionals, FunctionalsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("d").Attr<int32_t>("kin").Attr<int32_t>("hart").Attr<int32_t>("nuc")
                                  .Attr<int32_t>("x").Attr<int32_t>("c").Attr<float>("Ne").Attr<float>("R").Attr<float>("Z_alpha")
                                  .Attr<float>("Z_beta").Attr<ffi::Span<const float>>("coords")
                                  .Attr<ffi::Span<const float>>("z"));

// esum, xbar, xpbar <- all-pairs Hartree of one block (x rows) x (xp rows): mean over `norm_pairs` pairs (0: n*m)
static ffi::Error HartreeAllpairsImpl(cudaStream_t stream, F32 x, F32 xp, RF64 esum, RF32 xbar, RF32 xpbar, RU8 scratch,
                                      int32_t kind, int32_t d, float Ne, double norm_pairs) {
  const auto xd = x.dimensions();
  const auto pd = xp.dimensions();
  return as_error(ofdft_hartree_allpairs(stream, kind, d, Ne, x.typed_data(), (int32_t)xd[1], xp.typed_data(), (int32_t)pd[1],
                                         xd[0], pd[0], norm_pairs, esum->typed_data(), opt(*xbar), d, opt(*xpbar), d, 0,
                                         scratch->typed_data(), scratch->element_count()),
                  "ofdft_hartree_allpairs");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(OfdftHartreeAllpairs, HartreeAllpairsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<F32>().Arg<F32>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<F32>().Ret<F32>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("kind").Attr<int32_t>("d").Attr<float>("Ne").Attr<double>("norm_pairs"));
