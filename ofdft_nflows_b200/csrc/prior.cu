// prior.cu -- base-distribution kernels on the device (the callers either side of the
// hot path): the promolecular prior's log-density / score (ofdft_normflows/promolecular_distrax.py:52-62,86-89), the
// batch generators (ofdft_normflows/utils.py:76-143) and the last step of rho_rev (H20_mol_ofdft_min_equiv_flows.py:132-137).
// Elementwise / HBM-bound; Philox counter-based RNG from curand's device API (reproducible per (seed, row)).
#include <curand_kernel.h>

#include <cmath>

#include "common.cuh"

namespace ofdft {

constexpr int kPriorMaxNuc = 128;   // C27H46O (utils.py:327) has 74
struct PriorArgs {
  float coords[kPriorMaxNuc * 3];
  float logw[kPriorMaxNuc];     // log(Z_i / sum Z)
  float cdf[kPriorMaxNuc];
  int n_nuclei;
};

// log p(x) and grad log p(x) of the mixture of unit-variance Gaussians on the nuclei
__device__ __forceinline__ void promolecular(const PriorArgs& P, const float (&x)[3], float& logp, float (&score)[3]) {
  float mx = -INFINITY;
  for (int i = 0; i < P.n_nuclei; ++i) {
    const float d0 = x[0] - P.coords[3 * i], d1 = x[1] - P.coords[3 * i + 1], d2 = x[2] - P.coords[3 * i + 2];
    mx = fmaxf(mx, P.logw[i] - 0.5f * (d0 * d0 + d1 * d1 + d2 * d2));
  }
  float tot = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int i = 0; i < P.n_nuclei; ++i) {
    const float d0 = x[0] - P.coords[3 * i], d1 = x[1] - P.coords[3 * i + 1], d2 = x[2] - P.coords[3 * i + 2];
    const float e = expf(P.logw[i] - 0.5f * (d0 * d0 + d1 * d1 + d2 * d2) - mx);
    tot += e; s0 += e * d0; s1 += e * d1; s2 += e * d2;
  }
  logp = logf(tot) + mx - 2.756815599614018f;        // - (3/2) log(2 pi)
  const float inv = 1.0f / tot;
  score[0] = -s0 * inv; score[1] = -s1 * inv; score[2] = -s2 * inv;
}

__global__ void promolecular_kernel(const PriorArgs P, const float* __restrict__ x, int x_stride, long long n,
                                    const float* __restrict__ sub, int sub_stride, int exp_out,
                                    float* __restrict__ out, float* __restrict__ score) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv[3] = {x[i * x_stride], x[i * x_stride + 1], x[i * x_stride + 2]};
  float lp, sc[3];
  promolecular(P, xv, lp, sc);
  if (sub != nullptr) lp -= sub[i * sub_stride];
  if (out != nullptr) out[i] = exp_out ? expf(lp) : lp;
  if (score != nullptr) { score[3 * i] = sc[0]; score[3 * i + 1] = sc[1]; score[3 * i + 2] = sc[2]; }
}

__global__ void batch_generator_3d_kernel(const PriorArgs P, unsigned long long seed, unsigned long long offset,
                                          long long rows, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  curandStatePhilox4_32_10_t st;
  // Philox skip-ahead counts single 32-bit outputs: a row consumes 5 per call (1 uniform + normal4), so consecutive calls
  // are spaced 8 outputs apart -- every batch draws from a disjoint stretch of the row's subsequence
  curand_init(seed, (unsigned long long)i, offset * 8ull, &st);
  const float u = curand_uniform(&st);
  int c = 0;
  while (c + 1 < P.n_nuclei && u > P.cdf[c]) ++c;
  const float4 g = curand_normal4(&st);
  const float xv[3] = {P.coords[3 * c] + g.x, P.coords[3 * c + 1] + g.y, P.coords[3 * c + 2] + g.z};
  float lp, sc[3];
  promolecular(P, xv, lp, sc);
  float* o = out + i * 7;
  o[0] = xv[0]; o[1] = xv[1]; o[2] = xv[2]; o[3] = lp; o[4] = sc[0]; o[5] = sc[1]; o[6] = sc[2];
}

__global__ void batch_generator_1d_kernel(unsigned long long seed, unsigned long long offset, long long rows,
                                          float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)i, offset * 8ull, &st);   // 2 outputs per row and call; calls spaced 8 apart
  const float x = curand_normal(&st);
  out[3 * i] = x; out[3 * i + 1] = -0.5f * x * x - 0.9189385332046727f; out[3 * i + 2] = -x;
}

static int fill_prior(PriorArgs& P, const float* coords, const float* z, int n) {
  if (!coords || !z || n < 1) return OFDFT_EINVAL;
  if (n > kPriorMaxNuc) return OFDFT_EUNSUPPORTED;
  double tot = 0.0;
  for (int i = 0; i < n; ++i) tot += fabs((double)z[i]);
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    P.coords[3 * i] = coords[3 * i]; P.coords[3 * i + 1] = coords[3 * i + 1]; P.coords[3 * i + 2] = coords[3 * i + 2];
    const double w = (double)z[i] / tot;
    P.logw[i] = (float)log(w);
    acc += w;
    P.cdf[i] = (float)acc;
  }
  P.n_nuclei = n;
  return 0;
}

}  // namespace ofdft

using namespace ofdft;

extern "C" int ofdft_promolecular_logp_score(void* stream, const float* coords, const float* z, int32_t n_nuclei,
                                             const float* x, int32_t x_stride, int64_t n, const float* sub,
                                             int32_t sub_stride, int32_t exp_out, float* out, float* score) {
  if (!x || n <= 0 || x_stride < 3 || (!out && !score)) return OFDFT_EINVAL;
  PriorArgs P;
  int rc = fill_prior(P, coords, z, n_nuclei);
  if (rc) return rc;
  promolecular_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, x, x_stride, n, sub, sub_stride,
                                                                                      exp_out, out, score);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_batch_generator_3d(void* stream, uint64_t seed, uint64_t offset, const float* coords, const float* z,
                                        int32_t n_nuclei, int64_t bs, float* out) {
  if (!out || bs <= 0) return OFDFT_EINVAL;
  PriorArgs P;
  int rc = fill_prior(P, coords, z, n_nuclei);
  if (rc) return rc;
  const long long rows = 2 * bs;
  batch_generator_3d_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, seed, offset, rows, out);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_batch_generator_1d(void* stream, uint64_t seed, uint64_t offset, int64_t bs, float* out) {
  if (!out || bs <= 0) return OFDFT_EINVAL;
  const long long rows = 2 * bs;
  batch_generator_1d_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, offset, rows, out);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// optimiser step: optax.chain(clip_by_global_norm(max_norm), rmsprop(lr))
// as the drivers build it (H20_mol_ofdft_min_equiv_flows.py:111-116).  [third-party] optax semantics restated from
// its published algorithm: scale = min(1, max_norm / ||g||_2);  nu <- decay nu + (1-decay) g^2;
// theta <- theta - lr * g * rsqrt(nu + eps)   (eps inside the square root, initial nu = 0).
// One CTA: theta has 4.6k-527k entries, the step is latency-bound and runs after the gradient all-reduce.
// ------------------------------------------------------------------------------------------------
namespace ofdft {
__global__ void __launch_bounds__(1024) rmsprop_clip_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                                            float* __restrict__ nu, long long n, float lr, float decay,
                                                            float eps, float max_norm) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)grad[i] * (double)grad[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float gnorm = (float)sqrt(red[0]);
  const float scale = (max_norm > 0.f && gnorm > max_norm) ? max_norm / gnorm : 1.0f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float g = grad[i] * scale;
    const float v = decay * nu[i] + (1.0f - decay) * g * g;
    nu[i] = v;
    theta[i] -= lr * g * rsqrtf(v + eps);
  }
}

// optax.chain(clip_by_global_norm(max_norm), adam(lr)) as OFDFT_NF.py:104-108 builds it.  [third-party] optax.adam restated:
// mu <- b1 mu + (1-b1) g;  nu <- b2 nu + (1-b2) g^2;  theta <- theta - lr * (mu / (1-b1^t)) / (sqrt(nu / (1-b2^t)) + eps)
// (eps outside the square root, eps_root = 0, t = 1-based update count; defaults b1 = 0.9, b2 = 0.999, eps = 1e-8).
__global__ void __launch_bounds__(1024) adam_clip_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                                         float* __restrict__ mu, float* __restrict__ nu, long long n, float lr,
                                                         float b1, float b2, float eps, float max_norm, float c1, float c2) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)grad[i] * (double)grad[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) red[0] = v;
  }
  __syncthreads();
  const float gnorm = (float)sqrt(red[0]);
  const float scale = (max_norm > 0.f && gnorm > max_norm) ? max_norm / gnorm : 1.0f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float g = grad[i] * scale;
    const float m = fmaf(b1, mu[i], (1.0f - b1) * g);
    const float v = fmaf(b2, nu[i], (1.0f - b2) * g * g);
    mu[i] = m; nu[i] = v;
    theta[i] -= lr * (m * c1) / (sqrtf(v * c2) + eps);       // c1 = 1/(1-b1^t), c2 = 1/(1-b2^t)
  }
}
__global__ void ema_update_kernel(const double* __restrict__ values, double* __restrict__ state, double* __restrict__ out, int n,
                                  double decay, double debias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = decay * state[i] + (1.0 - decay) * values[i];
  state[i] = v;
  out[i] = v / debias;
}
}  // namespace ofdft

extern "C" int ofdft_ema_update(void* stream, const double* values, double* state, double* out, int32_t n, int64_t step,
                                double decay) {
  if (!values || !state || !out || n <= 0 || step < 1 || !(decay >= 0.0 && decay < 1.0)) return OFDFT_EINVAL;
  const double debias = 1.0 - pow(decay, (double)step);
  ofdft::ema_update_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(values, state, out, n, decay, debias);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_rmsprop_clip_step(void* stream, float* theta, const float* grad, float* nu, int64_t n, float lr,
                                       float decay, float eps, float max_norm) {
  if (!theta || !grad || !nu || n <= 0) return OFDFT_EINVAL;
  ofdft::rmsprop_clip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(theta, grad, nu, n, lr, decay, eps, max_norm);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_adam_clip_step(void* stream, float* theta, const float* grad, float* mu, float* nu, int64_t n, int64_t step,
                                    float lr, float b1, float b2, float eps, float max_norm) {
  if (!theta || !grad || !mu || !nu || n <= 0 || step < 1) return OFDFT_EINVAL;
  const float c1 = (float)(1.0 / (1.0 - pow((double)b1, (double)step)));
  const float c2 = (float)(1.0 / (1.0 - pow((double)b2, (double)step)));
  ofdft::adam_clip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(theta, grad, mu, nu, n, lr, b1, b2, eps, max_norm, c1, c2);
  return (int)cudaGetLastError();
}
