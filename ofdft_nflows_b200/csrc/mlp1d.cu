// mlp1d.cu -- fused fixed-step (RK4) integrator and discrete adjoint for the 1-D tanh-MLP flow
// g(x,t) = MLP([t', x])  (reference: ofdft_normflows/cn_flows.py:117-182 through jax_ode.py:9-110; LiH.py:64-65 uses
// features (512,512,512)).  sm_100a, FP32 FFMA pipe (round 1; the tcgen05 3xTF32 GEMM tiles are the next step).
//
// One CTA = one tile of 8 trajectories, H/2 threads: thread = (trajectory t, column group g of 16 hidden units).
// Per RHS evaluation the 2-jet (value, d/dx, d2/dx2) of every hidden layer is a [8 traj x 3 jets] x [H x H] GEMM:
// the activation jets live in shared memory, the weights are streamed through a double-buffered shared-memory ring
// with cp.async (16 rows per chunk; all CTAs sweep the same 1-2 MB of weights, which stays L2 resident).
//
// Backward = discrete adjoint of RK4 on the checkpointed stage states.  Per stage the sweep kernel recomputes the
// forward jets, runs the transposed GEMMs (W_l^T streamed from a pre-transposed copy) and writes, for every hidden
// layer, the jets (a, p', p'') and the pre-activation cotangent jets to HBM; the weight gradients
// dW_l = sum_rows A_{l-1}^T Pbar_l (rows = stage x trajectory x jet, ~2e5 for bs=512) are then ONE split-K GEMM per
// layer over those buffers -- a per-CTA accumulation would need 1 MB of accumulators per layer, and per-stage
// atomics would be ~50x slower than the FFMA work.
#include <cstdlib>

#include <mutex>

#include "common.cuh"
#include "tc.cuh"
#include "mlp1d_common.cuh"

namespace ofdft {
namespace mlp {

struct Args {
  const float* theta;
  const float* Wc;          // aligned copies of W_1..W_{L-1} (layer index i>=1), each H*H
  const float* WTc;         // transposes
  const float* y_in; float* y_out; float* ckpt;
  const float* ybar1; float* ybar0;
  float* acts; float* pbar; float* gsmall;   // backward scratch
  long long B, row_begin, rows;              // this launch handles rows [row_begin, row_begin+rows)
  int jets, stride, H, L, n_steps; float t0, t1, y0;
  int rhs_only; float t_rhs;                 // one evaluation of the vector field at t_rhs: y_out <- d/dt y
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

struct Smem {
  float *bufA, *bufB, *ws, *wx, *wt, *b1, *bl, *wl, *xs, *hpart, *hbar, *xpart;
  __device__ Smem(float* base, int H, int L) {
    bufA = base; base += 3 * H * kTT;
    bufB = base; base += 3 * H * kTT;
    ws = base; base += 2 * kKC * H;
    wx = base; base += H;
    wt = base; base += H;
    b1 = base; base += H;
    bl = base; base += (kMaxL - 1) * H;
    wl = base; base += H;
    xs = base; base += kTT;
    hpart = base; base += (H / 64) * 3 * kTT;
    hbar = base; base += 3 * kTT;
    xpart = base; base += (H / 64) * kTT;
  }
};
static size_t smem_bytes(int H) {
  return sizeof(float) * (size_t)(6 * H * kTT + 2 * kKC * H + 4 * H + (kMaxL - 1) * H + kTT + (H / 64) * 3 * kTT + 3 * kTT +
                                  (H / 64) * kTT);
}

__device__ __forceinline__ void issue_chunk(float* dst, const float* __restrict__ src, int H, int tid, int nthreads) {
  const int n4 = kKC * H / 4;
  for (int i = tid; i < n4; i += nthreads) cp_async16(dst + 4 * i, src + 4 * i);
  cp_async_commit();
}

// acc[jet][c] = sum_k actIn[jet][k][t] * W[k][col(c)],  col(c) = (H/4)*(c>>2) + 4*g + (c&3)
template <int NJ>
__device__ __forceinline__ void layer_gemm(float (&acc)[NJ][16], const float* __restrict__ actIn,
                                           const float* __restrict__ Wg, float* ws, int H, int t, int g, int tid, int nthreads) {
#pragma unroll
  for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[jt][c] = 0.f;
  const int nchunks = H / kKC;
  const int q = H / 4;
  __syncthreads();                       // writers of actIn and readers of the ring are done
  issue_chunk(ws, Wg, H, tid, nthreads);
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait_all();
    __syncthreads();
    if (c + 1 < nchunks) issue_chunk(ws + ((c + 1) & 1) * kKC * H, Wg + (size_t)(c + 1) * kKC * H, H, tid, nthreads);
    const float* wb = ws + (c & 1) * kKC * H + 4 * g;
    const float* ab = actIn + (size_t)(c * kKC) * kTT + t;
#pragma unroll 4
    for (int kk = 0; kk < kKC; ++kk) {
      float av[3];
      av[0] = ab[kk * kTT];
      if (NJ >= 2) av[1] = ab[(H + kk) * kTT];
      if (NJ == 3) av[2] = ab[(2 * H + kk) * kTT];
      const float4 w0 = *reinterpret_cast<const float4*>(wb + kk * H);
      const float4 w1 = *reinterpret_cast<const float4*>(wb + kk * H + q);
      const float4 w2 = *reinterpret_cast<const float4*>(wb + kk * H + 2 * q);
      const float4 w3 = *reinterpret_cast<const float4*>(wb + kk * H + 3 * q);
      const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) acc[jt][cc] = fmaf(av[jt], wv[cc], acc[jt][cc]);
    }
  }
}

__device__ __forceinline__ int col_of(int c, int g, int H) { return (H / 4) * (c >> 2) + 4 * g + (c & 3); }

// first layer: p = w_x x + (t' w_t + b1), p' = w_x, p'' = 0  ->  jets of a1 into act[jet][col][t]
template <int NJ>
__device__ __forceinline__ void first_layer(float* act, const Smem& sm, float x, float tprime, int H, int t, int g) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int col = col_of(c, g, H);
    const float w = sm.wx[col];
    const float a = tanh_f(fmaf(w, x, fmaf(tprime, sm.wt[col], sm.b1[col])));
    const float sg = fmaf(-a, a, 1.0f);
    const float a1 = sg * w;
    act[(size_t)col * kTT + t] = a;
    if (NJ >= 2) act[(size_t)(H + col) * kTT + t] = a1;
    if (NJ == 3) act[(size_t)(2 * H + col) * kTT + t] = -2.0f * a * a1 * w;
  }
}

// reduce per-thread partials over the column groups: in-warp over 4 groups, then across warps through smem
template <int N>
__device__ __forceinline__ void reduce_groups(float (&v)[N], float* part, int t, int g, int tid) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 8);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
  if ((tid & 31) < kTT) {
#pragma unroll
    for (int i = 0; i < N; ++i) part[((tid >> 5) * N + i) * kTT + t] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
template <int NJ>
__device__ void fwd_body(const Args& a, float* smem_base) {
  const int H = a.H, L = a.L;
  const int tid = threadIdx.x, nthreads = blockDim.x, t = tid & 7, g = tid >> 3, nwarps = nthreads >> 5;
  Smem sm(smem_base, H, L);
  const ThetaLayout TL(H, L);
  for (int i = tid; i < H; i += nthreads) {
    sm.wt[i] = a.theta[TL.w(0) + i];
    sm.wx[i] = a.theta[TL.w(0) + H + i];
    sm.b1[i] = a.theta[TL.b(0) + i];
    sm.wl[i] = a.theta[TL.w_last() + i];
    for (int l = 1; l < L; ++l) sm.bl[(l - 1) * H + i] = a.theta[TL.b(l) + i];
  }
  const float b_last = a.theta[0];
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  const long long n_tiles = (a.rows + kTT - 1) / kTT;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const bool owner = tid < kTT;
    const long long row = a.row_begin + tile * kTT + tid;
    const bool valid = owner && (tile * kTT + tid) < a.rows;
    const long long rowc = (tile * kTT + tid) < a.rows ? row : a.row_begin + a.rows - 1;
    float y[3] = {0, 0, 0}, accw[3] = {0, 0, 0}, kst[3] = {0, 0, 0}, xst = 0.f, sst = 0.f;
    if (owner) {
      const float* src = a.y_in + rowc * a.stride;
      y[0] = src[0];
      if (NJ >= 2) y[1] = src[1];
      if (NJ == 3) y[2] = src[2];
    }
    float acc[NJ][16];
    const int n_steps = a.rhs_only ? 1 : a.n_steps;
    for (int step = 0; step < n_steps; ++step) {
      for (int stage = 0; stage < (a.rhs_only ? 1 : 4); ++stage) {
        const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
        const float tprime = a.y0 * (a.rhs_only ? a.t_rhs : a.t0 + h * (float)step + cs);
        __syncthreads();
        if (owner) {
          xst = fmaf(cs, kst[0], y[0]);
          if (NJ == 3) sst = fmaf(cs, kst[2], y[2]);
          sm.xs[tid] = xst;
          if (a.ckpt != nullptr && valid) {
            float* ck = a.ckpt + ((long long)(step * 4 + stage) * 2) * a.B + row;
            ck[0] = xst;
            if (NJ == 3) ck[a.B] = sst;
          }
        }
        __syncthreads();
        first_layer<NJ>(sm.bufA, sm, sm.xs[t], tprime, H, t, g);
        float* in = sm.bufA; float* out = sm.bufB;
        float hp[3] = {0.f, 0.f, 0.f};
        for (int l = 1; l < L; ++l) {
          layer_gemm<NJ>(acc, in, a.Wc + (size_t)(l - 1) * H * H, sm.ws, H, t, g, tid, nthreads);
          const float* bl = sm.bl + (l - 1) * H;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col_of(c, g, H);
            const float av = tanh_f(acc[0][c] + bl[col]);
            const float sg = fmaf(-av, av, 1.0f);
            float a1 = 0.f, a2 = 0.f;
            if (NJ >= 2) a1 = sg * acc[1][c];
            if (NJ == 3) a2 = fmaf(sg, acc[2][c], -2.0f * av * a1 * acc[1][c]);
            if (l + 1 < L) {
              out[(size_t)col * kTT + t] = av;
              if (NJ >= 2) out[(size_t)(H + col) * kTT + t] = a1;
              if (NJ == 3) out[(size_t)(2 * H + col) * kTT + t] = a2;
            } else {
              const float w = sm.wl[col];
              hp[0] = fmaf(w, av, hp[0]); hp[1] = fmaf(w, a1, hp[1]); hp[2] = fmaf(w, a2, hp[2]);
            }
          }
          float* tmp = in; in = out; out = tmp;
        }
        reduce_groups<3>(hp, sm.hpart, t, g, tid);
        __syncthreads();
        if (owner) {
          float h0 = b_last, h1 = 0.f, h2 = 0.f;
          for (int w = 0; w < nwarps; ++w) {
            h0 += sm.hpart[(w * 3 + 0) * kTT + tid]; h1 += sm.hpart[(w * 3 + 1) * kTT + tid]; h2 += sm.hpart[(w * 3 + 2) * kTT + tid];
          }
          kst[0] = a.y0 * h0;
          if (NJ >= 2) kst[1] = -a.y0 * h1;
          if (NJ == 3) kst[2] = -a.y0 * (h1 * sst + h2);
          const float wgt = (stage == 0 || stage == 3) ? 1.0f : 2.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) accw[c] = fmaf(wgt, kst[c], accw[c]);
        }
      }
      if (owner && !a.rhs_only) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { y[c] = fmaf(h * (1.0f / 6.0f), accw[c], y[c]); accw[c] = 0.f; }
      }
    }
    if (valid) {
      float* dst = a.y_out + row * a.stride;
      const float* o = a.rhs_only ? kst : y;
      dst[0] = o[0];
      if (NJ >= 2) dst[1] = o[1];
      if (NJ == 3) dst[2] = o[2];
    }
  }
}

__global__ void __launch_bounds__(256, 1) mlp_fwd_kernel(const Args a) {
  extern __shared__ __align__(16) float smem_f[];
  if (a.jets == 3) fwd_body<3>(a, smem_f);
  else if (a.jets == 2) fwd_body<2>(a, smem_f);
  else fwd_body<1>(a, smem_f);
}

// aligned copies (and transposes) of the hidden kernels: theta offsets are odd, cp.async needs 16-byte sources
__global__ void mlp_prep_kernel(const float* __restrict__ theta, float* __restrict__ Wc, float* __restrict__ WTc, int H, int L,
                                float* __restrict__ WTlo = nullptr, float* __restrict__ Wclo = nullptr) {
  const ThetaLayout TL(H, L);
  const long long n = (long long)(L - 1) * H * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / ((long long)H * H)) + 1;
    const int r = (int)((i % ((long long)H * H)) / H), c = (int)(i % H);
    const float v = theta[TL.w(l) + r * H + c];
    Wc[i] = v;
    if (Wclo != nullptr) Wclo[i] = tc::lo_tf32(v);
    if (WTc != nullptr) WTc[(long long)(l - 1) * H * H + (long long)c * H + r] = v;
    if (WTlo != nullptr) WTlo[(long long)(l - 1) * H * H + (long long)c * H + r] = tc::lo_tf32(v);
  }
}

// ------------------------------------------------------------------------------------------------
// backward sweep kernel (jets = 3 only: the training path)
// ------------------------------------------------------------------------------------------------
// HBM block of one (stage, tile): acts[blk][layer 0..L-1][jet: a, p', p''][H][8], pbar[blk][layer 1..L-1][3][H][8]
__device__ __forceinline__ size_t act_blk(long long blk, int layer, int H, int L) { return ((size_t)blk * L + layer) * 3 * H * kTT; }
__device__ __forceinline__ size_t pb_blk(long long blk, int layer, int H, int L) { return ((size_t)blk * (L - 1) + (layer - 1)) * 3 * H * kTT; }

struct SmallGrad {          // per-thread partials for this thread's 16 columns (and trajectory t)
  float gb1[16], gwx[16], gwt[16], gwl[16], gbl[kMaxL - 1][16];
  float gblast;
};

__global__ void __launch_bounds__(256, 1) mlp_bwd_sweep_kernel(const Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const int H = a.H, L = a.L;
  const int tid = threadIdx.x, nthreads = blockDim.x, t = tid & 7, g = tid >> 3, nwarps = nthreads >> 5;
  Smem sm(smem_f, H, L);
  const ThetaLayout TL(H, L);
  for (int i = tid; i < H; i += nthreads) {
    sm.wt[i] = a.theta[TL.w(0) + i];
    sm.wx[i] = a.theta[TL.w(0) + H + i];
    sm.b1[i] = a.theta[TL.b(0) + i];
    sm.wl[i] = a.theta[TL.w_last() + i];
    for (int l = 1; l < L; ++l) sm.bl[(l - 1) * H + i] = a.theta[TL.b(l) + i];
  }
  const float b_last = a.theta[0];
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  const long long n_tiles = (a.rows + kTT - 1) / kTT;
  SmallGrad G;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    G.gb1[c] = G.gwx[c] = G.gwt[c] = G.gwl[c] = 0.f;
#pragma unroll
    for (int l = 0; l < kMaxL - 1; ++l) G.gbl[l][c] = 0.f;
  }
  G.gblast = 0.f;
  float acc[3][16];
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const bool owner = tid < kTT;
    const long long row = a.row_begin + tile * kTT + tid;
    const bool valid = owner && (tile * kTT + tid) < a.rows;
    const long long rowc = (tile * kTT + tid) < a.rows ? row : a.row_begin + a.rows - 1;
    float xb = 0.f, lb = 0.f, sb = 0.f;
    if (valid) { const float* src = a.ybar1 + row * a.stride; xb = src[0]; lb = src[1]; sb = src[2]; }
    float prevx = 0.f, prevs = 0.f;
    for (int step = a.n_steps - 1; step >= 0; --step) {
      float sumx = 0.f, sums = 0.f;
      for (int stage = 3; stage >= 0; --stage) {
        const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
        const float tprime = a.y0 * (a.t0 + h * (float)step + cs);
        const long long blk = (long long)(step * 4 + stage) * n_tiles + tile;
        float kx = 0.f, kl = 0.f, ks = 0.f, xst = 0.f, sst = 0.f;
        __syncthreads();
        if (owner) {
          const float ca = (stage == 0 || stage == 3) ? h * (1.0f / 6.0f) : h * (1.0f / 3.0f);
          const float cp = stage == 3 ? 0.f : (stage == 2 ? h : 0.5f * h);
          kx = fmaf(ca, xb, cp * prevx); ks = fmaf(ca, sb, cp * prevs); kl = ca * lb;
          const float* ck = a.ckpt + ((long long)(step * 4 + stage) * 2) * a.B + rowc;
          xst = ck[0]; sst = ck[a.B];
          sm.xs[tid] = xst;
          // dx = y0 h ; dlogp = -y0 h' ; ds = -y0 (h' s + h'')
          const float hb0 = a.y0 * kx, hb1 = -a.y0 * (kl + sst * ks), hb2 = -a.y0 * ks;
          sm.hbar[tid] = hb0; sm.hbar[kTT + tid] = hb1; sm.hbar[2 * kTT + tid] = hb2;
          G.gblast += hb0;
        }
        __syncthreads();
        const float x_t = sm.xs[t];
        const float hb[3] = {sm.hbar[t], sm.hbar[kTT + t], sm.hbar[2 * kTT + t]};
        // ---- forward recompute, storing (a, p', p'') of every hidden layer ----
        first_layer<3>(sm.bufA, sm, x_t, tprime, H, t, g);
        {
          float* dst = a.acts + act_blk(blk, 0, H, L);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col_of(c, g, H);
            dst[(size_t)col * kTT + t] = sm.bufA[(size_t)col * kTT + t];
            dst[(size_t)(H + col) * kTT + t] = sm.wx[col];
            dst[(size_t)(2 * H + col) * kTT + t] = 0.f;
          }
        }
        float* in = sm.bufA; float* out = sm.bufB;
        float hp[3] = {0.f, 0.f, 0.f};
        for (int l = 1; l < L; ++l) {
          layer_gemm<3>(acc, in, a.Wc + (size_t)(l - 1) * H * H, sm.ws, H, t, g, tid, nthreads);
          const float* bl = sm.bl + (l - 1) * H;
          float* dst = a.acts + act_blk(blk, l, H, L);
          float* pdst = a.pbar + pb_blk(blk, l, H, L);
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col_of(c, g, H);
            const float av = tanh_f(acc[0][c] + bl[col]);
            const float sg = fmaf(-av, av, 1.0f);
            const float p1 = acc[1][c], p2 = acc[2][c];
            const float a1 = sg * p1;
            const float a2 = fmaf(sg, p2, -2.0f * av * a1 * p1);
            dst[(size_t)col * kTT + t] = av;
            dst[(size_t)(H + col) * kTT + t] = p1;
            dst[(size_t)(2 * H + col) * kTT + t] = p2;
            if (l + 1 < L) {
              out[(size_t)col * kTT + t] = av; out[(size_t)(H + col) * kTT + t] = a1; out[(size_t)(2 * H + col) * kTT + t] = a2;
            } else {
              // last hidden layer: fused last-layer dot and its reverse
              const float w = sm.wl[col];
              hp[0] = fmaf(w, av, hp[0]); hp[1] = fmaf(w, a1, hp[1]); hp[2] = fmaf(w, a2, hp[2]);
              G.gwl[c] += av * hb[0] + a1 * hb[1] + a2 * hb[2];
              const float ab = w * hb[0], a1b = w * hb[1], a2b = w * hb[2];
              const float p2b = a2b * sg;
              const float p1b = sg * fmaf(-4.0f * av * p1, a2b, a1b);
              const float pzb = sg * (ab - 2.0f * av * p1 * a1b + a2b * fmaf(-2.0f * av, p2, -2.0f * fmaf(-3.0f * av, av, 1.0f) * p1 * p1));
#pragma unroll
              for (int li = 0; li < kMaxL - 1; ++li) if (li == l - 1) G.gbl[li][c] += pzb;
              out[(size_t)col * kTT + t] = pzb; out[(size_t)(H + col) * kTT + t] = p1b; out[(size_t)(2 * H + col) * kTT + t] = p2b;
              pdst[(size_t)col * kTT + t] = pzb; pdst[(size_t)(H + col) * kTT + t] = p1b; pdst[(size_t)(2 * H + col) * kTT + t] = p2b;
            }
          }
          float* tmp = in; in = out; out = tmp;
        }
        // `in` now holds Pbar_{L-1} (0-based layer index L-1); walk the hidden layers backwards
        float xpart[1] = {0.f};
        for (int l = L - 1; l >= 1; --l) {
          layer_gemm<3>(acc, in, a.WTc + (size_t)(l - 1) * H * H, sm.ws, H, t, g, tid, nthreads);
          const float* src = a.acts + act_blk(blk, l - 1, H, L);
          float* pdst = l - 1 >= 1 ? a.pbar + pb_blk(blk, l - 1, H, L) : nullptr;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = col_of(c, g, H);
            const float av = src[(size_t)col * kTT + t];
            const float p1 = src[(size_t)(H + col) * kTT + t];
            const float p2 = src[(size_t)(2 * H + col) * kTT + t];
            const float sg = fmaf(-av, av, 1.0f);
            const float ab = acc[0][c], a1b = acc[1][c], a2b = acc[2][c];
            const float p2b = a2b * sg;
            const float p1b = sg * fmaf(-4.0f * av * p1, a2b, a1b);
            const float pzb = sg * (ab - 2.0f * av * p1 * a1b + a2b * fmaf(-2.0f * av, p2, -2.0f * fmaf(-3.0f * av, av, 1.0f) * p1 * p1));
            if (l - 1 >= 1) {
#pragma unroll
              for (int li = 0; li < kMaxL - 1; ++li) if (li == l - 2) G.gbl[li][c] += pzb;
              out[(size_t)col * kTT + t] = pzb; out[(size_t)(H + col) * kTT + t] = p1b; out[(size_t)(2 * H + col) * kTT + t] = p2b;
              pdst[(size_t)col * kTT + t] = pzb; pdst[(size_t)(H + col) * kTT + t] = p1b; pdst[(size_t)(2 * H + col) * kTT + t] = p2b;
            } else {
              // first layer: p = w_x x + t' w_t + b1, p' = w_x (a parameter), p'' = 0
              G.gb1[c] += pzb;
              G.gwt[c] = fmaf(tprime, pzb, G.gwt[c]);
              G.gwx[c] += fmaf(pzb, x_t, p1b);
              xpart[0] = fmaf(pzb, sm.wx[col], xpart[0]);
            }
          }
          float* tmp = in; in = out; out = tmp;
        }
        reduce_groups<3>(hp, sm.hpart, t, g, tid);
        reduce_groups<1>(xpart, sm.xpart, t, g, tid);
        __syncthreads();
        if (owner) {
          float h1 = 0.f, xbar_net = 0.f;
          for (int w = 0; w < nwarps; ++w) { h1 += sm.hpart[(w * 3 + 1) * kTT + tid]; xbar_net += sm.xpart[w * kTT + tid]; }
          const float ysx = xbar_net;
          const float yss = -a.y0 * h1 * ks;
          sumx += ysx; prevx = ysx; sums += yss; prevs = yss;
        }
      }
      if (owner) { xb += sumx; sb += sums; }
    }
    if (valid && a.ybar0 != nullptr) { float* dst = a.ybar0 + row * a.stride; dst[0] = xb; dst[1] = lb; dst[2] = sb; }
  }
  (void)b_last;
  // ---- per-CTA small-parameter partials: sum over the 8 trajectories (lanes 0..7 of each group) ----
  float* out = a.gsmall + (size_t)blockIdx.x * TL.n_small();
  auto red8 = [](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
  };
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int col = col_of(c, g, H);
    const float v1 = red8(G.gb1[c]), v2 = red8(G.gwx[c]), v3 = red8(G.gwt[c]), v4 = red8(G.gwl[c]);
    if (t == 0) { out[TL.c_b0() + col] = v1; out[TL.c_w0() + H + col] = v2; out[TL.c_w0() + col] = v3; out[TL.c_wlast() + col] = v4; }
#pragma unroll
    for (int l = 1; l < kMaxL; ++l) {
      if (l < L) {
        const float v = red8(G.gbl[l - 1][c]);
        if (t == 0) out[TL.c_b(l) + col] = v;
      }
    }
  }
  const float gl = red8(G.gblast);        // owners are lanes 0..7 of warp 0
  if (tid == 0) out[0] = gl;
}

// ------------------------------------------------------------------------------------------------
// dW_l[k][j] = sum_blocks sum_{jet,t} A_{l-1}[blk][jet][k][t] * Pbar_l[blk][jet][j][t]   (split-K over blocks)
// 64x64 output tile per CTA, 256 threads x (4x4); A jets are rebuilt from (a, p', p'') while staging.
// ------------------------------------------------------------------------------------------------
constexpr int kDwTile = 64;
__global__ void __launch_bounds__(256) mlp_dw_gemm_kernel(const float* __restrict__ acts, const float* __restrict__ pbar,
                                                         float* __restrict__ part, int H, int L, int layer,
                                                         long long n_blocks, int splits) {
  __shared__ __align__(16) float As[3 * kTT][kDwTile + 4];
  __shared__ __align__(16) float Bs[3 * kTT][kDwTile + 4];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * kDwTile, j0 = blockIdx.y * kDwTile, sp = blockIdx.z;
  const int tk = (tid >> 4) * 4, tj = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long per = (n_blocks + splits - 1) / splits;
  const long long b_begin = sp * per, b_end = b_begin + per < n_blocks ? b_begin + per : n_blocks;
  for (long long blk = b_begin; blk < b_end; ++blk) {
    const float* asrc = acts + act_blk(blk, layer - 1, H, L);
    const float* psrc = pbar + pb_blk(blk, layer, H, L);
    __syncthreads();
    // stage: 64 units x 8 traj per jet; thread -> (unit u = tid>>2, traj pair tt = (tid&3)*2)
    {
      const int u = tid >> 2, tt = (tid & 3) * 2;
      const float2 av = *reinterpret_cast<const float2*>(asrc + (size_t)(k0 + u) * kTT + tt);
      const float2 p1 = *reinterpret_cast<const float2*>(asrc + (size_t)(H + k0 + u) * kTT + tt);
      const float2 p2 = *reinterpret_cast<const float2*>(asrc + (size_t)(2 * H + k0 + u) * kTT + tt);
      const float sgx = fmaf(-av.x, av.x, 1.0f), sgy = fmaf(-av.y, av.y, 1.0f);
      const float a1x = sgx * p1.x, a1y = sgy * p1.y;
      As[tt][u] = av.x; As[tt + 1][u] = av.y;
      As[kTT + tt][u] = a1x; As[kTT + tt + 1][u] = a1y;
      As[2 * kTT + tt][u] = fmaf(sgx, p2.x, -2.0f * av.x * a1x * p1.x);
      As[2 * kTT + tt + 1][u] = fmaf(sgy, p2.y, -2.0f * av.y * a1y * p1.y);
#pragma unroll
      for (int jt = 0; jt < 3; ++jt) {
        const float2 pv = *reinterpret_cast<const float2*>(psrc + (size_t)(jt * H + j0 + u) * kTT + tt);
        Bs[jt * kTT + tt][u] = pv.x; Bs[jt * kTT + tt + 1][u] = pv.y;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3 * kTT; ++r) {
      const float4 av = *reinterpret_cast<const float4*>(&As[r][tk]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[r][tj]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
  }
  float* out = part + (size_t)sp * H * H;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(size_t)(k0 + tk + i) * H + j0 + tj + j] = acc[i][j];
}

// ------------------------------------------------------------------------------------------------
// The same contraction on the tensor cores (tcgen05, error-compensated 3xTF32, accumulators in TMEM):
//   D[(hi k | lo k) 128][j 64] += A^T-stacked (128 x 8 traj) . (P_hi, then P_lo) (8 traj x 64)      per (block, jet)
// The HBM block layout [unit][8 traj] IS the K-major operand layout up to a regrouping of 16-byte chunks, so staging
// is float4 loads -> (jets, lo parts) -> float4 stores; operands go in unrounded (the tensor core truncates to tf32,
// the remainders are the "lo" parts).  Two 48 KB stages of two blocks each, one mbarrier per stage; the TMEM
// accumulator is drained into fp32 registers every kDwFlush stages because tensor-core accumulation truncates.
// Output: rows 0..63 (hi k) and 64..127 (lo k) go to two partial buffers, summed by mlp_dw_reduce_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kDwG = 2;           // HBM blocks per pipeline stage
constexpr int kDwFlush = 8;       // stages per TMEM accumulation chain (8 x 2 x 6 = 96 MMAs)
constexpr int kDwTcSplits = 4;    // 64 tiles x 4 = 256 CTAs, two per SM; 8 partial buffers
struct DwStage {
  float A[kDwG][3][2 * 128 * 4];  // [block][jet][chunk of 4 traj][row: hi k 0..63 | lo k 64..127][4]
  float Ph[kDwG][3][2 * 64 * 4];  // [block][jet][chunk][row j][4]
  float Pl[kDwG][3][2 * 64 * 4];
};
struct DwSmem {
  DwStage st[2];
  uint64_t bar[2];
  unsigned int tmem_slot;
};

__global__ void __launch_bounds__(256, 2) mlp_dw_gemm_tc_kernel(const float* __restrict__ acts, const float* __restrict__ pbar,
                                                               float* __restrict__ part, int H, int L, int layer,
                                                               long long n_blocks, int splits) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  DwSmem& sm = *reinterpret_cast<DwSmem*>(smem_raw);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * 64, j0 = blockIdx.y * 64, sp = blockIdx.z;
  if (tid == 0) { tc::mbar_init(&sm.bar[0], 1); tc::mbar_init(&sm.bar[1], 1); tc::fence_mbar_init(); }
  if (w == 0) tc::tmem_alloc(&sm.tmem_slot, 64);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
  const long long per = (n_blocks + splits - 1) / splits;
  const long long b_begin = sp * per, b_end = b_begin + per < n_blocks ? b_begin + per : n_blocks;
  const long long n_it = b_end > b_begin ? (b_end - b_begin + kDwG - 1) / kDwG : 0;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  uint32_t phase = 0, pending = 0;      // bit s: parity / outstanding commit of stage s
  int in_group = 0;
  // staging role: threads 0..127 build the A operand (unit u, trajectory chunk c), 128..255 the P operand
  const int role_p = tid >> 7, c = (tid >> 6) & 1, u = tid & 63;
  for (long long it = 0; it < n_it; ++it) {
    const int s = (int)(it & 1);
    if (pending & (1u << s)) { tc::mbar_wait(&sm.bar[s], (phase >> s) & 1u); phase ^= 1u << s; pending &= ~(1u << s); }
    DwStage& S = sm.st[s];
#pragma unroll
    for (int g = 0; g < kDwG; ++g) {
      const long long blk = b_begin + it * kDwG + g;
      if (blk < b_end) {
        if (!role_p) {
          const float* src = acts + act_blk(blk, layer - 1, H, L) + (size_t)(k0 + u) * kTT + 4 * c;
          const float4 av = *reinterpret_cast<const float4*>(src);
          const float4 p1 = *reinterpret_cast<const float4*>(src + (size_t)H * kTT);
          const float4 p2 = *reinterpret_cast<const float4*>(src + (size_t)2 * H * kTT);
          const float a0[4] = {av.x, av.y, av.z, av.w}, q1[4] = {p1.x, p1.y, p1.z, p1.w}, q2[4] = {p2.x, p2.y, p2.z, p2.w};
          float j1[4], j2[4], l0[4], l1[4], l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float sg = fmaf(-a0[e], a0[e], 1.0f);
            j1[e] = sg * q1[e];
            j2[e] = fmaf(sg, q2[e], -2.0f * a0[e] * j1[e] * q1[e]);
            l0[e] = tc::lo_tf32(a0[e]); l1[e] = tc::lo_tf32(j1[e]); l2[e] = tc::lo_tf32(j2[e]);
          }
          float* d0 = &S.A[g][0][(c * 128 + u) * 4];
          *reinterpret_cast<float4*>(d0) = av;
          *reinterpret_cast<float4*>(d0 + 64 * 4) = make_float4(l0[0], l0[1], l0[2], l0[3]);
          float* d1 = &S.A[g][1][(c * 128 + u) * 4];
          *reinterpret_cast<float4*>(d1) = make_float4(j1[0], j1[1], j1[2], j1[3]);
          *reinterpret_cast<float4*>(d1 + 64 * 4) = make_float4(l1[0], l1[1], l1[2], l1[3]);
          float* d2 = &S.A[g][2][(c * 128 + u) * 4];
          *reinterpret_cast<float4*>(d2) = make_float4(j2[0], j2[1], j2[2], j2[3]);
          *reinterpret_cast<float4*>(d2 + 64 * 4) = make_float4(l2[0], l2[1], l2[2], l2[3]);
        } else {
          const float* src = pbar + pb_blk(blk, layer, H, L) + (size_t)(j0 + u) * kTT + 4 * c;
#pragma unroll
          for (int jt = 0; jt < 3; ++jt) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)jt * H * kTT);
            *reinterpret_cast<float4*>(&S.Ph[g][jt][(c * 64 + u) * 4]) = v;
            *reinterpret_cast<float4*>(&S.Pl[g][jt][(c * 64 + u) * 4]) =
                make_float4(tc::lo_tf32(v.x), tc::lo_tf32(v.y), tc::lo_tf32(v.z), tc::lo_tf32(v.w));
          }
        }
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (w == 0) {
      tc::fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int g = 0; g < kDwG; ++g) {
          if (b_begin + it * kDwG + g < b_end) {
#pragma unroll
            for (int jt = 0; jt < 3; ++jt) {
              const uint64_t ad = tc::make_desc(tc::smem_u32(S.A[g][jt]), 128 * 16, 128);
              tc::mma_tf32(tmem, ad, tc::make_desc(tc::smem_u32(S.Ph[g][jt]), 64 * 16, 128), idesc, (in_group | g | jt) != 0);
              tc::mma_tf32(tmem, ad, tc::make_desc(tc::smem_u32(S.Pl[g][jt]), 64 * 16, 128), idesc, true);
            }
          }
        }
        tc::commit(&sm.bar[s]);
      }
      __syncwarp();
    }
    pending |= 1u << s;
    ++in_group;
    if (in_group == kDwFlush || it == n_it - 1) {
      // drain the chain into the fp32 accumulators: lane quadrant w&3 <-> rows, warp half w>>2 <-> 32 columns
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2)
        if (pending & (1u << s2)) { tc::mbar_wait(&sm.bar[s2], (phase >> s2) & 1u); phase ^= 1u << s2; pending &= ~(1u << s2); }
      __syncwarp();
      tc::fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float t[16];
        tc::tmem_ld16(tmem + ((uint32_t)(32 * (w & 3)) << 16) + 32 * (w >> 2) + c0, t);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[c0 + i] += t[i];
      }
      tc::fence_before_sync();
      __syncthreads();
      in_group = 0;
    }
  }
  {
    const int r = 32 * (w & 3) + lane;                       // accumulator row: hi k (r < 64) or lo k
    float* out = part + ((size_t)(2 * sp + (r >> 6)) * H + (k0 + (r & 63))) * H + j0 + 32 * (w >> 2);
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// 128 x 128 output tiles (H a multiple of 128), warp specialised.  With both operands in shared memory a (block, jet) slab of
// 8 trajectories costs 16 KB of operand stores and 24 KB of MMA operand reads per 192 tensor-core clocks -- 208 B/clk
// against the 128 B/clk of the shared-memory port -- so the activation operand goes through TENSOR MEMORY instead:
//   warps 0..7  (A role, thread = unit k = TMEM lane; the two warps of a lane quadrant take the even / the odd blocks):
//               (a, p', p'') of 8 trajectories -> operand jets (a, a', a'') and their tf32 remainders -> tcgen05.st, per
//               slot [jet][hi 8 | lo 8] = 48 columns; they also keep the 64 fp32 accumulators per thread;
//   warps 8..11 (P role): cotangent jets as stored -> (hi, lo) K-major operands in shared memory;
//   warp 12    : D[k 128][j 128] += A_hi . P_hi + A_lo . P_hi + A_hi . P_lo per (block, jet)  (TS-mode M128 N128 K8; the
//               lo.lo term is dropped), tcgen05.commit frees the slot.  setmaxnreg moves the register file to where it is
//               needed: 176 (A) / 120 (P) / 40 (MMA warpgroup).
// The HBM loads run 2 (A: 4 blocks of the stream) / 3 (P) blocks ahead in registers; four operand slots; two accumulator
// sets of 128 columns alternate every kDw3Flush blocks and are drained into fp32 registers by the A warps because
// tensor-core accumulation truncates.  Blocks whose cotangent jets p', p'' are exactly zero (rows whose logp / score the
// loss does not use; flagged per (evaluation, 8-row tile) by the sweep's head kernel) are contracted over jet 0 only.
// One CTA per SM: 16 tiles x 9 splits, or 16 x 5 beside the sweep.
// ------------------------------------------------------------------------------------------------
#ifndef OFDFT_DW_SLOTS
#define OFDFT_DW_SLOTS 4
#endif
constexpr int kDw3Slots = OFDFT_DW_SLOTS;
constexpr int kDw3Flush = 32;     // blocks per TMEM accumulation chain: 32 x 9 = 288 MMAs
constexpr int kDw3Splits = 9;     // 16 tiles x 9 = 144 CTAs on 148 SMs
constexpr int kDw3Threads = 16 * 32;   // two activation-operand warpgroups, the cotangent-operand warpgroup, the MMA-issue warpgroup
constexpr int kDw3ColA = 256;     // TMEM: accumulator sets at 0 and 128, operand ring from 256
// SL = true: the operands are read straight from per-evaluation layer buffers ([jet][H/4][rows_pad][4], what the layer GEMMs
// read and write anyway) instead of HBM blocks: block b = (evaluation b / n_tiles, rows 8 (b % n_tiles)..).  A thread's unit
// owns eight 4-byte words 16 bytes apart there; loading them as such costs 8 L1 tag lookups per warp instruction (192 per
// warp and block -- slower than the tensor core consumes them), so each lane quad loads the 4 units x 8 rows of its unit
// group as float4 (4 units of one row) and transposes 4x4 in registers with four shuffles.
struct DwSrc {
  const float* a; const float* p;                 // layer buffers of evaluation 0: activations of layer - 1, cotangents of layer
  size_t a_ev, p_ev;                              // strides between evaluations (floats)
  long long rows_pad, rows; int n_tiles;
  const unsigned char* live;                      // per block: 0 = the cotangent jets p', p'' of all 8 rows are exactly zero (rows
                                                  // whose logp / score outputs the loss does not use): only jet 0 is contracted
};
// 4x4 transpose across the four lanes of a quad: lane i holds row i (x, y, z, w = columns 0..3) and ends with column i
__device__ __forceinline__ float4 quad_transpose(float4 v, int i) {
  const bool hi = (i & 2) != 0, odd = (i & 1) != 0;
  // swap the off-diagonal 2x2 blocks between lanes i and i ^ 2
  const float r0 = __shfl_xor_sync(0xffffffffu, hi ? v.x : v.z, 2), r1 = __shfl_xor_sync(0xffffffffu, hi ? v.y : v.w, 2);
  if (hi) { v.x = r0; v.y = r1; } else { v.z = r0; v.w = r1; }
  // swap the off-diagonal elements of every 2x2 block between lanes i and i ^ 1
  const float q0 = __shfl_xor_sync(0xffffffffu, odd ? v.x : v.y, 1), q1 = __shfl_xor_sync(0xffffffffu, odd ? v.z : v.w, 1);
  if (odd) { v.x = q0; v.z = q1; } else { v.y = q0; v.w = q1; }
  return v;
}

constexpr int kDw3Sbo = 36;                       // floats between 8-unit core matrices of the cotangent operand (144 B)
constexpr int kDw3Lbo = 16 * kDw3Sbo;             // floats between its 4-row chunks (128 units)
struct Dw3Smem {
  // cotangent operand, K-major: [slot][jet][chunk of 4 rows][group of 8 units j: 36 floats][unit: 4 rows].  The 8-unit core
  // matrices are 144 B apart (SBO; 16 B of padding) so that the layer-buffer path can store a float4 = 4 units of one row
  // as four scalar stores without bank conflicts (lane quad -> row within the chunk, 8 quads -> 8 unit groups)
  float Ph[kDw3Slots][3][2 * kDw3Lbo];
  float Pl[kDw3Slots][3][2 * kDw3Lbo];
  uint64_t full[kDw3Slots], empty[kDw3Slots], accf[2], acce[2];
  unsigned int tmem_slot;
  int nj[kDw3Slots];                              // jets staged in the slot (staging warps -> MMA warp)
};

// staging loop of one warp.  ROLE_P = false: activation operand -> tensor memory; the two warps of a lane quadrant take the
// even / the odd blocks (one warp alone is latency-bound at ~2400 clocks per block against 576 of MMA) and both keep 64
// fp32 accumulators of the tile.  ROLE_P = true: cotangent operand -> shared memory, every block.
template <bool SL, bool ROLE_P>
__device__ __forceinline__ void dw3_staging(Dw3Smem& sm, const float* __restrict__ acts, const float* __restrict__ pbar,
                                            float* __restrict__ part, int H, int L, int layer, long long b_begin, int n_it,
                                            const DwSrc& src, int accumulate, uint32_t tmem, int w, int lane, int k0, int j0, int sp) {
  constexpr int PF = ROLE_P ? 3 : 2;                           // blocks of this warp in flight (HBM loads in registers)
  const int step = ROLE_P ? 1 : 2, it_first = ROLE_P ? 0 : (w >> 2);
  const int u = 32 * (w & 3) + lane;                           // unit row of the tile (A role: TMEM lane)
  const uint32_t tl = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  float acc[ROLE_P ? 1 : 64];                                  // A role: row 32 (w & 3) + lane, columns 64 (w >> 2) ..
#pragma unroll
  for (int i = 0; i < (ROLE_P ? 1 : 64); ++i) acc[i] = 0.f;
  const size_t blk_stride = ROLE_P ? (size_t)(L - 1) * 3 * H * kTT : (size_t)L * 3 * H * kTT;
  const int unit = (ROLE_P ? j0 : k0) + u;
  // SL: lane i of the quad loads rows i and i + 4 of the quad's unit group
  const float* src0 = SL ? (ROLE_P ? src.p : src.a) + ((size_t)(unit >> 2) * src.rows_pad + (lane & 3)) * 4
                         : (ROLE_P ? pbar + pb_blk(b_begin, layer, H, L) : acts + act_blk(b_begin, layer - 1, H, L)) + (size_t)unit * kTT;
  const size_t jet = SL ? (size_t)src.rows_pad * H : (size_t)H * kTT;
  const size_t ev_stride = ROLE_P ? src.p_ev : src.a_ev;
  float4 pf[PF][3][2];
  int pfn[PF];                                                 // jets of the block in each prefetch slot (1 or 3)
#pragma unroll
  for (int q = 0; q < PF; ++q) pfn[q] = 3;
  // (evaluation, tile) of the next block this warp loads: its blocks are loaded in order, one division up front
  int ld_ev = 0, ld_tile = 0;
  if (SL) { ld_ev = (int)((b_begin + it_first) / src.n_tiles); ld_tile = (int)(b_begin + it_first - (long long)ld_ev * src.n_tiles); }
  auto load_blk = [&](float4 (&R)[3][2], int& nj, int it) {
    nj = 3;
    if (SL) {
      const long long row0 = (long long)ld_tile * kTT;
      const float* sp_ = src0 + (size_t)ld_ev * ev_stride + (size_t)row0 * 4;
      const long long r_lo = row0 + (lane & 3);                // rows beyond the batch contribute nothing
      if (src.live != nullptr) {                               // one byte for the whole CTA; the shuffle tells the compiler so
        const int lv = __shfl_sync(0xffffffffu, (int)src.live[(size_t)ld_ev * src.n_tiles + ld_tile], 0);
        if (lv == 0) nj = 1;
      }
      ld_tile += step;
      while (ld_tile >= src.n_tiles) { ld_tile -= src.n_tiles; ++ld_ev; }
#pragma unroll
      for (int jt = 0; jt < 3; ++jt) {
        if (jt >= nj) break;
        R[jt][0] = r_lo < src.rows ? __ldg(reinterpret_cast<const float4*>(sp_ + jt * jet)) : make_float4(0.f, 0.f, 0.f, 0.f);
        R[jt][1] = r_lo + 4 < src.rows ? __ldg(reinterpret_cast<const float4*>(sp_ + jt * jet + 16)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      return;
    }
    const float* sp_ = src0 + (size_t)it * blk_stride;
#pragma unroll
    for (int jt = 0; jt < 3; ++jt) {
      R[jt][0] = __ldg(reinterpret_cast<const float4*>(sp_ + jt * jet));
      R[jt][1] = __ldg(reinterpret_cast<const float4*>(sp_ + jt * jet + 4));
    }
  };
  // accumulation chain c (blocks 32 c .. 32 c + 31, accumulator set c & 1) -> fp32 registers
  auto drain = [&](int c) {
    const int set = c & 1;
    tc::mbar_wait(&sm.accf[set], (uint32_t)(c >> 1) & 1u);
    __syncwarp();
    tc::fence_after_sync();
#pragma unroll
    for (int c0 = 0; c0 < (ROLE_P ? 0 : 64); c0 += 16) {
      float t[16];
      tc::tmem_ld16(tl + 128 * set + 64 * (w >> 2) + c0, t);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[c0 + i] += t[i];
    }
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&sm.acce[set]);
  };
  int drained = 0;
#pragma unroll
  for (int q = 0; q < PF; ++q)
    if (it_first + q * step < n_it) load_blk(pf[q], pfn[q], it_first + q * step);
  const tc::f32x2 one2 = tc::splat2(1.0f), m1 = tc::splat2(-1.0f), two2 = tc::splat2(2.0f);
  auto lo2x = [&](tc::f32x2 x) { return tc::fma2(x & 0xFFFFE000FFFFE000ull, m1, x); };      // x - trunc_tf32(x), two lanes
  // one block: registers R (A role: this thread's unit, 8 rows x 3 jets after the transposition) -> operand slot s
  auto stage = [&](float4 (&R)[3][2], int nj, int it) {
    using namespace tc;
    const int s = it % kDw3Slots;
    if (SL && !ROLE_P) {                                       // tensor memory wants lane = unit: transpose the quad's 4 x 4 tiles
#pragma unroll
      for (int jt = 0; jt < 3; ++jt) {
        if (jt >= nj) break;
        R[jt][0] = quad_transpose(R[jt][0], lane & 3); R[jt][1] = quad_transpose(R[jt][1], lane & 3);
      }
    }
    // use k = it / 4 of the slot: free once the MMAs of use k - 1 have completed
    if (it >= kDw3Slots) mbar_wait(&sm.empty[s], (uint32_t)(it / kDw3Slots - 1) & 1u);
    if (!ROLE_P) {
      __syncwarp();
      fence_after_sync();
      const uint32_t col = tl + kDw3ColA + 48 * s;
      if (nj == 1) {                                           // only the activations themselves are contracted
        float v0[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 A4 = R[0][e >> 1];
          const f32x2 av = (e & 1) ? pack2(A4.z, A4.w) : pack2(A4.x, A4.y);
          unpack2(av, v0[2 * e], v0[2 * e + 1]); unpack2(lo2x(av), v0[8 + 2 * e], v0[9 + 2 * e]);
        }
        tmem_st16(col, v0);
      } else {
        float v0[16], v1[16], v2[16];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 A4 = R[0][e >> 1], P4 = R[1][e >> 1], Q4 = R[2][e >> 1];
          const f32x2 av = (e & 1) ? pack2(A4.z, A4.w) : pack2(A4.x, A4.y);
          const f32x2 p1 = (e & 1) ? pack2(P4.z, P4.w) : pack2(P4.x, P4.y);
          const f32x2 p2 = (e & 1) ? pack2(Q4.z, Q4.w) : pack2(Q4.x, Q4.y);
          const f32x2 na = mul2(av, m1);
          const f32x2 sg = fma2(na, av, one2);                                  // 1 - a^2
          const f32x2 j1 = mul2(sg, p1);                                        // a' = (1 - a^2) p'
          const f32x2 j2 = fma2(sg, p2, mul2(mul2(mul2(na, two2), j1), p1));    // a'' = (1 - a^2) p'' - 2 a a' p'
          const f32x2 l0 = lo2x(av), l1 = lo2x(j1), l2 = lo2x(j2);
          unpack2(av, v0[2 * e], v0[2 * e + 1]); unpack2(l0, v0[8 + 2 * e], v0[9 + 2 * e]);
          unpack2(j1, v1[2 * e], v1[2 * e + 1]); unpack2(l1, v1[8 + 2 * e], v1[9 + 2 * e]);
          unpack2(j2, v2[2 * e], v2[2 * e + 1]); unpack2(l2, v2[8 + 2 * e], v2[9 + 2 * e]);
        }
        tmem_st16(col, v0);
        tmem_st16(col + 16, v1);
        tmem_st16(col + 32, v2);
      }
      tmem_st_wait();
      fence_before_sync();
    } else {
      if (w == 8 && lane == 0) sm.nj[s] = nj;                  // read by the MMA warp after this slot's full barrier
#pragma unroll
      for (int jt = 0; jt < 3; ++jt) {
        if (jt >= nj) break;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 v = R[jt][c];
          const f32x2 lxy = lo2x(pack2(v.x, v.y)), lzw = lo2x(pack2(v.z, v.w));
          const float4 l = make_float4(lo2(lxy), hi2(lxy), lo2(lzw), hi2(lzw));
          if (SL) {
            // v = units 4 (u >> 2) .. + 3 at row 4 c + (lane & 3): element (unit n, row) sits at chunk c, group n / 8, n % 8, row % 4
            const int n0 = u & ~3;
            float* ph = &sm.Ph[s][jt][c * kDw3Lbo + (n0 >> 3) * kDw3Sbo + (n0 & 7) * 4 + (lane & 3)];
            float* pl = &sm.Pl[s][jt][c * kDw3Lbo + (n0 >> 3) * kDw3Sbo + (n0 & 7) * 4 + (lane & 3)];
            ph[0] = v.x; ph[4] = v.y; ph[8] = v.z; ph[12] = v.w;
            pl[0] = l.x; pl[4] = l.y; pl[8] = l.z; pl[12] = l.w;
          } else {
            const int off = c * kDw3Lbo + (u >> 3) * kDw3Sbo + (u & 7) * 4;      // v = rows 4 c .. + 3 of unit u
            *reinterpret_cast<float4*>(&sm.Ph[s][jt][off]) = v;
            *reinterpret_cast<float4*>(&sm.Pl[s][jt][off]) = l;
          }
        }
      }
      fence_async_smem();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.full[s]);
    // chain c is drained a few blocks into chain c + 1, when its MMAs have long finished
    if (!ROLE_P) {
      while (it >= 3 && drained < (it - 3) / kDw3Flush) { drain(drained); ++drained; }
    }
  };
#pragma unroll 1
  for (int it0 = it_first; it0 < n_it; it0 += PF * step) {
#pragma unroll
    for (int q = 0; q < PF; ++q) {                             // static register indices: no rotation moves
      const int it = it0 + q * step;
      if (it < n_it) {
        // the loop is bound by the latency of these loads: the next ones go out BEFORE this block is staged (4.8 -> 4.2 ms
        // per step against issuing them afterwards; an additional L2 prefetch 8..32 blocks ahead changed nothing)
        float4 R[3][2];
        const int nj = pfn[q];
#pragma unroll
        for (int jt = 0; jt < 3; ++jt) { R[jt][0] = pf[q][jt][0]; R[jt][1] = pf[q][jt][1]; }
        if (it + PF * step < n_it) load_blk(pf[q], pfn[q], it + PF * step);
        stage(R, nj, it);
      }
    }
  }
  if (!ROLE_P) {
    const int n_chains = (n_it + kDw3Flush - 1) / kDw3Flush;
    while (drained < n_chains) { drain(drained); ++drained; }
    const int r = 32 * (w & 3) + lane;
    float* out = part + ((size_t)sp * H + (k0 + r)) * H + j0 + 64 * (w >> 2);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      float4 o = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      if (accumulate) {                                        // a later chunk of evaluations (same CTA <-> same slots, launches ordered)
        const float4 p = *reinterpret_cast<const float4*>(out + i);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      *reinterpret_cast<float4*>(out + i) = o;
    }
  }
}

template <bool SL>
__global__ void __launch_bounds__(kDw3Threads, 1) mlp_dw_gemm_ts_kernel(const float* __restrict__ acts, const float* __restrict__ pbar,
                                                                       float* __restrict__ part, int H, int L, int layer,
                                                                       long long blk0, long long n_blocks, int splits, const DwSrc src,
                                                                       int accumulate) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Dw3Smem& sm = *reinterpret_cast<Dw3Smem*>(smem_raw);
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role branches (and the shuffles
  // inside them) free of convergence barriers
  const int tid = threadIdx.x, w = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int k0 = blockIdx.x * 128, j0 = blockIdx.y * 128, sp = blockIdx.z;
  if (tid == 0) {
    for (int i = 0; i < kDw3Slots; ++i) { tc::mbar_init(&sm.full[i], 8); tc::mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&sm.accf[i], 1); tc::mbar_init(&sm.acce[i], 8); }
    tc::fence_mbar_init();
  }
  if (w == 0) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const long long per = (n_blocks + splits - 1) / splits;
  const long long b_begin = blk0 + sp * per, b_end = blk0 + (sp * per + per < n_blocks ? sp * per + per : n_blocks);
  const int n_it = b_end > b_begin ? (int)(b_end - b_begin) : 0;

  // warps 0..7: activation operand (two per lane quadrant, alternating blocks) + the fp32 accumulators; warps 8..11: cotangent
  // operand; warp 12: MMA issue (13..15 only donate their registers).  512 threads x 128 registers at launch; asking for more
  // than was allocated would make setmaxnreg.inc wait forever
  static_assert(256 * 176 + 128 * 120 + 128 * 40 <= 128 * kDw3Threads, "setmaxnreg budget exceeds the registers allocated at launch");
  if (w >= 12) {
    // ---------------- MMA issue warp ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const uint32_t idesc = tc::make_idesc_tf32(128, 128, 0, 0);
    uint32_t ph_full = 0, ph_acce = 0;
#pragma unroll 1
    for (int it = 0; it < (w == 12 ? n_it : 0); ++it) {
      const int s = it % kDw3Slots, set = (it / kDw3Flush) & 1;
      const bool first = (it % kDw3Flush) == 0;
      tc::mbar_wait(&sm.full[s], (ph_full >> s) & 1u); ph_full ^= 1u << s;
      if (first && it >= 2 * kDw3Flush) { tc::mbar_wait(&sm.acce[set], (ph_acce >> set) & 1u); ph_acce ^= 1u << set; }
      __syncwarp();
      tc::fence_after_sync();
      const int nj = *reinterpret_cast<volatile int*>(&sm.nj[s]);
      if (tc::elect_one()) {
        const uint32_t d = tmem + 128 * set;
#pragma unroll
        for (int jt = 0; jt < 3; ++jt) {
          if (jt >= nj) break;
          const uint32_t ahi = tmem + kDw3ColA + 48 * s + 16 * jt, alo = ahi + 8;
          const uint64_t ph = tc::make_desc(tc::smem_u32(sm.Ph[s][jt]), kDw3Lbo * 4, kDw3Sbo * 4);
          const uint64_t pl = tc::make_desc(tc::smem_u32(sm.Pl[s][jt]), kDw3Lbo * 4, kDw3Sbo * 4);
          tc::mma_tf32_ts(d, ahi, ph, idesc, !(first && jt == 0));
          tc::mma_tf32_ts(d, alo, ph, idesc, true);
          tc::mma_tf32_ts(d, ahi, pl, idesc, true);
        }
        tc::commit(&sm.empty[s]);
        if ((it + 1) % kDw3Flush == 0 || it + 1 == n_it) tc::commit(&sm.accf[set]);
      }
      __syncwarp();
    }
  } else if (w >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 120;");
    dw3_staging<SL, true>(sm, acts, pbar, part, H, L, layer, b_begin, n_it, src, accumulate, tmem, w, lane, k0, j0, sp);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    dw3_staging<SL, false>(sm, acts, pbar, part, H, L, layer, b_begin, n_it, src, accumulate, tmem, w, lane, k0, j0, sp);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (w == 0) tc::tmem_dealloc(tmem, 512);
}

// theta_bar[w(l) + i] (+)= sum_splits part[s][i]
__global__ void mlp_dw_reduce_kernel(const float* __restrict__ part, float* __restrict__ theta_bar, int off, int n, int splits,
                                     int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = accumulate ? (double)theta_bar[off + i] : 0.0;
  for (int k = 0; k < splits; ++k) s += (double)part[(size_t)k * n + i];
  theta_bar[off + i] = (float)s;
}

// small parameters: theta_bar[p] (+)= sum_ctas gsmall[c][p] for every p outside the hidden kernels
__global__ void mlp_small_reduce_kernel(const float* __restrict__ gsmall, float* __restrict__ theta_bar, int H, int L, int nblk,
                                        int accumulate) {
  const ThetaLayout TL(H, L);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = TL.n_small();
  if (c >= ns) return;
  const int p = TL.small_to_theta(c);                          // hidden kernels come from the dW GEMM
  double s = accumulate ? (double)theta_bar[p] : 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)gsmall[(size_t)b * ns + c];
  theta_bar[p] = (float)s;
}

// ---- weight-gradient GEMM overlapped with the adjoint sweep (activation-checkpoint path) ----------------------------------
// The sweep is a dependent chain of launches that keeps (rows / 128) x (H / 64) SMs busy; the rest of the GPU contracts the
// evaluations the sweep has already finished, on a second stream, in groups of kDwGroup evaluations.
constexpr int kDwGroup = 4;
// One library-owned stream and event pair per device, created on first use; `busy` serialises the host-side enqueueing of
// calls that share them (the events are re-recorded per group of evaluations; waits already enqueued keep their meaning).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; std::mutex busy; };
static SideStream* side_stream() {
  static SideStream side[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  SideStream& ss = side[dev];
  if (ss.s == nullptr) {
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) != cudaSuccess) { ss.s = nullptr; return nullptr; }
    if (cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &ss;
}
struct DwOverlapCtx {
  cudaStream_t main; SideStream* side; float* part; const float* actS; const float* pbar;
  int H, L, splits, launched; long long rows, n_tiles; const unsigned char* live;
};
static int dw_overlap_launch(void* vctx, int ev_lo, int ev_hi) {
  DwOverlapCtx& c = *static_cast<DwOverlapCtx*>(vctx);
  cudaError_t e = cudaEventRecord(c.side->fork, c.main);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(c.side->s, c.side->fork, 0);
  if (e != cudaSuccess) return (int)e;
  const long long rows_pad = (c.rows + 127) / 128 * 128;
  const size_t sbuf = al256(3 * (size_t)rows_pad * c.H * 4) / 4;
  const dim3 gg(c.H / 128, c.H / 128, c.splits);
#ifdef OFDFT_ABL_NODW       // timing ablation only (results invalid)
  return 0;
#endif
  for (int l = 1; l < c.L; ++l) {
    DwSrc src{c.actS + (size_t)(l - 1) * sbuf, c.pbar + (size_t)(l - 1) * sbuf, (size_t)c.L * sbuf, (size_t)(c.L - 1) * sbuf, rows_pad, c.rows,
              (int)c.n_tiles, c.live};
    mlp_dw_gemm_ts_kernel<true><<<gg, kDw3Threads, sizeof(Dw3Smem) + 1024, c.side->s>>>(
        nullptr, nullptr, c.part + (size_t)(l - 1) * c.splits * c.H * c.H, c.H, c.L, l, (long long)ev_lo * c.n_tiles,
        (long long)(ev_hi - ev_lo) * c.n_tiles, c.splits, src, c.launched);
  }
  c.launched = 1;
  return (int)cudaGetLastError();
}

static int check_shape(int64_t B, int jets, int stride, int H, int L, int n_steps) {
  if (B < 0 || n_steps <= 0 || jets < 1 || jets > 3) return OFDFT_EINVAL;
  if (stride < (jets == 3 ? 3 : jets)) return OFDFT_EINVAL;
  if (H < 64 || H > 512 || H % 64 || L < 2 || L > kMaxL) return OFDFT_EUNSUPPORTED;
  return 0;
}

}  // namespace mlp
}  // namespace ofdft

using namespace ofdft;
using namespace ofdft::mlp;

extern "C" size_t ofdft_cnf_mlp1d_ckpt_bytes(int64_t B, int32_t n_steps) {
  return (size_t)n_steps * 4 * 2 * (size_t)B * sizeof(float);
}
extern "C" size_t ofdft_cnf_mlp1d_act_ckpt_bytes(int64_t B, int32_t H, int32_t n_layers, int32_t n_steps) {
  if (B <= 0 || H < 64 || H > 512 || H % 64 || n_layers < 2 || n_layers > kMaxL || n_steps <= 0) return 0;
  const ActLayout AL(B, H, n_layers, n_steps);
  return AL.ok && H % 128 == 0 ? AL.total : 0;
}
extern "C" size_t ofdft_cnf_mlp1d_scratch_bytes(int64_t B, int32_t H, int32_t n_layers, int32_t n_steps) {
  if (B <= 0 || H <= 0 || n_layers < 2 || n_layers > kMaxL || n_steps <= 0) return 256;
  return ScratchLayout(B, H, n_layers, n_steps).total;
}

extern "C" int ofdft_cnf_mlp1d_fwd(void* stream, const float* theta, const float* y_in, float* y_out, float* ckpt,
                                   void* scratch, size_t scratch_bytes, int64_t B, int32_t jets, int32_t stride,
                                   int32_t H, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign,
                                   int32_t path) {
  int rc = check_shape(B, jets, stride, H, n_layers, n_steps);
  if (rc) return rc;
  if (B == 0) return 0;                                        // empty shard: nothing to integrate
  if (!theta || !y_in || !y_out || !scratch) return OFDFT_EINVAL;
  const ScratchLayout SL(B, H, n_layers, n_steps);
  // default: per-layer tcgen05 3xTF32 GEMM launches (mlp1d_tc.cu); OFDFT_PATH_FFMA: the fused FP32 kernel of this file
  const bool fwd_tc = !(path & OFDFT_PATH_FFMA);
  if (scratch_bytes < (fwd_tc ? SL.fwd_end : SL.wtc)) return OFDFT_ESCRATCH;
  const bool act = (path & OFDFT_PATH_ACT_CKPT) != 0;
  const ActLayout AL(B, H, n_layers, n_steps);
  if (act && H % 128) return OFDFT_EUNSUPPORTED;
  if (act && (!fwd_tc || !ckpt || jets != 3 || (path & OFDFT_PATH_FFMA_DW))) return OFDFT_EINVAL;
  if (act && !AL.ok) return OFDFT_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  float* Wc = (float*)((char*)scratch + SL.wc);
  float* WTc = (float*)((char*)scratch + SL.wtc);
  float* WTlo = (float*)((char*)scratch + SL.wtlo);
  mlp_prep_kernel<<<256, 256, 0, st>>>(theta, Wc, fwd_tc ? WTc : nullptr, H, n_layers, fwd_tc ? WTlo : nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (fwd_tc) {
    const int rcp = launch_mlp_prep_tc(stream, theta, WTc, WTlo, nullptr, nullptr, H, n_layers);
    if (rcp) return rcp;
    const size_t sbuf = al256(3 * (size_t)((B + 127) / 128 * 128) * H * 4);
    float* S0 = (float*)((char*)scratch + SL.fwd_s);
    float* S1 = (float*)((char*)scratch + SL.fwd_s + sbuf);
    float* state = (float*)((char*)scratch + SL.fwd_state);
    return launch_mlp_fwd_tc(stream, theta, WTc, WTlo, y_in, y_out, ckpt, S0, S1, state, B, jets, stride, H, n_layers, n_steps, t0, t1,
                             y0sign, act ? ckpt + AL.s / 4 : nullptr);
  }
  const size_t smem = smem_bytes(H);
  e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  Args a{};
  a.theta = theta; a.Wc = Wc; a.y_in = y_in; a.y_out = y_out; a.ckpt = ckpt; a.B = B; a.row_begin = 0; a.rows = B;
  a.jets = jets; a.stride = stride; a.H = H; a.L = n_layers; a.n_steps = n_steps; a.t0 = t0; a.t1 = t1; a.y0 = y0sign;
  const long long n_tiles = (B + kTT - 1) / kTT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  mlp_fwd_kernel<<<grid, H / 2, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// One evaluation of the vector field (Gen_CNFSimpleMLP.apply, cn_flows.py:179-182) or of the score-augmented right-hand side
// (jax_ode.py:80-98): the fused FP32 kernel with a single stage.
extern "C" int ofdft_cnf_mlp1d_rhs(void* stream, const float* theta, const float* y, float* dy, void* scratch, size_t scratch_bytes,
                                   int64_t B, int32_t jets, int32_t stride, int32_t H, int32_t n_layers, float t, float y0sign) {
  int rc = check_shape(B, jets, stride, H, n_layers, 1);
  if (rc) return rc;
  if (B == 0) return 0;
  if (!theta || !y || !dy || !scratch) return OFDFT_EINVAL;
  const ScratchLayout SL(B, H, n_layers, 1);
  if (scratch_bytes < SL.wtc) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  float* Wc = (float*)((char*)scratch + SL.wc);
  mlp_prep_kernel<<<256, 256, 0, st>>>(theta, Wc, nullptr, H, n_layers, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t smem = smem_bytes(H);
  e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  Args a{};
  a.theta = theta; a.Wc = Wc; a.y_in = y; a.y_out = dy; a.ckpt = nullptr; a.B = B; a.row_begin = 0; a.rows = B;
  a.jets = jets; a.stride = stride; a.H = H; a.L = n_layers; a.n_steps = 1; a.t0 = 0.f; a.t1 = 1.f; a.y0 = y0sign;
  a.rhs_only = 1; a.t_rhs = t;
  const long long n_tiles = (B + kTT - 1) / kTT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  mlp_fwd_kernel<<<grid, H / 2, smem, st>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_cnf_mlp1d_bwd(void* stream, const float* theta, const float* ckpt, const float* ybar1,
                                   float* theta_bar, float* ybar0, void* scratch, size_t scratch_bytes,
                                   int64_t B, int32_t jets, int32_t stride, int32_t H, int32_t n_layers,
                                   int32_t n_steps, float t0, float t1, float y0sign, int32_t path) {
  int rc = check_shape(B, jets, stride, H, n_layers, n_steps);
  if (rc) return rc;
  if (jets != 3) return OFDFT_EUNSUPPORTED;                    // the adjoint is instantiated for the training path
  if (B == 0) {                                                // empty shard: the gradient of an empty sum
    if (!theta_bar) return OFDFT_EINVAL;
    return (int)cudaMemsetAsync(theta_bar, 0, (size_t)ThetaLayout(H, n_layers).n * sizeof(float), (cudaStream_t)stream);
  }
  if (!theta || !ckpt || !ybar1 || !theta_bar || !scratch) return OFDFT_EINVAL;
  const int L = n_layers;
  const ScratchLayout SL(B, H, L, n_steps);
  if (scratch_bytes < SL.total) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  char* sc = (char*)scratch;
  float* Wc = (float*)(sc + SL.wc); float* WTc = (float*)(sc + SL.wtc);
  float* gsmall = (float*)(sc + SL.gsmall); float* part = (float*)(sc + SL.part);
  // default: adjoint sweep as per-layer tcgen05 GEMM launches (mlp1d_tc.cu); OFDFT_PATH_FFMA: fused FP32 sweep
  const bool bwd_tc = !(path & OFDFT_PATH_FFMA);
  const bool act = (path & OFDFT_PATH_ACT_CKPT) != 0;
  const ActLayout AL(B, H, L, n_steps);
  if (act && H % 128) return OFDFT_EUNSUPPORTED;
  if (act && (!bwd_tc || (path & OFDFT_PATH_FFMA_DW))) return OFDFT_EINVAL;
  if (act && !AL.ok) return OFDFT_EUNSUPPORTED;
  float* WTlo = (float*)(sc + SL.wtlo); float* Wclo = (float*)(sc + SL.wclo);
  mlp_prep_kernel<<<256, 256, 0, st>>>(theta, Wc, WTc, H, L, bwd_tc ? WTlo : nullptr, bwd_tc ? Wclo : nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const size_t smem = smem_bytes(H);
  e = cudaFuncSetAttribute(mlp_bwd_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms > 256) sms = 256;
  const ThetaLayout TL(H, L);
  int chunk_idx = 0;
  for (long long r0 = 0; r0 < B; r0 += kChunkRows, ++chunk_idx) {
    const long long rows = B - r0 < kChunkRows ? B - r0 : kChunkRows;
    const long long n_tiles = (rows + kTT - 1) / kTT;
    const long long n_blocks = (long long)n_steps * 4 * n_tiles;
    Args a{};
    a.theta = theta; a.Wc = Wc; a.WTc = WTc; a.ckpt = const_cast<float*>(ckpt); a.ybar1 = ybar1; a.ybar0 = ybar0;
    a.acts = (float*)(sc + SL.acts); a.pbar = (float*)(sc + SL.pbar); a.gsmall = gsmall;
    a.B = B; a.row_begin = r0; a.rows = rows; a.jets = 3; a.stride = stride; a.H = H; a.L = L; a.n_steps = n_steps;
    a.t0 = t0; a.t1 = t1; a.y0 = y0sign;
    int grid = (int)(n_tiles < sms ? n_tiles : sms);
    // activation-checkpoint path: the weight-gradient GEMM runs beside the sweep on the SMs the sweep leaves idle
    DwOverlapCtx ov{};
    MlpDwHook hook{};
    std::unique_lock<std::mutex> side_lock;
    if (act) {
      const int gemm_ctas = (int)((rows + 127) / 128) * (H / 64), tiles = (H / 128) * (H / 128);
      int ov_splits = (sms - gemm_ctas) / tiles;
      if (ov_splits > kDwSplits / (L - 1)) ov_splits = kDwSplits / (L - 1);
#ifdef OFDFT_DW_OV_SPLITS   // development: cap the SMs the overlapped weight-gradient GEMM takes
      if (ov_splits > OFDFT_DW_OV_SPLITS) ov_splits = OFDFT_DW_OV_SPLITS;
#endif
#ifdef OFDFT_ABL_DWEND      // timing ablation: weight-gradient GEMM after the sweep instead of beside it
      ov_splits = 0;
#endif
      // the GEMM only hides behind the sweep if it keeps up with it: at 1024 rows it needs 64 of its CTAs (measured: 4 splits
      // 4.15 ms, 3 splits 5.3 ms against 3.7 ms of sweep), and its work grows with the rows while the sweep's time does not
      if ((long long)ov_splits * 1024 < 4 * rows) ov_splits = 0;
      SideStream* side = ov_splits >= 1 ? side_stream() : nullptr;
      if (side != nullptr) {
        side_lock = std::unique_lock<std::mutex>(side->busy);
        e = cudaFuncSetAttribute(mlp_dw_gemm_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dw3Smem) + 1024);
        if (e != cudaSuccess) return (int)e;
        ov = DwOverlapCtx{st, side, part, ckpt + AL.s / 4, a.pbar, H, L, ov_splits, 0, rows, n_tiles, (const unsigned char*)(sc + SL.bwd_live)};
        hook = MlpDwHook{kDwGroup, &ov, dw_overlap_launch};
      }
    }
    if (bwd_tc) {
      if (chunk_idx == 0) {
        const int rcp = launch_mlp_prep_tc(stream, theta, WTc, WTlo, (float*)(sc + SL.wbi), Wclo, H, L);
        if (rcp) return rcp;
      }
      const size_t rows_pad_c = (size_t)((rows + 127) / 128 * 128);
      const size_t sb = al256(3 * rows_pad_c * H * 4);
      float* S[kMaxL];
      for (int l = 0; l < L; ++l) S[l] = (float*)(sc + SL.bwd_s + (size_t)l * sb);
      const int rc2 = launch_mlp_bwd_sweep_tc(stream, theta, (float*)(sc + SL.wbi), Wclo, WTc, WTlo, ckpt, ybar1, ybar0, S, (float*)(sc + SL.bwd_pb),
                                              (float*)(sc + SL.bwd_pb + sb), (float*)(sc + SL.bwd_bst), (float*)(sc + SL.bwd_xpart),
                                              gsmall, a.acts, a.pbar, B, r0, rows, stride, H, L, n_steps, t0, t1, y0sign,
                                              act ? ckpt + AL.s / 4 : nullptr, hook.launch != nullptr ? &hook : nullptr,
                                              act ? (unsigned char*)(sc + SL.bwd_live) : nullptr);
      if (rc2) return rc2;
      grid = (int)n_tiles;                       // small-gradient slots: one per 8-row tile (<= 256 per chunk)
    } else {
      mlp_bwd_sweep_kernel<<<grid, H / 2, smem, st>>>(a);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    mlp_small_reduce_kernel<<<(TL.n_small() + 255) / 256, 256, 0, st>>>(gsmall, theta_bar, H, L, grid, chunk_idx > 0);
    // default: tcgen05 3xTF32 weight-gradient GEMM; OFDFT_PATH_FFMA_DW: FP32-pipe split-K GEMM
    const bool dw_tc = !(path & OFDFT_PATH_FFMA_DW);
    if (dw_tc) {
      e = cudaFuncSetAttribute(mlp_dw_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DwSmem) + 1024);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(mlp_dw_gemm_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dw3Smem) + 1024);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(mlp_dw_gemm_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dw3Smem) + 1024);
      if (e != cudaSuccess) return (int)e;
    }
    if (hook.launch != nullptr) {                 // every evaluation has been contracted beside the sweep: join and reduce
      e = cudaEventRecord(ov.side->join, ov.side->s);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ov.side->join, 0);
      if (e != cudaSuccess) return (int)e;
      for (int l = 1; l < L; ++l)
        mlp_dw_reduce_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(part + (size_t)(l - 1) * ov.splits * H * H, theta_bar, TL.w(l), H * H,
                                                                   ov.splits, chunk_idx > 0);
    }
    for (int l = 1; l < L && hook.launch == nullptr; ++l) {
      if (dw_tc && H % 128 == 0) {
        dim3 gg(H / 128, H / 128, kDw3Splits);
        if (act) {      // operands straight from the per-evaluation layer buffers (activations: forward; cotangents: this sweep)
          const size_t sbuf = al256(3 * (size_t)((rows + 127) / 128 * 128) * H * 4) / 4;
          DwSrc src{ckpt + AL.s / 4 + (size_t)(l - 1) * sbuf, a.pbar + (size_t)(l - 1) * sbuf, (size_t)L * sbuf, (size_t)(L - 1) * sbuf,
                    (rows + 127) / 128 * 128, rows, (int)n_tiles, (const unsigned char*)(sc + SL.bwd_live)};
          mlp_dw_gemm_ts_kernel<true><<<gg, kDw3Threads, sizeof(Dw3Smem) + 1024, st>>>(nullptr, nullptr, part, H, L, l, 0, n_blocks, kDw3Splits, src, 0);
        } else {
          mlp_dw_gemm_ts_kernel<false><<<gg, kDw3Threads, sizeof(Dw3Smem) + 1024, st>>>(a.acts, a.pbar, part, H, L, l, 0, n_blocks, kDw3Splits, DwSrc{}, 0);
        }
        mlp_dw_reduce_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(part, theta_bar, TL.w(l), H * H, kDw3Splits, chunk_idx > 0);
      } else if (dw_tc) {
        dim3 gg(H / 64, H / 64, kDwTcSplits);
        mlp_dw_gemm_tc_kernel<<<gg, 256, sizeof(DwSmem) + 1024, st>>>(a.acts, a.pbar, part, H, L, l, n_blocks, kDwTcSplits);
        mlp_dw_reduce_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(part, theta_bar, TL.w(l), H * H, 2 * kDwTcSplits, chunk_idx > 0);
      } else {
        dim3 gg(H / kDwTile, H / kDwTile, kDwSplits);
        mlp_dw_gemm_kernel<<<gg, 256, 0, st>>>(a.acts, a.pbar, part, H, L, l, n_blocks, kDwSplits);
        mlp_dw_reduce_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(part, theta_bar, TL.w(l), H * H, kDwSplits, chunk_idx > 0);
      }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}
