// gnn_general.cu -- 3-D equivariant GNN flow with a radial MLP of 1 or 3 hidden layers of 64 units
// (`--nn 1|3` of OFDFT_NF.py:320-326 selects features (64,) / (64,64,64); equiv_flows.py:36-62 builds the stack).
//
// The (64, 64) network -- what every other driver of the reference hard-codes -- has the tensor-core tile kernels of
// gnn_tc.cu, (64, 64, 64) those of gnn_deep_tc.cu.  This file covers one hidden layer (no GEMM in that network) and, under
// OFDFT_PATH_FFMA, three with one plain FP32 kernel pair -- the second, independently written implementation the parity
// tests compare the tensor-core kernels of depth 3 with: thread = trajectory, the layer jets
// (a, a', a'') of one (trajectory, nucleus, RHS evaluation) live in the thread's local memory, weights are read through
// the read-only path with warp-uniform addresses, and the only cross-thread step is the weight-gradient update
//     dW_l[k][j] += sum over the tile's 128 trajectories and the jets of A_{l-1}[k] Pbar_l[j]
// which goes through two shared-memory row buffers (thread (j, k-half) keeps 32 accumulators per layer).  The small
// gradients are accumulated per trajectory and column-summed once per tile.  Same RK4 / discrete-adjoint scheme, stage
// checkpoint layout, tile schedule and per-tile partial gradients as the tile kernels, so results are bitwise reproducible.
//
// The kernel bodies use no warp intrinsics or inline PTX: `tests/hostsim/` compiles this file for the CPU (threads of a
// block = OS threads, __syncthreads = barrier) to check the algebra against the float64 oracle without a GPU.
#include "gnn_common.cuh"

namespace ofdft {

constexpr int kGT = kTP;               // trajectories = threads per tile
constexpr int kRow = kH + 1;           // padded row of a shared-memory row buffer (thread-private rows: conflict free)

__device__ __forceinline__ float tanh_g(float x) {
#ifdef OFDFT_HOST_SIM
  return tanhf(x);
#else
  return tanh_f(x);
#endif
}

struct SmemGenF {                    // forward: the tables only (5 KB: the CTAs per SM are not limited by shared memory)
  NucTables T;
  int tile;
};

struct SmemGen {
  NucTables T;
  float uA[kGT * kRow];
  float vP[kGT * kRow];
  float wQ[kGT * kRow];            // adjoint: last-layer kernel gradient terms of the current (nucleus, evaluation)
  int tile;
};

// jets of one (trajectory, nucleus): J[l] = (a, a', a'') of hidden layer l (derivatives with respect to r), Pd[l] = (P', P'')
// the pre-activation derivative jets of layer l >= 1 that the reverse sweep needs
template <int L>
struct LayerJets {
  float J[L][3][kH];
  float Pd[L][2][kH];
};

template <int L, int NJ>
__device__ void mlp_forward(const float* __restrict__ theta, const GnnThetaLayoutL& TL, const float* __restrict__ cbi, float r,
                            float tprime, LayerJets<L>& S, float (&f)[3]) {
  // first layer: p = w_r r + w_t t' + (b1 + onehot . W1[2:,:])
  for (int k = 0; k < kH; ++k) {
    const float wt = __ldg(theta + TL.W[0] + k), wr = __ldg(theta + TL.W[0] + kH + k);
    const float a = tanh_g(fmaf(wr, r, fmaf(tprime, wt, cbi[k])));
    const float a1 = fmaf(-a, a, 1.0f) * wr;
    S.J[0][0][k] = a;
    if (NJ >= 2) S.J[0][1][k] = a1;
    if (NJ == 3) S.J[0][2][k] = -2.0f * a * a1 * wr;
  }
  for (int l = 1; l < L; ++l) {
    const float* __restrict__ W = theta + TL.W[l];
    for (int j0 = 0; j0 < kH; j0 += 8) {
      float acc[3][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { acc[0][q] = __ldg(theta + TL.b[l] + j0 + q); acc[1][q] = 0.f; acc[2][q] = 0.f; }
      for (int k = 0; k < kH; ++k) {
        const float x0 = S.J[l - 1][0][k], x1 = NJ >= 2 ? S.J[l - 1][1][k] : 0.f, x2 = NJ == 3 ? S.J[l - 1][2][k] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float w = __ldg(W + k * kH + j0 + q);
          acc[0][q] = fmaf(x0, w, acc[0][q]);
          if (NJ >= 2) acc[1][q] = fmaf(x1, w, acc[1][q]);
          if (NJ == 3) acc[2][q] = fmaf(x2, w, acc[2][q]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        const float a = tanh_g(acc[0][q]);
        const float sg = fmaf(-a, a, 1.0f);
        S.J[l][0][j] = a;
        if (NJ >= 2) {
          const float a1 = sg * acc[1][q];
          S.J[l][1][j] = a1; S.Pd[l][0][j] = acc[1][q];
          if (NJ == 3) { S.J[l][2][j] = fmaf(sg, acc[2][q], -2.0f * a * a1 * acc[1][q]); S.Pd[l][1][j] = acc[2][q]; }
        }
      }
    }
  }
  f[0] = __ldg(theta + TL.b3); f[1] = 0.f; f[2] = 0.f;
  for (int j = 0; j < kH; ++j) {
    const float w = __ldg(theta + TL.w3 + j);
    f[0] = fmaf(w, S.J[L - 1][0][j], f[0]);
    if (NJ >= 2) f[1] = fmaf(w, S.J[L - 1][1][j], f[1]);
    if (NJ == 3) f[2] = fmaf(w, S.J[L - 1][2][j], f[2]);
  }
}

// one hidden layer: f = b3 + w3 . tanh-jet(first layer) unit by unit -- nothing is stored (the layer arrays of mlp_forward live
// in local memory, and with no GEMM to feed that traffic is all the kernel would do)
template <int NJ>
__device__ void mlp_forward_1(const float* __restrict__ theta, const GnnThetaLayoutL& TL, const float* __restrict__ cbi, float r,
                              float tprime, float (&f)[3]) {
  f[0] = __ldg(theta + TL.b3); f[1] = 0.f; f[2] = 0.f;
#pragma unroll 8
  for (int k = 0; k < kH; ++k) {          // unrolled: eight independent tanh chains in flight per thread
    const float wt = __ldg(theta + TL.W[0] + k), wr = __ldg(theta + TL.W[0] + kH + k), w = __ldg(theta + TL.w3 + k);
    const float a = tanh_g(fmaf(wr, r, fmaf(tprime, wt, cbi[k])));
    const float a1 = fmaf(-a, a, 1.0f) * wr;
    f[0] = fmaf(w, a, f[0]);
    if (NJ >= 2) f[1] = fmaf(w, a1, f[1]);
    if (NJ == 3) f[2] = fmaf(w, -2.0f * a * a1 * wr, f[2]);
  }
}

// explicit (geometry) part of one nucleus' contribution to d/dt (x, logp, s)   [same formulas as combine() in gnn.cu]
template <int NJ>
__device__ __forceinline__ void combine_nucleus(const float (&f)[3], const float* nc, const float (&xst)[3], const float (&sst)[3],
                                                float (&g)[3], float& dv, float (&ds)[3]) {
  const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
  g[0] = fmaf(f[0], d0, g[0]); g[1] = fmaf(f[0], d1, g[1]); g[2] = fmaf(f[0], d2, g[2]);
  if (NJ >= 2) {
    const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
    dv += fmaf(f[1], r, 3.0f * f[0]);
    if (NJ == 3) {
      const float q = fmaf(d0, sst[0], fmaf(d1, sst[1], d2 * sst[2]));
      const float amp = (fmaf(f[1], q + 4.0f, f[2] * r)) / r;
      ds[0] += fmaf(f[0], sst[0], amp * d0);
      ds[1] += fmaf(f[0], sst[1], amp * d1);
      ds[2] += fmaf(f[0], sst[2], amp * d2);
    }
  }
}

template <int L, int NJ>
__device__ void gen_fwd_tile(SmemGenF& sm, const GnnFwdArgs& a, const GnnThetaLayoutL& TL, long long row0, long long seg_end) {
  const int tid = threadIdx.x;
  const long long row = row0 + tid;
  const bool valid = row < seg_end;
  const long long rowc = valid ? row : seg_end - 1;
  float y[7] = {0, 0, 0, 0, 0, 0, 0}, accw[7] = {0, 0, 0, 0, 0, 0, 0}, kst[7] = {0, 0, 0, 0, 0, 0, 0};
  {
    const float* src = a.y_in + rowc * a.in_stride;
    y[0] = src[0]; y[1] = src[1]; y[2] = src[2];
    if (NJ >= 2) y[3] = src[3];
    if (NJ == 3) { y[4] = src[4]; y[5] = src[5]; y[6] = src[6]; }
  }
  const int n_steps = a.rhs_only ? 1 : a.n_steps;
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  for (int step = 0; step < n_steps; ++step) {
    for (int stage = 0; stage < (a.rhs_only ? 1 : 4); ++stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float ts = a.rhs_only ? a.t_rhs : a.t0 + h * (float)step + cs;
      const float tprime = a.y0 * ts;
      float xst[3], sst[3] = {0, 0, 0};
      for (int c = 0; c < 3; ++c) {
        xst[c] = fmaf(cs, kst[c], y[c]);
        if (NJ == 3) sst[c] = fmaf(cs, kst[4 + c], y[4 + c]);
      }
      if (a.ckpt != nullptr && valid) {
        float* ck = a.ckpt + ((long long)(step * 4 + stage) * 6) * a.B + row;
        for (int c = 0; c < 3; ++c) {
          ck[(long long)c * a.B] = xst[c];
          if (NJ == 3) ck[(long long)(3 + c) * a.B] = sst[c];
        }
      }
      float g[3] = {0, 0, 0}, dv = 0.f, ds[3] = {0, 0, 0};
      for (int i = 0; i < a.Na; ++i) {
        const float* nc = sm.T.nuc + i * 4;
        const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
        const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
        float f[3];
        if constexpr (L == 1) {
          mlp_forward_1<NJ>(a.theta, TL, sm.T.cb + sm.T.cls[i] * kH, r, tprime, f);
        } else {
          LayerJets<L> S;
          mlp_forward<L, NJ>(a.theta, TL, sm.T.cb + sm.T.cls[i] * kH, r, tprime, S, f);
        }
        combine_nucleus<NJ>(f, nc, xst, sst, g, dv, ds);
      }
      const float wgt = (stage == 0 || stage == 3) ? 1.0f : 2.0f;
      kst[0] = a.y0 * g[0]; kst[1] = a.y0 * g[1]; kst[2] = a.y0 * g[2];
      if (NJ >= 2) kst[3] = -a.y0 * dv;
      if (NJ == 3) { kst[4] = -a.y0 * ds[0]; kst[5] = -a.y0 * ds[1]; kst[6] = -a.y0 * ds[2]; }
      for (int c = 0; c < 7; ++c) accw[c] = fmaf(wgt, kst[c], accw[c]);
    }
    if (!a.rhs_only)
      for (int c = 0; c < 7; ++c) { y[c] = fmaf(h * (1.0f / 6.0f), accw[c], y[c]); accw[c] = 0.f; }
  }
  if (valid) {
    float* dst = a.y_out + row * a.out_stride;
    const float* o = a.rhs_only ? kst : y;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    if (NJ >= 2) dst[3] = o[3];
    if (NJ == 3) { dst[4] = o[4]; dst[5] = o[5]; dst[6] = o[6]; }
  }
}

template <int L>
__global__ void __launch_bounds__(kGT) gnn_general_fwd_kernel(const GnnFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemGenF& sm = *reinterpret_cast<SmemGenF*>(smem_raw);
  const GnnThetaLayoutL TL(a.n_types, L);
  const GnnThetaLayout L2(a.n_types);       // b1 / W1 offsets do not depend on the depth
  load_nuc_tables<kGT>(sm.T, a.theta, L2, a.nuclei, a.onehot, a.Na, a.n_types);
  const long long tiles_a = (a.n_a + kGT - 1) / kGT;
  const long long tiles_b = (a.B - a.n_a + kGT - 1) / kGT;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    if (tile < tiles_a) { row0 = tile * kGT; seg_end = a.n_a; nj = a.jets_a; }
    else { row0 = a.n_a + (tile - tiles_a) * kGT; seg_end = a.B; nj = a.jets_b; }
    if (nj == 3) gen_fwd_tile<L, 3>(sm, a, TL, row0, seg_end);
    else if (nj == 2) gen_fwd_tile<L, 2>(sm, a, TL, row0, seg_end);
    else gen_fwd_tile<L, 1>(sm, a, TL, row0, seg_end);
  }
}

// ----------------------------------------------------------------------------------------------------------------------
// discrete adjoint
// ----------------------------------------------------------------------------------------------------------------------
// small-gradient accumulators of thread (j = tid & 63, row half = tid >> 6): column j, summed over the thread's 64 rows of
// the tile.  Every (nucleus, evaluation) the 128 trajectories put their terms into shared-memory row buffers and each
// thread adds up its column half -- per-trajectory 64-vectors accumulated in local memory (the first version) made the
// adjoint L2-bound: 3 KB of read-modify-write per thread and evaluation.
template <int L>
struct SmallAcc {
  float w3;                        // d w3[j]
  float bl[L];                     // d b_l[j], l >= 1 (slot 0 unused)
  float wt, wr;                    // d W1[0][j], d W1[1][j]
  float S[kMaxClasses];            // per nucleus class: sum of first-layer pre-activation cotangents
  float b3;                        // per trajectory (thread = row)
  __device__ void zero() {
    w3 = 0.f; wt = 0.f; wr = 0.f; b3 = 0.f;
    for (int l = 0; l < L; ++l) bl[l] = 0.f;
    for (int c = 0; c < kMaxClasses; ++c) S[c] = 0.f;
  }
};

// weight-gradient accumulators of thread (j = tid & 63, k-half = tid >> 6): dW_l[32 kh + kk][j]
template <int L>
struct DwAcc { float v[L][32]; };

template <int L, int NJ>
__device__ void gen_bwd_tile(SmemGen& sm, const GnnBwdArgs& a, const GnnThetaLayoutL& TL, long long row0, long long seg_end,
                             SmallAcc<L>& G, DwAcc<L>& D) {
  const int tid = threadIdx.x;
  const long long row = row0 + tid;
  const bool valid = row < seg_end;
  const long long rowc = valid ? row : seg_end - 1;
  const int jcol = tid & 63, kh = tid >> 6;
  float xb[3] = {0, 0, 0}, lb = 0.f, sb[3] = {0, 0, 0};
  if (valid) {
    const float* src = a.ybar1 + row * a.bar_stride;
    xb[0] = src[0]; xb[1] = src[1]; xb[2] = src[2];
    if (NJ >= 2) lb = src[3];
    if (NJ == 3) { sb[0] = src[4]; sb[1] = src[5]; sb[2] = src[6]; }
  }
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  float prevx[3] = {0, 0, 0}, prevs[3] = {0, 0, 0};
  for (int step = a.n_steps - 1; step >= 0; --step) {
    float sumx[3] = {0, 0, 0}, sums[3] = {0, 0, 0};
    for (int stage = 3; stage >= 0; --stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float tprime = a.y0 * (a.t0 + h * (float)step + cs);
      // k4bar = h/6 ybar ; k3bar = h/3 ybar + h y4bar ; k2bar = h/3 ybar + h/2 y3bar ; k1bar = h/6 ybar + h/2 y2bar
      const float ca = (stage == 0 || stage == 3) ? h * (1.0f / 6.0f) : h * (1.0f / 3.0f);
      const float cp = stage == 3 ? 0.f : (stage == 2 ? h : 0.5f * h);
      float kx[3], kl = 0.f, ks[3] = {0, 0, 0}, xst[3], sst[3] = {0, 0, 0};
      const float* ck = a.ckpt + ((long long)(step * 4 + stage) * 6) * a.B + rowc;
      for (int c = 0; c < 3; ++c) {
        kx[c] = fmaf(ca, xb[c], cp * prevx[c]);
        if (NJ == 3) ks[c] = fmaf(ca, sb[c], cp * prevs[c]);
        xst[c] = ck[(long long)c * a.B];
        if (NJ == 3) sst[c] = ck[(long long)(3 + c) * a.B];
      }
      if (NJ >= 2) kl = ca * lb;
      float ysx[3] = {0, 0, 0}, yss[3] = {0, 0, 0};
      for (int i = 0; i < a.Na; ++i) {
        const float* nc = sm.T.nuc + i * 4;
        const int cls = sm.T.cls[i];
        const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
        const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
        const float rinv = 1.0f / r;
        float f[3];
        // cotangents of (f, f', f'')   [pre() of gnn.cu]
        float fb[3] = {0.f, 0.f, 0.f};
        float q = 0.f, gam = 0.f;
        {
          const float alpha = fmaf(d0, kx[0], fmaf(d1, kx[1], d2 * kx[2]));
          fb[0] = alpha;
          if (NJ >= 2) {
            fb[0] = fmaf(-3.0f, kl, fb[0]);
            fb[1] = -r * kl;
            if (NJ == 3) {
              q = fmaf(d0, sst[0], fmaf(d1, sst[1], d2 * sst[2]));
              const float beta = fmaf(ks[0], sst[0], fmaf(ks[1], sst[1], ks[2] * sst[2]));
              gam = fmaf(d0, ks[0], fmaf(d1, ks[1], d2 * ks[2])) * rinv;
              fb[0] -= beta;
              fb[1] = fmaf(-(q + 4.0f), gam, fb[1]);
              fb[2] = -r * gam;
            }
          }
          fb[0] *= a.y0; fb[1] *= a.y0; fb[2] *= a.y0;
          if (!valid) { fb[0] = 0.f; fb[1] = 0.f; fb[2] = 0.f; }
        }
        G.b3 += fb[0];
        __syncthreads();                                       // the previous evaluation's column sums have been read
        float rbar = 0.f;
        if constexpr (L == 1) {
          // one hidden layer: forward jet, last-layer gradient term and first-layer reverse unit by unit, nothing stored per thread
          const float* cbi = sm.T.cb + cls * kH;
          f[0] = __ldg(a.theta + TL.b3); f[1] = 0.f; f[2] = 0.f;
#pragma unroll 8
          for (int k = 0; k < kH; ++k) {
            const float wt = __ldg(a.theta + TL.W[0] + k), wv = __ldg(a.theta + TL.W[0] + kH + k), w = __ldg(a.theta + TL.w3 + k);
            const float aa = tanh_g(fmaf(wv, r, fmaf(tprime, wt, cbi[k])));
            const float sg = fmaf(-aa, aa, 1.0f);
            const float a1 = sg * wv;
            const float a2 = -2.0f * aa * a1 * wv;
            f[0] = fmaf(w, aa, f[0]);
            float gw = aa * fb[0];
            float t = w * fb[0], e = 0.f;
            if (NJ >= 2) {
              f[1] = fmaf(w, a1, f[1]);
              gw = fmaf(a1, fb[1], gw);
              const float c1 = w * fb[1];
              t = fmaf(-2.0f * aa * wv, c1, t);
              e = c1;
              if (NJ == 3) {
                f[2] = fmaf(w, a2, f[2]);
                gw = fmaf(a2, fb[2], gw);
                const float c2 = w * fb[2];
                t = fmaf(-2.0f * wv * wv * fmaf(-3.0f * aa, aa, 1.0f), c2, t);
                e = fmaf(-4.0f * aa * wv, c2, e);
              }
            }
            const float pb = sg * t;
            rbar = fmaf(pb, wv, rbar);
            sm.wQ[tid * kRow + k] = gw;
            sm.uA[tid * kRow + k] = pb;
            sm.vP[tid * kRow + k] = fmaf(pb, r, sg * e);
          }
        } else {
          LayerJets<L> S;
          float cot[3][kH], nxt[3][kH];
          mlp_forward<L, NJ>(a.theta, TL, sm.T.cb + cls * kH, r, tprime, S, f);
          // last layer: f = b3 + w3 . a_{L-1}
          for (int j = 0; j < kH; ++j) {
            const float w = __ldg(a.theta + TL.w3 + j);
            float gw = S.J[L - 1][0][j] * fb[0];
            cot[0][j] = w * fb[0];
            if (NJ >= 2) { gw = fmaf(S.J[L - 1][1][j], fb[1], gw); cot[1][j] = w * fb[1]; }
            if (NJ == 3) { gw = fmaf(S.J[L - 1][2][j], fb[2], gw); cot[2][j] = w * fb[2]; }
            sm.wQ[tid * kRow + j] = gw;
          }
          // hidden layers L-1 .. 1: reverse of the tanh jet, weight gradient, transposed mat-vec
          for (int l = L - 1; l >= 1; --l) {
            for (int j = 0; j < kH; ++j) {
              const float aa = S.J[l][0][j];
              const float sg = fmaf(-aa, aa, 1.0f);
              float pzb = cot[0][j];
              if (NJ >= 2) {
                const float p1 = S.Pd[l][0][j];
                const float a1b = cot[1][j];
                pzb = fmaf(-2.0f * aa * p1, a1b, pzb);
                float t1 = a1b;
                if (NJ == 3) {
                  const float p2 = S.Pd[l][1][j];
                  const float a2b = cot[2][j];
                  cot[2][j] = a2b * sg;
                  t1 = fmaf(-4.0f * aa * p1, a2b, t1);
                  pzb = fmaf(a2b, fmaf(-2.0f * aa, p2, -2.0f * fmaf(-3.0f * aa, aa, 1.0f) * p1 * p1), pzb);
                }
                cot[1][j] = sg * t1;
              }
              cot[0][j] = sg * pzb;
            }
            // dW_l[k][j] += sum over the tile's trajectories and jets of J[l-1][jet][k] * cot[jet][j]
            for (int jt = 0; jt < NJ; ++jt) {
              __syncthreads();
              for (int u = 0; u < kH; ++u) { sm.uA[tid * kRow + u] = S.J[l - 1][jt][u]; sm.vP[tid * kRow + u] = cot[jt][u]; }
              __syncthreads();
              float bsum = 0.f;
              for (int t = 0; t < kGT; ++t) {
                const float v = sm.vP[t * kRow + jcol];
                const float* ua = sm.uA + t * kRow + 32 * kh;
#pragma unroll
                for (int kk = 0; kk < 32; ++kk) D.v[l][kk] = fmaf(ua[kk], v, D.v[l][kk]);
                if ((t >> 6) == kh) bsum += v;                   // jet 0 of the cotangent is the bias gradient of layer l
              }
              if (jt == 0) G.bl[l] += bsum;
            }
            // cotangent jets of layer l-1: nxt[jet][k] = sum_j cot[jet][j] W_l[k][j]
            const float* __restrict__ W = a.theta + TL.W[l];
            for (int k0 = 0; k0 < kH; k0 += 8) {
              float acc[3][8];
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) { acc[0][qq] = 0.f; acc[1][qq] = 0.f; acc[2][qq] = 0.f; }
              for (int j = 0; j < kH; ++j) {
                const float c0 = cot[0][j], c1 = NJ >= 2 ? cot[1][j] : 0.f, c2 = NJ == 3 ? cot[2][j] : 0.f;
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                  const float w = __ldg(W + (k0 + qq) * kH + j);
                  acc[0][qq] = fmaf(c0, w, acc[0][qq]);
                  if (NJ >= 2) acc[1][qq] = fmaf(c1, w, acc[1][qq]);
                  if (NJ == 3) acc[2][qq] = fmaf(c2, w, acc[2][qq]);
                }
              }
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) { nxt[0][k0 + qq] = acc[0][qq]; nxt[1][k0 + qq] = acc[1][qq]; nxt[2][k0 + qq] = acc[2][qq]; }
            }
            for (int jt = 0; jt < NJ; ++jt)
              for (int k = 0; k < kH; ++k) cot[jt][k] = nxt[jt][k];
          }
          // first layer reverse   [epilogue_b1 of gnn.cu]
          rbar = 0.f;
          if (L > 1) __syncthreads();                            // the weight-gradient pass above has read uA / vP
          for (int k = 0; k < kH; ++k) {
            const float aa = S.J[0][0][k];
            const float wv = __ldg(a.theta + TL.W[0] + kH + k);
            const float sg = fmaf(-aa, aa, 1.0f);
            float t = cot[0][k], e = 0.f;
            if (NJ >= 2) {
              t = fmaf(-2.0f * aa * wv, cot[1][k], t);
              e = cot[1][k];
              if (NJ == 3) {
                t = fmaf(-2.0f * wv * wv * fmaf(-3.0f * aa, aa, 1.0f), cot[2][k], t);
                e = fmaf(-4.0f * aa * wv, cot[2][k], e);
              }
            }
            const float pb = sg * t;
            rbar = fmaf(pb, wv, rbar);
            sm.uA[tid * kRow + k] = pb;
            sm.vP[tid * kRow + k] = fmaf(pb, r, sg * e);
          }
        }
        // column sums of this (nucleus, evaluation) over the thread's half of the tile rows
        __syncthreads();
        {
          float s_pb = 0.f, s_wr = 0.f, s_w3 = 0.f;
          for (int t = 64 * kh; t < 64 * kh + 64; ++t) {
            s_pb += sm.uA[t * kRow + jcol]; s_wr += sm.vP[t * kRow + jcol]; s_w3 += sm.wQ[t * kRow + jcol];
          }
          G.S[cls] += s_pb;
          G.wt = fmaf(tprime, s_pb, G.wt);                     // t' is the same for every trajectory of the evaluation
          G.wr += s_wr;
          G.w3 += s_w3;
        }
        // explicit geometry   [post() of gnn.cu]
        {
          const float u0 = d0 * rinv, u1 = d1 * rinv, u2 = d2 * rinv;
          float db0 = a.y0 * f[0] * kx[0], db1 = a.y0 * f[0] * kx[1], db2 = a.y0 * f[0] * kx[2];
          if (NJ >= 2) rbar = fmaf(-a.y0 * f[1], kl, rbar);
          if (NJ == 3) {
            const float gamu = fmaf(u0, ks[0], fmaf(u1, ks[1], u2 * ks[2]));
            const float amp = fmaf(f[1], q + 4.0f, f[2] * r);
            rbar = fmaf(-a.y0 * f[2], gamu, rbar);
            const float qbar = -a.y0 * f[1] * gamu;
            const float ua = -a.y0 * amp;
            db0 = fmaf(qbar, sst[0], fmaf(ua * rinv, ks[0], db0));
            db1 = fmaf(qbar, sst[1], fmaf(ua * rinv, ks[1], db1));
            db2 = fmaf(qbar, sst[2], fmaf(ua * rinv, ks[2], db2));
            rbar = fmaf(-ua * gamu, rinv, rbar);
            yss[0] += fmaf(qbar, d0, -a.y0 * f[0] * ks[0]);
            yss[1] += fmaf(qbar, d1, -a.y0 * f[0] * ks[1]);
            yss[2] += fmaf(qbar, d2, -a.y0 * f[0] * ks[2]);
          }
          if (valid) {
            ysx[0] += fmaf(rbar, u0, db0);
            ysx[1] += fmaf(rbar, u1, db1);
            ysx[2] += fmaf(rbar, u2, db2);
          } else {
            yss[0] = yss[1] = yss[2] = 0.f;
          }
        }
      }
      for (int c = 0; c < 3; ++c) {
        sumx[c] += ysx[c]; prevx[c] = ysx[c];
        sums[c] += yss[c]; prevs[c] = yss[c];
      }
    }
    for (int c = 0; c < 3; ++c) { xb[c] += sumx[c]; sb[c] += sums[c]; }
  }
  if (valid && a.ybar0 != nullptr) {
    float* dst = a.ybar0 + row * a.bar_stride;
    dst[0] = xb[0]; dst[1] = xb[1]; dst[2] = xb[2];
    if (NJ >= 2) dst[3] = lb;
    if (NJ == 3) { dst[4] = sb[0]; dst[5] = sb[1]; dst[6] = sb[2]; }
  }
}

// both row halves of column j (threads j and 64 + j) -> the column sum over the whole tile
__device__ __forceinline__ float tile_halves_sum(SmemGen& sm, float v) {
  const int tid = threadIdx.x, jcol = tid & 63;
  __syncthreads();
  sm.vP[tid] = v;
  __syncthreads();
  return sm.vP[jcol] + sm.vP[64 + jcol];
}

template <int L>
__device__ void gen_flush_tile(SmemGen& sm, SmallAcc<L>& G, DwAcc<L>& D, float* __restrict__ out, const GnnThetaLayoutL& TL,
                               const float* __restrict__ onehot, int n_types) {
  const int tid = threadIdx.x, jcol = tid & 63, kh = tid >> 6;
  for (int l = 1; l < L; ++l)
    for (int kk = 0; kk < 32; ++kk) { out[TL.W[l] + (32 * kh + kk) * kH + jcol] = D.v[l][kk]; D.v[l][kk] = 0.f; }
  float s;
  s = tile_halves_sum(sm, G.w3); if (kh == 0) out[TL.w3 + jcol] = s;
  for (int l = 1; l < L; ++l) { s = tile_halves_sum(sm, G.bl[l]); if (kh == 0) out[TL.b[l] + jcol] = s; }
  s = tile_halves_sum(sm, G.wt); if (kh == 0) out[TL.W[0] + jcol] = s;
  s = tile_halves_sum(sm, G.wr); if (kh == 0) out[TL.W[0] + kH + jcol] = s;
  float sb1 = 0.f, st[kMaxTypes];
  for (int t = 0; t < kMaxTypes; ++t) st[t] = 0.f;
  for (int c = 0; c < sm.T.n_cls; ++c) {
    s = tile_halves_sum(sm, G.S[c]);
    sb1 += s;
    for (int t = 0; t < n_types; ++t) st[t] = fmaf(onehot[sm.T.rep[c] * n_types + t], s, st[t]);
  }
  if (kh == 0) {
    out[TL.b[0] + jcol] = sb1;
    for (int t = 0; t < n_types; ++t) out[TL.W[0] + (2 + t) * kH + jcol] = st[t];
  }
  // b3: sum of the per-trajectory scalars in fixed order
  __syncthreads();
  sm.uA[tid] = G.b3;
  __syncthreads();
  if (tid == 0) {
    float b = 0.f;
    for (int t = 0; t < kGT; ++t) b += sm.uA[t];
    out[TL.b3] = b;
  }
  __syncthreads();
  G.zero();
}

template <int L>
__global__ void __launch_bounds__(kGT) gnn_general_bwd_kernel(const GnnBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemGen& sm = *reinterpret_cast<SmemGen*>(smem_raw);
  const GnnThetaLayoutL TL(a.n_types, L);
  const GnnThetaLayout L2(a.n_types);
  load_nuc_tables<kGT>(sm.T, a.theta, L2, a.nuclei, a.onehot, a.Na, a.n_types);
  SmallAcc<L> G;
  DwAcc<L> D;
  G.zero();
  for (int l = 0; l < L; ++l)
    for (int kk = 0; kk < 32; ++kk) D.v[l][kk] = 0.f;
  const long long tiles_a = (a.n_a + kGT - 1) / kGT;
  const long long tiles_b = (a.B - a.n_a + kGT - 1) / kGT;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    if (tile < tiles_a) { row0 = tile * kGT; seg_end = a.n_a; nj = a.jets_a; }
    else { row0 = a.n_a + (tile - tiles_a) * kGT; seg_end = a.B; nj = a.jets_b; }
    if (nj == 3) gen_bwd_tile<L, 3>(sm, a, TL, row0, seg_end, G, D);
    else if (nj == 2) gen_bwd_tile<L, 2>(sm, a, TL, row0, seg_end, G, D);
    else gen_bwd_tile<L, 1>(sm, a, TL, row0, seg_end, G, D);
    gen_flush_tile<L>(sm, G, D, a.gpart + (size_t)tile * TL.n, TL, a.onehot, a.n_types);
  }
}

int gnn_general_theta_size(int n_types, int n_layers) { return GnnThetaLayoutL(n_types, n_layers).n; }

#ifndef OFDFT_HOST_SIM
int launch_gnn_general_fwd(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes, int n_layers) {
  if (scratch_bytes < kCounterBytes) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  a.counter = (int*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (a.n_a + kGT - 1) / kGT + (a.B - a.n_a + kGT - 1) / kGT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < 8LL * sms ? tiles : 8LL * sms);
  auto kern = n_layers == 1 ? gnn_general_fwd_kernel<1> : gnn_general_fwd_kernel<3>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemGenF));
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kGT, sizeof(SmemGenF), st>>>(a);
  return (int)cudaGetLastError();
}

// the caller has reset the tile counter and reduces gpart afterwards
int launch_gnn_general_bwd(void* stream, GnnBwdArgs& a, int n_layers) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long tiles = (a.n_a + kGT - 1) / kGT + (a.B - a.n_a + kGT - 1) / kGT;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);
  auto kern = n_layers == 1 ? gnn_general_bwd_kernel<1> : gnn_general_bwd_kernel<3>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemGen));
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kGT, sizeof(SmemGen), st>>>(a);
  return (int)cudaGetLastError();
}
#endif

}  // namespace ofdft
