// gnn.cu -- fused fixed-step (RK4) integrator and discrete adjoint for the 3-D permutation-equivariant
// GNN flow  g(x,t) = sum_i f_theta(t', |x-R_i|, onehot_i) (x - R_i)   (reference: ofdft_normflows/
// equiv_flows.py:36-127 evaluated through ofdft_normflows/jax_ode.py:9-110).  sm_100a, FP32 FFMA pipe.
//
// Work decomposition (one CTA = 256 threads = one tile of 128 trajectories, persistent over tiles):
//   * the radial MLP is evaluated as a 2-jet (value, d/dr, d2/dr2) -- SURVEY.md 3.4 -- so one RHS evaluation
//     per (trajectory, nucleus) is three 64x64 mat-vecs;  for a tile and one nucleus that is a
//     [128 traj x 3 jets] x [64 x 64] GEMM executed from shared memory with a 2x16x3 register tile per
//     thread (96 FFMA per 2 LDS.64/3 + 4 LDS.128);
//   * warp w = (trajectory half = w>>2, unit group = w&3), lane = trajectory pair; thread tid < 128 "owns"
//     trajectory tid: RK4 state, per-nucleus combination and the adjoint bookkeeping live in its registers;
//   * backward = discrete adjoint of RK4 on the stage states checkpointed by the forward pass:
//     recompute GEMM, transposed GEMM (W2^T) and the rank-(3*128) update of dW2 kept in a 4x4 register tile
//     per thread for the whole kernel; small-parameter gradients are reduced with warp butterflies.
#include <cstdlib>

#include "gnn_common.cuh"

namespace ofdft {

struct SmemFwd {
  float W2[kH * kH];
  float A[3 * kJetStride];           // layer-1 activation jets a, a', a''
  float wr[kH], wt[kH], b2[kH], w3[kH];
  NucTables T;                       // nuclei, per-class b1 + onehot_i . W1[2:,:]
  float xs[3 * kTP];                 // stage positions of the tile
  float Fpart[4 * 3 * kTP];          // per unit-group partial f, f', f''
  int tile;
};

struct SmemBwd {
  float W2[kH * kH];
  float W2T[kH * kH];
  float A[2 * kJetStride];           // a, a'   (a'' = -2 w_r a a' is recomputed)
  float PB[3 * kJetStride];          // cotangents of the layer-2 pre-activation jets
  float wr[kH], wt[kH], b2[kH], w3[kH], wr2[kH];
  NucTables T;
  float S[2 * kMaxClasses * kH];     // per (half, nucleus class): sum of layer-1 pre-activation cotangents
  float xs[3 * kTP];
  float fbar[3 * kTP];
  float Fpart[4 * 3 * kTP];
  float Rpart[4 * kTP];
  float red[2 * 4 * 32 * 2];
  int tile;
};

// ------------------------------------------------------------------------------------------------
// phase 1: first layer 2-jet for (2 trajectories x 16 units) per thread -> shared activation tile
// ------------------------------------------------------------------------------------------------
template <int NJ, bool STORE_J2>
__device__ __forceinline__ void first_layer(float* __restrict__ A, const float* __restrict__ xs,
                                            const float* __restrict__ wr, const float* __restrict__ wt,
                                            const float* __restrict__ cbi, const float* __restrict__ nuci,
                                            float tprime, int half, int ug, int lane, float (&r)[2]) {
  const int tp0 = half * 64 + 2 * lane;
  const float2 x0 = *reinterpret_cast<const float2*>(xs + tp0);
  const float2 x1 = *reinterpret_cast<const float2*>(xs + kTP + tp0);
  const float2 x2 = *reinterpret_cast<const float2*>(xs + 2 * kTP + tp0);
  const float R0 = nuci[0], R1 = nuci[1], R2 = nuci[2];
  {
    float d0 = x0.x - R0, d1 = x1.x - R1, d2 = x2.x - R2;
    r[0] = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
    d0 = x0.y - R0; d1 = x1.y - R1; d2 = x2.y - R2;
    r[1] = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
  }
#pragma unroll
  for (int kk = 0; kk < 16; kk += 4) {
    const int kb = ug * 16 + kk;
    const float4 w4 = *reinterpret_cast<const float4*>(wr + kb);
    const float4 t4 = *reinterpret_cast<const float4*>(wt + kb);
    const float4 c4 = *reinterpret_cast<const float4*>(cbi + kb);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
    const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
    const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kb + e;
      const float c = fmaf(tprime, tv[e], cv[e]);
      float a[2], a1[2], a2[2];
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        const float p = fmaf(wv[e], r[tr], c);
        a[tr] = tanh_f(p);
        const float sg = fmaf(-a[tr], a[tr], 1.0f);
        a1[tr] = sg * wv[e];
        a2[tr] = -2.0f * a[tr] * a1[tr] * wv[e];
      }
      const int off = act_off(k, tp0);
      *reinterpret_cast<float2*>(A + off) = make_float2(a[0], a[1]);
      if (NJ >= 2) *reinterpret_cast<float2*>(A + kJetStride + off) = make_float2(a1[0], a1[1]);
      if (NJ == 3 && STORE_J2) *reinterpret_cast<float2*>(A + 2 * kJetStride + off) = make_float2(a2[0], a2[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 64-deep GEMM from shared memory: acc[jet][tr][j] = sum_k A[jet][k][pair] * W[k][16*ug + j]
// ------------------------------------------------------------------------------------------------
template <int NJ, bool J2_SMEM>
__device__ __forceinline__ void gemm64(float (&acc)[NJ][2][16], const float* __restrict__ A,
                                       const float* __restrict__ W, const float* __restrict__ wr2,
                                       int half, int ug, int lane) {
#pragma unroll
  for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
    for (int tr = 0; tr < 2; ++tr)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[jt][tr][j] = 0.0f;
  const int chunk = half * 16 + (lane >> 1);
  const int sub = (lane & 1) * 2;
#pragma unroll 1
  for (int k4 = 0; k4 < kH; k4 += 4) {
    const int base = (((chunk ^ ((k4 >> 2) & 7)) << 2) | sub);
    float4 c2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NJ == 3 && !J2_SMEM) c2 = *reinterpret_cast<const float4*>(wr2 + k4);
    const float c2v[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k4 + e;
      const int off = k * kTP + base;
      float2 av[3];
      av[0] = *reinterpret_cast<const float2*>(A + off);
      if (NJ >= 2) av[1] = *reinterpret_cast<const float2*>(A + kJetStride + off);
      if (NJ == 3) {
        if (J2_SMEM) av[2] = *reinterpret_cast<const float2*>(A + 2 * kJetStride + off);
        else av[2] = make_float2(av[0].x * av[1].x * c2v[e], av[0].y * av[1].y * c2v[e]);
      }
      const float4* wrow = reinterpret_cast<const float4*>(W + k * kH + ug * 16);
      const float4 w0 = wrow[0], w1 = wrow[1], w2 = wrow[2], w3 = wrow[3];
      const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                            w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          acc[jt][0][j] = fmaf(av[jt].x, wv[j], acc[jt][0][j]);
          acc[jt][1][j] = fmaf(av[jt].y, wv[j], acc[jt][1][j]);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward epilogue: layer-2 tanh jet + last-layer partial dot -> Fpart[ug][jet][traj]
// ------------------------------------------------------------------------------------------------
template <int NJ>
__device__ __forceinline__ void epilogue_fwd(const float (&acc)[NJ][2][16], const float* __restrict__ b2,
                                             const float* __restrict__ w3, float* __restrict__ Fpart,
                                             int half, int ug, int lane, float* __restrict__ act, long long act_rows) {
  float fp[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  const long long jet_stride = (long long)kH * act_rows;
  // act points at (unit group 0, this tile's first row) of jet 0; this thread owns groups 4 ug .. 4 ug + 3 of two rows
  if (act != nullptr) act += act_group_off(ug * 4, half * 64 + 2 * lane, act_rows);
  float a0[2][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float b = b2[ug * 16 + j], w = w3[ug * 16 + j];
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
      const float a = tanh_f(acc[0][tr][j] + b);
      a0[tr][j & 3] = a;
      fp[0][tr] = fmaf(w, a, fp[0][tr]);
      if (NJ >= 2) {
        const float sg = fmaf(-a, a, 1.0f);
        const float p1 = acc[1][tr][j];
        const float a1 = sg * p1;
        fp[1][tr] = fmaf(w, a1, fp[1][tr]);
        if (NJ == 3) {
          const float a2 = fmaf(sg, acc[2][tr][j], -2.0f * a * a1 * p1);
          fp[2][tr] = fmaf(w, a2, fp[2][tr]);
        }
      }
    }
    if (act != nullptr && (j & 3) == 3) {
#pragma unroll
      for (int tr = 0; tr < 2; ++tr) {
        float* dst = act + tr * 4;
        __stcs(reinterpret_cast<float4*>(dst), make_float4(a0[tr][0], a0[tr][1], a0[tr][2], a0[tr][3]));
        if (NJ >= 2) __stcs(reinterpret_cast<float4*>(dst + jet_stride), make_float4(acc[1][tr][j - 3], acc[1][tr][j - 2], acc[1][tr][j - 1], acc[1][tr][j]));
        if (NJ == 3) __stcs(reinterpret_cast<float4*>(dst + 2 * jet_stride), make_float4(acc[2][tr][j - 3], acc[2][tr][j - 2], acc[2][tr][j - 1], acc[2][tr][j]));
      }
      act += act_rows * 4;
    }
  }
  const int tp0 = half * 64 + 2 * lane;
#pragma unroll
  for (int jt = 0; jt < NJ; ++jt)
    *reinterpret_cast<float2*>(Fpart + (ug * 3 + jt) * kTP + tp0) = make_float2(fp[jt][0], fp[jt][1]);
}

// ------------------------------------------------------------------------------------------------
// common setup of the small parameters
// ------------------------------------------------------------------------------------------------
template <class SM>
__device__ __forceinline__ void load_small(SM& sm, const float* __restrict__ theta, const float* __restrict__ nuclei,
                                           const float* __restrict__ onehot, int Na, int n_types) {
  const GnnThetaLayout L(n_types);
  const int tid = threadIdx.x;
  for (int i = tid; i < kH * kH; i += kThreads) sm.W2[i] = theta[L.W2 + i];
  if (tid < kH) {
    sm.wt[tid] = theta[L.W1 + tid];
    sm.wr[tid] = theta[L.W1 + kH + tid];
    sm.b2[tid] = theta[L.b2 + tid];
    sm.w3[tid] = theta[L.w3 + tid];
  }
  load_nuc_tables<kThreads>(sm.T, theta, L, nuclei, onehot, Na, n_types);
}

// ------------------------------------------------------------------------------------------------
// forward tile
// ------------------------------------------------------------------------------------------------
template <int NJ>
__device__ void fwd_tile(SmemFwd& sm, const GnnFwdArgs& a, long long row0, long long seg_end, float b3,
                         bool seg_b, long long seg_row0) {
  const ActLayout AL(a.act_rows_a, a.act_rows_b, a.jets_a, a.jets_b);
  const long long act_rows = seg_b ? a.act_rows_b : a.act_rows_a;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, half = w >> 2, ug = w & 3;
  const bool owner = tid < kTP;
  const long long row = row0 + tid;
  const bool valid = owner && row < seg_end;
  const long long rowc = row < seg_end ? row : seg_end - 1;
  float y[7] = {0, 0, 0, 0, 0, 0, 0}, accw[7] = {0, 0, 0, 0, 0, 0, 0}, kst[7] = {0, 0, 0, 0, 0, 0, 0};
  float xst[3] = {0, 0, 0}, sst[3] = {0, 0, 0};
  float g[3], dv, ds[3];
  if (owner) {
    const float* src = a.y_in + rowc * a.in_stride;
    y[0] = src[0]; y[1] = src[1]; y[2] = src[2];
    if (NJ >= 2) y[3] = src[3];
    if (NJ == 3) { y[4] = src[4]; y[5] = src[5]; y[6] = src[6]; }
  }
  const int n_steps = a.rhs_only ? 1 : a.n_steps;
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  float acc[NJ][2][16];
  float rr[2];

  auto combine = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    float f0 = b3, f1 = 0.f, f2 = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      f0 += sm.Fpart[(u * 3 + 0) * kTP + tid];
      if (NJ >= 2) f1 += sm.Fpart[(u * 3 + 1) * kTP + tid];
      if (NJ == 3) f2 += sm.Fpart[(u * 3 + 2) * kTP + tid];
    }
    const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
    g[0] = fmaf(f0, d0, g[0]); g[1] = fmaf(f0, d1, g[1]); g[2] = fmaf(f0, d2, g[2]);
    if (NJ >= 2) {
      const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
      dv += fmaf(f1, r, 3.0f * f0);
      if (NJ == 3) {
        const float q = fmaf(d0, sst[0], fmaf(d1, sst[1], d2 * sst[2]));
        const float amp = (fmaf(f1, q + 4.0f, f2 * r)) / r;
        ds[0] += fmaf(f0, sst[0], amp * d0);
        ds[1] += fmaf(f0, sst[1], amp * d1);
        ds[2] += fmaf(f0, sst[2], amp * d2);
      }
    }
  };

  for (int step = 0; step < n_steps; ++step) {
#pragma unroll 1
    for (int stage = 0; stage < 4; ++stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float ts = a.rhs_only ? a.t_rhs : a.t0 + h * (float)step + cs;
      const float tprime = a.y0 * ts;
      if (owner) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          xst[c] = fmaf(cs, kst[c], y[c]);
          sm.xs[c * kTP + tid] = xst[c];
          if (NJ == 3) sst[c] = fmaf(cs, kst[4 + c], y[4 + c]);
        }
        if (a.ckpt != nullptr && valid) {
          float* ck = a.ckpt + ((long long)(step * 4 + stage) * 6) * a.B + row;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ck[(long long)c * a.B] = xst[c];
            if (NJ == 3) ck[(long long)(3 + c) * a.B] = sst[c];
          }
        }
        g[0] = g[1] = g[2] = 0.f; dv = 0.f; ds[0] = ds[1] = ds[2] = 0.f;
      }
      __syncthreads();
      for (int i = 0; i < a.Na; ++i) {
        if (owner && i > 0) combine(i - 1);
        first_layer<NJ, true>(sm.A, sm.xs, sm.wr, sm.wt, sm.T.cb + sm.T.cls[i] * kH, sm.T.nuc + i * 4, tprime, half, ug, lane, rr);
        __syncthreads();
        gemm64<NJ, true>(acc, sm.A, sm.W2, nullptr, half, ug, lane);
        float* actp = nullptr;
        if (a.act != nullptr)
          actp = a.act + AL.seg_base((long long)(step * 4 + stage) * a.Na + i, seg_b) + (row0 - seg_row0) * 4;
        epilogue_fwd<NJ>(acc, sm.b2, sm.w3, sm.Fpart, half, ug, lane, actp, act_rows);
        __syncthreads();
      }
      if (owner) {
        combine(a.Na - 1);
        const float wgt = (stage == 0 || stage == 3) ? 1.0f : 2.0f;
        kst[0] = a.y0 * g[0]; kst[1] = a.y0 * g[1]; kst[2] = a.y0 * g[2];
        if (NJ >= 2) kst[3] = -a.y0 * dv;
        if (NJ == 3) { kst[4] = -a.y0 * ds[0]; kst[5] = -a.y0 * ds[1]; kst[6] = -a.y0 * ds[2]; }
#pragma unroll
        for (int c = 0; c < 7; ++c) accw[c] = fmaf(wgt, kst[c], accw[c]);
      }
      if (a.rhs_only) break;
    }
    if (owner && !a.rhs_only) {
#pragma unroll
      for (int c = 0; c < 7; ++c) { y[c] = fmaf(h * (1.0f / 6.0f), accw[c], y[c]); accw[c] = 0.f; }
    }
  }
  if (valid) {
    float* dst = a.y_out + row * a.out_stride;
    const float* o = a.rhs_only ? kst : y;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    if (NJ >= 2) dst[3] = o[3];
    if (NJ == 3) { dst[4] = o[4]; dst[5] = o[5]; dst[6] = o[6]; }
  }
}

__global__ void __launch_bounds__(kThreads, 1) gnn_fwd_kernel(const GnnFwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemFwd& sm = *reinterpret_cast<SmemFwd*>(smem_raw);
  load_small(sm, a.theta, a.nuclei, a.onehot, a.Na, a.n_types);
  const float b3 = a.theta[0];
  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    bool seg_b; long long seg_row0;
    if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; seg_b = false; seg_row0 = 0; }
    else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; seg_b = true; seg_row0 = a.n_a; }
    if (nj == 3) fwd_tile<3>(sm, a, row0, seg_end, b3, seg_b, seg_row0);
    else if (nj == 2) fwd_tile<2>(sm, a, row0, seg_end, b3, seg_b, seg_row0);
    else fwd_tile<1>(sm, a, row0, seg_end, b3, seg_b, seg_row0);
  }
}

// ------------------------------------------------------------------------------------------------
// backward pieces
// ------------------------------------------------------------------------------------------------
struct BwdAccum {
  float dW2[4][4];   // tile of dW2: rows 4*kg.., cols 4*jg..
  float g_b2w3;      // lane<16: d b2[16ug+lane]   else: d w3[16ug+lane-16]      (per traj-half)
  float g_wtwr;      // lane<16: d W1[0][16ug+lane] else: d W1[1][16ug+lane-16]
  float g_b3;        // owner threads: sum of fbar
};

// layer-2 tanh jet forward + reverse: writes Fpart and the pre-activation cotangent jets PB
template <int NJ, bool ACT>
__device__ __forceinline__ void epilogue_bwd(const float (&acc)[NJ][2][16], SmemBwd& sm, BwdAccum& G,
                                             int half, int ug, int lane) {
  const int tp0 = half * 64 + 2 * lane;
  float fb[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
  for (int jt = 0; jt < NJ; ++jt) {
    const float2 v = *reinterpret_cast<const float2*>(sm.fbar + jt * kTP + tp0);
    fb[jt][0] = v.x; fb[jt][1] = v.y;
  }
  float fp[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  float v32[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int jj = ug * 16 + j;
    const float b = sm.b2[jj], w = sm.w3[jj];
    float pz[2], p1b[2] = {0.f, 0.f}, p2b[2] = {0.f, 0.f};
    float gw = 0.f;
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
      const float a = ACT ? acc[0][tr][j] : tanh_f(acc[0][tr][j] + b);
      const float sg = fmaf(-a, a, 1.0f);
      fp[0][tr] = fmaf(w, a, fp[0][tr]);
      gw = fmaf(a, fb[0][tr], gw);
      const float ab = w * fb[0][tr];
      float pzb = ab;
      if (NJ >= 2) {
        const float p1 = acc[1][tr][j];
        const float a1 = sg * p1;
        fp[1][tr] = fmaf(w, a1, fp[1][tr]);
        gw = fmaf(a1, fb[1][tr], gw);
        const float a1b = w * fb[1][tr];
        pzb = fmaf(-2.0f * a * p1, a1b, pzb);
        float t1 = a1b;
        if (NJ == 3) {
          const float p2 = acc[2][tr][j];
          const float a2 = fmaf(sg, p2, -2.0f * a * a1 * p1);
          fp[2][tr] = fmaf(w, a2, fp[2][tr]);
          gw = fmaf(a2, fb[2][tr], gw);
          const float a2b = w * fb[2][tr];
          p2b[tr] = a2b * sg;
          t1 = fmaf(-4.0f * a * p1, a2b, t1);
          pzb = fmaf(a2b, fmaf(-2.0f * a, p2, -2.0f * fmaf(-3.0f * a, a, 1.0f) * p1 * p1), pzb);
        }
        p1b[tr] = sg * t1;
      }
      pz[tr] = sg * pzb;
    }
    const int off = act_off(jj, tp0);
    *reinterpret_cast<float2*>(sm.PB + off) = make_float2(pz[0], pz[1]);
    if (NJ >= 2) *reinterpret_cast<float2*>(sm.PB + kJetStride + off) = make_float2(p1b[0], p1b[1]);
    if (NJ == 3) *reinterpret_cast<float2*>(sm.PB + 2 * kJetStride + off) = make_float2(p2b[0], p2b[1]);
    v32[j] = pz[0] + pz[1];
    v32[16 + j] = gw;
  }
#pragma unroll
  for (int jt = 0; jt < NJ; ++jt)
    *reinterpret_cast<float2*>(sm.Fpart + (ug * 3 + jt) * kTP + tp0) = make_float2(fp[jt][0], fp[jt][1]);
  G.g_b2w3 += warp_vec32_sum(v32, lane);
}

// layer-1 reverse: acc = cotangent jets of (a, a', a'') for (2 traj x 16 units)
template <int NJ>
__device__ __forceinline__ void epilogue_b1(const float (&acc)[NJ][2][16], SmemBwd& sm, BwdAccum& G,
                                            const float (&r)[2], float tprime, int nucleus,
                                            int half, int ug, int lane) {
  const int tp0 = half * 64 + 2 * lane;
  float rp[2] = {0.f, 0.f};
  float v32[32];
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const int k = ug * 16 + kk;
    const float2 av = *reinterpret_cast<const float2*>(sm.A + act_off(k, tp0));
    const float w = sm.wr[k];
    const float aa[2] = {av.x, av.y};
    float ps = 0.f, pr = 0.f;
#pragma unroll
    for (int tr = 0; tr < 2; ++tr) {
      const float a = aa[tr];
      const float sg = fmaf(-a, a, 1.0f);
      float t = acc[0][tr][kk];
      float e = 0.f;
      if (NJ >= 2) {
        const float a1b = acc[1][tr][kk];
        t = fmaf(-2.0f * a * w, a1b, t);
        e = a1b;
        if (NJ == 3) {
          const float a2b = acc[2][tr][kk];
          t = fmaf(-2.0f * w * w * fmaf(-3.0f * a, a, 1.0f), a2b, t);
          e = fmaf(-4.0f * a * w, a2b, e);
        }
      }
      const float pb = sg * t;
      rp[tr] = fmaf(pb, w, rp[tr]);
      ps += pb;
      pr += fmaf(pb, r[tr], sg * e);
    }
    v32[kk] = ps;
    v32[16 + kk] = pr;
  }
  *reinterpret_cast<float2*>(sm.Rpart + ug * kTP + tp0) = make_float2(rp[0], rp[1]);
  const float red = warp_vec32_sum(v32, lane);
  if (lane < 16) {
    sm.S[(half * kMaxClasses + nucleus) * kH + ug * 16 + lane] += red;
    G.g_wtwr = fmaf(tprime, red, G.g_wtwr);
  } else {
    G.g_wtwr += red;
  }
}

// dW2[k][j] += sum_{jet, traj} A[jet][k][traj] * PB[jet][j][traj]
template <int NJ>
__device__ __forceinline__ void dw2_accum(BwdAccum& G, const SmemBwd& sm, int w, int lane) {
  const int kg = 4 * (w >> 1) + (lane & 3);
  const int jg = 8 * (w & 1) + (lane >> 2);
  const int k0 = 4 * kg, j0 = 4 * jg;
  const int swk = kg & 7, swj = jg & 7;
  float c2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c2[i] = sm.wr2[k0 + i];
#pragma unroll 2
  for (int c = 0; c < kTP / 4; ++c) {
    const int offk = ((c ^ swk) << 2), offj = ((c ^ swj) << 2);
    float4 av[3][4], pv[3][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      av[0][i] = *reinterpret_cast<const float4*>(sm.A + (k0 + i) * kTP + offk);
      if (NJ >= 2) av[1][i] = *reinterpret_cast<const float4*>(sm.A + kJetStride + (k0 + i) * kTP + offk);
      if (NJ == 3) {
        av[2][i].x = av[0][i].x * av[1][i].x * c2[i];
        av[2][i].y = av[0][i].y * av[1][i].y * c2[i];
        av[2][i].z = av[0][i].z * av[1][i].z * c2[i];
        av[2][i].w = av[0][i].w * av[1][i].w * c2[i];
      }
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt)
        pv[jt][i] = *reinterpret_cast<const float4*>(sm.PB + jt * kJetStride + (j0 + i) * kTP + offj);
    }
#pragma unroll
    for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
      for (int ik = 0; ik < 4; ++ik)
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
          float s = G.dW2[ik][ij];
          s = fmaf(av[jt][ik].x, pv[jt][ij].x, s);
          s = fmaf(av[jt][ik].y, pv[jt][ij].y, s);
          s = fmaf(av[jt][ik].z, pv[jt][ij].z, s);
          s = fmaf(av[jt][ik].w, pv[jt][ij].w, s);
          G.dW2[ik][ij] = s;
        }
  }
}

template <int NJ, bool ACT>
__device__ void bwd_tile(SmemBwd& sm, const GnnBwdArgs& a, long long row0, long long seg_end, float b3, BwdAccum& G,
                         bool seg_b, long long seg_row0) {
  const ActLayout AL(a.act_rows_a, a.act_rows_b, a.jets_a, a.jets_b);
  const long long act_rows = seg_b ? a.act_rows_b : a.act_rows_a;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, half = w >> 2, ug = w & 3;
  const bool owner = tid < kTP;
  const long long row = row0 + tid;
  const bool valid = owner && row < seg_end;
  const long long rowc = row < seg_end ? row : seg_end - 1;
  // cotangent at the end of the current step: xbar, lbar (constant along the trajectory), sbar
  float xb[3] = {0, 0, 0}, lb = 0.f, sb[3] = {0, 0, 0};
  if (valid) {
    const float* src = a.ybar1 + row * a.bar_stride;
    xb[0] = src[0]; xb[1] = src[1]; xb[2] = src[2];
    if (NJ >= 2) lb = src[3];
    if (NJ == 3) { sb[0] = src[4]; sb[1] = src[5]; sb[2] = src[6]; }
  }
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  float acc[NJ][2][16];
  float rr[2];
  float xst[3] = {0, 0, 0}, sst[3] = {0, 0, 0};
  float kx[3], kl = 0.f, ks[3] = {0, 0, 0};           // cotangent of the stage derivative k
  float ysx[3], yss[3];                              // cotangent of the stage state (accumulated over nuclei)
  float sumx[3], sums[3], prevx[3] = {0, 0, 0}, prevs[3] = {0, 0, 0};

  // owner: cotangents of f, f', f'' for nucleus i -> sm.fbar
  auto pre = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
    const float alpha = fmaf(d0, kx[0], fmaf(d1, kx[1], d2 * kx[2]));
    float f0b = alpha, f1b = 0.f, f2b = 0.f;
    if (NJ >= 2) {
      const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
      f0b = fmaf(-3.0f, kl, f0b);
      f1b = -r * kl;
      if (NJ == 3) {
        const float q = fmaf(d0, sst[0], fmaf(d1, sst[1], d2 * sst[2]));
        const float beta = fmaf(ks[0], sst[0], fmaf(ks[1], sst[1], ks[2] * sst[2]));
        const float gam = fmaf(d0, ks[0], fmaf(d1, ks[1], d2 * ks[2])) / r;
        f0b -= beta;
        f1b = fmaf(-(q + 4.0f), gam, f1b);
        f2b = -r * gam;
      }
    }
    f0b *= a.y0; f1b *= a.y0; f2b *= a.y0;
    sm.fbar[tid] = f0b;
    if (NJ >= 2) sm.fbar[kTP + tid] = f1b;
    if (NJ == 3) sm.fbar[2 * kTP + tid] = f2b;
    G.g_b3 += f0b;
  };
  // owner: explicit + network contributions of nucleus i to the stage-state cotangent
  auto post = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    float f0 = b3, f1 = 0.f, f2 = 0.f, rbar = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      f0 += sm.Fpart[(u * 3 + 0) * kTP + tid];
      if (NJ >= 2) f1 += sm.Fpart[(u * 3 + 1) * kTP + tid];
      if (NJ == 3) f2 += sm.Fpart[(u * 3 + 2) * kTP + tid];
      rbar += sm.Rpart[u * kTP + tid];
    }
    const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
    const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
    const float rinv = 1.0f / r;
    const float u0 = d0 * rinv, u1 = d1 * rinv, u2 = d2 * rinv;
    float db0 = a.y0 * f0 * kx[0], db1 = a.y0 * f0 * kx[1], db2 = a.y0 * f0 * kx[2];
    if (NJ >= 2) rbar = fmaf(-a.y0 * f1, kl, rbar);
    if (NJ == 3) {
      const float q = fmaf(d0, sst[0], fmaf(d1, sst[1], d2 * sst[2]));
      const float gam = fmaf(u0, ks[0], fmaf(u1, ks[1], u2 * ks[2]));
      const float amp = fmaf(f1, q + 4.0f, f2 * r);
      rbar = fmaf(-a.y0 * f2, gam, rbar);
      const float qbar = -a.y0 * f1 * gam;
      const float ua = -a.y0 * amp;                  // ubar = ua * ks
      db0 = fmaf(qbar, sst[0], fmaf(ua * rinv, ks[0], db0));
      db1 = fmaf(qbar, sst[1], fmaf(ua * rinv, ks[1], db1));
      db2 = fmaf(qbar, sst[2], fmaf(ua * rinv, ks[2], db2));
      rbar = fmaf(-ua * gam, rinv, rbar);
      yss[0] += fmaf(qbar, d0, -a.y0 * f0 * ks[0]);
      yss[1] += fmaf(qbar, d1, -a.y0 * f0 * ks[1]);
      yss[2] += fmaf(qbar, d2, -a.y0 * f0 * ks[2]);
    }
    ysx[0] += fmaf(rbar, u0, db0);
    ysx[1] += fmaf(rbar, u1, db1);
    ysx[2] += fmaf(rbar, u2, db2);
  };

  for (int step = a.n_steps - 1; step >= 0; --step) {
    sumx[0] = sumx[1] = sumx[2] = 0.f; sums[0] = sums[1] = sums[2] = 0.f;
#pragma unroll 1
    for (int stage = 3; stage >= 0; --stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float ts = a.t0 + h * (float)step + cs;
      const float tprime = a.y0 * ts;
      if (owner) {
        // k4bar = h/6 ybar ; k3bar = h/3 ybar + h y4bar ; k2bar = h/3 ybar + h/2 y3bar ; k1bar = h/6 ybar + h/2 y2bar
        const float ca = (stage == 0 || stage == 3) ? h * (1.0f / 6.0f) : h * (1.0f / 3.0f);
        const float cp = stage == 3 ? 0.f : (stage == 2 ? h : 0.5f * h);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          kx[c] = fmaf(ca, xb[c], cp * prevx[c]);
          if (NJ == 3) ks[c] = fmaf(ca, sb[c], cp * prevs[c]);
        }
        if (NJ >= 2) kl = ca * lb;
        const float* ck = a.ckpt + ((long long)(step * 4 + stage) * 6) * a.B + rowc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          xst[c] = ck[(long long)c * a.B];
          sm.xs[c * kTP + tid] = xst[c];
          if (NJ == 3) sst[c] = ck[(long long)(3 + c) * a.B];
        }
        ysx[0] = ysx[1] = ysx[2] = 0.f; yss[0] = yss[1] = yss[2] = 0.f;
        pre(0);
      }
      __syncthreads();
      for (int i = 0; i < a.Na; ++i) {
        if (ACT) {
          // issue the checkpoint loads first: their latency hides behind the first-layer phase
          const float* src = a.act + AL.seg_base((long long)(step * 4 + stage) * a.Na + i, seg_b)
                             + act_group_off(ug * 4, (row0 - seg_row0) + half * 64 + 2 * lane, act_rows);
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int tr = 0; tr < 2; ++tr) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (long long)jt * kH * act_rows + (long long)g * act_rows * 4 + tr * 4));
                acc[jt][tr][4 * g] = v.x; acc[jt][tr][4 * g + 1] = v.y; acc[jt][tr][4 * g + 2] = v.z; acc[jt][tr][4 * g + 3] = v.w;
              }
        }
        first_layer<NJ, false>(sm.A, sm.xs, sm.wr, sm.wt, sm.T.cb + sm.T.cls[i] * kH, sm.T.nuc + i * 4, tprime, half, ug, lane, rr);
        __syncthreads();
        if (!ACT) gemm64<NJ, false>(acc, sm.A, sm.W2, sm.wr2, half, ug, lane);
        epilogue_bwd<NJ, ACT>(acc, sm, G, half, ug, lane);
        __syncthreads();
        gemm64<NJ, true>(acc, sm.PB, sm.W2T, nullptr, half, ug, lane);
        epilogue_b1<NJ>(acc, sm, G, rr, tprime, sm.T.cls[i], half, ug, lane);
        dw2_accum<NJ>(G, sm, w, lane);
        __syncthreads();
        if (owner) {
          post(i);
          if (i + 1 < a.Na) pre(i + 1);
        }
      }
      if (owner) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sumx[c] += ysx[c]; prevx[c] = ysx[c];
          sums[c] += yss[c]; prevs[c] = yss[c];
        }
      }
    }
    if (owner) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { xb[c] += sumx[c]; sb[c] += sums[c]; }
    }
  }
  if (valid && a.ybar0 != nullptr) {
    float* dst = a.ybar0 + row * a.bar_stride;
    dst[0] = xb[0]; dst[1] = xb[1]; dst[2] = xb[2];
    if (NJ >= 2) dst[3] = lb;
    if (NJ == 3) { dst[4] = sb[0]; dst[5] = sb[1]; dst[6] = sb[2]; }
  }
}

// per-TILE partial gradient -> gpart[tile][n_theta]; resets the accumulators.  One partial per tile (not per CTA)
// makes the reduction order independent of the dynamic tile schedule: the gradient is bitwise reproducible.
__device__ void flush_tile_grad(SmemBwd& sm, BwdAccum& G, float* __restrict__ out, const GnnThetaLayout& L,
                                const float* __restrict__ onehot, int Na, int n_types) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, half = w >> 2, ug = w & 3;
  {
    const int kg = 4 * (w >> 1) + (lane & 3), jg = 8 * (w & 1) + (lane >> 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {             // theta offsets are odd: no vector stores
        out[L.W2 + (4 * kg + i) * kH + 4 * jg + j] = G.dW2[i][j];
        G.dW2[i][j] = 0.f;
      }
    }
  }
  // red[which][half][ug][lane]
  sm.red[((0 * 2 + half) * 4 + ug) * 32 + lane] = G.g_b2w3;
  sm.red[((1 * 2 + half) * 4 + ug) * 32 + lane] = G.g_wtwr;
  const float b3w = warp_sum(G.g_b3);
  G.g_b2w3 = 0.f; G.g_wtwr = 0.f; G.g_b3 = 0.f;
  __syncthreads();
  if (half == 0) {
    const float v0 = sm.red[((0 * 2 + 0) * 4 + ug) * 32 + lane] + sm.red[((0 * 2 + 1) * 4 + ug) * 32 + lane];
    const float v1 = sm.red[((1 * 2 + 0) * 4 + ug) * 32 + lane] + sm.red[((1 * 2 + 1) * 4 + ug) * 32 + lane];
    if (lane < 16) { out[L.b2 + ug * 16 + lane] = v0; out[L.W1 + ug * 16 + lane] = v1; }
    else { out[L.w3 + ug * 16 + lane - 16] = v0; out[L.W1 + kH + ug * 16 + lane - 16] = v1; }
  }
  __syncthreads();
  if (lane == 0) sm.red[w] = b3w;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int i = 0; i < 4; ++i) s += sm.red[i];     // owners are warps 0..3
    out[L.b3] = s;
  }
  if (tid < kH) {
    float sb1 = 0.f;
    for (int c = 0; c < sm.T.n_cls; ++c) sb1 += sm.S[(0 * kMaxClasses + c) * kH + tid] + sm.S[(1 * kMaxClasses + c) * kH + tid];
    out[L.b1 + tid] = sb1;
    for (int t = 0; t < n_types; ++t) {
      float s = 0.f;
      for (int c = 0; c < sm.T.n_cls; ++c)
        s = fmaf(onehot[sm.T.rep[c] * n_types + t], sm.S[(0 * kMaxClasses + c) * kH + tid] + sm.S[(1 * kMaxClasses + c) * kH + tid], s);
      out[L.W1 + (2 + t) * kH + tid] = s;
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * kMaxClasses * kH; i += kThreads) sm.S[i] = 0.f;
}

template <bool ACT>
__global__ void __launch_bounds__(kThreads, 1) gnn_bwd_kernel(const GnnBwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemBwd& sm = *reinterpret_cast<SmemBwd*>(smem_raw);
  const int tid = threadIdx.x;
  load_small(sm, a.theta, a.nuclei, a.onehot, a.Na, a.n_types);
  const GnnThetaLayout L(a.n_types);
  for (int i = tid; i < kH * kH; i += kThreads) sm.W2T[(i % kH) * kH + (i / kH)] = a.theta[L.W2 + i];
  if (tid < kH) sm.wr2[tid] = -2.0f * a.theta[L.W1 + kH + tid];
  for (int i = tid; i < 2 * kMaxClasses * kH; i += kThreads) sm.S[i] = 0.f;
  const float b3 = a.theta[0];
  BwdAccum G;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) G.dW2[i][j] = 0.f;
  G.g_b2w3 = 0.f; G.g_wtwr = 0.f; G.g_b3 = 0.f;

  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    bool seg_b; long long seg_row0;
    if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; seg_b = false; seg_row0 = 0; }
    else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; seg_b = true; seg_row0 = a.n_a; }
    if (nj == 3) bwd_tile<3, ACT>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0);
    else if (nj == 2) bwd_tile<2, ACT>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0);
    else bwd_tile<1, ACT>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0);
    __syncthreads();
    flush_tile_grad(sm, G, a.gpart + (size_t)tile * L.n, L, a.onehot, a.Na, a.n_types);
  }
}

// deterministic reduction of the per-tile partial gradients (float64 accumulation, fixed tile order)
__global__ void gnn_grad_reduce_kernel(const float* __restrict__ gpart, float* __restrict__ theta_bar, int n, int nblk) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)gpart[(size_t)b * n + p];
  theta_bar[p] = (float)s;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------

static int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

static int check_common(long long B, long long n_a, int ja, int jb, int Na, int nt, int L, int n_steps) {
  if (B < 0 || n_a < 0 || n_a > B || n_steps <= 0) return OFDFT_EINVAL;
  if (ja < 1 || ja > 3 || jb < 1 || jb > 3) return OFDFT_EINVAL;
  if (Na < 1 || Na > kMaxNa || nt < 0 || nt > kMaxTypes || L < 1 || L > kMaxLayers) return OFDFT_EUNSUPPORTED;
  return 0;
}

}  // namespace ofdft

using namespace ofdft;

extern "C" size_t ofdft_cnf_gnn_ckpt_bytes(int64_t B, int32_t n_steps) {
  return (size_t)n_steps * 4 * 6 * (size_t)B * sizeof(float);
}
extern "C" size_t ofdft_cnf_gnn_act_ckpt_bytes(int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b, int32_t Na,
                                               int32_t n_steps) {
  if (B <= 0 || n_a < 0 || n_a > B || Na <= 0 || n_steps <= 0) return 0;
  const ActLayout AL(pad_rows(n_a), pad_rows(B - n_a), jets_a, jets_b);
  return (size_t)n_steps * 4 * (size_t)Na * (size_t)AL.block * sizeof(float);
}
extern "C" size_t ofdft_cnf_gnn_fwd_scratch_bytes(void) { return kCounterBytes; }
static long long n_tiles_of(long long B, long long n_a) { return (n_a + kTP - 1) / kTP + (B - n_a + kTP - 1) / kTP; }
extern "C" int64_t ofdft_cnf_gnn_theta_size(int32_t n_types, int32_t n_layers) {
  if (n_types < 0 || n_types > kMaxTypes || n_layers < 1 || n_layers > kMaxLayers) return 0;
  return gnn_general_theta_size(n_types, n_layers);
}
extern "C" size_t ofdft_cnf_gnn_bwd_scratch_bytes(int64_t B, int64_t n_a, int32_t n_types, int32_t n_layers) {
  if (B <= 0 || n_a < 0 || n_a > B) return kCounterBytes;
  return kCounterBytes + kDeepWSplitBytes + (size_t)n_tiles_of(B, n_a) * (size_t)ofdft_cnf_gnn_theta_size(n_types, n_layers) * sizeof(float);
}

// path: OFDFT_PATH_DEFAULT (tcgen05 3xTF32 layer-2 GEMM, gnn_tc.cu) | OFDFT_PATH_FFMA (FP32-pipe kernel of this file)
static int launch_fwd(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes, int path, int n_layers) {
  if (!a.theta || !a.nuclei || !a.y_in || !a.y_out || !scratch) return OFDFT_EINVAL;
  if (a.n_types > 0 && !a.onehot) return OFDFT_EINVAL;
  if (scratch_bytes < kCounterBytes) return OFDFT_ESCRATCH;
  // three hidden layers (or two with OFDFT_PATH_DEEP): depth-generic tensor-core kernels (gnn_deep_tc.cu)
  if ((n_layers == 3 && !(path & OFDFT_PATH_FFMA)) || (n_layers == 2 && (path & OFDFT_PATH_DEEP)))
    return launch_gnn_deep_fwd_tc(stream, a, scratch, scratch_bytes, n_layers);
  if (n_layers != 2) return launch_gnn_general_fwd(stream, a, scratch, scratch_bytes, n_layers);   // gnn_general.cu
  if (!(path & OFDFT_PATH_FFMA)) return launch_gnn_fwd_tc(stream, a, scratch, scratch_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  a.counter = (int*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemFwd));
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (a.n_a + kTP - 1) / kTP + (a.B - a.n_a + kTP - 1) / kTP;
  int sms = num_sms(); if (sms <= 0) sms = 148;
  const int grid = (int)(tiles < sms ? tiles : sms);
  gnn_fwd_kernel<<<grid, kThreads, sizeof(SmemFwd), st>>>(a);
  return (int)cudaGetLastError();
}

extern "C" int ofdft_cnf_gnn_fwd(void* stream, const float* theta, const float* nuclei, const float* onehot,
                                 const float* y_in, float* y_out, float* ckpt, float* act_ckpt,
                                 void* scratch, size_t scratch_bytes,
                                 int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b,
                                 int32_t in_stride, int32_t out_stride,
                                 int32_t Na, int32_t n_types, int32_t n_layers, int32_t n_steps, float t0, float t1,
                                 float y0sign, int32_t path) {
  int rc = check_common(B, n_a, jets_a, jets_b, Na, n_types, n_layers, n_steps);
  if (rc) return rc;
  if (B == 0) return 0;   // empty shard: nothing to integrate
  const int jmax = jets_a > jets_b ? jets_a : jets_b;
  const int need = jmax == 3 ? 7 : (jmax == 2 ? 4 : 3);
  if (in_stride < need || out_stride < need) return OFDFT_EINVAL;
  GnnFwdArgs a{theta, nuclei, onehot, y_in, y_out, ckpt, nullptr, B, n_a, jets_a, jets_b, in_stride, out_stride,
               Na, n_types, n_steps, t0, t1, y0sign, 0, 0.f, n_layers == 2 ? act_ckpt : nullptr, pad_rows(n_a), pad_rows(B - n_a)};
  return launch_fwd(stream, a, scratch, scratch_bytes, path, n_layers);
}

extern "C" int ofdft_cnf_gnn_rhs(void* stream, const float* theta, const float* nuclei, const float* onehot,
                                 const float* y, float* dy, void* scratch, size_t scratch_bytes,
                                 int64_t B, int32_t jets, int32_t stride, int32_t Na, int32_t n_types, int32_t n_layers,
                                 float t, float y0sign, int32_t path) {
  int rc = check_common(B, B, jets, jets, Na, n_types, n_layers, 1);
  if (rc) return rc;
  if (B == 0) return 0;
  const int need = jets == 3 ? 7 : (jets == 2 ? 4 : 3);
  if (stride < need) return OFDFT_EINVAL;
  GnnFwdArgs a{theta, nuclei, onehot, y, dy, nullptr, nullptr, B, B, jets, jets, stride, stride,
               Na, n_types, 1, 0.f, 1.f, y0sign, 1, t, nullptr, 0, 0};
  return launch_fwd(stream, a, scratch, scratch_bytes, path, n_layers);
}

extern "C" int ofdft_cnf_gnn_bwd(void* stream, const float* theta, const float* nuclei, const float* onehot,
                                 const float* ckpt, const float* act_ckpt, const float* ybar1, float* theta_bar,
                                 float* ybar0, void* scratch, size_t scratch_bytes,
                                 int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b, int32_t bar_stride,
                                 int32_t Na, int32_t n_types, int32_t n_layers, int32_t n_steps, float t0, float t1,
                                 float y0sign, int32_t path) {
  int rc = check_common(B, n_a, jets_a, jets_b, Na, n_types, n_layers, n_steps);
  if (rc) return rc;
  const int n_theta = (int)ofdft_cnf_gnn_theta_size(n_types, n_layers);
  if (B == 0) {           // empty shard: the gradient of an empty sum
    if (!theta_bar) return OFDFT_EINVAL;
    return (int)cudaMemsetAsync(theta_bar, 0, (size_t)n_theta * sizeof(float), (cudaStream_t)stream);
  }
  if (!theta || !nuclei || !ckpt || !ybar1 || !theta_bar || !scratch) return OFDFT_EINVAL;
  if (n_types > 0 && !onehot) return OFDFT_EINVAL;
  const int jmax = jets_a > jets_b ? jets_a : jets_b;
  const int need = jmax == 3 ? 7 : (jmax == 2 ? 4 : 3);
  if (bar_stride < need) return OFDFT_EINVAL;
  if (scratch_bytes < ofdft_cnf_gnn_bwd_scratch_bytes(B, n_a, n_types, n_layers)) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gnn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemBwd));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gnn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemBwd));
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (n_a + kTP - 1) / kTP + (B - n_a + kTP - 1) / kTP;
  int sms = num_sms(); if (sms <= 0) sms = 148;
  const int grid = (int)(tiles < sms ? tiles : sms);
  float* w2split = (float*)((char*)scratch + kCounterBytes);
  float* gpart = (float*)((char*)scratch + kCounterBytes + kDeepWSplitBytes);
  GnnBwdArgs a{theta, nuclei, onehot, ckpt, ybar1, ybar0, gpart, (int*)scratch, B, n_a, jets_a, jets_b, bar_stride,
               Na, n_types, n_steps, t0, t1, y0sign, n_layers == 2 ? act_ckpt : nullptr, pad_rows(n_a), pad_rows(B - n_a), w2split};
  // default: tcgen05 3xTF32 GEMMs (without an activation checkpoint the layer-2 jets are recomputed on the tensor cores);
  // OFDFT_PATH_FFMA: FP32-pipe kernels (checkpoint-reading or recomputing)
  const bool tc_bwd = !(path & OFDFT_PATH_FFMA);
  if ((n_layers == 3 && tc_bwd) || (n_layers == 2 && (path & OFDFT_PATH_DEEP))) {
    a.act = nullptr;
    const int rc2 = launch_gnn_deep_bwd_tc(stream, a, grid, n_layers);   // gnn_deep_tc.cu (no activation checkpoint)
    if (rc2) return rc2;
  } else if (n_layers != 2) {
    const int rc2 = launch_gnn_general_bwd(stream, a, n_layers);      // gnn_general.cu (no activation checkpoint)
    if (rc2) return rc2;
  } else if (tc_bwd) {
    const int rc = launch_gnn_bwd_tc(stream, a, grid);
    if (rc) return rc;
  } else if (act_ckpt != nullptr) gnn_bwd_kernel<true><<<grid, kThreads, sizeof(SmemBwd), st>>>(a);
  else gnn_bwd_kernel<false><<<grid, kThreads, sizeof(SmemBwd), st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  gnn_grad_reduce_kernel<<<(n_theta + 127) / 128, 128, 0, st>>>(gpart, theta_bar, n_theta, (int)tiles);
  return (int)cudaGetLastError();
}
