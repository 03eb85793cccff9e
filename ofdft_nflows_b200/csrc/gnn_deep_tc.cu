// gnn_deep_tc.cu -- tensor-core (tcgen05 / TMEM) kernels of the 3-D equivariant GNN flow for a radial MLP of TWO OR THREE
// hidden layers of 64 units (`--nn 3` of OFDFT_NF.py:320-326 selects features (64, 64, 64); equiv_flows.py:36-62 builds the
// stack): forward integrator and discrete adjoint without any activation checkpoint.
//
// The (64, 64) network has the pipelined, warp-specialised kernels of gnn_tc.cu; this file is the depth-generic version:
// the same tiles (128 trajectories per CTA, thread (warp w, lane l) = trajectory m = 32 (w & 3) + l -- the TMEM lane its warp
// may access -- and unit half uh = w >> 2), the same operand layouts and the same error-compensated 3xTF32 GEMMs
//     x.w ~= x_hi.w_hi + x_hi.w_lo + x_lo.w_hi         (fp32 accumulation in tensor memory)
// but one phase sequence per (RK4 stage, nucleus) that every thread walks through: each hand-off to the tensor core is
// `stage operands -> __syncthreads -> one elected lane issues the MMAs + tcgen05.commit`, each completion one mbarrier that
// all threads wait on exactly once (two barriers for the two operand buffers of a GEMM -- the staging of jet jt + 1 runs under
// the MMAs of jet jt --, one for the weight-gradient batches; the last batch of a layer runs under the next reverse epilogue).
//
// Tensor memory (512 columns) is the layer-state store of the adjoint.  S(l) = 192 (l - 1) .. + 191 holds hidden layer l >= 1:
// the forward GEMM leaves the pre-activation jets (P, P', P'') there, the epilogue overwrites P by a = tanh(P + b_l) IN PLACE
// and (a, P', P'') is exactly what the reverse sweep needs (a' = (1 - a^2) P', a'' = (1 - a^2) P'' - 2 a a' P').  Once layer
// l has been reversed its columns are dead and receive the cotangent jets Abar_{l-1} = Pbar_l . W_l^T.  Columns 384..511 are
// the stacked [hi k | lo k] x [hi j | lo j] weight-gradient accumulator (one chain per (nucleus, evaluation, layer): 48 MMAs,
// folded into round-to-nearest fp32 sums in shared memory -- the tensor core accumulates with truncation); outside the
// weight-gradient GEMM they carry the two unrounded "hi" A-operand buffers of the TS-mode passes.
//
// Shared memory: one 129 KB operand region -- [A^T | Pbar^T] stacked K = trajectory operands of the weight gradient, or, in the
// GEMM phases, two tf32-remainder buffers of the A operand + two hi | lo weight-operand buffers, filled by cp.async one GEMM
// phase ahead from a prepared global buffer (forward / transposed orientation: 4 x 16 KB per layer) --, 16 KB of fp32
// weight-gradient sums per layer and the first-layer activations of the current (nucleus, evaluation).  The layer loops are
// real loops (one copy of the layer body): unrolled, the code did not fit the instruction cache.
#include "gnn_common.cuh"
#include "tc.cuh"

namespace ofdft {

constexpr int kDChunkA = kTP * 16;          // LBO of a K-major [128 rows] operand
constexpr int kDChunkB = kH * 16;           // LBO of a K-major [64 rows] operand
constexpr int kDTChunk = kTP * 4 + 4;       // floats per trajectory-chunk of a stacked 128-row K = trajectory operand (16 B pad)
constexpr int kDColDw = 384;                // weight-gradient accumulator (128 columns) / TS-mode "hi" operand (first 64)
__host__ __device__ constexpr int dcol_state(int l) { return 192 * (l - 1); }
constexpr int kDRegionFloats = 2 * 32 * kDTChunk;
constexpr int kDWopOff = 2 * kH * kTP;      // float offset of the two weight operand buffers (hi | lo each) inside the region; two remainder buffers in front
static_assert(kDWopOff + 2 * 2 * kH * kH <= kDRegionFloats, "GEMM-phase operands exceed the staging region");

struct SmemDeepF {
  float Alo[2 * kH * kTP];            // tf32 remainders of the jet being staged, two buffers
  float Whi[(kMaxLayers - 1) * kH * kH], Wlo[(kMaxLayers - 1) * kH * kH];   // [layer][k/4][j][k%4]
  float wr[kH], wt[kH], w3[kH], bl[(kMaxLayers - 1) * kH];
  NucTables T;
  float xs[3 * kTP];
  float Fpart[2 * 3 * kTP];
  uint64_t bar[3];
  unsigned int tmem_slot;
  int tile;
};

struct SmemDeepB {
  float R[kDRegionFloats];
  float DWs[(kMaxLayers - 1) * kH * kH];    // [layer - 1][j][k]
  float A0[kH * kTP];                       // first-layer activations of the current (nucleus, evaluation): [unit pair][trajectory][2]
  float wr[kH], wt[kH], w3[kH], bl[(kMaxLayers - 1) * kH];
  NucTables T;
  float S[8 * kMaxClasses * 32];            // per warp and nucleus class: sum over the warp's trajectories of pbar1[k = 32 uh + lane]
  float Fpart[2 * 3 * kTP];
  float Rpart[2 * kTP];
  float red[8 * 32 * (3 + kMaxLayers)];
  uint64_t bar[3];
  unsigned int tmem_slot;
  int tile;
};
static_assert(sizeof(SmemDeepB) + 1024 <= 227 * 1024, "backward tile state exceeds the 227 KB of shared memory per CTA");

// wait for completion barrier b (0 / 1: GEMM of the jet staged in operand buffer b, 2: weight-gradient batch; bit b of `phase`
// is its parity; every commit is followed by exactly one wait of all threads); a wait that lasts seconds is a protocol bug:
// trap instead of hanging the device
__device__ __forceinline__ void deep_wait(uint64_t* bars, uint32_t& phase, int b) {
  const uint32_t addr = tc::smem_u32(bars + b), par = (phase >> b) & 1u;
  auto probe = [&]() {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(par)
        : "memory");
    return done != 0;
  };
  if (!probe()) {                                   // the clock is only read when the batch has not completed yet
    const long long t0 = clock64();
    while (!probe())
      if (clock64() - t0 > 8000000000ll) __trap();
  }
  phase ^= 1u << b;
  __syncwarp();
  tc::fence_after_sync();
}

// The elementwise phases run on packed FP32 (FFMA2 / FMUL2 / FADD2: two hidden units per instruction, tc.cuh): they are
// issue-bound at two warps per scheduler, and half the instructions is also half the code the instruction cache has to hold.
// tf32 remainders x - trunc_tf32(x) = trunc * (-1) + x (exact) of two values
__device__ __forceinline__ tc::f32x2 lo2pk(tc::f32x2 p) { return tc::fma2(p & 0xFFFFE000FFFFE000ull, tc::splat2(-1.0f), p); }
// tanh of two values: 1 - 2 / (2^(2 x log2 e) + 1), one MUFU.EX2 + MUFU.RCP per value
__device__ __forceinline__ tc::f32x2 deep_tanh2(tc::f32x2 x) {
  using namespace tc;
  float x0, x1, e0, e1, d0, d1, r0, r1;
  unpack2(mul2(x, splat2(2.8853900817779268f)), x0, x1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
  unpack2(add2(pack2(e0, e1), splat2(1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return fma2(splat2(-2.0f), pack2(r0, r1), splat2(1.0f));
}
__device__ __forceinline__ tc::f32x2 ld2(const float* p) { const float2 v = *reinterpret_cast<const float2*>(p); return tc::pack2(v.x, v.y); }

// 32 units (32 uh ..) of trajectory m as the A operand of a [128 x 64] . [64 x 64] GEMM: the unrounded values go to tensor
// memory (the tensor core truncates them to tf32), the remainders x - trunc_tf32(x) to the K-major shared-memory buffer
__device__ __forceinline__ void deep_stage_A(const float (&x)[32], uint32_t tcol, float* Alo, int uh, int m) {
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 16) {
    float hv[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) hv[u] = x[c0 + u];
    tc::tmem_st16(tcol + c0, hv);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int off = ((8 * uh + c) * kTP + m) * 4;
    const tc::f32x2 l01 = lo2pk(tc::pack2(x[4 * c], x[4 * c + 1])), l23 = lo2pk(tc::pack2(x[4 * c + 2], x[4 * c + 3]));
    *reinterpret_cast<float4*>(Alo + off) = make_float4(tc::lo2(l01), tc::hi2(l01), tc::lo2(l23), tc::hi2(l23));
  }
}

// after the staging stores of all threads: barrier, then one elected lane of warp 0 issues D = A . B as 3xTF32 (24 MMAs)
__device__ __forceinline__ void deep_sync_issue_gemm(uint32_t d, uint32_t ahi, uint64_t dAlo, uint64_t dWh, uint64_t dWl, uint32_t idesc,
                                                     uint64_t* bar) {
  tc::tmem_st_wait();
  tc::tmem_ld_wait();
  asm volatile("cp.async.wait_all;" ::: "memory");     // weight operand copies of this thread (backward kernel)
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc::fence_after_sync();
    if (tc::elect_one()) {
      constexpr uint64_t stepA = (2 * kDChunkA) >> 4, stepB = (2 * kDChunkB) >> 4;
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, ahi + 8 * ks, dWh + ks * stepB, idesc, ks != 0);
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, ahi + 8 * ks, dWl + ks * stepB, idesc, true);
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) tc::mma_tf32(d, dAlo + ks * stepA, dWh + ks * stepB, idesc, true);
      tc::commit(bar);
    }
    __syncwarp();
  }
}

// first layer for (trajectory, units 32 uh ..): a = tanh(w_r r + w_t t' + cb)
__device__ __forceinline__ void deep_first_layer(float (&av)[32], const float* wr, const float* wt, const float* cbi, float r, float tprime) {
  using namespace tc;
  const f32x2 r2 = splat2(r), tp2 = splat2(tprime);
#pragma unroll
  for (int u = 0; u < 32; u += 2) unpack2(deep_tanh2(fma2(ld2(wr + u), r2, fma2(tp2, ld2(wt + u), ld2(cbi + u)))), av[u], av[u + 1]);
}
// jet jt of the first layer: a, a' = (1 - a^2) w_r, a'' = -2 a a' w_r
template <int JT>
__device__ __forceinline__ void deep_jet0(const float (&av)[32], const float* wr, float (&x)[32]) {
  using namespace tc;
#pragma unroll
  for (int u = 0; u < 32; u += 2) {
    const f32x2 a = pack2(av[u], av[u + 1]);
    f32x2 o = a;
    if (JT > 0) {
      const f32x2 w = ld2(wr + u);
      const f32x2 a1 = mul2(fma2(mul2(a, splat2(-1.0f)), a, splat2(1.0f)), w);
      o = JT == 1 ? a1 : mul2(mul2(mul2(a, splat2(-2.0f)), a1), w);
    }
    unpack2(o, x[u], x[u + 1]);
  }
}
// the adjoint keeps the first-layer activations in shared memory ([unit pair][trajectory][2]: a warp reads 256 contiguous
// bytes) instead of 32 registers per thread: a0 points at (unit pair 16 uh, trajectory m)
__device__ __forceinline__ void deep_first_layer_smem(float* a0, const float* wr, const float* wt, const float* cbi, float r, float tprime) {
  using namespace tc;
  const f32x2 r2 = splat2(r), tp2 = splat2(tprime);
#pragma unroll
  for (int u = 0; u < 32; u += 2) {
    float x0, x1;
    unpack2(deep_tanh2(fma2(ld2(wr + u), r2, fma2(tp2, ld2(wt + u), ld2(cbi + u)))), x0, x1);
    *reinterpret_cast<float2*>(a0 + (u >> 1) * (2 * kTP)) = make_float2(x0, x1);
  }
}
template <int JT>
__device__ __forceinline__ void deep_jet0_smem(const float* a0, const float* wr, float (&x)[32]) {
  using namespace tc;
#pragma unroll
  for (int u = 0; u < 32; u += 2) {
    const f32x2 a = ld2(a0 + (u >> 1) * (2 * kTP));
    f32x2 o = a;
    if (JT > 0) {
      const f32x2 w = ld2(wr + u);
      const f32x2 a1 = mul2(fma2(mul2(a, splat2(-1.0f)), a, splat2(1.0f)), w);
      o = JT == 1 ? a1 : mul2(mul2(mul2(a, splat2(-2.0f)), a1), w);
    }
    unpack2(o, x[u], x[u + 1]);
  }
}
// jet jt of hidden layer l >= 1 from its tensor-memory state (a, P', P''), units 32 uh .. of trajectory m
template <int JT>
__device__ __forceinline__ void deep_jet_state(uint32_t tstate, float (&x)[32]) {
  using namespace tc;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 16) {
    float a[16], p1[16], p2[16];
    tmem_ld16(tstate + c0, a);
    if (JT >= 1) tmem_ld16(tstate + kH + c0, p1);
    if (JT == 2) tmem_ld16(tstate + 2 * kH + c0, p2);
    tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 16; u += 2) {
      const f32x2 aa = pack2(a[u], a[u + 1]);
      f32x2 o = aa;
      if (JT > 0) {
        const f32x2 naa = mul2(aa, splat2(-1.0f));
        const f32x2 sg = fma2(naa, aa, splat2(1.0f));
        const f32x2 q1 = pack2(p1[u], p1[u + 1]);
        const f32x2 a1 = mul2(sg, q1);
        o = a1;
        if (JT == 2) o = fma2(sg, pack2(p2[u], p2[u + 1]), mul2(mul2(mul2(naa, splat2(2.0f)), a1), q1));   // sg p2 - 2 a a1 p1
      }
      unpack2(o, x[c0 + u], x[c0 + u + 1]);
    }
  }
}

// =================================================================================================================
// forward integrator
// =================================================================================================================
constexpr int kDFColAhi = 192;     // forward: accumulators of the three jets at columns 0..191, two TS-mode operand buffers behind them

template <int NJ, int L>
__device__ void deep_fwd_tile(SmemDeepF& sm, const GnnFwdArgs& a, long long row0, long long seg_end, float b3, uint32_t tmem,
                              uint32_t& phase) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m = 32 * (w & 3) + lane, uh = w >> 2;
  const bool owner = tid < kTP;
  const long long row = row0 + tid;
  const bool valid = owner && row < seg_end;
  const long long rowc = row < seg_end ? row : seg_end - 1;
  float y[7] = {0, 0, 0, 0, 0, 0, 0}, accw[7] = {0, 0, 0, 0, 0, 0, 0}, kst[7] = {0, 0, 0, 0, 0, 0, 0};
  float xst[3] = {0, 0, 0}, sst[3] = {0, 0, 0};
  float g[3] = {0, 0, 0}, dv = 0.f, ds[3] = {0, 0, 0};
  if (owner) {
    const float* src = a.y_in + rowc * a.in_stride;
    y[0] = src[0]; y[1] = src[1]; y[2] = src[2];
    if (NJ >= 2) y[3] = src[3];
    if (NJ == 3) { y[4] = src[4]; y[5] = src[5]; y[6] = src[6]; }
  }
  const int n_steps = a.rhs_only ? 1 : a.n_steps;
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  const uint32_t idesc = tc::make_idesc_tf32(kTP, kH, 0, 0);
  const uint32_t tlane = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  const uint64_t dAlo[2] = {tc::make_desc(tc::smem_u32(sm.Alo), kDChunkA, 128), tc::make_desc(tc::smem_u32(sm.Alo + kH * kTP), kDChunkA, 128)};

  auto combine = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    float f0 = b3 + sm.Fpart[0 * kTP + tid] + sm.Fpart[3 * kTP + tid], f1 = 0.f, f2 = 0.f;
    if (NJ >= 2) f1 = sm.Fpart[1 * kTP + tid] + sm.Fpart[4 * kTP + tid];
    if (NJ == 3) f2 = sm.Fpart[2 * kTP + tid] + sm.Fpart[5 * kTP + tid];
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
#pragma unroll 1
      for (int i = 0; i < a.Na; ++i) {
        float J0[32], J1[32], J2[32];
        {
          const float* nc = sm.T.nuc + i * 4;
          const float d0 = sm.xs[m] - nc[0], d1 = sm.xs[kTP + m] - nc[1], d2 = sm.xs[2 * kTP + m] - nc[2];
          const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
          deep_first_layer(J0, sm.wr + 32 * uh, sm.wt + 32 * uh, sm.T.cb + sm.T.cls[i] * kH + 32 * uh, r, tprime);
          if (NJ == 3) deep_jet0<2>(J0, sm.wr + 32 * uh, J2);
          if (NJ >= 2) deep_jet0<1>(J0, sm.wr + 32 * uh, J1);
        }
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          const uint64_t dWh = tc::make_desc(tc::smem_u32(sm.Whi + (l - 1) * kH * kH), kDChunkB, 128);
          const uint64_t dWl = tc::make_desc(tc::smem_u32(sm.Wlo + (l - 1) * kH * kH), kDChunkB, 128);
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {      // the staging of jet jt + 1 runs under the MMAs of jet jt (two operand buffers)
            const int buf = jt & 1;
            if (jt >= 2) deep_wait(sm.bar, phase, buf);
            deep_stage_A(jt == 0 ? J0 : (jt == 1 ? J1 : J2), tlane + kDFColAhi + kH * buf + 32 * uh, sm.Alo + buf * kH * kTP, uh, m);
            deep_sync_issue_gemm(tmem + jt * kH, tmem + kDFColAhi + kH * buf, dAlo[buf], dWh, dWl, idesc, &sm.bar[buf]);
          }
          const float* bl = sm.bl + (l - 1) * kH + 32 * uh;
          if constexpr (NJ == 3) {
            // jet 0 has completed (its barrier was waited for in front of the staging of jet 2): the tanh of the layer runs
            // under the MMAs of jets 1 and 2, the derivative jets follow as their accumulators complete
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float p0[16];
              tc::tmem_ld16(tlane + 32 * uh + c0, p0);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2)
                tc::unpack2(deep_tanh2(tc::add2(tc::pack2(p0[u], p0[u + 1]), ld2(bl + c0 + u))), J0[c0 + u], J0[c0 + u + 1]);
            }
            deep_wait(sm.bar, phase, 1);
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float p1[16];
              tc::tmem_ld16(tlane + kH + 32 * uh + c0, p1);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2) {
                using namespace tc;
                const f32x2 aa = pack2(J0[c0 + u], J0[c0 + u + 1]);
                unpack2(mul2(fma2(mul2(aa, splat2(-1.0f)), aa, splat2(1.0f)), pack2(p1[u], p1[u + 1])), J1[c0 + u], J1[c0 + u + 1]);
              }
            }
            deep_wait(sm.bar, phase, 0);
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float p1[16], p2[16];
              tc::tmem_ld16(tlane + kH + 32 * uh + c0, p1);
              tc::tmem_ld16(tlane + 2 * kH + 32 * uh + c0, p2);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2) {
                using namespace tc;
                const f32x2 aa = pack2(J0[c0 + u], J0[c0 + u + 1]);
                const f32x2 naa = mul2(aa, splat2(-1.0f));
                const f32x2 sg = fma2(naa, aa, splat2(1.0f));
                const f32x2 a1 = pack2(J1[c0 + u], J1[c0 + u + 1]);
                // sg p2 - 2 a a1 p1
                unpack2(fma2(sg, pack2(p2[u], p2[u + 1]), mul2(mul2(mul2(naa, splat2(2.0f)), a1), pack2(p1[u], p1[u + 1]))), J2[c0 + u], J2[c0 + u + 1]);
              }
            }
          } else {
#pragma unroll
          for (int jt = (NJ >= 2 ? NJ - 2 : 0); jt < NJ; ++jt) deep_wait(sm.bar, phase, jt & 1);
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float p0[16], p1[16], p2[16];
            tc::tmem_ld16(tlane + 0 * kH + 32 * uh + c0, p0);
            if (NJ >= 2) tc::tmem_ld16(tlane + 1 * kH + 32 * uh + c0, p1);
            if (NJ == 3) tc::tmem_ld16(tlane + 2 * kH + 32 * uh + c0, p2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              using namespace tc;
              const f32x2 aa = deep_tanh2(add2(pack2(p0[u], p0[u + 1]), ld2(bl + c0 + u)));
              unpack2(aa, J0[c0 + u], J0[c0 + u + 1]);
              if (NJ >= 2) {
                const f32x2 naa = mul2(aa, splat2(-1.0f));
                const f32x2 sg = fma2(naa, aa, splat2(1.0f));
                const f32x2 q1 = pack2(p1[u], p1[u + 1]);
                const f32x2 a1 = mul2(sg, q1);
                unpack2(a1, J1[c0 + u], J1[c0 + u + 1]);
                if (NJ == 3)
                  unpack2(fma2(sg, pack2(p2[u], p2[u + 1]), mul2(mul2(mul2(naa, splat2(2.0f)), a1), q1)), J2[c0 + u], J2[c0 + u + 1]);
              }
            }
          }
          }
        }
        {
          tc::f32x2 fpp[3] = {0ull, 0ull, 0ull};
#pragma unroll
          for (int u = 0; u < 32; u += 2) {
            const tc::f32x2 wv = ld2(sm.w3 + 32 * uh + u);
            fpp[0] = tc::fma2(wv, tc::pack2(J0[u], J0[u + 1]), fpp[0]);
            if (NJ >= 2) fpp[1] = tc::fma2(wv, tc::pack2(J1[u], J1[u + 1]), fpp[1]);
            if (NJ == 3) fpp[2] = tc::fma2(wv, tc::pack2(J2[u], J2[u + 1]), fpp[2]);
          }
          const float f0 = tc::lo2(fpp[0]) + tc::hi2(fpp[0]), f1 = tc::lo2(fpp[1]) + tc::hi2(fpp[1]), f2 = tc::lo2(fpp[2]) + tc::hi2(fpp[2]);
          sm.Fpart[(uh * 3 + 0) * kTP + m] = f0;
          if (NJ >= 2) sm.Fpart[(uh * 3 + 1) * kTP + m] = f1;
          if (NJ == 3) sm.Fpart[(uh * 3 + 2) * kTP + m] = f2;
        }
        tc::fence_before_sync();
        __syncthreads();
        if (owner) combine(i);
      }
      if (owner) {
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

template <int L>
__global__ void __launch_bounds__(kThreads, 1) gnn_deep_fwd_tc_kernel(const GnnFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemDeepF& sm = *reinterpret_cast<SmemDeepF*>(smem_raw);
  const int tid = threadIdx.x;
  const GnnThetaLayoutL TL(a.n_types, L);
  const GnnThetaLayout L2(a.n_types);       // b1 / W1 offsets do not depend on the depth
  for (int l = 1; l < L; ++l)
    for (int i = tid; i < kH * kH; i += kThreads) {
      const int k = i / kH, j = i % kH;
      float hi, lo;
      tc::split_tf32(a.theta[TL.W[l] + i], hi, lo);
      const int off = (l - 1) * kH * kH + ((k >> 2) * kH + j) * 4 + (k & 3);
      sm.Whi[off] = hi; sm.Wlo[off] = lo;
    }
  if (tid < kH) {
    sm.wt[tid] = a.theta[TL.W[0] + tid];
    sm.wr[tid] = a.theta[TL.W[0] + kH + tid];
    sm.w3[tid] = a.theta[TL.w3 + tid];
    for (int l = 1; l < L; ++l) sm.bl[(l - 1) * kH + tid] = a.theta[TL.b[l] + tid];
  }
  load_nuc_tables<kThreads>(sm.T, a.theta, L2, a.nuclei, a.onehot, a.Na, a.n_types);
  if (tid == 0) { tc::mbar_init(&sm.bar[0], 1); tc::mbar_init(&sm.bar[1], 1); tc::mbar_init(&sm.bar[2], 1); tc::fence_mbar_init(); }
  if (tid < 32) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  uint32_t phase = 0;
  const float b3 = a.theta[0];
  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; }
    else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; }
    if (nj == 3) deep_fwd_tile<3, L>(sm, a, row0, seg_end, b3, tmem, phase);
    else if (nj == 2) deep_fwd_tile<2, L>(sm, a, row0, seg_end, b3, tmem, phase);
    else deep_fwd_tile<1, L>(sm, a, row0, seg_end, b3, tmem, phase);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 512);
}

// =================================================================================================================
// discrete adjoint
// =================================================================================================================
template <int L>
struct DeepAccum {
  float g_w3, g_wt, g_wr, g_b3;   // lane <-> unit 32 uh + lane, summed over the warp's 32 trajectories (g_b3: per trajectory)
  float g_b[L];                   // bias gradients of the hidden layers l >= 1 (slot 0 unused)
  __device__ void zero() {
    g_w3 = g_wt = g_wr = g_b3 = 0.f;
#pragma unroll
    for (int l = 0; l < L; ++l) g_b[l] = 0.f;
  }
};

// reverse of the tanh jet of a hidden layer l >= 1 for two units, in place: (c0, c1, c2) = cotangents of (a, a', a'') ->
// cotangents of the pre-activation jets (P, P', P''); state = (a, P', P'')
template <int NJ>
__device__ __forceinline__ void deep_tanh_jet_reverse(tc::f32x2 aa, tc::f32x2 p1, tc::f32x2 p2, tc::f32x2& c0, tc::f32x2& c1, tc::f32x2& c2) {
  using namespace tc;
  const f32x2 one = splat2(1.0f), two = splat2(2.0f);
  const f32x2 naa = mul2(aa, splat2(-1.0f));            // -a
  const f32x2 sg = fma2(naa, aa, one);                  // 1 - a^2
  f32x2 pzb = c0;
  if (NJ >= 2) {
    const f32x2 nap1 = mul2(naa, p1);                   // -a p1
    pzb = fma2(mul2(nap1, two), c1, pzb);               // - 2 a p1 c1
    f32x2 t1 = c1;
    if (NJ == 3) {
      t1 = fma2(mul2(nap1, splat2(4.0f)), c2, t1);      // c1 - 4 a p1 c2
      const f32x2 c3 = fma2(mul2(aa, splat2(-3.0f)), aa, one);                                   // 1 - 3 a^2
      const f32x2 inner = fma2(mul2(naa, two), p2, mul2(mul2(c3, splat2(-2.0f)), mul2(p1, p1)));   // -2 a p2 - 2 (1 - 3 a^2) p1^2
      pzb = fma2(c2, inner, pzb);
      c2 = mul2(c2, sg);
    }
    c1 = mul2(sg, t1);
  }
  c0 = mul2(sg, pzb);
}

template <int NJ, int L>
__device__ void deep_bwd_tile(SmemDeepB& sm, const GnnBwdArgs& a, long long row0, long long seg_end, float b3, DeepAccum<L>& G,
                              uint32_t tmem, uint32_t& phase) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m = 32 * (w & 3) + lane, uh = w >> 2;
  const bool owner = uh == 0;
  const long long row = row0 + m;
  const bool valid = row < seg_end;
  const long long rowc = row < seg_end ? row : seg_end - 1;
  float xb[3] = {0, 0, 0}, lb = 0.f, sb[3] = {0, 0, 0};
  if (valid) {
    const float* src = a.ybar1 + row * a.bar_stride;
    xb[0] = src[0]; xb[1] = src[1]; xb[2] = src[2];
    if (NJ >= 2) lb = src[3];
    if (NJ == 3) { sb[0] = src[4]; sb[1] = src[5]; sb[2] = src[6]; }
  }
  const float h = (a.t1 - a.t0) / (float)a.n_steps;
  float xst[3] = {0, 0, 0}, sst[3] = {0, 0, 0};
  float kx[3] = {0, 0, 0}, kl = 0.f, ks[3] = {0, 0, 0};
  float ysx[3] = {0, 0, 0}, yss[3] = {0, 0, 0};
  float sumx[3], sums[3], prevx[3] = {0, 0, 0}, prevs[3] = {0, 0, 0};
  float fb0 = 0.f, fb1 = 0.f, fb2 = 0.f;
  const uint32_t tlane = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  const uint32_t idesc = tc::make_idesc_tf32(kTP, kH, 0, 0);
  const uint32_t idesc_dw = tc::make_idesc_tf32(kTP, 2 * kH, 0, 0);
  float* const Alo = sm.R;
  float* const Wop = sm.R + kDWopOff;
  float* const AT = sm.R;
  float* const PT = sm.R + 32 * kDTChunk;
  const uint64_t dAlo[2] = {tc::make_desc(tc::smem_u32(Alo), kDChunkA, 128), tc::make_desc(tc::smem_u32(Alo + kH * kTP), kDChunkA, 128)};
  const uint64_t dWh[2] = {tc::make_desc(tc::smem_u32(Wop), kDChunkB, 128), tc::make_desc(tc::smem_u32(Wop + 2 * kH * kH), kDChunkB, 128)};
  const uint64_t dWl[2] = {tc::make_desc(tc::smem_u32(Wop + kH * kH), kDChunkB, 128), tc::make_desc(tc::smem_u32(Wop + 3 * kH * kH), kDChunkB, 128)};
  int wb = 0;                 // weight operand buffer of the current GEMM phase (the next phase's operand is copied into the other one)
  const uint64_t dAT = tc::make_desc(tc::smem_u32(AT), kDTChunk * 4, 128);
  const uint64_t dPT = tc::make_desc(tc::smem_u32(PT), kDTChunk * 4, 128);
  const float* const wr = sm.wr + 32 * uh;
  float* const a0 = sm.A0 + (16 * uh * kTP + m) * 2;

  // cotangents of (f, f', f'') of nucleus i   [pre() of gnn.cu]
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
    fb0 = f0b * a.y0; fb1 = f1b * a.y0; fb2 = f2b * a.y0;
    if (!valid) { fb0 = 0.f; fb1 = 0.f; fb2 = 0.f; }
    if (owner) G.g_b3 += fb0;
  };
  // explicit geometry   [post() of gnn.cu]
  auto post = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    const float* Fp = sm.Fpart;
    float f0 = b3 + Fp[0 * kTP + m] + Fp[3 * kTP + m], f1 = 0.f, f2 = 0.f;
    if (NJ >= 2) f1 = Fp[1 * kTP + m] + Fp[4 * kTP + m];
    if (NJ == 3) f2 = Fp[2 * kTP + m] + Fp[5 * kTP + m];
    float rbar = sm.Rpart[m] + sm.Rpart[kTP + m];
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
      const float ua = -a.y0 * amp;
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
  // hi | lo weight operand of hidden layer l (orient 0: forward, B[n = j][k]; 1: transposed, B[n = k][j]) -> staging region
  // (asynchronous copies: cp.async, completed by the wait in front of the next MMA hand-off)
  auto load_W = [&](int l, int orient, int buf) {
    const float4* src = reinterpret_cast<const float4*>(a.w2split + (size_t)((l - 1) * 4 + 2 * orient) * kH * kH) + tid;
    const uint32_t dst = tc::smem_u32(Wop + buf * 2 * kH * kH) + tid * 16;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * kThreads * 16), "l"(src + q * kThreads) : "memory");
  };
  load_W(1, 0, wb);           // forward operand of the first hidden layer for the first nucleus

  for (int step = a.n_steps - 1; step >= 0; --step) {
    sumx[0] = sumx[1] = sumx[2] = 0.f; sums[0] = sums[1] = sums[2] = 0.f;
#pragma unroll 1
    for (int stage = 3; stage >= 0; --stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float tprime = a.y0 * (a.t0 + h * (float)step + cs);
      {
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
          if (NJ == 3) sst[c] = ck[(long long)(3 + c) * a.B];
        }
        ysx[0] = ysx[1] = ysx[2] = 0.f; yss[0] = yss[1] = yss[2] = 0.f;
      }
#pragma unroll 1
      for (int i = 0; i < a.Na; ++i) {
        pre(i);
        float r;
        {
          const float* nc = sm.T.nuc + i * 4;
          const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
          r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
          deep_first_layer_smem(a0, wr, sm.wt + 32 * uh, sm.T.cb + sm.T.cls[i] * kH + 32 * uh, r, tprime);
        }
        // ---------------- forward recompute: S(l) <- (a_l, P_l', P_l'') for l = 1 .. L-1 ----------------
#pragma unroll 1
        for (int l = 1; l < L; ++l) {
          // the next GEMM phase's weight operand (forward of layer l + 1, or transposed of the top layer) into the other buffer
          if (l + 1 < L) load_W(l + 1, 0, wb ^ 1); else load_W(L - 1, 1, wb ^ 1);
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {      // the staging of jet jt + 1 runs under the MMAs of jet jt (two operand buffers)
            const int buf = jt & 1;
            if (jt >= 2) deep_wait(sm.bar, phase, buf);
            float x[32];
            if (l == 1) { if (jt == 0) deep_jet0_smem<0>(a0, wr, x); else if (jt == 1) deep_jet0_smem<1>(a0, wr, x); else deep_jet0_smem<2>(a0, wr, x); }
            else {
              const uint32_t ts = tlane + dcol_state(l - 1) + 32 * uh;
              if (jt == 0) deep_jet_state<0>(ts, x); else if (jt == 1) deep_jet_state<1>(ts, x); else deep_jet_state<2>(ts, x);
            }
            deep_stage_A(x, tlane + kDColDw + kH * buf + 32 * uh, Alo + buf * kH * kTP, uh, m);
            deep_sync_issue_gemm(tmem + dcol_state(l) + jt * kH, tmem + kDColDw + kH * buf, dAlo[buf], dWh[wb], dWl[wb], idesc, &sm.bar[buf]);
          }
          wb ^= 1;
          // three jets: jet 0 has completed (waited for in front of the staging of jet 2), its tanh runs under the MMAs of jets 1, 2
          if (NJ < 3) {
#pragma unroll
            for (int jt = (NJ >= 2 ? NJ - 2 : 0); jt < NJ; ++jt) deep_wait(sm.bar, phase, jt & 1);
          }
          // P -> a = tanh(P + b_l) in place (the derivative jets stay as they are)
          const float* bl = sm.bl + (l - 1) * kH + 32 * uh;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float p0[16];
            tc::tmem_ld16(tlane + dcol_state(l) + 32 * uh + c0, p0);
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; u += 2) tc::unpack2(deep_tanh2(tc::add2(tc::pack2(p0[u], p0[u + 1]), ld2(bl + c0 + u))), p0[u], p0[u + 1]);
            tc::tmem_st16(tlane + dcol_state(l) + 32 * uh + c0, p0);
          }
          tc::tmem_st_wait();
          if (NJ == 3) { deep_wait(sm.bar, phase, 1); deep_wait(sm.bar, phase, 0); }
        }
        // ---------------- last layer and the reverse of the top hidden layer ----------------
        float q0[32], q1[32], q2[32];
        {
          tc::f32x2 fpp[3] = {0ull, 0ull, 0ull};
          float g32[32];
          const uint32_t ts = tlane + dcol_state(L - 1) + 32 * uh;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float s0[16], s1[16], s2[16];
            tc::tmem_ld16(ts + c0, s0);
            if (NJ >= 2) tc::tmem_ld16(ts + kH + c0, s1);
            if (NJ == 3) tc::tmem_ld16(ts + 2 * kH + c0, s2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              using namespace tc;
              const f32x2 wv = ld2(sm.w3 + 32 * uh + c0 + u);
              const f32x2 aa = pack2(s0[u], s0[u + 1]);
              const f32x2 sg = fma2(mul2(aa, splat2(-1.0f)), aa, splat2(1.0f));
              fpp[0] = fma2(wv, aa, fpp[0]);
              f32x2 gw = mul2(aa, splat2(fb0));
              f32x2 c0v = mul2(wv, splat2(fb0)), c1v = 0ull, c2v = 0ull, p1 = 0ull, p2 = 0ull;
              if (NJ >= 2) {
                p1 = pack2(s1[u], s1[u + 1]);
                const f32x2 a1 = mul2(sg, p1);
                fpp[1] = fma2(wv, a1, fpp[1]);
                gw = fma2(a1, splat2(fb1), gw);
                c1v = mul2(wv, splat2(fb1));
                if (NJ == 3) {
                  p2 = pack2(s2[u], s2[u + 1]);
                  const f32x2 a2 = fma2(sg, p2, mul2(mul2(mul2(aa, splat2(-2.0f)), a1), p1));
                  fpp[2] = fma2(wv, a2, fpp[2]);
                  gw = fma2(a2, splat2(fb2), gw);
                  c2v = mul2(wv, splat2(fb2));
                }
              }
              deep_tanh_jet_reverse<NJ>(aa, p1, p2, c0v, c1v, c2v);
              unpack2(c0v, q0[c0 + u], q0[c0 + u + 1]); unpack2(c1v, q1[c0 + u], q1[c0 + u + 1]); unpack2(c2v, q2[c0 + u], q2[c0 + u + 1]);
              unpack2(gw, g32[c0 + u], g32[c0 + u + 1]);
            }
          }
          const float f0 = tc::lo2(fpp[0]) + tc::hi2(fpp[0]), f1 = tc::lo2(fpp[1]) + tc::hi2(fpp[1]), f2 = tc::lo2(fpp[2]) + tc::hi2(fpp[2]);
          sm.Fpart[(uh * 3 + 0) * kTP + m] = f0;
          if (NJ >= 2) sm.Fpart[(uh * 3 + 1) * kTP + m] = f1;
          if (NJ == 3) sm.Fpart[(uh * 3 + 2) * kTP + m] = f2;
          G.g_w3 += warp_vec32_sum(g32, lane);
        }
        // ---------------- hidden layers L-1 .. 1 ----------------
#pragma unroll 1
        for (int l = L - 1; l >= 1; --l) {      // a real loop: one copy of the layer body keeps the code inside the instruction cache
          // Abar_{l-1}[jet] = Pbar_l[jet] . W_l^T  -> the (dead) state columns of layer l; the top layer's transposed operand was
          // copied under the forward GEMM, the others come in here (the weight-gradient operands overwrite the whole region)
          if (l < L - 1) load_W(l, 1, wb);
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {
            const int buf = jt & 1;
            if (jt >= 2) deep_wait(sm.bar, phase, buf);
            deep_stage_A(jt == 0 ? q0 : (jt == 1 ? q1 : q2), tlane + kDColDw + kH * buf + 32 * uh, Alo + buf * kH * kTP, uh, m);
            deep_sync_issue_gemm(tmem + dcol_state(l) + jt * kH, tmem + kDColDw + kH * buf, dAlo[buf], dWh[wb], dWl[wb], idesc, &sm.bar[buf]);
          }
          {
            // bias gradient of layer l (the butterfly runs under the MMAs)
            float t[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) t[u] = q0[u];
            const float s = warp_vec32_sum(t, lane);
#pragma unroll
            for (int ll = 1; ll < L; ++ll) G.g_b[ll] += ll == l ? s : 0.f;
          }
#pragma unroll
          for (int jt = (NJ >= 2 ? NJ - 2 : 0); jt < NJ; ++jt) deep_wait(sm.bar, phase, jt & 1);
          // dW_l[k][j] += sum over trajectories and jets of A_{l-1}[jet][k] Pbar_l[jet][j]   (K = trajectory, stacked hi | lo rows)
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {
            if (jt > 0) deep_wait(sm.bar, phase, 2);
            float x[32];
            if (l == 1) { if (jt == 0) deep_jet0_smem<0>(a0, wr, x); else if (jt == 1) deep_jet0_smem<1>(a0, wr, x); else deep_jet0_smem<2>(a0, wr, x); }
            else {
              const uint32_t ts = tlane + dcol_state(l - 1) + 32 * uh;
              if (jt == 0) deep_jet_state<0>(ts, x); else if (jt == 1) deep_jet_state<1>(ts, x); else deep_jet_state<2>(ts, x);
            }
            const float (&p)[32] = jt == 0 ? q0 : (jt == 1 ? q1 : q2);
            const int cbase = (m >> 2) * kDTChunk, csub = m & 3;
#pragma unroll
            for (int u = 0; u < 32; u += 2) {
              const tc::f32x2 xl = lo2pk(tc::pack2(x[u], x[u + 1])), pl = lo2pk(tc::pack2(p[u], p[u + 1]));
              float* at = AT + cbase + (32 * uh + u) * 4 + csub;
              float* pt = PT + cbase + (32 * uh + u) * 4 + csub;
              at[0] = x[u]; at[4] = x[u + 1]; at[kH * 4] = tc::lo2(xl); at[kH * 4 + 4] = tc::hi2(xl);
              pt[0] = p[u]; pt[4] = p[u + 1]; pt[kH * 4] = tc::lo2(pl); pt[kH * 4 + 4] = tc::hi2(pl);
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (tid < 32) {
              tc::fence_after_sync();
              if (tc::elect_one()) {
#pragma unroll 1
                for (int ksx = 0; ksx < 16; ++ksx)
                  tc::mma_tf32(tmem + kDColDw, dAT + ksx * ((2 * kDTChunk * 4) >> 4), dPT + ksx * ((2 * kDTChunk * 4) >> 4), idesc_dw,
                               (jt | ksx) != 0);
                tc::commit(&sm.bar[2]);
              }
              __syncwarp();
            }
          }
          if (l > 1) {
            // reverse of the tanh jet of layer l-1: cotangent jets Abar_{l-1} (columns of layer l) and the state of layer l-1
            const uint32_t tb = tlane + dcol_state(l) + 32 * uh, ts = tlane + dcol_state(l - 1) + 32 * uh;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float b0[16], b1[16], b2[16], s0[16], s1[16], s2[16];
              tc::tmem_ld16(tb + c0, b0); tc::tmem_ld16(ts + c0, s0);
              if (NJ >= 2) { tc::tmem_ld16(tb + kH + c0, b1); tc::tmem_ld16(ts + kH + c0, s1); }
              if (NJ == 3) { tc::tmem_ld16(tb + 2 * kH + c0, b2); tc::tmem_ld16(ts + 2 * kH + c0, s2); }
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2) {
                using namespace tc;
                f32x2 c0v = pack2(b0[u], b0[u + 1]), c1v = 0ull, c2v = 0ull, p1 = 0ull, p2 = 0ull;
                if (NJ >= 2) { c1v = pack2(b1[u], b1[u + 1]); p1 = pack2(s1[u], s1[u + 1]); }
                if (NJ == 3) { c2v = pack2(b2[u], b2[u + 1]); p2 = pack2(s2[u], s2[u + 1]); }
                deep_tanh_jet_reverse<NJ>(pack2(s0[u], s0[u + 1]), p1, p2, c0v, c1v, c2v);
                unpack2(c0v, q0[c0 + u], q0[c0 + u + 1]); unpack2(c1v, q1[c0 + u], q1[c0 + u + 1]); unpack2(c2v, q2[c0 + u], q2[c0 + u + 1]);
              }
            }
          } else {
            // first layer reverse   [epilogue_b1 of gnn.cu]
            tc::f32x2 rp2 = 0ull;
            float ps32[32], pr32[32];
            const uint32_t tb = tlane + dcol_state(1) + 32 * uh;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float b0[16], b1[16], b2[16];
              tc::tmem_ld16(tb + c0, b0);
              if (NJ >= 2) tc::tmem_ld16(tb + kH + c0, b1);
              if (NJ == 3) tc::tmem_ld16(tb + 2 * kH + c0, b2);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2) {
                using namespace tc;
                const f32x2 one = splat2(1.0f);
                const f32x2 aa = ld2(a0 + ((c0 + u) >> 1) * (2 * kTP));
                const f32x2 wv = ld2(wr + c0 + u);
                const f32x2 aw = mul2(aa, wv);                                  // a w
                const f32x2 sg = fma2(mul2(aa, splat2(-1.0f)), aa, one);        // 1 - a^2
                f32x2 t = pack2(b0[u], b0[u + 1]);
                f32x2 e = 0ull;
                if (NJ >= 2) {
                  const f32x2 c1 = pack2(b1[u], b1[u + 1]);
                  t = fma2(mul2(aw, splat2(-2.0f)), c1, t);                     // - 2 a w b1
                  e = c1;
                  if (NJ == 3) {
                    const f32x2 c2 = pack2(b2[u], b2[u + 1]);
                    const f32x2 c3 = fma2(mul2(aa, splat2(-3.0f)), aa, one);    // 1 - 3 a^2
                    t = fma2(mul2(mul2(mul2(wv, wv), splat2(-2.0f)), c3), c2, t);   // - 2 w^2 (1 - 3 a^2) b2
                    e = fma2(mul2(aw, splat2(-4.0f)), c2, e);                   // b1 - 4 a w b2
                  }
                }
                const f32x2 pb = mul2(sg, t);
                rp2 = fma2(pb, wv, rp2);
                unpack2(pb, ps32[c0 + u], ps32[c0 + u + 1]);
                unpack2(fma2(pb, splat2(r), mul2(sg, e)), pr32[c0 + u], pr32[c0 + u + 1]);
              }
            }
            sm.Rpart[uh * kTP + m] = tc::lo2(rp2) + tc::hi2(rp2);
            const float ps = warp_vec32_sum(ps32, lane);
            const float pr = warp_vec32_sum(pr32, lane);
            sm.S[(w * kMaxClasses + sm.T.cls[i]) * 32 + lane] += ps;
            G.g_wt = fmaf(tprime, ps, G.g_wt);
            G.g_wr += pr;
          }
          // the last weight-gradient batch of the layer ran under the reverse epilogue above
          deep_wait(sm.bar, phase, 2);
          // fold the stacked tile into the fp32 sums: dW_l[k][j] = D[k][j] + D[k][64 + j] + D[64 + k][j] (+ the lo.lo corner)
          {
            float v[32];
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float t[16], t2[16];
              tc::tmem_ld16(tlane + kDColDw + 32 * uh + c0, t);
              tc::tmem_ld16(tlane + kDColDw + kH + 32 * uh + c0, t2);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; ++u) v[c0 + u] = t[u] + t2[u];
            }
            float* dws = sm.DWs + (l - 1) * kH * kH + (m & 63);
            if (m < kH) {
#pragma unroll
              for (int u = 0; u < 32; ++u) dws[(32 * uh + u) * kH] += v[u];
            }
            tc::fence_before_sync();
            __syncthreads();
            if (m >= kH) {
#pragma unroll
              for (int u = 0; u < 32; ++u) dws[(32 * uh + u) * kH] += v[u];
            }
          }
        }
        load_W(1, 0, wb);       // forward operand of the first hidden layer for the next nucleus (the region is free again)
        tc::fence_before_sync();
        __syncthreads();
        post(i);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        sumx[c] += ysx[c]; prevx[c] = ysx[c];
        sums[c] += yss[c]; prevs[c] = yss[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { xb[c] += sumx[c]; sb[c] += sums[c]; }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");     // the operand copy issued for a next nucleus that does not exist
  if (owner && valid && a.ybar0 != nullptr) {
    float* dst = a.ybar0 + row * a.bar_stride;
    dst[0] = xb[0]; dst[1] = xb[1]; dst[2] = xb[2];
    if (NJ >= 2) dst[3] = lb;
    if (NJ == 3) { dst[4] = sb[0]; dst[5] = sb[1]; dst[6] = sb[2]; }
  }
}

// per-tile partial gradient (same layout and reduction kernel as the other paths)
template <int L>
__device__ void deep_flush_tile(SmemDeepB& sm, DeepAccum<L>& G, float* __restrict__ out, const GnnThetaLayoutL& TL,
                                const float* __restrict__ onehot, int n_types) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  sm.red[(0 * 8 + w) * 32 + lane] = G.g_w3; sm.red[(1 * 8 + w) * 32 + lane] = G.g_wt; sm.red[(2 * 8 + w) * 32 + lane] = G.g_wr;
#pragma unroll
  for (int l = 1; l < L; ++l) sm.red[((2 + l) * 8 + w) * 32 + lane] = G.g_b[l];
  const float b3w = warp_sum(G.g_b3);
  G.zero();
  __syncthreads();
  for (int l = 1; l < L; ++l)
    for (int i = tid; i < kH * kH; i += kThreads) {
      const int k = i / kH, j = i % kH;
      out[TL.W[l] + i] = sm.DWs[(l - 1) * kH * kH + j * kH + k];
    }
  if (tid < kH) {
    // unit tid: uh = tid >> 5, lane = tid & 31; the four warps uh*4 .. uh*4+3 hold partial sums over trajectories
    const int u2 = tid >> 5, l2 = tid & 31;
    float v[2 + kMaxLayers + 1];
    for (int q = 0; q < 2 + L; ++q) {
      v[q] = 0.f;
      for (int ww = 0; ww < 4; ++ww) v[q] += sm.red[(q * 8 + u2 * 4 + ww) * 32 + l2];
    }
    out[TL.w3 + tid] = v[0]; out[TL.W[0] + tid] = v[1]; out[TL.W[0] + kH + tid] = v[2];
    for (int l = 1; l < L; ++l) out[TL.b[l] + tid] = v[2 + l];
    float sb1 = 0.f, st[kMaxTypes];
    for (int t = 0; t < kMaxTypes; ++t) st[t] = 0.f;
    for (int c = 0; c < sm.T.n_cls; ++c) {
      float sn = 0.f;
      for (int ww = 0; ww < 4; ++ww) sn += sm.S[((u2 * 4 + ww) * kMaxClasses + c) * 32 + l2];
      sb1 += sn;
      for (int t = 0; t < n_types; ++t) st[t] = fmaf(onehot[sm.T.rep[c] * n_types + t], sn, st[t]);
    }
    out[TL.b[0] + tid] = sb1;
    for (int t = 0; t < n_types; ++t) out[TL.W[0] + (2 + t) * kH + tid] = st[t];
  }
  __syncthreads();
  if (lane == 0) sm.red[w] = b3w;
  __syncthreads();
  if (tid == 0) out[TL.b3] = sm.red[0] + sm.red[1] + sm.red[2] + sm.red[3];
  for (int i = tid; i < 8 * kMaxClasses * 32; i += kThreads) sm.S[i] = 0.f;
  for (int i = tid; i < (L - 1) * kH * kH; i += kThreads) sm.DWs[i] = 0.f;
  __syncthreads();
}

template <int L>
__global__ void __launch_bounds__(kThreads, 1) gnn_deep_bwd_tc_kernel(const GnnBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemDeepB& sm = *reinterpret_cast<SmemDeepB*>(smem_raw);
  const int tid = threadIdx.x;
  const GnnThetaLayoutL TL(a.n_types, L);
  const GnnThetaLayout L2(a.n_types);
  if (tid < kH) {
    sm.wt[tid] = a.theta[TL.W[0] + tid];
    sm.wr[tid] = a.theta[TL.W[0] + kH + tid];
    sm.w3[tid] = a.theta[TL.w3 + tid];
    for (int l = 1; l < L; ++l) sm.bl[(l - 1) * kH + tid] = a.theta[TL.b[l] + tid];
  }
  load_nuc_tables<kThreads>(sm.T, a.theta, L2, a.nuclei, a.onehot, a.Na, a.n_types);
  for (int i = tid; i < 8 * kMaxClasses * 32; i += kThreads) sm.S[i] = 0.f;
  for (int i = tid; i < (kMaxLayers - 1) * kH * kH; i += kThreads) sm.DWs[i] = 0.f;
  if (tid == 0) { tc::mbar_init(&sm.bar[0], 1); tc::mbar_init(&sm.bar[1], 1); tc::mbar_init(&sm.bar[2], 1); tc::fence_mbar_init(); }
  if (tid < 32) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  uint32_t phase = 0;
  const float b3 = a.theta[0];
  DeepAccum<L> G;
  G.zero();
  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; }
    else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; }
    if (nj == 3) deep_bwd_tile<3, L>(sm, a, row0, seg_end, b3, G, tmem, phase);
    else if (nj == 2) deep_bwd_tile<2, L>(sm, a, row0, seg_end, b3, G, tmem, phase);
    else deep_bwd_tile<1, L>(sm, a, row0, seg_end, b3, G, tmem, phase);
    deep_flush_tile<L>(sm, G, a.gpart + (size_t)tile * TL.n, TL, a.onehot, a.n_types);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 512);
}

// hidden-layer weights split hi / lo in both GEMM orientations: [layer - 1][fwd hi | fwd lo | transposed hi | transposed lo],
// forward B[n = j][k] as [k/4][j][k%4], transposed B[n = k][j] as [j/4][k][j%4]
__global__ void gnn_deep_wsplit_kernel(const float* __restrict__ theta, float* __restrict__ out, int n_types, int L) {
  const GnnThetaLayoutL TL(n_types, L);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = 1 + blockIdx.y;
  if (i >= kH * kH || l >= L) return;
  const int k = i / kH, j = i % kH;
  float hi, lo;
  tc::split_tf32(theta[TL.W[l] + i], hi, lo);
  float* o = out + (size_t)(l - 1) * 4 * kH * kH;
  const int of = ((k >> 2) * kH + j) * 4 + (k & 3), ot = ((j >> 2) * kH + k) * 4 + (j & 3);
  o[of] = hi; o[kH * kH + of] = lo; o[2 * kH * kH + ot] = hi; o[3 * kH * kH + ot] = lo;
}

int launch_gnn_deep_fwd_tc(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes, int n_layers) {
  if (scratch_bytes < kCounterBytes) return OFDFT_ESCRATCH;
  if (n_layers != 2 && n_layers != 3) return OFDFT_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  a.counter = (int*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (a.n_a + kTP - 1) / kTP + (a.B - a.n_a + kTP - 1) / kTP;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto kern = n_layers == 2 ? gnn_deep_fwd_tc_kernel<2> : gnn_deep_fwd_tc_kernel<3>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDeepF) + 1024);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)(tiles < sms ? tiles : sms);
  kern<<<grid, kThreads, sizeof(SmemDeepF) + 1024, st>>>(a);
  return (int)cudaGetLastError();
}

// the caller has reset the tile counter, provides kDeepWSplitBytes at a.w2split and reduces gpart afterwards
int launch_gnn_deep_bwd_tc(void* stream, GnnBwdArgs& a, int grid, int n_layers) {
  if (n_layers != 2 && n_layers != 3) return OFDFT_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  auto kern = n_layers == 2 ? gnn_deep_bwd_tc_kernel<2> : gnn_deep_bwd_tc_kernel<3>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDeepB) + 1024);
  if (e != cudaSuccess) return (int)e;
  gnn_deep_wsplit_kernel<<<dim3((kH * kH + 255) / 256, n_layers - 1), 256, 0, st>>>(a.theta, a.w2split, a.n_types, n_layers);
  kern<<<grid, kThreads, sizeof(SmemDeepB) + 1024, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace ofdft
