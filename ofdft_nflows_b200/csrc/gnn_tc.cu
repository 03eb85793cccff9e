// gnn_tc.cu -- tensor-core (tcgen05 / TMEM) kernels of the 3-D equivariant GNN flow: forward integrator and discrete adjoint.
//
// Same contract and results as the FP32-pipe kernels of gnn.cu; every [128 trajectories x jet] x [64 x 64] GEMM of a
// (tile, nucleus, RHS evaluation) runs on the 5th-generation tensor cores as error-compensated 3xTF32:
//     a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo      (fp32 accumulate in TMEM; the tensor core TRUNCATES fp32 operands to
//     tf32, so an unrounded value serves as its own "hi" part and lo = x - trunc(x))
// which holds fp32-level accuracy (measured 1.6e-6 max-relative on the diagnostic GEMM, tests/tc_diag.py).
//
//   * thread (warp w, lane l) owns trajectory m = 32*(w&3) + l -- the TMEM lane its warp may access -- and the unit half
//     uh = w>>2 (32 hidden units);
//   * operands are produced on chip by the threads themselves (no TMA: nothing is fetched from HBM but the 16 KB of
//     weights): the "hi" A operand goes to tensor memory (tcgen05.st, TS-mode MMA), remainders and the K = trajectory
//     operands of the weight gradient to shared memory in the canonical no-swizzle K-major UMMA layout
//     ([16-byte chunk of 4 k][row][4 floats]: SBO = 128 B, LBO = rows*16 B);
//   * forward: 256 threads, two co-resident CTAs per SM (one tile's staging / epilogue runs under the other's MMAs);
//   * backward: one CTA per SM, 256 compute threads + a dedicated MMA-issue warpgroup (setmaxnreg moves its registers to
//     the compute warpgroups), hand-offs by arrive-only named barriers, completions by mbarriers (see bwd_mma_tile).
#include <cstdlib>

#include "gnn_common.cuh"
#include "tc.cuh"

namespace ofdft {

constexpr int kChunkBytesA = kTP * 16;      // LBO of the activation operand (128 rows x 16 B)
constexpr int kChunkBytesB = kH * 16;       // LBO of the weight operand (64 rows x 16 B)
constexpr int kColAhi = 3 * kH;             // forward, two-CTA variant: TMEM columns of the activation "hi" operand

struct SmemTc1 {
  float Alo[kH * kTP];               // tf32 remainder of the jet being staged: [16 chunks][128 rows][4] (the "hi" part is in TMEM)
  float Whi[kH * kH], Wlo[kH * kH];  // B operand: [16 chunks of k][64 rows j][4]
  float wr[kH], wt[kH], b2[kH], w3[kH];
  NucTables T;
  float xs[3 * kTP];
  float Fpart[4 * 3 * kTP];
  uint64_t bar[3];
  unsigned int tmem_slot;
  int tile;
};
// tanh of two values: 1 - 2 / (2^(2 x log2 e) + 1) with the FP32 steps packed (one MUFU.EX2 + MUFU.RCP per value)
__device__ __forceinline__ tc::f32x2 tanh2(tc::f32x2 x) {
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

template <int NJ, int NU>
__device__ void fwd_tile_tc(SmemTc1& sm, const GnnFwdArgs& a, long long row0, long long seg_end, float b3,
                            bool seg_b, long long seg_row0, uint32_t tmem, uint32_t (&phase)[3]) {
  const ActLayout AL(a.act_rows_a, a.act_rows_b, a.jets_a, a.jets_b);
  const long long act_rows = seg_b ? a.act_rows_b : a.act_rows_a;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  constexpr int UPT = kH / NU;                      // hidden units per thread
  const int m = 32 * (w & 3) + lane, uh = w >> 2;   // uh: unit group 0..NU-1
  const bool owner = tid < kTP;                     // tid == m for w < 4
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
  const uint64_t dAlo = tc::make_desc(tc::smem_u32(sm.Alo), kChunkBytesA, 128);
  const uint64_t dWh = tc::make_desc(tc::smem_u32(sm.Whi), kChunkBytesB, 128);
  const uint64_t dWl = tc::make_desc(tc::smem_u32(sm.Wlo), kChunkBytesB, 128);

  auto combine = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    float f0 = b3, f1 = 0.f, f2 = 0.f;
#pragma unroll
    for (int q = 0; q < NU; ++q) {
      f0 += sm.Fpart[(q * 3 + 0) * kTP + tid];
      if (NJ >= 2) f1 += sm.Fpart[(q * 3 + 1) * kTP + tid];
      if (NJ == 3) f2 += sm.Fpart[(q * 3 + 2) * kTP + tid];
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
        // ---- first layer for (trajectory m, units 32*uh ..): a = tanh(w_r r + c), a' = (1-a^2) w_r, a'' = -2 a a' w_r ----
        float av[UPT];
        {
          const float* nc = sm.T.nuc + i * 4;
          const float d0 = sm.xs[m] - nc[0], d1 = sm.xs[kTP + m] - nc[1], d2 = sm.xs[2 * kTP + m] - nc[2];
          const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
          const float* cbi = sm.T.cb + sm.T.cls[i] * kH + UPT * uh;
          // two units per FP32 instruction (packed)
          const tc::f32x2 r2 = tc::splat2(r), tp2 = tc::splat2(tprime);
#pragma unroll
          for (int u = 0; u < UPT; u += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[UPT * uh + u]);
            const float2 t2 = *reinterpret_cast<const float2*>(&sm.wt[UPT * uh + u]);
            const float2 c2 = *reinterpret_cast<const float2*>(&cbi[u]);
            tc::unpack2(tanh2(tc::fma2(tc::pack2(w2.x, w2.y), r2, tc::fma2(tp2, tc::pack2(t2.x, t2.y), tc::pack2(c2.x, c2.y)))),
                        av[u], av[u + 1]);
          }
        }
        {
          // one jet at a time on one mbarrier.  The MMAs read 6 KB of operands per 32 clocks from shared memory when both
          // come from there -- more than its 128 B/clk -- so the activation "hi" part (read by two of the three passes)
          // goes to tensor memory (TS-mode MMA) and only the tf32 remainder is staged in shared memory.
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {
            if (jt > 0) { tc::mbar_wait(&sm.bar[0], phase[0]); phase[0] ^= 1; __syncwarp(); }
            float aj[UPT];
#pragma unroll
            for (int u = 0; u < UPT; u += 2) {
              using namespace tc;
              const f32x2 a2v = pack2(av[u], av[u + 1]);
              f32x2 ajp = a2v;
              if (jt > 0) {
                const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[UPT * uh + u]);
                const f32x2 wv = pack2(w2.x, w2.y);
                const f32x2 a1 = mul2(fma2(mul2(a2v, splat2(-1.0f)), a2v, splat2(1.0f)), wv);      // (1 - a^2) w
                ajp = jt == 1 ? a1 : mul2(mul2(mul2(a2v, splat2(-2.0f)), a1), wv);               // -2 a a' w
              }
              unpack2(ajp, aj[u], aj[u + 1]);
            }
#pragma unroll
            for (int c0 = 0; c0 < UPT; c0 += 16) {
              float hv[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) hv[u] = aj[c0 + u];
              tc::tmem_st16(tlane + kColAhi + UPT * uh + c0, hv);
            }
#pragma unroll
            for (int c = 0; c < UPT / 4; ++c) {
              const int off = (((UPT / 4) * uh + c) * kTP + m) * 4;
              // tf32 remainders x - trunc_tf32(x) = trunc * (-1) + x (exact), two per instruction
              const tc::f32x2 p01 = tc::pack2(aj[4 * c], aj[4 * c + 1]), p23 = tc::pack2(aj[4 * c + 2], aj[4 * c + 3]);
              const tc::f32x2 l01 = tc::fma2(p01 & 0xFFFFE000FFFFE000ull, tc::splat2(-1.0f), p01);
              const tc::f32x2 l23 = tc::fma2(p23 & 0xFFFFE000FFFFE000ull, tc::splat2(-1.0f), p23);
              *reinterpret_cast<float4*>(sm.Alo + off) = make_float4(tc::lo2(l01), tc::hi2(l01), tc::lo2(l23), tc::hi2(l23));
            }
            tc::tmem_st_wait();
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncthreads();
            if (w == 0) {
              tc::fence_after_sync();
              if (tc::elect_one()) {
                constexpr uint64_t stepA = (2 * kChunkBytesA) >> 4, stepB = (2 * kChunkBytesB) >> 4;
                const uint32_t d = tmem + jt * kH;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, tmem + kColAhi + 8 * ks, dWh + ks * stepB, idesc, ks != 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, tmem + kColAhi + 8 * ks, dWl + ks * stepB, idesc, true);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::mma_tf32(d, dAlo + ks * stepA, dWh + ks * stepB, idesc, true);
                tc::commit(&sm.bar[0]);
              }
              __syncwarp();
            }
          }
          tc::mbar_wait(&sm.bar[0], phase[0]); phase[0] ^= 1;
        }
        __syncwarp();                                   // tcgen05.ld is .sync.aligned: reconverge after the spin loops
        tc::fence_after_sync();
        tc::f32x2 fpp[3] = {0ull, 0ull, 0ull};
        float* actp = nullptr;
        if (a.act != nullptr)
          actp = a.act + AL.seg_base((long long)(step * 4 + stage) * a.Na + i, seg_b)
                 + act_group_off((UPT * uh) >> 2, (row0 - seg_row0) + m, act_rows);
        const long long jet_stride = (long long)kH * act_rows;
#pragma unroll
        for (int c0 = 0; c0 < UPT; c0 += 16) {
          float p0[16], p1[16], p2[16];
          tc::tmem_ld16(tlane + 0 * kH + UPT * uh + c0, p0);
          if (NJ >= 2) tc::tmem_ld16(tlane + 1 * kH + UPT * uh + c0, p1);
          if (NJ == 3) tc::tmem_ld16(tlane + 2 * kH + UPT * uh + c0, p2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; u += 2) {
            // layer-2 tanh jet and the last-layer dot, two units per FP32 instruction (packed)
            using namespace tc;
            const int j = UPT * uh + c0 + u;
            const float2 w2 = *reinterpret_cast<const float2*>(&sm.w3[j]);
            const float2 b2v = *reinterpret_cast<const float2*>(&sm.b2[j]);
            const f32x2 wv = pack2(w2.x, w2.y);
            const f32x2 aa = tanh2(add2(pack2(p0[u], p0[u + 1]), pack2(b2v.x, b2v.y)));
            unpack2(aa, p0[u], p0[u + 1]);
            fpp[0] = fma2(wv, aa, fpp[0]);
            if (NJ >= 2) {
              const f32x2 naa = mul2(aa, splat2(-1.0f));
              const f32x2 sg = fma2(naa, aa, splat2(1.0f));
              const f32x2 q1 = pack2(p1[u], p1[u + 1]);
              const f32x2 a1 = mul2(sg, q1);
              fpp[1] = fma2(wv, a1, fpp[1]);
              if (NJ == 3) {
                const f32x2 a2 = fma2(sg, pack2(p2[u], p2[u + 1]), mul2(mul2(mul2(naa, splat2(2.0f)), a1), q1));   // sg p2 - 2 a a1 p1
                fpp[2] = fma2(wv, a2, fpp[2]);
              }
            }
          }
#ifdef OFDFT_ABL_NOACTSTORE    // timing ablation only (the backward then reads garbage)
          if (false) {
#else
          if (actp != nullptr) {
#endif
            // layer-2 jets (a2, p2', p2'') -> HBM, four units per 16-byte store
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float* dst = actp + ((long long)((c0 >> 2) + g) * act_rows) * 4;
              __stcs(reinterpret_cast<float4*>(dst), make_float4(p0[4 * g], p0[4 * g + 1], p0[4 * g + 2], p0[4 * g + 3]));
              if (NJ >= 2) __stcs(reinterpret_cast<float4*>(dst + jet_stride), make_float4(p1[4 * g], p1[4 * g + 1], p1[4 * g + 2], p1[4 * g + 3]));
              if (NJ == 3) __stcs(reinterpret_cast<float4*>(dst + 2 * jet_stride), make_float4(p2[4 * g], p2[4 * g + 1], p2[4 * g + 2], p2[4 * g + 3]));
            }
          }
        }
#pragma unroll
        for (int jt = 0; jt < NJ; ++jt) sm.Fpart[(uh * 3 + jt) * kTP + m] = tc::lo2(fpp[jt]) + tc::hi2(fpp[jt]);
        tc::fence_before_sync();
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

template <int NU>
__global__ void __launch_bounds__(kTP * NU, 2) gnn_fwd_tc_kernel(const GnnFwdArgs a) {
  constexpr int kThr = kTP * NU;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemTc1& sm = *reinterpret_cast<SmemTc1*>(smem_raw);
  const int tid = threadIdx.x;
  const GnnThetaLayout L(a.n_types);
  // weight operand B[n = j][k] = W2[k][j], split hi/lo, layout [k/4][j][k%4]
  for (int i = tid; i < kH * kH; i += kThr) {
    const int k = i / kH, j = i % kH;
    float hi, lo;
    tc::split_tf32(a.theta[L.W2 + i], hi, lo);
    const int off = ((k >> 2) * kH + j) * 4 + (k & 3);
    sm.Whi[off] = hi; sm.Wlo[off] = lo;
  }
  if (tid < kH) {
    sm.wt[tid] = a.theta[L.W1 + tid];
    sm.wr[tid] = a.theta[L.W1 + kH + tid];
    sm.b2[tid] = a.theta[L.b2 + tid];
    sm.w3[tid] = a.theta[L.w3 + tid];
  }
  load_nuc_tables<kThr>(sm.T, a.theta, L, a.nuclei, a.onehot, a.Na, a.n_types);
  if (tid == 0) {
    tc::mbar_init(&sm.bar[0], 1); tc::mbar_init(&sm.bar[1], 1); tc::mbar_init(&sm.bar[2], 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(&sm.tmem_slot, 256);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  uint32_t phase[3] = {0, 0, 0};
  const float b3 = a.theta[0];
  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  long long tile;
  while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
    long long row0, seg_end; int nj;
    bool seg_b; long long seg_row0;
    if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; seg_b = false; seg_row0 = 0; }
    else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; seg_b = true; seg_row0 = a.n_a; }
    if (nj == 3) fwd_tile_tc<3, NU>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
    else if (nj == 2) fwd_tile_tc<2, NU>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
    else fwd_tile_tc<1, NU>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 256);
}


// =================================================================================================================
// tensor-core backward (discrete adjoint) -- requires the layer-2 activation checkpoint
// =================================================================================================================
// Per (tile, nucleus, RHS evaluation) and jet:
//   B1 :  Abar1[jet] (128 traj x 64 k)  = Pbar2[jet] . W2^T as 3xTF32, A operand in tensor memory (the parked jet itself for
//         the two "hi" passes, its tf32 remainder for the third)                                          24 MMAs (N = 64)
//   dW2:  D_dw[(hi|lo) k][(hi|lo) j] += A1[jet]^T Pbar2[jet]  with K = trajectory: both operands staged "transposed"
//         ([chunk of 4 traj][row][4], chunk stride padded by 16 B so the scalar stores are conflict free), A rows stacked
//         [hi k | lo k], B rows stacked [hi j | lo j]: one M128.N128.K8 MMA per K step gives all four cross terms   16 MMAs
// The tensor core accumulates with truncation: a D_dw chain spanning the whole tile (~2300 MMAs) biases the gradient by
// ~1e-4.  D_dw is therefore restarted every RK4 stage (Na x 48 MMAs) and folded into round-to-nearest fp32 accumulators in
// shared memory (see DESIGN.md for the measured errors).
constexpr int kA1TChunk = kTP * 4 + 4;      // floats per trajectory-chunk of the stacked A1^T operand (128 rows)
constexpr int kP2TChunk = 2 * kH * 4 + 4;   // floats per trajectory-chunk of the stacked [hi j | lo j] Pbar2^T operand (128 rows)
// tensor-memory map (512 columns): B1 accumulators of the three jets | dW2 accumulator | Pbar2 of the three jets (the
// unrounded values double as the "hi" operand of B1: the tensor core truncates) | tf32 remainder of the jet being staged
// The remainder of jet 0 borrows the accumulator columns of jet 2 (B1lo(0) has run before B1hi(2) is issued), the remainder
// of jet jt >= 1 the parked columns of jet jt - 1 (dead once B1(jt - 1) has completed): no dedicated remainder buffer.
constexpr int kColB1 = 0, kColDw = 192, kColP = 320;
__host__ __device__ constexpr int col_plo(int jt) { return jt == 0 ? kColB1 + 2 * kH : kColP + (jt - 1) * kH; }
#ifndef OFDFT_TC_FLUSH_PER_EVAL
#define OFDFT_TC_FLUSH_PER_EVAL 0
#endif
constexpr bool kFlushPerEval = OFDFT_TC_FLUSH_PER_EVAL != 0;
#ifndef OFDFT_TC_PREFETCH_L2
#define OFDFT_TC_PREFETCH_L2 1
#endif
constexpr bool kPrefetchL2 = OFDFT_TC_PREFETCH_L2 != 0;   // else: once per RK4 stage (Na evaluations, Na*96 MMAs)

struct SmemTcB {
  float A1T[32 * kA1TChunk];
  float P2T[32 * kP2TChunk];              // rows 0..63: Pbar2 (hi), rows 64..127: its tf32 remainder -- one N = 128 B operand
  float WTh[kH * kH], WTl[kH * kH];       // B operand of B1: [j/4][k][j%4] = W2[k][j]
  float wr[kH], wt[kH], b2[kH], w3[kH];
  NucTables T;
  float S[8 * kMaxClasses * 32];          // per warp, per nucleus class: sum over the warp's trajectories of pbar1[k = 32 uh + lane]
  float Fpart[2][2 * 3 * kTP];            // [nucleus parity][unit half][jet][trajectory]: exchanged between the two warps of a trajectory
  float Rpart[2][2 * kTP];
  float red[8 * 32 * 4];
  float DW[kTP * kH];                     // fp32 (round-to-nearest) accumulators of the stacked dW2 tile, [col j][row]
  uint64_t barB[3], barD[3];              // per jet: B1 batch done / dW2 batch done (each completes once per evaluation)
  uint64_t barR[3];                       // recompute mode: layer-2 GEMM of jet jt of the evaluation done
  unsigned int tmem_slot;
  int tile;
};

static_assert(sizeof(SmemTcB) + 1024 <= 227 * 1024, "backward tile state exceeds the 227 KB of shared memory per CTA");
static_assert(2 * (sizeof(SmemTc1) + 1024) <= 227 * 1024, "two forward CTAs must fit one SM");

// named barriers of the backward kernel (0 = __syncthreads of the whole CTA)
constexpr int kBarPair = 1;          // 1..4: the two warps of a trajectory quarter (64 threads)
constexpr int kBarCompute = 5;       // the 256 compute threads
constexpr int kBarStaged0 = 6;       // compute -> MMA warp: jet-0 operand of B1 is in tensor memory
constexpr int kBarStagedJet = 7;     // 7..9: compute -> MMA warp: weight-gradient operands of jet jt (and the B1 operand of jt + 1)
constexpr int kBarRecomp = 10;       // 10..12, recompute mode: compute -> MMA warp: layer-1 jet jt staged for the layer-2 GEMM
constexpr int kBwdThreads = kThreads + 128;      // two compute warpgroups + the MMA-issue warpgroup
// Recompute mode (no activation checkpoint): before the jets of an evaluation are staged the [A1T | P2T] operand region is
// idle and holds, for the layer-2 forward GEMM, the tf32 remainders of the three layer-1 jets (K-major [k/4][row][4], 32 KB
// each) and a copy of W2 in the forward orientation (hi, lo: [k/4][j][4], 16 KB each)
constexpr int kRcLoFloats = kH * kTP;                 // one remainder buffer
constexpr int kRcWOff = 3 * kRcLoFloats;              // float offset of W2 hi (lo follows)
static_assert((kRcWOff + 2 * kH * kH) * 4 <= (int)(sizeof(float) * 32 * (kA1TChunk + kP2TChunk)), "recompute operands exceed the staging region");
static_assert(offsetof(SmemTcB, P2T) == offsetof(SmemTcB, A1T) + sizeof(float) * 32 * kA1TChunk, "A1T and P2T must be contiguous");

#ifdef OFDFT_BWD_WAITPROF    // development only: clocks the compute threads spend in each tensor-core wait (thread 0 of every CTA)
__device__ unsigned long long g_waitprof[8];
#define WP_T0 const long long wp_t0_ = clock64();
#define WP_ADD(i) if (threadIdx.x == 0) atomicAdd(&g_waitprof[i], (unsigned long long)(clock64() - wp_t0_));
#else
#define WP_T0
#define WP_ADD(i)
#endif
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// The MMA-issue warp: mirrors the control flow of bwd_tile_tc and issues, per (RK4 stage, nucleus),
//   B1hi(0) | dW2(0) B1lo(0) B1hi(1) | dW2(1) B1lo(1) B1hi(2) | B1lo(2) dW2(2)
// as soon as all 256 compute threads have ARRIVED at the matching named barrier.  The compute warps never block on each
// other for an MMA hand-off (they only wait on the completion mbarriers they depend on), so they drift apart and their
// load / SFU / shared-memory phases overlap instead of running in lock-step.
template <int NJ, bool RC>
__device__ void bwd_mma_tile(SmemTcB& sm, const GnnBwdArgs& a, uint32_t tmem) {
  const uint32_t idesc = tc::make_idesc_tf32(kTP, kH, 0, 0);
  const uint64_t dA1T = tc::make_desc(tc::smem_u32(sm.A1T), kA1TChunk * 4, 128);
  const uint64_t dP2 = tc::make_desc(tc::smem_u32(sm.P2T), kP2TChunk * 4, 128);
  const uint32_t idesc_dw = tc::make_idesc_tf32(kTP, 2 * kH, 0, 0);      // M = 128 (hi k | lo k), N = 128 (hi j | lo j)
  const uint64_t dWTh = tc::make_desc(tc::smem_u32(sm.WTh), kH * 16, 128);
  const uint64_t dWTl = tc::make_desc(tc::smem_u32(sm.WTl), kH * 16, 128);
  constexpr uint64_t stepW = (2 * kH * 16) >> 4;
  // B1[jt] = Pbar2[jt] . W2^T as 3xTF32: the two passes that read the parked jet itself (x W_hi, x W_lo) ...
  auto issue_b1_hi = [&](int jt) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const uint64_t wb = p == 1 ? dWTl : dWTh;
#pragma unroll 1
      for (int ksx = 0; ksx < 8; ++ksx)
        tc::mma_tf32_ts(tmem + kColB1 + jt * kH, tmem + kColP + jt * kH + 8 * ksx, wb + ksx * stepW, idesc, (p | ksx) != 0);
    }
  };
  // ... and the pass that reads its tf32 remainder from the single staging buffer (x W_hi), once that jet is staged
  auto issue_b1_lo = [&](int jt) {
#pragma unroll 1
    for (int ksx = 0; ksx < 8; ++ksx)
      tc::mma_tf32_ts(tmem + kColB1 + jt * kH, tmem + col_plo(jt) + 8 * ksx, dWTh + ksx * stepW, idesc, true);
    tc::commit(&sm.barB[jt]);
  };
  const int n_eval = a.n_steps * 4;
#pragma unroll 1
  for (int ev = 0; ev < n_eval; ++ev) {
#pragma unroll 1
    for (int i = 0; i < a.Na; ++i) {
      if (RC) {
        // layer-2 forward GEMM of this evaluation: P2[jet] = A1[jet] . W2 as 3xTF32 (hi in tensor memory -- the columns the
        // cotangent jets are parked in later --, remainder and W2 in the idle staging region); result in the B1 accumulators
        // (one hand-off and one completion barrier per jet: the GEMM of jet jt runs under the staging of jet jt + 1)
        const uint32_t base = tc::smem_u32(sm.A1T);
        const uint64_t dWh = tc::make_desc(base + 4 * kRcWOff, kChunkBytesB, 128);
        const uint64_t dWl = tc::make_desc(base + 4 * (kRcWOff + kH * kH), kChunkBytesB, 128);
        constexpr uint64_t stepA = (2 * kChunkBytesA) >> 4, stepB = (2 * kChunkBytesB) >> 4;
#pragma unroll 1
        for (int jt = 0; jt < NJ; ++jt) {
          bar_sync(kBarRecomp + jt, kThreads + 32);
          tc::fence_after_sync();
          if (tc::elect_one()) {
            const uint32_t d = tmem + kColB1 + jt * kH, ahi = tmem + kColP + jt * kH;
            const uint64_t dLo = tc::make_desc(base + 4 * jt * kRcLoFloats, kChunkBytesA, 128);
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, ahi + 8 * ks, dWh + ks * stepB, idesc, ks != 0);
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks) tc::mma_tf32_ts(d, ahi + 8 * ks, dWl + ks * stepB, idesc, true);
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks) tc::mma_tf32(d, dLo + ks * stepA, dWh + ks * stepB, idesc, true);
            tc::commit(&sm.barR[jt]);
          }
          __syncwarp();
        }
      }
      bar_sync(kBarStaged0, kThreads + 32);
      tc::fence_after_sync();
      if (tc::elect_one()) issue_b1_hi(0);
      __syncwarp();
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt) {
        bar_sync(kBarStagedJet + jt, kThreads + 32);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          // dW2: D_dw += [A1_hi ; A1_lo]^T-stacked . (P_hi, then P_lo), K = 128 trajectories.  It goes first: the compute
          // warps restage shared memory as soon as it completes, the short B1lo pass behind it finishes under that staging
          // (last jet: B1lo first -- the layer-1 reverse waits for it while dW2 runs underneath)
          if (jt == NJ - 1) issue_b1_lo(jt);
#pragma unroll 1
          for (int ksx = 0; ksx < 16; ++ksx) {
            tc::mma_tf32(tmem + kColDw, dA1T + ksx * ((2 * kA1TChunk * 4) >> 4), dP2 + ksx * ((2 * kP2TChunk * 4) >> 4), idesc_dw,
                         (kFlushPerEval ? 0 : i) != 0 || (jt | ksx) != 0);
          }
          tc::commit(&sm.barD[jt]);
          if (jt < NJ - 1) issue_b1_lo(jt);
          // the hi passes of the next jet go behind the batch the compute warps are waiting for
          if (jt + 1 < NJ) issue_b1_hi(jt + 1);
        }
        __syncwarp();
      }
    }
  }
}

struct TcBwdAccum {
  float g_b2, g_w3;      // lane <-> unit j = 32 uh + lane, summed over this warp's 32 trajectories
  float g_wt, g_wr;      // lane <-> unit k = 32 uh + lane
  float g_b3;            // owner threads
};

template <int NJ, bool RC>
__device__ void bwd_tile_tc(SmemTcB& sm, const GnnBwdArgs& a, long long row0, long long seg_end, float b3, TcBwdAccum& G,
                            bool seg_b, long long seg_row0, uint32_t tmem, uint32_t& phase, bool& dw_pending) {
  const ActLayout AL(a.act_rows_a, a.act_rows_b, a.jets_a, a.jets_b);
  const long long act_rows = seg_b ? a.act_rows_b : a.act_rows_a;
  const long long jet_stride = (long long)kH * act_rows;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m = 32 * (w & 3) + lane, uh = w >> 2;
  // Both threads of a trajectory (unit halves uh = 0 / 1) carry its adjoint state and do the per-nucleus scalar algebra
  // redundantly: nothing is serialised on an "owner" warp, and the only exchange per nucleus -- the partial dot products of
  // the two unit halves -- needs a 64-thread named barrier between the two warps of the trajectory quarter, not the CTA.
  const bool owner = uh == 0;                       // writes results / accumulates per-trajectory scalars once
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
  float fb0 = 0.f, fb1 = 0.f, fb2 = 0.f;            // cotangents of (f, f', f'') of the current nucleus
  int par = 0;                                      // parity of the Fpart / Rpart exchange buffers
  const uint32_t tlane = tmem + ((uint32_t)(32 * (w & 3)) << 16);

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
    if (owner) G.g_b3 += fb0;
  };
  auto post = [&](int i) {
    const float* nc = sm.T.nuc + i * 4;
    const float* Fp = sm.Fpart[par];
    float f0 = b3 + Fp[0 * kTP + m] + Fp[3 * kTP + m], f1 = 0.f, f2 = 0.f;
    if (NJ >= 2) f1 = Fp[1 * kTP + m] + Fp[4 * kTP + m];
    if (NJ == 3) f2 = Fp[2 * kTP + m] + Fp[5 * kTP + m];
    float rbar = sm.Rpart[par][m] + sm.Rpart[par][kTP + m];
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

  for (int step = a.n_steps - 1; step >= 0; --step) {
    sumx[0] = sumx[1] = sumx[2] = 0.f; sums[0] = sums[1] = sums[2] = 0.f;
#pragma unroll 1
    for (int stage = 3; stage >= 0; --stage) {
      const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
      const float ts = a.t0 + h * (float)step + cs;
      const float tprime = a.y0 * ts;
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
        pre(0);
      }
      // x-only tiles (NJ == 1) have registers to spare: the four per-unit sums over the warp's trajectories are accumulated
      // over the nuclei of the stage and reduced by ONE butterfly each per stage (per run of equal nucleus class for the
      // first-layer pre-activation sum) instead of one per nucleus -- the butterflies are a third of such a tile's instructions
#ifndef OFDFT_STAGE_ACC
#define OFDFT_STAGE_ACC 1
#endif
      constexpr bool kStageAccR = NJ == 1 && (OFDFT_STAGE_ACC & 1);     // b2 / w3 sums (reverse epilogue)
      constexpr bool kStageAcc = NJ == 1 && (OFDFT_STAGE_ACC & 2);      // first-layer sums
      float b2acc[kStageAccR ? 32 : 1], w3acc[kStageAccR ? 32 : 1], pracc[kStageAcc ? 32 : 1], psacc[kStageAcc ? 32 : 1];
      int ps_cls = -1;
      if constexpr (kStageAccR) {
#pragma unroll
        for (int u = 0; u < 32; ++u) { b2acc[u] = 0.f; w3acc[u] = 0.f; }
      }
      if constexpr (kStageAcc) {
#pragma unroll
        for (int u = 0; u < 32; ++u) { pracc[u] = 0.f; psacc[u] = 0.f; }
      }
      auto flush_ps = [&]() {
        if constexpr (kStageAcc) { if (ps_cls >= 0) {
          const float ps = warp_vec32_sum(psacc, lane);
          sm.S[(w * kMaxClasses + ps_cls) * 32 + lane] += ps;
          G.g_wt = fmaf(tprime, ps, G.g_wt);
#pragma unroll
          for (int u = 0; u < 32; ++u) psacc[u] = 0.f;
        } }
      };
#pragma unroll 1
      for (int i = 0; i < a.Na; ++i) {
        // ---- layer-2 jets (a2, p2', p2'') of (trajectory m, units 32 uh ..) from the HBM checkpoint ----
        float q0[32], q1[32], q2[32];
        if constexpr (!RC) {
          const float* src = a.act + AL.seg_base((long long)(step * 4 + stage) * a.Na + i, seg_b)
                             + act_group_off(8 * uh, (row0 - seg_row0) + m, act_rows);
#pragma unroll
#ifdef OFDFT_ABL_NOLOAD     // timing ablation only (results invalid): the same 2 KB for every evaluation
          src = a.act + (threadIdx.x & 127) * 4;
#endif
          for (int g = 0; g < 8; ++g) {
            const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src + (long long)g * act_rows * 4));
            q0[4 * g] = v0.x; q0[4 * g + 1] = v0.y; q0[4 * g + 2] = v0.z; q0[4 * g + 3] = v0.w;
            if (NJ >= 2) {
              const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src + jet_stride + (long long)g * act_rows * 4));
              q1[4 * g] = v1.x; q1[4 * g + 1] = v1.y; q1[4 * g + 2] = v1.z; q1[4 * g + 3] = v1.w;
            }
            if (NJ == 3) {
              const float4 v2 = __ldcs(reinterpret_cast<const float4*>(src + 2 * jet_stride + (long long)g * act_rows * 4));
              q2[4 * g] = v2.x; q2[4 * g + 1] = v2.y; q2[4 * g + 2] = v2.z; q2[4 * g + 3] = v2.w;
            }
          }
          // pull the next evaluation's jets (backward order: next nucleus, else the previous stage) into L2 behind these
          // loads: per jet this warp reads 8 unit groups x 512 B = 32 lines, lane <-> (group, line)
          const long long ss = (long long)(step * 4 + stage);
          const long long en = (i + 1 < a.Na) ? ss * a.Na + i + 1 : (ss - 1) * a.Na;
          if (kPrefetchL2 && en >= 0) {
            const float* nx = a.act + AL.seg_base(en, seg_b)
                              + act_group_off(8 * uh + (lane >> 2), (row0 - seg_row0) + 32 * (w & 3) + 8 * (lane & 3), act_rows);
#pragma unroll
            for (int jt = 0; jt < NJ; ++jt) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + jt * jet_stride));
          }
        }
        // ---- first layer (recomputed): a1 = tanh(w_r r + c) ----
        float av[32];
        float r;
        {
          const float* nc = sm.T.nuc + i * 4;
          const float d0 = xst[0] - nc[0], d1 = xst[1] - nc[1], d2 = xst[2] - nc[2];
          r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
          const float* cbi = sm.T.cb + sm.T.cls[i] * kH + 32 * uh;
          // tanh(x) = 1 - 2 / (2^(2 x log2 e) + 1), two units per FP32 instruction (packed), one MUFU.EX2 + MUFU.RCP each
          using namespace tc;
          const f32x2 r2 = splat2(r), tp2 = splat2(tprime), one = splat2(1.0f), m2 = splat2(-2.0f), k2 = splat2(2.8853900817779268f);
#pragma unroll
          for (int u = 0; u < 32; u += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[32 * uh + u]);
            const float2 t2 = *reinterpret_cast<const float2*>(&sm.wt[32 * uh + u]);
            const float2 c2 = *reinterpret_cast<const float2*>(&cbi[u]);
            const f32x2 arg = mul2(fma2(pack2(w2.x, w2.y), r2, fma2(tp2, pack2(t2.x, t2.y), pack2(c2.x, c2.y))), k2);
            float x0, x1, e0, e1, q0r, q1r;
            unpack2(arg, x0, x1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(x0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(x1));
            const f32x2 den = add2(pack2(e0, e1), one);
            float d0, d1;
            unpack2(den, d0, d1);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q0r) : "f"(d0));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q1r) : "f"(d1));
            unpack2(fma2(m2, pack2(q0r, q1r), one), av[u], av[u + 1]);
          }
        }
        if constexpr (RC) {
          // ---- recompute mode: layer-2 jets (a2, p2', p2'') = tanh-jet of A1[jet] . W2 on the tensor cores ----
          using namespace tc;
          float* region = sm.A1T;
          // the staging region is free once the previous evaluation's last weight-gradient batch has completed
          if (dw_pending) { mbar_wait(&sm.barD[NJ - 1], (phase >> (3 + NJ - 1)) & 1u); phase ^= 1u << (3 + NJ - 1); dw_pending = false; __syncwarp(); }
          // W2 (forward orientation, hi | lo) from the prepared global copy: 8192 floats, 32 per thread
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int o = (q * kThreads + (int)threadIdx.x) * 4;
            *reinterpret_cast<float4*>(region + kRcWOff + o) = __ldg(reinterpret_cast<const float4*>(a.w2split + o));
          }
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float hv[16];
#pragma unroll
              for (int u = 0; u < 16; u += 4) {
                f32x2 p[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int uu = c0 + u + 2 * h;
                  const f32x2 a2v = pack2(av[uu], av[uu + 1]);
                  f32x2 ajp = a2v;
                  if (jt > 0) {
                    const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[32 * uh + uu]);
                    const f32x2 wv = pack2(w2.x, w2.y);
                    const f32x2 a1 = mul2(fma2(mul2(a2v, splat2(-1.0f)), a2v, splat2(1.0f)), wv);      // (1 - a^2) w
                    ajp = jt == 1 ? a1 : mul2(mul2(mul2(a2v, splat2(-2.0f)), a1), wv);               // -2 a a' w
                  }
                  p[h] = ajp;
                  unpack2(ajp, hv[u + 2 * h], hv[u + 2 * h + 1]);
                }
                const f32x2 l0 = fma2(p[0] & 0xFFFFE000FFFFE000ull, splat2(-1.0f), p[0]);
                const f32x2 l1 = fma2(p[1] & 0xFFFFE000FFFFE000ull, splat2(-1.0f), p[1]);
                const int chunk = 8 * uh + ((c0 + u) >> 2);
                *reinterpret_cast<float4*>(region + jt * kRcLoFloats + (chunk * kTP + m) * 4) = make_float4(lo2(l0), hi2(l0), lo2(l1), hi2(l1));
              }
              tmem_st16(tlane + kColP + jt * kH + 32 * uh + c0, hv);
            }
            tmem_st_wait();
            fence_async_smem();
            fence_before_sync();
            bar_arrive(kBarRecomp + jt, kThreads + 32);    // -> MMA warp: layer-2 GEMM of jet jt (runs under the next staging)
          }
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) {
            mbar_wait(&sm.barR[jt], (phase >> (6 + jt)) & 1u); phase ^= 1u << (6 + jt);
            __syncwarp();
            fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float t[16];
              tmem_ld16(tlane + kColB1 + jt * kH + 32 * uh + c0, t);
              tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; u += 2) {
                if (jt == 0) {
                  const float2 bb = *reinterpret_cast<const float2*>(&sm.b2[32 * uh + c0 + u]);
                  unpack2(tanh2(add2(pack2(t[u], t[u + 1]), pack2(bb.x, bb.y))), q0[c0 + u], q0[c0 + u + 1]);
                } else if (jt == 1) { q1[c0 + u] = t[u]; q1[c0 + u + 1] = t[u + 1]; }
                else { q2[c0 + u] = t[u]; q2[c0 + u + 1] = t[u + 1]; }
              }
            }
          }
          fence_before_sync();
        }
        // ---- layer-2 tanh-jet: forward partials of f and reverse to the pre-activation cotangents (in place) ----
        float g32[32];
        {
          // two hidden units per instruction (packed FP32): same formulas as the scalar epilogue of gnn.cu
          using namespace tc;
          const f32x2 one = splat2(1.0f), m1 = splat2(-1.0f), two = splat2(2.0f), four = splat2(4.0f), m3 = splat2(-3.0f), m2 = splat2(-2.0f);
          const f32x2 fb0p = splat2(fb0), fb1p = splat2(fb1), fb2p = splat2(fb2);
          f32x2 fpp[3] = {0ull, 0ull, 0ull};
#pragma unroll
          for (int u = 0; u < 32; u += 2) {
            const float2 w2 = *reinterpret_cast<const float2*>(&sm.w3[32 * uh + u]);
            const f32x2 wv = pack2(w2.x, w2.y);
            const f32x2 aa = pack2(q0[u], q0[u + 1]);
            const f32x2 naa = mul2(aa, m1);                       // -a
            const f32x2 sg = fma2(naa, aa, one);                  // 1 - a^2
            fpp[0] = fma2(wv, aa, fpp[0]);
            f32x2 gw = mul2(aa, fb0p);
            f32x2 pzb = mul2(wv, fb0p);
            if (NJ >= 2) {
              const f32x2 p1 = pack2(q1[u], q1[u + 1]);
              const f32x2 a1 = mul2(sg, p1);
              fpp[1] = fma2(wv, a1, fpp[1]);
              gw = fma2(a1, fb1p, gw);
              const f32x2 a1b = mul2(wv, fb1p);
              const f32x2 nap1 = mul2(naa, p1);                   // -a p1
              const f32x2 m2ap1 = mul2(nap1, two);                // -2 a p1
              pzb = fma2(m2ap1, a1b, pzb);
              f32x2 t1 = a1b;
              if (NJ == 3) {
                const f32x2 p2 = pack2(q2[u], q2[u + 1]);
                const f32x2 a2 = fma2(sg, p2, mul2(m2ap1, a1));   // (1 - a^2) p2 - 2 a a1 p1
                fpp[2] = fma2(wv, a2, fpp[2]);
                gw = fma2(a2, fb2p, gw);
                const f32x2 a2b = mul2(wv, fb2p);
                const f32x2 o2 = mul2(a2b, sg);
                unpack2(o2, q2[u], q2[u + 1]);
                t1 = fma2(mul2(nap1, four), a2b, t1);             // a1b - 4 a p1 a2b
                // -2 a p2 - 2 (1 - 3 a^2) p1^2
                const f32x2 c3 = fma2(mul2(aa, m3), aa, one);     // 1 - 3 a^2
                const f32x2 inner = fma2(mul2(naa, two), p2, mul2(mul2(c3, m2), mul2(p1, p1)));
                pzb = fma2(a2b, inner, pzb);
              }
              const f32x2 o1 = mul2(sg, t1);
              unpack2(o1, q1[u], q1[u + 1]);
            }
            const f32x2 o0 = mul2(sg, pzb);
            unpack2(o0, q0[u], q0[u + 1]);
            unpack2(gw, g32[u], g32[u + 1]);
          }
          float fp[3];
#pragma unroll
          for (int jt = 0; jt < 3; ++jt) { float x0, x1; unpack2(fpp[jt], x0, x1); fp[jt] = x0 + x1; }
#pragma unroll
          for (int jt = 0; jt < NJ; ++jt) sm.Fpart[par][(uh * 3 + jt) * kTP + m] = fp[jt];
        }
        // ---- per jet: B1 (A operand in tensor memory) and the dW2 update (operands in shared memory); issue order and
        // waits: see bwd_mma_tile ----
        // Park the three cotangent jets in tensor memory (they are the unrounded "hi" operand of B1) and let the MMA warp
        // start the B1 passes that only need them; from here on a jet lives in registers only while it is being staged.
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          float hv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) hv[u] = q0[c0 + u];
          tc::tmem_st16(tlane + kColP + 0 * kH + 32 * uh + c0, hv);
          if (NJ >= 2) {
#pragma unroll
            for (int u = 0; u < 16; ++u) hv[u] = q1[c0 + u];
            tc::tmem_st16(tlane + kColP + 1 * kH + 32 * uh + c0, hv);
          }
          if (NJ == 3) {
#pragma unroll
            for (int u = 0; u < 16; ++u) hv[u] = q2[c0 + u];
            tc::tmem_st16(tlane + kColP + 2 * kH + 32 * uh + c0, hv);
          }
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        bar_arrive(kBarStaged0, kThreads + 32);          // -> MMA warp: B1hi(0 .. NJ-1)
        if constexpr (kStageAccR) {
#pragma unroll
          for (int u = 0; u < 32; ++u) { w3acc[u] += g32[u]; b2acc[u] += q0[u]; }
        } else {
          G.g_w3 += warp_vec32_sum(g32, lane);
        }
#pragma unroll
        for (int jt = 0; jt < NJ; ++jt) {
          // the bias gradient sum_traj Pbar2[0] needs q0, which is dead once jet 0 has been staged: its butterfly runs in the
          // shadow of the first weight-gradient batch instead of in front of the staging
          if (jt == 1) G.g_b2 += warp_vec32_sum(q0, lane);
          // the shared-memory operands and the remainder buffer are free once the previous dW2 batch has completed
          // (its commit also covers the B1lo pass issued in front of it)
          { WP_T0
          if (jt > 0) { tc::mbar_wait(&sm.barD[jt - 1], (phase >> (3 + jt - 1)) & 1u); phase ^= 1u << (3 + jt - 1); __syncwarp(); }
          else if (dw_pending) { tc::mbar_wait(&sm.barD[NJ - 1], (phase >> (3 + NJ - 1)) & 1u); phase ^= 1u << (3 + NJ - 1); __syncwarp(); }
          WP_ADD(jt == 0 ? 0 : 1) }
          float pj[32];
          if (jt == 0) {
#pragma unroll
            for (int u = 0; u < 32; ++u) pj[u] = q0[u];
          } else {
            tc::fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
              float t16[16];
              tc::tmem_ld16(tlane + kColP + jt * kH + 32 * uh + c0, t16);
              tc::tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 16; ++u) pj[c0 + u] = t16[u];
            }
          }
          const int cbase = (m >> 2), csub = (m & 3);
          float lv[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
#pragma unroll
            for (int u2 = 0; u2 < 16; u2 += 2) {
              // the jet of the first-layer activation and the tf32 remainders, two units per instruction (packed FP32)
              using namespace tc;
              const f32x2 a2v = pack2(av[c0 + u2], av[c0 + u2 + 1]);
              f32x2 a1p = a2v;
              if (jt > 0) {
                const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[32 * uh + c0 + u2]);
                const f32x2 wv2 = pack2(w2.x, w2.y);
                const f32x2 sg1 = fma2(mul2(a2v, splat2(-1.0f)), a2v, splat2(1.0f));
                a1p = jt == 1 ? mul2(sg1, wv2) : mul2(mul2(mul2(a2v, sg1), splat2(-2.0f)), mul2(wv2, wv2));
              }
              const f32x2 pbp = pack2(pj[c0 + u2], pj[c0 + u2 + 1]);
              // remainder x - trunc_tf32(x) = trunc * (-1) + x (exact)
              const f32x2 a1lo = fma2(a1p & 0xFFFFE000FFFFE000ull, splat2(-1.0f), a1p);
              const f32x2 pblo = fma2(pbp & 0xFFFFE000FFFFE000ull, splat2(-1.0f), pbp);
              float a1j_[2], a1l_[2], pbj_[2];
              unpack2(a1p, a1j_[0], a1j_[1]); unpack2(a1lo, a1l_[0], a1l_[1]); unpack2(pbp, pbj_[0], pbj_[1]);
              unpack2(pblo, lv[c0 + u2], lv[c0 + u2 + 1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int uu = c0 + u2 + h;
              const float a1j = a1j_[h], pbj = pbj_[h];
              const float a1lo_s = a1l_[h];
              // operands go in unrounded: the tensor core truncates them to tf32, the remainders are the "lo" parts
#ifdef OFDFT_ABL_NOSTAGE     // timing ablation only (results invalid)
              if (a1j == 123.456f) sm.A1T[uu] = pbj;
#else
              sm.A1T[cbase * kA1TChunk + (32 * uh + uu) * 4 + csub] = a1j;
              sm.A1T[cbase * kA1TChunk + (kH + 32 * uh + uu) * 4 + csub] = a1lo_s;
              sm.P2T[cbase * kP2TChunk + (32 * uh + uu) * 4 + csub] = pbj;
              sm.P2T[cbase * kP2TChunk + (kH + 32 * uh + uu) * 4 + csub] = lv[uu];
#endif
            }
            }
          }
          // the remainder buffer is free once the previous jet's B1lo pass (issued behind its dW2 batch) has completed
          { WP_T0
          if (jt > 0) { tc::mbar_wait(&sm.barB[jt - 1], (phase >> (jt - 1)) & 1u); phase ^= 1u << (jt - 1); __syncwarp(); }
          WP_ADD(2) }
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float l16[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) l16[u] = lv[c0 + u];
            tc::tmem_st16(tlane + col_plo(jt) + 32 * uh + c0, l16);
          }
          tc::tmem_st_wait();
          tc::fence_async_smem();
          tc::fence_before_sync();
          bar_arrive(kBarStagedJet + jt, kThreads + 32);   // -> MMA warp: B1lo(jt), dW2(jt)
        }
        if (NJ == 1 && !kStageAccR) G.g_b2 += warp_vec32_sum(q0, lane);
        dw_pending = true;
        // the earlier B1 batches were retired before their remainder buffer was reused: wait for the last one
        { WP_T0
        tc::mbar_wait(&sm.barB[NJ - 1], (phase >> (NJ - 1)) & 1u); phase ^= 1u << (NJ - 1);
        WP_ADD(3) }
        __syncwarp();
        tc::fence_after_sync();
        // ---- layer-1 reverse from D_b1 (cotangent jets of a1 for trajectory m, units k = 32 uh ..) ----
        {
          float rp = 0.f;
          tc::f32x2 rp2 = 0ull;
          float ps32[32], pr32[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float b0[16], b1[16], b2v[16];
            tc::tmem_ld16(tlane + kColB1 + 0 * kH + 32 * uh + c0, b0);
            if (NJ >= 2) tc::tmem_ld16(tlane + kColB1 + 1 * kH + 32 * uh + c0, b1);
            if (NJ == 3) tc::tmem_ld16(tlane + kColB1 + 2 * kH + 32 * uh + c0, b2v);
            tc::tmem_ld_wait();
            // two units per instruction (packed FP32)
            using namespace tc;
            const f32x2 one = splat2(1.0f), m2 = splat2(-2.0f), m3 = splat2(-3.0f), m4 = splat2(-4.0f), rr2 = splat2(r);
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const f32x2 aa = pack2(av[c0 + u], av[c0 + u + 1]);
              const float2 w2 = *reinterpret_cast<const float2*>(&sm.wr[32 * uh + c0 + u]);
              const f32x2 wv = pack2(w2.x, w2.y);
              const f32x2 aw = mul2(aa, wv);                                  // a w
              const f32x2 sg = fma2(mul2(aa, splat2(-1.0f)), aa, one);        // 1 - a^2
              f32x2 t = pack2(b0[u], b0[u + 1]);
              f32x2 e = 0ull;
              if (NJ >= 2) {
                const f32x2 c1 = pack2(b1[u], b1[u + 1]);
                t = fma2(mul2(aw, m2), c1, t);                                // - 2 a w b1
                e = c1;
                if (NJ == 3) {
                  const f32x2 c2 = pack2(b2v[u], b2v[u + 1]);
                  const f32x2 c3 = fma2(mul2(aa, m3), aa, one);               // 1 - 3 a^2
                  t = fma2(mul2(mul2(mul2(wv, wv), m2), c3), c2, t);          // - 2 w^2 (1 - 3 a^2) b2
                  e = fma2(mul2(aw, m4), c2, e);                              // b1 - 4 a w b2
                }
              }
              const f32x2 pb = mul2(sg, t);
              rp2 = fma2(pb, wv, rp2);
              unpack2(pb, ps32[c0 + u], ps32[c0 + u + 1]);
              const f32x2 prv = fma2(pb, rr2, mul2(sg, e));
              unpack2(prv, pr32[c0 + u], pr32[c0 + u + 1]);
            }
          }
          {
            float r0, r1;
            tc::unpack2(rp2, r0, r1);
            rp = r0 + r1;
          }
          sm.Rpart[par][uh * kTP + m] = rp;
          // fold the dW2 tile into the fp32 accumulators (row m, columns 32 uh ..)
          if (kFlushPerEval || i == a.Na - 1) {
            { WP_T0
            tc::mbar_wait(&sm.barD[NJ - 1], (phase >> (3 + NJ - 1)) & 1u); phase ^= 1u << (3 + NJ - 1);
            WP_ADD(4) }
            dw_pending = false;
            __syncwarp();
            tc::fence_after_sync();
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            float t[16];
            float t2[16];
            tc::tmem_ld16(tlane + kColDw + 32 * uh + c0, t);            // columns j: A1 (hi|lo rows) x Pbar2 hi
            tc::tmem_ld16(tlane + kColDw + kH + 32 * uh + c0, t2);      // columns 64 + j:           x Pbar2 lo
            tc::tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 16; ++u) sm.DW[(32 * uh + c0 + u) * kTP + m] += t[u] + t2[u];
          }
          }
          if constexpr (kStageAcc) {
            const int cls = sm.T.cls[i];
            if (cls != ps_cls) { flush_ps(); ps_cls = cls; }
#pragma unroll
            for (int u = 0; u < 32; ++u) { psacc[u] += ps32[u]; pracc[u] += pr32[u]; }
          } else {
            const float ps = warp_vec32_sum(ps32, lane);
            const float pr = warp_vec32_sum(pr32, lane);
            sm.S[(w * kMaxClasses + sm.T.cls[i]) * 32 + lane] += ps;
            G.g_wt = fmaf(tprime, ps, G.g_wt);
            G.g_wr += pr;
          }
        }
        tc::fence_before_sync();
        // the partner warp (same trajectories, other unit half) has published its Fpart / Rpart of this nucleus
        bar_sync(kBarPair + (w & 3), 64);
        post(i);
        par ^= 1;
        if (i + 1 < a.Na) pre(i + 1);
      }
      if constexpr (kStageAccR) {
        G.g_b2 += warp_vec32_sum(b2acc, lane);
        G.g_w3 += warp_vec32_sum(w3acc, lane);
      }
      if constexpr (kStageAcc) {
        flush_ps();
        G.g_wr += warp_vec32_sum(pracc, lane);
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
  if (owner && valid && a.ybar0 != nullptr) {
    float* dst = a.ybar0 + row * a.bar_stride;
    dst[0] = xb[0]; dst[1] = xb[1]; dst[2] = xb[2];
    if (NJ >= 2) dst[3] = lb;
    if (NJ == 3) { dst[4] = sb[0]; dst[5] = sb[1]; dst[6] = sb[2]; }
  }
}

// per-tile partial gradient (same layout and reduction kernel as the FFMA path)
__device__ void flush_tile_grad_tc(SmemTcB& sm, TcBwdAccum& G, float* __restrict__ out, const GnnThetaLayout& L,
                                   const float* __restrict__ onehot, int Na, int n_types, uint32_t tmem, bool dw_started) {
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m = 32 * (w & 3) + lane, uh = w >> 2;
  const uint32_t tlane = tmem + ((uint32_t)(32 * (w & 3)) << 16);
  // DW rows 0..63 = (hi k) sums, rows 64..127 = (lo k) sums: dW2[k][j] = DW[k][j] + DW[64+k][j]
  (void)tmem; (void)dw_started; (void)uh;
  sm.red[(0 * 8 + w) * 32 + lane] = G.g_b2; sm.red[(1 * 8 + w) * 32 + lane] = G.g_w3;
  sm.red[(2 * 8 + w) * 32 + lane] = G.g_wt; sm.red[(3 * 8 + w) * 32 + lane] = G.g_wr;
  const float b3w = warp_sum(G.g_b3);
  G.g_b2 = G.g_w3 = G.g_wt = G.g_wr = G.g_b3 = 0.f;
  tc::fence_before_sync();
  bar_sync(kBarCompute, kThreads);
  for (int i = tid; i < kH * kH; i += kThreads) {
    const int k = i / kH, j = i % kH;
    out[L.W2 + i] = sm.DW[j * kTP + k] + sm.DW[j * kTP + kH + k];
  }
  if (tid < kH) {
    // unit j = k = tid: uh = tid >> 5, lane = tid & 31; the four warps (uh*4 .. uh*4+3) hold partial sums over trajectories
    const int u2 = tid >> 5, l2 = tid & 31;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < 4; ++q)
      for (int ww = 0; ww < 4; ++ww) v[q] += sm.red[(q * 8 + u2 * 4 + ww) * 32 + l2];
    out[L.b2 + tid] = v[0]; out[L.w3 + tid] = v[1]; out[L.W1 + tid] = v[2]; out[L.W1 + kH + tid] = v[3];
    float sb1 = 0.f;
    for (int c = 0; c < sm.T.n_cls; ++c) {
      float sn = 0.f;
      for (int ww = 0; ww < 4; ++ww) sn += sm.S[((u2 * 4 + ww) * kMaxClasses + c) * 32 + l2];
      sb1 += sn;
    }
    out[L.b1 + tid] = sb1;
    for (int t = 0; t < n_types; ++t) {
      float s = 0.f;
      for (int c = 0; c < sm.T.n_cls; ++c) {
        float sn = 0.f;
        for (int ww = 0; ww < 4; ++ww) sn += sm.S[((u2 * 4 + ww) * kMaxClasses + c) * 32 + l2];
        s = fmaf(onehot[sm.T.rep[c] * n_types + t], sn, s);
      }
      out[L.W1 + (2 + t) * kH + tid] = s;
    }
  }
  bar_sync(kBarCompute, kThreads);
  if (lane == 0) sm.red[w] = b3w;
  bar_sync(kBarCompute, kThreads);
  if (tid == 0) out[L.b3] = sm.red[0] + sm.red[1] + sm.red[2] + sm.red[3];
  for (int i = tid; i < 8 * kMaxClasses * 32; i += kThreads) sm.S[i] = 0.f;
  bar_sync(kBarCompute, kThreads);
  for (int i = tid; i < kTP * kH; i += kThreads) sm.DW[i] = 0.f;
  bar_sync(kBarCompute, kThreads);
}

template <bool RC>
__global__ void __launch_bounds__(kBwdThreads, 1) gnn_bwd_tc_kernel(const GnnBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemTcB& sm = *reinterpret_cast<SmemTcB*>(smem_raw);
  const int tid = threadIdx.x;
  const GnnThetaLayout L(a.n_types);
  for (int i = tid; i < kH * kH; i += kBwdThreads) {
    const int k = i / kH, j = i % kH;
    float hi, lo;
    tc::split_tf32(a.theta[L.W2 + i], hi, lo);
    const int off = ((j >> 2) * kH + k) * 4 + (j & 3);
    sm.WTh[off] = hi; sm.WTl[off] = lo;
  }
  if (tid < kH) {
    sm.wt[tid] = a.theta[L.W1 + tid];
    sm.wr[tid] = a.theta[L.W1 + kH + tid];
    sm.b2[tid] = a.theta[L.b2 + tid];
    sm.w3[tid] = a.theta[L.w3 + tid];
  }
  load_nuc_tables<kBwdThreads>(sm.T, a.theta, L, a.nuclei, a.onehot, a.Na, a.n_types);
  for (int i = tid; i < 8 * kMaxClasses * 32; i += kBwdThreads) sm.S[i] = 0.f;
  for (int i = tid; i < kTP * kH; i += kBwdThreads) sm.DW[i] = 0.f;
  if (tid == 0) {
    for (int j = 0; j < 3; ++j) { tc::mbar_init(&sm.barB[j], 1); tc::mbar_init(&sm.barD[j], 1); }
    for (int j = 0; j < 3; ++j) tc::mbar_init(&sm.barR[j], 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const long long tiles_a = (a.n_a + kTP - 1) / kTP;
  const long long tiles_b = (a.B - a.n_a + kTP - 1) / kTP;
  // warp roles: warpgroups 0-1 (256 threads) do the CUDA-core work of a tile with (almost) the whole register file; warp 8
  // of warpgroup 2 issues every tcgen05.mma; the register budget moves from the MMA warpgroup to the compute warpgroups
  // (the launch allocates 168 registers x 384 threads = 64 512; 256 x 232 + 128 x 40 uses exactly that pool -- a larger
  // request would make setmaxnreg.inc wait forever)
  static_assert(kThreads * 232 + 128 * 40 <= 168 * kBwdThreads, "setmaxnreg budget exceeds the registers allocated at launch");
  if (tid < kThreads) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    uint32_t phase = 0;            // bits 0..2: parity of barB[jet], bits 3..5: parity of barD[jet], bits 6..8: barR[jet]
    const float b3 = a.theta[0];
    TcBwdAccum G;
    G.g_b2 = G.g_w3 = G.g_wt = G.g_wr = G.g_b3 = 0.f;
    long long tile;
    while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
      long long row0, seg_end; int nj;
      bool seg_b; long long seg_row0;
      if (tile < tiles_a) { row0 = tile * kTP; seg_end = a.n_a; nj = a.jets_a; seg_b = false; seg_row0 = 0; }
      else { row0 = a.n_a + (tile - tiles_a) * kTP; seg_end = a.B; nj = a.jets_b; seg_b = true; seg_row0 = a.n_a; }
      bool dw_pending = false;
      if (nj == 3) bwd_tile_tc<3, RC>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0, tmem, phase, dw_pending);
      else if (nj == 2) bwd_tile_tc<2, RC>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0, tmem, phase, dw_pending);
      else bwd_tile_tc<1, RC>(sm, a, row0, seg_end, b3, G, seg_b, seg_row0, tmem, phase, dw_pending);
      bar_sync(kBarCompute, kThreads);
      flush_tile_grad_tc(sm, G, a.gpart + (size_t)tile * L.n, L, a.onehot, a.Na, a.n_types, tmem, dw_pending);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    long long tile;
    while (next_tile(a.counter, &sm.tile, tiles_a + tiles_b, tile)) {
      if (tid < kThreads + 32) {
        const int nj = tile < tiles_a ? a.jets_a : a.jets_b;
        if (nj == 3) bwd_mma_tile<3, RC>(sm, a, tmem);
        else if (nj == 2) bwd_mma_tile<2, RC>(sm, a, tmem);
        else bwd_mma_tile<1, RC>(sm, a, tmem);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 512);
}

// W2 in the forward GEMM orientation, split hi / lo: [k/4][j][4] (recompute mode: copied into shared memory per evaluation)
__global__ void gnn_w2_split_kernel(const float* __restrict__ theta, float* __restrict__ out, int n_types) {
  const GnnThetaLayout L(n_types);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kH * kH) return;
  const int k = i / kH, j = i % kH;
  float hi, lo;
  tc::split_tf32(theta[L.W2 + i], hi, lo);
  const int off = ((k >> 2) * kH + j) * 4 + (k & 3);
  out[off] = hi; out[kH * kH + off] = lo;
}

int launch_gnn_bwd_tc(void* stream, GnnBwdArgs& a, int grid) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool rc = a.act == nullptr;       // no activation checkpoint: recompute the layer-2 jets on the tensor cores
  cudaError_t e = rc ? cudaFuncSetAttribute(gnn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTcB) + 1024)
                     : cudaFuncSetAttribute(gnn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTcB) + 1024);
  if (e != cudaSuccess) return (int)e;
  if (rc) {
    gnn_w2_split_kernel<<<(kH * kH + 255) / 256, 256, 0, st>>>(a.theta, a.w2split, a.n_types);
    gnn_bwd_tc_kernel<true><<<grid, kBwdThreads, sizeof(SmemTcB) + 1024, st>>>(a);
  } else {
    gnn_bwd_tc_kernel<false><<<grid, kBwdThreads, sizeof(SmemTcB) + 1024, st>>>(a);
  }
  return (int)cudaGetLastError();
}

int launch_gnn_fwd_tc(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes) {
  if (scratch_bytes < kCounterBytes) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  a.counter = (int*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  // 256 threads, one 64 KB operand buffer, two co-resident CTAs per SM whose stage / MMA / epilogue phases interleave
  const long long tiles = (a.n_a + kTP - 1) / kTP + (a.B - a.n_a + kTP - 1) / kTP;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaFuncSetAttribute(gnn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTc1) + 1024);
  if (e != cudaSuccess) return (int)e;
  const int grid = (int)(tiles < 2 * sms ? tiles : 2 * sms);
  gnn_fwd_tc_kernel<2><<<grid, kTP * 2, sizeof(SmemTc1) + 1024, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace ofdft

#ifdef OFDFT_BWD_WAITPROF
extern "C" int ofdft_debug_waitprof(unsigned long long* host, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(host, ofdft::g_waitprof, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ofdft::g_waitprof, z, sizeof(z)); }
  return (int)e;
}
#endif
