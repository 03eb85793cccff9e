// gnn_tc.cu -- tensor-core (tcgen05 / TMEM) forward integrator for the 3-D equivariant GNN flow.
//
// Same contract and results as gnn_fwd_kernel (gnn.cu); the [128 traj x 3 jets] x [64 x 64] layer-2 GEMM of every
// (tile, nucleus, RHS evaluation) runs on the 5th-generation tensor cores as error-compensated 3xTF32:
//     a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo      (hi = top 19 bits, lo = exact remainder; fp32 accumulate in TMEM)
// which holds fp32-level accuracy (measured 1.6e-6 max-relative on the diagnostic GEMM, tests/tc_diag.py).
//
//   * thread (warp w, lane l) owns trajectory m = 32*(w&3) + l -- the TMEM lane its warp may read -- and the unit
//     half uh = w>>2 (32 hidden units); threads with w < 4 also own the RK4 state of trajectory m;
//   * operands are staged by the threads themselves in the canonical no-swizzle K-major UMMA layout
//     ([16-byte chunk of 4 k][row][4 floats]: SBO = 128 B, LBO = rows*16 B), no TMA (operands are produced on chip);
//   * one elected thread issues 24 tcgen05.mma (M=128, N=64, K=8) per jet and commits to an mbarrier; accumulators
//     live in TMEM (3 jets x 64 columns) and come back with tcgen05.ld (32x32b) for the tanh-jet epilogue;
//   * two jet operand buffers (2 x 64 KB): jets 0 and 1 are staged together, jet 2 reuses buffer 0 once the jet-0
//     MMAs have completed.
#include "gnn_common.cuh"
#include "tc.cuh"

namespace ofdft {

constexpr int kChunkBytesA = kTP * 16;      // LBO of the activation operand (128 rows x 16 B)
constexpr int kChunkBytesB = kH * 16;       // LBO of the weight operand (64 rows x 16 B)

struct SmemTc {
  float Abuf[2][2][kH * kTP];        // [jet buffer][hi/lo][16 chunks][128 rows][4]
  float Whi[kH * kH], Wlo[kH * kH];  // B operand: [16 chunks of k][64 rows j][4]
  float wr[kH], wt[kH], b2[kH], w3[kH];
  float cb[kMaxNa * kH];
  float nuc[kMaxNa * 4];
  float xs[3 * kTP];
  float Fpart[2 * 3 * kTP];
  uint64_t bar[3];
  unsigned int tmem_slot;
  int tile;
};

// stage one jet (32 units of one trajectory) as hi/lo tf32 operands
__device__ __forceinline__ void stage_jet(float* __restrict__ hi, float* __restrict__ lo, const float (&v)[32], int m, int uh) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float4 h, l;
    tc::split_tf32(v[4 * c + 0], h.x, l.x); tc::split_tf32(v[4 * c + 1], h.y, l.y);
    tc::split_tf32(v[4 * c + 2], h.z, l.z); tc::split_tf32(v[4 * c + 3], h.w, l.w);
    const int off = ((8 * uh + c) * kTP + m) * 4;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

// D[jet] = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo  (24 MMAs), issued by one elected lane of a converged warp.
// The four base descriptors are built once; a K step of 8 (two 16-byte chunks) only advances the address field.
__device__ __forceinline__ void issue_jet_mma(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t w_hi, uint64_t w_lo,
                                              uint32_t idesc) {
  constexpr uint64_t stepA = (2 * kChunkBytesA) >> 4, stepB = (2 * kChunkBytesB) >> 4;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const uint64_t ab = p == 1 ? a_lo : a_hi;
    const uint64_t wb = p == 2 ? w_lo : w_hi;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) tc::mma_tf32(d_tmem, ab + ks * stepA, wb + ks * stepB, idesc, (p | ks) != 0);
  }
}

template <int NJ>
__device__ void fwd_tile_tc(SmemTc& sm, const GnnFwdArgs& a, long long row0, long long seg_end, float b3,
                            bool seg_b, long long seg_row0, uint32_t tmem, uint32_t (&phase)[3]) {
  const ActLayout AL(a.act_rows_a, a.act_rows_b, a.jets_a, a.jets_b);
  const long long act_rows = seg_b ? a.act_rows_b : a.act_rows_a;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m = 32 * (w & 3) + lane, uh = w >> 2;
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
  const uint64_t dA00 = tc::make_desc(tc::smem_u32(sm.Abuf[0][0]), kChunkBytesA, 128);
  const uint64_t dA01 = tc::make_desc(tc::smem_u32(sm.Abuf[0][1]), kChunkBytesA, 128);
  const uint64_t dA10 = tc::make_desc(tc::smem_u32(sm.Abuf[1][0]), kChunkBytesA, 128);
  const uint64_t dA11 = tc::make_desc(tc::smem_u32(sm.Abuf[1][1]), kChunkBytesA, 128);
  const uint64_t dWh = tc::make_desc(tc::smem_u32(sm.Whi), kChunkBytesB, 128);
  const uint64_t dWl = tc::make_desc(tc::smem_u32(sm.Wlo), kChunkBytesB, 128);

  auto combine = [&](int i) {
    const float* nc = sm.nuc + i * 4;
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
      for (int i = 0; i < a.Na; ++i) {
        if (owner && i > 0) combine(i - 1);
        // ---- first layer for (trajectory m, units 32*uh ..): a = tanh(w_r r + c), a' = (1-a^2) w_r, a'' = -2 a a' w_r ----
        float av[32];
        {
          const float* nc = sm.nuc + i * 4;
          const float d0 = sm.xs[m] - nc[0], d1 = sm.xs[kTP + m] - nc[1], d2 = sm.xs[2 * kTP + m] - nc[2];
          const float r = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2)));
          const float* cbi = sm.cb + i * kH + 32 * uh;
#pragma unroll
          for (int u = 0; u < 32; ++u)
            av[u] = tanh_f(fmaf(sm.wr[32 * uh + u], r, fmaf(tprime, sm.wt[32 * uh + u], cbi[u])));
        }
        stage_jet(sm.Abuf[0][0], sm.Abuf[0][1], av, m, uh);
        if (NJ >= 2) {
          float a1[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) a1[u] = fmaf(-av[u], av[u], 1.0f) * sm.wr[32 * uh + u];
          stage_jet(sm.Abuf[1][0], sm.Abuf[1][1], a1, m, uh);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (w == 0) {
          tc::fence_after_sync();
          if (tc::elect_one()) {
            issue_jet_mma(tmem + 0 * kH, dA00, dA01, dWh, dWl, idesc);
            tc::commit(&sm.bar[0]);
            if (NJ >= 2) {
              issue_jet_mma(tmem + 1 * kH, dA10, dA11, dWh, dWl, idesc);
              tc::commit(&sm.bar[1]);
            }
          }
          __syncwarp();
        }
        if (NJ == 3) {
          float a2[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) {
            const float wv = sm.wr[32 * uh + u];
            a2[u] = -2.0f * av[u] * fmaf(-av[u], av[u], 1.0f) * wv * wv;
          }
          tc::mbar_wait(&sm.bar[0], phase[0]);           // jet-0 MMAs have finished reading buffer 0
          __syncwarp();
          stage_jet(sm.Abuf[0][0], sm.Abuf[0][1], a2, m, uh);
          tc::fence_async_smem();
          __syncthreads();
          if (w == 0) {
            tc::fence_after_sync();
            if (tc::elect_one()) {
              issue_jet_mma(tmem + 2 * kH, dA00, dA01, dWh, dWl, idesc);
              tc::commit(&sm.bar[2]);
            }
            __syncwarp();
          }
        }
        // ---- wait for the last commit (commits complete in order), then the tanh-jet epilogue from TMEM ----
        if (NJ == 3) { tc::mbar_wait(&sm.bar[2], phase[2]); phase[2] ^= 1; phase[0] ^= 1; tc::mbar_wait(&sm.bar[1], phase[1]); phase[1] ^= 1; }
        else if (NJ == 2) { tc::mbar_wait(&sm.bar[1], phase[1]); phase[1] ^= 1; tc::mbar_wait(&sm.bar[0], phase[0]); phase[0] ^= 1; }
        else { tc::mbar_wait(&sm.bar[0], phase[0]); phase[0] ^= 1; }
        __syncwarp();                                   // tcgen05.ld is .sync.aligned: reconverge after the spin loops
        tc::fence_after_sync();
        float fp[3] = {0.f, 0.f, 0.f};
        float* actp = nullptr;
        if (a.act != nullptr)
          actp = a.act + AL.seg_base((long long)(step * 4 + stage) * a.Na + i, seg_b) + (row0 - seg_row0)
                 + (long long)(32 * uh) * act_rows + m;
        const long long jet_stride = (long long)kH * act_rows;
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          float p0[16], p1[16], p2[16];
          tc::tmem_ld16(tlane + 0 * kH + 32 * uh + c0, p0);
          if (NJ >= 2) tc::tmem_ld16(tlane + 1 * kH + 32 * uh + c0, p1);
          if (NJ == 3) tc::tmem_ld16(tlane + 2 * kH + 32 * uh + c0, p2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = 32 * uh + c0 + u;
            const float wv = sm.w3[j];
            const float aa = tanh_f(p0[u] + sm.b2[j]);
            fp[0] = fmaf(wv, aa, fp[0]);
            if (NJ >= 2) {
              const float sg = fmaf(-aa, aa, 1.0f);
              const float a1 = sg * p1[u];
              fp[1] = fmaf(wv, a1, fp[1]);
              if (NJ == 3) fp[2] = fmaf(wv, fmaf(sg, p2[u], -2.0f * aa * a1 * p1[u]), fp[2]);
            }
            if (actp != nullptr) {
              float* dst = actp + (long long)(c0 + u) * act_rows;
              __stcs(dst, aa);
              if (NJ >= 2) __stcs(dst + jet_stride, p1[u]);
              if (NJ == 3) __stcs(dst + 2 * jet_stride, p2[u]);
            }
          }
        }
#pragma unroll
        for (int jt = 0; jt < NJ; ++jt) sm.Fpart[(uh * 3 + jt) * kTP + m] = fp[jt];
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

__global__ void __launch_bounds__(kThreads, 1) gnn_fwd_tc_kernel(const GnnFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemTc& sm = *reinterpret_cast<SmemTc*>(smem_raw);
  const int tid = threadIdx.x;
  const GnnThetaLayout L(a.n_types);
  // weight operand B[n = j][k] = W2[k][j], split hi/lo, layout [k/4][j][k%4]
  for (int i = tid; i < kH * kH; i += kThreads) {
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
  for (int i = tid; i < a.Na * kH; i += kThreads) {
    const int n = i / kH, k = i % kH;
    float c = a.theta[L.b1 + k];
    for (int t = 0; t < a.n_types; ++t) c = fmaf(a.onehot[n * a.n_types + t], a.theta[L.W1 + (2 + t) * kH + k], c);
    sm.cb[i] = c;
  }
  if (tid < a.Na * 3) sm.nuc[(tid / 3) * 4 + (tid % 3)] = a.nuclei[tid];
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
    if (nj == 3) fwd_tile_tc<3>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
    else if (nj == 2) fwd_tile_tc<2>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
    else fwd_tile_tc<1>(sm, a, row0, seg_end, b3, seg_b, seg_row0, tmem, phase);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 256);
}

int launch_gnn_fwd_tc(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes) {
  if (scratch_bytes < kCounterBytes) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  a.counter = (int*)scratch;
  cudaError_t e = cudaMemsetAsync(scratch, 0, kCounterBytes, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(gnn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTc) + 1024);
  if (e != cudaSuccess) return (int)e;
  const long long tiles = (a.n_a + kTP - 1) / kTP + (a.B - a.n_a + kTP - 1) / kTP;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)(tiles < sms ? tiles : sms);
  gnn_fwd_tc_kernel<<<grid, kThreads, sizeof(SmemTc) + 1024, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace ofdft
