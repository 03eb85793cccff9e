// mlp1d_tc.cu -- tensor-core (tcgen05 / TMEM) forward integrator of the 1-D tanh-MLP flow
// (reference: ofdft_normflows/cn_flows.py:117-182 evaluated through jax_ode.py:9-110; LiH.py:64-65 uses 3 x 512).
//
// A 128-trajectory tile of a 512-wide layer does not fit on chip with its three jets (768 KB of activations), so the
// integrator is a sequence of per-layer launches whose activations stay L2 resident (6 MB per layer at B = 1024):
//   glue kernel  (one warp per trajectory): last layer + RHS + RK4 bookkeeping of evaluation e, stage checkpoint and
//                first layer of evaluation e+1;
//   layer kernel (one CTA per 128 trajectories x 64 output units): P[jet] = A[jet] . W_l for the jets (a, a', a'')
//                as error-compensated 3xTF32 tcgen05 GEMMs with fp32 accumulators in tensor memory, tanh-jet epilogue.
// Layer buffers hold (a, p', p'') row-major [jet][row][unit]; the operand jets (a, a' = (1-a^2) p', a'' = ...) and their
// tf32 remainders are built while staging (float4 loads -> float4 stores into the K-major no-swizzle UMMA layout).
// Operands go in unrounded: the tensor core truncates to tf32 (tests/tc_trunc_probe.py).  The TMEM accumulation chain is
// kept short (two alternating accumulator sets, drained into fp32 registers every 128 k) because it truncates too.
#include "common.cuh"
#include "mlp1d_common.cuh"
#include "tc.cuh"

namespace ofdft {
namespace mlp {

// Programmatic dependent launch: every kernel of the per-layer chain is launched with the "programmatic stream
// serialization" attribute, calls pdl_wait() before it touches anything its predecessor wrote and pdl_trigger() once its
// successor may start being scheduled -- so a kernel's prologue (TMEM allocation, barrier init, scheduling) overlaps the
// tail of the one before it.  ~380 dependent launches per 1-D step make the fixed cost per launch matter.
#ifdef OFDFT_GEMM_TRACE     // development only: %globaltimer stamps of CTA 0 of every kernel of the chain
__device__ unsigned long long g_trace[1 << 16];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TRACE_BEGIN(kind) unsigned int tr_ = 0xffffffffu; if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == (kind == 0 ? 32 * kSW : 0)) { tr_ = atomicAdd(&g_trace_n, 1u) & 8191u; g_trace[tr_ * 8] = kind; g_trace[tr_ * 8 + 1] = gtime(); }
#define TRACE(i) if (tr_ != 0xffffffffu) g_trace[tr_ * 8 + (i)] = gtime();
#else
#define TRACE_BEGIN(kind)
#define TRACE(i)
#endif
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename Arg>
static cudaError_t launch_pdl(void (*kernel)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg& arg) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem)), "l"(gmem) : "memory");
}

constexpr int kGemmKC = 16;                       // k per pipeline stage (4 chunks of 4)
#ifndef OFDFT_GEMM_SG
#define OFDFT_GEMM_SG 2
#endif
static_assert(512 / 16 <= 32, "one TMEM accumulation chain per launch assumes H <= 512");
constexpr int kGemmGroup = 32;                    // stages per TMEM accumulation chain (256 k, 96 MMAs per jet); measured max
                                                  // state error vs the float64 oracle (3x512): 6e-7 / 1.1e-6 / 1.7e-6 for 8 / 16 / 32
constexpr int kGemmSlots = 3;                     // activation-operand ring in tensor memory
constexpr int kGemmRaw = 4;                       // shared-memory ring fed by TMA bulk copies: raw activation jets + weight stage
constexpr int kLboB = 64 * 16;                    // chunk stride of the weight operand (bytes): 64 rows x 16 B, as TMA lays it
constexpr int kColAcc = 0;                        // TMEM: accumulators 3 jets x 64 columns
constexpr int kColOp = 192;                       // TMEM: operand ring, per slot [jet][hi 16 | lo 16] = 96 columns
constexpr int kSG = OFDFT_GEMM_SG;                // staging groups of 8 warps (4 row quadrants x 2 chunk halves): group g stages the
                                                  // pipeline stages it = g (mod kSG), so kSG stages are in flight on the CUDA cores --
                                                  // one group alone is latency-bound at ~950 clocks per stage against 576 of MMA
constexpr int kSW = 8 * kSG;                      // staging warps
constexpr int kCT = 64 / (kSW / 4);               // output columns per staging thread (epilogue: all staging warps)
constexpr int kCH = 2;                            // 4-wide k chunks per staging thread and stage
constexpr int kGemmThreads = (kSW + 2) * 32;      // staging warps + MMA warp + TMA warp

// layer buffers: [jet: a, p', p''][unit / 4][row (padded to 128)][4 units] -- 32 consecutive rows x 4 units are 512
// contiguous bytes, which is what a warp whose lanes are rows (the TMEM lane constraint) loads or stores at once
__host__ __device__ __forceinline__ size_t sbuf_off(int unit_group, long long row, long long rows_pad) {
  return ((size_t)unit_group * rows_pad + row) * 4;
}

struct GemmSmem {
  float4 A[kGemmRaw][3][4][128];                  // [slot][jet][k chunk][row]: (a, p', p'') or cotangent jets as stored
  float W[kGemmRaw][2][4 * 64 * 4];               // [slot][hi/lo][k chunk][row n 64][4]: canonical K-major UMMA operand
  float red[1024];                                // adjoint epilogues: cross-warp column sums / x partials
  uint64_t rfull[kGemmRaw], rempty[kGemmRaw];     // TMA -> consumers (tx bytes) / consumers -> TMA (16 stager warps + 1 commit)
  uint64_t full[kGemmSlots], empty[kGemmSlots], accf, acce;
  unsigned int tmem_slot;
};

struct GemmArgs {
  const float* Sin; float* Sout;                  // layer buffers (A operand source / epilogue destination)
  const float* WT; const float* WTlo;             // B operand, interleaved [k/4][n][4] (unrounded) and its tf32 remainder
  const float* bias;                              // MODE 0: [H] (theta offset: may be unaligned -> scalar loads)
  long long B, rows_pad; int H;                   // B: valid rows of this chunk
  // adjoint sweep only
  float* blk_out; size_t blk_tile_stride;         // HBM blocks [tile of 8 rows][..][jet][H][8] the weight-gradient GEMM reads
  const float* Sprev;                             // MODE 1/2: (a, p', p'') of the layer whose tanh jet is reversed
  float* gsm; int n_small, off_b, off_w0;         // per-row-tile small-gradient partials (compact layout), offsets
  const float* wx; const float* xs; float tprime; // MODE 2: first-layer kernel row, stage states x of the chunk rows
  float* xpart;                                   // MODE 2: [H/64][rows_pad] partial x cotangents
};

// Layer GEMM P[jet] = A[jet] . W for the jets (a, a', a'') of 128 trajectories x 64 output units, 3xTF32 on tcgen05.
// An SS-mode M128.N64.K8 tf32 MMA reads 6 KB of operands per 32 clocks -- 192 B/clk against the 128 B/clk of the
// shared-memory port -- so the activation operand goes through TENSOR MEMORY (TS mode): warp w of the 16 staging warps
// owns rows 32 (w & 3).. (the TMEM lanes it may access) and the 4-wide k chunk w >> 2; each thread loads (a, p', p'')
// of its (row, chunk), builds the jets and their tf32 remainders and writes them with tcgen05.st; only the 8 KB
// weight stage goes through shared memory (cp.async, two stages ahead).  Warp 16 issues the MMAs.
//   full[s]  : staging warps -> MMA warp (16 arrivals)      empty[s] : tcgen05.commit -> stagers
//   accf     : tcgen05.commit -> stagers (chain finished)   acce     : stagers -> MMA warp (16 arrivals: chain drained)
// The accumulation chain is drained into fp32 registers every kGemmGroup stages because it truncates.
// MODE 0: forward layer (A jets built from (a, p', p''), epilogue tanh);  MODE 1: adjoint through a hidden layer
// (A = cotangent jets as they are, B = the kernel itself, epilogue = reverse of the previous layer's tanh jet);
// MODE 2: the same into the first layer (epilogue = first-layer gradients and the x cotangent).
template <int NJ, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1) mlp_layer_gemm_tc_kernel(const GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int H = g.H;
  const long long row0 = (long long)blockIdx.x * 128;
  const int j0 = blockIdx.y * 64;
  if (tid == 0) {
    for (int i = 0; i < kGemmSlots; ++i) { tc::mbar_init(&sm.full[i], 8); tc::mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < kGemmRaw; ++i) { tc::mbar_init(&sm.rfull[i], 1); tc::mbar_init(&sm.rempty[i], 8 + 1); }
    tc::mbar_init(&sm.accf, 1); tc::mbar_init(&sm.acce, kSW);
    tc::fence_mbar_init();
  }
  if (w == 0) tc::tmem_alloc(&sm.tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const int n_it = H / kGemmKC;
  const size_t jet_stride = (size_t)g.rows_pad * H;
  TRACE_BEGIN(0)
  pdl_wait();                                     // everything above overlapped the previous kernel's tail
  TRACE(2)

  if (w == kSW) {
    // ---------------- MMA issue warp ----------------
    const uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
    uint32_t ph_full = 0, ph_acce = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % kGemmSlots;
      const bool first = (it % kGemmGroup) == 0;
      tc::mbar_wait(&sm.full[s], (ph_full >> s) & 1u); ph_full ^= 1u << s;
      if (it == 0) { TRACE(3) }
      if (first && it > 0) { tc::mbar_wait(&sm.acce, ph_acce); ph_acce ^= 1u; }
      __syncwarp();
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint64_t wh = tc::make_desc(tc::smem_u32(sm.W[it % kGemmRaw][0]), kLboB, 128);
        const uint64_t wl = tc::make_desc(tc::smem_u32(sm.W[it % kGemmRaw][1]), kLboB, 128);
#pragma unroll
        for (int jt = 0; jt < NJ; ++jt) {
          const uint32_t ahi = tmem + kColOp + 96 * s + 32 * jt, alo = ahi + 16;
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              tc::mma_tf32_ts(tmem + kColAcc + jt * 64, (p == 1 ? alo : ahi) + 8 * ks, (p == 2 ? wl : wh) + ks * ((2 * kLboB) >> 4),
                              idesc, !(first && p == 0 && ks == 0));
        }
        tc::commit(&sm.empty[s]);
        tc::commit(&sm.rempty[it % kGemmRaw]);                 // the weight stage of this ring slot has been read
        if ((it + 1) % kGemmGroup == 0 || it + 1 == n_it) tc::commit(&sm.accf);
      }
      __syncwarp();
    }
    TRACE(4)
#ifdef OFDFT_GEMM_TRACE
    if (tr_ != 0xffffffffu) { tc::mbar_wait(&sm.accf, 0); TRACE(5) }
#endif
  } else if (w == kSW + 1) {
    // ---------------- TMA warp: bulk copies global/L2 -> shared-memory ring ----------------
    // per stage: 3 jets x 4 unit groups x (128 rows x 16 B = 2 KB contiguous in the interleaved layer buffer) and
    // (hi, lo) x 4 k groups x (64 rows x 16 B = 1 KB contiguous in the interleaved weight copy)
    if (tc::elect_one()) {
      uint32_t ph = 0;
      for (int it = 0; it < n_it; ++it) {
        const int s = it % kGemmRaw;
        if (it >= kGemmRaw) { tc::mbar_wait(&sm.rempty[s], (ph >> s) & 1u); ph ^= 1u << s; }
#ifdef OFDFT_ABL_NOTMA_A      // timing ablation only (results invalid): the activation operand is not fetched
        tc::mbar_arrive_expect_tx(&sm.rfull[s], (uint32_t)(2 * 4 * 1024));
#else
        tc::mbar_arrive_expect_tx(&sm.rfull[s], (uint32_t)(NJ * 4 * 2048 + 2 * 4 * 1024));
#pragma unroll
        for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tc::bulk_g2s(&sm.A[s][jt][c][0], g.Sin + jt * jet_stride + sbuf_off(4 * it + c, row0, g.rows_pad), 2048, &sm.rfull[s]);
#endif
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const size_t off = ((size_t)(4 * it + c) * H + j0) * 4;
          tc::bulk_g2s(&sm.W[s][0][c * 256], g.WT + off, 1024, &sm.rfull[s]);
          tc::bulk_g2s(&sm.W[s][1][c * 256], g.WTlo + off, 1024, &sm.rfull[s]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- staging warps ----------------
    // row (TMEM lane); staging: group grp handles stages it = grp (mod kSG), chunk half cs; epilogue: column group c
    const int r = 32 * (w & 3) + lane, c = w >> 2, cs = (w >> 2) & 1, grp = w >> 3;
    const uint32_t tl = tmem + ((uint32_t)(32 * (w & 3)) << 16);
    struct Regs { float4 a[3][kCH]; };
    auto load_regs = [&](Regs& R, int it) {
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
        for (int q = 0; q < kCH; ++q) R.a[jt][q] = sm.A[it % kGemmRaw][jt][kCH * cs + q][r];
    };
    // operand jets and their tf32 remainders, two values per FP32 instruction (packed FFMA2 / FMUL2)
    auto lo4 = [&](tc::f32x2 xy, tc::f32x2 zw, uint32_t col) {
      using namespace tc;
      const f32x2 m1 = splat2(-1.0f);
      const f32x2 l0 = fma2(xy & 0xFFFFE000FFFFE000ull, m1, xy), l1 = fma2(zw & 0xFFFFE000FFFFE000ull, m1, zw);   // x - trunc_tf32(x)
      tmem_st4(col, lo2(xy), hi2(xy), lo2(zw), hi2(zw));
      tmem_st4(col + 16, lo2(l0), hi2(l0), lo2(l1), hi2(l1));
    };
    auto store_tmem = [&](const Regs& R, int s) {
      using namespace tc;
#pragma unroll
      for (int q = 0; q < kCH; ++q) {
        const uint32_t base = tl + kColOp + 96 * s + 4 * (kCH * cs + q);
        const float4 a = R.a[0][q];
        const f32x2 a01 = pack2(a.x, a.y), a23 = pack2(a.z, a.w);
        lo4(a01, a23, base);
        if (MODE != 0) {
#pragma unroll
          for (int jt = 1; jt < NJ; ++jt) {
            const float4 v = R.a[jt][q];
            lo4(pack2(v.x, v.y), pack2(v.z, v.w), base + 32 * jt);
          }
        } else if (NJ >= 2) {
          const float4 p1 = R.a[1][q];
          const f32x2 p01 = pack2(p1.x, p1.y), p23 = pack2(p1.z, p1.w);
          const f32x2 one = splat2(1.0f), m1 = splat2(-1.0f);
          const f32x2 n01 = mul2(a01, m1), n23 = mul2(a23, m1);
          const f32x2 sg01 = fma2(n01, a01, one), sg23 = fma2(n23, a23, one);           // 1 - a^2
          const f32x2 b01 = mul2(sg01, p01), b23 = mul2(sg23, p23);                     // a' = (1 - a^2) p'
          lo4(b01, b23, base + 32);
          if (NJ == 3) {
            const float4 p2 = R.a[2][q];
            const f32x2 two = splat2(2.0f);
            // a'' = (1 - a^2) p'' - 2 a a' p'
            const f32x2 c01 = fma2(sg01, pack2(p2.x, p2.y), mul2(mul2(mul2(n01, two), b01), p01));
            const f32x2 c23 = fma2(sg23, pack2(p2.z, p2.w), mul2(mul2(mul2(n23, two), b23), p23));
            lo4(c01, c23, base + 64);
          }
        }
      }
    };
    // accumulators: warp w <-> rows 32 (w & 3).. and the kCT output columns kCT (w >> 2)..
    float acc[NJ][kCT];
#pragma unroll
    for (int jt = 0; jt < NJ; ++jt)
#pragma unroll
      for (int i = 0; i < kCT; ++i) acc[jt][i] = 0.f;
    uint32_t ph_accf = 0;
    auto drain = [&]() {
      tc::mbar_wait(&sm.accf, ph_accf); ph_accf ^= 1u;
      __syncwarp();
      tc::fence_after_sync();
#pragma unroll
      for (int jt = 0; jt < NJ; ++jt) {
#pragma unroll
        for (int c0 = 0; c0 < kCT; c0 += 16) {
          float t[16];
          tc::tmem_ld16(tl + kColAcc + jt * 64 + kCT * c + c0, t);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[jt][c0 + i] += t[i];
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sm.acce);
    };
    // one accumulation chain per launch (n_it <= kGemmGroup because H <= 512): no drain inside the loop
    for (int it = grp; it < n_it; it += kSG) {
      const int s = it % kGemmSlots, rs = it % kGemmRaw;
      tc::mbar_wait(&sm.rfull[rs], (uint32_t)(it / kGemmRaw) & 1u);                       // TMA data of this stage has landed
      Regs R;
      load_regs(R, it);
      // the TMEM slot of this stage is free once the MMAs of stage it-3 have completed
      if (it >= kGemmSlots) tc::mbar_wait(&sm.empty[s], (uint32_t)(it / kGemmSlots - 1) & 1u);
      __syncwarp();
      tc::fence_after_sync();
      store_tmem(R, s);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(&sm.full[s]); tc::mbar_arrive(&sm.rempty[rs]); }
    }
    drain();
    pdl_trigger();                                // the next kernel of the chain may be scheduled during this epilogue

    const long long row = row0 + r;
    const int jc = j0 + kCT * c;
    const long long rows8 = (g.B + 7) / 8 * 8;
    // HBM block slot of (row, unit): [tile = row / 8][jet][unit][row % 8]
    float* blk = g.blk_out != nullptr ? g.blk_out + (size_t)(row >> 3) * g.blk_tile_stride + (size_t)jc * kTT + (row & 7) : nullptr;
    if (MODE == 0) {
      // ---- epilogue: bias, tanh, store (a, p', p'') ----
      if (row < rows8) {
        float* dst = g.Sout + sbuf_off(jc >> 2, row, g.rows_pad);
#pragma unroll
        for (int i = 0; i < kCT; i += 4) {
          float4 a;
          a.x = tanh_f(acc[0][i] + g.bias[jc + i]); a.y = tanh_f(acc[0][i + 1] + g.bias[jc + i + 1]);
          a.z = tanh_f(acc[0][i + 2] + g.bias[jc + i + 2]); a.w = tanh_f(acc[0][i + 3] + g.bias[jc + i + 3]);
          if (row < g.B) {
            float* d = dst + sbuf_off(i >> 2, 0, g.rows_pad);
            *reinterpret_cast<float4*>(d) = a;
            if (NJ >= 2) *reinterpret_cast<float4*>(d + jet_stride) = make_float4(acc[1][i], acc[1][i + 1], acc[1][i + 2], acc[1][i + 3]);
            if (NJ == 3) *reinterpret_cast<float4*>(d + 2 * jet_stride) = make_float4(acc[2][i], acc[2][i + 1], acc[2][i + 2], acc[2][i + 3]);
          }
          if (blk != nullptr) {
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              blk[(size_t)(i + e) * kTT] = av[e];
              if (NJ >= 2) blk[(size_t)(g.H + i + e) * kTT] = acc[1][i + e];
              if (NJ == 3) blk[(size_t)(2 * g.H + i + e) * kTT] = acc[NJ - 1][i + e];
            }
          }
        }
      }
    } else {
      // ---- epilogue: reverse of the tanh jet of the layer below: (abar, abar', abar'') -> (pbar, pbar', pbar'') ----
      float q0[kCT], q1[kCT], q2[kCT];
      const float* sp = g.Sprev + sbuf_off(jc >> 2, row, g.rows_pad);
#pragma unroll
      for (int i = 0; i < kCT; i += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(sp + sbuf_off(i >> 2, 0, g.rows_pad));
        const float4 p14 = *reinterpret_cast<const float4*>(sp + sbuf_off(i >> 2, 0, g.rows_pad) + jet_stride);
        const float4 p24 = *reinterpret_cast<const float4*>(sp + sbuf_off(i >> 2, 0, g.rows_pad) + 2 * jet_stride);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, p1[4] = {p14.x, p14.y, p14.z, p14.w}, p2[4] = {p24.x, p24.y, p24.z, p24.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sg = fmaf(-av[e], av[e], 1.0f);
          const float ab = acc[0][i + e], a1b = acc[1][i + e], a2b = acc[NJ - 1][i + e];
          q2[i + e] = a2b * sg;
          q1[i + e] = sg * fmaf(-4.0f * av[e] * p1[e], a2b, a1b);
          q0[i + e] = sg * (ab - 2.0f * av[e] * p1[e] * a1b +
                            a2b * fmaf(-2.0f * av[e], p2[e], -2.0f * fmaf(-3.0f * av[e], av[e], 1.0f) * p1[e] * p1[e]));
        }
      }
      float* red = sm.red;                       // [4 row quadrants][64 units] (+ MODE 2: second quantity, x partials)
      if (MODE == 1) {
        if (row < rows8) {
          float* dst = g.Sout + sbuf_off(jc >> 2, row, g.rows_pad);
#pragma unroll
          for (int i = 0; i < kCT; i += 4) {
            float* d = dst + sbuf_off(i >> 2, 0, g.rows_pad);
            if (row < g.B) {
              *reinterpret_cast<float4*>(d) = make_float4(q0[i], q0[i + 1], q0[i + 2], q0[i + 3]);
              *reinterpret_cast<float4*>(d + jet_stride) = make_float4(q1[i], q1[i + 1], q1[i + 2], q1[i + 3]);
              *reinterpret_cast<float4*>(d + 2 * jet_stride) = make_float4(q2[i], q2[i + 1], q2[i + 2], q2[i + 3]);
            }
            if (blk != nullptr) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                blk[(size_t)(i + e) * kTT] = q0[i + e];
                blk[(size_t)(g.H + i + e) * kTT] = q1[i + e];
                blk[(size_t)(2 * g.H + i + e) * kTT] = q2[i + e];
              }
            }
          }
        }
        // bias gradient of the layer below: column sums over this tile's rows (rows >= B hold zeros)
#pragma unroll
        for (int i = 0; i < kCT; ++i) {
          const float v = warp_sum(q0[i]);
          if (lane == 0) red[(w & 3) * 64 + kCT * c + i] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSW * 32) : "memory");
        if (tid < 64) {
          float* gs = g.gsm + (size_t)blockIdx.x * g.n_small + g.off_b + j0 + tid;
          *gs += (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
        }
      } else {
        // first layer: p = w_x x + t' w_t + b1, p' = w_x (a parameter), p'' = 0
        const float x_row = row < g.B ? g.xs[row] : 0.f;
        float xp = 0.f;
#pragma unroll
        for (int i = 0; i < kCT; ++i) {
          xp = fmaf(q0[i], g.wx[jc + i], xp);
          const float v0 = warp_sum(q0[i]);
          const float v1 = warp_sum(fmaf(q0[i], x_row, q1[i]));
          if (lane == 0) { red[(w & 3) * 64 + kCT * c + i] = v0; red[256 + (w & 3) * 64 + kCT * c + i] = v1; }
        }
        red[512 + c * 128 + r] = xp;
        asm volatile("bar.sync 1, %0;" ::"n"(kSW * 32) : "memory");
        if (tid < 64) {
          const float sb = (red[tid] + red[64 + tid]) + (red[128 + tid] + red[192 + tid]);
          const float sx = (red[256 + tid] + red[320 + tid]) + (red[384 + tid] + red[448 + tid]);
          float* gs = g.gsm + (size_t)blockIdx.x * g.n_small;
          gs[g.off_b + j0 + tid] += sb;                        // b1
          gs[g.off_w0 + j0 + tid] = fmaf(g.tprime, sb, gs[g.off_w0 + j0 + tid]);     // w_t row
          gs[g.off_w0 + g.H + j0 + tid] += sx;                 // w_x row
        }
        if (tid < 128) {
          float xs = 0.f;
#pragma unroll
          for (int q = 0; q < kSW / 4; ++q) xs += red[512 + q * 128 + tid];
          g.xpart[(size_t)blockIdx.y * g.rows_pad + row0 + tid] = xs;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  TRACE(6)
  if (w == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// glue: one warp per trajectory.  eval >= 0: last layer + RHS + RK4 update of evaluation `eval` from the last hidden
// layer's buffer; then (if another evaluation follows) the stage checkpoint and first layer of evaluation eval+1.
// ------------------------------------------------------------------------------------------------
struct GlueArgs {
  const float* theta; const float* Slast; float* S0; float* state;
  const float* y_in; float* y_out; float* ckpt;
  long long B, rows_pad; int stride, H, L, n_steps, eval; float t0, t1, y0;
  float* acts0; size_t act_tile_stride;           // activation checkpoint: HBM blocks of layer 0 of evaluation eval+1 (or null)
};

// Thread mapping of the row kernels: one CTA per tile of 8 trajectories, thread = (row t = tid & 7, unit-group slot
// tid >> 3); a warp then touches 4 unit groups x (8 rows x 16 B = 128 contiguous bytes) of an interleaved layer buffer and
// 32-byte runs of an HBM block.  Sums over units: shuffles across the 4 lanes of a row, then across the 8 warps through
// shared memory; sums over the 8 rows of a tile: shuffles across lanes 0..7.
__device__ __forceinline__ float row_lanes_sum(float v) {      // over the 4 lanes that share a row within a warp
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float tile_rows_sum(float v) {      // over the 8 rows of the tile (lanes t = 0..7)
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <int NJ>
__global__ void __launch_bounds__(256) mlp_glue_tc_kernel(const GlueArgs g) {
  __shared__ float part[3][8][8];                 // [quantity][warp][row]
  __shared__ float bc[8][2];                      // per row: x and s of the next stage
  pdl_trigger();                                  // short kernel: let the next layer GEMM run its prologue underneath
  TRACE_BEGIN(1)
  pdl_wait();
  TRACE(2)
  const int tid = threadIdx.x, t = tid & 7, gs = tid >> 3, wp = tid >> 5;
  const long long row = (long long)blockIdx.x * 8 + t;
  const bool valid = row < g.B;
  const long long rowc = valid ? row : g.B - 1;
  const int H = g.H;
  const ThetaLayout TL(H, g.L);
  const size_t jet_stride = (size_t)g.rows_pad * H;
  const float h = (g.t1 - g.t0) / (float)g.n_steps;
  if (g.eval >= 0) {
    float hp[3] = {0.f, 0.f, 0.f};
    for (int grp = gs; grp < H / 4; grp += 32) {
      const int k = 4 * grp;
      const float* s0 = g.Slast + sbuf_off(grp, rowc, g.rows_pad);
      const float4 a = *reinterpret_cast<const float4*>(s0);
      float4 p1 = make_float4(0, 0, 0, 0), p2 = make_float4(0, 0, 0, 0);
      if (NJ >= 2) p1 = *reinterpret_cast<const float4*>(s0 + jet_stride);
      if (NJ == 3) p2 = *reinterpret_cast<const float4*>(s0 + 2 * jet_stride);
      const float av[4] = {a.x, a.y, a.z, a.w}, q1[4] = {p1.x, p1.y, p1.z, p1.w}, q2[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float wl = g.theta[TL.w_last() + k + e];
        const float sg = fmaf(-av[e], av[e], 1.0f);
        const float a1 = sg * q1[e];
        hp[0] = fmaf(wl, av[e], hp[0]);
        if (NJ >= 2) hp[1] = fmaf(wl, a1, hp[1]);
        if (NJ == 3) hp[2] = fmaf(wl, fmaf(sg, q2[e], -2.0f * av[e] * a1 * q1[e]), hp[2]);
      }
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float v = row_lanes_sum(hp[q]);
      if ((tid & 31) < 8) part[q][wp][t] = v;
    }
  }
  __syncthreads();
  if (tid < 8) {                                  // thread t owns the RK4 state of row t
    float* st = g.state + rowc * 8;               // y[3], accw[3], x_stage, s_stage
    float y[3] = {0, 0, 0}, accw[3] = {0, 0, 0}, xst, sst = 0.f;
    if (g.eval < 0) {
      const float* src = g.y_in + rowc * g.stride;
      y[0] = src[0];
      if (NJ >= 2) y[1] = src[1];
      if (NJ == 3) y[2] = src[2];
      xst = y[0]; sst = y[2];
    } else {
      float h0 = g.theta[0], h1 = 0.f, h2 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { h0 += part[0][q][t]; h1 += part[1][q][t]; h2 += part[2][q][t]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) { y[c] = st[c]; accw[c] = st[3 + c]; }
      sst = st[7];
      float k[3];
      k[0] = g.y0 * h0; k[1] = -g.y0 * h1; k[2] = -g.y0 * (h1 * sst + h2);
      const int stage = g.eval & 3;
      const float wgt = (stage == 0 || stage == 3) ? 1.0f : 2.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) accw[c] = fmaf(wgt, k[c], accw[c]);
      if (stage < 3) {
        const float cs = stage == 2 ? h : 0.5f * h;
        xst = fmaf(cs, k[0], y[0]);
        sst = fmaf(cs, k[2], y[2]);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { y[c] = fmaf(h * (1.0f / 6.0f), accw[c], y[c]); accw[c] = 0.f; }
        xst = y[0]; sst = y[2];
      }
    }
    bc[t][0] = xst; bc[t][1] = sst;
    if (valid) {
      const int ne = g.eval + 1;
      if (ne < 4 * g.n_steps) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { st[c] = y[c]; st[3 + c] = accw[c]; }
        st[6] = xst; st[7] = sst;
        if (g.ckpt != nullptr) {
          float* ck = g.ckpt + ((long long)ne * 2) * g.B + row;
          ck[0] = xst;
          if (NJ == 3) ck[g.B] = sst;
        }
      } else {
        float* dst = g.y_out + row * g.stride;
        dst[0] = y[0];
        if (NJ >= 2) dst[1] = y[1];
        if (NJ == 3) dst[2] = y[2];
      }
    }
  }
  const int ne = g.eval + 1;
  if (ne >= 4 * g.n_steps) return;
  __syncthreads();
  const float xst = bc[t][0];
  const int stage = ne & 3;
  const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
  const float tp = g.y0 * (g.t0 + h * (float)(ne >> 2) + cs);
  TRACE(6)
  if (!valid && g.acts0 == nullptr) return;
  float* blk = g.acts0 != nullptr ? g.acts0 + (size_t)blockIdx.x * g.act_tile_stride + t : nullptr;
  for (int grp = gs; grp < H / 4; grp += 32) {
    const int k = 4 * grp;
    float* d0 = g.S0 + sbuf_off(grp, row, g.rows_pad);
    float a[4], wx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      wx[e] = g.theta[TL.w(0) + H + k + e];
      a[e] = tanh_f(fmaf(wx[e], xst, fmaf(tp, g.theta[TL.w(0) + k + e], g.theta[TL.b(0) + k + e])));
      if (blk != nullptr) {                       // rows beyond B repeat row B-1 (finite; their cotangents are zero)
        blk[(size_t)(k + e) * kTT] = a[e];
        blk[(size_t)(H + k + e) * kTT] = wx[e];
        blk[(size_t)(2 * H + k + e) * kTT] = 0.f;
      }
    }
    if (!valid) continue;
    *reinterpret_cast<float4*>(d0) = make_float4(a[0], a[1], a[2], a[3]);
    if (NJ >= 2) *reinterpret_cast<float4*>(d0 + jet_stride) = make_float4(wx[0], wx[1], wx[2], wx[3]);
    if (NJ == 3) *reinterpret_cast<float4*>(d0 + 2 * jet_stride) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NJ>
static int run_fwd_tc(cudaStream_t st, const float* theta, const float* WTc, const float* WTlo, const float* y_in, float* y_out, float* ckpt,
                      float* S0, float* S1, float* state, long long B, int stride, int H, int L, int n_steps, float t0, float t1,
                      float y0sign, float* actS) {
  const long long rows_pad = (B + 127) / 128 * 128;
  const ThetaLayout TL(H, L);
  cudaError_t e = cudaFuncSetAttribute(mlp_layer_gemm_tc_kernel<NJ, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(GemmSmem) + 1024);
  if (e != cudaSuccess) return (int)e;
  // activation checkpoint (NJ == 3): every evaluation keeps its own layer buffers (actS: [eval][layer]); the adjoint then
  // skips the recomputation of the hidden layers and the weight-gradient GEMM contracts these buffers in place
  const size_t sbuf = al256(3 * (size_t)rows_pad * H * sizeof(float)) / sizeof(float);
  const long long n_tiles = (B + 7) / 8;
  const size_t act_tile = (size_t)L * 3 * H * kTT;
  auto sck = [&](int ev, int l) { return actS + ((size_t)ev * L + l) * sbuf; };
  // rows beyond B of the layer buffers are read by the last tile: keep them finite
  if (actS != nullptr) {
    if (B % 128 != 0 && (e = cudaMemsetAsync(actS, 0, (size_t)4 * n_steps * L * sbuf * sizeof(float), st)) != cudaSuccess) return (int)e;
  } else {
    e = cudaMemsetAsync(S0, 0, 3 * (size_t)rows_pad * H * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(S1, 0, 3 * (size_t)rows_pad * H * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }
  GlueArgs ga{theta, nullptr, S0, state, y_in, y_out, ckpt, B, rows_pad, stride, H, L, n_steps, -1, t0, t1, y0sign, nullptr, act_tile};
  const int glue_grid = (int)n_tiles;
  const dim3 gemm_grid((unsigned)(rows_pad / 128), (unsigned)(H / 64));
  float* bufs[2] = {S0, S1};
  for (int ev = -1; ev < 4 * n_steps; ++ev) {
    // the last hidden layer of evaluation ev ended in bufs[(L-1) & 1]
    ga.eval = ev; ga.Slast = bufs[(L - 1) & 1];
    const bool more = ev + 1 < 4 * n_steps;
    if (actS != nullptr) {
      ga.Slast = ev >= 0 ? sck(ev, L - 1) : nullptr;
      ga.S0 = more ? sck(ev + 1, 0) : nullptr;
    }
    if ((e = launch_pdl(mlp_glue_tc_kernel<NJ>, dim3(glue_grid), dim3(256), 0, st, ga)) != cudaSuccess) return (int)e;
    if (!more) break;
    for (int l = 1; l < L; ++l) {
      GemmArgs gg{};
      gg.Sin = bufs[(l - 1) & 1]; gg.Sout = bufs[l & 1]; gg.WT = WTc + (size_t)(l - 1) * H * H; gg.WTlo = WTlo + (size_t)(l - 1) * H * H;
      gg.bias = theta + TL.b(l); gg.B = B; gg.rows_pad = rows_pad; gg.H = H;
      if (actS != nullptr) { gg.Sin = sck(ev + 1, l - 1); gg.Sout = sck(ev + 1, l); }
      if ((e = launch_pdl(mlp_layer_gemm_tc_kernel<NJ, 0>, gemm_grid, dim3(kGemmThreads), sizeof(GemmSmem) + 1024, st, gg)) != cudaSuccess) return (int)e;
    }
  }
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// adjoint sweep on the same GEMM kernel (discrete adjoint of RK4 on the checkpointed stage states; same algebra as
// mlp_bwd_sweep_kernel in mlp1d.cu).  Per stage, in backward order:
//   tail/first : RK4 adjoint bookkeeping of the stage just finished, then the first layer of the next stage (from ckpt)
//   MODE 0 x(L-1): recompute the hidden layers, keeping every layer's (a, p', p'') and writing the HBM blocks
//   head       : last layer + its reverse + reverse of the last hidden tanh jet -> Pbar_{L-1}
//   MODE 1 x(L-2), MODE 2: transposed GEMMs with the tanh-jet reverse / first-layer gradients in the epilogue
// Small-parameter gradients accumulate in per-CTA slots (one writer per slot and parameter, launches are ordered),
// reduced at the end in fixed order: bitwise reproducible.  bst[row][8] = xb, lb, sb, prevx, prevs, sumx, sums, yss.
// ------------------------------------------------------------------------------------------------
struct SweepArgs {
  const float* theta; const float* ckpt; const float* ybar1; float* ybar0;
  float* S0; const float* Slast; float* Pb; float* bst; const float* xpart; float* gsm;
  float* acts0;            // HBM blocks of layer 0 for the stage being prepared: [tile][..]
  float* pbar_last;        // HBM blocks of the last hidden layer's cotangents for the stage being reversed
  size_t act_tile_stride, pb_tile_stride;
  long long Bglob, r0, rows, rows_pad; int stride, H, L, n_steps, n_xpart;
  int e_done, e_next;      // tail: stage just finished (-1: initialise), stage to prepare (-1: none -> write ybar0)
  int e;                   // head: stage being reversed
  float t0, t1, y0;
  int fused_tail;          // head: do the tail's row bookkeeping of stage e_done first (activation-checkpoint sweep)
  unsigned char* live;     // head: per (stage, 8-row tile) flag for the weight-gradient GEMM: any cotangent jet p', p'' nonzero?
};

// RK4 adjoint bookkeeping of the stage just finished (threads 0..7 of a tile: one row each)
__device__ __forceinline__ void sweep_tail_rows(const SweepArgs& g, long long row, bool valid) {
  float* st = g.bst + row * 8;
  if (g.e_done < 0) {
    const float* src = g.ybar1 + (g.r0 + (valid ? row : 0)) * g.stride;
    st[0] = valid ? src[0] : 0.f; st[1] = valid ? src[1] : 0.f; st[2] = valid ? src[2] : 0.f;
    st[3] = st[4] = st[5] = st[6] = st[7] = 0.f;
  } else {
    float ysx = 0.f;
    for (int nt = 0; nt < g.n_xpart; ++nt) ysx += g.xpart[(size_t)nt * g.rows_pad + row];
    const float yss = st[7];
    float sumx = st[5] + ysx, sums = st[6] + yss;
    st[3] = ysx; st[4] = yss;
    if ((g.e_done & 3) == 0) { st[0] += sumx; st[2] += sums; sumx = 0.f; sums = 0.f; }
    st[5] = sumx; st[6] = sums;
  }
  if (g.e_next < 0 && valid && g.ybar0 != nullptr) {
    float* dst = g.ybar0 + (g.r0 + row) * g.stride;
    dst[0] = st[0]; dst[1] = st[1]; dst[2] = st[2];
  }
}

__global__ void __launch_bounds__(256) mlp_sweep_tail_kernel(const SweepArgs g) {
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, t = tid & 7, gs = tid >> 3;
  const long long row = (long long)blockIdx.x * 8 + t;      // chunk-local
  const bool valid = row < g.rows;
  const int H = g.H;
  const ThetaLayout TL(H, g.L);
  const size_t jet_stride = (size_t)g.rows_pad * H;
  const float h = (g.t1 - g.t0) / (float)g.n_steps;
  if (tid < 8) sweep_tail_rows(g, row, valid);
  if (g.e_next < 0 || g.S0 == nullptr) return;
  // first layer of stage e_next: (a, p' = w_x, p'' = 0) -> layer buffer 0 and the HBM block of layer 0
  const int stage = g.e_next & 3;
  const float cs = stage == 0 ? 0.f : (stage == 3 ? h : 0.5f * h);
  const float tp = g.y0 * (g.t0 + h * (float)(g.e_next >> 2) + cs);
  const long long rowc = valid ? row : g.rows - 1;
  const float xst = g.ckpt[((long long)g.e_next * 2) * g.Bglob + g.r0 + rowc];
  float* blk = g.acts0 + (size_t)blockIdx.x * g.act_tile_stride + t;
  for (int grp = gs; grp < H / 4; grp += 32) {
    const int k = 4 * grp;
    float a[4], wx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      wx[e] = g.theta[TL.w(0) + H + k + e];
      a[e] = tanh_f(fmaf(wx[e], xst, fmaf(tp, g.theta[TL.w(0) + k + e], g.theta[TL.b(0) + k + e])));
      blk[(size_t)(k + e) * kTT] = a[e];
      blk[(size_t)(H + k + e) * kTT] = wx[e];
      blk[(size_t)(2 * H + k + e) * kTT] = 0.f;
    }
    if (valid) {
      float* d0 = g.S0 + sbuf_off(grp, row, g.rows_pad);
      *reinterpret_cast<float4*>(d0) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(d0 + jet_stride) = make_float4(wx[0], wx[1], wx[2], wx[3]);
      *reinterpret_cast<float4*>(d0 + 2 * jet_stride) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// head: one CTA per tile of 8 trajectories (= its small-gradient slot)
__global__ void __launch_bounds__(256) mlp_sweep_head_kernel(const SweepArgs g) {
  __shared__ float part[8][8];                    // h' partials [warp][row]
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x, t = tid & 7, gs = tid >> 3, wp = tid >> 5;
  const long long row = (long long)blockIdx.x * 8 + t;
  const bool valid = row < g.rows;
  const int H = g.H, L = g.L;
  const ThetaLayout TL(H, L);
  const size_t jet_stride = (size_t)g.rows_pad * H;
  const float h = (g.t1 - g.t0) / (float)g.n_steps;
  const int stage = g.e & 3;
  if (g.fused_tail) {                              // uniform per launch
    if (tid < 8) sweep_tail_rows(g, row, valid);
    __syncthreads();
  }
  const float* st = g.bst + row * 8;
  const float ca = (stage == 0 || stage == 3) ? h * (1.0f / 6.0f) : h * (1.0f / 3.0f);
  const float cp = stage == 3 ? 0.f : (stage == 2 ? h : 0.5f * h);
  const float xb = st[0], lb = st[1], sb = st[2], prevx = st[3], prevs = st[4];
  const float kx = fmaf(ca, xb, cp * prevx), ks = fmaf(ca, sb, cp * prevs), kl = ca * lb;
  const long long rowc = valid ? row : g.rows - 1;
  const float sst = g.ckpt[((long long)g.e * 2 + 1) * g.Bglob + g.r0 + rowc];
  // dx = y0 h ; dlogp = -y0 h' ; ds = -y0 (h' s + h'')
  const float hb0 = valid ? g.y0 * kx : 0.f, hb1 = valid ? -g.y0 * (kl + sst * ks) : 0.f, hb2 = valid ? -g.y0 * ks : 0.f;
  if (g.live != nullptr) {
    // rows whose logp / score cotangents are zero (the second half of a batch: the loss only uses their positions) keep
    // p'-bar = p''-bar = 0 exactly through every layer; a tile made of such rows is contracted over jet 0 only
    const int any = __syncthreads_or((hb1 != 0.f) | (hb2 != 0.f));
    if (tid == 0) g.live[(size_t)g.e * gridDim.x + blockIdx.x] = (unsigned char)(any != 0);
  }
  float h1 = 0.f;
  float* blk = g.pbar_last != nullptr ? g.pbar_last + (size_t)blockIdx.x * g.pb_tile_stride + t : nullptr;
  float* gsl = g.gsm + (size_t)blockIdx.x * TL.n_small();
  for (int grp = gs; grp < H / 4; grp += 32) {
    const int k = 4 * grp;
    const float* s0 = g.Slast + sbuf_off(grp, row, g.rows_pad);
    const float4 a4 = *reinterpret_cast<const float4*>(s0);
    const float4 p14 = *reinterpret_cast<const float4*>(s0 + jet_stride);
    const float4 p24 = *reinterpret_cast<const float4*>(s0 + 2 * jet_stride);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w}, p1[4] = {p14.x, p14.y, p14.z, p14.w}, p2[4] = {p24.x, p24.y, p24.z, p24.w};
    float q0[4], q1[4], q2[4], gwv[4], gbv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float wl = g.theta[TL.w_last() + k + e];
      const float sg = fmaf(-av[e], av[e], 1.0f);
      const float a1 = sg * p1[e];
      const float a2 = fmaf(sg, p2[e], -2.0f * av[e] * a1 * p1[e]);
      h1 = fmaf(wl, a1, h1);
      const float ab = wl * hb0, a1b = wl * hb1, a2b = wl * hb2;
      q2[e] = a2b * sg;
      q1[e] = sg * fmaf(-4.0f * av[e] * p1[e], a2b, a1b);
      q0[e] = sg * (ab - 2.0f * av[e] * p1[e] * a1b + a2b * fmaf(-2.0f * av[e], p2[e], -2.0f * fmaf(-3.0f * av[e], av[e], 1.0f) * p1[e] * p1[e]));
      if (blk != nullptr) {
        blk[(size_t)(k + e) * kTT] = q0[e];
        blk[(size_t)(H + k + e) * kTT] = q1[e];
        blk[(size_t)(2 * H + k + e) * kTT] = q2[e];
      }
      // last-layer kernel and last hidden bias gradients: sums over the 8 rows of the tile, one writer per unit
      gwv[e] = tile_rows_sum(av[e] * hb0 + a1 * hb1 + a2 * hb2);
      gbv[e] = tile_rows_sum(q0[e]);
    }
    if (t == 0) {          // read-modify-write of the tile's slot: all loads first, then the stores (compact offsets are odd: scalar)
      float* __restrict__ gw_p = gsl + TL.c_wlast() + k;
      float* __restrict__ gb_p = gsl + TL.c_b(L - 1) + k;
      const float w0 = gw_p[0], w1 = gw_p[1], w2 = gw_p[2], w3 = gw_p[3];
      const float b0 = gb_p[0], b1 = gb_p[1], b2 = gb_p[2], b3 = gb_p[3];
      gw_p[0] = w0 + gwv[0]; gw_p[1] = w1 + gwv[1]; gw_p[2] = w2 + gwv[2]; gw_p[3] = w3 + gwv[3];
      gb_p[0] = b0 + gbv[0]; gb_p[1] = b1 + gbv[1]; gb_p[2] = b2 + gbv[2]; gb_p[3] = b3 + gbv[3];
    }
    if (valid) {
      float* d0 = g.Pb + sbuf_off(grp, row, g.rows_pad);
      *reinterpret_cast<float4*>(d0) = make_float4(q0[0], q0[1], q0[2], q0[3]);
      *reinterpret_cast<float4*>(d0 + jet_stride) = make_float4(q1[0], q1[1], q1[2], q1[3]);
      *reinterpret_cast<float4*>(d0 + 2 * jet_stride) = make_float4(q2[0], q2[1], q2[2], q2[3]);
    }
  }
  h1 = row_lanes_sum(h1);
  if ((tid & 31) < 8) part[wp][t] = h1;
  const float gl = tile_rows_sum(hb0);            // last-layer bias gradient of the tile (all lanes take part)
  __syncthreads();
  if (tid < 8) {
    float hs = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) hs += part[q][t];
    g.bst[row * 8 + 7] = valid ? -g.y0 * hs * ks : 0.f;          // direct s cotangent of this stage
    if (t == 0) gsl[0] += gl;
  }
}

int launch_mlp_bwd_sweep_tc(void* stream, const float* theta, const float* Wc, const float* Wclo, const float* WTc,
                            const float* WTlo, const float* ckpt, const float* ybar1, float* ybar0, float* const* S,
                            float* Pb0, float* Pb1, float* bst, float* xpart, float* gsmall, float* acts, float* pbar,
                            long long Bglob, long long r0, long long rows, int stride, int H, int L, int n_steps, float t0,
                            float t1, float y0sign, const float* actS, const MlpDwHook* dw_hook, unsigned char* live) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows_pad = (rows + 127) / 128 * 128;
  const long long n_tiles = (rows + 7) / 8;
  const ThetaLayout TL(H, L);
  const size_t sbytes = 3 * (size_t)rows_pad * H * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(mlp_layer_gemm_tc_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem) + 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_layer_gemm_tc_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem) + 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_layer_gemm_tc_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem) + 1024);
  if (e != cudaSuccess) return (int)e;
  // activation checkpoint of the forward (single chunk: r0 == 0, rows == Bglob): every stage's layer buffers are read in
  // place, the hidden layers are not recomputed, and the tail's bookkeeping rides in the head kernel
  const size_t sbuf = al256(sbytes) / sizeof(float);
  auto sck = [&](int ev, int l) { return const_cast<float*>(actS) + ((size_t)ev * L + l) * sbuf; };
  // ... and the cotangent jets stay in per-evaluation layer buffers too (pbar: [eval][slot l-1 = cotangent of hidden layer l]),
  // which the weight-gradient GEMM reads in place; rows beyond the batch must hold zeros (bias-gradient column sums)
  auto pck = [&](int ev, int slot) { return pbar + ((size_t)ev * (L - 1) + slot) * sbuf; };
  if (actS == nullptr)
    for (int l = 0; l < L; ++l) if ((e = cudaMemsetAsync(S[l], 0, sbytes, st)) != cudaSuccess) return (int)e;
  if (actS != nullptr && rows % 128 != 0 &&
      (e = cudaMemsetAsync(pbar, 0, (size_t)4 * n_steps * (L - 1) * sbuf * sizeof(float), st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(Pb0, 0, sbytes, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(Pb1, 0, sbytes, st)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(gsmall, 0, (size_t)256 * TL.n_small() * sizeof(float), st)) != cudaSuccess) return (int)e;
  const size_t act_tile = (size_t)L * 3 * H * kTT, pb_tile = (size_t)(L - 1) * 3 * H * kTT;
  const int n_evals = 4 * n_steps;
  const float hstep = (t1 - t0) / (float)n_steps;
  const int glue_grid = (int)n_tiles;
  const dim3 gemm_grid((unsigned)(rows_pad / 128), (unsigned)(H / 64));
  SweepArgs sa{};
  sa.theta = theta; sa.ckpt = ckpt; sa.ybar1 = ybar1; sa.ybar0 = ybar0; sa.S0 = S[0]; sa.Slast = S[L - 1]; sa.Pb = Pb0; sa.bst = bst;
  sa.xpart = xpart; sa.gsm = gsmall; sa.act_tile_stride = act_tile; sa.pb_tile_stride = pb_tile; sa.Bglob = Bglob; sa.r0 = r0;
  sa.rows = rows; sa.rows_pad = rows_pad; sa.stride = stride; sa.H = H; sa.L = L; sa.n_steps = n_steps; sa.n_xpart = H / 64;
  sa.t0 = t0; sa.t1 = t1; sa.y0 = y0sign; sa.live = live;
  int dw_next_hi = n_evals;
  for (int ev = n_evals - 1; ev >= -1; --ev) {
    // finish stage ev+1 (or initialise), prepare stage ev
    sa.e_done = ev + 1 < n_evals ? ev + 1 : -1; sa.e_next = ev;
    sa.acts0 = ev >= 0 ? acts + (size_t)ev * n_tiles * act_tile : nullptr;                       // layer 0 slot of the tile block
    sa.fused_tail = actS != nullptr && ev >= 0;
    if (actS != nullptr) { sa.S0 = nullptr; sa.Slast = ev >= 0 ? sck(ev, L - 1) : nullptr; }
    if (!sa.fused_tail && (e = launch_pdl(mlp_sweep_tail_kernel, dim3(glue_grid), dim3(256), 0, st, sa)) != cudaSuccess) return (int)e;
    if (ev < 0) break;
    for (int l = 1; l < L && actS == nullptr; ++l) {
      GemmArgs gg{};
      gg.Sin = S[l - 1]; gg.Sout = S[l]; gg.WT = WTc + (size_t)(l - 1) * H * H; gg.WTlo = WTlo + (size_t)(l - 1) * H * H;
      gg.bias = theta + TL.b(l); gg.B = rows; gg.rows_pad = rows_pad; gg.H = H;
      gg.blk_out = acts + (size_t)ev * n_tiles * act_tile + (size_t)l * 3 * H * kTT; gg.blk_tile_stride = act_tile;
      if ((e = launch_pdl(mlp_layer_gemm_tc_kernel<3, 0>, gemm_grid, dim3(kGemmThreads), sizeof(GemmSmem) + 1024, st, gg)) != cudaSuccess) return (int)e;
    }
    sa.e = ev;
    sa.pbar_last = pbar + (size_t)ev * n_tiles * pb_tile + (size_t)(L - 2) * 3 * H * kTT;
    if (actS != nullptr) { sa.pbar_last = nullptr; sa.Pb = pck(ev, L - 2); }
    if ((e = launch_pdl(mlp_sweep_head_kernel, dim3(glue_grid), dim3(256), 0, st, sa)) != cudaSuccess) return (int)e;
    const int stage = ev & 3;
    const float cs = stage == 0 ? 0.f : (stage == 3 ? hstep : 0.5f * hstep);
    float* pin = Pb0; float* pout = Pb1;
    for (int l = L - 1; l >= 1; --l) {
      GemmArgs gg{};
      if (actS != nullptr) { pin = pck(ev, l - 1); pout = l >= 2 ? pck(ev, l - 2) : nullptr; }
      gg.Sin = pin; gg.Sout = pout; gg.WT = Wc + (size_t)(l - 1) * H * H; gg.WTlo = Wclo + (size_t)(l - 1) * H * H;
      gg.B = rows; gg.rows_pad = rows_pad; gg.H = H; gg.Sprev = actS != nullptr ? sck(ev, l - 1) : S[l - 1]; gg.gsm = gsmall; gg.n_small = TL.n_small();
      if (l >= 2) {
        gg.off_b = TL.c_b(l - 1);
        if (actS == nullptr) { gg.blk_out = pbar + (size_t)ev * n_tiles * pb_tile + (size_t)(l - 2) * 3 * H * kTT; gg.blk_tile_stride = pb_tile; }
        if ((e = launch_pdl(mlp_layer_gemm_tc_kernel<3, 1>, gemm_grid, dim3(kGemmThreads), sizeof(GemmSmem) + 1024, st, gg)) != cudaSuccess) return (int)e;
        float* t = pin; pin = pout; pout = t;
      } else {
        gg.off_b = TL.c_b0(); gg.off_w0 = TL.c_w0(); gg.wx = theta + TL.w(0) + H;
        gg.xs = ckpt + ((long long)ev * 2) * Bglob + r0; gg.tprime = y0sign * (t0 + hstep * (float)(ev >> 2) + cs); gg.xpart = xpart;
        if ((e = launch_pdl(mlp_layer_gemm_tc_kernel<3, 2>, gemm_grid, dim3(kGemmThreads), sizeof(GemmSmem) + 1024, st, gg)) != cudaSuccess) return (int)e;
      }
    }
    // hand finished evaluations to the weight-gradient GEMM: in groups, and one by one at the end of the sweep (what is
    // handed over last runs after the sweep, exposed)
    if (dw_hook != nullptr && (ev % dw_hook->group == 0 || ev < dw_hook->group)) {
      const int rc = dw_hook->launch(dw_hook->ctx, ev, dw_next_hi);
      if (rc) return rc;
      dw_next_hi = ev;
    }
    // the head of the next stage writes Pb0 again; MODE 1 outputs alternate between the two buffers
  }
  return (int)cudaGetLastError();
}

__global__ void mlp_prep_tc_kernel(const float* __restrict__ theta, float* __restrict__ Wf, float* __restrict__ Wf_lo,
                                   float* __restrict__ Wb, float* __restrict__ Wb_lo, int H, int L) {
  const ThetaLayout TL(H, L);
  const long long n = (long long)(L - 1) * H * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i / ((long long)H * H)) + 1;
    const int k = (int)((i % ((long long)H * H)) / H), j = (int)(i % H);       // W_l[k][j], (in, out)
    const float v = theta[TL.w(l) + k * H + j];
    const size_t base = (size_t)(l - 1) * H * H;
    const size_t of = base + ((size_t)(k >> 2) * H + j) * 4 + (k & 3);         // forward: contraction over k, rows n = j
    const size_t ob = base + ((size_t)(j >> 2) * H + k) * 4 + (j & 3);         // adjoint: contraction over j, rows n = k
    Wf[of] = v; Wf_lo[of] = tc::lo_tf32(v);
    if (Wb != nullptr) { Wb[ob] = v; Wb_lo[ob] = tc::lo_tf32(v); }
  }
}

int launch_mlp_prep_tc(void* stream, const float* theta, float* Wf, float* Wf_lo, float* Wb, float* Wb_lo, int H, int L) {
  mlp_prep_tc_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(theta, Wf, Wf_lo, Wb, Wb_lo, H, L);
  return (int)cudaGetLastError();
}

int launch_mlp_fwd_tc(void* stream, const float* theta, const float* WTc, const float* WTlo, const float* y_in, float* y_out, float* ckpt,
                      float* S0, float* S1, float* state, long long B, int jets, int stride, int H, int L, int n_steps,
                      float t0, float t1, float y0sign, float* actS) {
  cudaStream_t st = (cudaStream_t)stream;
  if (jets == 3) return run_fwd_tc<3>(st, theta, WTc, WTlo, y_in, y_out, ckpt, S0, S1, state, B, stride, H, L, n_steps, t0, t1, y0sign, actS);
  if (actS != nullptr) return OFDFT_EINVAL;                    // the activation checkpoint belongs to the full-jet training path
  if (jets == 2) return run_fwd_tc<2>(st, theta, WTc, WTlo, y_in, y_out, ckpt, S0, S1, state, B, stride, H, L, n_steps, t0, t1, y0sign, nullptr);
  return run_fwd_tc<1>(st, theta, WTc, WTlo, y_in, y_out, ckpt, S0, S1, state, B, stride, H, L, n_steps, t0, t1, y0sign, nullptr);
}

}  // namespace mlp
}  // namespace ofdft

#ifdef OFDFT_GEMM_TRACE
extern "C" int ofdft_debug_trace_dump(unsigned long long* host, int n_words) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host, ofdft::mlp::g_trace, (size_t)n_words * 8);
}
#endif
