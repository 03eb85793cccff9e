// tc_test.cu -- diagnostic: D[M x 64] = A[M x 64] . W[64 x 64] on tcgen05 (kind::tf32, 1 or 3 passes) with operands
// staged by threads in the no-swizzle canonical layout; dumps all 128 TMEM lanes so the M=64 lane mapping can be read off.
#include "common.cuh"
#include "tc.cuh"

namespace ofdft {

__global__ void __launch_bounds__(128, 1) tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                             float* __restrict__ D, int M, int passes, int mn_major, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* Ahi = reinterpret_cast<float*>(smem_raw);         // [16 kc][128 m][4]
  float* Alo = Ahi + 64 * 128;
  float* Whi = Alo + 64 * 128;                             // [16 kc][64 n][4]
  float* Wlo = Whi + 64 * 64;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(Ahi + 32 * (128 * 4 + 4) + 2 * 32 * (64 * 4 + 4) + 256);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int kc = 0; kc < 16; ++kc) {
    float4 h = make_float4(0, 0, 0, 0), l = make_float4(0, 0, 0, 0);
    if (tid < M) {
      const float4 v = *reinterpret_cast<const float4*>(A + tid * 64 + 4 * kc);
      tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
      if (mode == 3) h = v;        // raw fp32 bits: does the tensor core truncate or round its tf32 inputs?
    }
    *reinterpret_cast<float4*>(Ahi + (kc * 128 + tid) * 4) = h;
    *reinterpret_cast<float4*>(Alo + (kc * 128 + tid) * 4) = l;
    if (tid < 64 && mn_major) {
      const float4 v = *reinterpret_cast<const float4*>(W + tid * 64 + 4 * kc);
      float4 wh, wl;
      tc::split_tf32(v.x, wh.x, wl.x); tc::split_tf32(v.y, wh.y, wl.y); tc::split_tf32(v.z, wh.z, wl.z); tc::split_tf32(v.w, wh.w, wl.w);
      *reinterpret_cast<float4*>(Whi + (kc * 64 + tid) * 4) = wh;
      *reinterpret_cast<float4*>(Wlo + (kc * 64 + tid) * 4) = wl;
    }
    if (tid < 64 && !mn_major) {
      float4 wh, wl;
      tc::split_tf32(W[(4 * kc + 0) * 64 + tid], wh.x, wl.x); tc::split_tf32(W[(4 * kc + 1) * 64 + tid], wh.y, wl.y);
      tc::split_tf32(W[(4 * kc + 2) * 64 + tid], wh.z, wl.z); tc::split_tf32(W[(4 * kc + 3) * 64 + tid], wh.w, wl.w);
      if (mode == 3) wh = make_float4(W[(4 * kc + 0) * 64 + tid], W[(4 * kc + 1) * 64 + tid], W[(4 * kc + 2) * 64 + tid], W[(4 * kc + 3) * 64 + tid]);
      *reinterpret_cast<float4*>(Whi + (kc * 64 + tid) * 4) = wh;
      *reinterpret_cast<float4*>(Wlo + (kc * 64 + tid) * 4) = wl;
    }
  }
  if (tid == 0) { tc::mbar_init(mbar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 256);
  if (mode == 2) {
    // K = trajectory operands, padded K-major: [traj chunk (32)][row][4 traj], LBO = rows*16 + 16 (bank-conflict-free scalar stores)
    // A: rows 0..63 = hi of unit k, rows 64..127 = lo of unit k ; P[traj][j] = 0.5 * A[traj][(j+17)%64] -> hi in Whi-area, lo in Wlo-area
    __syncthreads();                                   // the generic staging above is dead in this mode
    float* At = Ahi;                                   // 32 * (128*4 + 4) floats
    float* Ph = Ahi + 32 * (128 * 4 + 4); float* Pl = Ph + 32 * (64 * 4 + 4);
    for (int k = 0; k < 64; ++k) {
      float hi, lo;
      tc::split_tf32(A[tid * 64 + k], hi, lo);
      At[(tid >> 2) * (128 * 4 + 4) + k * 4 + (tid & 3)] = hi;
      At[(tid >> 2) * (128 * 4 + 4) + (64 + k) * 4 + (tid & 3)] = lo;
      tc::split_tf32(0.5f * A[tid * 64 + ((k + 17) & 63)], hi, lo);
      Ph[(tid >> 2) * (64 * 4 + 4) + k * 4 + (tid & 3)] = hi;
      Pl[(tid >> 2) * (64 * 4 + 4) + k * 4 + (tid & 3)] = lo;
    }
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (mode == 1) {
    // TS mode: A hi -> TMEM columns [64,128), A lo -> [128,192); thread = row
    for (int c0 = 0; c0 < 64; c0 += 16) {
      float h[16], l[16];
      for (int i = 0; i < 16; ++i) tc::split_tf32(A[tid * 64 + c0 + i], h[i], l[i]);
      tc::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + 64 + c0, h);
      tc::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + 128 + c0, l);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (tid == 0) {
      const uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
      bool acc = false;
      for (int p = 0; p < 3; ++p) {
        const uint32_t ac = p == 1 ? 128 : 64;
        const float* Wp = (p == 2) ? Wlo : Whi;
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bd = tc::make_desc(tc::smem_u32(Wp) + ks * 2 * 1024, 1024, 128);
          tc::mma_tf32_ts(tmem, tmem + ac + ks * 8, bd, idesc, acc);
          acc = true;
        }
      }
      tc::commit(mbar);
    }
  }
  if (mode == 2 && tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, 64, 0, 0);
    const uint32_t lboA = 128 * 16 + 16, lboP = 64 * 16 + 16;
    bool acc = false;
    for (int p = 0; p < 2; ++p) {
      const float* Pp = Ahi + 32 * (128 * 4 + 4) + (p == 0 ? 0 : 32 * (64 * 4 + 4));
      for (int ks = 0; ks < 16; ++ks) {               // K = 128 trajectories
        const uint64_t ad = tc::make_desc(tc::smem_u32(Ahi) + ks * 2 * lboA, lboA, 128);
        const uint64_t bd = tc::make_desc(tc::smem_u32(Pp) + ks * 2 * lboP, lboP, 128);
        tc::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = true;
      }
    }
    tc::commit(mbar);
  }
  if (tid == 0 && mn_major && mode == 0) {
    // D[[hi;lo] k (128)][j (64)] = sum_traj A[traj][k] * Wt[traj][j]: both operands MN-major, K = trajectory index.
    // Here "W" is read as a (64 traj x 64 j) matrix and A's first 64 rows are the trajectories (zero padded to 128).
    const uint32_t idesc = tc::make_idesc_tf32(128, 64, 1, 1);
    bool acc = false;
    for (int p = 0; p < 2; ++p) {
      const float* Bp = p == 0 ? Whi : Wlo;        // [jc(16)][traj(64)][4]: SBO = 64*16 = 1024, LBO = 128
      for (int ks = 0; ks < 8; ++ks) {             // K = 64 trajectories, 8 per MMA
        const uint64_t ad = tc::make_desc(tc::smem_u32(Ahi) + ks * 128, 128, 2048);   // [hi|lo] contiguous: 32 M-chunks
        const uint64_t bd = tc::make_desc(tc::smem_u32(Bp) + ks * 128, 128, 1024);
        tc::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = true;
      }
    }
    tc::commit(mbar);
  }
  if (tid == 0 && !mn_major && (mode == 0 || mode == 3)) {
    const uint32_t idesc = tc::make_idesc_tf32(M, 64, 0, 0);
    bool acc = false;
    for (int p = 0; p < passes; ++p) {
      const float* Ap = (p == 1) ? Alo : Ahi;
      const float* Wp = (p == 2) ? Wlo : Whi;
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = tc::make_desc(tc::smem_u32(Ap) + ks * 2 * 2048, 2048, 128);
        const uint64_t bd = tc::make_desc(tc::smem_u32(Wp) + ks * 2 * 1024, 1024, 128);
        tc::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = true;
      }
    }
    tc::commit(mbar);
  }
  tc::mbar_wait(mbar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(32 * warp + lane) * 64 + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace ofdft

extern "C" int ofdft_tc_gemm_test(void* stream, const float* A, const float* W, float* D, int32_t M, int32_t passes) {
  using namespace ofdft;
  // passes == 0: MN-major probe (unsupported layout for tf32, kept as a negative test);
  // passes == -1: A operand from tensor memory (tcgen05.st + TS-mode MMA);
  // passes == -3: one pass on raw (unrounded) fp32 bits -- shows how the tensor core narrows its inputs;
  // passes == -2: K = trajectory operands in the padded K-major layout, stacked [hi;lo] rows (weight-gradient GEMM)
  if (!A || !W || !D || (M != 64 && M != 128) || passes < -3 || passes > 3) return OFDFT_EINVAL;
  const size_t smem = sizeof(float) * (32 * (128 * 4 + 4) + 2 * 32 * (64 * 4 + 4) + 256) + 64 + 4096;
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_gemm_test_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, D, M, passes == -3 ? 1 : (passes < 0 ? 0 : passes), passes == 0 ? 1 : 0, passes < 0 ? -passes : 0);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// diagnostic: clocks per tcgen05.mma (kind::tf32, M = 128, K = 8) issued back to back, by N and by where the A operand
// lives (shared memory descriptor / tensor memory).  out[0] = clocks for `reps` MMAs (issue + completion).
// ------------------------------------------------------------------------------------------------
namespace ofdft {
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(long long* out, int N, int ts, int reps, int distinct_b, int nacc) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* A = reinterpret_cast<float*>(smem_raw);               // 128 x 8 tf32, K-major: 2 chunks x 128 rows x 16 B = 4 KB
  float* B = A + 1024;                                         // up to 8 tiles of 256 x 8: 8 KB each
  uint64_t* mbar = reinterpret_cast<uint64_t*>(B + 8 * 2048);
  uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 1024 + 8 * 2048; i += 128) A[i] = 0.001f * (float)(i & 63);
  if (tid == 0) { tc::mbar_init(mbar, 1); tc::fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  {
    float v[16];
    for (int i = 0; i < 16; ++i) v[i] = 0.5f;
    tc::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + 448, v);     // A operand columns 448..463
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 0 && tc::elect_one()) {      // elect.sync: the compiler keeps descriptors in uniform registers, as in the kernels
    const uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
    const uint64_t ad = tc::make_desc(tc::smem_u32(A), 128 * 16, 128);
    const uint64_t bd0 = tc::make_desc(tc::smem_u32(B), N * 16, 128);
    const uint64_t bd1 = tc::make_desc(tc::smem_u32(B + (distinct_b ? 2048 : 0)), N * 16, 128);
    // nacc accumulators of N columns used round-robin: consecutive MMAs into the SAME accumulator depend on each other
    const uint32_t stride = nacc > 1 ? (uint32_t)N : 0u;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 12) {            // 12 back-to-back issues per loop trip, descriptors precomputed
#pragma unroll
      for (int q = 0; q < 12; ++q) {
        const uint32_t d = tmem + stride * (uint32_t)(nacc == 2 ? (q & 1) : (nacc == 3 ? (q % 3) : (nacc == 4 ? (q & 3) : 0)));
        if (ts) tc::mma_tf32_ts(d, tmem + 448 + 8 * (q & 1), (q & 1) ? bd1 : bd0, idesc, true);
        else tc::mma_tf32(d, ad, (q & 1) ? bd1 : bd0, idesc, true);
      }
    }
    tc::commit(mbar);
    tc::mbar_wait(mbar, 0);
    out[0] = clock64() - t0;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace ofdft

extern "C" int ofdft_tc_mma_rate(void* stream, long long* out, int32_t N, int32_t ts, int32_t reps, int32_t distinct_b,
                                 int32_t nacc) {
  using namespace ofdft;
  if (!out || N < 8 || N > 256 || N % 8 || reps < 1 || nacc < 1 || nacc > 4 || nacc * N > 448) return OFDFT_EINVAL;
  const size_t smem = sizeof(float) * (1024 + 8 * 2048) + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  tc_mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out, N, ts, reps, distinct_b, nacc);
  return (int)cudaGetLastError();
}
