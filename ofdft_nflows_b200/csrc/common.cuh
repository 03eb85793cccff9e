// common.cuh -- shared device helpers for the ofdft_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ofdft_b200.h"

namespace ofdft {

constexpr int kH = 64;          // hidden width of the GNN radial MLP (features = (64, 64))
constexpr int kTP = 128;        // trajectories per CTA tile
constexpr int kThreads = 256;   // threads per CTA: warp = (traj-half, unit-group), lane = trajectory pair
constexpr int kMaxNa = 128;     // nuclei per launch: C27H46O (utils.py:327) has 74
constexpr int kMaxTypes = 8;
constexpr int kMaxClasses = 8;  // distinct one-hot rows: the first-layer offset b1 + onehot_i . W1[2:,:] is kept on chip per class
constexpr int kMaxLayers = 3;   // hidden layers: --nn 1|2|3 selects (64,), (64,64), (64,64,64) (OFDFT_NF.py:320-326)

// theta offsets of the GNN (H = 64, I = 2 + n_types); see include/ofdft_b200.h
struct GnnThetaLayout {
  int b3, w3, b1, W1, b2, W2, n;
  __host__ __device__ explicit GnnThetaLayout(int n_types) {
    const int I = 2 + n_types;
    b3 = 0; w3 = 1; b1 = 1 + kH; W1 = b1 + kH; b2 = W1 + I * kH; W2 = b2 + kH; n = W2 + kH * kH;
  }
};

#ifndef OFDFT_HOST_SIM   // (tests/hostsim compiles gnn_general.cu for the CPU: no PTX, no warp intrinsics there)
// tanh with absolute error ~1e-7: 1 - 2/(2^(2x log2 e) + 1) on the SFU (MUFU.EX2 + MUFU.RCP).
// tanh.approx.f32 (one MUFU, 2^-11 relative) is too coarse for the 1e-5 energy parity bar.
__device__ __forceinline__ float tanh_f(float x) {
#ifdef OFDFT_TANH_PRECISE
  return tanhf(x);
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Butterfly reduction of a 32-vector held by every lane: on return lane l holds sum over lanes of v[l]
// (31 shuffles instead of 160).
__device__ __forceinline__ float warp_vec32_sum(float (&v)[32], int lane) {
#ifdef OFDFT_ABL_NOBFLY      // timing ablation only (results invalid)
  return v[lane & 31 ? 1 : 0];
#endif
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    float keep = up ? v[i + 16] : v[i];
    float send = up ? v[i] : v[i + 16];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    float keep = up ? v[i + 8] : v[i];
    float send = up ? v[i] : v[i + 8];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    float keep = up ? v[i + 4] : v[i];
    float send = up ? v[i] : v[i + 4];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    float keep = up ? v[i + 2] : v[i];
    float send = up ? v[i] : v[i + 2];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    float keep = up ? v[1] : v[0];
    float send = up ? v[0] : v[1];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

#endif  // OFDFT_HOST_SIM

// activation tiles in shared memory: [jet][unit 0..63][trajectory 0..127] with the 16-byte trajectory
// chunk XOR-swizzled by the unit index so that both access patterns are conflict-free:
//   GEMM phases : fixed unit, lanes read consecutive float2          (any in-row permutation is fine)
//   dW2 phase   : fixed chunk, lanes read 4/8 units 4 rows apart     (chunk ^ ((unit>>2)&7) spreads banks)
__device__ __forceinline__ int act_off(int unit, int traj) {
  return unit * kTP + ((((traj >> 2) ^ ((unit >> 2) & 7)) << 2) | (traj & 3));
}
constexpr int kJetStride = kH * kTP;

}  // namespace ofdft
