// gnn_common.cuh -- argument blocks and helpers shared by the FFMA (gnn.cu) and tensor-core (gnn_tc.cu) GNN kernels.
#pragma once
#include "common.cuh"

namespace ofdft {

struct GnnFwdArgs {
  const float* theta; const float* nuclei; const float* onehot;
  const float* y_in; float* y_out; float* ckpt; int* counter;
  long long B, n_a; int jets_a, jets_b; int in_stride, out_stride;
  int Na, n_types, n_steps; float t0, t1, y0; int rhs_only; float t_rhs;
  float* act; long long act_rows_a, act_rows_b;   // optional layer-2 activation checkpoint (rows padded to kTP)
};

struct GnnBwdArgs {
  const float* theta; const float* nuclei; const float* onehot;
  const float* ckpt; const float* ybar1; float* ybar0; float* gpart; int* counter;
  long long B, n_a; int jets_a, jets_b; int bar_stride;
  int Na, n_types, n_steps; float t0, t1, y0;
  const float* act; long long act_rows_a, act_rows_b;
  float* w2split;   // kDeepWSplitBytes of caller scratch for prepared weight operands: W2 hi | lo in the forward GEMM orientation
                    // (tensor-core recompute mode of gnn_tc.cu: the first 2 x 4096 floats) or [layer][fwd hi | fwd lo | transposed hi |
                    // transposed lo] (gnn_deep_tc.cu)
};

// theta offsets for L hidden layers (key-sorted ravel_pytree order; L = 2 reproduces GnnThetaLayout)
struct GnnThetaLayoutL {
  int b3, w3, b[kMaxLayers], W[kMaxLayers], n;
  __host__ __device__ GnnThetaLayoutL(int n_types, int L) {
    const int I = 2 + n_types;
    b3 = 0; w3 = 1;
    b[0] = 1 + kH; W[0] = b[0] + kH;
    int end = W[0] + I * kH;
    for (int l = 1; l < kMaxLayers; ++l) {
      b[l] = end; W[l] = end + kH;
      if (l < L) end = W[l] + kH * kH;
    }
    n = end;
  }
};

// Layer-2 activation checkpoint (optional; trades 180 GB of HBM for the recompute GEMM of the backward pass):
// for every RHS evaluation e = (step*4+stage)*Na + nucleus and both row segments, the jets (a2, p2', p2'') as
// [jet][group of 4 units][row][4 units]: a thread holding consecutive units of one trajectory moves them as float4, and
// a warp (32 consecutive rows) touches 512 contiguous bytes per instruction.
struct ActLayout {
  long long rows_a, rows_b, block; int ja, jb;
  __host__ __device__ ActLayout(long long ra, long long rb, int ja_, int jb_) : rows_a(ra), rows_b(rb), ja(ja_), jb(jb_) {
    block = (long long)kH * (ja * ra + jb * rb);
  }
  __host__ __device__ long long seg_base(long long e, bool seg_b) const {
    return e * block + (seg_b ? (long long)kH * ja * rows_a : 0);
  }
};

// Per-nucleus tables on chip.  Nuclei with identical one-hot rows share the first-layer offset cb = b1 + onehot . W1[2:,:]
// (a "class": one per species), so 128 nuclei cost 2.5 KB of shared memory instead of 32 KB; the adjoint accumulates the
// first-layer pre-activation cotangents per class as well.  More than kMaxClasses distinct rows trap.
struct alignas(16) NucTables {
  float nuc[kMaxNa * 4];
  float cb[kMaxClasses * kH];
  int cls[kMaxNa];        // nucleus -> class
  int rep[kMaxClasses];   // class -> first nucleus carrying that one-hot row
  int n_cls;
};

template <int THREADS>
__device__ __forceinline__ void load_nuc_tables(NucTables& T, const float* __restrict__ theta, const GnnThetaLayout& L,
                                                const float* __restrict__ nuclei, const float* __restrict__ onehot, int Na,
                                                int n_types) {
  const int tid = threadIdx.x;
  for (int n = tid; n < Na; n += THREADS) {
    T.nuc[n * 4 + 0] = nuclei[n * 3 + 0]; T.nuc[n * 4 + 1] = nuclei[n * 3 + 1]; T.nuc[n * 4 + 2] = nuclei[n * 3 + 2];
    T.nuc[n * 4 + 3] = 0.f;
    int first = n;
    for (int m = 0; m < n && first == n; ++m) {
      bool same = true;
      for (int t = 0; t < n_types; ++t) same = same && (onehot[m * n_types + t] == onehot[n * n_types + t]);
      if (same) first = m;
    }
    T.cls[n] = first;       // provisional: index of the first nucleus with the same row
  }
  __syncthreads();
  if (tid == 0) {
    int c = 0;
    for (int n = 0; n < Na; ++n) {
      if (T.cls[n] == n) {
        if (c == kMaxClasses) __trap();
        T.rep[c] = n; T.cls[n] = c++;
      } else {
        T.cls[n] = T.cls[T.cls[n]];
      }
    }
    T.n_cls = c;
  }
  __syncthreads();
  for (int i = tid; i < T.n_cls * kH; i += THREADS) {
    const int n = T.rep[i / kH], k = i % kH;
    float c = theta[L.b1 + k];
    for (int t = 0; t < n_types; ++t) c = fmaf(onehot[n * n_types + t], theta[L.W1 + (2 + t) * kH + k], c);
    T.cb[i] = c;
  }
}

// offset of (unit group g = unit/4, row) inside one jet block of `rows` trajectories
__host__ __device__ __forceinline__ long long act_group_off(int g, long long row, long long rows) { return ((long long)g * rows + row) * 4; }

__device__ __forceinline__ bool next_tile(int* counter, int* slot, long long n_tiles, long long& tile) {
  __syncthreads();
  if (threadIdx.x == 0) *slot = atomicAdd(counter, 1);
  __syncthreads();
  tile = *slot;
  return tile < n_tiles;
}


constexpr size_t kCounterBytes = 256;
constexpr size_t kW2SplitBytes = 2 * kH * kH * sizeof(float);
inline long long pad_rows(long long n) { return (n + kTP - 1) / kTP * kTP; }

// tensor-core forward launcher (gnn_tc.cu); returns a cudaError_t / OFDFT_E* code
int launch_gnn_fwd_tc(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes);
// tensor-core backward (a.act == nullptr: the layer-2 jets are recomputed on the tensor cores); the caller has reset the
// tile counter and reduces gpart afterwards
int launch_gnn_bwd_tc(void* stream, GnnBwdArgs& a, int grid);

// tensor-core kernels for two or three hidden layers, no activation checkpoint (gnn_deep_tc.cu)
constexpr size_t kDeepWSplitBytes = (size_t)(kMaxLayers - 1) * 4 * kH * kH * sizeof(float);
int launch_gnn_deep_fwd_tc(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes, int n_layers);
int launch_gnn_deep_bwd_tc(void* stream, GnnBwdArgs& a, int grid, int n_layers);

// general-depth kernels (gnn_general.cu): 1 or 3 hidden layers
int gnn_general_theta_size(int n_types, int n_layers);
int launch_gnn_general_fwd(void* stream, GnnFwdArgs& a, void* scratch, size_t scratch_bytes, int n_layers);
int launch_gnn_general_bwd(void* stream, GnnBwdArgs& a, int n_layers);

}  // namespace ofdft
