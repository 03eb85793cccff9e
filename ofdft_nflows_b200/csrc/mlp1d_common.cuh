// mlp1d_common.cuh -- shared declarations of the 1-D tanh-MLP flow kernels (mlp1d.cu, mlp1d_tc.cu)
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/ofdft_b200.h"

namespace ofdft {
namespace mlp {

constexpr int kTT = 8;        // trajectories per tile
constexpr int kKC = 16;       // weight rows per streamed chunk
constexpr int kMaxL = 4;

struct ThetaLayout {
  int H, L, n;
  __host__ __device__ ThetaLayout(int H_, int L_) : H(H_), L(L_) { n = 1 + H + 3 * H + (L - 1) * (H + H * H); }
  __host__ __device__ int b_last() const { return 0; }
  __host__ __device__ int w_last() const { return 1; }
  __host__ __device__ int b(int i) const { return i == 0 ? 1 + H : 1 + 4 * H + (i - 1) * (H + H * H); }
  __host__ __device__ int w(int i) const { return b(i) + H; }     // layer 0: (2,H) rows [t, x]; else (H,H)
  // compact layout of everything except the hidden kernels (per-CTA partial buffers of the backward sweep)
  __host__ __device__ int n_small() const { return 1 + 4 * H + (L - 1) * H; }
  __host__ __device__ int c_wlast() const { return 1; }
  __host__ __device__ int c_b0() const { return 1 + H; }
  __host__ __device__ int c_w0() const { return 1 + 2 * H; }
  __host__ __device__ int c_b(int l) const { return 1 + 4 * H + (l - 1) * H; }
  // compact index -> theta offset
  __host__ __device__ int small_to_theta(int c) const {
    if (c < 1 + H) return c;                              // b_last, w_last
    if (c < 1 + 4 * H) return c;                          // b0, w0 share the same offsets (1+H .. 1+4H)
    const int l = (c - (1 + 4 * H)) / H + 1, i = (c - (1 + 4 * H)) % H;
    return b(l) + i;
  }
};

constexpr long long kChunkRows = 2048;     // backward row chunk (bounds the HBM activation scratch)
constexpr int kDwSplits = 16;
inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// Activation checkpoint of the tensor-core path (OFDFT_PATH_ACT_CKPT; one backward row chunk, i.e. B <= kChunkRows, H a
// multiple of 128): appended to the stage checkpoint, per (evaluation, layer) buffers [jet][H/4][rows padded to 128][4] -- the
// layout the layer GEMMs read and write, and the one the weight-gradient GEMM contracts in place (no HBM blocks).
struct ActLayout {
  size_t s, total; bool ok;
  ActLayout(long long B, int H, int L, int n_steps) {
    ok = B > 0 && B <= kChunkRows;
    const size_t rows_pad = (size_t)((B + 127) / 128 * 128);
    s = al256((size_t)n_steps * 4 * 2 * (size_t)B * 4);
    total = s + (size_t)n_steps * 4 * L * al256(3 * rows_pad * H * 4);
  }
};

struct ScratchLayout {
  size_t wc, wtc, wtlo, wclo, wbi, fwd_s, fwd_state, fwd_end, bwd_s, bwd_pb, bwd_bst, bwd_xpart, bwd_live, gsmall, part, acts, pbar, total;
  ScratchLayout(long long B, int H, int L, int n_steps) {
    const long long rows = B < kChunkRows ? B : kChunkRows;
    const long long n_tiles = (rows + kTT - 1) / kTT;
    const long long n_blocks = (long long)n_steps * 4 * n_tiles;
    size_t o = 0;
    wc = o; o += al256((size_t)(L - 1) * H * H * 4);
    wtc = o; o += al256((size_t)(L - 1) * H * H * 4);
    wtlo = o; o += al256((size_t)(L - 1) * H * H * 4);    // tf32 remainders of the transposed kernels
    // tensor-core forward: two ping-pong layer buffers [jet: a, p', p''][H/4][B rows padded to 128][4 units] and the RK4 row state
    const size_t rows_pad = (size_t)((B + 127) / 128 * 128);
    fwd_s = o; o += 2 * al256(3 * rows_pad * H * 4);
    fwd_state = o; o += al256(rows_pad * 8 * 4);
    fwd_end = o;
    // tensor-core adjoint sweep (per 2048-row chunk): L layer buffers, two cotangent buffers, row state, x partials
    const size_t rows_pad_c = (size_t)((rows + 127) / 128 * 128);
    wclo = o; o += al256((size_t)(L - 1) * H * H * 4);
    wbi = o; o += al256((size_t)(L - 1) * H * H * 4);
    bwd_s = o; o += (size_t)L * al256(3 * rows_pad_c * H * 4);
    bwd_pb = o; o += 2 * al256(3 * rows_pad_c * H * 4);
    bwd_bst = o; o += al256(rows_pad_c * 8 * 4);
    bwd_xpart = o; o += al256((size_t)(H / 64 > 0 ? H / 64 : 1) * rows_pad_c * 4);
    bwd_live = o; o += al256((size_t)n_blocks);           // one byte per (evaluation, 8-row tile): are the cotangent jets p', p'' live?
    gsmall = o; o += al256((size_t)256 * ThetaLayout(H, L).n_small() * 4);
    part = o; o += al256((size_t)kDwSplits * H * H * 4);
    acts = o; o += al256((size_t)n_blocks * L * 3 * H * kTT * 4);
    {   // HBM blocks, or (activation-checkpoint sweep) per-evaluation cotangent layer buffers padded to 128 rows
      const size_t blocks = al256((size_t)n_blocks * (L - 1) * 3 * H * kTT * 4);
      const size_t bufs = (size_t)n_steps * 4 * (L - 1) * al256(3 * rows_pad_c * H * 4);
      pbar = o; o += blocks > bufs ? blocks : bufs;
    }
    total = o;
  }
};


// tensor-core forward integrator (mlp1d_tc.cu): per-layer tcgen05 GEMM launches; returns a cudaError_t / OFDFT_E* code
// interleaved [k/4][n][4] copies of the hidden kernels for the TMA-fed GEMM: forward operand (n = output unit), adjoint
// operand (n = input unit), each with its tf32 remainder
int launch_mlp_prep_tc(void* stream, const float* theta, float* Wf, float* Wf_lo, float* Wb, float* Wb_lo, int H, int L);
int launch_mlp_fwd_tc(void* stream, const float* theta, const float* WTc, const float* WTlo, const float* y_in, float* y_out, float* ckpt,
                      float* S0, float* S1, float* state, long long B, int jets, int stride, int H, int L, int n_steps,
                      float t0, float t1, float y0sign, float* actS);

// called by the sweep after the launches of evaluation `ev_lo` (evaluations run downwards) whenever ev_lo is a multiple of
// `group`: the cotangents of evaluations [ev_lo, ev_hi) are complete on the sweep's stream
struct MlpDwHook { int group; void* ctx; int (*launch)(void* ctx, int ev_lo, int ev_hi); };

// tensor-core adjoint sweep of one row chunk (mlp1d_tc.cu): fills the HBM blocks the weight-gradient GEMM contracts and
// the per-CTA small-gradient partials; returns a cudaError_t / OFDFT_E* code
int launch_mlp_bwd_sweep_tc(void* stream, const float* theta, const float* Wc, const float* Wclo, const float* WTc,
                            const float* WTlo, const float* ckpt, const float* ybar1, float* ybar0, float* const* S,
                            float* Pb0, float* Pb1, float* bst, float* xpart, float* gsmall, float* acts, float* pbar,
                            long long Bglob, long long r0, long long rows, int stride, int H, int L, int n_steps, float t0,
                            float t1, float y0sign, const float* actS, const MlpDwHook* dw_hook, unsigned char* live);

}  // namespace mlp
}  // namespace ofdft
