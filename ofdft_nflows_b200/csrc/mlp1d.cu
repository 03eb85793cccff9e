// mlp1d.cu -- 1-D tanh-MLP flow (Gen_CNFSimpleMLP, ofdft_normflows/cn_flows.py:117-182).
// Placeholder entry points: implemented after the GNN path.
#include "common.cuh"

extern "C" size_t ofdft_cnf_mlp1d_ckpt_bytes(int64_t B, int32_t n_steps) {
  return (size_t)n_steps * 4 * 2 * (size_t)B * sizeof(float);
}
extern "C" size_t ofdft_cnf_mlp1d_scratch_bytes(int64_t, int32_t, int32_t) { return 256; }
extern "C" int ofdft_cnf_mlp1d_fwd(void*, const float*, const float*, float*, float*, void*, size_t, int64_t, int32_t,
                                   int32_t, int32_t, int32_t, int32_t, float, float, float) {
  return OFDFT_EUNSUPPORTED;
}
extern "C" int ofdft_cnf_mlp1d_bwd(void*, const float*, const float*, const float*, float*, float*, void*, size_t,
                                   int64_t, int32_t, int32_t, int32_t, int32_t, int32_t, float, float, float) {
  return OFDFT_EUNSUPPORTED;
}
