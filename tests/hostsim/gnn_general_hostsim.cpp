// CPU execution of ofdft_nflows_b200/csrc/gnn_general.cu (test infrastructure only).
// The kernel source is compiled UNCHANGED as C++: the threads of a block become OS threads, __syncthreads a pthread barrier,
// shared memory a static buffer, blocks run one after the other (the kernels are persistent: one block walks every tile).
// This checks the algebra and the indexing of the general-depth GNN kernels against the float64 oracle without a GPU;
// tests/test_hostsim.py builds it with g++ and drives it through ctypes.
#define OFDFT_HOST_SIM 1
#include <cuda_runtime.h>
#include <pthread.h>
#define __launch_bounds__(...)

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static thread_local uint3 tl_threadIdx;
#define threadIdx tl_threadIdx
static pthread_barrier_t g_bar;
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline void __trap() { abort(); }
namespace ofdft { alignas(16) unsigned char smem_raw[1 << 18]; }

#include "../../ofdft_nflows_b200/csrc/gnn_general.cu"

using namespace ofdft;

template <class K, class A>
static void run_block(K kern, const A& a) {
  pthread_barrier_init(&g_bar, nullptr, kGT);
  std::vector<std::thread> th;
  for (int t = 0; t < kGT; ++t)
    th.emplace_back([=]() { tl_threadIdx.x = t; tl_threadIdx.y = 0; tl_threadIdx.z = 0; kern(a); });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&g_bar);
}

extern "C" int hostsim_theta_size(int n_types, int n_layers) { return gnn_general_theta_size(n_types, n_layers); }

extern "C" int hostsim_gnn_general_fwd(const float* theta, const float* nuclei, const float* onehot, const float* y_in, float* y_out,
                                       float* ckpt, long long B, long long n_a, int jets_a, int jets_b, int stride, int Na,
                                       int n_types, int n_layers, int n_steps, float t0, float t1, float y0, int rhs_only, float t_rhs) {
  int counter = 0;
  GnnFwdArgs a{theta, nuclei, onehot, y_in, y_out, ckpt, &counter, B, n_a, jets_a, jets_b, stride, stride,
               Na, n_types, n_steps, t0, t1, y0, rhs_only, t_rhs, nullptr, 0, 0};
  if (n_layers == 1) run_block(gnn_general_fwd_kernel<1>, a);
  else if (n_layers == 3) run_block(gnn_general_fwd_kernel<3>, a);
  else return -2;
  return 0;
}

// gpart: (n_tiles, n_theta) per-tile partial gradients (the caller sums them)
extern "C" int hostsim_gnn_general_bwd(const float* theta, const float* nuclei, const float* onehot, const float* ckpt,
                                       const float* ybar1, float* ybar0, float* gpart, long long B, long long n_a, int jets_a,
                                       int jets_b, int stride, int Na, int n_types, int n_layers, int n_steps, float t0, float t1,
                                       float y0) {
  int counter = 0;
  GnnBwdArgs a{theta, nuclei, onehot, ckpt, ybar1, ybar0, gpart, &counter, B, n_a, jets_a, jets_b, stride,
               Na, n_types, n_steps, t0, t1, y0, nullptr, 0, 0};
  if (n_layers == 1) run_block(gnn_general_bwd_kernel<1>, a);
  else if (n_layers == 3) run_block(gnn_general_bwd_kernel<3>, a);
  else return -2;
  return 0;
}
