/* ofdft_b200.h -- C ABI of the B200-native (sm_100a) Monte-Carlo energy-estimator hot path.
 *
 * The reference (ChemAI-Lab/ofdft_nflows) is pure Python/JAX and has no FFI; the "interface" each
 * entry point replaces is therefore a Python call signature, cited per function below
 * (paths relative to the reference checkout).  A JAX host binds these through jax.ffi custom calls
 * wrapped in jax.custom_vjp; the in-repo host mirror binds them with ctypes.
 *
 * Conventions (all entry points):
 *   - plain C: device pointers + PODs only; `stream` is a cudaStream_t passed as void*;
 *   - asynchronous on `stream`; never allocates, never synchronises, never owns memory;
 *     scratch is caller-provided and sized by the companion *_scratch_bytes();
 *   - returns 0 on success, a positive cudaError_t from the launch, or a negative OFDFT_E* code;
 *   - re-entrant / thread-safe (no mutable global state);
 *   - all floating-point buffers are float32 unless stated; energy sums are float64.
 *
 * Flat parameter vector `theta` (key-sorted ravel_pytree order of the Flax pytree):
 *   [last_layer.bias(n_out) | last_layer.kernel(H_L*n_out) | layers_0.bias(H) | layers_0.kernel(I*H) |
 *    layers_1.bias(H) | layers_1.kernel(H*H) | ...],  kernels (in,out) row-major.
 *   GNN field: I = 2 + n_types, input order [t, r, onehot]  (ofdft_normflows/equiv_flows.py:51)
 *   MLP field: I = 1 + d,       input order [t, x]          (ofdft_normflows/cn_flows.py:133-134)
 *
 * State rows ("trajectories"), row-major: 3-D [x0 x1 x2 | logp | s0 s1 s2], 1-D [x | logp | s]
 *   (ofdft_normflows/utils.py:98-99,133-134).
 */
#ifndef OFDFT_B200_H_
#define OFDFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDFT_B200_ABI_VERSION 3

/* error codes (negative) */
#define OFDFT_EINVAL      (-1)   /* bad argument (null pointer, size <= 0, ...)           */
#define OFDFT_EUNSUPPORTED (-2)  /* shape outside what the kernels are instantiated for   */
#define OFDFT_ESCRATCH    (-3)   /* scratch buffer too small                               */

/* kernel-path selector (last argument of the integrator entry points).  The default runs every GEMM-shaped phase on the
 * tcgen05 tensor cores as 3xTF32; the FP32-pipe kernels are kept as an independently written second implementation that
 * the parity tests run against the same oracle (and as the recompute path of the GNN adjoint). */
#define OFDFT_PATH_DEFAULT 0
#define OFDFT_PATH_FFMA    1     /* integrator / adjoint sweep on the FP32 FFMA pipe                       */
#define OFDFT_PATH_FFMA_DW 2     /* 1-D flow: weight-gradient GEMM on the FP32 pipe (bit-or with the above) */
#define OFDFT_PATH_ACT_CKPT 4    /* 1-D flow, tensor-core path: `ckpt` is ofdft_cnf_mlp1d_act_ckpt_bytes() large; the forward
                                    also keeps every evaluation's layer activations there and the adjoint reads them
                                    instead of recomputing the hidden layers (pass the flag to both calls)          */
#define OFDFT_PATH_DEEP    8     /* 3-D flow, two hidden layers: run the depth-generic tensor-core kernels (the default of
                                    three hidden layers) instead of the pipelined two-layer ones -- a cross-check      */

/* jets carried per trajectory */
#define OFDFT_JETS_X      1      /* x only                  (second-half rows: only xp is used, H20_...py:154-156) */
#define OFDFT_JETS_XLOGP  2      /* x, logp                 (neural_ode, jax_ode.py:9-49)                           */
#define OFDFT_JETS_FULL   3      /* x, logp, score          (neural_ode_score, jax_ode.py:51-110)                   */

/* functional selectors (lower-cased string keys of functionals.py factories) */
enum {
  OFDFT_KIN_NONE = 0, OFDFT_KIN_TF = 1, OFDFT_KIN_W = 2, OFDFT_KIN_TF_W = 3,   /* functionals.py:21-43   */
  OFDFT_KIN_TF1D = 4, OFDFT_KIN_TFW_1D = 5
};
enum {
  OFDFT_HART_NONE = 0, OFDFT_HART_COULOMB = 1,      /* 'hartree'    functionals.py:463-486 (eps per component) */
  OFDFT_HART_SOFTC = 2,                             /* 'softc'      functionals.py:439-461 (1-D, no 1/2)        */
  OFDFT_HART_MT = 3                                 /* 'hartree_mt' functionals.py:489-496                      */
};
enum {
  OFDFT_NUC_NONE = 0, OFDFT_NUC_POINT = 1,          /* 'nuclei_potential' functionals.py:556-587 (eps on r)    */
  OFDFT_NUC_HGH = 2,                                /* 'hgh' functionals.py:617-655 (H parameters for all)     */
  OFDFT_NUC_ATTRACTION = 3,                         /* 'attraction' functionals.py:523-549 (1-D)               */
  OFDFT_NUC_HGH_SPECIES = 4                         /* NON-reference: H/O tables by Z (hgh_pseudopotentials.py:8-11) */
};
enum {
  OFDFT_X_NONE = 0, OFDFT_X_LDA = 1,                /* 'lda' functionals.py:391-417 ((3/pi)**1/3 as coded)     */
  OFDFT_X_B88 = 2,                                  /* 'b88_x_e' functionals.py:335-389                        */
  OFDFT_X_DIRAC_B88 = 3                             /* 'dirac_b88_x_e' = lda + b88_x_e (functionals.py:181-183) */
};
enum {
  OFDFT_C_NONE = 0, OFDFT_C_VWN = 1,                /* 'vwn_c_e'  functionals.py:185-245                       */
  OFDFT_C_PW92 = 2,                                 /* 'pw92_c_e' functionals.py:247-292                       */
  OFDFT_C_XC1D = 3                                  /* 'xc_1d'    functionals.py:294-332 (1-D, LiH.py:217)     */
};

/* energy-term slots of the float64 sums written by ofdft_functionals_fwd_bwd (F_values, H20_...py:43-49) */
#define OFDFT_TERM_KIN  0
#define OFDFT_TERM_NUC  1
#define OFDFT_TERM_HART 2
#define OFDFT_TERM_X    3
#define OFDFT_TERM_C    4
#define OFDFT_N_TERMS   8

typedef struct {
  int32_t d;             /* 3 or 1                                                           */
  int32_t kin, hart, nuc, x;
  int32_t c;             /* c-type functional (den, Ne): OFDFT_C_*                           */
  int32_t n_nuclei;      /* 3-D: rows of coords/z                                            */
  float   Ne;
  float   R, Z_alpha, Z_beta;     /* 1-D attraction (LiH.py:218-223)                         */
} ofdft_energy_spec;

/* ---- library ---------------------------------------------------------------------------------- */
int ofdft_b200_abi_version(void);
/* 0 if a device of compute capability 10.x is current, negative otherwise */
int ofdft_b200_check_device(void);
const char* ofdft_b200_strerror(int code);

/* ---- 3-D equivariant GNN flow: fused fixed-step (RK4) integrator -------------------------------
 * Replaces neural_ode / neural_ode_score (ofdft_normflows/jax_ode.py:9-49, 51-110) evaluated on
 * Gen_EqvFlow (ofdft_normflows/equiv_flows.py:108-127) with features = (64,) * n_layers, n_layers in {1, 2, 3}
 * (OFDFT_NF.py:320-326; every other driver hard-codes (64, 64)).  n_layers = 2 runs on the pipelined tcgen05 tile kernels,
 * n_layers = 3 on the depth-generic tcgen05 kernels (layer states in tensor memory, no activation checkpoint: act_ckpt is
 * ignored), n_layers = 1 (no hidden GEMM) on a plain FP32 kernel pair; OFDFT_PATH_FFMA selects FP32-pipe kernels for 2 and 3.
 *   theta   : ofdft_cnf_gnn_theta_size(n_types, n_layers) floats (4289 + 64*(2+n_types) for two layers)
 *   nuclei  : (Na,3)     onehot : (Na,n_types)
 *   y_in    : (B, in_stride)  rows [x | logp | s]; columns beyond what the row's jets need are ignored
 *   y_out   : (B, out_stride) same layout at t1 (only the carried columns are written)
 *   rows [0, n_a) carry jets_a, rows [n_a, B) carry jets_b (one launch, heavy tiles first)
 *   ckpt    : NULL, or ofdft_cnf_gnn_ckpt_bytes(B, n_steps) bytes receiving every RK4 stage state
 *             (x, s) for the backward pass
 *   act_ckpt: NULL, or ofdft_cnf_gnn_act_ckpt_bytes(...) bytes receiving the layer-2 activation jets of every
 *             RHS evaluation (spends HBM capacity -- ~100 KB per full-jet trajectory at Na=3, RK4x16 -- to
 *             remove the recompute GEMM from the backward pass; NULL selects recomputation)
 *   y0sign  : +1 (bool_neg=False) or -1 (bool_neg=True): t' = y0*t, output scaled by y0
 *   scratch : ofdft_cnf_gnn_fwd_scratch_bytes() bytes
 *   B = 0 (an empty shard) is a no-op returning OFDFT_OK; the adjoint of an empty shard writes a zero theta_bar.
 *   Na <= 128 (C27H46O, utils.py:327, has 74), n_types <= 8 and at most 8 DISTINCT one-hot rows (the kernels keep the
 *   first-layer offset b1 + onehot_i . W1[2:,:] on chip per distinct row); larger shapes return OFDFT_EUNSUPPORTED
 *   (more than 8 distinct rows trap in the kernel: the rows live in device memory).
 */
int64_t ofdft_cnf_gnn_theta_size(int32_t n_types, int32_t n_layers);
size_t ofdft_cnf_gnn_ckpt_bytes(int64_t B, int32_t n_steps);
size_t ofdft_cnf_gnn_act_ckpt_bytes(int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b, int32_t Na, int32_t n_steps);
size_t ofdft_cnf_gnn_fwd_scratch_bytes(void);
int ofdft_cnf_gnn_fwd(void* stream, const float* theta, const float* nuclei, const float* onehot,
                      const float* y_in, float* y_out, float* ckpt, float* act_ckpt,
                      void* scratch, size_t scratch_bytes,
                      int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b,
                      int32_t in_stride, int32_t out_stride,
                      int32_t Na, int32_t n_types, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign,
                      int32_t path);

/* Discrete adjoint of ofdft_cnf_gnn_fwd (replaces the odeint custom_vjp the reference reaches through
 * jax.value_and_grad, H20_mol_ofdft_min_equiv_flows.py:177-178).
 *   act_ckpt  : the buffer the forward pass filled, or NULL (recompute)
 *   ybar1     : (B, bar_stride) cotangent of y(t1), rows [xbar | logpbar | sbar]
 *   theta_bar : n_theta floats, OVERWRITTEN with d<ybar1, y(t1)>/d theta
 *   ybar0     : NULL or (B, bar_stride) cotangent of y(t0)
 *   scratch   : ofdft_cnf_gnn_bwd_scratch_bytes(B, n_a, n_types, n_layers) bytes (tile counter, 128 KB of prepared hi / lo weight
 *               operands, one partial gradient per 128-row tile reduced in fixed order: theta_bar is bitwise reproducible)
 */
size_t ofdft_cnf_gnn_bwd_scratch_bytes(int64_t B, int64_t n_a, int32_t n_types, int32_t n_layers);
int ofdft_cnf_gnn_bwd(void* stream, const float* theta, const float* nuclei, const float* onehot,
                      const float* ckpt, const float* act_ckpt, const float* ybar1, float* theta_bar, float* ybar0,
                      void* scratch, size_t scratch_bytes,
                      int64_t B, int64_t n_a, int32_t jets_a, int32_t jets_b, int32_t bar_stride,
                      int32_t Na, int32_t n_types, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign,
                      int32_t path);

/* One evaluation of the vector field, f.apply(params, t, states) (equiv_flows.py:124-127) or the
 * score-augmented right-hand side _evol_fn_i (jax_ode.py:80-94):  dy (B, stride) <- d/dt y. */
int ofdft_cnf_gnn_rhs(void* stream, const float* theta, const float* nuclei, const float* onehot,
                      const float* y, float* dy, void* scratch, size_t scratch_bytes,
                      int64_t B, int32_t jets, int32_t stride, int32_t Na, int32_t n_types, int32_t n_layers, float t,
                      float y0sign, int32_t path);

/* ---- 1-D tanh-MLP flow (Gen_CNFSimpleMLP, cn_flows.py:166-182; LiH.py:64-65 uses (512,512,512)) --
 * Same contract as the GNN entry points with d = 1: rows [x | logp | s]; theta per the layout above
 * with I = 2.  Supported: n_layers in [2,4], hidden width H multiple of 64 up to 512; the backward pass is
 * instantiated for jets == OFDFT_JETS_FULL.  scratch (ofdft_cnf_mlp1d_scratch_bytes) holds 16-byte-aligned
 * (and transposed) copies of the hidden kernels and, for the backward pass, the per-stage activation /
 * cotangent jets of a 2048-row chunk that the weight-gradient GEMM contracts (~2 MB per row at H=512, RK4x16).
 * ofdft_cnf_mlp1d_act_ckpt_bytes: size of `ckpt` under OFDFT_PATH_ACT_CKPT (stage states + ~1.2 MB per row of layer
 * activations at 3 x 512, RK4x16); 0 when the shape does not support it (B above one 2048-row backward chunk, or H not
 * a multiple of 128).  With the flag the 1-D adjoint also forks its weight-gradient GEMM onto a library-owned stream
 * and joins it before returning (stream order as seen by the caller is unchanged).
 */
size_t ofdft_cnf_mlp1d_ckpt_bytes(int64_t B, int32_t n_steps);
size_t ofdft_cnf_mlp1d_act_ckpt_bytes(int64_t B, int32_t H, int32_t n_layers, int32_t n_steps);
size_t ofdft_cnf_mlp1d_scratch_bytes(int64_t B, int32_t H, int32_t n_layers, int32_t n_steps);
int ofdft_cnf_mlp1d_fwd(void* stream, const float* theta, const float* y_in, float* y_out, float* ckpt,
                        void* scratch, size_t scratch_bytes, int64_t B, int32_t jets, int32_t stride,
                        int32_t H, int32_t n_layers, int32_t n_steps, float t0, float t1, float y0sign, int32_t path);
/* One evaluation of the vector field, f.apply(params, t, states) (cn_flows.py:179-182), or of the score-augmented
 * right-hand side _evol_fn_i (jax_ode.py:80-94):  dy (B, stride) <- d/dt y at time t. */
int ofdft_cnf_mlp1d_rhs(void* stream, const float* theta, const float* y, float* dy, void* scratch, size_t scratch_bytes,
                        int64_t B, int32_t jets, int32_t stride, int32_t H, int32_t n_layers, float t, float y0sign);
int ofdft_cnf_mlp1d_bwd(void* stream, const float* theta, const float* ckpt, const float* ybar1,
                        float* theta_bar, float* ybar0, void* scratch, size_t scratch_bytes,
                        int64_t B, int32_t jets, int32_t stride, int32_t H, int32_t n_layers,
                        int32_t n_steps, float t0, float t1, float y0sign, int32_t path);

/* ---- functionals: fused per-term reductions + cotangents ----------------------------------------
 * Replaces the functional calls + jnp.mean of `loss` (H20_mol_ofdft_min_equiv_flows.py:150-173,
 * LiH.py:118-138) and their VJP.
 *   y1     : (2*bs, stride) state at t1; rows [0,bs) = x/logp/score, rows [bs,2bs) supply xp
 *   coords : (n_nuclei,3), z : (n_nuclei)  (mol_info, H20_...py:76); HOST pointers (static molecule
 *            tables -- XLA attributes under jax.ffi -- copied into the launch arguments; n_nuclei <= 128);
 *            ignored when d == 1
 *   sums   : OFDFT_N_TERMS float64: per-term MEANS [kin, vnuc, hart, x, c, total, 0, 0]
 *   ybar1  : NULL or (2*bs, stride): d(total mean)/d y1 (all columns written)
 *   scratch: ofdft_functionals_scratch_bytes(bs)
 */
size_t ofdft_functionals_scratch_bytes(int64_t bs);
int ofdft_functionals_fwd_bwd(void* stream, const ofdft_energy_spec* spec, const float* coords, const float* z,
                              const float* y1, int32_t stride, int64_t bs, double* sums, float* ybar1,
                              void* scratch, size_t scratch_bytes);

/* Per-sample integrands (bs,1) of the string-keyed factories _kinetic/_hartree/_nuclear/
 * _exchange_correlation (functionals.py:21-43,164-183,425-436,504-521):
 *   e[term*bs + i] for term in {KIN, NUC, HART, X, C} (5*bs floats); inputs as ofdft_functionals_fwd_bwd. */
int ofdft_functionals_integrands(void* stream, const ofdft_energy_spec* spec, const float* coords, const float* z,
                                 const float* y1, int32_t stride, int64_t bs, float* e);

/* One functional of the string-keyed factories on separate array arguments, with its derivatives (the reference's
 * per-term call signatures, functionals.py):
 *   family 0: _kinetic(name)(den, score, Ne)              a0 = den (n), a1 = score (n,d)   g0 = d/d den, g1 = d/d score
 *   family 1: _exchange_correlation(name)(den, score, Ne)  same arguments (x-type: spec.x)
 *   family 4: _exchange_correlation(name)(den, Ne)          a0 = den (n)  (c-type: spec.c; a1 ignored)   g0 = d/d den
 *   family 2: _hartree(name)(x, xp, Ne)                   a0 = x (n,d), a1 = xp (n,d)      g0 = d/dx, g1 = d/dxp
 *   family 3: _nuclear(name)(x, Ne, mol_info) / attraction a0 = x (n,d)                     g0 = d/dx
 * out (n) receives the per-sample integrand; g0 / g1 may be NULL.  The functional is selected by the matching
 * field of `spec` (the others are ignored); coords / z are HOST pointers as above. */
int ofdft_functional_eval(void* stream, int32_t family, const ofdft_energy_spec* spec, const float* coords, const float* z,
                          const float* a0, const float* a1, int64_t n, float* out, float* g0, float* g1);

/* ---- all-pairs Hartree (north-star generalisation of the paired estimator) ----------------------
 * sum_{i<n, j<m} w / sqrt(|x_i - xp_j|^2 + d*eps)   (kind COULOMB, d = 3, w = 0.5 Ne^2)
 * sum_{i,j}      w / sqrt(1 + (x_i - xp_j)^2)       (kind SOFTC,   d = 1, w = Ne^2)
 *   x (n, x_stride), xp (m, xp_stride): first d columns are coordinates (state rows as they lie, or packed coordinates)
 *   norm_pairs : number of pairs the MEAN is taken over; 0 selects n*m.  A rank that evaluates one block of a larger
 *                pair set (its strip of the all-gathered batch) passes the size of the whole set
 *   esum   : 1 float64, (1/norm_pairs) * sum
 *   xbar   : NULL or (n, xbar_stride) d(esum)/dx in the first d columns;  xpbar : NULL or (m, xpbar_stride)
 *   accumulate != 0 ADDS into esum / xbar / xpbar (e.g. the x columns of a state cotangent) instead of overwriting
 *   scratch: ofdft_hartree_allpairs_scratch_bytes(n, m)
 */
size_t ofdft_hartree_allpairs_scratch_bytes(int64_t n, int64_t m);
int ofdft_hartree_allpairs(void* stream, int32_t kind, int32_t d, float Ne, const float* x, int32_t x_stride,
                           const float* xp, int32_t xp_stride, int64_t n, int64_t m, double norm_pairs, double* esum,
                           float* xbar, int32_t xbar_stride, float* xpbar, int32_t xpbar_stride, int32_t accumulate,
                           void* scratch, size_t scratch_bytes);

/* ---- callers either side of the path --------------------------------------------
 * Promolecular prior (ofdft_normflows/promolecular_distrax.py:52-62,86-89): mixture of unit-variance Gaussians on the
 * nuclei with weights Z_i / sum Z.  x (n, x_stride) -> out (n) = log p(x) [ - sub[i*sub_stride] ] [ exp() if exp_out ],
 * score (n,3) = grad log p(x); out or score may be NULL.  With sub = Delta logp of the reverse flow and exp_out = 1
 * this is the last step of rho_rev (H20_mol_ofdft_min_equiv_flows.py:132-137).  coords / z are HOST pointers. */
int ofdft_promolecular_logp_score(void* stream, const float* coords, const float* z, int32_t n_nuclei, const float* x,
                                  int32_t x_stride, int64_t n, const float* sub, int32_t sub_stride, int32_t exp_out,
                                  float* out, float* score);
/* batch_generator (ofdft_normflows/utils.py:76-108): out (2*bs, 7) rows [x, log rho0(x), grad log rho0(x)], two
 * independent draws; Philox stream (seed, subsequence = row, skip-ahead = 8*offset outputs): pass offset = call index
 * (0, 1, 2, ...) -- every call then reads a disjoint stretch of each row's stream (a row consumes 5 outputs per call).  1-D variant
 * (utils.py:110-143 with the N(0,1) prior of LiH.py:78): out (2*bs, 3). */
int ofdft_batch_generator_3d(void* stream, uint64_t seed, uint64_t offset, const float* coords, const float* z,
                             int32_t n_nuclei, int64_t bs, float* out);
int ofdft_batch_generator_1d(void* stream, uint64_t seed, uint64_t offset, int64_t bs, float* out);

/* Optimiser step of the drivers: optax.chain(clip_by_global_norm(max_norm), rmsprop(lr))
 * (H20_mol_ofdft_min_equiv_flows.py:111-116; optax defaults decay = 0.9, eps = 1e-8 inside the square root, nu0 = 0).
 * In place on theta (n) and the second-moment state nu (n); max_norm <= 0 disables the clip. */
int ofdft_rmsprop_clip_step(void* stream, float* theta, const float* grad, float* nu, int64_t n, float lr, float decay,
                            float eps, float max_norm);

/* optax.chain(clip_by_global_norm(max_norm), adam(lr)) as OFDFT_NF.py:104-108 builds it (optax defaults b1 = 0.9, b2 = 0.999,
 * eps = 1e-8 outside the square root, eps_root = 0).  In place on theta and the moment states mu, nu (n each, initially zero);
 * step is the 1-based update count (bias correction). */
int ofdft_adam_clip_step(void* stream, float* theta, const float* grad, float* mu, float* nu, int64_t n, int64_t step, float lr,
                         float b1, float b2, float eps, float max_norm);

/* Running average of the per-term energy means, optax.ema(decay, debias=True) as the drivers log it
 * (H20_mol_ofdft_min_equiv_flows.py:117-118, 193-196): state <- decay*state + (1-decay)*values, out = state / (1 - decay^step).
 * values / state / out are DEVICE doubles (values is the `sums` output of ofdft_functionals_fwd_bwd), so a training loop can
 * append to a device-resident history (out = history + step*n) without a host synchronisation per epoch.  step is 1-based. */
int ofdft_ema_update(void* stream, const double* values, double* state, double* out, int32_t n, int64_t step, double decay);

/* FP32 FFMA / SFU peak microbenchmarks (roofline denominators; bench.py).  Run `iters` dependent
 * rounds on every SM; return via *out the number of FFMA (or MUFU.RSQ) lane-operations issued. */
int ofdft_microbench_ffma(void* stream, float* sink, int32_t iters, double* ops_out);
int ofdft_microbench_sfu(void* stream, float* sink, int32_t iters, double* ops_out);

/* Diagnostic: D (128 x 64, all TMEM lanes dumped) = A (M x 64) . W (64 x 64) on tcgen05 kind::tf32 with
 * `passes` = 1 (plain TF32) or 3 (3xTF32 error-compensated); M in {64, 128}.  Used by tests to pin the hand-written
 * descriptor / TMEM plumbing that the tensor-core integrator builds on. */
int ofdft_tc_gemm_test(void* stream, const float* A, const float* W, float* D, int32_t M, int32_t passes);
/* diagnostic: clocks for `reps` back-to-back tcgen05.mma (kind::tf32, M = 128, K = 8) of width N, A operand from shared
 * memory (ts = 0) or tensor memory (ts = 1), accumulating round-robin into nacc (1..4) accumulators -- consecutive MMAs
 * into the same accumulator are dependent; out = one int64 on the device. */
int ofdft_tc_mma_rate(void* stream, long long* out, int32_t N, int32_t ts, int32_t reps, int32_t distinct_b, int32_t nacc);

#ifdef __cplusplus
}
#endif
#endif /* OFDFT_B200_H_ */
