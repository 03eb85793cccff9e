// functionals.cu -- fused OF-DFT functional reductions + cotangents, per-sample integrands, and the tiled
// all-pairs Hartree kernel.  Reference: ofdft_normflows/functionals.py (per-function lines cited below) and the
// `loss` closures H20_mol_ofdft_min_equiv_flows.py:150-173 / LiH.py:118-138.
// The local reductions are HBM-bound (read x, xp, logp, s; write cotangents); the all-pairs kernel is bound by the
// FP32/SFU pipes (12 B per point amortised over a 256-point shared-memory tile).
#include "common.cuh"

namespace ofdft {

constexpr int kFnThreads = 256;
constexpr int kFnMaxNuc = 128;

struct FnArgs {
  ofdft_energy_spec spec;
  float coords[kFnMaxNuc * 3];
  float z[kFnMaxNuc];
  const float* y1; int stride; long long bs;
  double* partial;    // [grid][5]
  float* ybar1;       // (2bs, stride) or null
  float* integrands;  // (5, bs) or null
};

// HGH local pseudopotential parameters (ofdft_normflows/hgh_pseudopotentials.py:8-11)
struct Hgh { float zion, rloc, c1, c2, c3, c4; };
__device__ __forceinline__ Hgh hgh_H() { return {1.f, 0.2f, -4.180237f, 0.725075f, 0.f, 0.f}; }
__device__ __forceinline__ Hgh hgh_O() { return {6.f, 0.247621f, -16.580318f, 2.395701f, 0.f, 0.f}; }

struct Terms { float kin, nuc, hart, x, c; float gx[3], gxp[3], gl, gs[3]; };

// one first-half row i with its partner row bs+i: integrands and their derivatives (un-normalised)
__device__ __forceinline__ Terms eval_row(const FnArgs& a, long long i) {
  const ofdft_energy_spec& sp = a.spec;
  const int d = sp.d;
  const float Ne = sp.Ne;
  const float* row = a.y1 + i * a.stride;
  const float* rowp = a.y1 + (a.bs + i) * a.stride;
  Terms t;
  t.kin = t.nuc = t.hart = t.x = t.c = 0.f; t.gl = 0.f;
  float x[3] = {0, 0, 0}, xp[3] = {0, 0, 0}, s[3] = {0, 0, 0};
  for (int c = 0; c < d; ++c) { x[c] = row[c]; xp[c] = rowp[c]; s[c] = row[d + 1 + c]; }
  const float logp = row[d];
#pragma unroll
  for (int c = 0; c < 3; ++c) { t.gx[c] = 0.f; t.gxp[c] = 0.f; t.gs[c] = 0.f; }
  // ---- kinetic (functionals.py:70-158) ----
  const int kin = sp.kin;
  if (kin == OFDFT_KIN_TF || kin == OFDFT_KIN_TF_W) {
    const float cTF = 2.871234000188191f;                         // (3/10)(3 pi^2)^(2/3)
    const float e = cTF * powf(Ne, 5.0f / 3.0f) * expf((2.0f / 3.0f) * logp);   // c Ne^{5/3} den^{2/3}
    t.kin += e; t.gl += (2.0f / 3.0f) * e;
  }
  if (kin == OFDFT_KIN_TF1D || kin == OFDFT_KIN_TFW_1D) {
    const float e = 0.41123351671205655f * Ne * Ne * Ne * expf(2.0f * logp);    // (pi^2/24) Ne^3 den^2
    t.kin += e; t.gl += 2.0f * e;
  }
  if (kin == OFDFT_KIN_W || kin == OFDFT_KIN_TF_W || kin == OFDFT_KIN_TFW_1D) {
    const float cw = 0.2f * Ne * 0.125f;                          // lambda0 Ne / 8
    t.kin += cw * (s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int c = 0; c < d; ++c) t.gs[c] = 2.0f * cw * s[c];
  }
  // ---- exchange: lda (functionals.py:391-417; (3/pi)**1/3 == 1/pi as coded) ----
  if (sp.x == OFDFT_X_LDA || sp.x == OFDFT_X_DIRAC_B88) {
    const float l = -0.75f * powf(Ne, 4.0f / 3.0f) * 0.3183098861837907f;
    const float e = l * expf((1.0f / 3.0f) * logp);
    t.x = e; t.gl += (1.0f / 3.0f) * e;
  }
  // ---- exchange: B88 gradient correction (functionals.py:335-389): -beta Ne^{2/3} S^2 den^{2/3} / (1 + 6 beta x asinh x),
  //      x = S den^{-1/3}, S = |score|  (the reference evaluates it through log2/2** of clipped arguments) ----
  if (sp.x == OFDFT_X_B88 || sp.x == OFDFT_X_DIRAC_B88) {
    const float beta = 0.0042f;
    const float S2 = s[0] * s[0] + s[1] * s[1] + s[2] * s[2];
    const float S = sqrtf(S2);
    const float d23 = expf((2.0f / 3.0f) * logp), dm13 = expf((-1.0f / 3.0f) * logp);
    const float xs = S * dm13;
    const float ash = asinhf(xs);
    const float D = 1.0f + 6.0f * beta * xs * ash;
    const float Nn = S2 * d23;
    const float pre = -beta * powf(Ne, 2.0f / 3.0f);
    const float e = pre * Nn / D;
    const float dDdx = 6.0f * beta * (ash + xs / sqrtf(1.0f + xs * xs));
    t.x += e;
    t.gl += pre * ((2.0f / 3.0f) * Nn / D + Nn / (D * D) * dDdx * (xs * (1.0f / 3.0f)));
    if (S > 0.f) {
      const float cs = pre * (2.0f * d23 / D - Nn / (D * D) * dDdx * dm13 / S);
      for (int c = 0; c < d; ++c) t.gs[c] += cs * s[c];
    }
  }
  // ---- correlation (c-type functionals of the density only) ----
  if (sp.c == OFDFT_C_VWN) {            // functionals.py:185-245
    const float A = 0.0621814f, b = 3.72744f, cc = 12.9352f, x0 = -0.10498f;
    const float rs = 0.6203504908994f * expf((-1.0f / 3.0f) * logp);       // (3/(4 pi))^{1/3} den^{-1/3}
    const float x = sqrtf(rs);
    const float X = x * x + b * x + cc, X0 = x0 * x0 + b * x0 + cc, Q = sqrtf(4.0f * cc - b * b);
    const float at = atanf(Q / (2.0f * x + b));
    const float e = 0.5f * A * (2.0f * logf(x) - logf(X) + 2.0f * b / Q * at
                                - b * x0 / X0 * (logf((x - x0) * (x - x0) / X) + 2.0f * (2.0f * x0 + b) / Q * at));
    const float Xp = 2.0f * x + b;
    const float dat = -2.0f * Q / (Xp * Xp + Q * Q);
    const float dedx = 0.5f * A * (2.0f / x - Xp / X + 2.0f * b / Q * dat
                                   - b * x0 / X0 * (2.0f / (x - x0) - Xp / X + 2.0f * (2.0f * x0 + b) / Q * dat));
    t.c = Ne * e;
    t.gl += Ne * dedx * (-x * (1.0f / 6.0f));
  } else if (sp.c == OFDFT_C_PW92) {    // functionals.py:247-292
    const float A = 0.031091f, a1 = 0.21370f, b1 = 7.5957f, b2 = 3.5876f, b3 = 1.6382f, b4 = 0.49294f;
    const float rs = 0.6203504908994f * expf((-1.0f / 3.0f) * logp);
    const float sr = sqrtf(rs);
    const float G = 2.0f * A * (b1 * sr + b2 * rs + b3 * rs * sr + b4 * rs * rs);
    const float Gp = 2.0f * A * (0.5f * b1 / sr + b2 + 1.5f * b3 * sr + 2.0f * b4 * rs);
    const float L = log1pf(1.0f / G);
    const float e = -2.0f * A * (1.0f + a1 * rs) * L;
    const float dedrs = -2.0f * A * a1 * L + 2.0f * A * (1.0f + a1 * rs) * Gp / (G * G + G);
    t.c = Ne * e;
    t.gl += Ne * dedrs * (-rs * (1.0f / 3.0f));
  } else if (sp.c == OFDFT_C_XC1D) {    // functionals.py:294-332
    const float a0 = -0.8862269f, b0 = -2.1414101f, c0 = 0.4721355f, d0 = 2.81423f, e0 = 0.529891f, f0 = 0.458513f;
    const float g0 = -0.202642f, h0 = 0.470876f, al = 0.104435f, be = 4.11613f;
    const float rs = 1.0f / (2.0f * Ne) * expf(-logp);
    const float n1 = a0 + rs * (b0 + c0 * rs), d1 = 1.0f + rs * (d0 + rs * (e0 + f0 * rs));
    const float n1p = b0 + 2.0f * c0 * rs, d1p = d0 + rs * (2.0f * e0 + 3.0f * f0 * rs);
    const float rb = powf(rs, be);
    const float u = rs + al * rb;
    const float lu = logf(u);
    const float n2 = g0 * rs * lu, d2 = 1.0f + h0 * rs * rs;
    const float n2p = g0 * lu + g0 * rs * (1.0f + al * be * rb / rs) / u, d2p = 2.0f * h0 * rs;
    t.c = Ne * (n1 / d1 + n2 / d2);
    t.gl += Ne * ((n1p * d1 - n1 * d1p) / (d1 * d1) + (n2p * d2 - n2 * d2p) / (d2 * d2)) * (-rs);
  }
  // ---- Hartree, row-paired (functionals.py:439-496) ----
  if (sp.hart != OFDFT_HART_NONE) {
    float df[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]};
    float zz, w;
    if (sp.hart == OFDFT_HART_SOFTC) { zz = 1.0f + df[0] * df[0]; w = Ne * Ne; }
    else {
      const float eps = sp.hart == OFDFT_HART_COULOMB ? 1e-5f : 0.f;   // eps per component inside the sum
      zz = 0.f;
      for (int c = 0; c < d; ++c) zz += df[c] * df[c] + eps;
      w = 0.5f * Ne * Ne;
    }
    const float rs = 1.0f / sqrtf(zz);
    t.hart = w * rs;
    const float g = -w * rs / zz;
    for (int c = 0; c < d; ++c) { t.gx[c] += g * df[c]; t.gxp[c] -= g * df[c]; }
  }
  // ---- nuclear attraction (functionals.py:523-655) ----
  if (sp.nuc == OFDFT_NUC_ATTRACTION) {
    const float A = x[0] + 0.5f * sp.R, B = x[0] - 0.5f * sp.R;
    const float ia = 1.0f / sqrtf(1.0f + A * A), ib = 1.0f / sqrtf(1.0f + B * B);
    t.nuc = Ne * (-sp.Z_alpha * ia - sp.Z_beta * ib);
    t.gx[0] += Ne * (sp.Z_alpha * A * ia * ia * ia + sp.Z_beta * B * ib * ib * ib);
  } else if (sp.nuc != OFDFT_NUC_NONE) {
    float e = 0.f;
    for (int n = 0; n < sp.n_nuclei; ++n) {
      const float d0 = x[0] - a.coords[3 * n], d1 = x[1] - a.coords[3 * n + 1], d2 = x[2] - a.coords[3 * n + 2];
      const float r = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
      float dvdr;
      if (sp.nuc == OFDFT_NUC_POINT) {
        const float re = r + 1e-4f;                               // eps on r, not r^2
        e -= a.z[n] / re;
        dvdr = a.z[n] / (re * re);
      } else {
        const Hgh p = (sp.nuc == OFDFT_NUC_HGH_SPECIES && a.z[n] == 8.0f) ? hgh_O() : hgh_H();
        const float rr = r / p.rloc, rr2 = rr * rr;
        const float ex = expf(-0.5f * rr2);
        const float er = erff(rr * 0.7071067811865476f);
        const float poly = p.c1 + rr2 * (p.c2 + rr2 * (p.c3 + rr2 * p.c4));
        e += (-p.zion / r) * er + ex * poly;
        const float dpoly = rr * (2.0f * p.c2 + rr2 * (4.0f * p.c3 + rr2 * 6.0f * p.c4)) / p.rloc;
        dvdr = p.zion / (r * r) * er - (p.zion / r) * (0.7978845608028654f / p.rloc) * ex
               + ex * (dpoly - rr / p.rloc * poly);
      }
      const float g = Ne * dvdr / r;
      t.gx[0] += g * d0; t.gx[1] += g * d1; t.gx[2] += g * d2;
    }
    t.nuc = Ne * e;
  }
  return t;
}

__global__ void __launch_bounds__(kFnThreads) functionals_kernel(const FnArgs a) {
  const long long i = (long long)blockIdx.x * kFnThreads + threadIdx.x;
  double acc[5] = {0, 0, 0, 0, 0};
  if (i < a.bs) {
    const Terms t = eval_row(a, i);
    acc[0] = t.kin; acc[1] = t.nuc; acc[2] = t.hart; acc[3] = t.x; acc[4] = t.c;
    const int d = a.spec.d;
    if (a.integrands) {
      a.integrands[0 * a.bs + i] = t.kin; a.integrands[1 * a.bs + i] = t.nuc;
      a.integrands[2 * a.bs + i] = t.hart; a.integrands[3 * a.bs + i] = t.x; a.integrands[4 * a.bs + i] = t.c;
    }
    if (a.ybar1) {
      const float inv = 1.0f / (float)a.bs;
      float* o = a.ybar1 + i * a.stride;
      float* op = a.ybar1 + (a.bs + i) * a.stride;
      for (int c = 0; c < a.stride; ++c) { o[c] = 0.f; op[c] = 0.f; }
      for (int c = 0; c < d; ++c) { o[c] = t.gx[c] * inv; op[c] = t.gxp[c] * inv; o[d + 1 + c] = t.gs[c] * inv; }
      o[d] = t.gl * inv;
    }
  }
  if (a.partial == nullptr) return;
  __shared__ double red[5][kFnThreads / 32];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int w = 0; w < kFnThreads / 32; ++w) v += red[threadIdx.x][w];
    a.partial[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
  }
}

__global__ void functionals_final_kernel(const double* __restrict__ partial, int nblk, long long bs, double* sums) {
  __shared__ double red[8];
  const int k = threadIdx.x >> 5, l = threadIdx.x & 31;   // 256 threads: warp k reduces term k (5 used)
  double v = 0.0;
  if (k < 5) for (int b = l; b < nblk; b += 32) v += partial[(size_t)b * 5 + k];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (l == 0) red[k] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int t = 0; t < 5; ++t) { const double m = red[t] / (double)bs; sums[t] = m; tot += m; }
    sums[5] = tot; sums[6] = 0.0; sums[7] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// all-pairs Hartree: E = (w/(n m)) sum_ij 1/sqrt(|x_i - xp_j|^2 + c), force on x_i
// ------------------------------------------------------------------------------------------------
constexpr int kHpThreads = 256;
constexpr int kHpPts = 4;                 // i-points per thread
constexpr int kHpTile = 256;              // j-points per shared-memory tile

// packed FP32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: two independent fp32 operations per issue slot)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// One (i-block, j-split) CTA: every thread keeps kHpPts = 4 i-points as two PACKED pairs, so a pair evaluation costs
// 6 issue slots (+ 1 MUFU.RSQ) with forces and 3.5 (+ 1 MUFU) without, instead of 12 / 7 scalar FP32 instructions: the
// kernel is then bounded by the FP32 pipe (forces) or the SFU (energy only), not by the issue rate.  The j-tile holds the
// NEGATED xp coordinates so that the difference is one packed add.
template <int D, bool FORCES, bool ENERGY>
__global__ void __launch_bounds__(kHpThreads) hartree_pairs_kernel(const float* __restrict__ x, const float* __restrict__ xp,
                                                                  int stride, int stride_p, long long n, long long m, float cst,
                                                                  int splits, double* __restrict__ epart,
                                                                  float* __restrict__ fpart) {
  static_assert(kHpPts == 4, "two packed pairs of i-points per thread");
  __shared__ float4 tile[kHpTile];
  const long long i0 = ((long long)blockIdx.x * kHpThreads + threadIdx.x) * kHpPts;
  f32x2 xi[2][3], f[2][3], e[2];
#pragma unroll
  for (int pp = 0; pp < 2; ++pp) {
    const long long ia = i0 + 2 * pp < n ? i0 + 2 * pp : n - 1, ib = i0 + 2 * pp + 1 < n ? i0 + 2 * pp + 1 : n - 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xi[pp][c] = c < D ? pack2(x[ia * stride + c], x[ib * stride + c]) : pack2(0.f, 0.f);
      f[pp][c] = pack2(0.f, 0.f);
    }
    e[pp] = pack2(0.f, 0.f);
  }
  const f32x2 cst2 = pack2(cst, cst);
  const long long per = (m + splits - 1) / splits;
  const long long j_begin = (long long)blockIdx.y * per;
  const long long j_end = j_begin + per < m ? j_begin + per : m;
  double esum = 0.0;
  for (long long jt = j_begin; jt < j_end; jt += kHpTile) {
    __syncthreads();
    {
      const long long j = jt + threadIdx.x;
      float4 v = make_float4(-1e18f, -1e18f, -1e18f, 0.f);   // padded points sit "at infinity": rsqrt -> 0, force -> 0
      if (j < j_end) {
        v.x = -xp[j * stride_p];
        if (D == 3) { v.y = -xp[j * stride_p + 1]; v.z = -xp[j * stride_p + 2]; } else { v.y = 0.f; v.z = 0.f; }
      }
      tile[threadIdx.x] = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = 0; jj < kHpTile; ++jj) {
      const float4 q = tile[jj];
      const f32x2 q0 = pack2(q.x, q.x), q1 = pack2(q.y, q.y), q2 = pack2(q.z, q.z);
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        const f32x2 d0 = add2(xi[pp][0], q0);
        f32x2 zz = fma2(d0, d0, cst2);
        f32x2 d1 = 0, d2 = 0;
        if (D == 3) { d1 = add2(xi[pp][1], q1); d2 = add2(xi[pp][2], q2); zz = fma2(d1, d1, zz); zz = fma2(d2, d2, zz); }
        float za, zb, ra, rb;
        unpack2(zz, za, zb);
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(za));    // MUFU.RSQ without the denormal fix-up path
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(zb));
        const f32x2 rs = pack2(ra, rb);
        if (ENERGY) e[pp] = add2(e[pp], rs);
        if (FORCES) {
          const f32x2 g = mul2(mul2(rs, rs), rs);
          f[pp][0] = fma2(g, d0, f[pp][0]);
          if (D == 3) { f[pp][1] = fma2(g, d1, f[pp][1]); f[pp][2] = fma2(g, d2, f[pp][2]); }
        }
      }
    }
    // flush the fp32 tile sums into the float64 accumulator every tile
    if (ENERGY) {
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        float ea, eb;
        unpack2(e[pp], ea, eb);
        if (i0 + 2 * pp < n) esum += (double)ea;
        if (i0 + 2 * pp + 1 < n) esum += (double)eb;
        e[pp] = pack2(0.f, 0.f);
      }
    }
  }
  if (fpart != nullptr) {
#pragma unroll
    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        float fa, fb;
        unpack2(f[pp][c], fa, fb);
        if (i0 + 2 * pp < n) fpart[((size_t)blockIdx.y * n + (i0 + 2 * pp)) * D + c] = fa;
        if (i0 + 2 * pp + 1 < n) fpart[((size_t)blockIdx.y * n + (i0 + 2 * pp + 1)) * D + c] = fb;
      }
  }
  if (ENERGY) {
    __shared__ double red[kHpThreads / 32];
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = esum;
    __syncthreads();
    if (threadIdx.x == 0 && epart != nullptr) {
      double v = 0.0;
      for (int w = 0; w < kHpThreads / 32; ++w) v += red[w];
      epart[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = v;
    }
  }
}

__global__ void hartree_force_reduce_kernel(const float* __restrict__ fpart, float* __restrict__ out, long long n_el,
                                            int splits, float scale, int d, int out_stride, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_el) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += fpart[(size_t)k * n_el + i];
  const long long o = (i / d) * out_stride + (i % d);
  const float v = -scale * s;     // d/dx_i of 1/sqrt(z) = -(x_i - x_j)/z^{3/2}
  out[o] = accumulate ? out[o] + v : v;
}

__global__ void hartree_energy_reduce_kernel(const double* __restrict__ epart, int nblk, double scale, double* out, int accumulate) {
  __shared__ double red[8];
  double v = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 256) v += epart[b];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int w = 0; w < 8; ++w) s += red[w]; *out = accumulate ? *out + s * scale : s * scale; }
}

// ------------------------------------------------------------------------------------------------
// string-keyed functional factories (functionals.py:21-43,164-183,425-436,504-521): per-sample integrand (n,1)
// and its derivatives w.r.t. the two array arguments, on separate (den, score) / (x, xp) / (x) inputs
// ------------------------------------------------------------------------------------------------
struct FevArgs {
  ofdft_energy_spec spec;
  float coords[kFnMaxNuc * 3];
  float z[kFnMaxNuc];
  int family;                 // 0 kinetic(den, score)  1 exchange(den, score)  2 hartree(x, xp)  3 nuclear(x)  4 correlation(den)
  const float* a0; const float* a1; long long n;
  float* out; float* g0; float* g1;
};

__global__ void __launch_bounds__(kFnThreads) functional_eval_kernel(const FevArgs a) {
  const long long i = (long long)blockIdx.x * kFnThreads + threadIdx.x;
  if (i >= a.n) return;
  const int d = a.spec.d;
  // pack the separate arguments into the fused-path row layout and reuse eval_row's arithmetic
  float row[7] = {0, 0, 0, 0, 0, 0, 0}, rowp[7] = {0, 0, 0, 0, 0, 0, 0};
  float den = 1.0f;
  if (a.family <= 1 || a.family == 4) {
    den = a.a0[i];
    row[d] = logf(den);
    if (a.family <= 1) for (int c = 0; c < d; ++c) row[d + 1 + c] = a.a1[i * d + c];
  } else {
    for (int c = 0; c < d; ++c) { row[c] = a.a0[i * d + c]; if (a.family == 2) rowp[c] = a.a1[i * d + c]; }
  }
  FnArgs f;
  f.spec = a.spec;
  if (a.family != 0) f.spec.kin = OFDFT_KIN_NONE;
  if (a.family != 1) f.spec.x = OFDFT_X_NONE;
  if (a.family != 2) f.spec.hart = OFDFT_HART_NONE;
  if (a.family != 3) f.spec.nuc = OFDFT_NUC_NONE;
  if (a.family != 4) f.spec.c = OFDFT_C_NONE;
  for (int k = 0; k < a.spec.n_nuclei * 3; ++k) f.coords[k] = a.coords[k];
  for (int k = 0; k < a.spec.n_nuclei; ++k) f.z[k] = a.z[k];
  float buf[14];
  for (int c = 0; c < 7; ++c) { buf[c] = row[c]; buf[7 + c] = rowp[c]; }
  f.y1 = buf; f.stride = 7; f.bs = 1;
  // eval_row indexes the partner at (bs + i) * stride: bs = 1, i = 0 -> buf + 7
  Terms t = eval_row(f, 0);
  const float e = a.family == 0 ? t.kin : (a.family == 1 ? t.x : (a.family == 2 ? t.hart : (a.family == 3 ? t.nuc : t.c)));
  a.out[i] = e;
  if (a.family <= 1 || a.family == 4) {
    if (a.g0) a.g0[i] = t.gl / den;                              // d/d den = (d/d logp) / den
    if (a.g1 && a.family <= 1) for (int c = 0; c < d; ++c) a.g1[i * d + c] = t.gs[c];
  } else {
    if (a.g0) for (int c = 0; c < d; ++c) a.g0[i * d + c] = t.gx[c];
    if (a.g1 && a.family == 2) for (int c = 0; c < d; ++c) a.g1[i * d + c] = t.gxp[c];
  }
}

// j-range splits per i-block.  Every CTA does the same amount of work, so the grid is made MANY residency waves deep
// (148 SMs x 8 CTAs): with a grid of only one to two waves a bs/8 x bs strip left 14 % of the SM slots idle
// (640 CTAs = 4.3 per SM -> the SMs holding 4 waited for those holding 5).  A split keeps at least 8 tiles of j points.
static int hp_splits(long long n, long long m) {
  const long long blocks_i = (n + kHpThreads * kHpPts - 1) / (kHpThreads * kHpPts);
  long long s = (8 * 1184 + blocks_i - 1) / blocks_i;
  const long long max_s = (m + 8 * kHpTile - 1) / (8 * kHpTile);
  if (s > max_s) s = max_s;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return (int)s;
}
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// peak microbenchmarks (roofline denominators)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) microbench_ffma_kernel(float* sink, int iters) {
  float a[16];
  const float b = 1.0f + 1e-7f * threadIdx.x, c = 1e-9f * blockIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (float)i * 0.25f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) microbench_sfu_kernel(float* sink, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0f + (float)i + 1e-3f * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = rsqrtf(a[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace ofdft

using namespace ofdft;

extern "C" size_t ofdft_functionals_scratch_bytes(int64_t bs) {
  const size_t nblk = (size_t)((bs + kFnThreads - 1) / kFnThreads);
  return align256(nblk * 5 * sizeof(double));
}

static int fill_fn_args(FnArgs& a, const ofdft_energy_spec* spec, const float* coords_dev_or_host, const float* z) {
  (void)coords_dev_or_host; (void)z;
  if (!spec) return OFDFT_EINVAL;
  if (spec->d != 1 && spec->d != 3) return OFDFT_EUNSUPPORTED;
  a.spec = *spec;
  return 0;
}

// coords / z are HOST pointers (tiny molecule tables, copied into the kernel argument block)
static int launch_functionals(void* stream, const ofdft_energy_spec* spec, const float* coords, const float* z,
                              const float* y1, int32_t stride, int64_t bs, double* sums, float* ybar1, float* integrands,
                              void* scratch, size_t scratch_bytes) {
  FnArgs a;
  int rc = fill_fn_args(a, spec, coords, z);
  if (rc) return rc;
  if (!y1 || bs <= 0 || stride < 2 * spec->d + 1) return OFDFT_EINVAL;
  const bool needs_mol = spec->d == 3 && spec->nuc != OFDFT_NUC_NONE;
  if (needs_mol) {
    if (!coords || !z || spec->n_nuclei < 1) return OFDFT_EINVAL;
    if (spec->n_nuclei > kFnMaxNuc) return OFDFT_EUNSUPPORTED;
    for (int i = 0; i < spec->n_nuclei * 3; ++i) a.coords[i] = coords[i];
    for (int i = 0; i < spec->n_nuclei; ++i) a.z[i] = z[i];
  }
  a.y1 = y1; a.stride = stride; a.bs = bs; a.ybar1 = ybar1; a.integrands = integrands;
  const int nblk = (int)((bs + kFnThreads - 1) / kFnThreads);
  a.partial = nullptr;
  if (sums) {
    if (!scratch) return OFDFT_EINVAL;
    if (scratch_bytes < ofdft_functionals_scratch_bytes(bs)) return OFDFT_ESCRATCH;
    a.partial = (double*)scratch;
  }
  cudaStream_t st = (cudaStream_t)stream;
  functionals_kernel<<<nblk, kFnThreads, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (sums) {
    functionals_final_kernel<<<1, 256, 0, st>>>(a.partial, nblk, bs, sums);
    e = cudaGetLastError();
  }
  return (int)e;
}

extern "C" int ofdft_functionals_fwd_bwd(void* stream, const ofdft_energy_spec* spec, const float* coords, const float* z,
                                         const float* y1, int32_t stride, int64_t bs, double* sums, float* ybar1,
                                         void* scratch, size_t scratch_bytes) {
  if (!sums) return OFDFT_EINVAL;
  return launch_functionals(stream, spec, coords, z, y1, stride, bs, sums, ybar1, nullptr, scratch, scratch_bytes);
}

extern "C" int ofdft_functionals_integrands(void* stream, const ofdft_energy_spec* spec, const float* coords, const float* z,
                                            const float* y1, int32_t stride, int64_t bs, float* e) {
  if (!e) return OFDFT_EINVAL;
  return launch_functionals(stream, spec, coords, z, y1, stride, bs, nullptr, nullptr, e, nullptr, 0);
}

extern "C" int ofdft_functional_eval(void* stream, int32_t family, const ofdft_energy_spec* spec, const float* coords,
                                     const float* z, const float* a0, const float* a1, int64_t n, float* out, float* g0,
                                     float* g1) {
  if (!spec || !a0 || !out || n <= 0 || family < 0 || family > 4) return OFDFT_EINVAL;
  if ((family <= 2) && !a1) return OFDFT_EINVAL;
  if (spec->d != 1 && spec->d != 3) return OFDFT_EUNSUPPORTED;
  FevArgs a;
  a.spec = *spec;
  const bool needs_mol = family == 3 && spec->d == 3 && spec->nuc != OFDFT_NUC_NONE;
  if (needs_mol) {
    if (!coords || !z || spec->n_nuclei < 1) return OFDFT_EINVAL;
    if (spec->n_nuclei > kFnMaxNuc) return OFDFT_EUNSUPPORTED;
    for (int i = 0; i < spec->n_nuclei * 3; ++i) a.coords[i] = coords[i];
    for (int i = 0; i < spec->n_nuclei; ++i) a.z[i] = z[i];
  } else {
    a.spec.n_nuclei = 0;
  }
  a.family = family; a.a0 = a0; a.a1 = a1; a.n = n; a.out = out; a.g0 = g0; a.g1 = g1;
  functional_eval_kernel<<<(unsigned)((n + kFnThreads - 1) / kFnThreads), kFnThreads, 0, (cudaStream_t)stream>>>(a);
  return (int)cudaGetLastError();
}

extern "C" size_t ofdft_hartree_allpairs_scratch_bytes(int64_t n, int64_t m) {
  if (n <= 0 || m <= 0) return 256;
  const size_t s1 = (size_t)hp_splits(n, m), s2 = (size_t)hp_splits(m, n);
  const size_t blocks1 = (size_t)((n + kHpThreads * kHpPts - 1) / (kHpThreads * kHpPts));
  const size_t f1 = s1 * (size_t)n * 3 * sizeof(float), f2 = s2 * (size_t)m * 3 * sizeof(float);
  return align256(s1 * blocks1 * sizeof(double)) + align256(f1 > f2 ? f1 : f2);
}

extern "C" int ofdft_hartree_allpairs(void* stream, int32_t kind, int32_t d, float Ne, const float* x, int32_t x_stride,
                                      const float* xp, int32_t xp_stride, int64_t n, int64_t m, double norm_pairs, double* esum,
                                      float* xbar, int32_t xbar_stride, float* xpbar, int32_t xpbar_stride, int32_t accumulate,
                                      void* scratch, size_t scratch_bytes) {
  if (!x || !xp || !esum || !scratch || n <= 0 || m <= 0 || x_stride < d || xp_stride < d) return OFDFT_EINVAL;
  if ((xbar && xbar_stride < d) || (xpbar && xpbar_stride < d) || norm_pairs < 0.0) return OFDFT_EINVAL;
  float cst, w;
  if (kind == OFDFT_HART_COULOMB && d == 3) { cst = 3e-5f; w = 0.5f * Ne * Ne; }
  else if (kind == OFDFT_HART_SOFTC && d == 1) { cst = 1.0f; w = Ne * Ne; }
  else return OFDFT_EUNSUPPORTED;
  if (scratch_bytes < ofdft_hartree_allpairs_scratch_bytes(n, m)) return OFDFT_ESCRATCH;
  cudaStream_t st = (cudaStream_t)stream;
  const double scale = (double)w / (norm_pairs > 0.0 ? norm_pairs : (double)n * (double)m);
  for (int pass = 0; pass < 2; ++pass) {
    const float* a = pass == 0 ? x : xp;
    const float* b = pass == 0 ? xp : x;
    const long long na = pass == 0 ? n : m, nb = pass == 0 ? m : n;
    float* out = pass == 0 ? xbar : xpbar;
    const int sa = pass == 0 ? x_stride : xp_stride, sb = pass == 0 ? xp_stride : x_stride;
    const int bar_stride = pass == 0 ? xbar_stride : xpbar_stride;
    if (pass == 1 && out == nullptr) break;
    const int splits = hp_splits(na, nb);
    const unsigned blocks = (unsigned)((na + kHpThreads * kHpPts - 1) / (kHpThreads * kHpPts));
    double* epart = (double*)scratch;
    float* fpart = (float*)((char*)scratch + align256((size_t)hp_splits(n, m) *
                            (size_t)((n + kHpThreads * kHpPts - 1) / (kHpThreads * kHpPts)) * sizeof(double)));
    dim3 grid(blocks, (unsigned)splits);
    double* ep = pass == 0 ? epart : nullptr;
    float* fp = out ? fpart : nullptr;
    // pass 0 carries the energy (and the forces on x when requested), pass 1 only the forces on xp
    if (d == 3 && fp && ep) hartree_pairs_kernel<3, true, true><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    else if (d == 3 && fp) hartree_pairs_kernel<3, true, false><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    else if (d == 3) hartree_pairs_kernel<3, false, true><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    else if (fp && ep) hartree_pairs_kernel<1, true, true><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    else if (fp) hartree_pairs_kernel<1, true, false><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    else hartree_pairs_kernel<1, false, true><<<grid, kHpThreads, 0, st>>>(a, b, sa, sb, na, nb, cst, splits, ep, fp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (pass == 0) {
      hartree_energy_reduce_kernel<<<1, 256, 0, st>>>(epart, (int)(blocks * splits), scale, esum, accumulate);
      e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
    }
    if (out) {
      const long long n_el = na * d;
      hartree_force_reduce_kernel<<<(unsigned)((n_el + 255) / 256), 256, 0, st>>>(fpart, out, n_el, splits, (float)scale, d, bar_stride, accumulate);
      e = cudaGetLastError();
      if (e != cudaSuccess) return (int)e;
    }
  }
  return 0;
}

extern "C" int ofdft_microbench_ffma(void* stream, float* sink, int32_t iters, double* ops_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8;
  microbench_ffma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sink, iters);
  if (ops_out) *ops_out = (double)grid * 256.0 * (double)iters * 256.0;
  return (int)cudaGetLastError();
}

extern "C" int ofdft_microbench_sfu(void* stream, float* sink, int32_t iters, double* ops_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8;
  microbench_sfu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sink, iters);
  if (ops_out) *ops_out = (double)grid * 256.0 * (double)iters * 64.0;
  return (int)cudaGetLastError();
}

extern "C" int ofdft_b200_abi_version(void) { return OFDFT_B200_ABI_VERSION; }

extern "C" int ofdft_b200_check_device(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return OFDFT_EUNSUPPORTED;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return OFDFT_EUNSUPPORTED;
  return major == 10 ? 0 : OFDFT_EUNSUPPORTED;
}

extern "C" const char* ofdft_b200_strerror(int code) {
  if (code == 0) return "ok";
  if (code == OFDFT_EINVAL) return "ofdft_b200: invalid argument";
  if (code == OFDFT_EUNSUPPORTED) return "ofdft_b200: unsupported shape/configuration";
  if (code == OFDFT_ESCRATCH) return "ofdft_b200: scratch buffer too small";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "ofdft_b200: unknown error";
}
