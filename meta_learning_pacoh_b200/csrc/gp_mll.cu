// Batched GP marginal log-likelihood with analytic gradient, one warp per (particle, task) matrix, sm_100a.
//
// Replaces, for every (p, t) pair at once, what the reference does per task through gpytorch
// (meta_learn/random_gp.py:54-89 VectorizedGP.forward -> SEKernelLight models.py:428-446,
//  GaussianLikelihoodLight models.py:463-487, ExactMarginalLogLikelihood random_gp.py:83-85) and the
// autograd reverse pass through the Cholesky (svgd.py:16):
//
//   Kt   = s * exp(-1/2 |(z_a - z_b)/l|^2) + sigma^2 I            (s = 1 on the SVGD/VI path)
//   mll  = [ -1/2 r^T Kt^-1 r - 1/2 log det Kt - n/2 log 2pi ] / n,   r = y - m
//   dL/dKt = 1/2 (alpha alpha^T - Kt^-1),  dL/dm = alpha,             alpha = Kt^-1 r
//
// Data layout.  Lane i of the warp owns rows i and i+32 of the unit-diagonal-normalised matrix in REGISTERS (NC
// columns each) plus the augmented column r.  The factorisation is a BLOCKED symmetric Gauss-Jordan sweep with
// 4 pivots per block (13 blocks for n = 50):
//   1. every lane copies its 4 block-column entries t0 (by symmetry = the 4 pivot rows) out of the register
//      matrix (one jump-table switch, no dynamic register indexing) and publishes them to shared memory,
//   2. after ONE __syncwarp every lane inverts the 4x4 pivot block redundantly in registers (its pivots are
//      the squared Cholesky diagonal d_k = L_kk^2: same log det, same positive-definiteness test),
//   3. the rank-4 update  A_i. -= w_i . T0  runs over the whole row with broadcast LDS.128 operands:
//      416 independent FFMAs per block, no bank conflicts.
// After all blocks the registers hold -Kt^-1 and the augmented column holds Kt^-1 r.  The Gram matrix is never
// stored: exp(-d2) is recomputed (MUFU.EX2) for the gradient contraction.
#include <math_constants.h>
#include "common.cuh"
#include "kernels.cuh"

namespace pacoh {

namespace {

constexpr int kGpWarps = 4;
#ifndef PACOH_GP_MINB
#define PACOH_GP_MINB(NC) ((NC) > 32 ? 3 : 4)   // min CTAs per SM (EXPERIMENT: 3 -> 168 regs, some spills)
#endif
constexpr int kPubStride = 68;            // floats per published block column (64 rows + pad)
constexpr int kPubFloats = 4 * kPubStride + 8;   // 4 columns + 4 augmented entries of the pivot rows (+ pad)
constexpr float kFar = 1.0e18f;           // scaled feature of a padding row: exp2(-(1e18)^2) == 0 exactly

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_newton(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r * fmaf(-d, r, 2.0f);
}

// t[s][j] = A[s][4c + j] with a runtime block index c: a switch over compile-time register indices, so the
// register matrix never needs dynamic indexing (read only: writes would turn every A register into a phi).
template <int NC, int NS>
__device__ __forceinline__ void block_columns(const float (&A)[NS][NC], float (&t)[NS][4], int c) {
#define PACOH_BLK_CASE(I)                                              \
  case I:                                                              \
    if constexpr (4 * (I) < NC) {                                      \
      _Pragma("unroll") for (int s = 0; s < NS; ++s)                   \
          _Pragma("unroll") for (int j = 0; j < 4; ++j) {              \
        t[s][j] = A[s][4 * (I) + j];                                   \
      }                                                                \
    }                                                                  \
    break;
  switch (c) {
    PACOH_BLK_CASE(0) PACOH_BLK_CASE(1) PACOH_BLK_CASE(2) PACOH_BLK_CASE(3) PACOH_BLK_CASE(4) PACOH_BLK_CASE(5)
    PACOH_BLK_CASE(6) PACOH_BLK_CASE(7) PACOH_BLK_CASE(8) PACOH_BLK_CASE(9) PACOH_BLK_CASE(10) PACOH_BLK_CASE(11)
    PACOH_BLK_CASE(12) PACOH_BLK_CASE(13) PACOH_BLK_CASE(14) PACOH_BLK_CASE(15)
    default: break;
  }
#undef PACOH_BLK_CASE
}

// ---- building blocks of the blocked sweep (all register / warp-local) -----------------------------------------
// Publish the 4 block columns (= the 4 pivot rows, by symmetry) and the pivot rows' augmented entries.  The pivot
// block's own diagonal is published MINUS 1: with X = B0 - I in place of B0 the uniform rank-4 update also produces
// the swept values of the block columns themselves, so the register matrix never needs a dynamic-index write-back.
template <int NS>
__device__ __forceinline__ void publish_block(float* pb, float (&t0)[NS][4], const float (&dadd)[NS], const float (&aug)[NS],
                                              int lane, int c) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int rel = lane + 32 * s - 4 * c;       // position of this row inside the block (0..3) if it is a pivot row
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (rel == j) t0[s][j] += dadd[s];         // true diagonal (deferred unit-diagonal term)
      pb[j * kPubStride + lane + 32 * s] = rel == j ? t0[s][j] - 1.0f : t0[s][j];
    }
    if (rel >= 0 && rel < 4) pb[4 * kPubStride + rel] = aug[s];
  }
}

__device__ __forceinline__ void load_pivot_block(const float* pb, int c, float (&B)[4][4], float4& augB) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = lds4(pb + j * kPubStride + 4 * c);
    B[j][0] = v.x; B[j][1] = v.y; B[j][2] = v.z; B[j][3] = v.w;
    B[j][j] += 1.0f;
  }
  augB = lds4(pb + 4 * kPubStride);
}

// In-register symmetric sweep of the 4x4 pivot block: B <- -B^-1; its pivots are the Schur diagonals d_k = L_kk^2.
__device__ __forceinline__ void invert4(float (&B)[4][4], bool& ok, float& logdet2) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float d = B[q4][q4];
    ok = ok && (d > 1e-12f);
    logdet2 += lg2_approx(d);
    const float inv = rcp_newton(d);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i == q4) continue;
      const float f = B[i][q4] * inv;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j != q4) B[i][j] = fmaf(-f, B[q4][j], B[i][j]);
      B[i][q4] = f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j != q4) B[q4][j] *= inv;
    B[q4][q4] = -inv;
  }
}

// Row multipliers of the rank-4 update (Binv = -B): non-pivot rows w = t0 Binv; pivot rows w = e_rel - Binv[rel][:].
template <int NS>
__device__ __forceinline__ void multipliers(float (&w)[NS][4], const float (&t0)[NS][4], const float (&B)[4][4], int lane, int c) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int rel = lane + 32 * s - 4 * c;
    const bool inb = rel >= 0 && rel < 4;
    float uu[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uu[j] = -(t0[s][0] * B[0][j] + t0[s][1] * B[1][j] + t0[s][2] * B[2][j] + t0[s][3] * B[3][j]);
      if (inb) uu[j] = rel == j ? 1.0f : 0.0f;   // B0[rel][:] Binv = e_rel exactly (no cond(B0) rounding)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float uB = -(uu[0] * B[0][j] + uu[1] * B[1][j] + uu[2] * B[2][j] + uu[3] * B[3][j]);
      w[s][j] = inb ? uu[j] - uB : uu[j];
    }
  }
}

// A[s][col] -= sum_j w[s][j] * X[j][col] over the whole row (a pivot row's own diagonal register ends up at
// -Binv_rr + 2 - dadd: undone when the diagonal is read).
template <int NC, int NS>
__device__ __forceinline__ void rank4_update(float (&A)[NS][NC], const float (&w)[NS][4], const float* pb) {
#pragma unroll
  for (int c4 = 0; c4 < NC / 4; ++c4) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = lds4(pb + j * kPubStride + 4 * c4);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        A[s][4 * c4] = fmaf(-w[s][j], v[j].x, A[s][4 * c4]);
        A[s][4 * c4 + 1] = fmaf(-w[s][j], v[j].y, A[s][4 * c4 + 1]);
        A[s][4 * c4 + 2] = fmaf(-w[s][j], v[j].z, A[s][4 * c4 + 2]);
        A[s][4 * c4 + 3] = fmaf(-w[s][j], v[j].w, A[s][4 * c4 + 3]);
      }
    }
  }
}

template <int NC, int FT>
__global__ void __launch_bounds__(kGpWarps * 32, PACOH_GP_MINB(NC)) gp_mll_kernel(GpArgs a) {
  constexpr int NS = (NC + 31) / 32;
  constexpr int NB = NC / 4;                   // pivot blocks
  constexpr int RS = ((FT + 1 + 3) / 4) * 4;   // smem feature row: FT scaled features, then alpha
  __shared__ __align__(16) float s_pub[kGpWarps][2][kPubFloats];
  __shared__ __align__(16) float s_feat[kGpWarps][NC][RS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kGpWarps + warp;
  if (pair >= a.P * a.T) return;
  const int p = pair / a.T, t = pair - p * a.T;
  const int ns = a.n, F = a.F, Q = a.T * a.n;               // ns: row stride of the (padded) task arrays
  const int src = __ldg(a.task_idx + t);
  const int n = a.task_n != nullptr ? __ldg(a.task_n + src) : ns;   // ragged batches: this task's own number of points
  const float* th = a.theta + (size_t)p * a.D;
  float(*sf)[RS] = s_feat[warp];

  // ---- hyper-parameters (random_gp.py:69-73; MAP: GPR_meta_mll.py:54-55,218)
  const float kC = 0.84932180028801904272f;   // sqrt(0.5 * log2(e)):  exp(-d2/2) = 2^(-|kC (z_a-z_b)/l|^2)
  float inv_ls[FT], ls[FT];
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    ls[f] = f < F ? softplus_f(__ldg(th + a.off_ls + f)) : 1.0f;
    inv_ls[f] = f < F ? kC / ls[f] : 0.0f;
  }
  const float raw_noise = __ldg(th + a.off_noise);
  const float sig2 = a.noise_floor + softplus_f(raw_noise);
  const float raw_os = a.has_oscale ? __ldg(th + a.off_oscale) : 0.0f;
  const float osc = a.has_oscale ? softplus_f(raw_os) : 1.0f;
  const float cmean = a.mean_kind == PACOH_MEAN_CONSTANT ? __ldg(th + a.off_const_mean) : 0.0f;

  // ---- this lane's rows: residual and scaled features (padding rows sit "infinitely far" away => zero Gram rows)
  float r[NS], u[NS][FT];
  bool valid[NS];
  for (int i = lane; i < 2 * kPubFloats; i += 32) s_pub[warp][0][i] = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int row = lane + 32 * s;
    valid[s] = row < n;
    const size_t q = (size_t)p * Q + (size_t)t * ns + row;
    float m = cmean;
    if (a.mean != nullptr && valid[s]) m = __ldg(a.mean + q);
    r[s] = valid[s] ? __ldg(a.y + (size_t)src * ns + row) - m : 0.0f;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
      float z = 0.0f;
      if (valid[s] && f < F)
        z = a.feat != nullptr ? __ldg(a.feat + q * F + f) : __ldg(a.x + ((size_t)src * ns + row) * a.d + f);
      u[s][f] = valid[s] ? z * inv_ls[f] : kFar;
    }
    if (row < NC) {
#pragma unroll
      for (int f = 0; f < RS; ++f) sf[row][f] = f < FT ? u[s][f] : 0.0f;
    }
  }
  __syncwarp();

  float A[NS][NC], aug[NS];
  float tot = 0.0f, rho = 0.0f, logdet2 = 0.0f;
  int status = -1;
#pragma unroll 1
  for (int attempt = 0; attempt < 4 && status < 0; ++attempt) {
    const float jit = attempt == 0 ? 0.0f : (attempt == 1 ? 1e-6f : (attempt == 2 ? 1e-5f : 1e-4f));   // gpytorch psd_safe_cholesky
    tot = osc + sig2 + jit;
    rho = osc / tot;
    // ---- normalised Gram  A = Kt / tot; the "+ (1 - rho)" of the unit diagonal is added when a row's block is extracted
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      if (b < n) {
        float fb[RS];
#pragma unroll
        for (int c4 = 0; c4 < RS / 4; ++c4) {
          const float4 v = lds4(&sf[b][4 * c4]);
          fb[4 * c4] = v.x; fb[4 * c4 + 1] = v.y; fb[4 * c4 + 2] = v.z; fb[4 * c4 + 3] = v.w;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          float e = 0.0f;
#pragma unroll
          for (int f = 0; f < FT; ++f) {
            const float du = u[s][f] - fb[f];
            e = fmaf(-du, du, e);
          }
          A[s][b] = rho * ex2_approx(e);
        }
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) A[s][b] = 0.0f;
      }
    }
    float dadd[NS];   // deferred diagonal: real rows 1 - rho, identity padding rows (n <= row < NC) 1
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      aug[s] = r[s];
      dadd[s] = valid[s] ? 1.0f - rho : 1.0f;
    }

    // ---- blocked symmetric Gauss-Jordan sweep, 4 pivots per block, software pipelined:
    //      iteration c publishes + inverts block c+1 (a latency-bound dependent chain) in the same basic block as the
    //      416 independent FFMAs of block c's rank-4 update, so ptxas interleaves the two.
    bool ok = true;
    logdet2 = 0.0f;
    const int nblk = (n + 3) >> 2;
    float w[NS][4];
    float4 augB;
    {
      float t0[NS][4], B[4][4];
      block_columns<NC, NS>(A, t0, 0);
      publish_block<NS>(s_pub[warp][0], t0, dadd, aug, lane, 0);
      __syncwarp();
      load_pivot_block(s_pub[warp][0], 0, B, augB);
      invert4(B, ok, logdet2);
      multipliers<NS>(w, t0, B, lane, 0);
    }
#pragma unroll 1
    for (int c = 0; c < nblk && ok; ++c) {
      const float* pb = s_pub[warp][c & 1];
#pragma unroll
      for (int s = 0; s < NS; ++s) aug[s] -= w[s][0] * augB.x + w[s][1] * augB.y + w[s][2] * augB.z + w[s][3] * augB.w;
      if (c + 1 < nblk) {
        float* pbn = s_pub[warp][(c + 1) & 1];
        float tn[NS][4], B[4][4];
        block_columns<NC, NS>(A, tn, c + 1);
        // bring the next block's columns up to date with block c's update before everything else
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 x4 = lds4(pb + j * kPubStride + 4 * (c + 1));
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            tn[s][0] = fmaf(-w[s][j], x4.x, tn[s][0]); tn[s][1] = fmaf(-w[s][j], x4.y, tn[s][1]);
            tn[s][2] = fmaf(-w[s][j], x4.z, tn[s][2]); tn[s][3] = fmaf(-w[s][j], x4.w, tn[s][3]);
          }
        }
        publish_block<NS>(pbn, tn, dadd, aug, lane, c + 1);
        __syncwarp();
        float4 augBn;
        load_pivot_block(pbn, c + 1, B, augBn);
        invert4(B, ok, logdet2);                      // dependent chain ...
        rank4_update<NC, NS>(A, w, pb);               // ... overlapped with block c's independent FFMAs
        multipliers<NS>(w, tn, B, lane, c + 1);
        augB = augBn;
      } else {
        rank4_update<NC, NS>(A, w, pb);
      }
      __syncwarp();   // all lanes are done reading pb before the buffer is republished two blocks later
    }
    __syncwarp();
    if (ok) status = attempt;
  }

  float* mll_out = a.mll + (size_t)p * a.T + t;
  float* hyp = a.dhyp + ((size_t)p * a.T + t) * gp_hyp_stride(F);
  if (a.info != nullptr && lane == 0) a.info[(size_t)p * a.T + t] = status;
  if (status < 0) {   // reference: gpytorch raises NotPSDError; the host wrapper does the same from `info`
    if (lane == 0) {
      *mll_out = CUDART_NAN_F;
      for (int f = 0; f < F + 3; ++f) hyp[f] = 0.0f;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int row = lane + 32 * s;
      if (row < ns) {
        const size_t q = (size_t)p * Q + (size_t)t * ns + row;
        if (a.dmean != nullptr) a.dmean[q] = 0.0f;
        if (a.dfeat != nullptr)
          for (int f = 0; f < F; ++f) a.dfeat[q * F + f] = 0.0f;
      }
    }
    return;
  }

  // ---- value: alpha_hat = A_hat^-1 r lives in aug; alpha = alpha_hat / tot
  const float inv_tot = 1.0f / tot;
  const float inv_n = 1.0f / (float)n;
  float quad = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) quad = fmaf(r[s], aug[s], quad);
  quad = warp_sum(quad) * inv_tot;
  const float logdet = (float)n * logf(tot) + logdet2 * 0.69314718055994530942f;
  const float mll = (-0.5f * quad - 0.5f * logdet - 0.5f * (float)n * 1.83787706640934548356f) * inv_n;
  if (lane == 0) *mll_out = mll;

  // ---- diagonal of -Khat^-1 (register index == lane: predicated copies) and alpha broadcast rows
  float dg[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    dg[s] = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (32 * s + j < NC && lane == j) dg[s] = A[s][(32 * s + j) < NC ? (32 * s + j) : 0] - 1.0f - rho;   // - (2 - dadd)
    const int row = lane + 32 * s;
    if (row < NC) sf[row][FT] = aug[s];
  }
  __syncwarp();

  // ---- gradient contraction: w_ab = (beta_a alpha_hat_b - Khat^-1_ab) k_ab ; everything scaled at the end
  float beta[NS], S1[NS][FT], Sk = 0.0f, Str = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    beta[s] = aug[s] * inv_tot;
    if (valid[s]) Str += fmaf(beta[s], aug[s], dg[s]);
#pragma unroll
    for (int f = 0; f < FT; ++f) S1[s][f] = 0.0f;
  }
#pragma unroll
  for (int b = 0; b < NC; ++b) {
    if (b < n) {
      float fb[RS];
#pragma unroll
      for (int c4 = 0; c4 < RS / 4; ++c4) {
        const float4 v = lds4(&sf[b][4 * c4]);
        fb[4 * c4] = v.x; fb[4 * c4 + 1] = v.y; fb[4 * c4 + 2] = v.z; fb[4 * c4 + 3] = v.w;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float du[FT], e = 0.0f;
#pragma unroll
        for (int f = 0; f < FT; ++f) {
          du[f] = u[s][f] - fb[f];
          e = fmaf(-du[f], du[f], e);
        }
        const float wv = fmaf(beta[s], fb[FT], A[s][b]) * ex2_approx(e);
        Sk += wv;
#pragma unroll
        for (int f = 0; f < FT; ++f) S1[s][f] = fmaf(wv, du[f], S1[s][f]);
      }
    }
  }

  // ---- write-out.  G = g' / (2 tot);  W = rho/2 g' k;  all gradients are of mll = L / n.
  //      sum_ab w_ab du_ab^2 = 2 sum_a u_a S1_a   (w symmetric, du antisymmetric) gives the lengthscale gradient.
  // The stationary kernel is invariant to a common shift of the features: sum_a S1_a = 0 exactly.  Rounding leaves a
  // residue (w is not bit-symmetric); removing its mean projects the gradient back onto that invariance, so that the
  // kernel net's output-bias gradient (= sum over the points) cancels as far as the later fp32 sums allow.
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    float t = 0.0f;
#pragma unroll
    for (int s = 0; s < NS; ++s) t += (lane + 32 * s < n) ? S1[s][f] : 0.0f;
    t = warp_sum(t) * inv_n;
#pragma unroll
    for (int s = 0; s < NS; ++s) S1[s][f] -= t;
  }
  float S2[FT];
#pragma unroll
  for (int f = 0; f < FT; ++f) S2[f] = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int row = lane + 32 * s;
    if (row >= n && row < ns) {   // padding rows of a ragged batch: the MLP backward must see exact zeros
      const size_t q = (size_t)p * Q + (size_t)t * ns + row;
      if (a.dmean != nullptr) a.dmean[q] = 0.0f;
      if (a.dfeat != nullptr)
        for (int f = 0; f < F; ++f) a.dfeat[q * F + f] = 0.0f;
    }
    if (row < n) {
      const size_t q = (size_t)p * Q + (size_t)t * ns + row;
      if (a.dmean != nullptr) a.dmean[q] = beta[s] * inv_n;
#pragma unroll
      for (int f = 0; f < FT; ++f) {
        if (f < F) {
          // sum_a S1_a = 0, so any common offset may be removed from u_a: centre on row 0 against cancellation
          S2[f] = fmaf(2.0f * (u[s][f] - sf[0][f]), S1[s][f], S2[f]);
          if (a.dfeat != nullptr) a.dfeat[q * F + f] = -rho * inv_n * inv_ls[f] / (kC * kC) * S1[s][f];
        }
      }
    }
  }
  float dmean_sum = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) dmean_sum += beta[s];
  dmean_sum = warp_sum(dmean_sum) * inv_n;
  Sk = warp_sum(Sk) - (float)n * (1.0f + rho);   // the diagonal registers carry +2 - dadd = 1 + rho  (k_aa = 1)
  Str = warp_sum(Str);
#pragma unroll
  for (int f = 0; f < FT; ++f) S2[f] = warp_sum(S2[f]);
  if (lane == 0) {
#pragma unroll
    for (int f = 0; f < FT; ++f)
      if (f < F) {
        // dL/dl_f = rho/(2 n) * S2 / (kC^2 l_f) ; chain through softplus
        hyp[f] = 0.5f * rho * inv_n * S2[f] / (kC * kC * ls[f]) * sigmoid_f(__ldg(th + a.off_ls + f));
      }
    hyp[F] = 0.5f * inv_tot * inv_n * Str * sigmoid_f(raw_noise);
    hyp[F + 1] = a.has_oscale ? 0.5f * inv_tot * inv_n * Sk * sigmoid_f(raw_os) : 0.0f;
    hyp[F + 2] = dmean_sum;
  }
}

template <int NC, int FT>
int launch_one(const GpArgs& a, cudaStream_t st) {
  const int pairs = a.P * a.T;
  const int blocks = (pairs + kGpWarps - 1) / kGpWarps;
  gp_mll_kernel<NC, FT><<<blocks, kGpWarps * 32, 0, st>>>(a);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

template <int FT>
int dispatch_nc16(const GpArgs& a, cudaStream_t st) {
  const int nc = (a.n + 15) / 16 * 16;
  switch (nc) {
    case 16: return launch_one<16, FT>(a, st);
    case 32: return launch_one<32, FT>(a, st);
    case 48: return launch_one<48, FT>(a, st);
    case 64: return launch_one<64, FT>(a, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

int dispatch_nc4_f2(const GpArgs& a, cudaStream_t st) {
  const int nc = (a.n + 3) / 4 * 4;
  switch (nc) {
#define PACOH_NC_CASE(N) case N: return launch_one<N, 2>(a, st);
    PACOH_NC_CASE(4) PACOH_NC_CASE(8) PACOH_NC_CASE(12) PACOH_NC_CASE(16) PACOH_NC_CASE(20) PACOH_NC_CASE(24)
    PACOH_NC_CASE(28) PACOH_NC_CASE(32) PACOH_NC_CASE(36) PACOH_NC_CASE(40) PACOH_NC_CASE(44) PACOH_NC_CASE(48)
    PACOH_NC_CASE(52) PACOH_NC_CASE(56) PACOH_NC_CASE(60) PACOH_NC_CASE(64)
#undef PACOH_NC_CASE
  }
  return PACOH_ERR_UNSUPPORTED;
}

}  // namespace

int launch_gp_mll(const GpArgs& a, cudaStream_t st) {
  if (a.n < 1 || a.n > kMaxGpN || a.F < 1 || a.F > kMaxGpF) return PACOH_ERR_UNSUPPORTED;
  // 32 < n <= 64: the tensor-memory / tcgen05 kernel (gp_tc.cu) unless PACOH_GP=warp selects the register kernel
  // (kept for A/B measurements; both are parity-tested).
  static int tc = -1;
  if (tc < 0) {
    const char* e = getenv("PACOH_GP");
    tc = (e != nullptr && strcmp(e, "warp") == 0) ? 0 : 1;
  }
  if ((tc == 1 || a.n > 64) && a.n > 32 && a.F <= 4) return launch_gp_mll_tc(a, st);
  if (a.n > 64) return PACOH_ERR_UNSUPPORTED;          // the register kernel holds at most 64 columns per lane pair
  if (a.F <= 2) return dispatch_nc4_f2(a, st);
  if (a.F <= 4) return dispatch_nc16<4>(a, st);
  return dispatch_nc16<16>(a, st);
}

}  // namespace pacoh
