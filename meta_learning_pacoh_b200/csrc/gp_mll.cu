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
// Data layout.  Lane i of the warp owns rows i and i+32 of the (unit-diagonal-normalised) matrix in REGISTERS
// (NC columns each) plus the augmented column r.  The factorisation is a symmetric Gauss-Jordan sweep: at
// step k the owner lane publishes pivot row k to shared memory, every lane reads its multiplier from that
// row (symmetry) and applies the rank-1 update with broadcast LDS.128 operands -- one FFMA per matrix
// entry per step, no dynamic register indexing, no bank conflicts, one __syncwarp per step.  The pivots are
// the squared Cholesky diagonal (d_k = L_kk^2), so log det and the positive-definiteness test are the
// Cholesky ones; after n sweeps the registers hold -Kt^-1 and the augmented column holds Kt^-1 r.
// The Gram matrix is never stored: exp(-d2) is recomputed (MUFU.EX2) for the gradient contraction.
#include <math_constants.h>
#include "common.cuh"
#include "kernels.cuh"

namespace pacoh {

namespace {

constexpr int kGpWarps = 4;
constexpr int kRowBuf = 72;   // 64 columns + [64] augmented entry + [65] pivot

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

template <int NC, int FT>
__global__ void __launch_bounds__(kGpWarps * 32) gp_mll_kernel(GpArgs a) {
  constexpr int NS = (NC + 31) / 32;
  constexpr int RS = ((FT + 1 + 3) / 4) * 4;   // smem feature row: FT scaled features, then alpha
  __shared__ __align__(16) float s_row[kGpWarps][2][kRowBuf];
  __shared__ __align__(16) float s_feat[kGpWarps][NC][RS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * kGpWarps + warp;
  if (pair >= a.P * a.T) return;
  const int p = pair / a.T, t = pair - p * a.T;
  const int n = a.n, F = a.F, Q = a.T * a.n;
  const int src = __ldg(a.task_idx + t);
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

  // ---- this lane's rows: residual and scaled features
  float r[NS], u[NS][FT];
  bool valid[NS];
  for (int i = lane; i < 2 * kRowBuf; i += 32) s_row[warp][0][i] = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int row = lane + 32 * s;
    valid[s] = row < n;
    const size_t q = (size_t)p * Q + (size_t)t * n + row;
    float m = cmean;
    if (a.mean != nullptr && valid[s]) m = __ldg(a.mean + q);
    r[s] = valid[s] ? __ldg(a.y + (size_t)src * n + row) - m : 0.0f;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
      float z = 0.0f;
      if (valid[s] && f < F)
        z = a.feat != nullptr ? __ldg(a.feat + q * F + f) : __ldg(a.x + ((size_t)src * n + row) * a.d + f);
      u[s][f] = z * inv_ls[f];
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
  const float jitter[4] = {0.0f, 1e-6f, 1e-5f, 1e-4f};   // gpytorch psd_safe_cholesky ladder (fp32)
  for (int attempt = 0; attempt < 4 && status < 0; ++attempt) {
    tot = osc + sig2 + jitter[attempt];
    rho = osc / tot;
    // ---- normalised Gram  A = Kt / tot  (unit diagonal), augmented with r
#pragma unroll
    for (int b = 0; b < NC; ++b) {
      if (b < n) {
        float fb[RS];
#pragma unroll
        for (int c = 0; c < RS / 4; ++c) {
          const float4 v = lds4(&sf[b][4 * c]);
          fb[4 * c] = v.x; fb[4 * c + 1] = v.y; fb[4 * c + 2] = v.z; fb[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          float e = 0.0f;
#pragma unroll
          for (int f = 0; f < FT; ++f) {
            const float du = u[s][f] - fb[f];
            e = fmaf(-du, du, e);
          }
          const float kv = rho * ex2_approx(e);
          A[s][b] = valid[s] ? ((lane + 32 * s == b) ? 1.0f : kv) : 0.0f;
        }
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) A[s][b] = 0.0f;
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) aug[s] = r[s];

    // ---- symmetric Gauss-Jordan sweep over the n pivots
    bool ok = true;
    logdet2 = 0.0f;
    for (int k = 0; k < n; ++k) {
      float* rb = s_row[warp][k & 1];
      if (lane == (k & 31)) {
        if (NS == 1 || k < 32) {
#pragma unroll
          for (int c = 0; c < NC / 4; ++c) sts4(rb + 4 * c, make_float4(A[0][4 * c], A[0][4 * c + 1], A[0][4 * c + 2], A[0][4 * c + 3]));
          rb[64] = aug[0];
        } else {
#pragma unroll
          for (int c = 0; c < NC / 4; ++c)
            sts4(rb + 4 * c, make_float4(A[NS - 1][4 * c], A[NS - 1][4 * c + 1], A[NS - 1][4 * c + 2], A[NS - 1][4 * c + 3]));
          rb[64] = aug[NS - 1];
        }
        const float dk = rb[k];
        rb[65] = dk;
        rb[k] = dk - 1.0f;   // makes the uniform update below produce A_ik <- A_ik / d  (and the pivot row / d)
      }
      __syncwarp();
      const float dk = rb[65];
      if (!(dk > 1e-12f)) { ok = false; break; }   // warp-uniform: not positive definite at this jitter level
      logdet2 += lg2_approx(dk);
      const float inv = rcp_newton(dk);
      float fm[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) fm[s] = -rb[lane + 32 * s] * inv;
#pragma unroll
      for (int c = 0; c < NC / 4; ++c) {
        const float4 v = lds4(rb + 4 * c);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          A[s][4 * c] = fmaf(fm[s], v.x, A[s][4 * c]);
          A[s][4 * c + 1] = fmaf(fm[s], v.y, A[s][4 * c + 1]);
          A[s][4 * c + 2] = fmaf(fm[s], v.z, A[s][4 * c + 2]);
          A[s][4 * c + 3] = fmaf(fm[s], v.w, A[s][4 * c + 3]);
        }
      }
      const float va = rb[64];
#pragma unroll
      for (int s = 0; s < NS; ++s) aug[s] = fmaf(fm[s], va, aug[s]);
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
      if (row < n) {
        const size_t q = (size_t)p * Q + (size_t)t * n + row;
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

  // ---- diagonal of the swept matrix (carries a +2 offset) and alpha broadcast rows
  float dg[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    dg[s] = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (32 * s + j < NC && lane == j) dg[s] = A[s][(32 * s + j) < NC ? (32 * s + j) : 0];
    const int row = lane + 32 * s;
    if (row < NC) sf[row][FT] = aug[s];
  }
  __syncwarp();

  // ---- gradient contraction: w_ab = (beta_a alpha_hat_b - Khat^-1_ab) k_ab ; everything scaled at the end
  float beta[NS], S1[NS][FT], S2[FT], Sk = 0.0f, Str = 0.0f;
#pragma unroll
  for (int f = 0; f < FT; ++f) S2[f] = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    beta[s] = aug[s] * inv_tot;
    if (valid[s]) Str += fmaf(beta[s], aug[s], dg[s]) - 2.0f;
#pragma unroll
    for (int f = 0; f < FT; ++f) S1[s][f] = 0.0f;
  }
#pragma unroll
  for (int b = 0; b < NC; ++b) {
    if (b < n) {
      float fb[RS];
#pragma unroll
      for (int c = 0; c < RS / 4; ++c) {
        const float4 v = lds4(&sf[b][4 * c]);
        fb[4 * c] = v.x; fb[4 * c + 1] = v.y; fb[4 * c + 2] = v.z; fb[4 * c + 3] = v.w;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        float du[FT], e = 0.0f;
#pragma unroll
        for (int f = 0; f < FT; ++f) {
          du[f] = u[s][f] - fb[f];
          e = fmaf(-du[f], du[f], e);
        }
        const float w = fmaf(beta[s], fb[FT], A[s][b]) * ex2_approx(e);
        Sk += w;
#pragma unroll
        for (int f = 0; f < FT; ++f) {
          const float tdu = w * du[f];
          S1[s][f] += tdu;
          S2[f] = fmaf(tdu, du[f], S2[f]);
        }
      }
    }
  }

  // ---- write-out.  G = g' / (2 tot);  W = rho/2 g' k;  all gradients are of mll = L / n.
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int row = lane + 32 * s;
    if (row < n) {
      const size_t q = (size_t)p * Q + (size_t)t * n + row;
      if (a.dmean != nullptr) a.dmean[q] = beta[s] * inv_n;
      if (a.dfeat != nullptr) {
#pragma unroll
        for (int f = 0; f < FT; ++f)
          if (f < F) a.dfeat[q * F + f] = -rho * inv_n * inv_ls[f] / (kC * kC) * S1[s][f];
      }
    }
  }
  // rows beyond n contribute exact zeros to the sums (beta = 0, A row = 0)
  float dmean_sum = 0.0f;
#pragma unroll
  for (int s = 0; s < NS; ++s) dmean_sum += beta[s];
  dmean_sum = warp_sum(dmean_sum) * inv_n;
  Sk = warp_sum(Sk) - 2.0f * (float)n;     // remove the +2 diagonal offset (k_aa = 1)
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
  if (a.F <= 2) return dispatch_nc4_f2(a, st);
  if (a.F <= 4) return dispatch_nc16<4>(a, st);
  return dispatch_nc16<16>(a, st);
}

}  // namespace pacoh
