// Batched GP marginal log-likelihood with analytic gradient -- tensor-memory / tcgen05 version for 32 < n <= 128, sm_100a.
//
// Same mathematics and outputs as gp_mll_kernel (gp_mll.cu; reference: meta_learn/random_gp.py:54-89, models.py:428-487,
// gpytorch ExactMarginalLogLikelihood at random_gp.py:83-85 and its autograd reverse pass, svgd.py:16), different mapping:
//
//   CTA = 128 threads = TWO (particle, task) matrices for n <= 64 (ONE for 64 < n <= 128: same code, MT = 1);
//   thread (m = tid >> 6, row = tid & 63) owns one matrix row.
//   The normalised matrix lives in TENSOR MEMORY: row r of matrix m is TMEM lane 64 m + r, columns 0..63 (fp32).
//   Blocked symmetric Gauss-Jordan, 4 pivots per block.  Per block
//     1. every thread reads its 4 block-column entries with tcgen05.ld at a RUNTIME column offset (tensor memory is
//        addressed, registers are not: no register-indexing switch), adds the deferred diagonal term and publishes them
//        as one 16-byte chunk of the UMMA B operand X[64 x 8] (K-slots 0-3: matrix 0, 4-7: matrix 1; hi / lo split);
//     2. after a CTA barrier every thread inverts its matrix's 4x4 pivot block in registers (pivots = squared
//        Cholesky diagonal: same log det / positive-definiteness test), forms its row multipliers w and publishes -w
//        as its row of the UMMA A operand W[128 x 8] (zeros in the other matrix's K-slots);
//     3. ONE thread issues the rank-4 update of BOTH matrices, D[128 x 64] += W X^T, as 3 tcgen05.mma (3xTF32,
//        fp32 accumulate in place in TMEM) -- the 2 x 416 FFMAs per block of the register version disappear.
//   The `X - I` trick (pivot block diagonal published minus 1) makes the same uniform update produce the swept block
//   columns; a pivot row's own diagonal ends up offset by 2 - dadd, undone when the diagonal is read.
#include <math_constants.h>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace pacoh {

namespace {

using namespace tc;

constexpr int kGT = 128;           // threads per CTA
constexpr float kFar = 1.0e18f;    // scaled feature of a padding row: exp2(-(1e18)^2) == 0 exactly
#ifndef PACOH_GPTC_MINB
#define PACOH_GPTC_MINB 8   // CTAs per SM the register allocation is capped for (tensor memory allows 8)
#endif

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpn(float d) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d)); return r * fmaf(-d, r, 2.0f); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&f)[8]) {
  uint32_t v[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);
}

// hi part of the 3xTF32 split: truncation (one LOP; the tensor core would truncate anyway).  Round-to-nearest splitting
// was tried for the one-matrix-per-CTA variant (n <= 128) and did not improve its accuracy (the error there comes from the
// in-place Gauss-Jordan sweep itself), so both variants share this.
template <int MT>
__device__ __forceinline__ float split_hi(float v) {
  return tf32_hi(v);
}

// In-register symmetric sweep of the 4x4 pivot block: B <- -B^-1 (only the upper triangle B[i][j], i <= j, is read or
// written; the swept matrix stays symmetric); its pivots are the Schur diagonals d_k = L_kk^2.
#define PACOH_SYM(i, j) B[(i) < (j) ? (i) : (j)][(i) < (j) ? (j) : (i)]
__device__ __forceinline__ void invert4_sym(float (&B)[4][4], bool& ok, float& logdet2) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float d = B[q4][q4];
    ok = ok && (d > 1e-12f);
    logdet2 += lg2a(d);
    const float inv = rcpn(d);
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = i == q4 ? 0.0f : PACOH_SYM(i, q4) * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j)
        if (i != q4 && j != q4) B[i][j] = fmaf(-f[i], PACOH_SYM(q4, j), B[i][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i != q4) PACOH_SYM(i, q4) = f[i];
    B[q4][q4] = -inv;
  }
}

// MT = matrices per CTA: 2 (n <= 64: 64 rows / TMEM columns each, 8 CTAs per SM) or 1 (64 < n <= 128: 128 rows and
// columns, 4 CTAs per SM; the K-slots 4-7 of both MMA operands are a constant zero plane).
// RAG: ragged batches (per-task n from a.task_n); the dense instantiation keeps n a launch constant.
template <int FT, int MT, bool RAG>
__global__ void __launch_bounds__(kGT, (MT == 2 ? PACOH_GPTC_MINB : 4)) gp_tc_kernel(GpArgs a) {
  constexpr int RS = ((FT + 1 + 3) / 4) * 4;   // smem feature row: FT scaled features, then alpha
  constexpr int kRows = kGT / MT;              // rows (= TMEM columns) per matrix
  constexpr int WPM = 4 / MT;                  // warps per matrix
  constexpr int kCols = kRows;                 // tensor-memory columns of the CTA
  __shared__ __align__(1024) float s_x_hi[kRows * 8], s_x_lo[kRows * 8];     // UMMA B operand X [kRows x 8], K-major
  __shared__ __align__(1024) float s_w_hi[kGT * 8], s_w_lo[kGT * 8];         // UMMA A operand W [128 x 8], K-major
  __shared__ __align__(16) float s_t0[MT][kRows][4];                          // published block columns (true values)
  __shared__ __align__(16) float s_aug[MT][4];
  __shared__ __align__(16) float s_feat[MT][kRows][RS];
  __shared__ float s_red[MT][WPM][8];                                        // [matrix][warp-in-matrix][slot]
  __shared__ float s_hyp[MT][8][2];                                           // hyper-parameters, one per designated thread
  __shared__ float s_proj[MT][WPM][4];                                        // per-warp sums of S1 (shift-invariance projection)
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtx = tid / kRows, row = tid % kRows, wim = warp % WPM;
  const int kplane = MT == 2 ? mtx : 0;              // K-slot plane (4 slots) this matrix owns in the MMA operands
  const int p = blockIdx.y;                          // particle
  const bool pvalid = blockIdx.x * MT + mtx < a.T;   // odd T: the last CTA's second matrix is a dummy
  const int t = pvalid ? blockIdx.x * MT + mtx : a.T - 1;
  const int ns = a.n, F = a.F, Q = a.T * a.n;               // ns: row stride of the (padded) task arrays
  const int src = __ldg(a.task_idx + t);
  // ragged batches: every matrix has its own number of points n <= ns; the CTA-uniform loops run to the larger of the two
  int n = ns, nloop = ns;
  if (RAG) {
    n = __ldg(a.task_n + src);
    nloop = n;
    if (MT == 2) {
      const int t_other = min((int)blockIdx.x * 2 + (1 - mtx), a.T - 1);
      nloop = max(n, __ldg(a.task_n + __ldg(a.task_idx + t_other)));
    }
  }
  const float* th = a.theta + (size_t)p * a.D;
  float(*sf)[RS] = s_feat[mtx];

  if (warp == 0) tmem_alloc<kCols>(&tmem_base_s);
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  // the K-slots of the OTHER matrix are zero in this row of W, for the whole kernel
  sts4(s_w_hi + ((tid + (1 - kplane) * kGT) << 2), make_float4(0.f, 0.f, 0.f, 0.f));
  sts4(s_w_lo + ((tid + (1 - kplane) * kGT) << 2), make_float4(0.f, 0.f, 0.f, 0.f));
  if (MT == 1) {   // one matrix: the second K plane of X is zero as well
    sts4(s_x_hi + ((row + kRows) << 2), make_float4(0.f, 0.f, 0.f, 0.f));
    sts4(s_x_lo + ((row + kRows) << 2), make_float4(0.f, 0.f, 0.f, 0.f));
  }

  // ---- hyper-parameters (random_gp.py:69-73; MAP: GPR_meta_mll.py:54-55,218)
  //      computed once per matrix: rows 0..FT-1 take a lengthscale each, row 4 the noise, row 5 the output scale;
  //      slot [.][0] = the value the sweep needs, slot [.][1] = the softplus chain factor of its gradient
  const float kC = 0.84932180028801904272f;   // sqrt(0.5 * log2(e))
  if (row < 6) {
    float v0 = 0.0f, v1 = 0.0f;
    if (row < FT) {
      if (row < F) {
        const float raw = __ldg(th + a.off_ls + row), ls = softplus_f(raw);
        v0 = kC / ls;                                  // scaled inverse lengthscale
        v1 = 0.5f * sigmoid_f(raw) / (kC * kC * ls);   // dL/dl_f = rho/(2 n) * S2 / (kC^2 l_f) ; chain through softplus
      }
    } else if (row == 4) {
      const float raw = __ldg(th + a.off_noise);
      v0 = a.noise_floor + softplus_f(raw);
      v1 = sigmoid_f(raw);
    } else if (row == 5) {
      const float raw = a.has_oscale ? __ldg(th + a.off_oscale) : 0.0f;
      v0 = a.has_oscale ? softplus_f(raw) : 1.0f;
      v1 = a.has_oscale ? sigmoid_f(raw) : 0.0f;
    }
    s_hyp[mtx][row][0] = v0;
    s_hyp[mtx][row][1] = v1;
  }

  // ---- this thread's row: residual and scaled features (padding rows sit "infinitely far" away => zero Gram rows)
  const bool valid = row < n;
  const size_t q = (size_t)p * Q + (size_t)t * ns + row;
  float r = 0.0f, u[FT];
  {
    float m = a.mean_kind == PACOH_MEAN_CONSTANT ? __ldg(th + a.off_const_mean) : 0.0f;
    if (a.mean != nullptr && valid) m = __ldg(a.mean + q);
    r = valid ? __ldg(a.y + (size_t)src * ns + row) - m : 0.0f;
#pragma unroll
    for (int f = 0; f < FT; ++f) {
      u[f] = 0.0f;
      if (valid && f < F) u[f] = a.feat != nullptr ? __ldg(a.feat + q * F + f) : __ldg(a.x + ((size_t)src * ns + row) * a.d + f);
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const float sig2 = s_hyp[mtx][4][0], osc = s_hyp[mtx][5][0];
#pragma unroll
  for (int f = 0; f < FT; ++f) u[f] = valid ? u[f] * s_hyp[mtx][f][0] : kFar;
#pragma unroll
  for (int f = 0; f < RS; ++f) sf[row][f] = f < FT ? u[f] : 0.0f;
  __syncthreads();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);   // provably warp-uniform: MMA operands stay in uniform registers
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t bar = smem_u32(&mbar);
  uint32_t parity = 0;
  const int nblk = (nloop + 3) >> 2;   // 4-pivot blocks
  const int ngrp = (nloop + 7) >> 3;   // 8-column groups of the row that are ever read
  const uint32_t idesc = umma_idesc_tf32(128, ((nloop + 15) >> 4) << 4);

  float aug = 0.0f, tot = 0.0f, rho = 0.0f, logdet2 = 0.0f, dadd = 0.0f;
  int lvl = 0, status = -1;     // jitter level of this thread's matrix (uniform within the matrix)
#pragma unroll 1
  for (int attempt = 0; attempt < 4; ++attempt) {
    const float jit = lvl == 0 ? 0.0f : (lvl == 1 ? 1e-6f : (lvl == 2 ? 1e-5f : 1e-4f));   // gpytorch psd_safe_cholesky ladder
    tot = osc + sig2 + jit;
    rho = osc / tot;
    // ---- normalised Gram row -> tensor memory, 8 columns at a time (the "+ (1 - rho)" of the unit diagonal is added
    //      when the row's pivot block is read).  rho k = 2^(log2 rho - |du|^2); padding rows / columns give exactly 0.
    const float e0 = valid ? lg2a(rho) : -CUDART_INF_F;
#pragma unroll 1
    for (int g8 = 0; g8 < ngrp; ++g8) {
      uint32_t g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 fb = lds4(&sf[8 * g8 + j][0]);
        float e = e0;
        { const float du = u[0] - fb.x; e = fmaf(-du, du, e); }
        if (FT > 1) { const float du = u[1] - fb.y; e = fmaf(-du, du, e); }
        if (FT > 2) { const float du = u[2] - fb.z; e = fmaf(-du, du, e); }
        if (FT > 3) { const float du = u[3] - fb.w; e = fmaf(-du, du, e); }
        g[j] = __float_as_uint(ex2a(e));
      }
      tmem_st8(lane_base + 8 * g8, g);
    }
    tmem_st_wait();
    aug = r;
    dadd = valid ? 1.0f - rho : 1.0f;
    bool ok = true;
    logdet2 = 0.0f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // ---- blocked symmetric Gauss-Jordan sweep, 4 pivots per block; the rank-4 update runs on the tensor cores
#pragma unroll 1
    for (int c = 0; c < nblk; ++c) {
      if (c > 0) {
        mbar_wait_hint(bar, parity, 2000u);
        parity ^= 1;
        fence_after_sync();
      }
      float t0[4];
      tmem_ld4(lane_base + 4 * c, t0);
      const int rel = row - 4 * c;
      const bool inb = (unsigned)rel < 4u;
      {
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (rel == j) t0[j] += dadd;                 // true diagonal
          xv[j] = rel == j ? t0[j] - 1.0f : t0[j];     // X = T0 - I on the pivot block diagonal
        }
        *reinterpret_cast<float4*>(&s_t0[mtx][row][0]) = make_float4(t0[0], t0[1], t0[2], t0[3]);
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { hi[j] = split_hi<MT>(xv[j]); lo[j] = xv[j] - hi[j]; }
        sts4(s_x_hi + ((row + kplane * kRows) << 2), make_float4(hi[0], hi[1], hi[2], hi[3]));
        sts4(s_x_lo + ((row + kplane * kRows) << 2), make_float4(lo[0], lo[1], lo[2], lo[3]));
        if (inb) s_aug[mtx][rel] = aug;
      }
      __syncthreads();
      float B[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(&s_t0[mtx][4 * c + i][0]);
        B[i][0] = v.x; B[i][1] = v.y; B[i][2] = v.z; B[i][3] = v.w;
      }
      const float4 augB = *reinterpret_cast<const float4*>(&s_aug[mtx][0]);
      invert4_sym(B, ok, logdet2);
      // negated multipliers nw = -w (Binv = -B): non-pivot rows w = t0 Binv, i.e. nw = t0 B
      float nw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        nw[j] = fmaf(t0[3], PACOH_SYM(3, j), fmaf(t0[2], PACOH_SYM(2, j), fmaf(t0[1], PACOH_SYM(1, j), t0[0] * PACOH_SYM(0, j))));
      if (inb) {   // pivot rows: B0[rel][:] Binv = e_rel exactly (no cond(B0) rounding), w = e_rel - Binv[rel][:]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float b01 = (rel & 1) ? PACOH_SYM(1, j) : PACOH_SYM(0, j), b23 = (rel & 1) ? PACOH_SYM(3, j) : PACOH_SYM(2, j);
          const float brj = (rel & 2) ? b23 : b01;
          nw[j] = (rel == j ? -1.0f : 0.0f) - brj;
        }
      }
      aug = fmaf(nw[0], augB.x, fmaf(nw[1], augB.y, fmaf(nw[2], augB.z, fmaf(nw[3], augB.w, aug))));
      {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { hi[j] = split_hi<MT>(nw[j]); lo[j] = nw[j] - hi[j]; }
        sts4(s_w_hi + ((tid + kplane * kGT) << 2), make_float4(hi[0], hi[1], hi[2], hi[3]));
        sts4(s_w_lo + ((tid + kplane * kGT) << 2), make_float4(lo[0], lo[1], lo[2], lo[3]));
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint64_t dwh = umma_desc(smem_u32(s_w_hi), kGT * 16, 128), dwl = umma_desc(smem_u32(s_w_lo), kGT * 16, 128);
        const uint64_t dxh = umma_desc(smem_u32(s_x_hi), kRows * 16, 128), dxl = umma_desc(smem_u32(s_x_lo), kRows * 16, 128);
        // D[128 x N] += W X^T : lo*hi, hi*lo, hi*hi
        umma_tf32(tmem, dwl, dxh, idesc, 1);
        umma_tf32(tmem, dwh, dxl, idesc, 1);
        umma_tf32(tmem, dwh, dxh, idesc, 1);
        umma_commit(bar);
      }
    }
    mbar_wait_hint(bar, parity, 2000u);
    parity ^= 1;
    fence_after_sync();
    if (ok && status < 0) status = lvl;
    const bool retry = !ok && lvl < 3 && pvalid;
    if (retry) ++lvl;
    if (!__syncthreads_or(retry ? 1 : 0)) break;     // both matrices done (or out of jitter levels)
  }
  if (!pvalid) status = 0;
  const bool failed = status < 0;

  // ---- gradient contraction over the row of -Khat^-1, read back from tensor memory 8 columns at a time:
  //      w_ab = (beta_a alpha_hat_b - Khat^-1_ab) k_ab ; everything scaled at the end
  sf[row][FT] = aug;
  const float inv_tot = 1.0f / tot;
  const float inv_n = 1.0f / (float)n;
  const float beta = aug * inv_tot;
  __syncthreads();
  float S1[FT], Sk = 0.0f, dg = 0.0f;
#pragma unroll
  for (int f = 0; f < FT; ++f) S1[f] = 0.0f;
#pragma unroll 1
  for (int g8 = 0; g8 < ngrp; ++g8) {
    float A[8];
    tmem_ld8(lane_base + 8 * g8, A);
    if ((row >> 3) == g8) {   // this row's diagonal entry: a select tree on the low 3 bits of the row index
      const bool b0 = row & 1, b1 = row & 2, b2 = row & 4;
      const float s0 = b0 ? A[1] : A[0], s1 = b0 ? A[3] : A[2], s2 = b0 ? A[5] : A[4], s3 = b0 ? A[7] : A[6];
      const float t0 = b1 ? s1 : s0, t1 = b1 ? s3 : s2;
      dg = b2 ? t1 : t0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 fb = lds4(&sf[8 * g8 + j][0]);
      const float fbv[4] = {fb.x, fb.y, fb.z, fb.w};
      float du[FT], e = 0.0f;
#pragma unroll
      for (int f = 0; f < FT; ++f) {
        du[f] = u[f] - fbv[f];
        e = fmaf(-du[f], du[f], e);
      }
      float alpha_b;
      if (FT < 4) alpha_b = fbv[FT < 4 ? FT : 0];
      else alpha_b = sf[8 * g8 + j][FT];
      const float wv = fmaf(beta, alpha_b, A[j]) * ex2a(e);
      Sk += wv;
#pragma unroll
      for (int f = 0; f < FT; ++f) S1[f] = fmaf(wv, du[f], S1[f]);
    }
  }
  // ---- per-matrix sums over its two warps: quad, Sk, Str, dmean, S2[f]
  //      sum_ab w_ab du_ab^2 = 2 sum_a u_a S1_a (w symmetric, du antisymmetric); sum_a S1_a = 0, so centre u on row 0
  float red[4 + FT];
  red[0] = warp_sum(r * aug);
  red[1] = warp_sum(valid ? Sk : 0.0f);
  red[2] = warp_sum(valid ? fmaf(beta, aug, dg - 1.0f - rho) : 0.0f);     // diagonal entries carry + (2 - dadd) = 1 + rho
  red[3] = warp_sum(valid ? beta : 0.0f);
#pragma unroll
  for (int f = 0; f < FT; ++f) red[4 + f] = warp_sum(valid ? 2.0f * (u[f] - sf[0][f]) * S1[f] : 0.0f);
  // sum_a S1_a = 0 exactly (the kernel is invariant to a common feature shift); its rounding residue is removed below
  float prj[FT];
#pragma unroll
  for (int f = 0; f < FT; ++f) prj[f] = warp_sum(valid ? S1[f] : 0.0f);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4 + FT; ++i) s_red[mtx][wim][i] = red[i];
#pragma unroll
    for (int f = 0; f < FT; ++f) s_proj[mtx][wim][f] = prj[f];
  }
  __syncthreads();
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < WPM; ++w) t += s_proj[mtx][w][f];
    S1[f] -= t * inv_n;
  }

  if (pvalid) {
    float* hyp = a.dhyp + ((size_t)p * a.T + t) * gp_hyp_stride(F);
    if (!valid && row < ns) {   // padding rows of a ragged batch: the MLP backward must see exact zeros
      if (a.dmean != nullptr) a.dmean[q] = 0.0f;
      if (a.dfeat != nullptr)
        for (int f = 0; f < F; ++f) a.dfeat[q * F + f] = 0.0f;
    }
    if (valid) {
      if (a.dmean != nullptr) a.dmean[q] = failed ? 0.0f : beta * inv_n;
      if (a.dfeat != nullptr) {
#pragma unroll
        for (int f = 0; f < FT; ++f)
          if (f < F) a.dfeat[q * F + f] = failed ? 0.0f : -rho * inv_n * s_hyp[mtx][f][0] / (kC * kC) * S1[f];
      }
    }
    if (row == 0) {
      float* mll_out = a.mll + (size_t)p * a.T + t;
      if (a.info != nullptr) a.info[(size_t)p * a.T + t] = status;
      if (failed) {   // reference: gpytorch raises NotPSDError; the host wrapper does the same from `info`
        *mll_out = CUDART_NAN_F;
        for (int f = 0; f < F + 3; ++f) hyp[f] = 0.0f;
      } else {
        float tot_red[4 + FT];
#pragma unroll
        for (int i = 0; i < 4 + FT; ++i) {
          tot_red[i] = 0.0f;
#pragma unroll
          for (int w = 0; w < WPM; ++w) tot_red[i] += s_red[mtx][w][i];
        }
        const float quad = tot_red[0] * inv_tot;
        const float Skt = tot_red[1] - (float)n * (1.0f + rho);
        const float Str = tot_red[2];
        const float dms = tot_red[3] * inv_n;
        const float logdet = (float)n * logf(tot) + logdet2 * 0.69314718055994530942f;
        *mll_out = (-0.5f * quad - 0.5f * logdet - 0.5f * (float)n * 1.83787706640934548356f) * inv_n;
#pragma unroll
        for (int f = 0; f < FT; ++f)
          if (f < F) hyp[f] = rho * inv_n * tot_red[4 + f] * s_hyp[mtx][f][1];
        hyp[F] = 0.5f * inv_tot * inv_n * Str * s_hyp[mtx][4][1];
        hyp[F + 1] = 0.5f * inv_tot * inv_n * Skt * s_hyp[mtx][5][1];
        hyp[F + 2] = dms;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kCols>(tmem);
}

}  // namespace

// Tensor-memory GP kernel: 32 < n <= 64 (two matrices per CTA) and 64 < n <= 128 (one), F <= 4 (FT + 4 reduction slots <= 8).
int launch_gp_mll_tc(const GpArgs& a, cudaStream_t st) {
  if (a.n <= 32 || a.n > 128 || a.F < 1 || a.F > 4) return PACOH_ERR_UNSUPPORTED;
  if (a.P > 65535) return PACOH_ERR_UNSUPPORTED;
  const bool rag = a.task_n != nullptr;
  const dim3 grid(a.n <= 64 ? (a.T + 1) / 2 : a.T, a.P);
  // 8 CTAs / SM need 8 x 17 KB of shared memory: ask for the large carve-out (L1 is not used)
#define PACOH_GPTC_LAUNCH(FT_, MT_, RAG_)                                                                               \
  do {                                                                                                                  \
    static bool once = false;                                                                                           \
    if (!once) {                                                                                                        \
      PACOH_CUDA_CHECK(cudaFuncSetAttribute(gp_tc_kernel<FT_, MT_, RAG_>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                            cudaSharedmemCarveoutMaxShared));                                           \
      once = true;                                                                                                      \
    }                                                                                                                   \
    gp_tc_kernel<FT_, MT_, RAG_><<<grid, kGT, 0, st>>>(a);                                                              \
  } while (0)
  if (a.n <= 64) {
    if (a.F <= 2) { if (rag) PACOH_GPTC_LAUNCH(2, 2, true); else PACOH_GPTC_LAUNCH(2, 2, false); }
    else { if (rag) PACOH_GPTC_LAUNCH(4, 2, true); else PACOH_GPTC_LAUNCH(4, 2, false); }
  } else {
    if (a.F <= 2) { if (rag) PACOH_GPTC_LAUNCH(2, 1, true); else PACOH_GPTC_LAUNCH(2, 1, false); }
    else { if (rag) PACOH_GPTC_LAUNCH(4, 1, true); else PACOH_GPTC_LAUNCH(4, 1, false); }
  }
#undef PACOH_GPTC_LAUNCH
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

}  // namespace pacoh
