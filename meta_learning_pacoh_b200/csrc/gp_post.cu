// Eval-mode GP posterior and evaluation metrics on the device, batched over parameter vectors (particles /
// posterior samples) AND test tasks.
//
// Reference path replaced: get_pred_dist (GPR_meta_svgd.py:203-212, GPR_meta_vi.py:229-252, GPR_meta_mll.py:174-183) ->
// gpytorch ExactGP.eval() + likelihood, then abstract.py:134-181 (eval / eval_datasets: joint test log-likelihood, RMSE,
// calibration error) and models.py:74-140 (equally weighted mixture over the particles):
//
//   mu*   = m(X*) + K*c Kt^-1 (y_c - m(X_c))                 Sigma* = K** - K*c Kt^-1 Kc* + sigma^2 I
//   LL    = log N(y* | mu*, Sigma*)  (the full n* x n* covariance)
//
// How it is computed here (normalised matrix Khat = Kt / tot, tot = s + sigma^2, rho = s / tot, as in the MLL kernels):
//   * one CTA per (parameter vector, test task): Khat_cc = L L^T (n_c <= 128) is factorised in shared memory (potrf + trtri,
//     chol128.cuh), v = L^-1 r; then one thread per test point, forward substitution w_j = L^-1 khat_j:
//        mu_j = m_j + w_j . v                     var_j = tot (1 - w_j . w_j)                     khat_j = rho k(x_j, X_c)
//   * the joint log-likelihood needs no n* x n* factorisation of Sigma*: by the chain rule of Gaussians
//        log N(y* | mu*, Sigma*) = log N([y_c; y*] | joint prior) - log N(y_c | prior)
//     i.e. two calls of the batched marginal-log-likelihood kernels (values only) on the packed [context; test] point sets,
//     which run the tensor-core Cholesky for large n* -- same kernels, same jitter ladder as training.
//   * mixture mean / RMSE / calibration error / mixture log-likelihood: one small kernel per test task (pacoh_pred_metrics).
#include <math_constants.h>
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "kernels.cuh"
#include "chol128.cuh"

namespace pacoh {

namespace {

constexpr int NBP = 128;
constexpr float kFarP = 1.0e18f;
constexpr float kCP = 0.84932180028801904272f;   // sqrt(0.5 * log2(e))

__device__ __forceinline__ float ex2p(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float khat_p(const float4& ur, const float4& uc, float e0) {
  float e = e0;
  { const float du = ur.x - uc.x; e = fmaf(-du, du, e); }
  { const float du = ur.y - uc.y; e = fmaf(-du, du, e); }
  { const float du = ur.z - uc.z; e = fmaf(-du, du, e); }
  { const float du = ur.w - uc.w; e = fmaf(-du, du, e); }
  return ex2p(e);
}

struct PostArgs {
  const float* theta; int D, P, Tt, d, F;
  int mean_kind, has_oscale; float noise_floor;
  int off_ls, off_noise, off_oscale, off_const_mean;
  const float* xj; const float* yj;        // packed [context; test] points per task: (Tt, nj_max, d), (Tt, nj_max)
  const int* ncv; const int* njv;          // context / joint point counts per task
  int nj_max, nc_max, ns_max;
  const float* mean; const float* feat;    // (P, Tt * nj_max), (P, Tt * nj_max, F) from the nets, or nullptr
  float* mu; float* var;                   // (P, Tt, ns_max)
  int* info;                               // (P, Tt) or nullptr
  float* kcs;                              // (P, Tt, nc_max, ns_max) scratch for the full covariance: W = L^-1 Khat_c*, or nullptr
  float* cov;                              // (P, Tt, ns_max, ns_max) or nullptr
};

// [context rows | test rows | zero padding] per task, plus the per-task counts and the identity task index.
__global__ void post_pack_kernel(const float* __restrict__ xc, const float* __restrict__ yc, const int* __restrict__ nc_in,
                                 const float* __restrict__ xs, const int* __restrict__ ns_in, const float* __restrict__ ys,
                                 int nc_max, int ns_max, int d, float* __restrict__ xj, float* __restrict__ yj,
                                 int* __restrict__ njv, int* __restrict__ ncv, int* __restrict__ idx) {
  const int t = blockIdx.x, nj_max = nc_max + ns_max;
  const int nc = nc_in != nullptr ? nc_in[t] : nc_max, ns = ns_in != nullptr ? ns_in[t] : ns_max;
  for (int i = threadIdx.x; i < nj_max; i += blockDim.x) {
    float yv = 0.0f;
    const float* src = nullptr;
    if (i < nc) { src = xc + ((size_t)t * nc_max + i) * d; yv = yc[(size_t)t * nc_max + i]; }
    else if (i < nc + ns) { src = xs + ((size_t)t * ns_max + (i - nc)) * d; yv = ys != nullptr ? ys[(size_t)t * ns_max + (i - nc)] : 0.0f; }
    for (int k = 0; k < d; ++k) xj[((size_t)t * nj_max + i) * d + k] = src != nullptr ? src[k] : 0.0f;
    yj[(size_t)t * nj_max + i] = yv;
  }
  if (threadIdx.x == 0) { njv[t] = nc + ns; ncv[t] = nc; idx[t] = t; }
}

// One CTA per (test task, parameter vector), 256 threads, dynamic shared memory: two 128 x 132 tiles + vectors.
__global__ void __launch_bounds__(256, 1) gp_post_kernel(PostArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* T = sm;                          // Khat_cc -> L (lower triangle, kept for the forward substitutions)
  float* X = sm + NBP * LDT;              // U = L^-T (for v = U^T r), later the khat vectors of 128 test points
  float4* ucs = reinterpret_cast<float4*>(X + NBP * LDT);   // scaled context features
  float* rs = reinterpret_cast<float*>(ucs + NBP);          // residuals -> alphahat
  float* vs = rs + NBP;
  float* flag = vs + NBP;
  const int tid = threadIdx.x, r = tid & 127, h = tid >> 7;
  const int t = blockIdx.x, p = blockIdx.y;
  const int nc = a.ncv[t], ns = a.njv[t] - nc;
  const float* th = a.theta + (size_t)p * a.D;
  float sc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int f = 0; f < a.F && f < 4; ++f) sc[f] = kCP / softplus_f(__ldg(th + a.off_ls + f));
  const float sig2 = a.noise_floor + softplus_f(__ldg(th + a.off_noise));
  const float osc = a.has_oscale ? softplus_f(__ldg(th + a.off_oscale)) : 1.0f;
  const float cm = a.mean_kind == PACOH_MEAN_CONSTANT ? __ldg(th + a.off_const_mean) : 0.0f;
  const size_t Q = (size_t)a.Tt * a.nj_max;

  auto point = [&](int row, float4& u, float& m) {       // scaled features and prior mean of packed row `row` of task t
    const size_t q = (size_t)p * Q + (size_t)t * a.nj_max + row;
    float f4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int f = 0; f < a.F && f < 4; ++f)
      f4[f] = (a.feat != nullptr ? __ldg(a.feat + q * a.F + f) : __ldg(a.xj + ((size_t)t * a.nj_max + row) * a.d + f)) * sc[f];
    u = make_float4(f4[0], f4[1], f4[2], f4[3]);
    m = a.mean != nullptr ? __ldg(a.mean + q) : cm;
  };

  if (tid < NBP) {
    float4 u = make_float4(kFarP, kFarP, kFarP, kFarP);
    float res = 0.0f;
    if (tid < nc) {
      float m;
      point(tid, u, m);
      res = __ldg(a.yj + (size_t)t * a.nj_max + tid) - m;
    }
    ucs[tid] = u;
    rs[tid] = res;
  }
  __syncthreads();
  const float4 urow = ucs[r];
  float tot = osc + sig2, rho = osc / tot;
  int status = -1;
  for (int lvl = 0; lvl < 4; ++lvl) {                    // gpytorch psd_safe_cholesky: retry with jitter 1e-6, 1e-5, 1e-4
    const float jit = lvl == 0 ? 0.0f : (lvl == 1 ? 1e-6f : (lvl == 2 ? 1e-5f : 1e-4f));
    tot = osc + sig2 + jit;
    rho = osc / tot;
    const float e0 = r < nc ? log2f(rho) : -CUDART_INF_F;
    conv_sync();
#pragma unroll 4
    for (int i = 0; i < 64; ++i) {
      const int cc = 64 * h + i;
      float kv = khat_p(urow, ucs[cc], e0);
      if (cc == r) kv = 1.0f;
      T[r * LDT + cc] = kv;
    }
    const bool ok = potrf_trtri_128(T, X, tid, flag);
    if (ok) { status = lvl; break; }
  }
  if (a.info != nullptr && tid == 0) a.info[(size_t)p * a.Tt + t] = status;
  // v = L^-1 r = U^T r
  if (tid < NBP) {
    float v = 0.0f;
    for (int i = 0; i <= tid && i < nc; ++i) v = fmaf(X[i * LDT + tid], rs[i], v);
    vs[tid] = tid < nc ? v : 0.0f;
  }
  conv_sync();
  // ---- test points, 128 at a time (threads 0..127).  X is free now: it holds the vectors khat_j -> w_j = L^-1 khat_j
  //      ([i][thread], forward substitution in place -- the backward-stable form; k^T Khat^-1 k through an explicit inverse
  //      loses the 1e-4 bar on the variance where 1 - w.w cancels):  mu_j = m_j + w_j . v,  var_j = tot (1 - w_j . w_j)
  float* ks = X;
  const float lg2rho = log2f(rho);
  const bool failed = status < 0;
  for (int j0 = 0; j0 < ns; j0 += NBP) {
    const int j = j0 + tid;
    if (tid < NBP && j < ns) {
      float4 uj;
      float mj;
      point(nc + j, uj, mj);
      for (int i = 0; i < nc; ++i) ks[i * NBP + tid] = khat_p(uj, ucs[i], lg2rho);
      float mean = 0.0f, q = 0.0f;
      const size_t so = (((size_t)p * a.Tt + t) * a.nc_max) * a.ns_max + j;
      for (int i = 0; i < nc; ++i) {
        float sacc = ks[i * NBP + tid];
        const float* Li = T + i * LDT;
        for (int m = 0; m < i; ++m) sacc = fmaf(-Li[m], ks[m * NBP + tid], sacc);
        const float w = sacc / Li[i];
        ks[i * NBP + tid] = w;
        mean = fmaf(w, vs[i], mean);
        q = fmaf(w, w, q);
        if (a.kcs != nullptr) a.kcs[so + (size_t)i * a.ns_max] = w;
      }
      const size_t o = ((size_t)p * a.Tt + t) * a.ns_max + j;
      a.mu[o] = failed ? CUDART_NAN_F : mj + mean;
      a.var[o] = failed ? CUDART_NAN_F : tot * (1.0f - q);
    }
    __syncthreads();
  }
  // padding entries of mu / var
  for (int j = ns + tid; j < a.ns_max; j += blockDim.x) {
    const size_t o = ((size_t)p * a.Tt + t) * a.ns_max + j;
    a.mu[o] = 0.0f;
    a.var[o] = 1.0f;
  }
}

// Full predictive covariance (only on request: predict(return_density=True).covariance_matrix):
//   Sigma*_jk = tot (rho k(x_j, x_k) + (1 - rho) delta_jk - w_j . w_k),  w = L^-1 khat.   grid (ceil(ns/16), ceil(ns/16), P * Tt)
__global__ void gp_post_cov_kernel(PostArgs a) {
  const int pt = blockIdx.z, p = pt / a.Tt, t = pt - p * a.Tt;
  const int nc = a.ncv[t], ns = a.njv[t] - nc;
  const int j = blockIdx.y * 16 + (threadIdx.x >> 4), k = blockIdx.x * 16 + (threadIdx.x & 15);
  if (j >= a.ns_max || k >= a.ns_max) return;
  float* out = a.cov + (((size_t)pt * a.ns_max) + j) * a.ns_max + k;
  if (j >= ns || k >= ns) { *out = j == k ? 1.0f : 0.0f; return; }
  const float* th = a.theta + (size_t)p * a.D;
  const float sig2 = a.noise_floor + softplus_f(__ldg(th + a.off_noise));
  const float osc = a.has_oscale ? softplus_f(__ldg(th + a.off_oscale)) : 1.0f;
  const float tot = osc + sig2, rho = osc / tot;      // (a jitter level used by the context factorisation is ignored here: <= 1e-4)
  const size_t Q = (size_t)a.Tt * a.nj_max;
  float d2 = 0.0f;
  for (int f = 0; f < a.F; ++f) {
    const float scf = kCP / softplus_f(__ldg(th + a.off_ls + f));
    const size_t qj = (size_t)p * Q + (size_t)t * a.nj_max + nc + j, qk = (size_t)p * Q + (size_t)t * a.nj_max + nc + k;
    const float fj = a.feat != nullptr ? a.feat[qj * a.F + f] : a.xj[((size_t)t * a.nj_max + nc + j) * a.d + f];
    const float fk = a.feat != nullptr ? a.feat[qk * a.F + f] : a.xj[((size_t)t * a.nj_max + nc + k) * a.d + f];
    const float du = (fj - fk) * scf;
    d2 = fmaf(du, du, d2);
  }
  float s = 0.0f;
  const size_t so = ((size_t)pt * a.nc_max) * a.ns_max;
  for (int i = 0; i < nc; ++i) s = fmaf(a.kcs[so + (size_t)i * a.ns_max + j], a.kcs[so + (size_t)i * a.ns_max + k], s);
  *out = tot * ((j == k ? 1.0f : rho * ex2p(-d2)) - s);
}

// joint_ll[p, t] = n_j mll_joint - n_c mll_context  (the MLL kernels return log N(.) / n)
__global__ void post_jll_kernel(const float* __restrict__ mllj, const float* __restrict__ mllc, const int* __restrict__ njv,
                                const int* __restrict__ ncv, const int* __restrict__ infoj, int* __restrict__ info, int P, int Tt,
                                float* __restrict__ jll) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * Tt) return;
  const int t = i % Tt;
  jll[i] = (float)njv[t] * mllj[i] - (float)ncv[t] * mllc[i];
  if (info != nullptr && infoj != nullptr && info[i] >= 0) info[i] = infoj[i] < 0 ? -1 : max(info[i], infoj[i]);
}

// One block per test task: mixture statistics over the P parameter vectors (abstract.py:157-161, models.py:90-126).
__global__ void pred_metrics_kernel(int P, int Tt, int ns_max, const float* __restrict__ mu, const float* __restrict__ var,
                                    const int* __restrict__ ns_in, const float* __restrict__ yn, const float* __restrict__ jll,
                                    float y_std, float* __restrict__ out) {
  __shared__ float red[32];
  __shared__ int cnt[20];
  const int t = blockIdx.x, ns = ns_in != nullptr ? ns_in[t] : ns_max;
  if (threadIdx.x < 20) cnt[threadIdx.x] = 0;
  __syncthreads();
  float se = 0.0f;
  for (int j = threadIdx.x; j < ns; j += blockDim.x) {
    const float y = yn[(size_t)t * ns_max + j];
    float m = 0.0f, c = 0.0f;
    for (int p = 0; p < P; ++p) {
      const size_t o = ((size_t)p * Tt + t) * ns_max + j;
      const float mp = mu[o], sp = sqrtf(var[o]);
      m += mp;
      c += 0.5f * (1.0f + erff((y - mp) / (sp * 1.41421356237f)));      // Normal cdf; the affine un-normalisation cancels
    }
    m /= (float)P; c /= (float)P;
    se = fmaf(m - y, m - y, se);
    for (int k = 0; k < 20; ++k) {
      const float conf = 0.05f + (0.95f - 0.05f) * (float)k / 19.0f;   // torch.linspace(0.05, 0.95, 20)
      if (c <= conf) atomicAdd(&cnt[k], 1);
    }
  }
  se = warp_sum(se);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float rmse = y_std * sqrtf(tot / (float)ns);
    float cal = 0.0f;
    for (int k = 0; k < 20; ++k) {
      const float conf = 0.05f + (0.95f - 0.05f) * (float)k / 19.0f;
      const float e = (float)cnt[k] / (float)ns - conf;
      cal = fmaf(e, e, cal);
    }
    cal = sqrtf(cal / 20.0f);
    float ll = CUDART_NAN_F;
    if (jll != nullptr) {       // mixture: logsumexp over the particles - log P; affine Jacobian - n* log y_std; per point
      float mx = -CUDART_INF_F;
      for (int p = 0; p < P; ++p) mx = fmaxf(mx, jll[(size_t)p * Tt + t]);
      float s = 0.0f;
      for (int p = 0; p < P; ++p) s += expf(jll[(size_t)p * Tt + t] - mx);
      ll = (mx + logf(s) - logf((float)P) - (float)ns * logf(y_std)) / (float)ns;
    }
    out[t * 3 + 0] = ll; out[t * 3 + 1] = rmse; out[t * 3 + 2] = cal;
  }
}

struct PostPlan {
  ModelDev m;
  int nj_max;
  bool big;
  size_t off_xj, off_yj, off_nj, off_nc, off_idx, off_mean, off_feat, off_mllj, off_mllc, off_hyp, off_infoj, off_kcs,
      off_fwd, fwd_bytes, off_big, big_bytes, total;     // byte offsets
};

size_t al256p(size_t v) { return (v + 255) & ~(size_t)255; }

int post_plan(const pacoh_arch_t* arch, int P, int Tt, int nc_max, int ns_max, bool want_cov, PostPlan* pl) {
  if (!build_model(arch, &pl->m)) { set_error("invalid architecture descriptor"); return PACOH_ERR_INVALID; }
  if (P < 1 || Tt < 1 || nc_max < 1 || ns_max < 1) { set_error("pacoh_gp_posterior: P, tasks, n_c, n* must be positive"); return PACOH_ERR_INVALID; }
  const ModelDev& m = pl->m;
  if (nc_max > NBP) { set_error("pacoh_gp_posterior: context sets of more than %d points are not supported (n_c = %d)", NBP, nc_max); return PACOH_ERR_UNSUPPORTED; }
  if (m.F > 4) { set_error("pacoh_gp_posterior: feature dim %d > 4 not supported", m.F); return PACOH_ERR_UNSUPPORTED; }
  pl->nj_max = nc_max + ns_max;
  pl->big = gp_use_big(pl->nj_max, m.F);
  if (pl->nj_max > (pl->big ? kMaxBigN : kMaxGpN)) { set_error("pacoh_gp_posterior: n_c + n* = %d points not supported", pl->nj_max); return PACOH_ERR_UNSUPPORTED; }
  const size_t pts = (size_t)Tt * pl->nj_max, PT = (size_t)P * Tt;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = al256p(off + bytes); return o; };
  pl->off_xj = take(pts * m.d * 4); pl->off_yj = take(pts * 4);
  pl->off_nj = take((size_t)Tt * 4); pl->off_nc = take((size_t)Tt * 4); pl->off_idx = take((size_t)Tt * 4);
  pl->off_mean = take(m.mean_kind == PACOH_MEAN_NN ? (size_t)P * pts * 4 : 0);
  pl->off_feat = take(m.covar_kind == PACOH_COVAR_NN ? (size_t)P * pts * m.F * 4 : 0);
  pl->off_mllj = take(PT * 4); pl->off_mllc = take(PT * 4);
  pl->off_hyp = take(PT * gp_hyp_stride(m.F) * 4);
  pl->off_infoj = take(PT * 4);
  pl->off_kcs = take(want_cov ? PT * nc_max * ns_max * 4 : 0);
  const int64_t fb = pacoh_gp_forward_workspace_bytes(arch, P, (int32_t)pts);
  if (fb < 0) return (int)fb;
  pl->fwd_bytes = (size_t)fb;
  pl->off_fwd = take(pl->fwd_bytes);
  pl->big_bytes = pl->big ? gp_big_workspace_bytes(pl->nj_max, (long long)PT) : 0;
  pl->off_big = take(pl->big_bytes);
  pl->total = off;
  return PACOH_OK;
}

}  // namespace
}  // namespace pacoh

using namespace pacoh;

extern "C" int64_t pacoh_gp_posterior_workspace_bytes(const pacoh_arch_t* arch, int32_t P, int32_t Tt, int32_t nc_max, int32_t ns_max,
                                                      int32_t want_cov) {
  PostPlan pl;
  const int rc = post_plan(arch, P, Tt, nc_max, ns_max, want_cov != 0, &pl);
  return rc != PACOH_OK ? rc : (int64_t)pl.total;
}

extern "C" int pacoh_gp_posterior(const pacoh_arch_t* arch, int32_t P, int32_t Tt, int32_t nc_max, int32_t ns_max, const float* theta,
                                  const float* x_c, const float* y_c, const int32_t* n_c, const float* x_s, const int32_t* n_s,
                                  const float* y_s, float* mu, float* var, float* cov, float* joint_ll, int32_t* info,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  PostPlan pl;
  int rc = post_plan(arch, P, Tt, nc_max, ns_max, cov != nullptr, &pl);
  if (rc != PACOH_OK) return rc;
  if (!theta || !x_c || !y_c || !x_s || !mu || !var || !workspace || (joint_ll && !y_s)) {
    set_error("pacoh_gp_posterior: null pointer argument (joint_ll needs y_s)");
    return PACOH_ERR_INVALID;
  }
  if ((size_t)workspace_bytes < pl.total) { set_error("pacoh_gp_posterior: workspace %lld < %zu bytes", (long long)workspace_bytes, pl.total); return PACOH_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const ModelDev& m = pl.m;
  uint8_t* ws = (uint8_t*)workspace;
  float* xj = (float*)(ws + pl.off_xj); float* yj = (float*)(ws + pl.off_yj);
  int* njv = (int*)(ws + pl.off_nj); int* ncv = (int*)(ws + pl.off_nc); int* idx = (int*)(ws + pl.off_idx);
  float* mean = m.mean_kind == PACOH_MEAN_NN ? (float*)(ws + pl.off_mean) : nullptr;
  float* feat = m.covar_kind == PACOH_COVAR_NN ? (float*)(ws + pl.off_feat) : nullptr;

  post_pack_kernel<<<Tt, 128, 0, st>>>(x_c, y_c, n_c, x_s, n_s, y_s, nc_max, ns_max, m.d, xj, yj, njv, ncv, idx);
  PACOH_CUDA_CHECK(cudaGetLastError());
  if (mean != nullptr || feat != nullptr) {
    rc = pacoh_gp_forward(arch, P, Tt * pl.nj_max, theta, xj, mean, feat, ws + pl.off_fwd, (int64_t)pl.fwd_bytes, stream);
    if (rc != PACOH_OK) return rc;
  }
  PostArgs a;
  memset(&a, 0, sizeof(a));
  a.theta = theta; a.D = m.D; a.P = P; a.Tt = Tt; a.d = m.d; a.F = m.F;
  a.mean_kind = m.mean_kind; a.has_oscale = m.has_oscale; a.noise_floor = m.noise_floor;
  a.off_ls = m.off_ls; a.off_noise = m.off_noise; a.off_oscale = m.off_oscale; a.off_const_mean = m.off_const_mean;
  a.xj = xj; a.yj = yj; a.ncv = ncv; a.njv = njv; a.nj_max = pl.nj_max; a.nc_max = nc_max; a.ns_max = ns_max;
  a.mean = mean; a.feat = feat; a.mu = mu; a.var = var; a.info = info; a.cov = cov;
  a.kcs = cov != nullptr ? (float*)(ws + pl.off_kcs) : nullptr;
  const size_t smem = sizeof(float) * (2 * NBP * LDT + 4 * NBP + 2 * NBP + 64);
  static bool once = false;
  if (!once) { PACOH_CUDA_CHECK(cudaFuncSetAttribute(gp_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
  gp_post_kernel<<<dim3(Tt, P), 256, smem, st>>>(a);
  PACOH_CUDA_CHECK(cudaGetLastError());
  if (cov != nullptr) {
    gp_post_cov_kernel<<<dim3((ns_max + 15) / 16, (ns_max + 15) / 16, P * Tt), 256, 0, st>>>(a);
    PACOH_CUDA_CHECK(cudaGetLastError());
  }
  if (joint_ll != nullptr) {
    GpArgs g;
    memset(&g, 0, sizeof(g));
    g.theta = theta; g.x = xj; g.y = yj; g.task_idx = idx; g.mean = mean; g.feat = feat;
    g.dhyp = (float*)(ws + pl.off_hyp); g.info = (int*)(ws + pl.off_infoj);
    g.P = P; g.T = Tt; g.n = pl.nj_max; g.d = m.d; g.F = m.F; g.D = m.D;
    g.mean_kind = m.mean_kind; g.covar_kind = m.covar_kind; g.has_oscale = m.has_oscale; g.noise_floor = m.noise_floor;
    g.off_ls = m.off_ls; g.off_noise = m.off_noise; g.off_oscale = m.off_oscale; g.off_const_mean = m.off_const_mean;
    g.values_only = 1;
    for (int pass = 0; pass < 2; ++pass) {       // joint [context; test] sets, then the context sets alone (same arrays, fewer rows)
      g.task_n = pass == 0 ? njv : ncv;
      g.mll = (float*)(ws + (pass == 0 ? pl.off_mllj : pl.off_mllc));
      rc = pl.big ? launch_gp_mll_big(g, ws + pl.off_big, pl.big_bytes, st) : launch_gp_mll(g, st);
      if (rc != PACOH_OK) return rc;
      if (pass == 0) g.info = nullptr;
    }
    post_jll_kernel<<<(P * Tt + 255) / 256, 256, 0, st>>>((float*)(ws + pl.off_mllj), (float*)(ws + pl.off_mllc), njv, ncv,
                                                         (int*)(ws + pl.off_infoj), info, P, Tt, joint_ll);
    PACOH_CUDA_CHECK(cudaGetLastError());
  }
  return PACOH_OK;
}

extern "C" int pacoh_pred_metrics(int32_t P, int32_t Tt, int32_t ns_max, const float* mu, const float* var, const int32_t* n_s,
                                  const float* y_s, const float* joint_ll, float y_std, float* out, void* stream) {
  if (P < 1 || Tt < 1 || ns_max < 1 || !mu || !var || !y_s || !out) { set_error("pacoh_pred_metrics: invalid argument"); return PACOH_ERR_INVALID; }
  pred_metrics_kernel<<<Tt, 256, 0, (cudaStream_t)stream>>>(P, Tt, ns_max, mu, var, n_s, y_s, joint_ll, y_std, out);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}
