/*
 * pacoh_b200.h -- C ABI of the B200-native engine for PACOH's meta-training hot path.
 *
 * The reference (jonasrothfuss/meta_learning_pacoh) is pure Python and has no FFI; the seam this
 * library sits behind is where its inference objectives (L3) call the "random GP" log-density (L2).
 * Every entry point names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch types.  All tensor pointers are DEVICE pointers to
 *     contiguous fp32 (int32 where noted); the caller owns every buffer.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no allocation and no
 *     host synchronisation happens inside, scratch comes from the caller-provided workspace.
 *   - return value: PACOH_OK (0) or a negative PACOH_ERR_* code; per-matrix numerical status is
 *     written to `info` on the device (0 clean, 1..3 jitter level used, <0 not positive definite),
 *     mirroring gpytorch's psd_safe_cholesky jitter ladder 1e-6/1e-5/1e-4.
 *   - parameter vectors use the reference's flat layout (meta_learn/models.py:266-277,319-323):
 *     mean module, covariance module, noise_raw; MLP layers fc_1..fc_L,out; bias before weight;
 *     weight row-major (out,in).  `outputscale_raw` (PACOH-MAP only) is appended last.
 */
#ifndef PACOH_B200_H
#define PACOH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PACOH_ABI_VERSION 1
#define PACOH_MAX_LAYERS 8

enum {
  PACOH_OK = 0,
  PACOH_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, ...)            */
  PACOH_ERR_UNSUPPORTED = -2,  /* architecture / size outside what the kernels implement     */
  PACOH_ERR_WORKSPACE = -3,    /* workspace too small: call pacoh_workspace_bytes            */
  PACOH_ERR_CUDA = -4          /* a CUDA runtime call failed (see pacoh_last_error)          */
};

enum { PACOH_MEAN_ZERO = 0, PACOH_MEAN_CONSTANT = 1, PACOH_MEAN_NN = 2 };
enum { PACOH_COVAR_SE = 0, PACOH_COVAR_NN = 1 };
enum { PACOH_SVGD_RBF = 0, PACOH_SVGD_IMQ = 1 };

/* Architecture of one "vectorised GP" (meta_learn/random_gp.py:22-52 VectorizedGP.__init__;
 * PACOH-MAP: meta_learn/GPR_meta_mll.py:207-251 with has_outputscale=1, noise_floor=1e-3). */
typedef struct {
  int32_t input_dim;                        /* d                                              */
  int32_t mean_kind;                        /* PACOH_MEAN_*                                   */
  int32_t covar_kind;                       /* PACOH_COVAR_*                                  */
  int32_t n_mean_layers;                    /* hidden layers of the mean MLP                  */
  int32_t mean_layers[PACOH_MAX_LAYERS];    /* their widths                                   */
  int32_t n_kernel_layers;
  int32_t kernel_layers[PACOH_MAX_LAYERS];
  int32_t feature_dim;                      /* F: kernel-MLP output dim (ignored for SE: F=d) */
  int32_t has_outputscale;                  /* 1: K = softplus(outputscale_raw) * exp(...)    */
  float noise_floor;                        /* sigma^2 = noise_floor + softplus(noise_raw)    */
} pacoh_arch_t;

int pacoh_abi_version(void);
const char* pacoh_last_error(void);

/* Length D of one parameter vector for `arch` (RandomGPMeta.parameter_shapes, random_gp.py:186-190). */
int64_t pacoh_param_count(const pacoh_arch_t* arch);

/* Per-coordinate (mu, sigma) of the factorised Gaussian hyper-prior, written to HOST arrays of length D
 * (_RandomGPBase.__init__, random_gp.py:118-157). */
int pacoh_hyper_prior_params(const pacoh_arch_t* arch, float weight_prior_std, float bias_prior_std,
                             float* mu_host, float* sigma_host);

/* Scratch bytes needed by pacoh_meta_mll_fwd_bwd for P parameter vectors, T batch tasks of n points (n <= 4096; more than 4
 * kernel features only for n <= 64).  Up to 64 points per task the matrices live on chip and the scratch holds the hand-over
 * buffers between the MLP and GP kernels; above, it also holds the Cholesky factors of the blocked tensor-core path
 * (~ 1.13 n^2 floats per (parameter vector, task), processed in passes of at most 24 GB: see pacoh_debug_big_layout). */
int64_t pacoh_workspace_bytes(const pacoh_arch_t* arch, int32_t P, int32_t T, int32_t n);

/*
 * Batched GP marginal log-likelihood, forward + backward, for every (particle, task) pair.
 * Replaces RandomGPMeta._log_prob_likelihood + autograd (random_gp.py:206-219, svgd.py:15-16,
 * GPR_meta_vi.py:221) and the PACOH-MAP loop body (GPR_meta_mll.py:109-113):
 *
 *   mll[p,t]      = log N(y_t | m_p(x_t), K_p(x_t,x_t) + sigma_p^2 I) / n        (dense Cholesky for every n: the reference's
 *                   gpytorch switches to CG / Lanczos above settings.max_cholesky_size; parity is defined on the dense path)
 *   mll_sum[p]    = sum_t mll[p,t]                      (duplicates in task_idx count twice)
 *   dtheta_lik    = d mll_sum[p] / d theta[p,:]         (P, D)
 *
 *   theta     (P, D)            parameter vectors
 *   x         (T_total, n, d)   all tasks' normalised inputs;  y (T_total, n) targets
 *   task_idx  (T) int32         indices into the T_total tasks, in batch order, may repeat
 *   mll       (P, T) or NULL ;  mll_sum (P) ;  dtheta_lik (P, D) ;  info (P, T) int32 or NULL
 */
int pacoh_meta_mll_fwd_bwd(const pacoh_arch_t* arch, int32_t P, int32_t T, int32_t n,
                           const float* theta, const float* x, const float* y, const int32_t* task_idx,
                           float* mll, float* mll_sum, float* dtheta_lik, int32_t* info,
                           void* workspace, int64_t workspace_bytes, void* stream);

/*
 * The same for RAGGED task sets (tasks with different numbers of points, which the reference's per-task loop handles
 * implicitly: random_gp.py:214-217): x / y are padded to n = max_t n_t rows per task and task_n (T_total) int32 holds
 * every task's own n_t (1 <= n_t <= n).  Rows n_t .. n-1 of a task are ignored (any finite padding values), mll[p,t]
 * is divided by n_t, and the padding rows receive exactly zero gradient.  task_n == NULL: all tasks have n points.
 * The harmonic-mean pre_factor of random_gp.py:209-212 is the caller's (pacoh_logprob_finalize takes it as a scalar).
 */
int pacoh_meta_mll_fwd_bwd_ragged(const pacoh_arch_t* arch, int32_t P, int32_t T, int32_t n,
                                  const float* theta, const float* x, const float* y, const int32_t* task_n,
                                  const int32_t* task_idx, float* mll, float* mll_sum, float* dtheta_lik, int32_t* info,
                                  void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Forward pass of the learned mean / kernel-feature nets for every parameter vector at `npts` points
 * (LearnedGPRegressionModel.forward's NN calls, models.py:505-514; used by the eval-mode posterior,
 * GPR_meta_svgd.py:203-212).  x (npts, d) -> mean (P, npts) [mean_kind NN, else untouched / may be NULL],
 * feat (P, npts, F) [covar_kind NN, else untouched / may be NULL].
 */
int64_t pacoh_gp_forward_workspace_bytes(const pacoh_arch_t* arch, int32_t P, int32_t npts);
int pacoh_gp_forward(const pacoh_arch_t* arch, int32_t P, int32_t npts, const float* theta, const float* x,
                     float* mean, float* feat, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Eval-mode exact GP posterior for every parameter vector and a BATCH of test tasks, everything in normalised space
 * (get_pred_dist: GPR_meta_svgd.py:203-212, GPR_meta_vi.py:229-252, GPR_meta_mll.py:174-183 -> gpytorch ExactGP.eval +
 * likelihood; abstract.py:134-163):
 *
 *   mu[p,t,j]  = m_p(x*_j) + K*c Kt^-1 (y_c - m_p(X_c))            var[p,t,j] = diag(K** - K*c Kt^-1 Kc*) + sigma_p^2
 *   joint_ll[p,t] = log N(y*_t | mu, Sigma* incl. noise)  with the FULL n* x n* predictive covariance, computed as
 *                   log N([y_c; y*]) - log N(y_c) with the batched marginal-likelihood kernels (no n* x n* factorisation)
 *
 *   x_c (Tt, nc_max, d), y_c (Tt, nc_max), n_c (Tt) int32 or NULL (all nc_max): context sets, zero padded;  nc_max <= 128
 *   x_s (Tt, ns_max, d), n_s (Tt) int32 or NULL, y_s (Tt, ns_max) or NULL: test inputs / targets (targets only for joint_ll)
 *   mu, var (P, Tt, ns_max);  cov (P, Tt, ns_max, ns_max) or NULL (full predictive covariance incl. noise, on request);
 *   joint_ll (P, Tt) or NULL;  info (P, Tt) int32 or NULL (jitter level used / -1, as for the MLL)
 */
int64_t pacoh_gp_posterior_workspace_bytes(const pacoh_arch_t* arch, int32_t P, int32_t Tt, int32_t nc_max, int32_t ns_max,
                                           int32_t want_cov);
int pacoh_gp_posterior(const pacoh_arch_t* arch, int32_t P, int32_t Tt, int32_t nc_max, int32_t ns_max, const float* theta,
                       const float* x_c, const float* y_c, const int32_t* n_c, const float* x_s, const int32_t* n_s,
                       const float* y_s, float* mu, float* var, float* cov, float* joint_ll, int32_t* info, void* workspace,
                       int64_t workspace_bytes, void* stream);

/*
 * Evaluation metrics of the equally weighted mixture over the P parameter vectors, one row per test task
 * (abstract.py:157-161, 260-272; models.py:90-126):  out (Tt, 3) = [ avg test log-likelihood, RMSE, calibration error ]
 *   ll    = ( logsumexp_p joint_ll[p,t] - log P - n* log y_std ) / n*          (NaN when joint_ll == NULL)
 *   rmse  = y_std sqrt( mean_j (mean_p mu[p,t,j] - y_s[t,j])^2 )               (y_s: NORMALISED targets)
 *   calib = sqrt( mean_k ( mean_j [ mean_p Phi((y_j - mu_pj)/sd_pj) <= c_k ] - c_k )^2 ),  c = linspace(0.05, 0.95, 20)
 */
int pacoh_pred_metrics(int32_t P, int32_t Tt, int32_t ns_max, const float* mu, const float* var, const int32_t* n_s,
                       const float* y_s, const float* joint_ll, float y_std, float* out, void* stream);

/*
 * Hyper-prior log-density + combination (RandomGPMeta.log_prob, random_gp.py:179-180,221-222;
 * CatDist.log_prob, models.py:159-181):
 *
 *   logp[p]      = prior_factor * sum_k logN(theta[p,k]; mu_k, sigma_k) + pre_factor * mll_sum[p]
 *   dtheta[p,k]  = prior_factor * (-(theta[p,k]-mu_k)/sigma_k^2)       + pre_factor * dtheta_lik[p,k]
 *
 * mll_sum / dtheta_lik are the (all-reduced, when task-sharded) outputs of pacoh_meta_mll_fwd_bwd;
 * pre_factor = n_h / (n_h + T_global) is computed by the caller (random_gp.py:209-212).
 */
int pacoh_logprob_finalize(int32_t P, int64_t D, const float* theta, const float* prior_mu,
                           const float* prior_sigma, float prior_factor, float pre_factor,
                           const float* mll_sum, const float* dtheta_lik, float* logp, float* dtheta,
                           void* stream);

/*
 * Task-sharded runs (one process per GPU of an NVLink / NVSwitch node): pacoh_logprob_finalize with the all-reduce of
 * the packed likelihood buffer (dtheta_lik (P, D) | mll_sum (P), P*D + P floats) FUSED in, over peer memory instead of
 * an NCCL launch.  peer_bufs[r] / peer_flags[r] (host arrays of `world` device pointers, world <= PACOH_MAX_PEERS) are
 * rank r's packed buffer for this step and rank r's flag array (>= world uint32, zero-initialised once), both in
 * memory every rank of the node has mapped (CUDA IPC / symmetric memory).  `token` must increase by one per call on
 * every rank (start at 1) and the caller must alternate between two packed buffers (token parity): rank r's kernel
 * releases `token` into slot r of every peer's flags, acquires all of its own slots, then sums the `world` buffers in
 * rank order (bitwise identical on all ranks) while applying prior and pre_factor exactly like pacoh_logprob_finalize.
 * Replaces the autograd accumulation over the per-task loop of random_gp.py:200-222 across devices.
 */
#define PACOH_MAX_PEERS 8
int pacoh_peer_allreduce_finalize(int32_t world, int32_t rank, const void* const* peer_bufs, void* const* peer_flags,
                                  uint32_t token, int32_t P, int64_t D, const float* theta, const float* prior_mu,
                                  const float* prior_sigma, float prior_factor, float pre_factor, float* logp,
                                  float* dtheta, void* stream);

/* The same with the token taken from the device: token_dev (int32, device, may be NULL) is the step counter of a
 * CUDA-graph-captured training loop (pacoh_step_prepare) and the token of this call is *token_dev + token.  err_flag
 * (int32, device, may be NULL) is set non-zero if a peer never announced within the bounded wait (a dead rank): the
 * kernel then returns garbage sums instead of hanging the node. */
int pacoh_peer_allreduce_finalize_dev(int32_t world, int32_t rank, const void* const* peer_bufs, void* const* peer_flags,
                                      uint32_t token, const void* token_dev, int32_t* err_flag, int32_t P, int64_t D,
                                      const float* theta, const float* prior_mu, const float* prior_sigma, float prior_factor,
                                      float pre_factor, float* logp, float* dtheta, void* stream);

/*
 * SVGD direction (SVGD.phi tail + RBF_Kernel, svgd.py:18-21,32-59,103-107):
 *   d2_ij = |theta_i|^2 + |theta_j|^2 - 2 theta_i.theta_j ;  gamma = 1/(1e-8 + 2 h^2)
 *   bandwidth h > 0 fixed, or h <= 0: median heuristic  h^2 = median(d2 over all P*P) / (2 log(P+1))
 *   phi = (K score + 2 gamma (rowsum(K) * theta - K theta)) / P
 * gamma_out (1 float, device) receives gamma.  workspace: pacoh_svgd_workspace_bytes(P, D).
 *
 * kernel_kind PACOH_SVGD_IMQ (IMQSteinKernel(alpha=0.5, beta=-0.5), svgd.py:63-99, as GPR_meta_svgd.py:176-177 builds it):
 *   K_ij = (alpha + sum_d (theta_jd - theta_id)^2 / h_d)^beta ;  bandwidth > 0: h_d = bandwidth, else
 *   h_d = lower median over the pairs i < j of (theta_jd - theta_id)^2, divided by log(P+1)  (one on-device sort per d)
 *   phi = (K score - d sum(K)/d theta) / P, the derivative taken like SVGD.phi takes it (first kernel argument only,
 *   THROUGH the median: svgd.py:18-19).  gamma_out is set to 0.  P >= 2.
 *
 * The kernel matrix depends on the particles only, not on the scores, so the call is also exposed in two stages:
 * pacoh_svgd_kernel_matrix (distances, median, K, row sums -> workspace, gamma) may run on a side stream WHILE the
 * batched MLL forward/backward computes the scores; pacoh_svgd_phi_apply then only contracts K with score / theta.
 * pacoh_svgd_phi = both stages back to back on one stream.
 */
int64_t pacoh_svgd_workspace_bytes(int32_t P, int64_t D);
int pacoh_svgd_phi(int32_t P, int64_t D, const float* theta, const float* score, float bandwidth,
                   int32_t kernel_kind, float* phi, float* gamma_out, void* workspace,
                   int64_t workspace_bytes, void* stream);
int pacoh_svgd_kernel_matrix(int32_t P, int64_t D, const float* theta, float bandwidth, int32_t kernel_kind,
                             float* gamma_out, void* workspace, int64_t workspace_bytes, void* stream);
int pacoh_svgd_phi_apply(int32_t P, int64_t D, const float* theta, const float* score, int32_t kernel_kind,
                         float* phi, const float* gamma, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Diagonal-Gaussian VI posterior (RandomGPPosterior cov_type='diag', random_gp.py:244-263;
 * get_neg_elbo, GPR_meta_vi.py:216-224).
 *   sample:  theta[s,k] = loc[k] + exp(scale[k]) * eps[s,k] ;  logq[s] = sum_k logN(theta; loc, exp(scale))
 *   grad  :  loss = -(1/S) sum_s (logp_s - prior_factor*logq_s) with g = dlogp/dtheta (S, D):
 *            dloc[k] = -(1/S) sum_s g[s,k] ;  dscale[k] = -(1/S) sum_s g[s,k] eps[s,k] exp(scale[k]) - prior_factor
 */
int pacoh_vi_sample(int32_t S, int64_t D, const float* loc, const float* scale, const float* eps,
                    float* theta, float* logq, void* stream);
int pacoh_vi_grad(int32_t S, int64_t D, const float* scale, const float* eps, const float* g,
                  float prior_factor, float* dloc, float* dscale, void* stream);

/*
 * Fused Adam update of the (P, D) particle matrix (torch.optim.Adam semantics, no amsgrad / weight decay), replacing
 * `particles.grad = -phi; optim.step()` (svgd.py:27-28; GPR_meta_svgd.py:217-225).  `grad_sign` = -1 feeds grad = -phi
 * directly.  step is the 1-based step count AFTER this update.  All buffers have `count` floats, updated in place.
 */
int pacoh_adam_step(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg, float* exp_avg_sq,
                    float lr, float beta1, float beta2, float eps, int64_t step, void* stream);

/*
 * Device-side step state for training loops captured in a CUDA graph (the meta_fit loops GPR_meta_svgd.py:100-104,
 * GPR_meta_vi.py:103-110, GPR_meta_mll.py:104-117: sample a batch, step, optimizer.step(), lr_scheduler.step()).
 * `state` = 8 x 4 bytes on the device, zero-initialised: [0] int32 number of completed steps, [2] float learning rate,
 * [3] float lr / (1 - beta1^step), [4] float 1 / sqrt(1 - beta2^step).
 * pacoh_step_prepare: step += 1; lr = lr0 * gamma^((step - 1) / decay_every) (torch StepLR; decay_every <= 0: constant);
 * copies slot (old step mod K) of idx_stream (K, T) int32 to idx_out (T) and of fstream (K, F) to fout (F) (either may be
 * NULL): the host uploads K steps' worth of sampled task indices (numpy RandomState.choice stream) / normal draws at once.
 * pacoh_adam_step_dev: pacoh_adam_step with lr and bias corrections read from `state`.
 */
int pacoh_step_prepare(void* state, int32_t K, int32_t T, const int32_t* idx_stream, int32_t* idx_out, int64_t F,
                       const float* fstream, float* fout, float lr0, float gamma, int32_t decay_every, float beta1, float beta2,
                       void* stream);
int pacoh_adam_step_dev(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg, float* exp_avg_sq,
                        float beta1, float beta2, float eps, const void* state, void* stream);
/* torch.optim.AdamW on a flat parameter buffer (PACOH-MAP: GPR_meta_mll.py:253-257, decay on every group): decoupled weight
 * decay p *= 1 - lr * weight_decay, then the Adam update; `mask` (count bytes, may be NULL) freezes the entries that are 0. */
int pacoh_adamw_step_dev(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg, float* exp_avg_sq,
                         float beta1, float beta2, float eps, float weight_decay, const uint8_t* mask, const void* state,
                         void* stream);

/*
 * Host-only helper: the persistent schedule of the tensor-core MLP backward kernel (one CTA per SM per wave, the
 * nets * P * ceil(points / 128) tiles cut into equal ranges): CTAs, tiles per CTA, and the number of partial-gradient
 * slots a (net, particle) can be split into.  No device work; lets the schedule arithmetic be tested on a CPU-only host.
 */
int pacoh_mlp_bwd_schedule(int32_t P, int32_t nets, int64_t points, int32_t* grid, int32_t* tiles_per_cta, int32_t* slots);

/*
 * Optional per-stage device timing of pacoh_meta_mll_fwd_bwd (used by bench.py for the per-kernel roofline):
 * enable, run calls, then read the accumulated milliseconds per stage since the last read
 * (stages: 0 mlp_fwd, 1 gp_mll, 2 mlp_bwd, 3 reductions).  Reading synchronises the recorded events.
 */
#define PACOH_NUM_STAGES 4
int pacoh_stage_timing_enable(int32_t on);
int pacoh_stage_timing_read(float* ms_out, int32_t* calls_out);

/*
 * Diagnostics for the large-n path (64 < n <= 4096 points per task: blocked Cholesky / inverse on the tensor cores,
 * csrc/gp_big.cu).  out[14] (host): [0] byte offset of the large-n scratch inside the workspace of
 * pacoh_meta_mll_fwd_bwd (0 and all-zero when (n, F) takes a small-matrix kernel), [1] tiles per side, [2] padded n,
 * [3] matrices per pass, then the byte offsets (relative to [0]) of: [4] L / U tiles (batch, npad, npad), [5] inverted
 * diagonal tiles (batch, nb, 2, 128, 128), [6] scaled features, [7] residuals, [8] v = L^-1 r, [9] alphahat,
 * [10] log2-det partials, [11] gradient partials, [12] per-matrix state, [13] total bytes.  Lets the tests compare the
 * factors themselves with an fp64 Cholesky.
 */
int pacoh_debug_big_layout(const pacoh_arch_t* arch, int32_t P, int32_t T, int32_t n, int64_t* out);

/* FP32 FFMA micro-benchmark used by bench.py for the roofline denominator.  iters > 0: immediate-operand dependent
 * chains (the classic peak test); iters < 0: |iters| iterations of a GEMM-shaped body whose FFMAs read three distinct
 * registers (acc += w * v), the form the kernels actually issue.  `sink` needs >= 20 floats.  The number of
 * floating-point operations issued is returned in *flops_out (host). */
int pacoh_ffma_peak_launch(int32_t iters, float* sink, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PACOH_B200_H */
