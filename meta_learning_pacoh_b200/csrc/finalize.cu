// Deterministic reductions and the small elementwise kernels around the batched MLL:
//   * sum of per-CTA parameter-gradient partials              (second stage of the mlp_bwd reduction)
//   * sum over tasks of mll and of the GP hyper-parameter gradients
//   * hyper-prior log-density + gradient and the final combination   (random_gp.py:179-180, 221-222;
//                                                                     CatDist.log_prob models.py:159-181)
//   * diagonal-Gaussian VI reparameterised sample / log q / gradient epilogue
//                                                                    (random_gp.py:244-263, GPR_meta_vi.py:216-224)
#include "common.cuh"
#include "kernels.cuh"

namespace pacoh {

namespace {

// dtheta[p, dst_off + i] = sum_c partial[c, p, i]    (fixed summation order => bitwise repeatable)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int chunks, int P, int total,
                                       float* __restrict__ dtheta, int D, int dst_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (i >= total) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += partial[((size_t)c * P + p) * total + i];
  dtheta[(size_t)p * D + dst_off + i] = s;
}

// One CTA per particle: mll_sum[p] = sum_t mll[p,t]; hyper-parameter gradients summed over tasks.
__global__ void reduce_hyp_kernel(GpArgs a, float* __restrict__ dtheta, float* __restrict__ mll_sum) {
  __shared__ float red[32];
  const int p = blockIdx.x;
  const int H = a.F + 3;
  const float* hyp = a.dhyp + (size_t)p * a.T * H;
  for (int c = -1; c < H; ++c) {
    float s = 0.0f;
    if (c < 0) for (int t = threadIdx.x; t < a.T; t += blockDim.x) s += a.mll[(size_t)p * a.T + t];
    else       for (int t = threadIdx.x; t < a.T; t += blockDim.x) s += hyp[(size_t)t * H + c];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
      v = warp_sum(v);
      if (threadIdx.x == 0) {
        float* row = dtheta + (size_t)p * a.D;
        if (c < 0) mll_sum[p] = v;
        else if (c < a.F) row[a.off_ls + c] = v;
        else if (c == a.F) row[a.off_noise] = v;
        else if (c == a.F + 1) { if (a.has_oscale) row[a.off_oscale] = v; }
        else if (a.mean_kind == PACOH_MEAN_CONSTANT) row[a.off_const_mean] = v;
      }
    }
    __syncthreads();
  }
}

__global__ void logprob_finalize_kernel(int P, int64_t D, const float* __restrict__ theta, const float* __restrict__ mu,
                                        const float* __restrict__ sigma, float prior_factor, float pre_factor,
                                        const float* __restrict__ mll_sum, const float* __restrict__ dlik,
                                        float* __restrict__ logp, float* __restrict__ dtheta) {
  __shared__ float red[32];
  const int p = blockIdx.x;
  float lp = 0.0f;
  for (int64_t k = threadIdx.x; k < D; k += blockDim.x) {
    const float s = sigma[k];
    const float zc = (theta[p * D + k] - mu[k]) / s;
    lp += -0.5f * zc * zc - logf(s) - 0.91893853320467274178f;
    if (dtheta != nullptr) dtheta[p * D + k] = fmaf(pre_factor, dlik[p * D + k], -prior_factor * zc / s);
  }
  lp = warp_sum(lp);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && logp != nullptr) logp[p] = fmaf(prior_factor, v, pre_factor * mll_sum[p]);
  }
}

// ---- all-reduce over NVLink peer memory FUSED into the finalize step (task-sharded runs) ---------------------------
// Every rank's packed likelihood buffer (dtheta_lik (P, D) | mll_sum (P)) lives in symmetric memory that all ranks of
// the node have mapped.  One kernel per rank: announce "my buffer for step `token` is complete" by writing the token
// into slot `rank` of every peer's flag array (system-scope release), wait until every peer has announced to us
// (acquire), then read all `world` buffers directly over NVLink, sum them in rank order (bitwise identical on every
// rank) and apply the hyper-prior / pre-factor on the fly.  No NCCL launch, no extra pass over the data; the buffers are
// double-buffered by step parity on the host side, which together with the monotone tokens makes reuse safe.
struct PeerPtrs {
  const float* buf[PACOH_MAX_PEERS];
  unsigned int* flag[PACOH_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {      // peer data: never through the (incoherent) L1
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// `token_dev` (optional): the step counter of a CUDA-graph-captured training loop (pacoh_step_prepare); the token of this
// call is then *token_dev + token (token = offset).  The wait is bounded: a peer that never announces (a rank that died or
// raised) trips `*err_flag` after 30 s instead of hanging the node; the sums are then garbage and the
// host raises on its next check.
__global__ void peer_sum_finalize_kernel(PeerPtrs pp, int world, int rank, unsigned int token, const int* __restrict__ token_dev,
                                         int* __restrict__ err_flag, int P, int64_t D,
                                         const float* __restrict__ theta, const float* __restrict__ mu,
                                         const float* __restrict__ sigma, float prior_factor, float pre_factor,
                                         float* __restrict__ logp, float* __restrict__ dtheta) {
  __shared__ float red[32];
  if (token_dev != nullptr) token += (unsigned int)*token_dev;
  if (threadIdx.x < world) {
    if (blockIdx.x == 0) {
      __threadfence_system();                                     // the kernels before us on this stream wrote the buffer
      st_release_sys(pp.flag[threadIdx.x] + rank, token);
    }
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(ld_acquire_sys(pp.flag[rank] + threadIdx.x) - token) < 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 30000000000ull) { if (err_flag != nullptr) atomicExch(err_flag, 1 + (int)threadIdx.x); break; }   // 30 s
    }
  }
  __syncthreads();
  const int p = blockIdx.x;
  float lp = 0.0f;
  for (int64_t k = threadIdx.x; k < D; k += blockDim.x) {
    float v[PACOH_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < PACOH_MAX_PEERS; ++r) v[r] = r < world ? ld_peer(pp.buf[r] + p * D + k) : 0.0f;   // all in flight
    float dl = v[0];
#pragma unroll
    for (int r = 1; r < PACOH_MAX_PEERS; ++r) dl += v[r];                                                  // rank order
    const float s = sigma[k];
    const float zc = (theta[p * D + k] - mu[k]) / s;
    lp += -0.5f * zc * zc - logf(s) - 0.91893853320467274178f;
    dtheta[p * D + k] = fmaf(pre_factor, dl, -prior_factor * zc / s);
  }
  lp = warp_sum(lp);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      float ms = 0.0f;
      for (int r = 0; r < world; ++r) ms += ld_peer(pp.buf[r] + (int64_t)P * D + p);
      logp[p] = fmaf(prior_factor, v, pre_factor * ms);
    }
  }
}

__global__ void vi_sample_kernel(int S, int64_t D, const float* __restrict__ loc, const float* __restrict__ scale,
                                 const float* __restrict__ eps, float* __restrict__ theta, float* __restrict__ logq) {
  __shared__ float red[32];
  const int s = blockIdx.x;
  float lq = 0.0f;
  for (int64_t k = threadIdx.x; k < D; k += blockDim.x) {
    const float e = eps[s * D + k], sc = scale[k];
    theta[s * D + k] = fmaf(expf(sc), e, loc[k]);
    lq += -0.5f * e * e - sc - 0.91893853320467274178f;
  }
  lq = warp_sum(lq);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    v = warp_sum(v);
    if (threadIdx.x == 0 && logq != nullptr) logq[s] = v;
  }
}

__global__ void vi_grad_kernel(int S, int64_t D, const float* __restrict__ scale, const float* __restrict__ eps,
                               const float* __restrict__ g, float prior_factor, float* __restrict__ dloc,
                               float* __restrict__ dscale) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= D) return;
  float gl = 0.0f, gs = 0.0f;
  for (int s = 0; s < S; ++s) {
    const float gv = g[s * D + k];
    gl += gv;
    gs = fmaf(gv, eps[s * D + k], gs);
  }
  const float invS = 1.0f / (float)S;
  dloc[k] = -gl * invS;
  dscale[k] = -gs * invS * expf(scale[k]) - prior_factor;
}

// torch.optim.Adam (single tensor, no amsgrad, no weight decay): same operation order as torch/optim/adam.py.
__global__ void adam_kernel(int64_t count, float* __restrict__ p, const float* __restrict__ g, float gsign,
                            float* __restrict__ m, float* __restrict__ v, float beta1, float beta2, float eps,
                            float step_size, float inv_sqrt_bc2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float gr = gsign * g[i];
  const float mi = m[i] + (gr - m[i]) * (1.0f - beta1);          // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gr * gr);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] -= step_size * (mi / denom);
}

// Same update with the step-dependent scalars read from the device-side step state (CUDA-graph-captured loops).
__global__ void adam_dev_kernel(int64_t count, float* __restrict__ p, const float* __restrict__ g, float gsign,
                                float* __restrict__ m, float* __restrict__ v, float beta1, float beta2, float eps,
                                const float* __restrict__ state) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float step_size = state[3], inv_sqrt_bc2 = state[4];
  const float gr = gsign * g[i];
  const float mi = m[i] + (gr - m[i]) * (1.0f - beta1);
  const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gr * gr);
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] -= step_size * (mi / denom);
}

// torch.optim.AdamW (decoupled weight decay: p *= 1 - lr wd before the Adam update) with the step-dependent scalars on the
// device; `mask` (optional) freezes entries (PACOH-MAP learning modes that train only the mean or only the kernel).
__global__ void adamw_dev_kernel(int64_t count, float* __restrict__ p, const float* __restrict__ g, float gsign,
                                 float* __restrict__ m, float* __restrict__ v, float beta1, float beta2, float eps, float wd,
                                 const unsigned char* __restrict__ mask, const float* __restrict__ state) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count || (mask != nullptr && mask[i] == 0)) return;
  const float lr = state[2], step_size = state[3], inv_sqrt_bc2 = state[4];
  const float gr = gsign * g[i];
  const float pi = p[i] * (1.0f - lr * wd);
  const float mi = m[i] + (gr - m[i]) * (1.0f - beta1);
  const float vi = fmaf(beta2, v[i], (1.0f - beta2) * gr * gr);
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] = pi - step_size * (mi / denom);
}

// One block: advances the step counter, derives the learning rate (StepLR) and Adam's bias corrections for the new step,
// and copies slot (old step mod K) of the pre-uploaded index / float streams into the fixed buffers the step's kernels read.
__global__ void step_prepare_kernel(int* __restrict__ state, int K, int T, const int* __restrict__ idx_stream, int* __restrict__ idx_out,
                                    int64_t F, const float* __restrict__ fstream, float* __restrict__ fout,
                                    float lr0, float gamma, int decay_every, float beta1, float beta2) {
  const int s = state[0];
  __syncthreads();
  const int slot = K > 0 ? s % K : 0;
  if (threadIdx.x == 0) {
    const int step = s + 1;
    state[0] = step;
    double lr = (double)lr0;
    if (decay_every > 0 && gamma != 1.0f) lr *= pow((double)gamma, (double)((step - 1) / decay_every));
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    float* fs = reinterpret_cast<float*>(state);
    fs[2] = (float)lr;
    fs[3] = (float)(lr / bc1);
    fs[4] = (float)(1.0 / sqrt(bc2));
  }
  if (idx_stream != nullptr)
    for (int t = threadIdx.x; t < T; t += blockDim.x) idx_out[t] = idx_stream[(size_t)slot * T + t];
  if (fstream != nullptr)
    for (int64_t i = threadIdx.x; i < F; i += blockDim.x) fout[i] = fstream[(size_t)slot * F + i];
}

}  // namespace

int launch_reduce_partials(const float* partial, int chunks, int P, int total, float* dtheta, int D, int dst_off,
                           cudaStream_t st) {
  dim3 grid((total + 255) / 256, P);
  reduce_partials_kernel<<<grid, 256, 0, st>>>(partial, chunks, P, total, dtheta, D, dst_off);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

int launch_reduce_hyp(const GpArgs& a, float* dtheta, float* mll_sum, cudaStream_t st) {
  reduce_hyp_kernel<<<a.P, 256, 0, st>>>(a, dtheta, mll_sum);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

}  // namespace pacoh

using namespace pacoh;

extern "C" int pacoh_logprob_finalize(int32_t P, int64_t D, const float* theta, const float* prior_mu,
                                      const float* prior_sigma, float prior_factor, float pre_factor,
                                      const float* mll_sum, const float* dtheta_lik, float* logp, float* dtheta,
                                      void* stream) {
  if (P < 1 || D < 1 || !theta || !prior_mu || !prior_sigma || !mll_sum || (dtheta && !dtheta_lik)) {
    set_error("pacoh_logprob_finalize: invalid argument");
    return PACOH_ERR_INVALID;
  }
  logprob_finalize_kernel<<<P, 256, 0, (cudaStream_t)stream>>>(P, D, theta, prior_mu, prior_sigma, prior_factor, pre_factor,
                                                               mll_sum, dtheta_lik, logp, dtheta);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_peer_allreduce_finalize_dev(int32_t world, int32_t rank, const void* const* peer_bufs, void* const* peer_flags,
                                                 uint32_t token, const void* token_dev, int32_t* err_flag, int32_t P, int64_t D,
                                                 const float* theta, const float* prior_mu, const float* prior_sigma,
                                                 float prior_factor, float pre_factor, float* logp, float* dtheta, void* stream) {
  if (world < 1 || world > PACOH_MAX_PEERS || rank < 0 || rank >= world || !peer_bufs || !peer_flags || P < 1 || D < 1 ||
      !theta || !prior_mu || !prior_sigma || !logp || !dtheta) {
    set_error("pacoh_peer_allreduce_finalize: invalid argument");
    return PACOH_ERR_INVALID;
  }
  PeerPtrs pp;
  for (int r = 0; r < PACOH_MAX_PEERS; ++r) {
    pp.buf[r] = r < world ? (const float*)peer_bufs[r] : nullptr;
    pp.flag[r] = r < world ? (unsigned int*)peer_flags[r] : nullptr;
    if (r < world && (!pp.buf[r] || !pp.flag[r])) { set_error("pacoh_peer_allreduce_finalize: null peer pointer"); return PACOH_ERR_INVALID; }
  }
  peer_sum_finalize_kernel<<<P, 1024, 0, (cudaStream_t)stream>>>(pp, world, rank, token, (const int*)token_dev, err_flag, P, D, theta,
                                                                prior_mu, prior_sigma, prior_factor, pre_factor, logp, dtheta);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_peer_allreduce_finalize(int32_t world, int32_t rank, const void* const* peer_bufs, void* const* peer_flags,
                                             uint32_t token, int32_t P, int64_t D, const float* theta, const float* prior_mu,
                                             const float* prior_sigma, float prior_factor, float pre_factor, float* logp,
                                             float* dtheta, void* stream) {
  return pacoh_peer_allreduce_finalize_dev(world, rank, peer_bufs, peer_flags, token, nullptr, nullptr, P, D, theta, prior_mu,
                                           prior_sigma, prior_factor, pre_factor, logp, dtheta, stream);
}

extern "C" int pacoh_vi_sample(int32_t S, int64_t D, const float* loc, const float* scale, const float* eps, float* theta,
                               float* logq, void* stream) {
  if (S < 1 || D < 1 || !loc || !scale || !eps || !theta) {
    set_error("pacoh_vi_sample: invalid argument");
    return PACOH_ERR_INVALID;
  }
  vi_sample_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(S, D, loc, scale, eps, theta, logq);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_vi_grad(int32_t S, int64_t D, const float* scale, const float* eps, const float* g, float prior_factor,
                             float* dloc, float* dscale, void* stream) {
  if (S < 1 || D < 1 || !scale || !eps || !g || !dloc || !dscale) {
    set_error("pacoh_vi_grad: invalid argument");
    return PACOH_ERR_INVALID;
  }
  vi_grad_kernel<<<(unsigned)((D + 255) / 256), 256, 0, (cudaStream_t)stream>>>(S, D, scale, eps, g, prior_factor, dloc, dscale);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_adam_step(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg,
                               float* exp_avg_sq, float lr, float beta1, float beta2, float eps, int64_t step, void* stream) {
  if (count < 1 || !param || !grad || !exp_avg || !exp_avg_sq || step < 1) {
    set_error("pacoh_adam_step: invalid argument");
    return PACOH_ERR_INVALID;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      count, param, grad, grad_sign, exp_avg, exp_avg_sq, beta1, beta2, eps, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)));
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_step_prepare(void* state, int32_t K, int32_t T, const int32_t* idx_stream, int32_t* idx_out, int64_t F,
                                  const float* fstream, float* fout, float lr0, float gamma, int32_t decay_every, float beta1,
                                  float beta2, void* stream) {
  if (!state || K < 0 || T < 0 || F < 0 || (idx_stream && (!idx_out || K < 1)) || (fstream && (!fout || K < 1))) {
    set_error("pacoh_step_prepare: invalid argument");
    return PACOH_ERR_INVALID;
  }
  step_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((int*)state, K, T, idx_stream, idx_out, F, fstream, fout, lr0, gamma,
                                                           decay_every, beta1, beta2);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_adam_step_dev(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg,
                                   float* exp_avg_sq, float beta1, float beta2, float eps, const void* state, void* stream) {
  if (count < 1 || !param || !grad || !exp_avg || !exp_avg_sq || !state) {
    set_error("pacoh_adam_step_dev: invalid argument");
    return PACOH_ERR_INVALID;
  }
  adam_dev_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, param, grad, grad_sign, exp_avg, exp_avg_sq,
                                                                                    beta1, beta2, eps, (const float*)state);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

extern "C" int pacoh_adamw_step_dev(int64_t count, float* param, const float* grad, float grad_sign, float* exp_avg,
                                    float* exp_avg_sq, float beta1, float beta2, float eps, float weight_decay,
                                    const uint8_t* mask, const void* state, void* stream) {
  if (count < 1 || !param || !grad || !exp_avg || !exp_avg_sq || !state) {
    set_error("pacoh_adamw_step_dev: invalid argument");
    return PACOH_ERR_INVALID;
  }
  adamw_dev_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, param, grad, grad_sign, exp_avg, exp_avg_sq,
                                                                                     beta1, beta2, eps, weight_decay, mask,
                                                                                     (const float*)state);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}
