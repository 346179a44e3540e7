// Fused SVGD direction: pairwise squared distances, exact median-heuristic bandwidth on device, RBF kernel,
// driving + repulsive term.  Replaces SVGD.phi's tail and RBF_Kernel (meta_learn/svgd.py:18-21, 32-59, 103-107):
//
//   d2_ij = -2 <x_i,x_j> + |x_i|^2 + |x_j|^2          (same association order as norm_sq, diag exactly 0)
//   h     = median(d2 over ALL P*P entries) / (2 log(P+1))      (np.median: mean of the two middle values)
//   gamma = 1 / (1e-8 + 2 h)    [fixed bandwidth b: 1 / (1e-8 + 2 b^2)]
//   phi   = (K s + 2 gamma (rowsum(K) x - K x)) / P,   K = exp(-gamma d2)
//
// The reference syncs with the host for np.median; here the P*P distances are bitonic-sorted in shared memory
// by one CTA, so the whole step stays on the stream (CUDA-graph capturable).
#include "common.cuh"

namespace pacoh {

namespace {

constexpr int kColTile = 32;

// partial[chunk][i][j] = sum_{c in chunk} x_i[c] x_j[c]
__global__ void svgd_gram_partial_kernel(int P, int64_t D, const float* __restrict__ x, float* __restrict__ partial) {
  extern __shared__ float sx[];   // [P][kColTile + 1]
  const int64_t c0 = (int64_t)blockIdx.x * kColTile;
  for (int i = threadIdx.x; i < P * kColTile; i += blockDim.x) {
    const int r = i / kColTile, c = i - r * kColTile;
    sx[r * (kColTile + 1) + c] = (c0 + c < D) ? x[(int64_t)r * D + c0 + c] : 0.0f;
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * P * P;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const int i = e / P, j = e - i * P;
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kColTile; ++c) s = fmaf(sx[i * (kColTile + 1) + c], sx[j * (kColTile + 1) + c], s);
    out[e] = s;
  }
}

// Single CTA: reduce Gram partials, distances, exact median, gamma, kernel matrix and its row sums.
__global__ void svgd_kernel_matrix_kernel(int P, int chunks, const float* __restrict__ partial, float bandwidth,
                                          float* __restrict__ Kmat, float* __restrict__ rowsum, float* __restrict__ gamma_out,
                                          int npow2) {
  extern __shared__ float sm[];   // d2 [P*P] | sort buffer [npow2] | diag [P]
  float* d2 = sm;
  float* srt = sm + P * P;
  float* diag = srt + npow2;
  __shared__ float s_gamma;
  const int N = P * P;
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    float s = 0.0f;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * N + e];
    d2[e] = s;   // Gram for now
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) diag[i] = d2[i * P + i];
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    const int i = e / P, j = e - i * P;
    const float v = (-2.0f * d2[e] + diag[i]) + diag[j];
    srt[e] = v;
  }
  for (int e = N + threadIdx.x; e < npow2; e += blockDim.x) srt[e] = 3.4e38f;
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += blockDim.x) d2[e] = srt[e];
  __syncthreads();
  if (bandwidth <= 0.0f) {
    for (int k = 2; k <= npow2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float a = srt[i], b = srt[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { srt[i] = b; srt[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) {
      const float med = (N & 1) ? srt[(N - 1) / 2] : 0.5f * (srt[N / 2 - 1] + srt[N / 2]);
      const double h = (double)med / (2.0 * log((double)P + 1.0));
      const double bw = sqrt(h);
      s_gamma = (float)(1.0 / (1e-8 + 2.0 * bw * bw));
    }
  } else if (threadIdx.x == 0) {
    s_gamma = (float)(1.0 / (1e-8 + 2.0 * (double)bandwidth * (double)bandwidth));
  }
  __syncthreads();
  const float gamma = s_gamma;
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    const float kv = expf(-gamma * d2[e]);
    d2[e] = kv;
    Kmat[e] = kv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    float s = 0.0f;
    for (int j = 0; j < P; ++j) s += d2[i * P + j];
    rowsum[i] = s;
  }
  if (threadIdx.x == 0) *gamma_out = gamma;
}

// phi[p][c] = (sum_j K[p][j] (s_j[c] - 2 gamma x_j[c]) + 2 gamma rowsum[p] x_p[c]) / P
__global__ void svgd_phi_kernel(int P, int64_t D, const float* __restrict__ x, const float* __restrict__ score,
                                const float* __restrict__ Kmat, const float* __restrict__ rowsum,
                                const float* __restrict__ gamma_p, float* __restrict__ phi) {
  extern __shared__ float sm[];   // K [P*P] | v [P][kColTile]
  float* sK = sm;
  float* sv = sm + P * P;
  const float gamma = *gamma_p;
  const int64_t c0 = (int64_t)blockIdx.x * kColTile;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) sK[e] = Kmat[e];
  for (int i = threadIdx.x; i < P * kColTile; i += blockDim.x) {
    const int r = i / kColTile, c = i - r * kColTile;
    float v = 0.0f;
    if (c0 + c < D) v = fmaf(-2.0f * gamma, x[(int64_t)r * D + c0 + c], score[(int64_t)r * D + c0 + c]);
    sv[i] = v;
  }
  __syncthreads();
  const int c = threadIdx.x & (kColTile - 1);
  const float invP = 1.0f / (float)P;
  for (int p = threadIdx.x / kColTile; p < P; p += blockDim.x / kColTile) {
    float s = 0.0f;
    for (int j = 0; j < P; ++j) s = fmaf(sK[p * P + j], sv[j * kColTile + c], s);
    if (c0 + c < D) {
      const float xv = x[(int64_t)p * D + c0 + c];
      phi[(int64_t)p * D + c0 + c] = fmaf(2.0f * gamma * rowsum[p], xv, s) * invP;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// IMQ Stein kernel (IMQSteinKernel, meta_learn/svgd.py:63-99, alpha = 0.5, beta = -0.5 as GPR_meta_svgd.py:176-177
// constructs it):  K_ij = (alpha + sum_d (x_jd - x_id)^2 / h_d)^beta,  h_d = fixed bandwidth, or the LOWER median
// (torch.median) over the pairs i < j of (x_jd - x_id)^2, divided by log(P + 1).  SVGD.phi differentiates K(X, X.detach())
// with respect to its first argument THROUGH the median (svgd.py:18-19), which adds, for the pair (i*_d < j*_d) whose
// distance is the median of dimension d, a term on particle j*_d only (closed form: oracle.svgd_phi_imq):
//   d sum(K)/d x_md = (2/h_d) sum_i A_im (x_md - x_id) - [m == j*_d] (sum_ij A_ij ns_ijd / h_d^2) 2 (x_j*d - x_i*d) / log(P+1)
//   A = beta base^(beta-1),  phi = (K s - d sum(K)/dx) / P.
constexpr float kImqAlpha = 0.5f, kImqBeta = -0.5f;

// One CTA per parameter dimension: bitonic sort of the P(P-1)/2 pair distances (key = value bits << 32 | pair index).
__global__ void imq_bandwidth_kernel(int P, int64_t D, const float* __restrict__ x, float* __restrict__ h,
                                     int* __restrict__ istar, int* __restrict__ jstar, int npow2) {
  extern __shared__ unsigned long long skey[];   // [npow2] keys, then the column as floats
  float* col = reinterpret_cast<float*>(skey + npow2);
  const int64_t d = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) col[i] = x[(int64_t)i * D + d];
  __syncthreads();
  const int N = P * (P - 1) / 2;
  for (int e = threadIdx.x; e < npow2; e += blockDim.x) skey[e] = ~0ull;
  __syncthreads();
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const int i = e / P, j = e - i * P;
    if (j > i) {
      const float diff = col[j] - col[i];
      const float v = diff * diff;                                  // >= 0: its bit pattern orders like the value
      const int slot = i * P - (i * (i + 1)) / 2 + (j - i - 1);     // row-major position among the pairs i < j
      skey[slot] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)e;
    }
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = skey[i], b = skey[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { skey[i] = b; skey[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const unsigned long long m = skey[(N - 1) / 2];
    const float med = __uint_as_float((unsigned int)(m >> 32));
    const int e = (int)(m & 0xffffffffu);
    h[d] = med / logf((float)P + 1.0f);
    istar[d] = e / P;
    jstar[d] = e - (e / P) * P;
  }
}

__global__ void imq_fill_bandwidth_kernel(int64_t D, float bandwidth, float* __restrict__ h, int* __restrict__ istar,
                                          int* __restrict__ jstar) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) { h[d] = bandwidth; istar[d] = -1; jstar[d] = -1; }
}

// partial[chunk][i][j] = sum_{c in chunk} (x_j[c] - x_i[c])^2 / h_c
__global__ void imq_base_partial_kernel(int P, int64_t D, const float* __restrict__ x, const float* __restrict__ h,
                                        float* __restrict__ partial) {
  extern __shared__ float sx[];   // [P][kColTile + 1] | inv_h [kColTile]
  float* sih = sx + P * (kColTile + 1);
  const int64_t c0 = (int64_t)blockIdx.x * kColTile;
  for (int i = threadIdx.x; i < P * kColTile; i += blockDim.x) {
    const int r = i / kColTile, c = i - r * kColTile;
    sx[r * (kColTile + 1) + c] = (c0 + c < D) ? x[(int64_t)r * D + c0 + c] : 0.0f;
  }
  for (int c = threadIdx.x; c < kColTile; c += blockDim.x) sih[c] = (c0 + c < D) ? 1.0f / h[c0 + c] : 0.0f;
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * P * P;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
    const int i = e / P, j = e - i * P;
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kColTile; ++c) {
      const float diff = sx[j * (kColTile + 1) + c] - sx[i * (kColTile + 1) + c];
      s = fmaf(diff * diff, sih[c], s);
    }
    out[e] = s;
  }
}

// Single CTA: base = alpha + sum of the partials, K = base^beta, A = beta base^(beta - 1), column sums of A.
__global__ void imq_matrix_kernel(int P, int chunks, const float* __restrict__ partial, float* __restrict__ Kmat,
                                  float* __restrict__ Amat, float* __restrict__ colsumA) {
  const int N = P * P;
  for (int e = threadIdx.x; e < N; e += blockDim.x) {
    float s = kImqAlpha;
    for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * N + e];
    const float kv = expf(kImqBeta * logf(s));
    Kmat[e] = kv;
    Amat[e] = kImqBeta * kv / s;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < P; m += blockDim.x) {
    float s = 0.0f;
    for (int i = 0; i < P; ++i) s += Amat[i * P + m];
    colsumA[m] = s;
  }
}

// phi[m][c] = (sum_j K[m][j] s_j[c] - (2/h_c)(colsumA[m] x_m[c] - sum_i A[m][i] x_i[c]) + median term) / P
__global__ void imq_phi_kernel(int P, int64_t D, const float* __restrict__ x, const float* __restrict__ score,
                               const float* __restrict__ Kmat, const float* __restrict__ Amat,
                               const float* __restrict__ colsumA, const float* __restrict__ h,
                               const int* __restrict__ istar, const int* __restrict__ jstar, float* __restrict__ phi) {
  extern __shared__ float sm[];   // K [P*P] | A [P*P] | x [P][kColTile] | s [P][kColTile] | cred [8][kColTile]
  float* sK = sm;
  float* sA = sK + P * P;
  float* sxv = sA + P * P;
  float* ssv = sxv + P * kColTile;
  float* cred = ssv + P * kColTile;
  const int64_t c0 = (int64_t)blockIdx.x * kColTile;
  for (int e = threadIdx.x; e < P * P; e += blockDim.x) { sK[e] = Kmat[e]; sA[e] = Amat[e]; }
  for (int i = threadIdx.x; i < P * kColTile; i += blockDim.x) {
    const int r = i / kColTile, c = i - r * kColTile;
    const bool ok = c0 + c < D;
    sxv[i] = ok ? x[(int64_t)r * D + c0 + c] : 0.0f;
    ssv[i] = ok ? score[(int64_t)r * D + c0 + c] : 0.0f;
  }
  __syncthreads();
  const int c = threadIdx.x & (kColTile - 1), grp = threadIdx.x / kColTile, ngrp = blockDim.x / kColTile;
  // c_c = sum_ij A_ij (x_jc - x_ic)^2 (only needed where the bandwidth is the median)
  float cs = 0.0f;
  const bool med = (c0 + c < D) && jstar[c0 + c] >= 0;
  if (med) {
    for (int i = grp; i < P; i += ngrp) {
      const float xi = sxv[i * kColTile + c];
      for (int j = 0; j < P; ++j) {
        const float diff = sxv[j * kColTile + c] - xi;
        cs = fmaf(sA[i * P + j], diff * diff, cs);
      }
    }
  }
  cred[grp * kColTile + c] = cs;
  __syncthreads();
  float cc = 0.0f;
  for (int g = 0; g < ngrp; ++g) cc += cred[g * kColTile + c];
  const float invP = 1.0f / (float)P;
  if (c0 + c < D) {
    const float hc = h[c0 + c];
    const int js = jstar[c0 + c], is = istar[c0 + c];
    float medterm = 0.0f;
    if (js >= 0) medterm = (cc / (hc * hc)) * 2.0f * (sxv[js * kColTile + c] - sxv[is * kColTile + c]) / logf((float)P + 1.0f);
    for (int m = grp; m < P; m += ngrp) {
      float ks = 0.0f, ax = 0.0f;
      for (int j = 0; j < P; ++j) {
        ks = fmaf(sK[m * P + j], ssv[j * kColTile + c], ks);
        ax = fmaf(sA[m * P + j], sxv[j * kColTile + c], ax);
      }
      float grad = (2.0f / hc) * (colsumA[m] * sxv[m * kColTile + c] - ax);
      if (m == js) grad -= medterm;
      phi[(int64_t)m * D + c0 + c] = (ks - grad) * invP;
    }
  }
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace
}  // namespace pacoh

using namespace pacoh;

extern "C" int64_t pacoh_svgd_workspace_bytes(int32_t P, int64_t D) {
  if (P < 1 || D < 1) return -1;
  const int64_t chunks = (D + kColTile - 1) / kColTile;
  // RBF: partials | K | rowsum.  IMQ: partials | K | A | colsum(A) | h (D) | i* (D) | j* (D).  Sized for either.
  return (int64_t)sizeof(float) * (chunks * P * P + 2 * (int64_t)P * P + P + 3 * D + 64);
}

// Argument checks shared by the three SVGD entry points.
static int svgd_check(const char* fn, int32_t P, int64_t D, int32_t kernel_kind, const void* workspace, int64_t workspace_bytes) {
  if (P < 1 || D < 1 || !workspace) {
    set_error("%s: invalid argument", fn);
    return PACOH_ERR_INVALID;
  }
  if (kernel_kind != PACOH_SVGD_RBF && kernel_kind != PACOH_SVGD_IMQ) {
    set_error("%s: unknown Stein kernel %d", fn, kernel_kind);
    return PACOH_ERR_UNSUPPORTED;
  }
  if (P > 128) {
    set_error("%s: P=%d > 128 particles not supported by the single-CTA median", fn, P);
    return PACOH_ERR_UNSUPPORTED;
  }
  if (kernel_kind == PACOH_SVGD_IMQ && P < 2) {
    set_error("%s: the IMQ kernel needs at least two particles", fn);
    return PACOH_ERR_INVALID;
  }
  if (workspace_bytes < pacoh_svgd_workspace_bytes(P, D)) {
    set_error("%s: workspace too small", fn);
    return PACOH_ERR_WORKSPACE;
  }
  return PACOH_OK;
}

namespace {
struct ImqWs {
  float *partial, *Kmat, *Amat, *colsum, *h;
  int *istar, *jstar;
};
ImqWs imq_carve(void* workspace, int P, int64_t D) {
  const int64_t chunks = (D + kColTile - 1) / kColTile;
  ImqWs w;
  w.partial = (float*)workspace;
  w.Kmat = w.partial + (size_t)chunks * P * P;
  w.Amat = w.Kmat + (size_t)P * P;
  w.colsum = w.Amat + (size_t)P * P;
  w.h = w.colsum + P;
  w.istar = (int*)(w.h + D);
  w.jstar = w.istar + D;
  return w;
}
}  // namespace

// Stage 1 (depends on the particles only): pairwise squared distances, bandwidth (median heuristic), K and what the
// gradient of K needs (RBF: row sums; IMQ: A = beta base^(beta-1), its column sums, h_d and the median pairs).
extern "C" int pacoh_svgd_kernel_matrix(int32_t P, int64_t D, const float* theta, float bandwidth, int32_t kernel_kind,
                                        float* gamma_out, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = svgd_check("pacoh_svgd_kernel_matrix", P, D, kernel_kind, workspace, workspace_bytes);
  if (rc != PACOH_OK) return rc;
  if (!theta || !gamma_out) { set_error("pacoh_svgd_kernel_matrix: invalid argument"); return PACOH_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = (int)((D + kColTile - 1) / kColTile);
  if (kernel_kind == PACOH_SVGD_IMQ) {
    ImqWs w = imq_carve(workspace, P, D);
    if (bandwidth <= 0.0f) {
      const int np2 = next_pow2(P * (P - 1) / 2);
      const size_t smem0 = sizeof(unsigned long long) * np2 + sizeof(float) * P;
      PACOH_CUDA_CHECK(cudaFuncSetAttribute(imq_bandwidth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
      imq_bandwidth_kernel<<<(unsigned)D, 256, smem0, st>>>(P, D, theta, w.h, w.istar, w.jstar, np2);
    } else {
      imq_fill_bandwidth_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(D, bandwidth, w.h, w.istar, w.jstar);
    }
    PACOH_CUDA_CHECK(cudaGetLastError());
    const size_t smem1 = sizeof(float) * (P * (kColTile + 1) + kColTile);
    imq_base_partial_kernel<<<chunks, 256, smem1, st>>>(P, D, theta, w.h, w.partial);
    PACOH_CUDA_CHECK(cudaGetLastError());
    imq_matrix_kernel<<<1, 1024, 0, st>>>(P, chunks, w.partial, w.Kmat, w.Amat, w.colsum);
    PACOH_CUDA_CHECK(cudaGetLastError());
    PACOH_CUDA_CHECK(cudaMemsetAsync(gamma_out, 0, sizeof(float), st));   // no scalar gamma for this kernel
    return PACOH_OK;
  }
  float* partial = (float*)workspace;
  float* Kmat = partial + (size_t)chunks * P * P;
  float* rowsum = Kmat + (size_t)P * P;
  const size_t smem1 = sizeof(float) * P * (kColTile + 1);
  svgd_gram_partial_kernel<<<chunks, 256, smem1, st>>>(P, D, theta, partial);
  PACOH_CUDA_CHECK(cudaGetLastError());
  const int np2 = next_pow2(P * P);
  const size_t smem2 = sizeof(float) * ((size_t)P * P + np2 + P);
  PACOH_CUDA_CHECK(cudaFuncSetAttribute(svgd_kernel_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  svgd_kernel_matrix_kernel<<<1, 1024, smem2, st>>>(P, chunks, partial, bandwidth, Kmat, rowsum, gamma_out, np2);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

// Stage 2: phi from the K / gradient pieces stage 1 left in `workspace` / `gamma` (same particles!).
extern "C" int pacoh_svgd_phi_apply(int32_t P, int64_t D, const float* theta, const float* score, int32_t kernel_kind,
                                    float* phi, const float* gamma, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = svgd_check("pacoh_svgd_phi_apply", P, D, kernel_kind, workspace, workspace_bytes);
  if (rc != PACOH_OK) return rc;
  if (!theta || !score || !phi || !gamma) { set_error("pacoh_svgd_phi_apply: invalid argument"); return PACOH_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = (int)((D + kColTile - 1) / kColTile);
  if (kernel_kind == PACOH_SVGD_IMQ) {
    ImqWs w = imq_carve(workspace, P, D);
    const size_t smem3 = sizeof(float) * (2 * (size_t)P * P + 2 * (size_t)P * kColTile + 8 * kColTile);
    PACOH_CUDA_CHECK(cudaFuncSetAttribute(imq_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    imq_phi_kernel<<<chunks, 256, smem3, st>>>(P, D, theta, score, w.Kmat, w.Amat, w.colsum, w.h, w.istar, w.jstar, phi);
    PACOH_CUDA_CHECK(cudaGetLastError());
    return PACOH_OK;
  }
  const float* Kmat = (const float*)workspace + (size_t)chunks * P * P;
  const float* rowsum = Kmat + (size_t)P * P;
  const size_t smem3 = sizeof(float) * ((size_t)P * P + (size_t)P * kColTile);
  PACOH_CUDA_CHECK(cudaFuncSetAttribute(svgd_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
  svgd_phi_kernel<<<chunks, 256, smem3, st>>>(P, D, theta, score, Kmat, rowsum, gamma, phi);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

// Both stages back to back on one stream.
extern "C" int pacoh_svgd_phi(int32_t P, int64_t D, const float* theta, const float* score, float bandwidth,
                              int32_t kernel_kind, float* phi, float* gamma_out, void* workspace, int64_t workspace_bytes,
                              void* stream) {
  if (!score || !phi) { set_error("pacoh_svgd_phi: invalid argument"); return PACOH_ERR_INVALID; }
  int rc = pacoh_svgd_kernel_matrix(P, D, theta, bandwidth, kernel_kind, gamma_out, workspace, workspace_bytes, stream);
  if (rc != PACOH_OK) return rc;
  return pacoh_svgd_phi_apply(P, D, theta, score, kernel_kind, phi, gamma_out, workspace, workspace_bytes, stream);
}
