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
  return (int64_t)sizeof(float) * (chunks * P * P + (int64_t)P * P + P + 64);
}

// Argument checks shared by the three SVGD entry points.
static int svgd_check(const char* fn, int32_t P, int64_t D, int32_t kernel_kind, const void* workspace, int64_t workspace_bytes) {
  if (P < 1 || D < 1 || !workspace) {
    set_error("%s: invalid argument", fn);
    return PACOH_ERR_INVALID;
  }
  if (kernel_kind != PACOH_SVGD_RBF) {
    set_error("%s: only the RBF Stein kernel is implemented (IMQ: SURVEY 8(f).3)", fn);
    return PACOH_ERR_UNSUPPORTED;
  }
  if (P > 128) {
    set_error("%s: P=%d > 128 particles not supported by the single-CTA median", fn, P);
    return PACOH_ERR_UNSUPPORTED;
  }
  if (workspace_bytes < pacoh_svgd_workspace_bytes(P, D)) {
    set_error("%s: workspace too small", fn);
    return PACOH_ERR_WORKSPACE;
  }
  return PACOH_OK;
}

// Stage 1 (depends on the particles only): pairwise squared distances, median-heuristic bandwidth, K and its row sums.
extern "C" int pacoh_svgd_kernel_matrix(int32_t P, int64_t D, const float* theta, float bandwidth, int32_t kernel_kind,
                                        float* gamma_out, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = svgd_check("pacoh_svgd_kernel_matrix", P, D, kernel_kind, workspace, workspace_bytes);
  if (rc != PACOH_OK) return rc;
  if (!theta || !gamma_out) { set_error("pacoh_svgd_kernel_matrix: invalid argument"); return PACOH_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = (int)((D + kColTile - 1) / kColTile);
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

// Stage 2: phi = (K score + 2 gamma (rowsum(K) * theta - K theta)) / P from the K / row sums / gamma stage 1 left in
// `workspace` / `gamma` (same particles!).
extern "C" int pacoh_svgd_phi_apply(int32_t P, int64_t D, const float* theta, const float* score, int32_t kernel_kind,
                                    float* phi, const float* gamma, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = svgd_check("pacoh_svgd_phi_apply", P, D, kernel_kind, workspace, workspace_bytes);
  if (rc != PACOH_OK) return rc;
  if (!theta || !score || !phi || !gamma) { set_error("pacoh_svgd_phi_apply: invalid argument"); return PACOH_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = (int)((D + kColTile - 1) / kColTile);
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
