// Shared device/host definitions for the PACOH B200 engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pacoh_b200.h"

namespace pacoh {

constexpr int kMaxLayers = PACOH_MAX_LAYERS;
constexpr int kHid = 32;      // hidden width of the register-tiled MLP path
constexpr int kTileP = 32;    // points per warp tile
constexpr int kSRow = 36;     // smem row stride (floats) of a [feature][point] activation tile
constexpr int kMaxOut = 8;    // max MLP output dim on the fast path (mean: 1, kernel features: F)
constexpr int kMaxDin = 8;    // max input dim on the fast path

// One MLP inside a particle's flat parameter row (reference layout: bias before weight, weight (out,in) row-major).
struct NetDev {
  int n_hidden;                 // L
  int width[kMaxLayers];        // hidden widths
  int out_dim;
  int off_b[kMaxLayers + 1];    // layer l = 0..L-1 hidden, l = L output
  int off_w[kMaxLayers + 1];
  int total;                    // number of parameters of this net
};

struct ModelDev {
  int d, F, D;
  int mean_kind, covar_kind, has_oscale;
  float noise_floor;
  NetDev mean, kern;
  int off_const_mean, off_ls, off_noise, off_oscale;
};

// Host: derive offsets from the public arch descriptor.  Returns false if the descriptor is malformed.
bool build_model(const pacoh_arch_t* a, ModelDev* m);
// Fast register-tiled path: every hidden width == 32, 1 <= L <= 2, d <= kMaxDin, out <= kMaxOut.
bool net_is_fast(const NetDev& n, int d);

void set_error(const char* fmt, ...);

#define PACOH_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      pacoh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PACOH_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tanh with ~1e-7 absolute error: 1 - 2 / (exp(2x) + 1); saturates cleanly at +-1.
__device__ __forceinline__ float tanh_fast(float x) {
  float t = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, t + 1.0f);
}

// torch.nn.functional.softplus (beta=1, threshold=20) and its derivative.
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sts4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace pacoh
