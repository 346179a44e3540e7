// General-shape MLP path (any depth <= 8, any widths, any input / output dim): used when a net does not fit the
// register-tiled kernels of mlp.cu (hidden width > 32, input dim > 4, output dim > 4 or more than 4 hidden layers;
// e.g. the 4 x 128 nets of experiments/meta_GPR_mll_base_exp.py:29-30).  Same semantics as mlp.cu
// (meta_learn/models.py:295-317, 343-349 and their reverse pass), laid out for coalescing rather than for peak FLOP/s:
// activations and deltas live in a caller-provided scratch as [particle][feature][point], one thread per point in the
// per-point kernels and one warp per weight entry in the weight-gradient kernel.
#include "common.cuh"
#include "kernels.cuh"

namespace pacoh {

namespace {

__device__ __forceinline__ int sum_widths(const NetDev& n) {
  int s = 0;
  for (int l = 0; l < n.n_hidden; ++l) s += n.width[l];
  return s;
}

__device__ __forceinline__ int gather_src(const MlpArgs& a, int q) {
  const int t = q / a.n;
  return (a.task_idx != nullptr ? __ldg(a.task_idx + t) : t) * a.n + (q - t * a.n);
}

// acts[p][foff_l + j][q] = tanh(b_l[j] + sum_k W_l[j][k] acts[p][foff_{l-1} + k][q]);  out = Wout h_L + bout
__global__ void generic_fwd_kernel(MlpArgs a, int zi, float* __restrict__ acts, bool write_out) {
  const NetDev& net = a.net[zi];
  const int Q = a.T * a.n, p = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float* th = a.theta + (size_t)p * a.D;
  const int SW = sum_widths(net);
  float* A = acts + (size_t)p * SW * Q;
  const int src = gather_src(a, q);
  int foff_prev = 0, foff = 0, win = a.d;
  for (int l = 0; l < net.n_hidden; ++l) {
    const int w = net.width[l];
    for (int j = 0; j < w; ++j) {
      float acc = __ldg(th + net.off_b[l] + j);
      const float* wr = th + net.off_w[l] + (size_t)j * win;
      if (l == 0) for (int k = 0; k < win; ++k) acc = fmaf(__ldg(wr + k), __ldg(a.x + (size_t)src * a.d + k), acc);
      else        for (int k = 0; k < win; ++k) acc = fmaf(__ldg(wr + k), A[(size_t)(foff_prev + k) * Q + q], acc);
      A[(size_t)(foff + j) * Q + q] = tanh_fast(acc);
    }
    foff_prev = foff; foff += w; win = w;
  }
  if (write_out) {
    const int L = net.n_hidden;
    for (int o = 0; o < net.out_dim; ++o) {
      float acc = __ldg(th + net.off_b[L] + o);
      const float* wr = th + net.off_w[L] + (size_t)o * win;
      for (int k = 0; k < win; ++k) acc = fmaf(__ldg(wr + k), A[(size_t)(foff_prev + k) * Q + q], acc);
      a.out[zi][((size_t)p * Q + q) * net.out_dim + o] = acc;
    }
  }
}

// deltas[p][foff_l + j][q] = dL/d(pre-activation); computed top-down, one thread per point.
__global__ void generic_delta_kernel(MlpArgs a, int zi, const float* __restrict__ acts, float* __restrict__ deltas) {
  const NetDev& net = a.net[zi];
  const int Q = a.T * a.n, p = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float* th = a.theta + (size_t)p * a.D;
  const int SW = sum_widths(net), L = net.n_hidden;
  const float* A = acts + (size_t)p * SW * Q;
  float* Dl = deltas + (size_t)p * SW * Q;
  int foffs[kMaxLayers];
  { int f = 0; for (int l = 0; l < L; ++l) { foffs[l] = f; f += net.width[l]; } }
  const float* dout = a.dout[zi] + ((size_t)p * Q + q) * net.out_dim;
  {
    const int w = net.width[L - 1];
    for (int k = 0; k < w; ++k) {
      float s = 0.0f;
      for (int o = 0; o < net.out_dim; ++o) s = fmaf(__ldg(th + net.off_w[L] + (size_t)o * w + k), __ldg(dout + o), s);
      const float h = A[(size_t)(foffs[L - 1] + k) * Q + q];
      Dl[(size_t)(foffs[L - 1] + k) * Q + q] = s * fmaf(-h, h, 1.0f);
    }
  }
  for (int l = L - 1; l >= 1; --l) {   // delta of hidden layer index l-1 from layer index l
    const int w = net.width[l], wprev = net.width[l - 1];
    for (int k = 0; k < wprev; ++k) {
      float s = 0.0f;
      for (int j = 0; j < w; ++j) s = fmaf(__ldg(th + net.off_w[l] + (size_t)j * wprev + k), Dl[(size_t)(foffs[l] + j) * Q + q], s);
      const float h = A[(size_t)(foffs[l - 1] + k) * Q + q];
      Dl[(size_t)(foffs[l - 1] + k) * Q + q] = s * fmaf(-h, h, 1.0f);
    }
  }
}

// One warp per parameter of the net: dot product over the Q points (coalesced rows).
__global__ void generic_wgrad_kernel(MlpArgs a, int zi, const float* __restrict__ acts, const float* __restrict__ deltas) {
  const NetDev& net = a.net[zi];
  const int Q = a.T * a.n, p = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // net-local parameter index
  if (i >= net.total) return;
  const int SW = sum_widths(net), L = net.n_hidden;
  const float* A = acts + (size_t)p * SW * Q;
  const float* Dl = deltas + (size_t)p * SW * Q;
  const int g = net.off_b[0] + i;
  // locate (layer, kind, j, k)
  int layer = 0; bool is_w = false; int j = 0, k = 0;
  for (int l = L; l >= 0; --l) {
    if (g >= net.off_w[l]) { layer = l; is_w = true; const int win = l == 0 ? a.d : net.width[l - 1]; j = (g - net.off_w[l]) / win; k = (g - net.off_w[l]) % win; break; }
    if (g >= net.off_b[l]) { layer = l; is_w = false; j = g - net.off_b[l]; break; }
  }
  // delta rows of hidden layer `layer` start at foff; its inputs (hidden layer layer-1) start at foff_prev
  int foff = 0, foff_prev = 0;
  for (int l = 0; l < layer; ++l) { foff_prev = foff; foff += net.width[l]; }
  float s = 0.0f;
  for (int q = lane; q < Q; q += 32) {
    float dv;
    if (layer == L) dv = __ldg(a.dout[zi] + ((size_t)p * Q + q) * net.out_dim + j);
    else dv = Dl[(size_t)(foff + j) * Q + q];
    float hv = 1.0f;
    if (is_w) {
      if (layer == 0) hv = __ldg(a.x + (size_t)gather_src(a, q) * a.d + k);
      else hv = A[(size_t)(foff_prev + k) * Q + q];
    }
    s = fmaf(dv, hv, s);
  }
  s = warp_sum(s);
  if (lane == 0) a.partial[zi][(size_t)p * net.total + i] = s;   // chunk 0
}

}  // namespace

size_t mlp_generic_scratch_floats(const NetDev& net, int P, int Q) {
  size_t sw = 0;
  for (int l = 0; l < net.n_hidden; ++l) sw += net.width[l];
  return 2 * sw * (size_t)P * Q;
}

int launch_mlp_generic(const MlpArgs& a, int zi, int chunks, bool bwd, float* scratch, cudaStream_t st) {
  const NetDev& net = a.net[zi];
  const int Q = a.T * a.n;
  size_t sw = 0;
  for (int l = 0; l < net.n_hidden; ++l) sw += net.width[l];
  float* acts = scratch;
  float* deltas = scratch + sw * (size_t)a.P * Q;
  dim3 grid((Q + 127) / 128, a.P);
  generic_fwd_kernel<<<grid, 128, 0, st>>>(a, zi, acts, !bwd);
  PACOH_CUDA_CHECK(cudaGetLastError());
  if (bwd) {
    generic_delta_kernel<<<grid, 128, 0, st>>>(a, zi, acts, deltas);
    PACOH_CUDA_CHECK(cudaGetLastError());
    if (chunks > 1)
      PACOH_CUDA_CHECK(cudaMemsetAsync(a.partial[zi] + (size_t)a.P * net.total, 0, sizeof(float) * (size_t)(chunks - 1) * a.P * net.total, st));
    dim3 g2((net.total + 7) / 8, a.P);
    generic_wgrad_kernel<<<g2, 256, 0, st>>>(a, zi, acts, deltas);
    PACOH_CUDA_CHECK(cudaGetLastError());
  }
  return PACOH_OK;
}

}  // namespace pacoh
