// Tensor-core (tcgen05 / TMEM) forward pass of the per-particle MLPs, sm_100a.
//
// Same semantics as mlp_fwd_kernel in mlp.cu (meta_learn/models.py:295-317, 343-349), but the hidden 32x32 layers run as
// UMMA instructions issued by one thread per CTA:
//
//   tile   = 128 points of the flattened task batch; thread t of the CTA owns point t for the elementwise work
//   A      = H_{l-1} tile (128 x 32): each thread writes its row into ITS TENSOR-MEMORY LANE with tcgen05.st (hi / lo parts),
//            so activations never touch shared memory (no proxy fence, no bank conflicts)
//   B      = W_l (32 x 32, K-major SWIZZLE_NONE core-matrix layout in shared memory) staged once per CTA
//   D      = 128 lanes x 32 fp32 columns in tensor memory; read back with tcgen05.ld (lane == point), bias + tanh in registers
//
// Precision: the north-star asks for 1e-4 relative parity on gradients in fp32, which plain TF32 (10-bit mantissa)
// cannot give.  Every operand is therefore split into hi = tf32(v) and lo = v - hi and the product is accumulated as
// lo*hi + hi*lo + hi*hi in fp32 (3xTF32): ~4e-7 relative error measured against fp64 (tools/ubench/tc_ts_test.cu),
// at 12 tcgen05.mma instructions per layer and tile (~17.6 cycles each), and it frees the CUDA cores for tanh /
// splitting / the output layer.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace pacoh {

namespace {

using namespace tc;

constexpr int kTcThreads = 128;
constexpr int kTcTile = 128;
constexpr int kFwdMinSmemBytes = 50 * 1024;   // caps residency at 4 CTAs/SM = 4 x 128 TMEM columns (the whole tensor memory); 4 x (50 + 1) KB fit, 5 do not

template <int L, int DIN, int OUT>
struct TcSmem {
  // floats
  static constexpr int B = 0;                                      // per hidden layer l = 2..L: W hi [32x32], W lo [32x32]
  static constexpr int W1 = B + (L - 1) * 2 * kHid * kHid;         // [32][DIN]
  static constexpr int B1 = W1 + kHid * DIN;
  static constexpr int BH = B1 + kHid;                             // biases of layers 2..L
  static constexpr int WOUT = BH + (L - 1) * kHid;                 // [OUT][32]
  static constexpr int BOUT = WOUT + OUT * kHid;
  static constexpr int END = BOUT + 4;
};

template <int L, int DIN, int OUT>
__global__ void __launch_bounds__(kTcThreads) mlp_tc_fwd_kernel(MlpArgs a) {
  using S = TcSmem<L, DIN, OUT>;
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const NetDev& net = a.net[blockIdx.z];
  const int p = blockIdx.y, tid = threadIdx.x, warp = tid >> 5;
  const float* th = a.theta + (size_t)p * a.D;

  if (warp == 0) tmem_alloc<kTmemCols>(&tmem_base_s);
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  // ---- stage the particle's weights: hidden-layer matrices as K-major UMMA B operands (hi / lo), the rest as plain arrays
  const int w0 = net.width[0];
  for (int i = tid; i < kHid * DIN; i += kTcThreads) {
    const int j = i / DIN, dd = i - j * DIN;
    smem[S::W1 + i] = (j < w0 && dd < a.d) ? th[net.off_w[0] + j * a.d + dd] : 0.0f;
  }
  for (int i = tid; i < kHid; i += kTcThreads) smem[S::B1 + i] = i < w0 ? th[net.off_b[0] + i] : 0.0f;
#pragma unroll
  for (int l = 2; l <= L; ++l) {
    const int win = net.width[l - 2], wout = net.width[l - 1];
    float* bh = smem + S::B + (l - 2) * 2 * kHid * kHid;
    for (int i = tid; i < kHid * kHid; i += kTcThreads) {
      const int j = i >> 5, k = i & 31;
      const float w = (j < wout && k < win) ? th[net.off_w[l - 1] + j * win + k] : 0.0f;
      const float h = tf32_hi(w);
      bh[ktile_off(j, k, kHid)] = h;
      bh[kHid * kHid + ktile_off(j, k, kHid)] = w - h;
    }
    for (int i = tid; i < kHid; i += kTcThreads) smem[S::BH + (l - 2) * kHid + i] = i < wout ? th[net.off_b[l - 1] + i] : 0.0f;
  }
  const int wl = net.width[L - 1];
  for (int i = tid; i < OUT * kHid; i += kTcThreads) {
    const int o = i >> 5, k = i & 31;
    smem[S::WOUT + i] = (o < net.out_dim && k < wl) ? th[net.off_w[L] + o * wl + k] : 0.0f;
  }
  if (tid < 4) smem[S::BOUT + tid] = tid < net.out_dim ? th[net.off_b[L] + tid] : 0.0f;
  fence_async_smem();            // generic-proxy writes of the B operands -> visible to the tensor-core proxy
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_s, 0);   // provably warp-uniform: MMA operands stay in uniform registers
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
  const uint32_t bar = smem_u32(&mbar);
  uint32_t parity = 0;

  const int Q = a.T * a.n;
  const int tiles = (Q + kTcTile - 1) / kTcTile;
  // balanced split: chunk sizes differ by at most one tile
  const int t0 = (int)(((long long)tiles * blockIdx.x) / gridDim.x), t1 = (int)(((long long)tiles * (blockIdx.x + 1)) / gridDim.x);
  float* outp = a.out[blockIdx.z] + (size_t)p * Q * net.out_dim;

  // inputs of the NEXT tile are loaded one tile ahead (global latency hidden behind the current tile)
  auto load_x = [&](int tile_i, float (&xo)[DIN]) {
    const int qq = tile_i * kTcTile + tid;
    const bool ok = tile_i < t1 && qq < Q;
    int src = 0;
    if (ok) {
      const int t = qq / a.n;
      src = (a.task_idx != nullptr ? __ldg(a.task_idx + t) : t) * a.n + (qq - t * a.n);
    }
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) xo[dd] = (ok && dd < a.d) ? __ldg(a.x + (size_t)src * a.d + dd) : 0.0f;
  };
  float xn[DIN];
  load_x(t0, xn);

  for (int tile = t0; tile < t1; ++tile) {
    const int q = tile * kTcTile + tid;
    const bool valid = q < Q;
    float x[DIN];
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) x[dd] = xn[dd];
    load_x(tile + 1, xn);
    // ---- layer 1 (input dim <= 4): registers only
    float h[kHid];
#pragma unroll
    for (int j4 = 0; j4 < kHid; j4 += 4) {
      const float4 b = lds4(smem + S::B1 + j4);
      float acc[4] = {b.x, b.y, b.z, b.w};
      float wv[4 * DIN];                            // the 4 features' first-layer weights: DIN vector loads, not 4 DIN scalar ones
#pragma unroll
      for (int qv = 0; qv < DIN; ++qv) {
        const float4 t4 = lds4(smem + S::W1 + (j4) * DIN + 4 * qv);
        wv[4 * qv] = t4.x; wv[4 * qv + 1] = t4.y; wv[4 * qv + 2] = t4.z; wv[4 * qv + 3] = t4.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int dd = 0; dd < DIN; ++dd) acc[e] = fmaf(wv[e * DIN + dd], x[dd], acc[e]);
        h[j4 + e] = tanh_fast(acc[e]);
      }
    }
    // ---- hidden layers 2..L on the tensor cores
#pragma unroll
    for (int l = 2; l <= L; ++l) {
      store_a_tmem(lane_base, h);          // this point's activations -> its TMEM lane (hi / lo)
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint32_t b_hi = smem_u32(smem + S::B + (l - 2) * 2 * kHid * kHid);
        gemm128x32x32_3xtf32_ts(tmem, b_hi, b_hi + kHid * kHid * 4, bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      fence_after_sync();
      uint32_t v[32];
      tmem_ld32(lane_base + kTmemD, v);
#pragma unroll
      for (int j4 = 0; j4 < kHid; j4 += 4) {
        const float4 b = lds4(smem + S::BH + (l - 2) * kHid + j4);
        h[j4] = tanh_fast(__uint_as_float(v[j4]) + b.x);
        h[j4 + 1] = tanh_fast(__uint_as_float(v[j4 + 1]) + b.y);
        h[j4 + 2] = tanh_fast(__uint_as_float(v[j4 + 2]) + b.z);
        h[j4 + 3] = tanh_fast(__uint_as_float(v[j4 + 3]) + b.w);
      }
    }
    // ---- output layer in registers
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      float acc = smem[S::BOUT + o];
#pragma unroll
      for (int j4 = 0; j4 < kHid; j4 += 4) {
        const float4 w = lds4(smem + S::WOUT + o * kHid + j4);
        acc = fmaf(w.x, h[j4], acc); acc = fmaf(w.y, h[j4 + 1], acc); acc = fmaf(w.z, h[j4 + 2], acc); acc = fmaf(w.w, h[j4 + 3], acc);
      }
      if (valid && o < net.out_dim) outp[(size_t)q * net.out_dim + o] = acc;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

template <int L, int DIN, int OUT>
int launch_tc(const MlpArgs& a, int chunks, int nets, cudaStream_t st) {
  using S = TcSmem<L, DIN, OUT>;
  size_t smem = sizeof(float) * S::END;
  if (smem < (size_t)kFwdMinSmemBytes) smem = kFwdMinSmemBytes;
  PACOH_CUDA_CHECK(cudaFuncSetAttribute(mlp_tc_fwd_kernel<L, DIN, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(chunks, a.P, nets);
  mlp_tc_fwd_kernel<L, DIN, OUT><<<grid, kTcThreads, smem, st>>>(a);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

template <int L, int DIN>
int tc_dispatch_out(const MlpArgs& a, int out_pad, int chunks, int nets, cudaStream_t st) {
  switch (out_pad) {
    case 1: return launch_tc<L, DIN, 1>(a, chunks, nets, st);
    case 2: return launch_tc<L, DIN, 2>(a, chunks, nets, st);
    case 4: return launch_tc<L, DIN, 4>(a, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

template <int L>
int tc_dispatch_din(const MlpArgs& a, int din_pad, int out_pad, int chunks, int nets, cudaStream_t st) {
  switch (din_pad) {
    case 1: return tc_dispatch_out<L, 1>(a, out_pad, chunks, nets, st);
    case 2: return tc_dispatch_out<L, 2>(a, out_pad, chunks, nets, st);
    case 4: return tc_dispatch_out<L, 4>(a, out_pad, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

int tc_pad_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : 4); }

}  // namespace

// Forward pass of the `nets` nets in a.net[] (same depth, widths <= 32, d <= 4, out <= 4) on tcgen05.
int launch_mlp_tc_fwd(const MlpArgs& a, int nets, int chunks, cudaStream_t st) {
  const int din = tc_pad_pow2(a.d);
  int out_pad = 1;
  for (int z = 0; z < nets; ++z) out_pad = max(out_pad, tc_pad_pow2(a.net[z].out_dim));
  switch (a.net[0].n_hidden) {
    case 1: return tc_dispatch_din<1>(a, din, out_pad, chunks, nets, st);
    case 2: return tc_dispatch_din<2>(a, din, out_pad, chunks, nets, st);
    case 3: return tc_dispatch_din<3>(a, din, out_pad, chunks, nets, st);
    case 4: return tc_dispatch_din<4>(a, din, out_pad, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

}  // namespace pacoh
