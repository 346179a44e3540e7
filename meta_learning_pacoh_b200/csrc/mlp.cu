// Per-particle MLP (mean net / kernel-feature net) forward and recompute-backward kernels, sm_100a.
//
// Replaces LinearVectorized.forward / NeuralNetworkVectorized.forward (meta_learn/models.py:295-317, 343-349)
// and the autograd reverse pass through them (svgd.py:16).  The reference evaluates the nets once per task
// with torch.bmm over P copies of the task; here the sampled task batch is flattened to Q = T*n points and a
// CTA owns (chunk of 32-point tiles, particle, net): the particle's weights are staged once in shared memory
// and every warp streams its own tiles, so activations never touch HBM.
//
// Register tiling (per warp, 32 points x 32 features): lane = (pb = lane & 7, ob = lane >> 3) owns points
// 4pb..4pb+3 and features 8ob..8ob+7.  Activation tiles live in shared memory as [feature][point] with a
// row stride of 36 floats, so every operand fetch is one conflict-free LDS.128.  Hidden widths below 32 are
// zero-padded to 32 (padded units are exactly 0 and receive exactly 0 gradient).
//
// Backward recomputes the forward (cheaper than spilling 32 L floats per point to HBM) and keeps the
// parameter-gradient accumulators in registers across all tiles of the CTA; one deterministic in-CTA
// reduction at the end writes a (chunk, particle) partial that reduce_partials_kernel sums.
#include "common.cuh"
#include "kernels.cuh"

namespace pacoh {

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTileF = kHid * kSRow;   // floats of one [32][36] activation tile

template <int L, int DIN, int OUT>
struct SmemLayout {
  // CTA-shared weights (floats), all hidden widths padded to 32
  static constexpr int W1 = 0;                               // [32][DIN]
  static constexpr int B1 = W1 + kHid * DIN;                 // [32]
  static constexpr int HW = B1 + kHid;                       // per hidden layer l = 2..L: W [j][k], WT [k][j], B
  static constexpr int HSTRIDE = 2 * kHid * kHid + kHid;
  static constexpr int WOUT = HW + (L - 1) * HSTRIDE;        // [OUT][32]
  static constexpr int BOUT = WOUT + OUT * kHid;             // [OUT] padded to 4
  static constexpr int WEND = BOUT + 4 * ((OUT + 3) / 4);
  __host__ __device__ static constexpr int w(int l) { return HW + (l - 2) * HSTRIDE; }                       // l >= 2
  __host__ __device__ static constexpr int wt(int l) { return HW + (l - 2) * HSTRIDE + kHid * kHid; }
  __host__ __device__ static constexpr int b(int l) { return HW + (l - 2) * HSTRIDE + 2 * kHid * kHid; }
  // per-warp region
  static constexpr int X = 0;                                // [DIN][32]
  static constexpr int DOUT = X + DIN * kTileP;              // [OUT][32]
  static constexpr int H = DOUT + OUT * kTileP;              // H_1 .. H_{L-1}: (L-1) tiles
  static constexpr int DA = H + (L - 1) * kTileF;            // one dA tile (backward only)
  static constexpr int WARP_FWD = DA;
  static constexpr int WARP_BWD = DA + (L > 1 ? kTileF : 0);
  // padded accumulator layout used for the in-CTA reduction: b1, W1, (b_l, W_l) l = 2..L, bout, Wout
  static constexpr int A_B1 = 0;
  static constexpr int A_W1 = kHid;
  static constexpr int A_H = A_W1 + kHid * DIN;              // per hidden layer: b (32) then W (1024)
  static constexpr int A_HSTRIDE = kHid + kHid * kHid;
  static constexpr int A_BO = A_H + (L - 1) * A_HSTRIDE;
  static constexpr int A_WO = A_BO + OUT;
  static constexpr int A_END = A_WO + OUT * kHid;
};

// Stage one particle's weights for one net into shared memory (zero padded to 32 / DIN / OUT).
template <int L, int DIN, int OUT>
__device__ __forceinline__ void stage_weights(float* sw, const float* __restrict__ th, const NetDev& net, int d) {
  using S = SmemLayout<L, DIN, OUT>;
  const int tid = threadIdx.x;
  const int w0 = net.width[0];
  for (int i = tid; i < kHid * DIN; i += kThreads) {
    int j = i / DIN, dd = i - j * DIN;
    sw[S::W1 + i] = (j < w0 && dd < d) ? th[net.off_w[0] + j * d + dd] : 0.0f;
  }
  for (int i = tid; i < kHid; i += kThreads) sw[S::B1 + i] = i < w0 ? th[net.off_b[0] + i] : 0.0f;
#pragma unroll
  for (int l = 2; l <= L; ++l) {
    const int win = net.width[l - 2], wout = net.width[l - 1];
    for (int i = tid; i < kHid * kHid; i += kThreads) {
      int j = i >> 5, k = i & 31;
      float w = (j < wout && k < win) ? th[net.off_w[l - 1] + j * win + k] : 0.0f;
      sw[S::w(l) + i] = w;
      sw[S::wt(l) + k * kHid + j] = w;
    }
    for (int i = tid; i < kHid; i += kThreads) sw[S::b(l) + i] = i < wout ? th[net.off_b[l - 1] + i] : 0.0f;
  }
  const int wl = net.width[L - 1];
  for (int i = tid; i < OUT * kHid; i += kThreads) {
    int o = i >> 5, k = i & 31;
    sw[S::WOUT + i] = (o < net.out_dim && k < wl) ? th[net.off_w[L] + o * wl + k] : 0.0f;
  }
  for (int i = tid; i < 4 * ((OUT + 3) / 4); i += kThreads) sw[S::BOUT + i] = i < net.out_dim ? th[net.off_b[L] + i] : 0.0f;
}

// Gather the tile's inputs (and, for backward, the incoming output gradients) into the warp's smem.
template <int DIN, int OUT, bool BWD>
__device__ __forceinline__ void load_tile(float* sx, float* sdout, const MlpArgs& a, const NetDev& net, int p, int tile,
                                          int lane) {
  const int Q = a.T * a.n;
  const int q = tile * kTileP + lane;
  const bool valid = q < Q;
  int src = 0;
  if (valid) {
    int t = q / a.n;
    src = (a.task_idx != nullptr ? __ldg(a.task_idx + t) : t) * a.n + (q - t * a.n);
  }
#pragma unroll
  for (int dd = 0; dd < DIN; ++dd) sx[dd * kTileP + lane] = (valid && dd < a.d) ? __ldg(a.x + (size_t)src * a.d + dd) : 0.0f;
  if (BWD) {
    const float* dsrc = a.dout[blockIdx.z] + ((size_t)p * Q + q) * net.out_dim;
#pragma unroll
    for (int o = 0; o < OUT; ++o) sdout[o * kTileP + lane] = (valid && o < net.out_dim) ? __ldg(dsrc + o) : 0.0f;
  }
}

// Layer 1: h[e][i] = tanh(b1 + W1 x) for features 8ob+e, points 4pb+i.  xr[dd] holds the 4 points' inputs.
template <int DIN>
__device__ __forceinline__ void layer1(float (&h)[8][4], const float* sw1, const float* sb1, const float4 (&xr)[DIN], int ob) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int j = 8 * ob + e;
    const float b = sb1[j];
    float a0 = b, a1 = b, a2 = b, a3 = b;
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) {
      const float w = sw1[j * DIN + dd];
      a0 = fmaf(w, xr[dd].x, a0); a1 = fmaf(w, xr[dd].y, a1); a2 = fmaf(w, xr[dd].z, a2); a3 = fmaf(w, xr[dd].w, a3);
    }
    h[e][0] = tanh_fast(a0); h[e][1] = tanh_fast(a1); h[e][2] = tanh_fast(a2); h[e][3] = tanh_fast(a3);
  }
}

// acc[e][i] += sum_k Wk[k][8ob+e] * A[k][4pb+i]  -- the 32x32x32 warp GEMM, operands from smem.
__device__ __forceinline__ void gemm_tile(float (&acc)[8][4], const float* __restrict__ sA, const float* __restrict__ sWk,
                                          int pb, int ob) {
#pragma unroll 8
  for (int k = 0; k < kHid; ++k) {
    const float4 a = lds4(sA + k * kSRow + 4 * pb);
    const float4 w0 = lds4(sWk + k * kHid + 8 * ob);
    const float4 w1 = lds4(sWk + k * kHid + 8 * ob + 4);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e][0] = fmaf(w[e], a.x, acc[e][0]); acc[e][1] = fmaf(w[e], a.y, acc[e][1]);
      acc[e][2] = fmaf(w[e], a.z, acc[e][2]); acc[e][3] = fmaf(w[e], a.w, acc[e][3]);
    }
  }
}

__device__ __forceinline__ void store_tile(float* sH, const float (&h)[8][4], int pb, int ob) {
#pragma unroll
  for (int e = 0; e < 8; ++e) sts4(sH + (8 * ob + e) * kSRow + 4 * pb, make_float4(h[e][0], h[e][1], h[e][2], h[e][3]));
}

// Forward through the hidden layers; returns the last hidden activations in registers and leaves
// H_1 .. H_{L-1} in the warp's shared memory.
template <int L, int DIN, int OUT>
__device__ __forceinline__ void forward_hidden(float (&h)[8][4], const float* sw, float* swarp, float4 (&xr)[DIN], int pb, int ob) {
  using S = SmemLayout<L, DIN, OUT>;
#pragma unroll
  for (int dd = 0; dd < DIN; ++dd) xr[dd] = lds4(swarp + S::X + dd * kTileP + 4 * pb);
  layer1<DIN>(h, sw + S::W1, sw + S::B1, xr, ob);
#pragma unroll
  for (int l = 2; l <= L; ++l) {
    float* sH = swarp + S::H + (l - 2) * kTileF;
    store_tile(sH, h, pb, ob);
    __syncwarp();
    const float4 b0 = lds4(sw + S::b(l) + 8 * ob), b1 = lds4(sw + S::b(l) + 8 * ob + 4);
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e][0] = h[e][1] = h[e][2] = h[e][3] = b[e];
    gemm_tile(h, sH, sw + S::wt(l), pb, ob);
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int i = 0; i < 4; ++i) h[e][i] = tanh_fast(h[e][i]);
  }
}

// ------------------------------------------------------------------------------------------- forward
template <int L, int DIN, int OUT>
__global__ void __launch_bounds__(kThreads) mlp_fwd_kernel(MlpArgs a) {
  using S = SmemLayout<L, DIN, OUT>;
  extern __shared__ __align__(16) float smem[];
  const NetDev& net = a.net[blockIdx.z];
  const int p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pb = lane & 7, ob = lane >> 3;
  float* sw = smem;
  float* swarp = smem + S::WEND + warp * S::WARP_FWD;
  stage_weights<L, DIN, OUT>(sw, a.theta + (size_t)p * a.D, net, a.d);
  __syncthreads();

  const int Q = a.T * a.n;
  const int tiles = (Q + kTileP - 1) / kTileP;
  // balanced split: chunk sizes differ by at most one tile
  const int t0 = (int)(((long long)tiles * blockIdx.x) / gridDim.x), t1 = (int)(((long long)tiles * (blockIdx.x + 1)) / gridDim.x);
  float* outp = a.out[blockIdx.z] + (size_t)p * Q * net.out_dim;

  for (int tile = t0 + warp; tile < t1; tile += kWarps) {
    load_tile<DIN, OUT, false>(swarp + S::X, nullptr, a, net, p, tile, lane);
    __syncwarp();
    float h[8][4];
    float4 xr[DIN];
    forward_hidden<L, DIN, OUT>(h, sw, swarp, xr, pb, ob);
    // output layer: reduce the 8-feature partials over the four ob lanes that share these points
    float o_[OUT][4];
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      const float4 w0 = lds4(sw + S::WOUT + o * kHid + 8 * ob), w1 = lds4(sw + S::WOUT + o * kHid + 8 * ob + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(w[e], h[e][i], s);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        o_[o][i] = s + sw[S::BOUT + o];
      }
    }
    if (ob == 0) {
      const int q0 = tile * kTileP + 4 * pb;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (q0 + i < Q) {
#pragma unroll
          for (int o = 0; o < OUT; ++o)
            if (o < net.out_dim) outp[(size_t)(q0 + i) * net.out_dim + o] = o_[o][i];
        }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------- backward
template <int L, int DIN, int OUT>
__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_kernel(MlpArgs a) {
  using S = SmemLayout<L, DIN, OUT>;
  extern __shared__ __align__(16) float smem[];
  const NetDev& net = a.net[blockIdx.z];
  const int p = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pb = lane & 7, ob = lane >> 3;  // (4 points) x (8 features) tiling
  float* sw = smem;
  float* swarp = smem + S::WEND + warp * S::WARP_BWD;
  stage_weights<L, DIN, OUT>(sw, a.theta + (size_t)p * a.D, net, a.d);
  __syncthreads();

  const int Q = a.T * a.n;
  const int tiles = (Q + kTileP - 1) / kTileP;
  // balanced split: chunk sizes differ by at most one tile
  const int t0 = (int)(((long long)tiles * blockIdx.x) / gridDim.x), t1 = (int)(((long long)tiles * (blockIdx.x + 1)) / gridDim.x);

  // persistent per-lane accumulators (partial sums over this warp's tiles)
  constexpr int LH = L > 1 ? L - 1 : 1;
  float accW[LH][8][4];   // hidden layer l (2..L): dW_l[j = jb + 4e][k = kb + 8f], jb = lane >> 3, kb = lane & 7
  float accb[L][8];       // db_l[8ob + e], partial over pb
  float accWo[OUT][8];    // dWout[o][8ob + e], partial over pb
  float accbo[OUT];       // dbout[o], partial over pb (identical on all ob)
  float accW1[8][DIN];    // dW1[8ob + e][dd]
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int l = 0; l < L; ++l) accb[l][e] = 0.0f;
#pragma unroll
    for (int l = 0; l < LH; ++l)
#pragma unroll
      for (int f = 0; f < 4; ++f) accW[l][e][f] = 0.0f;
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) accW1[e][dd] = 0.0f;
  }
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    accbo[o] = 0.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) accWo[o][e] = 0.0f;
  }

  for (int tile = t0 + warp; tile < t1; tile += kWarps) {
    load_tile<DIN, OUT, true>(swarp + S::X, swarp + S::DOUT, a, net, p, tile, lane);
    __syncwarp();
    float h[8][4];
    float4 xr[DIN];
    forward_hidden<L, DIN, OUT>(h, sw, swarp, xr, pb, ob);

    // ---- output layer backward: h = last hidden activations (registers)
    float4 dr[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) dr[o] = lds4(swarp + S::DOUT + o * kTileP + 4 * pb);
    float da[8][4];
#pragma unroll
    for (int e = 0; e < 8; ++e) da[e][0] = da[e][1] = da[e][2] = da[e][3] = 0.0f;
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      accbo[o] += (dr[o].x + dr[o].y) + (dr[o].z + dr[o].w);
      const float4 w0 = lds4(sw + S::WOUT + o * kHid + 8 * ob), w1 = lds4(sw + S::WOUT + o * kHid + 8 * ob + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        accWo[o][e] += fmaf(dr[o].x, h[e][0], fmaf(dr[o].y, h[e][1], fmaf(dr[o].z, h[e][2], dr[o].w * h[e][3])));
        da[e][0] = fmaf(w[e], dr[o].x, da[e][0]); da[e][1] = fmaf(w[e], dr[o].y, da[e][1]);
        da[e][2] = fmaf(w[e], dr[o].z, da[e][2]); da[e][3] = fmaf(w[e], dr[o].w, da[e][3]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int i = 0; i < 4; ++i) da[e][i] *= fmaf(-h[e][i], h[e][i], 1.0f);

#pragma unroll
    for (int l = L; l >= 2; --l) {
      float* sDA = swarp + S::DA;
      const float* sH = swarp + S::H + (l - 2) * kTileF;   // H_{l-1}
#pragma unroll
      for (int e = 0; e < 8; ++e) accb[l - 1][e] += (da[e][0] + da[e][1]) + (da[e][2] + da[e][3]);
      store_tile(sDA, da, pb, ob);
      __syncwarp();
      // ---- dW_l += dA_l^T H_{l-1} over the tile's 32 points; lane owns j = jb + 4e, k = kb + 8f
      {
        const int jb = lane >> 3, kb = lane & 7;
        const float* sA = sDA + jb * kSRow;
        const float* sB = sH + kb * kSRow;
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
          float4 bv[4];
#pragma unroll
          for (int f = 0; f < 4; ++f) bv[f] = lds4(sB + 8 * f * kSRow + 4 * g);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 av = lds4(sA + 4 * e * kSRow + 4 * g);
#pragma unroll
            for (int f = 0; f < 4; ++f)
              accW[l - 2][e][f] += fmaf(av.x, bv[f].x, fmaf(av.y, bv[f].y, fmaf(av.z, bv[f].z, av.w * bv[f].w)));
          }
        }
      }
      // ---- dH_{l-1} = dA_l W_l ; dA_{l-1} = dH_{l-1} * (1 - H_{l-1}^2)
#pragma unroll
      for (int e = 0; e < 8; ++e) da[e][0] = da[e][1] = da[e][2] = da[e][3] = 0.0f;
      gemm_tile(da, sDA, sw + S::w(l), pb, ob);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 hv = lds4(sH + (8 * ob + e) * kSRow + 4 * pb);
        da[e][0] *= fmaf(-hv.x, hv.x, 1.0f); da[e][1] *= fmaf(-hv.y, hv.y, 1.0f);
        da[e][2] *= fmaf(-hv.z, hv.z, 1.0f); da[e][3] *= fmaf(-hv.w, hv.w, 1.0f);
      }
      __syncwarp();   // all lanes are done with sDA before the next layer overwrites it
    }
    // ---- first layer: db1, dW1
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      accb[0][e] += (da[e][0] + da[e][1]) + (da[e][2] + da[e][3]);
#pragma unroll
      for (int dd = 0; dd < DIN; ++dd)
        accW1[e][dd] += fmaf(da[e][0], xr[dd].x, fmaf(da[e][1], xr[dd].y, fmaf(da[e][2], xr[dd].z, da[e][3] * xr[dd].w)));
    }
    __syncwarp();
  }

  // ---- reduce lanes that share a feature block (xor over the pb bits), then warps in a fixed order
#pragma unroll
  for (int s = 1; s <= 4; s <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
      for (int l = 0; l < L; ++l) accb[l][e] += __shfl_xor_sync(0xffffffffu, accb[l][e], s);
#pragma unroll
      for (int dd = 0; dd < DIN; ++dd) accW1[e][dd] += __shfl_xor_sync(0xffffffffu, accW1[e][dd], s);
#pragma unroll
      for (int o = 0; o < OUT; ++o) accWo[o][e] += __shfl_xor_sync(0xffffffffu, accWo[o][e], s);
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) accbo[o] += __shfl_xor_sync(0xffffffffu, accbo[o], s);
  }
  __syncthreads();                  // everyone is done with the activation tiles: reuse them as the CTA accumulator
  float* sacc = smem + S::WEND;     // A_END floats, padded net-local parameter order
  for (int i = threadIdx.x; i < S::A_END; i += kThreads) sacc[i] = 0.0f;
  __syncthreads();
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
      {
        const int jb = lane >> 3, kb = lane & 7;
#pragma unroll
        for (int l = 2; l <= L; ++l)
#pragma unroll
          for (int e = 0; e < 8; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f)
              sacc[S::A_H + (l - 2) * S::A_HSTRIDE + kHid + (jb + 4 * e) * kHid + kb + 8 * f] += accW[l - 2][e][f];
      }
      if (pb == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = 8 * ob + e;
          sacc[S::A_B1 + j] += accb[0][e];
#pragma unroll
          for (int l = 2; l <= L; ++l) sacc[S::A_H + (l - 2) * S::A_HSTRIDE + j] += accb[l - 1][e];
#pragma unroll
          for (int dd = 0; dd < DIN; ++dd) sacc[S::A_W1 + j * DIN + dd] += accW1[e][dd];
#pragma unroll
          for (int o = 0; o < OUT; ++o) sacc[S::A_WO + o * kHid + j] += accWo[o][e];
        }
        if (ob == 0) {
#pragma unroll
          for (int o = 0; o < OUT; ++o) sacc[S::A_BO + o] += accbo[o];
        }
      }
    }
    __syncthreads();
  }
  // ---- un-pad into the reference's flat order (bias before weight per layer) and write this CTA's partial
  float* dst = a.partial[blockIdx.z] + ((size_t)blockIdx.x * a.P + p) * net.total;
  const int base = net.off_b[0];
  for (int i = threadIdx.x; i < net.total; i += kThreads) {
    const int g = base + i;   // offset inside the particle row
    int src = 0;
    if (g >= net.off_w[L]) { int r = g - net.off_w[L]; int wl = net.width[L - 1]; src = S::A_WO + (r / wl) * kHid + (r % wl); }
    else if (g >= net.off_b[L]) src = S::A_BO + (g - net.off_b[L]);
    else {
      src = -1;
#pragma unroll
      for (int l = L; l >= 2; --l) {
        if (src < 0 && g >= net.off_w[l - 1]) { int r = g - net.off_w[l - 1]; int win = net.width[l - 2];
          src = S::A_H + (l - 2) * S::A_HSTRIDE + kHid + (r / win) * kHid + (r % win); }
        else if (src < 0 && g >= net.off_b[l - 1]) src = S::A_H + (l - 2) * S::A_HSTRIDE + (g - net.off_b[l - 1]);
      }
      if (src < 0) {
        if (g >= net.off_w[0]) { int r = g - net.off_w[0]; src = S::A_W1 + (r / a.d) * DIN + (r % a.d); }
        else src = S::A_B1 + (g - net.off_b[0]);
      }
    }
    dst[i] = sacc[src];
  }
}

// ------------------------------------------------------------------------------------------- host dispatch
template <int L, int DIN, int OUT>
int launch_kernel(const MlpArgs& a, int chunks, int nets, bool bwd, cudaStream_t st) {
  using S = SmemLayout<L, DIN, OUT>;
  dim3 grid(chunks, a.P, nets), block(kThreads);
  if (!bwd) {
    size_t smem = sizeof(float) * (S::WEND + kWarps * S::WARP_FWD);
    PACOH_CUDA_CHECK(cudaFuncSetAttribute(mlp_fwd_kernel<L, DIN, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_fwd_kernel<L, DIN, OUT><<<grid, block, smem, st>>>(a);
  } else {
    size_t warp_floats = (size_t)kWarps * S::WARP_BWD;
    if (warp_floats < (size_t)S::A_END) warp_floats = S::A_END;
    size_t smem = sizeof(float) * (S::WEND + warp_floats);
    PACOH_CUDA_CHECK(cudaFuncSetAttribute(mlp_bwd_kernel<L, DIN, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_bwd_kernel<L, DIN, OUT><<<grid, block, smem, st>>>(a);
  }
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

template <int L, int DIN>
int dispatch_out(const MlpArgs& a, int out_pad, int chunks, int nets, bool bwd, cudaStream_t st) {
  switch (out_pad) {
    case 1: return launch_kernel<L, DIN, 1>(a, chunks, nets, bwd, st);
    case 2: return launch_kernel<L, DIN, 2>(a, chunks, nets, bwd, st);
    case 4: return launch_kernel<L, DIN, 4>(a, chunks, nets, bwd, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

template <int L>
int dispatch_din(const MlpArgs& a, int din_pad, int out_pad, int chunks, int nets, bool bwd, cudaStream_t st) {
  switch (din_pad) {
    case 1: return dispatch_out<L, 1>(a, out_pad, chunks, nets, bwd, st);
    case 2: return dispatch_out<L, 2>(a, out_pad, chunks, nets, bwd, st);
    case 4: return dispatch_out<L, 4>(a, out_pad, chunks, nets, bwd, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

int pad_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : 4); }

}  // namespace

bool net_is_fast(const NetDev& n, int d) {
  if (n.n_hidden < 1 || n.n_hidden > 4 || d > 4 || n.out_dim > 4) return false;
  for (int l = 0; l < n.n_hidden; ++l)
    if (n.width[l] < 1 || n.width[l] > kHid) return false;
  return true;
}

// Launches the register-tiled kernels for the `nets` nets in a.net[] (they must have the same depth).
int launch_mlp_fast(const MlpArgs& a, int nets, int chunks, bool bwd, cudaStream_t st) {
  const int din = pad_pow2(a.d);
  int out_pad = 1;
  for (int z = 0; z < nets; ++z) out_pad = max(out_pad, pad_pow2(a.net[z].out_dim));
  switch (a.net[0].n_hidden) {
    case 1: return dispatch_din<1>(a, din, out_pad, chunks, nets, bwd, st);
    case 2: return dispatch_din<2>(a, din, out_pad, chunks, nets, bwd, st);
    case 3: return dispatch_din<3>(a, din, out_pad, chunks, nets, bwd, st);
    case 4: return dispatch_din<4>(a, din, out_pad, chunks, nets, bwd, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

}  // namespace pacoh
