// Tensor-core backward pass of the per-particle MLPs (recompute-forward + reverse pass), sm_100a.
//
// Same semantics as mlp_bwd_kernel in mlp.cu (the autograd reverse pass through NeuralNetworkVectorized,
// meta_learn/models.py:295-317, 343-349, svgd.py:16).  Work split per 128-point tile (thread t owns point t):
//
//   tensor cores (tcgen05, 3xTF32, A and D in TMEM)    CUDA cores
//   -------------------------------------------          -----------------------------------------------------------------
//   H_l   = tanh(H_{l-1} W_l^T + b_l)   (recompute)      layer 1, tanh, hi/lo splitting, output-layer backward
//   dH_{l-1} = dA_l W_l                                  dW_l += dA_l^T H_{l-1}  as a 32x32x32 warp GEMM over the warp's own 32
//                                                        points -- issued while the dH MMA of the same layer is in flight
//
// Every hidden layer keeps one per-warp [feature][point] tile (first H_l^T, later overwritten in place by dA_l^T); the
// bias / first-layer / output-layer gradients are lane-per-feature row sums over those tiles.  Accumulators persist in
// registers over all tiles of the CTA and are reduced once, in a fixed order, into a (chunk, particle) partial.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace pacoh {

namespace {

using namespace tc;

constexpr int kThreads = 128;
constexpr int kWarps = 4;
constexpr int kTile = 128;
constexpr int kTF = kHid * kSRow;   // floats of one per-warp [32 features][36] tile

template <int L, int DIN, int OUT>
struct BwdSmem {
  static constexpr int B = 0;                                      // per layer l = 2..L: W hi, W lo, W^T hi, W^T lo (1024 floats each)
  static constexpr int W1 = B + (L - 1) * 4 * kHid * kHid;         // [32][DIN]
  static constexpr int B1 = W1 + kHid * DIN;
  static constexpr int BH = B1 + kHid;                             // biases of layers 2..L
  static constexpr int WOUT = BH + (L - 1) * kHid;                 // [OUT][32]
  static constexpr int BOUT = WOUT + OUT * kHid;
  static constexpr int WARP = BOUT + 4;
  // per-warp region
  static constexpr int T = 0;                                      // L tiles [32][36]: H_l^T, later dA_l^T
  static constexpr int X = T + L * kTF;                            // [DIN][32]
  static constexpr int DOUT = X + DIN * 32;                        // [OUT][32]
  static constexpr int WARP_SIZE = DOUT + 4 * ((OUT * 32 + 3) / 4);
  static constexpr int END = WARP + kWarps * WARP_SIZE;
  // padded accumulator layout for the in-CTA reduction (aliases the per-warp regions): b1, W1, (b_l, W_l) l = 2..L, bout, Wout
  static constexpr int R_B1 = 0;
  static constexpr int R_W1 = kHid;
  static constexpr int R_H = R_W1 + kHid * DIN;
  static constexpr int R_HSTRIDE = kHid + kHid * kHid;
  static constexpr int R_BO = R_H + (L - 1) * R_HSTRIDE;
  static constexpr int R_WO = R_BO + OUT;
  static constexpr int R_END = R_WO + OUT * kHid;
  static_assert(R_END <= kWarps * WARP_SIZE, "reduction buffer must fit in the per-warp regions");
};

template <int L, int DIN, int OUT>
__global__ void __launch_bounds__(kThreads, 3) mlp_tc_bwd_kernel(MlpArgs a) {
  using S = BwdSmem<L, DIN, OUT>;
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const NetDev& net = a.net[blockIdx.z];
  const int p = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* th = a.theta + (size_t)p * a.D;
  float* sw = smem + S::WARP + warp * S::WARP_SIZE;   // this warp's region

  if (warp == 0) tmem_alloc<kTmemCols>(&tmem_base_s);
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  // ---- stage the particle's weights
  const int w0 = net.width[0];
  for (int i = tid; i < kHid * DIN; i += kThreads) {
    const int j = i / DIN, dd = i - j * DIN;
    smem[S::W1 + i] = (j < w0 && dd < a.d) ? th[net.off_w[0] + j * a.d + dd] : 0.0f;
  }
  for (int i = tid; i < kHid; i += kThreads) smem[S::B1 + i] = i < w0 ? th[net.off_b[0] + i] : 0.0f;
#pragma unroll
  for (int l = 2; l <= L; ++l) {
    const int win = net.width[l - 2], wout = net.width[l - 1];
    float* bl = smem + S::B + (l - 2) * 4 * kHid * kHid;
    for (int i = tid; i < kHid * kHid; i += kThreads) {
      const int j = i >> 5, k = i & 31;
      const float w = (j < wout && k < win) ? th[net.off_w[l - 1] + j * win + k] : 0.0f;
      const float h = tf32_hi(w);
      bl[ktile_off(j, k, kHid)] = h;                                   // forward operand: rows = outputs j, K = inputs k
      bl[kHid * kHid + ktile_off(j, k, kHid)] = w - h;
      bl[2 * kHid * kHid + ktile_off(k, j, kHid)] = h;                 // dH operand: rows = inputs k, K = outputs j
      bl[3 * kHid * kHid + ktile_off(k, j, kHid)] = w - h;
    }
    for (int i = tid; i < kHid; i += kThreads) smem[S::BH + (l - 2) * kHid + i] = i < wout ? th[net.off_b[l - 1] + i] : 0.0f;
  }
  const int wl = net.width[L - 1];
  for (int i = tid; i < OUT * kHid; i += kThreads) {
    const int o = i >> 5, k = i & 31;
    smem[S::WOUT + i] = (o < net.out_dim && k < wl) ? th[net.off_w[L] + o * wl + k] : 0.0f;
  }
  if (tid < 4) smem[S::BOUT + tid] = 0.0f;
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
  const uint32_t bar = smem_u32(&mbar);
  uint32_t parity = 0;

  // ---- persistent accumulators
  constexpr int LH = L > 1 ? L - 1 : 1;
  float accW[LH][8][4];    // dW_l[j = jb + 4e][k = kb + 8f], jb = lane >> 3, kb = lane & 7 (this warp's points)
  float accb[L];           // db_l[lane]
  float accWo[OUT];        // dWout[o][lane]
  float accbo[OUT];        // dbout[o], per-thread partial over its own points
  float accW1[DIN];        // dW1[lane][dd]
#pragma unroll
  for (int l = 0; l < LH; ++l)
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
      for (int f = 0; f < 4; ++f) accW[l][e][f] = 0.0f;
#pragma unroll
  for (int l = 0; l < L; ++l) accb[l] = 0.0f;
#pragma unroll
  for (int o = 0; o < OUT; ++o) { accWo[o] = 0.0f; accbo[o] = 0.0f; }
#pragma unroll
  for (int dd = 0; dd < DIN; ++dd) accW1[dd] = 0.0f;

  const int Q = a.T * a.n;
  const int tiles = (Q + kTile - 1) / kTile;
  const int per = (tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(tiles, t0 + per);

  // this point's inputs / output gradients; the loads for the NEXT tile are issued one tile ahead (latency hidden)
  auto load_point = [&](int tile_i, float (&xo)[DIN], float (&dro)[OUT]) {
    const int qq = tile_i * kTile + tid;
    const bool ok = tile_i < t1 && qq < Q;
    int src = 0;
    if (ok) {
      const int t = qq / a.n;
      src = (a.task_idx != nullptr ? __ldg(a.task_idx + t) : t) * a.n + (qq - t * a.n);
    }
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) xo[dd] = (ok && dd < a.d) ? __ldg(a.x + (size_t)src * a.d + dd) : 0.0f;
    const float* dsrc = a.dout[blockIdx.z] + ((size_t)p * Q + qq) * net.out_dim;
#pragma unroll
    for (int o = 0; o < OUT; ++o) dro[o] = (ok && o < net.out_dim) ? __ldg(dsrc + o) : 0.0f;
  };
  float xn[DIN], drn[OUT];
  load_point(t0, xn, drn);

  for (int tile = t0; tile < t1; ++tile) {
    float x[DIN], dr[OUT];
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) {
      x[dd] = xn[dd];
      sw[S::X + dd * 32 + lane] = x[dd];
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      dr[o] = drn[o];
      sw[S::DOUT + o * 32 + lane] = dr[o];
      accbo[o] += dr[o];
    }
    load_point(tile + 1, xn, drn);
    // ---- forward recompute: layer 1 in registers, layers 2..L on the tensor cores; H_l^T kept in the warp tiles
    float h[kHid];
#pragma unroll
    for (int j4 = 0; j4 < kHid; j4 += 4) {
      const float4 b = lds4(smem + S::B1 + j4);
      float acc[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int dd = 0; dd < DIN; ++dd) acc[e] = fmaf(smem[S::W1 + (j4 + e) * DIN + dd], x[dd], acc[e]);
        h[j4 + e] = tanh_fast(acc[e]);
      }
    }
#pragma unroll
    for (int l = 2; l <= L; ++l) {
      float* Tprev = sw + S::T + (l - 2) * kTF;
#pragma unroll
      for (int k = 0; k < kHid; ++k) Tprev[k * kSRow + lane] = h[k];
      store_a_tmem(lane_base, h);                   // A operand of the recompute GEMM: this point's row -> its TMEM lane
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint32_t b_hi = smem_u32(smem + S::B + (l - 2) * 4 * kHid * kHid);
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
    // ---- output layer backward.  dWout[o][j] += sum_pt H_L[pt][j] dout[pt][o] as a lane-per-feature row sum
    float* TL = sw + S::T + (L - 1) * kTF;
#pragma unroll
    for (int k = 0; k < kHid; ++k) TL[k * kSRow + lane] = h[k];
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 hv = lds4(TL + lane * kSRow + 4 * g);
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float4 dv = lds4(sw + S::DOUT + o * 32 + 4 * g);
        accWo[o] += fmaf(hv.x, dv.x, fmaf(hv.y, dv.y, fmaf(hv.z, dv.z, hv.w * dv.w)));
      }
    }
    float da[kHid];
#pragma unroll
    for (int j4 = 0; j4 < kHid; j4 += 4) {
      float dh[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float4 w = lds4(smem + S::WOUT + o * kHid + j4);
        dh[0] = fmaf(w.x, dr[o], dh[0]); dh[1] = fmaf(w.y, dr[o], dh[1]); dh[2] = fmaf(w.z, dr[o], dh[2]); dh[3] = fmaf(w.w, dr[o], dh[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) da[j4 + e] = dh[e] * fmaf(-h[j4 + e], h[j4 + e], 1.0f);
    }
    // ---- hidden layers L..2
#pragma unroll
    for (int l = L; l >= 2; --l) {
      float* Tl = sw + S::T + (l - 1) * kTF;        // H_l^T  -> dA_l^T
      float* Tp = sw + S::T + (l - 2) * kTF;        // H_{l-1}^T
      __syncwarp();                                 // every lane is done reading H_l^T
#pragma unroll
      for (int k = 0; k < kHid; ++k) Tl[k * kSRow + lane] = da[k];
      store_a_tmem(lane_base, da);
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint32_t bt_hi = smem_u32(smem + S::B + (l - 2) * 4 * kHid * kHid + 2 * kHid * kHid);
        gemm128x32x32_3xtf32_ts(tmem, bt_hi, bt_hi + kHid * kHid * 4, bar);              // dH_{l-1} = dA_l W_l
      }
      // ---- while the MMA runs: dW_l += dA_l^T H_{l-1} over this warp's 32 points, and db_l
      {
        const int jb = lane >> 3, kb = lane & 7;
        const float* sA = Tl + jb * kSRow;
        const float* sB = Tp + kb * kSRow;
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
        float s = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 v4 = lds4(Tl + lane * kSRow + 4 * g);
          s += (v4.x + v4.y) + (v4.z + v4.w);
        }
        accb[l - 1] += s;
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      fence_after_sync();
      uint32_t v[32];
      tmem_ld32(lane_base + kTmemD, v);
      __syncwarp();                                 // the warp GEMM above is done reading H_{l-1}^T
#pragma unroll
      for (int k = 0; k < kHid; ++k) {
        const float hp = Tp[k * kSRow + lane];
        da[k] = __uint_as_float(v[k]) * fmaf(-hp, hp, 1.0f);
      }
    }
    // ---- first layer: db1, dW1 as lane-per-feature row sums over dA_1^T
    {
      float* T1 = sw + S::T;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < kHid; ++k) T1[k * kSRow + lane] = da[k];
      __syncwarp();
      float s = 0.0f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 v4 = lds4(T1 + lane * kSRow + 4 * g);
        s += (v4.x + v4.y) + (v4.z + v4.w);
#pragma unroll
        for (int dd = 0; dd < DIN; ++dd) {
          const float4 xv = lds4(sw + S::X + dd * 32 + 4 * g);
          accW1[dd] += fmaf(v4.x, xv.x, fmaf(v4.y, xv.y, fmaf(v4.z, xv.z, v4.w * xv.w)));
        }
      }
      accb[0] += s;
      __syncwarp();                                 // before the next tile overwrites X / DOUT / T
    }
  }

  // ---- reduce: warps in a fixed order into the padded layout (aliasing the A tiles), then un-pad to the flat order
#pragma unroll
  for (int o = 0; o < OUT; ++o) accbo[o] = warp_sum(accbo[o]);
  fence_before_sync();
  __syncthreads();
  float* sacc = smem + S::WARP;
  for (int i = tid; i < S::R_END; i += kThreads) sacc[i] = 0.0f;
  __syncthreads();
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
      const int jb = lane >> 3, kb = lane & 7;
#pragma unroll
      for (int l = 2; l <= L; ++l) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) sacc[S::R_H + (l - 2) * S::R_HSTRIDE + kHid + (jb + 4 * e) * kHid + kb + 8 * f] += accW[l - 2][e][f];
        sacc[S::R_H + (l - 2) * S::R_HSTRIDE + lane] += accb[l - 1];
      }
      sacc[S::R_B1 + lane] += accb[0];
#pragma unroll
      for (int dd = 0; dd < DIN; ++dd) sacc[S::R_W1 + lane * DIN + dd] += accW1[dd];
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        sacc[S::R_WO + o * kHid + lane] += accWo[o];
        if (lane == 0) sacc[S::R_BO + o] += accbo[o];
      }
    }
    __syncthreads();
  }
  float* dst = a.partial[blockIdx.z] + ((size_t)blockIdx.x * a.P + p) * net.total;
  const int base = net.off_b[0];
  for (int i = tid; i < net.total; i += kThreads) {
    const int g = base + i;
    int src = -1;
    if (g >= net.off_w[L]) { const int r = g - net.off_w[L]; src = S::R_WO + (r / wl) * kHid + (r % wl); }
    else if (g >= net.off_b[L]) src = S::R_BO + (g - net.off_b[L]);
    else {
#pragma unroll
      for (int l = L; l >= 2; --l) {
        if (src < 0 && g >= net.off_w[l - 1]) { const int r = g - net.off_w[l - 1]; const int win = net.width[l - 2];
          src = S::R_H + (l - 2) * S::R_HSTRIDE + kHid + (r / win) * kHid + (r % win); }
        else if (src < 0 && g >= net.off_b[l - 1]) src = S::R_H + (l - 2) * S::R_HSTRIDE + (g - net.off_b[l - 1]);
      }
      if (src < 0) {
        if (g >= net.off_w[0]) { const int r = g - net.off_w[0]; src = S::R_W1 + (r / a.d) * DIN + (r % a.d); }
        else src = S::R_B1 + (g - net.off_b[0]);
      }
    }
    dst[i] = sacc[src];
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

template <int L, int DIN, int OUT>
int launch_tc_bwd(const MlpArgs& a, int chunks, int nets, cudaStream_t st) {
  using S = BwdSmem<L, DIN, OUT>;
  const size_t smem = sizeof(float) * S::END;
  PACOH_CUDA_CHECK(cudaFuncSetAttribute(mlp_tc_bwd_kernel<L, DIN, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(chunks, a.P, nets);
  mlp_tc_bwd_kernel<L, DIN, OUT><<<grid, kThreads, smem, st>>>(a);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

template <int L, int DIN>
int bwd_dispatch_out(const MlpArgs& a, int out_pad, int chunks, int nets, cudaStream_t st) {
  switch (out_pad) {
    case 1: return launch_tc_bwd<L, DIN, 1>(a, chunks, nets, st);
    case 2: return launch_tc_bwd<L, DIN, 2>(a, chunks, nets, st);
    case 4: return launch_tc_bwd<L, DIN, 4>(a, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

template <int L>
int bwd_dispatch_din(const MlpArgs& a, int din_pad, int out_pad, int chunks, int nets, cudaStream_t st) {
  switch (din_pad) {
    case 1: return bwd_dispatch_out<L, 1>(a, out_pad, chunks, nets, st);
    case 2: return bwd_dispatch_out<L, 2>(a, out_pad, chunks, nets, st);
    case 4: return bwd_dispatch_out<L, 4>(a, out_pad, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

int bwd_pad_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : 4); }

}  // namespace

// Backward pass of the `nets` nets in a.net[] (same depth <= 3, widths <= 32, d <= 4, out <= 4) on tcgen05 + CUDA cores.
int launch_mlp_tc_bwd(const MlpArgs& a, int nets, int chunks, cudaStream_t st) {
  const int din = bwd_pad_pow2(a.d);
  int out_pad = 1;
  for (int z = 0; z < nets; ++z) out_pad = max(out_pad, bwd_pad_pow2(a.net[z].out_dim));
  switch (a.net[0].n_hidden) {
    case 1: return bwd_dispatch_din<1>(a, din, out_pad, chunks, nets, st);
    case 2: return bwd_dispatch_din<2>(a, din, out_pad, chunks, nets, st);
    case 3: return bwd_dispatch_din<3>(a, din, out_pad, chunks, nets, st);
    case 4: return bwd_dispatch_din<4>(a, din, out_pad, chunks, nets, st);
  }
  return PACOH_ERR_UNSUPPORTED;
}

}  // namespace pacoh
