// Tensor-core backward pass of the per-particle MLPs (recompute-forward + reverse pass), sm_100a.
//
// Same semantics as mlp_bwd_kernel in mlp.cu (the autograd reverse pass through NeuralNetworkVectorized,
// meta_learn/models.py:295-317, 343-349, svgd.py:16).  One CTA = THREE independent 256-thread warpgroups that share the
// staged weights; each warpgroup walks its own 128-point tiles with its own tensor-memory columns, mbarriers and named
// barrier, so the three tile pipelines hide each other's MMA / barrier round trips.  TWO threads share a point (16 of
// the 32 features each; warps w and w + 4 address the same 32 TMEM lanes): 24 warps per SM instead of 12 at half the
// registers per thread -- the elementwise chains (tanh, hi/lo splits, column stores) are latency-bound otherwise.
//
//   tensor cores (tcgen05, 3xTF32, accumulators in TMEM)             CUDA cores
//   ------------------------------------------------------------     ------------------------------------------------
//   H_l      = tanh(H_{l-1} W_l^T + b_l)  (recompute; A in TMEM)     layer 1, tanh, hi/lo splitting, output-layer backward,
//   dH_{l-1} = dA_l W_l                   (A in TMEM)                bias / first-layer / output-layer gradients as
//   dW_l    += dA_l^T H_{l-1}             (both operands in smem,    lane-per-feature row sums
//              K = the 128 points of the tile; the accumulator
//              stays in tensor memory over ALL tiles of the warpgroup)
//
// The weight-gradient GEMM needs its operands K-major with K = points, i.e. TRANSPOSED tiles [feature][point].  Each
// thread writes its point's column of dA_l^T / H_{l-1}^T straight into a SWIZZLE_128B K-major tile: a warp's 32 points
// are one 128-byte swizzle row per feature (conflict-free column stores AND conflict-free lane-per-feature row reads),
// one 8 KB K-block per warp.  The lo parts are STACKED under the values along M / N (rows 0-31: the unsplit fp32
// value -- the tensor core truncates it to tf32 itself --, rows 32-63: value - tf32(value)), so ONE M64 N64 MMA per
// 8 points yields hi*hi, hi*lo and lo*hi as separate 32x32 blocks of the accumulator (summed once per CTA):
// 16 MMAs per tile instead of 48.  (Layout and truncation validated by tools/ubench/tc_sw128_test.cu.)
// The dW MMAs are issued by a second thread and committed to their own mbarrier, so they overlap the dH round trip;
// the operand tiles are only waited for right before they are overwritten.
#include <cstdlib>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace pacoh {

namespace {

using namespace tc;

constexpr int kWG = 3;                   // warpgroups (independent tile pipelines) per CTA
constexpr int kWGThreads = 256;          // TWO threads per point: thread (point, half) owns 16 of the 32 features
constexpr int kFH = 16;
constexpr int kThreads = kWG * kWGThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 128;
constexpr int kStackF = 8192;            // floats of one stacked transposed tile [64 rows][128 points] (32 KB)
// tensor-memory columns of one warpgroup: D 32 | A hi 32 | A lo 32 (tc_common.cuh) | dW accumulator 64
constexpr uint32_t kTmemDW = 96;
constexpr uint32_t kTmemWG = 160;
constexpr int kBwdMaxTilesPerAcc = 120;  // persistent schedule: tiles one warpgroup may sum into its TMEM accumulator (see mlp_tc_bwd_slots)

// K-major SWIZZLE_128B shared-memory descriptor: 8-row groups 1024 B apart, start address may advance by 32 B per K-step
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;      // stride byte offset
  d |= (uint64_t)1 << 46;                // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int L, int DIN, int OUT>
struct BwdSmem {
  static constexpr int LH = L > 1 ? L - 1 : 0;
  // per-warpgroup stacked transposed tiles first (1024-byte aligned)
  static constexpr int WG_TILES = (1 + LH) * kStackF;              // dA_l^T, then H_l^T for l = 1..L-1
  static constexpr int B = kWG * WG_TILES;                         // per layer l = 2..L: W hi, W lo, W^T hi, W^T lo (1024 floats each)
  static constexpr int W1 = B + LH * 4 * kHid * kHid;              // [32][DIN]
  static constexpr int B1 = W1 + kHid * DIN;
  static constexpr int BH = B1 + kHid;                             // biases of layers 2..L
  static constexpr int WOUT = BH + LH * kHid;                      // [OUT][32]
  static constexpr int BOUT = WOUT + OUT * kHid;
  static constexpr int WARP = BOUT + 4;
  // per-warp region
  static constexpr int X = 0;                                      // [DIN][32]
  static constexpr int DOUT = X + DIN * 32;                        // [OUT][32]
  static constexpr int WARP_SIZE = DOUT + 4 * ((OUT * 32 + 3) / 4);
  static constexpr int END = WARP + kWarps * WARP_SIZE;
  // padded accumulator layout for the in-CTA reduction (aliases warpgroup 0's tiles): b1, W1, (b_l, W_l) l = 2..L, bout, Wout
  static constexpr int R_B1 = 0;
  static constexpr int R_W1 = kHid;
  static constexpr int R_H = R_W1 + kHid * DIN;
  static constexpr int R_HSTRIDE = kHid + kHid * kHid;
  static constexpr int R_BO = R_H + LH * R_HSTRIDE;
  static constexpr int R_WO = R_BO + OUT;
  static constexpr int R_END = R_WO + OUT * kHid;
  static constexpr int R_PAD = ((R_END + 31) / 32) * 32 + 1;       // per-warp copy stride (odd: the row-major dW stores spread over the banks)
  static_assert((kWarps + 1) * R_PAD <= kWG * WG_TILES, "reduction buffers must fit in the transposed tiles");
  static_assert(L <= 2, "deeper nets do not fit three warpgroups of transposed tiles in shared memory");
};

template <int L, int DIN, int OUT>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_bwd_kernel(MlpArgs a, int nets, int per_cta) {
  using S = BwdSmem<L, DIN, OUT>;
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t mbar[kWG][2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = tid >> 8, wt = tid & 255;                       // warpgroup, thread inside it
  const int wq = (wt >> 5) & 3, fh = wt >> 7;                    // TMEM lane quadrant (points 32 wq + lane), feature half
  const int f0 = kFH * fh;                                       // this thread's features f0 .. f0 + 15
  float* sw = smem + S::WARP + warp * S::WARP_SIZE;              // this warp's region

  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (tid < kWG) { mbar_init(smem_u32(&mbar[tid][0]), 1); mbar_init(smem_u32(&mbar[tid][1]), 1); }
  uint32_t parity = 0, parity_dw = 0;
  bool dw_pending = false;     // dW MMAs in flight still read the transposed tiles

  // ---- PERSISTENT CTA: the (net, particle, tile) space is linearised and cut into gridDim.x equal ranges of `per_cta`
  //      tiles, so every SM gets the same work whatever P, nets and the batch size are (no wave quantisation).  A range
  //      is processed as segments of one (net, particle) each: stage its weights, walk its tiles, reduce, and write the
  //      partial gradient into slot (this CTA - first CTA touching that (net, particle)) of the zero-initialised
  //      (slots, P, D_net) partial buffer.
  const int Q = a.T * a.n;
  const int tiles = (Q + kTile - 1) / kTile;
  const long long total_tiles = (long long)nets * a.P * tiles;
  const long long g_begin = (long long)blockIdx.x * per_cta;
  const long long g_end = g_begin + per_cta < total_tiles ? g_begin + per_cta : total_tiles;
  for (long long g_seg = g_begin; g_seg < g_end;) {
  const int pn = (int)(g_seg / tiles);                           // (net, particle) of this segment
  const int zn = pn / a.P, p = pn - zn * a.P;
  const int t0 = (int)(g_seg - (long long)pn * tiles);
  const int t1 = (int)((g_end < (long long)(pn + 1) * tiles ? g_end : (long long)(pn + 1) * tiles) - (long long)pn * tiles);
  const int slot = (int)(blockIdx.x - ((long long)pn * tiles) / per_cta);
  g_seg += t1 - t0;
  const NetDev& net = a.net[zn];
  const float* th = a.theta + (size_t)p * a.D;
  // ---- stage the particle's weights (shared by the three warpgroups)
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
  const uint32_t tmem = tmem_base_s + kTmemWG * wg;                    // this warpgroup's columns
  // the same values as provably warp-uniform registers: the MMA issue sequences (descriptor arithmetic, UTCHMMA
  // operands) then stay on the uniform datapath instead of paying one R2UR per operand per instruction
  // (recomputed right before every issue: two shuffles are cheaper than four more live registers)
#define PACOH_UNIFORM_CTX()                                                                     \
  const int wg_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 8), 0);                       \
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base_s, 0) + kTmemWG * wg_u;           \
  const uint32_t bar_u = smem_u32(&mbar[wg_u][0]), bar_dw_u = smem_u32(&mbar[wg_u][1]);        \
  const uint32_t at_u = smem_u32(smem + wg_u * S::WG_TILES);                                   \
  (void)bar_dw_u; (void)at_u;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);       // this warp's 32 TMEM lanes
  const uint32_t bar = smem_u32(&mbar[wg][0]), bar_dw = smem_u32(&mbar[wg][1]);
  auto wg_sync = [&]() { asm volatile("bar.sync %0, 256;" :: "r"(wg + 1) : "memory"); };
  // ---- zero the dW accumulator (it is only ever accumulated into): each feature half zeroes 32 of the 64 columns
  if (L > 1) {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_st16(lane_base + kTmemDW + 32 * fh, z);
    tmem_st16(lane_base + kTmemDW + 32 * fh + 16, z);
    tmem_st_wait();
  }

  // ---- persistent accumulators.  Row sums are taken lane-per-feature over this warp's OWN 16 rows: lane l owns row
  //      f0 + (l & 15) and the point chunks 4 (l >> 4) .. + 3 of the warp's 32 points; halves are combined at the end.
  float accb[L];           // db_l[row]
  float accWo[OUT];        // dWout[o][row]
  float accbo[OUT];        // dbout[o], per-thread partial over its own points (feature half 0 only)
  float accW1[DIN];        // dW1[row][dd]
#pragma unroll
  for (int l = 0; l < L; ++l) accb[l] = 0.0f;
#pragma unroll
  for (int o = 0; o < OUT; ++o) { accWo[o] = 0.0f; accbo[o] = 0.0f; }
#pragma unroll
  for (int dd = 0; dd < DIN; ++dd) accW1[dd] = 0.0f;

  // this point's inputs / output gradients; the loads for the NEXT tile are issued one tile ahead (latency hidden)
  const int pt = wq * 32 + lane;
  auto load_point = [&](int tile_i, float (&xo)[DIN], float (&dro)[OUT]) {
    const int qq = tile_i * kTile + pt;
    const bool ok = tile_i < t1 && qq < Q;
    int src = 0;
    if (ok) {
      const int t = qq / a.n;
      src = (a.task_idx != nullptr ? __ldg(a.task_idx + t) : t) * a.n + (qq - t * a.n);
    }
#pragma unroll
    for (int dd = 0; dd < DIN; ++dd) xo[dd] = (ok && dd < a.d) ? __ldg(a.x + (size_t)src * a.d + dd) : 0.0f;
    const float* dsrc = a.dout[zn] + ((size_t)p * Q + qq) * net.out_dim;
#pragma unroll
    for (int o = 0; o < OUT; ++o) dro[o] = (ok && o < net.out_dim) ? __ldg(dsrc + o) : 0.0f;
  };
  float xn[DIN], drn[OUT];
  load_point(t0 + wg, xn, drn);

  // ---- addressing of the stacked SWIZZLE_128B tiles.  Row r, K-block wq, point = lane:
  //      float offset = wq * 2048 + (r >> 3) * 256 + (r & 7) * 32 + (lane ^ ((r & 7) << 2))
  float* AT = smem + wg * S::WG_TILES;               // dA_l^T stack (values rows 0-31, lo rows 32-63)
  float* kblk = AT + wq * 2048 + (f0 >> 3) * 256;    // this warp's K-block of the dA^T stack, at its first row group
  // swizzled position of this point inside row r: (((lane >> 2) ^ (r & 7)) << 2) + (lane & 3) == lane ^ ((r & 7) << 2)
#define PACOH_XO(j) (lane ^ ((j) << 2))
  const int rl = lane & 15;                                   // row-sum lane mapping: row f0 + rl ...
  const int rrow = (rl >> 3) * 256 + (rl & 7) * 32;           // ... relative to kblk
  const int rsw = (rl & 7) << 2;                              // its swizzle: logical chunk c sits at float offset ((4 c) ^ rsw)
  const int c0 = (lane >> 4) * 4;                             // ... and point chunks c0 .. c0 + 3
  // this point's 16-feature column slice -> rows f0.. (value) and 32 + f0.. (lo) of a stacked tile's K-block
  auto publish_col = [&](float* kb, const float (&v)[kFH]) {
#pragma unroll
    for (int e = 0; e < kFH; ++e) {
      const int o = (e >> 3) * 256 + (e & 7) * 32 + PACOH_XO(e & 7);
      kb[o] = v[e];
      kb[o + 1024] = v[e] - tf32_hi(v[e]);
    }
  };
  // ... and -> hi / lo halves of the TMEM A operand (columns f0 .. f0 + 15 of this point's lane)
  auto store_a_half = [&](const float (&v)[kFH]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float hv = tf32_hi(v[e]);
      hi[e] = __float_as_uint(hv);
      lo[e] = __float_as_uint(v[e] - hv);
    }
    tmem_st16(lane_base + kTmemAhi + f0, hi);
    tmem_st16(lane_base + kTmemAlo + f0, lo);
    tmem_st_wait();
  };
  auto park_col = [&](const float (&v)[kFH]) {      // values only (lane-per-feature row sums)
#pragma unroll
    for (int e = 0; e < kFH; ++e) kblk[(e >> 3) * 256 + (e & 7) * 32 + PACOH_XO(e & 7)] = v[e];
  };
  auto wait_dw = [&]() {
    if (dw_pending) {
      mbar_wait(bar_dw, parity_dw);
      parity_dw ^= 1;
      dw_pending = false;
    }
  };

  if (wg > 0) __nanosleep(2000u * wg);   // stagger the three tile pipelines by ~1/3 tile: they start in lockstep otherwise
  for (int tile = t0 + wg; tile < t1; tile += kWG) {
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
      if (fh == 0) accbo[o] += dr[o];
    }
    load_point(tile + kWG, xn, drn);
    // ---- forward recompute: layer 1 in registers, layers 2..L on the tensor cores; H_l^T kept in the transposed tiles
    float h[kFH];
#pragma unroll
    for (int j4 = 0; j4 < kFH; j4 += 4) {
      const float4 b = lds4(smem + S::B1 + f0 + j4);
      float acc[4] = {b.x, b.y, b.z, b.w};
      float wv[4 * DIN];                            // the 4 features' first-layer weights: DIN vector loads, not 4 DIN scalar ones
#pragma unroll
      for (int qv = 0; qv < DIN; ++qv) {
        const float4 t4 = lds4(smem + S::W1 + (f0 + j4) * DIN + 4 * qv);
        wv[4 * qv] = t4.x; wv[4 * qv + 1] = t4.y; wv[4 * qv + 2] = t4.z; wv[4 * qv + 3] = t4.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int dd = 0; dd < DIN; ++dd) acc[e] = fmaf(wv[e * DIN + dd], x[dd], acc[e]);
        h[j4 + e] = tanh_fast(acc[e]);
      }
    }
    wait_dw();                                      // the previous tile's dW MMAs are done with the transposed tiles
#pragma unroll
    for (int l = 2; l <= L; ++l) {
      publish_col(kblk + (l - 1) * kStackF, h);     // H_{l-1}^T
      store_a_half(h);                              // A operand of the recompute GEMM: this point's row -> its TMEM lane
      fence_before_sync();
      wg_sync();
      PACOH_UNIFORM_CTX();
      if (wt == 0) {
        fence_after_sync();
        const uint32_t b_hi = smem_u32(smem + S::B + (l - 2) * 4 * kHid * kHid);
        gemm128x32x32_3xtf32_ts(tmem_u, b_hi, b_hi + kHid * kHid * 4, bar_u);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      fence_after_sync();
      uint32_t v[16];
      tmem_ld16(lane_base + kTmemD + f0, v);
#pragma unroll
      for (int j4 = 0; j4 < kFH; j4 += 4) {
        const float4 b = lds4(smem + S::BH + (l - 2) * kHid + f0 + j4);
        h[j4] = tanh_fast(__uint_as_float(v[j4]) + b.x);
        h[j4 + 1] = tanh_fast(__uint_as_float(v[j4 + 1]) + b.y);
        h[j4 + 2] = tanh_fast(__uint_as_float(v[j4 + 2]) + b.z);
        h[j4 + 3] = tanh_fast(__uint_as_float(v[j4 + 3]) + b.w);
      }
    }
    // ---- output layer backward.  dWout[o][j] += sum_pt H_L[pt][j] dout[pt][o] as a lane-per-feature row sum over
    //      H_L^T, parked in the value rows of this warp's (idle) slice of the dA^T K-block
    park_col(h);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 hv = lds4(kblk + rrow + ((4 * (c0 + c)) ^ rsw));
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float4 dv = lds4(sw + S::DOUT + o * 32 + 4 * (c0 + c));
        accWo[o] += fmaf(hv.x, dv.x, fmaf(hv.y, dv.y, fmaf(hv.z, dv.z, hv.w * dv.w)));
      }
    }
    float da[kFH];
#pragma unroll
    for (int j4 = 0; j4 < kFH; j4 += 4) {
      float dh[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        const float4 w = lds4(smem + S::WOUT + o * kHid + f0 + j4);
        dh[0] = fmaf(w.x, dr[o], dh[0]); dh[1] = fmaf(w.y, dr[o], dh[1]); dh[2] = fmaf(w.z, dr[o], dh[2]); dh[3] = fmaf(w.w, dr[o], dh[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) da[j4 + e] = dh[e] * fmaf(-h[j4 + e], h[j4 + e], 1.0f);
    }
    __syncwarp();                                   // every lane is done with the H_L^T row sums
    // ---- hidden layers L..2
#pragma unroll
    for (int l = L; l >= 2; --l) {
      float* HTk = kblk + (l - 1) * kStackF;        // this warp's slice of the H_{l-1}^T stack
      wait_dw();
      publish_col(kblk, da);                        // dA_l^T
      store_a_half(da);
      fence_async_smem();                           // transposed tiles (generic-proxy stores) -> visible to the MMA
      fence_before_sync();
      wg_sync();
      PACOH_UNIFORM_CTX();
      if (wt == 0) {
        fence_after_sync();
        const uint32_t bt_hi = smem_u32(smem + S::B + (l - 2) * 4 * kHid * kHid + 2 * kHid * kHid);
        gemm128x32x32_3xtf32_ts(tmem_u, bt_hi, bt_hi + kHid * kHid * 4, bar_u);          // dH_{l-1} = dA_l W_l
      } else if (wt == 32) {                        // a second issuer: the weight-gradient GEMM, on its own barrier
        fence_after_sync();
        const uint32_t idesc = umma_idesc_tf32(64, 64);   // M = 64: exactly the stacked rows, no over-read of the A tile
        const uint64_t da0 = umma_desc_sw128(at_u), db0 = umma_desc_sw128(at_u + (l - 1) * kStackF * 4);
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)            // 8 points = 32 B inside the 128-byte swizzle row; K-block = 8 KB
            umma_tf32(tmem_u + kTmemDW, da0 + (uint64_t)(w * 512 + ks * 2), db0 + (uint64_t)(w * 512 + ks * 2), idesc, 1);
        umma_commit(bar_dw_u);
      }
      dw_pending = true;
      // ---- while the MMAs run: db_l[row] = row sum of dA_l^T over this lane's half of the warp's points
      {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 v4 = lds4(kblk + rrow + ((4 * (c0 + c)) ^ rsw));
          s += (v4.x + v4.y) + (v4.z + v4.w);
        }
        accb[l - 1] += s;
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      fence_after_sync();
      uint32_t v[16];
      tmem_ld16(lane_base + kTmemD + f0, v);
#pragma unroll
      for (int e = 0; e < kFH; ++e) {
        const float hp = HTk[(e >> 3) * 256 + (e & 7) * 32 + PACOH_XO(e & 7)];
        da[e] = __uint_as_float(v[e]) * fmaf(-hp, hp, 1.0f);
      }
    }
    // ---- first layer: db1, dW1 as lane-per-feature row sums over dA_1^T (parked in the dA^T K-block again)
    {
      wait_dw();                                    // the dW MMAs are done reading the dA^T stack
      park_col(da);
      __syncwarp();
      float s = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 v4 = lds4(kblk + rrow + ((4 * (c0 + c)) ^ rsw));
        s += (v4.x + v4.y) + (v4.z + v4.w);
#pragma unroll
        for (int dd = 0; dd < DIN; ++dd) {
          const float4 xv = lds4(sw + S::X + dd * 32 + 4 * (c0 + c));
          accW1[dd] += fmaf(v4.x, xv.x, fmaf(v4.y, xv.y, fmaf(v4.z, xv.z, v4.w * xv.w)));
        }
      }
      accb[0] += s;
      __syncwarp();                                 // before the next tile overwrites X / DOUT / the K-block
    }
  }

  // ---- reduce: warps in a fixed order into the padded layout (aliasing warpgroup 0's tiles), then un-pad to the flat order
  wait_dw();
  fence_after_sync();
  // combine the two point halves of every row sum; lanes 0-15 then hold row f0 + lane
#pragma unroll
  for (int l = 0; l < L; ++l) accb[l] += __shfl_xor_sync(0xffffffffu, accb[l], 16);
#pragma unroll
  for (int dd = 0; dd < DIN; ++dd) accW1[dd] += __shfl_xor_sync(0xffffffffu, accW1[dd], 16);
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
    accWo[o] += __shfl_xor_sync(0xffffffffu, accWo[o], 16);
    accbo[o] = warp_sum(accbo[o]);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  // every warp drops its contributions into ITS OWN padded copy (kWarps x R_PAD floats over the now idle tiles), then
  // each entry is summed over the warps in a fixed order: 3 CTA barriers instead of one per warp, bitwise repeatable
  float* sall = smem;
  for (int i = tid; i < kWarps * S::R_PAD; i += kThreads) sall[i] = 0.0f;
  __syncthreads();
  {
    float* mine = sall + warp * S::R_PAD;
    if (L > 1 && fh == 0) {
      // M = 64 accumulator: row i of the stacked product sits in TMEM lane (i / 16) * 32 + i % 16 (probed by
      // tools/ubench/tc_m64_probe.cu): quadrants 0, 1 hold the value rows j = 16 wq + lane (hi*hi | hi*lo), quadrants
      // 2, 3 the lo rows j = 16 (wq - 2) + lane (lo*hi); lanes 16-31 of every quadrant are unused.
      uint32_t v[32];
      tmem_ld32(lane_base + kTmemDW, v);                 // (.sync.aligned: the whole warp executes it)
      const int j = 16 * (wq & 1) + lane;
      float* dstw = mine + S::R_H + kHid + j * kHid;
      if (lane < 16) {
#pragma unroll
        for (int k = 0; k < kHid; ++k) dstw[k] = __uint_as_float(v[k]);
      }
      if (wq < 2) {
        tmem_ld32(lane_base + kTmemDW + 32, v);
        if (lane < 16) {
#pragma unroll
          for (int k = 0; k < kHid; ++k) dstw[k] += __uint_as_float(v[k]);
        }
      }
    }
    if (lane < 16) {
      const int row = f0 + lane;
#pragma unroll
      for (int l = 2; l <= L; ++l) mine[S::R_H + (l - 2) * S::R_HSTRIDE + row] = accb[l - 1];
      mine[S::R_B1 + row] = accb[0];
#pragma unroll
      for (int dd = 0; dd < DIN; ++dd) mine[S::R_W1 + row * DIN + dd] = accW1[dd];
#pragma unroll
      for (int o = 0; o < OUT; ++o) mine[S::R_WO + o * kHid + row] = accWo[o];
    }
    if (lane == 0 && fh == 0) {
#pragma unroll
      for (int o = 0; o < OUT; ++o) mine[S::R_BO + o] = accbo[o];
    }
  }
  __syncthreads();
  float* sacc = sall + kWarps * S::R_PAD;          // the summed vector, after the per-warp copies
  for (int i = tid; i < S::R_END; i += kThreads) {
    float sum = 0.0f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sum += sall[w * S::R_PAD + i];
    sacc[i] = sum;
  }
  __syncthreads();
  float* dst = a.partial[zn] + ((size_t)slot * a.P + p) * net.total;
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
  fence_before_sync();
  __syncthreads();               // the next segment restages the weights and reuses the tiles the reduction aliased
  fence_after_sync();
  }  // segments
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

template <int L, int DIN, int OUT>
int launch_tc_bwd(const MlpArgs& a, int chunks, int nets, cudaStream_t st) {
  using S = BwdSmem<L, DIN, OUT>;
  const size_t smem = sizeof(float) * S::END;
  // wide input / output layers (d > 2 together with out > 2 ...) do not leave room for three warpgroups of transposed
  // tiles in the 227 KB of shared memory: those shapes run on the CUDA-core kernel (mlp.cu)
  if (smem + 256 > 227 * 1024) return launch_mlp_fast(a, nets, chunks, true, st);
  int grid = 0, per_cta = 0;
  if (mlp_tc_bwd_slots(a.P, nets, a.T * a.n, &grid, &per_cta) > chunks) return PACOH_ERR_WORKSPACE;   // partial slots
  // unused slots of the partial buffers must read as zero in the fixed-order reduction
  for (int z = 0; z < nets; ++z)
    PACOH_CUDA_CHECK(cudaMemsetAsync(a.partial[z], 0, sizeof(float) * (size_t)chunks * a.P * a.net[z].total, st));
  PACOH_CUDA_CHECK(cudaFuncSetAttribute(mlp_tc_bwd_kernel<L, DIN, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp_tc_bwd_kernel<L, DIN, OUT><<<grid, kThreads, smem, st>>>(a, nets, per_cta);
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

// Persistent schedule of the tensor-core backward kernel: one CTA per SM, the nets * P * tiles tile space cut into equal
// ranges.  Returns the number of partial-gradient slots a (net, particle) can be split into (= CTAs touching it), which
// sizes the (slots, P, D_net) partial buffers.
int mlp_tc_bwd_slots(int P, int nets, int Q, int* grid_out, int* per_cta_out) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    else { sms = 148; cudaGetLastError(); }
  }
  // One CTA per SM would be ideal for the schedule, but a warpgroup sums ALL its tiles into one fp32 TMEM accumulator:
  // measured on config #4 (tests/manual/fullsize_fp64_check.py), 461 tiles per accumulator leave 6e-5 relative error in the
  // weight gradients, 115 tiles 1.7e-5.  So the launch is cut into `waves` equal CTAs per SM such that no warpgroup
  // accumulates more than kBwdMaxTilesPerAcc tiles (each extra wave costs ~18 us of CTA start / reduction time).
  static int waves_env = -1;
  if (waves_env < 0) {
    const char* e = getenv("PACOH_BWD_WAVES");
    waves_env = e != nullptr && atoi(e) > 0 ? atoi(e) : 0;
  }
  const long long tiles = (Q + kTile - 1) / kTile, total = (long long)nets * P * tiles;
  const long long per1 = (total + sms - 1) / sms;
  long long waves = waves_env > 0 ? waves_env : (per1 + kWG * kBwdMaxTilesPerAcc - 1) / (kWG * kBwdMaxTilesPerAcc);
  if (waves < 1) waves = 1;
  const long long ctas = (long long)sms * waves;
  long long per = (total + ctas - 1) / ctas;
  // at least two tiles per warpgroup when there is enough work to fill the machine anyway (below that the fixed per-CTA cost
  // dominates the throughput); small launches (configs #1 - #3: a few dozen tiles) spread over as many SMs as they have
  // tiles instead -- their cost is latency, and six (net, particle) segments in a row on one SM cost ~100 us
  if (per < 6 && total >= 6LL * sms) per = 6;
  if (per < 1) per = 1;
  const int grid = (int)((total + per - 1) / per);
  int slots = 1;
  for (long long pn = 0; pn < (long long)nets * P; ++pn) {
    const int segs = (int)(((pn + 1) * tiles - 1) / per - (pn * tiles) / per + 1);
    slots = segs > slots ? segs : slots;
  }
  if (grid_out) *grid_out = grid;
  if (per_cta_out) *per_cta_out = (int)per;
  return slots;
}

// Backward pass of the `nets` nets in a.net[] (same depth <= 3, widths <= 32, d <= 4, out <= 4) on tcgen05 + CUDA cores.
int launch_mlp_tc_bwd(const MlpArgs& a, int nets, int chunks, cudaStream_t st) {
  const int din = bwd_pad_pow2(a.d);
  int out_pad = 1;
  for (int z = 0; z < nets; ++z) out_pad = max(out_pad, bwd_pad_pow2(a.net[z].out_dim));
  switch (a.net[0].n_hidden) {
    case 1: return bwd_dispatch_din<1>(a, din, out_pad, chunks, nets, st);
    case 2: return bwd_dispatch_din<2>(a, din, out_pad, chunks, nets, st);
    default: return launch_mlp_fast(a, nets, chunks, true, st);   // deeper nets: the CUDA-core kernel (mlp.cu)
  }
  return PACOH_ERR_UNSUPPORTED;
}

}  // namespace pacoh
