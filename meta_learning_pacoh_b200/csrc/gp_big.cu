// Batched GP marginal log-likelihood with analytic gradient for LARGE tasks (64 < n <= 4096 points), sm_100a:
// a left-looking blocked Cholesky / triangular inverse / K^-1 pipeline whose 128 x 128 x 128 tile products run on the
// tensor cores (tcgen05, 3xTF32) with TMA-staged operands.  BASELINE config #5 (PACOH-MAP, 512-2048 points per task).
//
// Same mathematics and outputs as gp_mll_kernel / gp_tc_kernel (reference: meta_learn/random_gp.py:54-89,
// models.py:428-487, gpytorch ExactMarginalLogLikelihood at random_gp.py:83-85 / GPR_meta_mll.py:104-119 and the
// autograd reverse pass through the Cholesky):
//
//   Khat = (s k + (sigma^2 + jitter) I) / tot,  tot = s + sigma^2 + jitter        (unit diagonal, rho = s / tot)
//   Khat = L L^T                       left-looking, block column k:   C_ik = Khat_ik - sum_{j<k} L_ij L_kj^T
//                                      diagonal tile: big_kernel<M_DIAG> writes C_kk, big_potrf_kernel factorises and inverts
//                                      it in shared memory (potrf + trtri in place); panel: L_ik = C_ik L_kk^-T (fused solve)
//   U = L^-T                           U_ab = -(sum_{j=a}^{b-1} U_aj L_bj^T) L_bb^-T        (one launch per distance b - a)
//   Khat^-1 = U U^T                    lower block triangle only: tile (a, b <= a) = sum_{m >= a} U_am U_bm^T, contracted on the
//                                      fly with the Gram derivative (never stored):  w_rc = (alphahat_r alphahat_c / tot -
//                                      Khat^-1_rc) k_rc feeds the rows of block a directly and, for b < a, the rows of block b
//                                      through column sums
//   v = L^-1 r (block by block: the diagonal-tile kernels), alphahat = U v + one step of iterative refinement,
//   quad = r . alphahat / tot, log det = n log tot + 2 sum log L_ii
//
// Every tile product is the same persistent, warp-specialised kernel (big_kernel<MODE>, one CTA per SM, 212 KB shared memory,
// all 512 tensor-memory columns):
//   warp 8  (one lane)  TMA producer: cp.async.bulk.tensor boxes of 128 rows x 32 fp32 (SWIZZLE_128B) into a 4-stage ring
//   warps 0-7           converters: the A rows go to TENSOR MEMORY as tf32 hi / lo parts (tcgen05.st), the B tile's lo part
//                       is written beside the raw tile (which IS the hi operand: the tensor core truncates fp32 to tf32)
//   warp 9  (one lane)  issues D[128 x 128] (+)= A B^T as 3 x 4 tcgen05.mma.kind::tf32 per 32-wide K chunk (lo.hi, hi.lo, hi.hi)
//   warps 0-7 again     after every 128-wide K block the accumulator (two TMEM buffers, alternating) is added into fp32
//                       REGISTERS: the tensor core's accumulate truncates, so sums are kept short (48 MMAs) and the long
//                       sum over K blocks is rounded to nearest on the CUDA cores.
// The triangular solves with the diagonal tile are one more K block of the same pipeline whose A operand comes from the
// register accumulator (C -> hi / lo -> tensor memory) and whose B operand is the inverted diagonal tile; result tiles are
// stored straight from the registers.
#include <cuda.h>
#include <math_constants.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include "chol128.cuh"

namespace pacoh {

namespace {

using namespace tc;

constexpr int NB = 128;                 // tile edge
constexpr int KC = 32;                  // K chunk = one TMA box / one pipeline stage
constexpr int NSTAGE = 4;
constexpr int kConv = 256;              // converter / flush / epilogue threads (warps 0-7)
constexpr int kBigThreads = 320;        // + warp 8 (TMA producer) + warp 9 (MMA issuer, tensor-memory owner)
constexpr float kFarB = 1.0e18f;
constexpr float kCB = 0.84932180028801904272f;   // sqrt(0.5 * log2(e))

constexpr uint32_t kStageBytes = 3 * 16384;                // raw A | raw B | lo B
constexpr uint32_t kOffTile = NSTAGE * kStageBytes;        // 196608: scratch of the epilogues (column sums, row exchange)
constexpr uint32_t kTileBytes = 2048 * 4 + 1024 * 4 + 256;
constexpr uint32_t kOffUcol = kOffTile + kTileBytes;       // float4[128]
constexpr uint32_t kOffAcol = kOffUcol + 2048;             // float[128]
constexpr uint32_t kOffMisc = kOffAcol + 512;              // float[128] scratch
constexpr uint32_t kOffBar = kOffMisc + 512;               // mbarriers
constexpr uint32_t kBigSmem = kOffBar + 256;

constexpr uint32_t kTmemAcol = 256;     // D0 [0,128) | D1 [128,256) | A stages [256 + 64 s): hi 32 | lo 32

enum { M_DIAG = 0, M_PANEL = 1, M_UINV = 2, M_GRAD = 3 };

__device__ __forceinline__ float ex2b(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 3xTF32 operand split with round-to-nearest parts: v = hi + lo + O(2^-23 |v|), errors symmetric around zero (a
// truncating split leaves every operand short by up to 2^-21 |v| with the sign of v, a bias that sums coherently over
// K and is what trace-like contractions of Khat^-1 are most sensitive to).
// (cvt.rna.tf32.f32 is emulated with ~5 instructions on sm_100; adding half an ulp before the mask does the same in 2 for
// finite values.  A lo part only needs the "+ half ulp": the tensor core's own truncation completes the rounding.)
__device__ __forceinline__ float rn_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float pre_round(float v) { return __uint_as_float(__float_as_uint(v) + 0x1000u); }
// B operand: the raw fp32 tile is the hi part (the tensor core truncates it), so lo = v - trunc(v), rounded to tf32
__device__ __forceinline__ float lo_of_raw(float v) { return pre_round(v - tf32_hi(v)); }

// Khat entry: rho 2^(-|du|^2) (e0 = log2 rho, or -inf for a padding row); one definition so that the factorisation and
// the residual of the iterative refinement see bit-identical matrices
__device__ __forceinline__ float khat_entry(const float4& ur, const float4& uc, float e0) {
  float e = e0;
  { const float du = ur.x - uc.x; e = fmaf(-du, du, e); }
  { const float du = ur.y - uc.y; e = fmaf(-du, du, e); }
  { const float du = ur.z - uc.z; e = fmaf(-du, du, e); }
  { const float du = ur.w - uc.w; e = fmaf(-du, du, e); }
  return ex2b(e);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

}  // namespace

// Per-matrix state shared by all launches of one pass.
struct BigMat {
  float tot, rho, lg2rho, pad0;
  int n, nb, lvl, fail;        // points, tiles per side, jitter level, "a pivot was not positive" flag of the current attempt
  int status, src, p, t;       // final status (jitter level used or -1), source task, particle, batch slot
};

struct BigArgs {
  int B, nb_max, npad;
  float* Lbuf;                 // (B, npad, npad)  lower block triangle: L; strictly upper block triangle: U = L^-T
  float* SB;                   // (B, nb_max, 2, 128, 128)  [0] = L_kk^-1, [1] = U_kk = L_kk^-T
  float4* ubuf;                // (B, npad)  scaled features (zero-padded to 4), padding rows far away
  float* rbuf;                 // (B, npad)  residuals y - m (0 on padding rows)
  float* vbuf;                 // (B, npad)  v = L^-1 r
  float* abuf;                 // (B, npad)  alphahat = Khat^-1 r
  float* wbuf;                 // (B, npad)  scratch of the iterative refinement
  float* ldet;                 // (B, nb_max) sum of log2 L_ii per diagonal tile
  float* part;                 // (B, nb_max, 12) per-tile partial sums: quad tot, -, trace, sum beta, S2[4], sum S1[4]
  float* rowpart;              // (B, npad, 8)   per-row sums over the tiles of the row's own block row (+ diagonal of Khat^-1)
  float* colpart;              // (B, npairs, 128, 4) column sums of the tiles (a, b), b < a: contributions to the rows of block b
  int npairs;                  // nb_max (nb_max - 1) / 2
  BigMat* mat;                 // (B)
  const int* list;             // retry passes: indices of the matrices to redo (nullptr: all B matrices)
  const int* count;            //               and how many
  GpArgs g;                    // outputs (dmean, dfeat, mll, dhyp, info) and the hyper-parameter offsets
};

namespace {

// ---------------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kBigThreads, 1)
big_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapS, const BigArgs a, const int step) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  // barriers: full_raw[NSTAGE] | conv_done[NSTAGE] | mma_done[NSTAGE] | d_full[2] | d_free[2] | (spare)
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto CONV = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto MMAD = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto DFULL = [&](int d) { return bar0 + 8u * (3 * NSTAGE + d); };
  auto DFREE = [&](int d) { return bar0 + 8u * (3 * NSTAGE + 2 + d); };
  const uint32_t EPI = bar0 + 8u * (3 * NSTAGE + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 8 * (3 * NSTAGE + 5));
  if ((smem_u32(smem) & 1023u) != 0) __trap();      // SWIZZLE_128B tiles need a 1024-byte aligned base

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(FULL(s), 1); mbar_init(CONV(s), kConv); mbar_init(MMAD(s), 1); }
    for (int d = 0; d < 2; ++d) { mbar_init(DFULL(d), 1); mbar_init(DFREE(d), kConv); }
    mbar_init(EPI, kConv);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // ---- work decomposition, identical for every role
  const int per = MODE == M_DIAG ? 1 : (MODE == M_PANEL ? a.nb_max - step - 1 : (MODE == M_UINV ? a.nb_max - step : a.nb_max));
  const int nmat = a.list != nullptr ? __ldg(a.count) : a.B;
  const int total = nmat * per;
  constexpr bool kSolve = MODE == M_PANEL || MODE == M_UINV;
  constexpr bool kSameAB = MODE == M_DIAG;

  struct Item { int m, ti, tj, nbm; bool skip; };
  auto decode = [&](int item) {
    Item it;
    const int ml = item / per, tl = item - ml * per;
    it.m = a.list != nullptr ? __ldg(a.list + ml) : ml;
    it.nbm = a.mat[it.m].nb;
    if (MODE == M_DIAG) { it.ti = it.tj = step; it.skip = step >= it.nbm; }
    else if (MODE == M_PANEL) { it.ti = step + 1 + tl; it.tj = step; it.skip = it.ti >= it.nbm; }
    else if (MODE == M_UINV) { it.ti = tl; it.tj = tl + step; it.skip = it.tj >= it.nbm; }
    else { it.ti = tl; it.tj = 0; it.skip = it.ti >= it.nbm; }
    return it;
  };
  // number of sub-tiles of an item and K blocks of a sub-tile
  // GRAD: block row a computes the tiles (a, b) for b <= a only (Khat^-1 is symmetric): every tile feeds the rows of
  // block a directly and, for b < a, the rows of block b through column sums
  auto nsub_of = [&](const Item& it) { return MODE == M_GRAD ? it.ti + 1 : 1; };
  auto nblk_of = [&](const Item& it, int sub) {
    if (MODE == M_DIAG || MODE == M_PANEL) return step;
    if (MODE == M_UINV) return step;
    return it.nbm - it.ti;
  };

  if (warp == 8) {
    // =============================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0, nitem = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const Item it = decode(item);
        if (it.skip) continue;
        ++nitem;
        const int sbz = (it.m * a.nb_max) * 2;
        const int ns = nsub_of(it);
        for (int sub = 0; sub < ns; ++sub) {
          const int nblk = nblk_of(it, sub);
          for (int q = 0; q < nblk + (kSolve ? 1 : 0); ++q) {
            const bool solve = kSolve && q == nblk;
            // operand tiles of this K block
            bool a_sb = false, b_sb = false;
            int a_row = 0, a_col = 0, a_z = 0, b_row = 0, b_col = 0, b_z = 0;
            if (solve) {
              b_sb = true; b_z = sbz + 2 * it.tj;                              // L_jj^-1 (row-major)
            } else if (MODE == M_DIAG) {
              b_row = step * NB; b_col = q * NB;
            } else if (MODE == M_PANEL) {
              a_row = it.ti * NB; a_col = q * NB; b_row = step * NB; b_col = q * NB;
            } else if (MODE == M_UINV) {
              const int j = it.ti + q;
              if (j == it.ti) { a_sb = true; a_z = sbz + 2 * it.ti + 1; } else { a_row = it.ti * NB; a_col = j * NB; }
              b_row = it.tj * NB; b_col = j * NB;
            } else {
              const int mm = it.ti + q;
              if (mm == it.ti) { a_sb = true; a_z = sbz + 2 * it.ti + 1; } else { a_row = it.ti * NB; a_col = mm * NB; }
              if (mm == sub) { b_sb = true; b_z = sbz + 2 * sub + 1; } else { b_row = sub * NB; b_col = mm * NB; }
            }
            for (int c = 0; c < 4; ++c, ++g) {
              const uint32_t s = g % NSTAGE;
              if (g >= NSTAGE) mbar_wait(MMAD(s), ((g / NSTAGE) - 1) & 1);
              const uint32_t sbase = smem_u32(smem) + s * kStageBytes;
              const bool loadA = !solve && !kSameAB;
              mbar_expect_tx(FULL(s), loadA ? 32768u : 16384u);
              if (loadA) {
                if (a_sb) tma_load_3d(sbase, &mapS, c * KC, 0, a_z, FULL(s));
                else tma_load_3d(sbase, &mapL, a_col + c * KC, a_row, it.m, FULL(s));
              }
              if (b_sb) tma_load_3d(sbase + 16384, &mapS, c * KC, 0, b_z, FULL(s));
              else tma_load_3d(sbase + 16384, &mapL, b_col + c * KC, b_row, it.m, FULL(s));
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      uint32_t g = 0, bb = 0;
      const uint32_t idesc = umma_idesc_tf32(128, 128);
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const Item it = decode(item);
        if (it.skip) continue;
        const int ns = nsub_of(it);
        for (int sub = 0; sub < ns; ++sub) {
          const int nblk = nblk_of(it, sub) + (kSolve ? 1 : 0);
          for (int q = 0; q < nblk; ++q, ++bb) {
            const uint32_t d = bb & 1;
            if (bb >= 2) mbar_wait(DFREE(d), ((bb >> 1) - 1) & 1);
            fence_after_sync();
            const uint32_t dcol = tmem + d * 128;
            for (int c = 0; c < 4; ++c, ++g) {
              const uint32_t s = g % NSTAGE;
              mbar_wait(CONV(s), (g / NSTAGE) & 1);
              fence_after_sync();
              const uint32_t sbase = smem_u32(smem) + s * kStageBytes;
              const uint64_t bhi = desc_sw128(sbase + 16384), blo = desc_sw128(sbase + 32768);
              const uint32_t ahi = tmem + kTmemAcol + 64 * s, alo = ahi + 32;
#pragma unroll
              for (int ps = 0; ps < 3; ++ps) {          // lo.hi, hi.lo, hi.hi
                const uint32_t acol = ps == 0 ? alo : ahi;
                const uint64_t bd = ps == 1 ? blo : bhi;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_tf32_ts(dcol, acol + ks * 8, bd + (uint64_t)(ks * 2), idesc, (c | ps | ks) != 0 ? 1u : 0u);
              }
              umma_commit(MMAD(s));
            }
            umma_commit(DFULL(d));
          }
        }
      }
    }
  } else {
    // =============================================================== converters / accumulators / epilogue (256 threads)
    const int r = tid & 127, h = tid >> 7;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* T = reinterpret_cast<float*>(smem + kOffTile);
    float4* ucol = reinterpret_cast<float4*>(smem + kOffUcol);
    float* acolv = reinterpret_cast<float*>(smem + kOffAcol);
    float* misc = reinterpret_cast<float*>(smem + kOffMisc);
    uint32_t g = 0, bb = 0;
    float acc[64];
    int swz[4];                                             // byte offsets of this thread's four 16-byte pieces in a SWIZZLE_128B tile
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) swz[j4] = r * 128 + (((4 * h + j4) ^ (r & 7)) << 4);

    auto flush = [&](uint32_t b) {       // acc += accumulator of K block b (this thread's 64 columns of its row)
      const uint32_t d = b & 1;
      mbar_wait(DFULL(d), (b >> 1) & 1);
      fence_after_sync();
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        uint32_t v[32];
        tmem_ld32(lane_base + d * 128 + 64 * h + 32 * q2, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[32 * q2 + i] += __uint_as_float(v[i]);
      }
      fence_before_sync();
      mbar_arrive(DFREE(d));
    };

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const Item it = decode(item);
      if (it.skip) continue;
      const BigMat mt = a.mat[it.m];
      const size_t vecoff = (size_t)it.m * a.npad;
      const int row = it.ti * NB + r;                       // this thread's row of the matrix
      const bool rvalid = row < mt.n;
      const float4 urow = a.ubuf[vecoff + row];
      // GRAD: per-row accumulators over the whole block row
      float S1[4] = {0.0f, 0.0f, 0.0f, 0.0f}, trc = 0.0f, alpha_r = 0.0f, beta_r = 0.0f;
      if (MODE == M_GRAD) { alpha_r = a.abuf[vecoff + row]; beta_r = alpha_r / mt.tot; }
      const float inv_tot = 1.0f / mt.tot;
      float tdot = 0.0f;                                    // DIAG: sum_j L_kj v_j for this row (this thread's k-half)
      const int ns = nsub_of(it);
      for (int sub = 0; sub < ns; ++sub) {
        const int nblk = nblk_of(it, sub);
        const int ctile = MODE == M_GRAD ? sub : it.tj;     // column tile whose features / alphas the epilogue needs
        if (MODE != M_UINV) {
          conv_sync();                                      // previous users of ucol / acolv / T are done
          if (tid < NB) {
            ucol[tid] = a.ubuf[vecoff + ctile * NB + tid];
            if (MODE == M_GRAD) acolv[tid] = a.abuf[vecoff + ctile * NB + tid];
          }
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
        for (int q = 0; q < nblk; ++q, ++bb) {
          const float* vsrc = MODE == M_DIAG ? a.vbuf + vecoff + q * NB + 16 * h : nullptr;
          for (int c = 0; c < 4; ++c, ++g) {
            const uint32_t s = g % NSTAGE;
            mbar_wait(FULL(s), (g / NSTAGE) & 1);
            const uint8_t* st = smem + s * kStageBytes;
            const uint8_t* rawA = st + (kSameAB ? 16384 : 0);
            const uint8_t* rawB = st + 16384;
            uint8_t* loB = smem + s * kStageBytes + 32768;
            // A: this thread's row, k-half h (16 of the chunk's 32 values) -> hi / lo -> tensor memory
            {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 v = *reinterpret_cast<const float4*>(rawA + swz[j4]);
                const float vv[4] = {v.x, v.y, v.z, v.w};
                if (MODE == M_DIAG) {
                  const float4 w = __ldg(reinterpret_cast<const float4*>(vsrc + c * KC + 4 * j4));
                  tdot = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, tdot))));
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float hv = rn_tf32(vv[e]);
                  hi[4 * j4 + e] = __float_as_uint(hv);
                  lo[4 * j4 + e] = __float_as_uint(vv[e] - hv);
                }
              }
              const uint32_t ab = lane_base + kTmemAcol + 64 * s + 16 * h;
              tmem_st16(ab, hi);
              tmem_st16(ab + 32, lo);
            }
            // B: lo part of row r, same k-half, written beside the raw tile at the same swizzled position
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int off = swz[j4];
              const float4 v = *reinterpret_cast<const float4*>(rawB + off);
              *reinterpret_cast<float4*>(loB + off) = make_float4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w));
            }
            tmem_st_wait();
            fence_before_sync();
            fence_async_smem();
            mbar_arrive(CONV(s));
          }
          if (q > 0) flush(bb - 1);
        }
        if (nblk > 0) flush(bb - 1);

        // ---------------------------------------------------------------- C = base - acc, then the fused triangular solve
        if (MODE != M_UINV) conv_sync();                  // ucol / acolv of this sub-tile are visible
        if (MODE == M_DIAG || MODE == M_PANEL) {
          // base = Khat tile: rho 2^(-|du|^2) off the diagonal, 1 on it, 0 against padding rows
          const float e0 = rvalid ? mt.lg2rho : -CUDART_INF_F;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const int cc = 64 * h + i;
            float kv = khat_entry(urow, ucol[cc], e0);
            if (MODE == M_DIAG && cc == r) kv = 1.0f;
            acc[i] = kv - acc[i];
          }
        }
        if (MODE == M_UINV) {
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = -acc[i];
        }
        if (kSolve) {
          // one more K block: A = C (from registers), B = L_jj^-1 chunks
#pragma unroll
          for (int c = 0; c < 4; ++c, ++g) {
            const uint32_t s = g % NSTAGE;
            mbar_wait(FULL(s), (g / NSTAGE) & 1);
            const uint8_t* rawB = smem + s * kStageBytes + 16384;
            uint8_t* loB = smem + s * kStageBytes + 32768;
            if (h == (c >> 1)) {
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float v = acc[32 * (c & 1) + 16 * q2 + e], hv = rn_tf32(v);
                  hi[e] = __float_as_uint(hv);
                  lo[e] = __float_as_uint(v - hv);
                }
                const uint32_t ab = lane_base + kTmemAcol + 64 * s + 16 * q2;
                tmem_st16(ab, hi);
                tmem_st16(ab + 32, lo);
              }
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int off = swz[j4];
              const float4 v = *reinterpret_cast<const float4*>(rawB + off);
              *reinterpret_cast<float4*>(loB + off) = make_float4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w));
            }
            tmem_st_wait();
            fence_before_sync();
            fence_async_smem();
            mbar_arrive(CONV(s));
          }
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
          flush(bb);
          ++bb;
          // result tile -> global straight from the registers: this thread owns 64 consecutive floats of row r (a warp's store
          // covers 32 rows x 16 bytes; the L2 merges the sector halves).  No staging tile: its 66 KB buy a fourth pipeline stage.
          {
            float* dst = a.Lbuf + ((size_t)it.m * a.npad + (size_t)it.ti * NB + r) * a.npad + (size_t)it.tj * NB + 64 * h;
#pragma unroll
            for (int i4 = 0; i4 < 16; ++i4)
              *reinterpret_cast<float4*>(dst + 4 * i4) = make_float4(acc[4 * i4], acc[4 * i4 + 1], acc[4 * i4 + 2], acc[4 * i4 + 3]);
          }
        }

        if (MODE == M_DIAG) {
          // ------------------------------------------------------------ diagonal tile: write C_kk = Khat_kk - sum_j L_kj L_kj^T
          //      and rhs_k = r_k - sum_j L_kj v_j; big_potrf_kernel factorises it (3 CTAs per SM: its serial sections
          //      would otherwise stall this kernel's tensor pipeline)
          conv_sync();
          if (h == 1) misc[r] = tdot;
          conv_sync();
          if (h == 0) a.vbuf[vecoff + row] = a.rbuf[vecoff + row] - (tdot + misc[r]);
          float* dst = a.Lbuf + ((size_t)it.m * a.npad + (size_t)step * NB + r) * a.npad + (size_t)step * NB + 64 * h;
#pragma unroll
          for (int i4 = 0; i4 < 16; ++i4)
            *reinterpret_cast<float4*>(dst + 4 * i4) = make_float4(acc[4 * i4], acc[4 * i4 + 1], acc[4 * i4 + 2], acc[4 * i4 + 3]);
        }

        if (MODE == M_GRAD) {
          // ------------------------------------------------------------ gradient contraction with the tile of Khat^-1:
          //   w_rc = (alphahat_r alphahat_c / tot - Khat^-1_rc) k_rc          (bit-symmetric in r <-> c)
          //   rows of block a:     S1_r += sum_c w_rc (u_r - u_c)
          //   rows of block b < a: S1_c += sum_r w_rc (u_c - u_r)             (column sums of the same products)
          const bool offdiag = sub < it.ti;
          const int F = a.g.F;
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            float x0[32], x1[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int cc = 64 * h + 32 * g2 + i;
              const float4 uc = ucol[cc];
              const float d0 = urow.x - uc.x, d1 = urow.y - uc.y, d2 = urow.z - uc.z, d3 = urow.w - uc.w;
              const float e = fmaf(-d0, d0, fmaf(-d1, d1, fmaf(-d2, d2, -d3 * d3)));
              const float wv = fmaf(alpha_r * acolv[cc], inv_tot, -acc[32 * g2 + i]) * ex2b(e);
              S1[0] = fmaf(wv, d0, S1[0]); S1[1] = fmaf(wv, d1, S1[1]); S1[2] = fmaf(wv, d2, S1[2]); S1[3] = fmaf(wv, d3, S1[3]);
              x0[i] = wv * d0; x1[i] = wv * d1;
              if (!offdiag && cc == r) trc = fmaf(alpha_r * alpha_r, inv_tot, -acc[32 * g2 + i]);
            }
            if (offdiag) {
              // column sums over the warp's 32 rows: transpose-reduce (31 shuffles per 32 columns), lane l <- column l
#pragma unroll
              for (int f = 0; f < 4; ++f) {
                if (f >= F) break;
                float v[32];
                if (f < 2) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = f == 0 ? x0[i] : x1[i];
                } else {        // features 2, 3 (rare): recompute the products
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const int cc = 64 * h + 32 * g2 + i;
                    const float4 uc = ucol[cc];
                    const float d0 = urow.x - uc.x, d1 = urow.y - uc.y, d2 = urow.z - uc.z, d3 = urow.w - uc.w;
                    const float e = fmaf(-d0, d0, fmaf(-d1, d1, fmaf(-d2, d2, -d3 * d3)));
                    const float wv = fmaf(alpha_r * acolv[cc], inv_tot, -acc[32 * g2 + i]) * ex2b(e);
                    v[i] = wv * (f == 2 ? d2 : d3);
                  }
                }
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) {
                  const bool up = (lane & sft) != 0;
#pragma unroll
                  for (int i = 0; i < sft; ++i) {
                    const float send = up ? v[i] : v[i + sft];
                    const float keep = up ? v[i + sft] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
                  }
                }
                // T scratch: [warp][g2][f][lane]
                T[((warp * 2 + g2) * 4 + f) * 32 + lane] = v[0];
              }
            }
          }
          if (offdiag) {
            conv_sync();
            // column c = 64 h' + 32 g2 + l of tile `sub`: sum over the four warps with the same h' (fixed order), negated
            for (int o = tid; o < NB * 4; o += kConv) {
              const int cc = o >> 2, f = o & 3;
              float sv = 0.0f;
              if (f < F) {
                const int hh = cc >> 6, g2 = (cc >> 5) & 1, l = cc & 31;
                const float* src = T + ((hh * 4 * 2 + g2) * 4 + f) * 32 + l;
                sv = -((src[0] + src[2 * 4 * 32]) + (src[2 * 2 * 4 * 32] + src[3 * 2 * 4 * 32]));
              }
              a.colpart[(((size_t)it.m * a.npairs + (size_t)it.ti * (it.ti - 1) / 2 + sub) * NB + cc) * 4 + f] = sv;
            }
          }
        }
      }

      if (MODE == M_GRAD) {
        // the row sums of this block row (two column halves per row) and the diagonal of Khat^-1
        conv_sync();
        if (h == 1) { float* o = T + 2048 + r * 8; o[1] = S1[0]; o[2] = S1[1]; o[3] = S1[2]; o[4] = S1[3]; o[5] = trc; }
        conv_sync();
        if (h == 0) {
          const float* o = T + 2048 + r * 8;
          float* dst = a.rowpart + (vecoff + row) * 8;
          *reinterpret_cast<float4*>(dst) = make_float4(S1[0] + o[1], S1[1] + o[2], S1[2] + o[3], S1[3] + o[4]);
          dst[4] = trc + o[5];
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// Diagonal tile k of every matrix: C_kk (written by big_kernel<M_DIAG>) -> L_kk, L_kk^-1, U_kk = L_kk^-T, v_k = L_kk^-1 rhs_k,
// sum log2 L_ii.  One CTA (256 threads, one 67.5 KB tile: L below the diagonal, U above it) per matrix, three CTAs per SM:
// the factorisation is a chain of short serial sections (32-column panels) that only other resident CTAs can hide.
__global__ void __launch_bounds__(256, 3) big_potrf_kernel(BigArgs a, int step) {
  extern __shared__ __align__(16) float smp[];
  float* T = smp;
  float* rhs = smp + NB * LDT;
  float* red = rhs + NB;
  float* flag = red + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nmat = a.list != nullptr ? __ldg(a.count) : a.B;
  for (int ml = blockIdx.x; ml < nmat; ml += gridDim.x) {
    const int m = a.list != nullptr ? __ldg(a.list + ml) : ml;
    if (step >= a.mat[m].nb) continue;
    const size_t vecoff = (size_t)m * a.npad;
    float* tileL = a.Lbuf + ((size_t)m * a.npad + (size_t)step * NB) * a.npad + (size_t)step * NB;
    conv_sync();
    for (int rr = warp; rr < NB; rr += 8) sts4(T + rr * LDT + 4 * lane, *reinterpret_cast<const float4*>(tileL + (size_t)rr * a.npad + 4 * lane));
    if (tid < NB) rhs[tid] = a.vbuf[vecoff + step * NB + tid];
    const bool ok = potrf_trtri_128_inplace(T, tid, flag);
    if (!ok && tid == 0) atomicExch(&a.mat[m].fail, 1);
    if (tid < NB) {                     // v_k[c] = sum_{r <= c} U[r][c] rhs_r ; log2 det
      const int c = tid;
      const float ucc = 1.0f / T[c * LDT + c];
      float vk = ucc * rhs[c];
      for (int rr = 0; rr < c; ++rr) vk = fmaf(T[rr * LDT + c], rhs[rr], vk);
      a.vbuf[vecoff + step * NB + c] = vk;
      float lg = log2f(T[c * LDT + c]);
      lg = warp_sum(lg);
      if (lane == 0) red[warp] = lg;
    }
    conv_sync();
    if (tid == 0) a.ldet[(size_t)m * a.nb_max + step] = (red[0] + red[1]) + (red[2] + red[3]);
    // L_kk (zeros above the diagonal) -> Lbuf ; U_kk -> SB[.][k][1] ; L_kk^-1 = U_kk^T -> SB[.][k][0]
    float* dstI = a.SB + ((size_t)m * a.nb_max + step) * 2 * NB * NB;
    float* dstU = dstI + NB * NB;
    for (int rr = warp; rr < NB; rr += 8) {
      const float4 t4 = lds4(T + rr * LDT + 4 * lane);
      const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
      const float inv_d = 1.0f / T[rr * LDT + rr];
      float lo[4], up[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cc = 4 * lane + e;
        lo[e] = cc <= rr ? tv[e] : 0.0f;
        up[e] = cc > rr ? tv[e] : (cc == rr ? inv_d : 0.0f);
      }
      *reinterpret_cast<float4*>(tileL + (size_t)rr * a.npad + 4 * lane) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<float4*>(dstU + rr * NB + 4 * lane) = make_float4(up[0], up[1], up[2], up[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {      // row rr of L^-1 = column rr of U
        const int cr = lane + 32 * e;    // U[cr][rr]
        dstI[rr * NB + cr] = cr < rr ? T[cr * LDT + rr] : (cr == rr ? inv_d : 0.0f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Small CUDA-core kernels around the tile pipeline.

// One CTA per matrix: scaled features / residuals / per-matrix constants.  grid = B, 128 threads.
__global__ void big_prep_kernel(BigArgs a, int b0) {
  const GpArgs& g = a.g;
  const int ml = blockIdx.x, gm = b0 + ml;                 // gm = p * T + t
  const int p = gm / g.T, t = gm - p * g.T;
  const int src = __ldg(g.task_idx + t);
  const int n = g.task_n != nullptr ? __ldg(g.task_n + src) : g.n;
  const float* th = g.theta + (size_t)p * g.D;
  if (threadIdx.x == 0) {
    BigMat mt;
    const float sig2 = g.noise_floor + softplus_f(__ldg(th + g.off_noise));
    const float osc = g.has_oscale ? softplus_f(__ldg(th + g.off_oscale)) : 1.0f;
    mt.tot = osc + sig2; mt.rho = osc / mt.tot; mt.lg2rho = log2f(mt.rho); mt.pad0 = 0.0f;
    mt.n = n; mt.nb = (n + NB - 1) / NB; mt.lvl = 0; mt.fail = 0; mt.status = 0; mt.src = src; mt.p = p; mt.t = t;
    a.mat[ml] = mt;
  }
  float sc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int f = 0; f < g.F && f < 4; ++f) sc[f] = kCB / softplus_f(__ldg(th + g.off_ls + f));
  const float cm = g.mean_kind == PACOH_MEAN_CONSTANT ? __ldg(th + g.off_const_mean) : 0.0f;
  const size_t Q = (size_t)g.T * g.n;
  for (int row = threadIdx.x; row < a.npad; row += blockDim.x) {
    float4 u = make_float4(kFarB, kFarB, kFarB, kFarB);
    float res = 0.0f;
    if (row < n) {
      const size_t q = (size_t)p * Q + (size_t)t * g.n + row;
      float f4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int f = 0; f < g.F && f < 4; ++f)
        f4[f] = (g.feat != nullptr ? __ldg(g.feat + q * g.F + f) : __ldg(g.x + ((size_t)src * g.n + row) * g.d + f)) * sc[f];
      u = make_float4(f4[0], f4[1], f4[2], f4[3]);
      const float m = g.mean != nullptr ? __ldg(g.mean + q) : cm;
      res = __ldg(g.y + (size_t)src * g.n + row) - m;
    }
    a.ubuf[(size_t)ml * a.npad + row] = u;
    a.rbuf[(size_t)ml * a.npad + row] = res;
  }
}

// After a factorisation attempt: matrices whose attempt failed move one jitter level up (gpytorch psd_safe_cholesky:
// 1e-6, 1e-5, 1e-4) and are appended to the retry list; matrices that succeeded keep status = level used.
__global__ void big_retry_kernel(BigArgs a, int* list_out, int* count_out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.B) return;
  BigMat& mt = a.mat[m];
  if (mt.status < 0 || !mt.fail) return;                 // gave up earlier / this attempt succeeded: status == lvl stays
  if (mt.lvl >= 3) { mt.status = -1; return; }
  mt.lvl += 1;
  mt.status = mt.lvl;
  mt.fail = 0;
  const float jit = mt.lvl == 1 ? 1e-6f : (mt.lvl == 2 ? 1e-5f : 1e-4f);
  const float* th = a.g.theta + (size_t)mt.p * a.g.D;
  const float sig2 = a.g.noise_floor + softplus_f(__ldg(th + a.g.off_noise));
  const float osc = a.g.has_oscale ? softplus_f(__ldg(th + a.g.off_oscale)) : 1.0f;
  mt.tot = osc + sig2 + jit; mt.rho = osc / mt.tot; mt.lg2rho = log2f(mt.rho);
  list_out[atomicAdd(count_out, 1)] = m;
}

// y (+)= U x: one CTA per (matrix, block row a), warp per row, lanes over the columns >= the row's block.
// alphahat = U v, and the correction U (U^T r') of the iterative refinement.
__global__ void big_ux_kernel(BigArgs a, const float* __restrict__ xin, float* __restrict__ yout, int add) {
  const int m = blockIdx.x / a.nb_max, ta = blockIdx.x % a.nb_max;
  const BigMat mt = a.mat[m];
  if (ta >= mt.nb) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* v = xin + (size_t)m * a.npad;
  const float* Ud = a.SB + (((size_t)m * a.nb_max + ta) * 2 + 1) * NB * NB;
  for (int rr = warp; rr < NB; rr += nw) {
    const int row = ta * NB + rr;
    double s = 0.0;
    for (int c = lane; c < NB; c += 32) s += (double)(Ud[rr * NB + c] * v[ta * NB + c]);
    const float* Urow = a.Lbuf + ((size_t)m * a.npad + row) * a.npad;
    for (int c = (ta + 1) * NB + lane; c < mt.nb * NB; c += 32) s += (double)(Urow[c] * v[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      float* y = yout + (size_t)m * a.npad + row;
      *y = add ? (float)((double)*y + s) : (float)s;
    }
  }
}

// w = U^T x = L^-1 x: one CTA per (matrix, block column b), thread per column, rows <= the column.
__global__ void big_utx_kernel(BigArgs a, const float* __restrict__ xin, float* __restrict__ wout) {
  __shared__ float xs[NB];
  const int m = blockIdx.x / a.nb_max, tb = blockIdx.x % a.nb_max;
  const BigMat mt = a.mat[m];
  if (tb >= mt.nb) return;
  const int c = threadIdx.x;
  double s = 0.0;
  for (int ta = 0; ta <= tb; ++ta) {
    __syncthreads();
    xs[c] = xin[(size_t)m * a.npad + ta * NB + c];
    __syncthreads();
    if (ta < tb) {
      const float* U = a.Lbuf + ((size_t)m * a.npad + (size_t)ta * NB) * a.npad + (size_t)tb * NB + c;
      for (int rr = 0; rr < NB; ++rr) s += (double)(U[(size_t)rr * a.npad] * xs[rr]);
    } else {
      const float* Ud = a.SB + (((size_t)m * a.nb_max + tb) * 2 + 1) * NB * NB + c;
      for (int rr = 0; rr <= c; ++rr) s += (double)(Ud[rr * NB] * xs[rr]);
    }
  }
  wout[(size_t)m * a.npad + tb * NB + c] = (float)s;
}

// r' = r - Khat alphahat with Khat regenerated from the features (bit-identical to the factorised matrix), fp64
// accumulation: one CTA per (matrix, block row), thread per row.  One step of iterative refinement brings alphahat
// from the accuracy of the explicit inverse U (1e-5 .. 1e-4 at n = 2048) to that of a backward-stable fp32 solve.
__global__ void big_resid_kernel(BigArgs a, float* __restrict__ rout) {
  __shared__ float4 us[NB];
  __shared__ float as[NB];
  const int m = blockIdx.x / a.nb_max, ta = blockIdx.x % a.nb_max;
  const BigMat mt = a.mat[m];
  if (ta >= mt.nb) return;
  const int tid = threadIdx.x, row = ta * NB + tid;
  const size_t vecoff = (size_t)m * a.npad;
  const float4 ur = a.ubuf[vecoff + row];
  const bool rvalid = row < mt.n;
  const float e0 = rvalid ? mt.lg2rho : -CUDART_INF_F;
  double s = 0.0;
  for (int tb = 0; tb < mt.nb; ++tb) {
    __syncthreads();
    us[tid] = a.ubuf[vecoff + tb * NB + tid];
    as[tid] = a.abuf[vecoff + tb * NB + tid];
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < NB; ++c) {
      float kv = khat_entry(ur, us[c], e0);
      if (tb == ta && c == tid) kv = 1.0f;
      s += (double)(kv * as[c]);
    }
  }
  rout[vecoff + row] = rvalid ? (float)((double)a.rbuf[vecoff + row] - s) : 0.0f;
}

// Per (matrix, tile t), thread per row: total S1 = own block row's sums + the column sums of the tiles (a, t), a > t;
// per-point gradients dmean / dfeat and the tile's partial sums for the hyper-parameter gradients.
__global__ void big_gradfin_kernel(BigArgs a) {
  __shared__ float sred[4][12];
  const int m = blockIdx.x / a.nb_max, tt = blockIdx.x % a.nb_max;
  const BigMat mt = a.mat[m];
  if (tt >= mt.nb) return;
  const int tid = threadIdx.x, row = tt * NB + tid, warp = tid >> 5, lane = tid & 31;
  const size_t vecoff = (size_t)m * a.npad;
  const bool rvalid = row < mt.n;
  const float* rp = a.rowpart + (vecoff + row) * 8;
  float4 S = *reinterpret_cast<const float4*>(rp);
  const float trc = rp[4];
  for (int ta = tt + 1; ta < mt.nb; ++ta) {
    const float4 c = *reinterpret_cast<const float4*>(a.colpart + (((size_t)m * a.npairs + (size_t)ta * (ta - 1) / 2 + tt) * NB + tid) * 4);
    S.x += c.x; S.y += c.y; S.z += c.z; S.w += c.w;
  }
  const float S1[4] = {S.x, S.y, S.z, S.w};
  *reinterpret_cast<float4*>(a.rowpart + (vecoff + row) * 8) = S;         // total, for big_dfeat_kernel
  const float alpha_r = a.abuf[vecoff + row], beta_r = alpha_r / mt.tot;
  const float4 urow = a.ubuf[vecoff + row], u0 = a.ubuf[vecoff];
  float red[12];
  red[0] = rvalid ? a.rbuf[vecoff + row] * alpha_r : 0.0f;      // quad * tot
  red[1] = 0.0f;
  red[2] = rvalid ? trc : 0.0f;
  red[3] = rvalid ? beta_r : 0.0f;
  red[4] = rvalid ? 2.0f * (urow.x - u0.x) * S1[0] : 0.0f;
  red[5] = rvalid ? 2.0f * (urow.y - u0.y) * S1[1] : 0.0f;
  red[6] = rvalid ? 2.0f * (urow.z - u0.z) * S1[2] : 0.0f;
  red[7] = rvalid ? 2.0f * (urow.w - u0.w) * S1[3] : 0.0f;
#pragma unroll
  for (int f = 0; f < 4; ++f) red[8 + f] = rvalid ? S1[f] : 0.0f;
#pragma unroll
  for (int i = 0; i < 12; ++i) red[i] = warp_sum(red[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) sred[warp][i] = red[i];
  }
  __syncthreads();
  if (tid < 12) a.part[((size_t)m * a.nb_max + tt) * 12 + tid] = (sred[0][tid] + sred[1][tid]) + (sred[2][tid] + sred[3][tid]);
}

// Per (matrix, tile), thread per row: per-point gradients dmean / dfeat.  sum_a S1_a = 0 exactly (the kernel is invariant
// to a common shift of the features); the rounding residue of the matrix-wide sum is removed here, which projects the
// feature gradient back onto that invariance (the kernel net's output-bias gradient is this sum over all points).
__global__ void big_dfeat_kernel(BigArgs a) {
  const int m = blockIdx.x / a.nb_max, tt = blockIdx.x % a.nb_max;
  const BigMat mt = a.mat[m];
  if (tt >= mt.nb) return;
  const int tid = threadIdx.x, row = tt * NB + tid;
  const GpArgs& ga = a.g;
  if (row >= ga.n) return;
  const size_t vecoff = (size_t)m * a.npad;
  const bool ok = row < mt.n && mt.status >= 0;
  const float inv_n = 1.0f / (float)mt.n;
  const size_t qpt = (size_t)mt.p * ga.T * ga.n + (size_t)mt.t * ga.n + row;
  if (ga.dmean != nullptr) ga.dmean[qpt] = ok ? a.abuf[vecoff + row] / mt.tot * inv_n : 0.0f;
  if (ga.dfeat != nullptr) {
    const float* th = ga.theta + (size_t)mt.p * ga.D;
    const float* S1 = a.rowpart + (vecoff + row) * 8;
    for (int f = 0; f < ga.F; ++f) {
      float tot = 0.0f;
      for (int tb = 0; tb < mt.nb; ++tb) tot += a.part[((size_t)m * a.nb_max + tb) * 12 + 8 + f];
      const float sc = kCB / softplus_f(__ldg(th + ga.off_ls + f));     // scaled inverse lengthscale
      ga.dfeat[qpt * ga.F + f] = ok ? -mt.rho * inv_n * sc / (kCB * kCB) * (S1[f] - tot * inv_n) : 0.0f;
    }
  }
}

// Values-only calls (predictive path): quad * tot = v.v with v = L^-1 r from the diagonal-tile kernels; no U, no gradient.
__global__ void big_quad_kernel(BigArgs a) {
  __shared__ float sred[8];
  const int m = blockIdx.x;
  const BigMat mt = a.mat[m];
  float s = 0.0f;
  for (int i = threadIdx.x; i < mt.nb * NB; i += blockDim.x) { const float v = a.vbuf[(size_t)m * a.npad + i]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  for (int i = threadIdx.x; i < mt.nb * 12; i += blockDim.x) {
    float v = 0.0f;
    if (i == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sred[w];
    a.part[(size_t)m * a.nb_max * 12 + i] = v;
  }
}

// Per matrix: mll, hyper-parameter gradients, status (same formulas as gp_tc_kernel's row-0 epilogue).
__global__ void big_finish_kernel(BigArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.B) return;
  const GpArgs& g = a.g;
  const BigMat mt = a.mat[m];
  const int F = g.F;
  float* hyp = g.dhyp + ((size_t)mt.p * g.T + mt.t) * gp_hyp_stride(F);
  float* mll_out = g.mll + (size_t)mt.p * g.T + mt.t;
  if (g.info != nullptr) g.info[(size_t)mt.p * g.T + mt.t] = mt.status;
  if (mt.status < 0) {
    *mll_out = CUDART_NAN_F;
    for (int f = 0; f < F + 3; ++f) hyp[f] = 0.0f;
    return;
  }
  float red[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ld2 = 0.0f;
  for (int tb = 0; tb < mt.nb; ++tb) {
    for (int i = 0; i < 8; ++i) red[i] += a.part[((size_t)m * a.nb_max + tb) * 12 + i];
    ld2 += a.ldet[(size_t)m * a.nb_max + tb];
  }
  const float n = (float)mt.n, inv_n = 1.0f / n, inv_tot = 1.0f / mt.tot;
  const float quad = red[0] * inv_tot;
  const float logdet = n * logf(mt.tot) + 2.0f * ld2 * 0.69314718055994530942f;
  *mll_out = (-0.5f * quad - 0.5f * logdet - 0.5f * n * 1.83787706640934548356f) * inv_n;
  const float* th = g.theta + (size_t)mt.p * g.D;
  for (int f = 0; f < F && f < 4; ++f) {
    const float raw = __ldg(th + g.off_ls + f), ls = softplus_f(raw);
    hyp[f] = mt.rho * inv_n * red[4 + f] * (0.5f * sigmoid_f(raw) / (kCB * kCB * ls));
  }
  hyp[F] = 0.5f * inv_tot * inv_n * red[2] * sigmoid_f(__ldg(th + g.off_noise));
  // output scale: sum_ab (beta_a alphahat_b - Khat^-1_ab) k_ab with k = (Khat - (1 - rho) I) / rho collapses to
  // (quad - n - (1 - rho) Str) / rho  (tr(Khat^-1 Khat) = n, beta . Khat alphahat = quad): no element-wise cancellation
  const float Skt = (quad - n - (1.0f - mt.rho) * red[2]) / mt.rho;
  hyp[F + 1] = g.has_oscale ? 0.5f * inv_tot * inv_n * Skt * sigmoid_f(__ldg(th + g.off_oscale)) : 0.0f;
  hyp[F + 2] = red[3] * inv_n;
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else cudaGetLastError();
  }
  return fn;
}

// 3-D fp32 tensor (inner, rows, batch), boxes of 32 x 128 x 1 elements, 128-byte swizzle.
bool make_map(CUtensorMap* map, void* base, uint64_t inner, uint64_t rows, uint64_t batch) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[3] = {inner, rows, batch};
  const cuuint64_t strides[2] = {inner * 4, inner * rows * 4};
  const cuuint32_t box[3] = {KC, NB, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct BigLayout {
  int nb, npad, batch;                 // tiles per side, padded size, matrices per pass
  size_t off_L, off_SB, off_u, off_r, off_v, off_a, off_w, off_rowp, off_colp, off_ldet, off_part, off_mat, off_list, total;   // bytes
};

size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

BigLayout big_layout(int n, long long matrices) {
  BigLayout L;
  L.nb = (n + NB - 1) / NB;
  L.npad = L.nb * NB;
  const size_t per = (size_t)L.npad * L.npad * 4 + (size_t)L.nb * 2 * NB * NB * 4 + (size_t)L.npad * 72 + (size_t)(L.nb * (L.nb - 1) / 2) * NB * 16;
  const size_t budget = (size_t)24 << 30;
  long long batch = std::max<long long>(1, std::min<long long>(matrices, (long long)(budget / per)));
  L.batch = (int)batch;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = al256(off + bytes); return o; };
  L.off_L = take((size_t)L.batch * L.npad * L.npad * 4);
  L.off_SB = take((size_t)L.batch * L.nb * 2 * NB * NB * 4);
  L.off_u = take((size_t)L.batch * L.npad * 16);
  L.off_r = take((size_t)L.batch * L.npad * 4);
  L.off_v = take((size_t)L.batch * L.npad * 4);
  L.off_a = take((size_t)L.batch * L.npad * 4);
  L.off_w = take((size_t)L.batch * L.npad * 4);
  L.off_rowp = take((size_t)L.batch * L.npad * 32);
  L.off_colp = take((size_t)L.batch * (L.nb * (L.nb - 1) / 2) * NB * 16);
  L.off_ldet = take((size_t)L.batch * L.nb * 4);
  L.off_part = take((size_t)L.batch * L.nb * 12 * 4);
  L.off_mat = take((size_t)L.batch * sizeof(BigMat));
  L.off_list = take((size_t)(L.batch + 1) * 3 * 4);      // three retry lists + their counters
  L.total = off;
  return L;
}

int grid_for(long long items) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  else cudaGetLastError();
  return (int)std::max<long long>(1, std::min<long long>(items, sms));
}

template <int MODE>
int launch_big(const CUtensorMap& mL, const CUtensorMap& mS, const BigArgs& a, int step, int per, cudaStream_t st) {
  static bool once = false;
  if (!once) {
    PACOH_CUDA_CHECK(cudaFuncSetAttribute(big_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmem));
    once = true;
  }
  if (per <= 0) return PACOH_OK;
  const long long items = (long long)a.B * per;
  // retry passes (list != nullptr) normally have nothing to do: a small grid that exits at once
  const int grid = a.list != nullptr ? std::min(grid_for(items), 64) : grid_for(items);
  big_kernel<MODE><<<grid, kBigThreads, kBigSmem, st>>>(mL, mS, a, step);
  PACOH_CUDA_CHECK(cudaGetLastError());
  return PACOH_OK;
}

}  // namespace

size_t gp_big_workspace_bytes(int n, long long matrices) { return big_layout(n, matrices).total; }

// diagnostics (pacoh_debug_big_layout): byte offsets of the scratch buffers inside the large-n workspace
void gp_big_layout_debug(int n, long long matrices, long long* out) {
  const BigLayout L = big_layout(n, matrices);
  out[0] = L.nb; out[1] = L.npad; out[2] = L.batch; out[3] = (long long)L.off_L; out[4] = (long long)L.off_SB;
  out[5] = (long long)L.off_u; out[6] = (long long)L.off_r; out[7] = (long long)L.off_v; out[8] = (long long)L.off_a;
  out[9] = (long long)L.off_ldet; out[10] = (long long)L.off_part; out[11] = (long long)L.off_mat; out[12] = (long long)L.total;
}

// Large-n path: 64 < n (every n works; the small kernels are faster below).  `ws` = gp_big_workspace_bytes(n, P * T) bytes.
int launch_gp_mll_big(const GpArgs& g, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g.F < 1 || g.F > 4) return PACOH_ERR_UNSUPPORTED;
  const long long matrices = (long long)g.P * g.T;
  const BigLayout L = big_layout(g.n, matrices);
  if (ws == nullptr || ws_bytes < L.total) { set_error("large-n GP path: workspace too small"); return PACOH_ERR_WORKSPACE; }
  uint8_t* base = (uint8_t*)ws;
  BigArgs a;
  memset(&a, 0, sizeof(a));
  a.nb_max = L.nb; a.npad = L.npad;
  a.Lbuf = (float*)(base + L.off_L); a.SB = (float*)(base + L.off_SB);
  a.ubuf = (float4*)(base + L.off_u); a.rbuf = (float*)(base + L.off_r);
  a.vbuf = (float*)(base + L.off_v); a.abuf = (float*)(base + L.off_a); a.wbuf = (float*)(base + L.off_w);
  a.ldet = (float*)(base + L.off_ldet); a.part = (float*)(base + L.off_part);
  a.rowpart = (float*)(base + L.off_rowp); a.colpart = (float*)(base + L.off_colp); a.npairs = L.nb * (L.nb - 1) / 2;
  a.mat = (BigMat*)(base + L.off_mat);
  int* lists = (int*)(base + L.off_list);
  a.g = g;
  CUtensorMap mL, mS;
  int rc;
  // PACOH_BIG_TIMING=1: per-phase device times of every pass on stderr (diagnostics; synchronises the stream)
  static int timing = -1;
  if (timing < 0) { const char* e = getenv("PACOH_BIG_TIMING"); timing = (e != nullptr && e[0] == '1') ? 1 : 0; }
  cudaEvent_t ev[6];
  if (timing) for (auto& e : ev) cudaEventCreate(&e);
  auto mark = [&](int i) { if (timing) cudaEventRecord(ev[i], st); };
  for (long long b0 = 0; b0 < matrices; b0 += L.batch) {
    a.B = (int)std::min<long long>(L.batch, matrices - b0);
    a.list = nullptr; a.count = nullptr;
    if (!make_map(&mL, a.Lbuf, (uint64_t)L.npad, (uint64_t)L.npad, (uint64_t)a.B) ||
        !make_map(&mS, a.SB, NB, NB, (uint64_t)a.B * L.nb * 2)) {
      set_error("large-n GP path: cuTensorMapEncodeTiled failed");
      return PACOH_ERR_CUDA;
    }
    mark(0);
    PACOH_CUDA_CHECK(cudaMemsetAsync(lists, 0, (size_t)(L.batch + 1) * 3 * 4, st));
    big_prep_kernel<<<a.B, 128, 0, st>>>(a, (int)b0);
    PACOH_CUDA_CHECK(cudaGetLastError());
    // ---- factorisation, up to 4 attempts (attempts 1-3 only touch the matrices of the retry list: normally none)
    for (int attempt = 0; attempt < 4; ++attempt) {
      if (attempt > 0) {
        int* list = lists + (size_t)(attempt - 1) * (L.batch + 1);
        BigArgs all = a; all.list = nullptr; all.count = nullptr;
        big_retry_kernel<<<(a.B + 127) / 128, 128, 0, st>>>(all, list + 1, list);
        PACOH_CUDA_CHECK(cudaGetLastError());
        a.list = list + 1; a.count = list;
      }
      for (int k = 0; k < L.nb; ++k) {
        if ((rc = launch_big<M_DIAG>(mL, mS, a, k, 1, st)) != PACOH_OK) return rc;
        {
          static bool once = false;
          const int psmem = (NB * LDT + NB + 16) * 4;
          if (!once) { PACOH_CUDA_CHECK(cudaFuncSetAttribute(big_potrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem)); once = true; }
          const int pgrid = a.list != nullptr ? std::min(a.B, 64) : std::min(a.B, 3 * grid_for(1 << 30));
          big_potrf_kernel<<<pgrid, 256, psmem, st>>>(a, k);
          PACOH_CUDA_CHECK(cudaGetLastError());
        }
        if ((rc = launch_big<M_PANEL>(mL, mS, a, k, L.nb - k - 1, st)) != PACOH_OK) return rc;
      }
    }
    {   // matrices that failed the last attempt
      BigArgs all = a; all.list = nullptr; all.count = nullptr;
      int* list = lists + (size_t)2 * (L.batch + 1);
      big_retry_kernel<<<(a.B + 127) / 128, 128, 0, st>>>(all, list + 1, list);   // lvl == 3 and failed -> status -1
      PACOH_CUDA_CHECK(cudaGetLastError());
    }
    a.list = nullptr; a.count = nullptr;
    mark(1);
    if (g.values_only) {
      big_quad_kernel<<<a.B, 256, 0, st>>>(a);
      big_finish_kernel<<<(a.B + 127) / 128, 128, 0, st>>>(a);
      PACOH_CUDA_CHECK(cudaGetLastError());
      continue;
    }
    for (int s = 1; s < L.nb; ++s)
      if ((rc = launch_big<M_UINV>(mL, mS, a, s, L.nb - s, st)) != PACOH_OK) return rc;
    mark(2);
    big_ux_kernel<<<a.B * L.nb, 256, 0, st>>>(a, a.vbuf, a.abuf, 0);                 // alphahat = U v
    big_resid_kernel<<<a.B * L.nb, NB, 0, st>>>(a, a.vbuf);                          // r' = r - Khat alphahat   (v is dead)
    big_utx_kernel<<<a.B * L.nb, NB, 0, st>>>(a, a.vbuf, a.wbuf);                     // w = L^-1 r'
    big_ux_kernel<<<a.B * L.nb, 256, 0, st>>>(a, a.wbuf, a.abuf, 1);                 // alphahat += U w
    PACOH_CUDA_CHECK(cudaGetLastError());
    mark(3);
    if ((rc = launch_big<M_GRAD>(mL, mS, a, 0, L.nb, st)) != PACOH_OK) return rc;
    mark(4);
    big_gradfin_kernel<<<a.B * L.nb, NB, 0, st>>>(a);
    big_dfeat_kernel<<<a.B * L.nb, NB, 0, st>>>(a);
    big_finish_kernel<<<(a.B + 127) / 128, 128, 0, st>>>(a);
    PACOH_CUDA_CHECK(cudaGetLastError());
    mark(5);
    if (timing) {
      cudaEventSynchronize(ev[5]);
      float ms[5];
      for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
      fprintf(stderr, "[pacoh big] n=%d matrices=%d: cholesky %.3f ms, U=L^-T %.3f, alpha+refine %.3f, K^-1+grad %.3f, finish %.3f\n",
              g.n, a.B, ms[0], ms[1], ms[2], ms[3], ms[4]);
    }
  }
  return PACOH_OK;
}

}  // namespace pacoh
