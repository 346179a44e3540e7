// tcgen05 / TMEM helpers shared by the tensor-core MLP kernels (sm_100a inline PTX; no CUTLASS dependency).
// Descriptor bit layouts follow cute::UMMA::SmemDescriptor / InstrDescriptor; validated by tools/ubench/tc_gemm_test.cu.
#pragma once
#include <stdint.h>

namespace pacoh {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_NONE K-major shared-memory matrix descriptor: start address, leading (K-chunk) and stride (8-row group) byte offsets.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;
}

// kind::tf32 instruction descriptor: fp32 accumulate, both operands K-major, M x N tile.
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

// Same with a suspend-time hint (ns): the warp sleeps in hardware instead of spinning through the issue slots.
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
  }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot_in_smem)), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS));
}

// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Store 16 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
         "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// float offset of element (row r, k) in a K-major SWIZZLE_NONE tile with R rows and 32 K-elements:
// 16-byte chunk index = r + (k / 4) * R   (8-row groups contiguous: SBO = 128 B; K-chunk planes: LBO = 16 R bytes)
__device__ __forceinline__ int ktile_off(int r, int k, int R) { return ((r + (k >> 2) * R) << 2) + (k & 3); }

// D[128 x 32] (TMEM) = A[128 x 32] * B[32 x 32]^T in 3xTF32: lo*hi + hi*lo + hi*hi, 12 tcgen05.mma, then commit to `bar`.
// a_hi/a_lo: 128-row K-major tiles; b_hi/b_lo: 32-row K-major tiles (shared-memory byte addresses).
__device__ __forceinline__ void gemm128x32x32_3xtf32(uint32_t tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t bar) {
  const uint32_t idesc = umma_idesc_tf32(128, 32);
  uint32_t acc = 0;
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {
    const uint32_t sa = ps == 0 ? a_lo : a_hi, sb = ps == 1 ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {   // K = 32 = 4 x 8
      umma_tf32(tmem, umma_desc(sa + ks * 2 * 128 * 16, 128 * 16, 128), umma_desc(sb + ks * 2 * 32 * 16, 32 * 16, 128), idesc, acc);
      acc = 1;
    }
  }
  umma_commit(bar);
}

// ---- A operand in TENSOR MEMORY ("TS" form): row r of the 128 x 32 A tile lives in TMEM lane r, columns [col, col+32).
// TMEM column map used by the MLP kernels (128 columns allocated per CTA):
constexpr uint32_t kTmemD = 0;      // accumulator D (32 fp32 columns)
constexpr uint32_t kTmemAhi = 32;   // A operand, tf32 hi parts
constexpr uint32_t kTmemAlo = 64;   // A operand, lo parts
constexpr int kTmemCols = 128;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// This thread's 32 activations -> hi / lo split -> its TMEM lane (columns kTmemAhi.. / kTmemAlo..): no shared memory,
// no proxy fence, no bank conflicts.  `lane_base` = tmem base + (first lane of the warp << 16).
__device__ __forceinline__ void store_a_tmem(uint32_t lane_base, const float (&h)[32]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float v = h[16 * c + e], hv = tf32_hi(v);
      hi[e] = __float_as_uint(hv);
      lo[e] = __float_as_uint(v - hv);
    }
    tmem_st16(lane_base + kTmemAhi + 16 * c, hi);
    tmem_st16(lane_base + kTmemAlo + 16 * c, lo);
  }
  tmem_st_wait();
}

// D[128 x 32] = A[128 x 32] * B[32 x 32]^T in 3xTF32 with A in TMEM, B (hi / lo, 32-row K-major tiles) in shared memory.
__device__ __forceinline__ void gemm128x32x32_3xtf32_ts(uint32_t tmem, uint32_t b_hi, uint32_t b_lo, uint32_t bar) {
  const uint32_t idesc = umma_idesc_tf32(128, 32);
  const uint64_t dhi = umma_desc(b_hi, 32 * 16, 128), dlo = umma_desc(b_lo, 32 * 16, 128);
  uint32_t acc = 0;
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {                    // lo*hi, hi*lo, hi*hi
    const uint32_t a_col = ps == 0 ? kTmemAlo : kTmemAhi;
    const uint64_t db = ps == 1 ? dlo : dhi;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {                  // K = 32 = 4 x 8 (8 TMEM columns / two 32-row K-chunk planes per step)
      umma_tf32_ts(tmem + kTmemD, tmem + a_col + ks * 8, db + (uint64_t)(ks * ((2 * 32 * 16) >> 4)), idesc, acc);
      acc = 1;
    }
  }
  umma_commit(bar);
}

}  // namespace tc
}  // namespace pacoh
