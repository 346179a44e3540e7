// Validation of the weight-gradient MMA of mlp_tc_bwd.cu:  dW[32 x 32] = dA^T[32 x 128 pts] * H[128 pts x 32]
// with K = points.  Both operands live in shared memory as K-major SWIZZLE_128B tiles written column-by-column by the
// thread that owns the point; hi / lo parts are STACKED along M / N (rows 0-31 hi, 32-63 lo) so that ONE M128 N64 MMA
// per K-step yields hi*hi, hi*lo and lo*hi as separate 32 x 32 blocks of D (summed once at the end).
// mode 0: hi = truncated tf32; mode 1: hi = round-to-nearest tf32; mode 2: "hi" rows hold the unsplit fp32 value
// (tests whether the tensor core truncates the low 13 mantissa bits).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {   // K-major SWIZZLE_128B, 8-row groups 1024 B apart
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
// byte offset of (row r, point k = 32 w + l) inside a stacked [64 rows][128 points] tile: K-block w = 8 KB
__device__ __forceinline__ int sw_off(int r, int w, int l) { return w * 8192 + (r >> 3) * 1024 + (r & 7) * 128 + ((((l >> 2) ^ (r & 7))) << 4) + (l & 3) * 4; }

__global__ void __launch_bounds__(128) tc_sw(const float* __restrict__ dA, const float* __restrict__ H, float* __restrict__ out, int mode, int M, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;            // 32 KB
  unsigned char* sB = smem + 32768;    // 32 KB
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  for (int f = 0; f < 32; ++f) {
    const float a = dA[tid * 32 + f], h = H[tid * 32 + f];
    float ah, al, hh, hl;
    if (mode == 0) { ah = tf32_trunc(a); al = a - ah; hh = tf32_trunc(h); hl = h - hh; }
    else if (mode == 1) { ah = tf32_rn(a); al = a - ah; hh = tf32_rn(h); hl = h - hh; }
    else { ah = a; al = a - tf32_trunc(a); hh = h; hl = h - tf32_trunc(h); }
    *reinterpret_cast<float*>(sA + sw_off(f, warp, lane)) = ah;
    *reinterpret_cast<float*>(sA + sw_off(32 + f, warp, lane)) = al;
    *reinterpret_cast<float*>(sB + sw_off(f, warp, lane)) = hh;
    *reinterpret_cast<float*>(sB + sw_off(32 + f, warp, lane)) = hl;
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, 64);
    uint32_t acc = 0;
    for (int w = 0; w < 4; ++w)
      for (int ks = 0; ks < 4; ++ks) {        // 8 points = 32 B per K-step inside the 128-byte swizzle row
        mma_ss(tmem, desc_sw128(smem_u32(sA) + w * 8192 + ks * 32), desc_sw128(smem_u32(sB) + w * 8192 + ks * 32), idesc, acc);
        acc = 1;
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
  }
  mbar_wait(smem_u32(&mbar), 0);
  if (cyc != nullptr) {   // timing: 64 rounds of the same 16 MMAs into a scratch accumulator (columns 64..127 are not allocated: reuse D, results garbage)
    __syncthreads();
    long long t0 = clock64();
    if (tid == 0) {
      const uint32_t idesc = make_idesc(M, 64);
      for (int rep = 0; rep < 64; ++rep)
        for (int w = 0; w < 4; ++w)
          for (int ks = 0; ks < 4; ++ks)
            mma_ss(tmem, desc_sw128(smem_u32(sA) + w * 8192 + ks * 32), desc_sw128(smem_u32(sB) + w * 8192 + ks * 32), idesc, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
    }
    mbar_wait(smem_u32(&mbar), 1);
    long long t1 = clock64();
    if (tid == 0) *cyc = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(lane_addr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * 64 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}
int main() {
  srand(3);
  std::vector<float> A(128 * 32), H(128 * 32), out(128 * 64);
  for (auto& v : A) v = ((float)rand() / RAND_MAX * 2 - 1) * 3.0f;
  for (auto& v : H) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dH, *dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dH, H.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 65536 + 16384;
  CK(cudaFuncSetAttribute(tc_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* dC; CK(cudaMalloc(&dC, 8));
  // 1. reference run M = 128 to know the right answers, rows 0..63 x 64 columns
  tc_sw<<<1, 128, smem>>>(dA, dH, dO, 0, 128, nullptr);
  CK(cudaDeviceSynchronize());
  std::vector<float> ref(128 * 64);
  CK(cudaMemcpy(ref.data(), dO, ref.size() * 4, cudaMemcpyDeviceToHost));
  // 2. M = 64: where do rows 0..63 land?
  CK(cudaMemset(dO, 0, out.size() * 4));
  tc_sw<<<1, 128, smem>>>(dA, dH, dO, 0, 64, nullptr);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  for (int r = 0; r < 64; ++r) {
    int found = -1;
    for (int l = 0; l < 128; ++l) {
      bool same = true;
      for (int c = 0; c < 64 && same; ++c) same = fabsf(out[l * 64 + c] - ref[r * 64 + c]) <= 1e-5f * (1.0f + fabsf(ref[r * 64 + c]));
      if (same) { found = l; break; }
    }
    printf("%d->%d ", r, found);
  }
  printf("\n");
  for (int M : {128, 64}) {
    long long c = 0;
    tc_sw<<<1, 128, smem>>>(dA, dH, dO, 0, M, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
    printf("M=%d N=64 K=8: %.1f cycles per MMA (1024 back-to-back, one CTA)\n", M, (double)c / 1024.0);
  }
  return 0;
}
