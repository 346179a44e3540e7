// Validation of tcgen05.mma with the A operand in TENSOR MEMORY (written by tcgen05.st, lane = row) and B in shared
// memory (K-major, SWIZZLE_NONE): D[128 x 32] = A[128 x 32] * B[32 x 32]^T in 3xTF32, compared with an fp64 CPU reference.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46; return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ int kmajor_off(int r, int k, int R) { return ((r + (k >> 2) * R) << 2) + (k & 3); }
#define ST32(taddr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
  :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), \
     "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")
__global__ void __launch_bounds__(128) tc_ts(const float* __restrict__ Ain, const float* __restrict__ Bin, float* __restrict__ out) {
  __shared__ __align__(1024) float sB_hi[32 * 32], sB_lo[32 * 32];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  for (int i = tid; i < 32 * 32; i += 128) { int r = i >> 5, k = i & 31; float v = Bin[i], h = tf32_hi(v); sB_hi[kmajor_off(r, k, 32)] = h; sB_lo[kmajor_off(r, k, 32)] = v - h; }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  // A row of this thread -> TMEM columns [32,64) hi and [64,96) lo of its lane
  uint32_t hi[32], lo[32];
  for (int k = 0; k < 32; ++k) { float v = Ain[tid * 32 + k], h = tf32_hi(v); hi[k] = __float_as_uint(h); lo[k] = __float_as_uint(v - h); }
  ST32(lane_addr + 32, hi);
  ST32(lane_addr + 64, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = make_idesc(128, 32);
    uint32_t acc = 0;
    for (int ps = 0; ps < 3; ++ps) {
      const uint32_t a_col = ps == 0 ? 64 : 32;            // lo*hi, hi*lo, hi*hi
      const float* b = ps == 1 ? sB_lo : sB_hi;
      for (int ks = 0; ks < 4; ++ks) {
        mma_ts(tmem, tmem + a_col + ks * 8, make_desc(smem_u32(b) + ks * 2 * 32 * 16, 32 * 16, 128), idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(lane_addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem));
}
int main() {
  srand(2);
  std::vector<float> A(128 * 32), B(32 * 32), out(128 * 32);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dO;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  tc_ts<<<1, 128>>>(dA, dB, dO);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) { double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)A[r * 32 + k] * B[n * 32 + k]; maxerr = fmax(maxerr, fabs(ref - out[r * 32 + n])); maxref = fmax(maxref, fabs(ref)); }
  printf("A-in-TMEM (TS) 128x32x32 3xTF32: max abs err %.3e (max |ref| %.3f)\n", maxerr, maxref);
  return 0;
}
