// Stage-1 validation of hand-written tcgen05 (UMMA) TF32 GEMMs from shared memory with SWIZZLE_NONE canonical layouts.
//   test 1: D[128 x 32] = A[128 x 32] * B[32 x 32]^T   A, B K-major           (MLP forward:  H * W^T)
//   test 2: G[64 x 40]  = P[128 x 64]^T * Q[128 x 40]   both operands MN-major  (weight gradients: dA^T * [H | x 1])
// Each with plain TF32 (1 pass) and 3xTF32 (hi/lo split, 3 passes); results are compared with an fp64 CPU reference.
// Prints where the M=64 accumulator rows land in TMEM.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): addresses / offsets in 16-byte units.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                        // c_format = F32
  d |= 2u << 7;                        // a_format = TF32
  d |= 2u << 10;                       // b_format = TF32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
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

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// K-major tile, R rows x 32 K-elements (8 chunks of 16 B): offset16(r, kc) = r + kc * R      [SBO = 128 B, LBO = R*16 B]
__device__ __forceinline__ int kmajor_off(int r, int k, int R) { return ((r + (k >> 2) * R) << 2) + (k & 3); }

template <int MODE>   // 0: K-major 128x32x32, 1: MN-major 64x40x128
__global__ void __launch_bounds__(128) tc_test(const float* __restrict__ Ain, const float* __restrict__ Bin, float* __restrict__ out,
                                               int three_pass_in) {
  const int three_pass = three_pass_in & 1, variant = three_pass_in >> 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sA_hi = (float*)smem_raw;                 // up to 128 x 64 floats = 32 KB
  float* sA_lo = sA_hi + 128 * 64;
  float* sB_hi = sA_lo + 128 * 64;                 // up to 128 x 40 floats = 20 KB
  float* sB_lo = sB_hi + 128 * 40;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // ---- stage operands into the canonical layouts (hi / lo split)
  if (MODE == 0) {
    for (int i = tid; i < 128 * 32; i += 128) {    // A[r][k], 128 rows
      int r = i >> 5, k = i & 31;
      float v = Ain[i], h = tf32_hi(v);
      sA_hi[kmajor_off(r, k, 128)] = h; sA_lo[kmajor_off(r, k, 128)] = v - h;
    }
    for (int i = tid; i < 32 * 32; i += 128) {     // B[n][k], 32 rows
      int r = i >> 5, k = i & 31;
      float v = Bin[i], h = tf32_hi(v);
      sB_hi[kmajor_off(r, k, 32)] = h; sB_lo[kmajor_off(r, k, 32)] = v - h;
    }
  } else {
    // MN-major: P[pt][m] (128 points x 64), Q[pt][n] (128 x 40): the SAME physical layout as a K-major tile with rows = points:
    // offset16(pt, chunk) = pt + chunk * 128    -> as MN-major operand: SBO (chunk stride) = 2048 B, LBO (8-point group) = 128 B
    for (int i = tid; i < 128 * 64; i += 128) {
      int r = i >> 6, m = i & 63;
      float v = Ain[i], h = tf32_hi(v);
      sA_hi[kmajor_off(r, m, 128)] = h; sA_lo[kmajor_off(r, m, 128)] = v - h;
    }
    for (int i = tid; i < 128 * 40; i += 128) {
      int r = i / 40, n = i - r * 40;
      float v = Bin[i], h = tf32_hi(v);
      sB_hi[kmajor_off(r, n, 128)] = h; sB_lo[kmajor_off(r, n, 128)] = v - h;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;");   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const int passes = three_pass ? 3 : 1;
    uint32_t acc = 0;
    for (int ps = 0; ps < passes; ++ps) {
      // 3xTF32: lo*hi + hi*lo first, hi*hi last
      const float* a = (passes == 1 || ps == 2) ? sA_hi : (ps == 0 ? sA_lo : sA_hi);
      const float* b = (passes == 1 || ps == 2) ? sB_hi : (ps == 0 ? sB_hi : sB_lo);
      if (MODE == 0) {
        const uint32_t idesc = make_idesc(128, 32, 0, 0);
        for (int ks = 0; ks < 4; ++ks) {           // K = 32 = 4 x 8
          uint64_t ad = make_desc(smem_u32(a) + ks * 2 * 128 * 16, 128 * 16, 128);
          uint64_t bd = make_desc(smem_u32(b) + ks * 2 * 32 * 16, 32 * 16, 128);
          mma_tf32(tmem, ad, bd, idesc, acc);
          acc = 1;
        }
      } else {
        const int swap = (variant >> 0) & 1, m128 = (variant >> 1) & 1;
        const uint32_t idesc = make_idesc(m128 ? 128 : 64, m128 ? 48 : 40, 1, 1);
        for (int ks = 0; ks < 16; ++ks) {          // K = 128 points = 16 x 8
          uint64_t ad = swap ? make_desc(smem_u32(a) + ks * 128, 2048, 128) : make_desc(smem_u32(a) + ks * 128, 128, 2048);
          uint64_t bd = swap ? make_desc(smem_u32(b) + ks * 128, 2048, 128) : make_desc(smem_u32(b) + ks * 128, 128, 2048);
          mma_tf32(tmem, ad, bd, idesc, acc);
          acc = 1;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)));
  }
  mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  // ---- read back all 128 lanes x 64 columns
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + half * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tmem));
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main() {
  srand(1);
  float *dA, *dB, *dO;
  CK(cudaMalloc(&dA, 128 * 64 * 4)); CK(cudaMalloc(&dB, 128 * 40 * 4)); CK(cudaMalloc(&dO, 128 * 64 * 4));
  const size_t smem = (2 * 128 * 64 + 2 * 128 * 40) * 4;
  CK(cudaFuncSetAttribute(tc_test<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(tc_test<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> out(128 * 64);
  // ---- test 1
  {
    std::vector<float> A(128 * 32), B(32 * 32);
    for (auto& v : A) v = (float)frand();
    for (auto& v : B) v = (float)frand();
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int three = 0; three < 2; ++three) {
      CK(cudaMemset(dO, 0, 128 * 64 * 4));
      tc_test<0><<<1, 128, smem>>>(dA, dB, dO, three);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) {
        double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)A[r * 32 + k] * B[n * 32 + k];
        maxerr = fmax(maxerr, fabs(ref - out[r * 64 + n])); maxref = fmax(maxref, fabs(ref));
      }
      printf("test1 K-major 128x32x32 %s: max abs err %.3e (max |ref| %.3f)\n", three ? "3xTF32" : "1xTF32", maxerr, maxref);
    }
  }
  // ---- test 2
  {
    std::vector<float> P(128 * 64), Q(128 * 40);
    for (auto& v : P) v = (float)frand();
    for (auto& v : Q) v = (float)frand();
    CK(cudaMemcpy(dA, P.data(), P.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
    std::vector<double> ref(64 * 40);
    for (int m = 0; m < 64; ++m) for (int n = 0; n < 40; ++n) { double s = 0; for (int r = 0; r < 128; ++r) s += (double)P[r * 64 + m] * Q[r * 40 + n]; ref[m * 40 + n] = s; }
    for (int var = 0; var < 4; ++var)
    for (int three = 0; three < 1; ++three) {
      printf("variant swap=%d m128=%d\n", var & 1, var >> 1);
      CK(cudaMemset(dO, 0, 128 * 64 * 4));
      tc_test<1><<<1, 128, smem>>>(dA, dB, dO, three | (var << 1));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
      // find which TMEM lane holds row m: match column 0..39 pattern
      int lane_of_row[64]; double maxerr = 0;
      for (int m = 0; m < 64; ++m) {
        int best = -1; double berr = 1e30;
        for (int l = 0; l < 128; ++l) { double e = 0; for (int n = 0; n < 40; ++n) e = fmax(e, fabs(ref[m * 40 + n] - out[l * 64 + n])); if (e < berr) { berr = e; best = l; } }
        lane_of_row[m] = best; maxerr = fmax(maxerr, berr);
      }
      printf("test2 MN-major 64x40x128 %s: max abs err %.3e ; row->lane:", three ? "3xTF32" : "1xTF32", maxerr);
      for (int m = 0; m < 64; m += 1) if (m < 4 || m % 16 == 0 || m == 63) printf(" %d->%d", m, lane_of_row[m]);
      printf("\n");
      if (!three) {   // where does G[m][n] land?  search every (lane, column)
        const int probes[][2] = {{0, 0}, {0, 1}, {0, 8}, {1, 0}, {8, 0}, {15, 0}, {16, 0}, {31, 0}, {32, 0}, {48, 0}, {63, 39}};
        for (auto& pr : probes) {
          printf("   G[%d][%d]=%.4f found at:", pr[0], pr[1], ref[pr[0] * 40 + pr[1]]);
          for (int l = 0; l < 128; ++l) for (int c = 0; c < 64; ++c)
            if (fabs(out[l * 64 + c] - ref[pr[0] * 40 + pr[1]]) < 3e-2) printf(" (lane %d, col %d)", l, c);
          printf("\n");
        }
        int nz = 0; for (int l = 0; l < 128; ++l) { bool any = false; for (int c = 0; c < 64; ++c) any |= out[l * 64 + c] != 0.f; nz += any; }
        printf("   lanes with any nonzero: %d; lane0: %.3f %.3f %.3f %.3f | lane16: %.3f %.3f | lane 32: %.3f %.3f\n", nz, out[0], out[1], out[2], out[3],
               out[16 * 64], out[16 * 64 + 1], out[32 * 64], out[32 * 64 + 1]);
      }
    }
  }
  return 0;
}
