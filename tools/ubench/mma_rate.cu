// Micro-benchmark: issue rate of legacy mma.sync TF32 / BF16 shapes on sm_100a (one-off measurement tool).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, float* sink) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (MODE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 1234.5f) sink[0] = s;
}
template <int MODE> void run(const char* name, double macs_per_mma, int warps_per_sm) {
  float* sink; cudaMalloc(&sink, 64);
  int iters = 20000, sms = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps_per_sm * 32>>>(100, sink);
  cudaEventRecord(e0); k<MODE><<<sms, warps_per_sm * 32>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas = (double)sms * warps_per_sm * iters * 8;
  printf("%-22s warps/SM=%2d: %.1f TFLOP/s, %.2f mma/clk/SM (at 1.92 GHz)\n", name, warps_per_sm, 2 * mmas * macs_per_mma / ms / 1e9,
         mmas / sms / (ms * 1e-3 * 1.92e9));
}
int main() {
  for (int w : {4, 8, 16}) { run<0>("tf32 m16n8k8", 16 * 8 * 8, w); run<1>("tf32 m16n8k4", 16 * 8 * 4, w); run<2>("bf16 m16n8k16", 16 * 8 * 16, w); }
  return 0;
}
