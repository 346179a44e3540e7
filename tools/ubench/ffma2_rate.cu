// Micro-benchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) issue behaviour on sm_100a, alone and mixed with integer ALU work.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, float* sink) {
  float2 acc[16], w[4], v[4];
  int ia = threadIdx.x, ib = 3;
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(0.f, 0.f);
  for (int i = 0; i < 4; ++i) { w[i] = make_float2(sink[i] + 1e-6f * threadIdx.x, sink[i + 4]); v[i] = make_float2(sink[8 + i], sink[12 + i]); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == 0 || MODE == 2) {   // scalar FFMA: 2 instructions per pair
            acc[i * 4 + j].x = fmaf(w[i].x, v[j].x, acc[i * 4 + j].x);
            acc[i * 4 + j].y = fmaf(w[i].y, v[j].y, acc[i * 4 + j].y);
          } else {                        // packed
            acc[i * 4 + j] = __ffma2_rn(w[i], v[j], acc[i * 4 + j]);
          }
          if (MODE >= 2) { ia = ia * 3 + ib; ib ^= ia; }   // two extra ALU/IMAD instructions per FMA pair
        }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 1234.5f || ia == 77) sink[0] = s + ib;
}
template <int MODE> void run(const char* name) {
  float* sink; cudaMalloc(&sink, 256); cudaMemset(sink, 0, 256);
  int iters = 20000, sms = 148, thr = 256, blocks = sms * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, thr>>>(100, sink);
  cudaEventRecord(e0); k<MODE><<<blocks, thr>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = (double)blocks * thr * iters * 2 * 16 * 2;
  printf("%-34s %.1f TFLOP/s  (%.2f ms)\n", name, 2 * fmas / ms / 1e9, ms);
}
int main() {
  run<0>("FFMA scalar"); run<1>("FFMA2 packed"); run<2>("FFMA scalar + 2 int ops / pair"); run<3>("FFMA2 packed + 2 int ops / pair");
  return 0;
}
