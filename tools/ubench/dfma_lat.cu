// Micro-benchmark: FP64 FMA throughput vs (warps per SM, independent chains per thread).  Development tool.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *sink, int iters)
{
  double a[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) a[c] = threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int c = 0; c < CH; c++) a[c] = fma(a[c], m, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += a[c];
  if (s == 123.456) sink[0] = s;
}
template <int CH>
void run(int warps_per_sm, double *sink)
{
  int sms = 148;
  int threads = 32 * (warps_per_sm >= 4 ? 4 : warps_per_sm);       // warps spread over the 4 SMSPs
  int blocks_per_sm = warps_per_sm * 32 / threads;
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CH><<<sms * blocks_per_sm, threads>>>(sink, 100);
  cudaEventRecord(e0);
  k<CH><<<sms * blocks_per_sm, threads>>>(sink, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma_total = (double)sms * blocks_per_sm * threads * iters * 16.0 * CH;
  double tf = 2 * fma_total / (ms * 1e-3) * 1e-12;
  // cycles per dependent DFMA per warp at 1.965 GHz
  double cyc = ms * 1e-3 * 1.965e9 / (iters * 16.0);
  printf("warps/SM %2d chains %d : %.2f TFLOP/s (%.1f%% of 37.2)  cycles per chain step %.2f\n", warps_per_sm, CH, tf, tf / 37.2 * 100, cyc);
}
int main()
{
  double *sink; cudaMalloc(&sink, 8);
  for (int w : {4, 8, 16, 24, 32, 48, 64}) { run<1>(w, sink); run<2>(w, sink); run<4>(w, sink); }
  return 0;
}
