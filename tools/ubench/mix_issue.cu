// Micro-benchmark: does an FP64 instruction (2 pipe cycles per warp on B200: 16 lanes per scheduler) also hold the scheduler's
// issue slot for its second cycle?  K integer-ALU instructions are interleaved with every DFMA (independent chains, no memory):
// if the loop time stays at 2 cycles per DFMA for K = 1 the integer instruction rides in the free slot; if it grows to 2 + K
// the two add up and the hot kernel's bound is (2 x FP64 + other) issue cycles per warp-evaluation.  Development tool.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[4] = {1.0000001, 1e-9, 0.9999999, 2e-9};

template <int K, int KIND>
__global__ void __launch_bounds__(128, 4) k(double *sink, const double *in, const int *iin, int iters)
{
  double a[6], b[6];
  int x[6], y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i] + 1.0; x[i] = iin[threadIdx.x + i]; y[i] = iin[threadIdx.x + 8 + i] | 1; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 6; i++) {
        a[i] = fma(a[i], b[i], kc[1]);                                   // reg, reg, const: 98 % of the FP64 rate on its own
#pragma unroll
        for (int j = 0; j < K; j++) {
          if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[i]) : "r"(y[(i + j) % 6]), "r"(y[(i + j + 1) % 6]));   // LOP3 (majority: does not fold)

          else asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[(i + j) % 6]), "r"(y[(i + j + 1) % 6]));   // IMAD
        }
      }
    }
  }
  double s = 0; int t = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) { s += a[i]; t ^= x[i]; }
  if (s == 123.456 || t == 0x12345678) sink[0] = s + t;
}
template <int K, int KIND>
void run(double *sink, double *in, int *iin, const char *name)
{
  int iters = 10000;
  int blocks = 148 * 4, threads = 128;       // 4 warps per scheduler
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<K, KIND><<<blocks, threads>>>(sink, in, iin, 100);
  cudaEventRecord(e0);
  k<K, KIND><<<blocks, threads>>>(sink, in, iin, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // warp-DFMAs per scheduler: 4 warps x iters x 48; cycles at 1965 MHz
  double cyc = ms * 1e-3 * 1.965e9 / (4.0 * iters * 48.0);
  printf("%-28s K=%d  %.3f ms  %.2f cycles per (DFMA + K int) per scheduler\n", name, K, ms, cyc);
}
int main()
{
  double *sink, *in; int *iin; cudaMalloc(&sink, 8); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&iin, 4096); cudaMemset(iin, 1, 4096);
  run<0, 0>(sink, in, iin, "DFMA only");
  run<1, 0>(sink, in, iin, "DFMA + LOP3");
  run<2, 0>(sink, in, iin, "DFMA + LOP3");
  run<3, 0>(sink, in, iin, "DFMA + LOP3");
  run<1, 2>(sink, in, iin, "DFMA + IMAD");
  run<2, 2>(sink, in, iin, "DFMA + IMAD");
  return 0;
}
