// Measures the FP64 FMA issue peak of the device (the real ceiling of this workload: SURVEY.md F5).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 8);
  const int iters = 20000; constexpr int ILP = 8;
  for (int tpb : {128, 256, 512, 1024}) {
    for (int bps : {1, 2, 4}) {
      if (tpb * bps > 2048) continue;
      int nb = p.multiProcessorCount * bps;
      dfma<ILP><<<nb, tpb>>>(d, 100, 1.0000001, 1e-9);
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      dfma<ILP><<<nb, tpb>>>(d, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = (double)nb * tpb * iters * ILP;
      printf("{\"tpb\": %d, \"blocks_per_sm\": %d, \"dfma_per_s\": %.4e, \"tflops\": %.2f, \"ms\": %.3f, \"sms\": %d}\n", tpb, bps, fmas / (ms * 1e-3), 2 * fmas / (ms * 1e-3) / 1e12, ms, p.multiProcessorCount);
    }
  }
  return 0;
}
