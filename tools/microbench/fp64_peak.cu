// fp64_peak.cu -- sustained FP64 FMA throughput of this GPU (MEASURED_PEAKS.json has no FP64
// figure; the nominal one is 148 SMs x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s).  Independent DFMA chains
// in registers, no memory traffic; reports TFLOP/s for a few occupancies.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;   // keeps the chains alive
}

template <int CHAINS>
void run(double *out, int blocks_per_sm, int sms) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * blocks_per_sm;
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k_dfma<CHAINS><<<grid, 256>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  const double flops = 2.0 * CHAINS * (double)iters * 256.0 * grid;
  printf("{\"chains_per_thread\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"fp64_TFLOPs\": %.2f}\n", CHAINS,
         blocks_per_sm * 8, best, flops / best / 1e9);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *out;
  cudaMalloc(&out, 8);
  run<8>(out, 1, p.multiProcessorCount);
  run<8>(out, 2, p.multiProcessorCount);
  run<8>(out, 4, p.multiProcessorCount);
  run<16>(out, 2, p.multiProcessorCount);
  run<2>(out, 2, p.multiProcessorCount);
  printf("{\"sms\": %d, \"clock_MHz\": %d}\n", p.multiProcessorCount, p.clockRate / 1000);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
