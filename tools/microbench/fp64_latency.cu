// fp64_latency.cu -- dependent-issue latency of the FP64 operations the Kalman recursions chain (one warp, one SM):
// DFMA, IEEE division (1 / x), rsqrt().  The batched solves with few, long systems (Lorenz63: 49,152 chains of
// 5,000 steps) are bound by these chains, not by throughput.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cyc, double x0, double a, int iters) {
  double x = x0 + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (OP == 0) x = fma(x, a, 1e-3);
      else if (OP == 1) x = 1.0 / x + 0.5;        // division + one add
      else if (OP == 2) x = rsqrt(x) + 0.5;       // rsqrt + one add
      else x = x + a;                              // DADD
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double *out;
  long long *cyc, h;
  cudaMalloc(&out, 32 * 8);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char *names[4] = {"DFMA", "1/x + DADD", "rsqrt + DADD", "DADD"};
  for (int op = 0; op < 4; op++) {
    for (int rep = 0; rep < 2; rep++) {
      if (op == 0) chain<0><<<1, 32>>>(out, cyc, 1.0, 0.999, iters);
      if (op == 1) chain<1><<<1, 32>>>(out, cyc, 1.3, 0.999, iters);
      if (op == 2) chain<2><<<1, 32>>>(out, cyc, 1.3, 0.999, iters);
      if (op == 3) chain<3><<<1, 32>>>(out, cyc, 1.3, 1e-9, iters);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_op\": %.2f}\n", names[op], (double)h / (iters * 16.0));
  }
  return 0;
}
