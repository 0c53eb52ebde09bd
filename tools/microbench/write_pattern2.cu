// write_pattern2.cu -- sensitivity of the output write pattern to the trajectory length T1 (= stride between
// systems) and to the tile -> system mapping (consecutive systems per warp, or systems interleaved
// across the batch).  Companion experiment of write_pattern.cu.
#include <cstdio>
#include <cuda_runtime.h>

template <int W, bool INTERLEAVE>
__global__ void __launch_bounds__(64) k_write(double *out, long B, int T1) {
  const int lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * 2 + threadIdx.x / 32;
  const long ntiles = B / 32;
  if (tile >= ntiles) return;
  constexpr int C = W / 2;
  for (int t = T1 - 1; t >= 0; t--) {
#pragma unroll 4
    for (int q = lane; q < 32 * C; q += 32) {
      const int s = q / C, c = q - s * C;
      const long b = INTERLEAVE ? tile + (long)s * ntiles : tile * 32 + s;
      *reinterpret_cast<double2 *>(out + (b * T1 + t) * W + 2 * c) = make_double2((double)t, (double)c);
    }
  }
}

template <int W, bool INTERLEAVE>
void run(double *out, long B, int T1) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned grid = (unsigned)((B / 32 + 1) / 2);
  float best = 1e30f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_write<W, INTERLEAVE><<<grid, 64>>>(out, B, T1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  printf("{\"W\": %d, \"T1\": %d, \"interleave\": %d, \"ms\": %.3f, \"GBps\": %.0f}\n", W, T1, (int)INTERLEAVE, best,
         (double)B * T1 * W * 8 / best / 1e6);
}

int main() {
  const long B = 65536;
  double *out;
  if (cudaMalloc(&out, (size_t)B * 1024 * 36 * 8) != cudaSuccess) return 1;
  const int t1s[] = {400, 401, 384, 512, 399, 1001};
  for (int T1 : t1s) {
    run<36, false>(out, B, T1);
    run<36, true>(out, B, T1);
  }
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
