// write_pattern.cu -- how fast can B200 absorb the smoother's OUTPUT PATTERN with no compute at all?
// 65,536 systems, each a contiguous trajectory of T1 x W doubles (var: W = 36); at "time step" t every warp
// writes the W-double run of its 32 systems (16-byte stores, neighbouring lanes = neighbouring chunks of one
// run), t descending -- the access pattern of kalman_backward_staged's variance output.  KT time steps are
// written per system before moving to the next system group to see what run length buys.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_pattern write_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

// store variants: 0 plain, 1 st.cs (streaming), 2 st.wt, 3 L2 evict_last hint, 4 L2 evict_first hint
template <int MODE>
__device__ __forceinline__ void store16(double *p, double2 v, unsigned long long pol) {
  if (MODE == 1) __stcs(reinterpret_cast<double2 *>(p), v);
  else if (MODE == 2) __stwt(reinterpret_cast<double2 *>(p), v);
  else if (MODE >= 3)
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
  else *reinterpret_cast<double2 *>(p) = v;
}

template <int W, int MODE>
__global__ void __launch_bounds__(64) k_write_hint(double *out, long B, int T1) {
  const int lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * 2 + threadIdx.x / 32;
  const long b0 = tile * 32;
  if (b0 >= B) return;
  unsigned long long pol = 0;
  if (MODE == 3) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  if (MODE == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  constexpr int C = W / 2;
  for (int t = T1 - 1; t >= 0; t--) {
#pragma unroll 4
    for (int q = lane; q < 32 * C; q += 32) {
      const int s = q / C, c = q - s * C;
      store16<MODE>(out + ((b0 + s) * T1 + t) * W + 2 * c, make_double2((double)t, (double)c), pol);
    }
  }
}

template <int W, int MODE>
void run_hint(double *out, long B, int T1, const char *name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned grid = (unsigned)((B / 32 + 1) / 2);
  float best = 1e30f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_write_hint<W, MODE><<<grid, 64>>>(out, B, T1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  printf("{\"pattern\": \"%s\", \"W\": %d, \"ms\": %.3f, \"GBps\": %.0f}\n", name, W, best, (double)B * T1 * W * 8 / best / 1e6);
}

template <int W, int KT>
__global__ void __launch_bounds__(64) k_write(double *out, long B, int T1) {
  const int lane = threadIdx.x % 32;
  const long tile = (long)blockIdx.x * 2 + threadIdx.x / 32;
  const long b0 = tile * 32;
  if (b0 >= B) return;
  constexpr int C = W * KT / 2;   // 16-byte chunks per system per iteration
  for (int t = T1 - KT; t >= 0; t -= KT) {
#pragma unroll 4
    for (int q = lane; q < 32 * C; q += 32) {
      const int s = q / C, c = q - s * C;
      double2 v = make_double2((double)t, (double)c);
      *reinterpret_cast<double2 *>(out + ((b0 + s) * T1 + t) * W + 2 * c) = v;
    }
  }
}

template <int W, int KT>
void run(double *out, long B, int T1, const char *name) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned grid = (unsigned)((B / 32 + 1) / 2);
  float best = 1e30f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0);
    k_write<W, KT><<<grid, 64>>>(out, B, T1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it && ms < best) best = ms;
  }
  const double bytes = (double)B * (T1 / KT * KT) * W * 8;
  printf("{\"pattern\": \"%s\", \"W\": %d, \"steps_per_visit\": %d, \"ms\": %.3f, \"GBps\": %.0f}\n", name, W, KT, best,
         bytes / best / 1e6);
}

int main(int argc, char **argv) {
  const bool quick = argc > 1;   // --quick: only the first line (bench.py)
  const long B = 65536;
  const int T1 = 400;
  double *out;
  if (cudaMalloc(&out, (size_t)B * T1 * 48 * 8) != cudaSuccess) return 1;
  run<36, 1>(out, B, T1, "var runs, one step per visit");
  if (quick) return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
  run<36, 2>(out, B, T1, "var runs, two steps per visit");
  run<36, 4>(out, B, T1, "var runs, four steps per visit");
  run<36, 8>(out, B, T1, "var runs, eight steps per visit");
  run_hint<36, 1>(out, B, T1, "var runs, st.global.cs");
  run_hint<36, 2>(out, B, T1, "var runs, st.global.wt");
  run_hint<36, 3>(out, B, T1, "var runs, L2 evict_last hint");
  run_hint<36, 4>(out, B, T1, "var runs, L2 evict_first hint");
  run<6, 1>(out, B, T1, "mean runs, one step per visit");
  run<6, 2>(out, B, T1, "mean runs, two steps per visit");
  run<6, 4>(out, B, T1, "mean runs, four steps per visit");
  run<6, 8>(out, B, T1, "mean runs, eight steps per visit");
  run<32, 1>(out, B, T1, "256-byte line-aligned runs");
  run<16, 1>(out, B, T1, "128-byte line-aligned runs");
  run<48, 1>(out, B, T1, "384-byte line-aligned runs");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
