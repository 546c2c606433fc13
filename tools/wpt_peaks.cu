// wpt_peaks — the two roofs SURVEY.md 8(d) asks to be measured on the box next to the path tracer's numbers:
//   * L2 -> SM read bandwidth: every block streams through a buffer that fits the L2 (32 MiB) again and again with
//     128-bit loads (the node / shape fetch width of the traversal), HBM only sees the first pass; once bypassing the L1
//     (ld.global.cg) and once through it (each SM's share of the buffer fits its L1: that is the L1 roof);
//   * sustained FP32 rate: dependent FFMA chains, 8 independent accumulators per thread, all SMs resident;
//   * (for reference) HBM read bandwidth: the same kernel over a 4 GiB buffer.
// Prints one JSON line. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/wpt_peaks.cu -o tools/wpt_peaks
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); std::exit(1); } } while (0)

template <bool BYPASS_L1>
__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ buf, size_t n4, int passes, float* sink) {
  float acc = 0.0f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; p++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 v = BYPASS_L1 ? __ldcg(buf + i) : __ldg(buf + i);   // ld.global.cg: cached in L2 only
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;   // never true: keeps the loads alive
}

__global__ void __launch_bounds__(256) k_fma(int iters, float* sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float m = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll 8
    for (int k = 0; k < 8; k++) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456f) *sink = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float* sink; CK(cudaMalloc(&sink, 4));
  // ---- L2: 32 MiB buffer, 64 passes
  double l2_gbs = 0, l1_gbs = 0, hbm_gbs = 0, tflops = 0;
  {
    size_t bytes = 32ull << 20; float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    int passes = 64, grid = sms * 8;
    k_read<true><<<grid, 256>>>(buf, bytes / 16, 2, sink);   // warm the L2
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_read<true><<<grid, 256>>>(buf, bytes / 16, passes, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      double g = (double)bytes * passes / (time_ms(e0, e1) * 1e-3) / 1e9; if (g > l2_gbs) l2_gbs = g;
    }
    // the same with ld.global.nc: every SM re-reads its own 221 KB share, which stays in its L1
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_read<false><<<grid, 256>>>(buf, bytes / 16, passes, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      double g = (double)bytes * passes / (time_ms(e0, e1) * 1e-3) / 1e9; if (g > l1_gbs) l1_gbs = g;
    }
    CK(cudaFree(buf));
  }
  // ---- HBM: 4 GiB buffer, one pass
  {
    size_t bytes = 4ull << 30; float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
    int grid = sms * 8;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_read<true><<<grid, 256>>>(buf, bytes / 16, 1, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      double g = (double)bytes / (time_ms(e0, e1) * 1e-3) / 1e9; if (g > hbm_gbs) hbm_gbs = g;
    }
    CK(cudaFree(buf));
  }
  // ---- FP32: 8 blocks x 256 threads per SM, 8 chains per thread
  {
    int iters = 4096, grid = sms * 8;
    k_fma<<<grid, 256>>>(16, sink);
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaEventRecord(e0)); k_fma<<<grid, 256>>>(iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      double flops = 2.0 * 64.0 * iters * (double)grid * 256.0;
      double t = flops / (time_ms(e0, e1) * 1e-3) / 1e12; if (t > tflops) tflops = t;
    }
  }
  int clock_khz = 0; CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, \"l2_read_gbs\": %.1f, \"l1_read_gbs\": %.1f, \"hbm_read_gbs\": %.1f, \"fp32_fma_tflops\": %.2f, "
              "\"fp32_nominal_tflops\": %.2f, \"how\": \"L2: 32 MiB buffer x 64 passes of 128-bit ld.global.cg (L1 bypassed), L1: the same with ld.global.nc (each SM re-reads its 221 KB share), best of 3; HBM: 4 GiB x 1 pass; FP32: 8 FFMA chains per thread, 2048 threads per SM, best of 3\"}\n",
              prop.name, sms, clock_khz / 1e3, l2_gbs, l1_gbs, hbm_gbs, tflops, sms * 128.0 * 2.0 * clock_khz * 1e3 / 1e12);
  return 0;
}
