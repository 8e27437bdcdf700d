// Pure-register FP32 pipe peak on sm_100a: FFMA (3-register, with operand reuse) vs FFMA2 (fma.rn.f32x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: FFMA 8x8 outer product, 1: FFMA2 col pairs with scalar broadcast, 2: FFMA2 both packed
__global__ void __launch_bounds__(256) k_peak(float* out, int iters, float seed) {
  float a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = seed - i * 0.5f; }
  float2 acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 0) {
          acc[i][j].x = fmaf(a[i], b[2 * j], acc[i][j].x);
          acc[i][j].y = fmaf(a[i], b[2 * j + 1], acc[i][j].y);
        } else if (MODE == 1) {
          acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), make_float2(b[2 * j], b[2 * j + 1]), acc[i][j]);
        } else {
          acc[i][j] = __ffma2_rn(make_float2(a[i], a[(i + 1) & 7]), make_float2(b[2 * j], b[2 * j + 1]), acc[i][j]);
        }
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a[i] * 0.999f;  // keep the loop from being hoisted
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d; cudaMalloc(&d, sizeof(float) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas = 1; ctas <= 8; ctas *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_peak<0><<<sms * ctas, 256>>>(d, iters, 1.f);
        if (mode == 1) k_peak<1><<<sms * ctas, 256>>>(d, iters, 1.f);
        if (mode == 2) k_peak<2><<<sms * ctas, 256>>>(d, iters, 1.f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
      }
      const double fma = (double)sms * ctas * 256 * iters * 64;
      printf("mode %d (%s) ctas/SM %d: %.3f ms  %.2f TFLOP/s  %.1f FMA/clk/SM @1.965GHz\n", mode,
             mode == 0 ? "FFMA" : mode == 1 ? "FFMA2 bcast-a" : "FFMA2 packed", ctas, best, 2 * fma / best / 1e9,
             fma / sms / (best * 1e-3 * 1.965e9));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
