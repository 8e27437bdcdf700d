// DFMA issue rate per SM in the shapes the VFE front stage uses: 8 warps per SM, 16 independent accumulators per
// thread, 6 k-steps. Variants: register operand, kernel-parameter (constant bank -> uniform register) operand as in
// vfe_kernel's P.w1, shared-memory operand. Also FFMA for scale.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/dfma_probe.cu -o tools/dfma_probe
#include <cstdio>
struct W { double w[6][16]; };
template <int VAR>
__global__ void kd(double* out, long long* cyc, const __grid_constant__ W P, double w0) {
  __shared__ double sw[6][16];
  if (threadIdx.x < 96) sw[threadIdx.x / 16][threadIdx.x % 16] = P.w[threadIdx.x / 16][threadIdx.x % 16];
  double f[6];
  for (int k = 0; k < 6; ++k) f[k] = threadIdx.x * 0.001 + k;
  double tot = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < 64; ++it) {
    double d[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) d[j] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const double w = VAR == 0 ? w0 + j : VAR == 1 ? P.w[k][j] : sw[k][j];
        d[j] = fma(f[k], w, d[j]);
      }
#pragma unroll
    for (int j = 0; j < 16; ++j) tot += d[j];
    for (int k = 0; k < 6; ++k) f[k] += 1e-3 * tot;
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double* od; long long* c;
  cudaMalloc(&od, 148 * 512 * 8); cudaMalloc(&c, 148 * 8);
  W P;
  for (int k = 0; k < 6; ++k) for (int j = 0; j < 16; ++j) P.w[k][j] = 0.01 * (k + j);
  const char* names[3] = {"register operand", "kernel-param (uniform register) operand", "shared-memory operand"};
  for (int var = 0; var < 3; ++var) {
    long long h[148];
    if (var == 0) kd<0><<<148, 256>>>(od, c, P, 0.5);
    if (var == 1) kd<1><<<148, 256>>>(od, c, P, 0.5);
    if (var == 2) kd<2><<<148, 256>>>(od, c, P, 0.5);
    cudaDeviceSynchronize();
    cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-42s: %6lld cycles per 64 x (96 DFMA + 16 DADD) per thread, 8 warps/SM -> %.1f DP lanes/clk/SM\n", names[var], h[0],
           256.0 * 64 * 112 / h[0]);
  }
  return 0;
}
