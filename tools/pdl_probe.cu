// What a kernel boundary costs under programmatic dependent launch: chains of 5 near-empty kernels with the grouping
// chain's grid sizes (782 / 2500 / 2500 / 3200 / 3125 blocks of 256 threads), each thread doing one dependent L2 round
// trip on data its predecessor wrote, timed back to back with CUDA events. Prints us per chain and per kernel for plain
// stream order and for programmatic launches.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) hop(const int* __restrict__ in, int* __restrict__ out, int n, int pdl) {
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldcg(in + i) + 1;
}

static float run(int pdl, int reps, int* a, int* b, int n, const int* blocks) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto chain = [&]() {
    for (int k = 0; k < 5; ++k) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(blocks[k]);
      cfg.blockDim = dim3(256);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl ? 1 : 0;
      const int* in = (k & 1) ? b : a;
      int* out = (k & 1) ? a : b;
      cudaLaunchKernelEx(&cfg, hop, in, out, n, pdl);
    }
  };
  for (int i = 0; i < 20; ++i) chain();
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) chain();
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1000.f / reps;
}

int main() {
  const int n = 800000;
  int *a, *b;
  cudaMalloc(&a, n * 4);
  cudaMalloc(&b, n * 4);
  cudaMemset(a, 0, n * 4);
  const int chain_blocks[5] = {782, 2500, 2500, 3200, 3125}, small[5] = {148, 148, 148, 148, 148};
  for (int pdl = 0; pdl < 2; ++pdl) {
    const float t = run(pdl, 2000, a, b, n, chain_blocks), s = run(pdl, 2000, a, b, n, small);
    printf("%s: chain-sized grids %.2f us per 5 kernels (%.2f per kernel); 148-block grids %.2f us (%.2f per kernel)\n",
           pdl ? "programmatic launch" : "plain stream order ", t, t / 5, s, s / 5);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
