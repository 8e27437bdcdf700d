// Probe: can a K-major SWIZZLE_128B operand start at a row that is NOT a multiple of 8 (not 1 KB-aligned), with 8-row
// groups a non-multiple of 1 KB apart? That is what reusing one haloed input box across the kw taps of a convolution
// needs (conv.cu: tap kw = the same box read from row kw on, tile rows 8 wide inside 10-row box lines -> SBO = 1280 B).
// One CTA, A = 512 rows x 64 bf16 written the way TMA writes a box (16-byte chunk c of row r at chunk c ^ (r & 7)),
// B = 64 x 64 identity, so D[m][n] = A[row read for m][n]. For every (row shift, SBO, base_offset) variant the program
// reports how many of the 128 x 64 outputs equal A[shift + (m / 8) * (SBO / 128) + m % 8][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lisec_b200/csrc tools/umma_shift_probe.cu -o tools/umma_shift_probe
#include <cuda_bf16.h>

#include <cstdio>
#include <vector>

#include "umma.cuh"

using namespace lisec::umma;

constexpr int ROWS = 512;
__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 7 + k * 3) % 251); }

__device__ inline uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ inline void mma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int shift, int sbo, int base_off, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // 512 x 128 B = 64 KB
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + ROWS * 128);    // 64 x 128 B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ROWS * 128 + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < ROWS * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    A[r * 64 + ((((k >> 3) ^ (r & 7)) << 3) | (k & 7))] = __float2bfloat16(a_val(r, k));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    B[n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))] = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  fence_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init_fence();
  }
  if (threadIdx.x < 32) tmem_alloc<64>(slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < 4; ++j)
      mma_bf16(tmem, make_desc(base + shift * 128 + 32 * j, sbo, base_off), make_desc(base + ROWS * 128 + 32 * j, 1024, 0),
               idesc, j != 0);
    mma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    float x[32];
    tmem_ld_32x32(tmem + ((uint32_t)(32 * warp) << 16) + c0, x);
    for (int i = 0; i < 32; ++i) out[(32 * warp + lane) * 64 + c0 + i] = x[i];
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<64>(tmem);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  const int smem = ROWS * 128 + 64 * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 64);
  const int shifts[] = {0, 1, 3, 10, 21}, sbos[] = {1024, 1280};
  for (int sbo : sbos)
    for (int shift : shifts)
      for (int mode = 0; mode < 2; ++mode) {
        const int bo = mode ? (shift & 7) : 0;
        probe<<<1, 128, smem>>>(shift, sbo, bo, d);
        if (cudaDeviceSynchronize() != cudaSuccess) {
          printf("sbo %d shift %d base_off %d: CUDA error %s\n", sbo, shift, bo, cudaGetErrorString(cudaGetLastError()));
          return 1;
        }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 0, rows_ok = 0;
        for (int m = 0; m < 128; ++m) {
          const int r = shift + (m / 8) * (sbo / 128) + m % 8;
          int row_ok = 0;
          for (int n = 0; n < 64; ++n) row_ok += h[m * 64 + n] == a_val(r, n);
          ok += row_ok;
          rows_ok += row_ok == 64;
        }
        printf("sbo %4d  row shift %2d  base_offset %d : %4d / 8192 values, %3d / 128 rows as a plain linear read\n", sbo,
               shift, bo, ok, rows_ok);
      }
  return 0;
}
