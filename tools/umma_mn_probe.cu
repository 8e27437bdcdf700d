// Probe for the weight-gradient GEMM of the training step: both operands MN-major.
//   dW[co][ci] = sum over positions p of dY[p][co] * X[p][ci]
// K = positions; dY and X are channels-last, so a TMA box of one operand is [positions][64 channels] bf16 with the
// 128-byte swizzle: 128-byte rows indexed by k, the MN index inside the row — the "Major-MN, SWIZZLE_128B" canonical
// layout ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)) of cute/atom/mma_traits_sm100.hpp. This program fills two such operands
// (A: K x 128 as two 64-channel boxes, B: K x 128 likewise) exactly the way TMA would, runs K/16 tcgen05.mma.kind::f16
// with a_major = b_major = 1 for every assignment of (LBO, SBO) and k-step, and reports which one yields A^T B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lisec_b200/csrc tools/umma_mn_probe.cu -o tools/umma_mn_probe
#include <cuda_bf16.h>

#include <cstdio>
#include <vector>

#include "umma.cuh"

using namespace lisec::umma;

constexpr int KP = 64;                      // positions (K)
constexpr int AROWS = KP + 32;              // the A boxes are taller: the K rows may start at any row `shift` of them
constexpr int BOX = AROWS * 128;            // bytes of one [AROWS][64] bf16 box (B uses the first KP rows of its boxes)
__host__ __device__ inline float a_val(int k, int m) { return (float)((k * 5 + m * 3) % 7); }
__host__ __device__ inline float b_val(int k, int n) { return (float)((k * 3 + n * 11) % 5) - 2.f; }

__device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
__device__ inline void mma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

// element (k, mn) of an operand made of 64-channel boxes: box mn / 64, row k, 16-byte chunk (mn % 64) / 8 swizzled with k % 8
__device__ inline void put(__nv_bfloat16* base, int k, int mn, float v) {
  const int box = mn >> 6, c = mn & 63;
  base[box * (BOX / 2) + k * 64 + ((((c >> 3) ^ (k & 7)) << 3) | (c & 7))] = __float2bfloat16(v);
}

__global__ void __launch_bounds__(128, 1) probe(int lbo, int sbo, int kstep, int shift, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // two boxes
  __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 2 * BOX);  // two boxes
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * BOX);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < AROWS * 128; i += 128) {
    const int k = i >> 7, mn = i & 127;
    put(A, k, mn, a_val(k, mn));          // row k of the A boxes holds "position" k; the MMA reads rows shift .. shift + KP - 1
    if (k < KP) put(B, k, mn, b_val(k, mn));
  }
  fence_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init_fence();
  }
  if (threadIdx.x < 32) tmem_alloc<128>(slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    // kind::f16, bf16 x bf16 -> f32, a_major = b_major = MN (bits 15, 16), N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < KP / 16; ++j)
      mma_bf16(tmem, make_desc(base + shift * 128 + kstep * j, lbo, sbo), make_desc(base + 2 * BOX + kstep * j, lbo, sbo), idesc, j != 0);
    mma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    float x[32];
    tmem_ld_32x32(tmem + ((uint32_t)(32 * warp) << 16) + c0, x);
    for (int i = 0; i < 32; ++i) out[(32 * warp + lane) * 128 + c0 + i] = x[i];
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<128>(tmem);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 128 * 4);
  const int smem = 4 * BOX + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 128), want(128 * 128);
  auto reference = [&](int shift) {
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        float s = 0.f;
        for (int k = 0; k < KP; ++k) s += a_val(shift + k, m) * b_val(k, n);
        want[m * 128 + n] = s;
      }
  };
  // Measured on B200: the first line is exact (16384 / 16384); the others are not. (Candidates that step past the
  // operand, e.g. SBO = 2048, fault — they are not in the list.)
  const int cands[][3] = {{BOX, 1024, 2048}, {1024, BOX, 2048}, {BOX, 1024, 256}, {1024, BOX, 256}};
  // row shifts: 0 (aligned), 8 (a whole swizzle atom), and starts in the middle of an atom — what reusing one haloed X
  // box across the (kh, kw) taps of the weight gradient needs (as conv_halo_kernel does for the K-major forward operand)
  // Overlapping M blocks: a small LBO would make rows 64..127 of the M = 128 operand the SAME 64 channels read a few rows
  // further down — two filter taps of one haloed box stacked into one MMA. Measured on B200: block 0 is exact, block 1 is
  // NOT, for 128 B, 1 KB, 2304 B and 3 KB alike: M blocks must be separate boxes. A halo-box weight gradient therefore
  // issues M = 64 MMAs per tap (or loads the box twice).
  for (int lbo : {128, 18 * 128, 8 * 128, 24 * 128}) {
    const int shift = 3;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        float acc = 0.f;
        const int extra = m >= 64 ? lbo / 128 : 0;
        for (int k = 0; k < KP; ++k) acc += a_val(shift + extra + k, m & 63) * b_val(k, n);
        want[m * 128 + n] = acc;
      }
    probe<<<1, 128, smem>>>(lbo, 1024, 2048, shift, d);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("overlapping blocks, LBO %d: CUDA error %s\n", lbo, cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int ok = 0;
    for (int i = 0; i < 128 * 128; ++i) ok += h[i] == want[i];
    printf("overlapping M blocks: row shift %d, LBO %4d B (block 1 = block 0 moved %2d rows down): %5d / 16384 exact\n", shift, lbo,
           lbo / 128, ok);
  }
  const int shifts[] = {0, 8, 1, 3, 19};
  for (int shift : shifts)
  for (auto& c : cands) {
    if (shift && &c != &cands[0]) continue;
    reference(shift);
    probe<<<1, 128, smem>>>(c[0], c[1], c[2], shift, d);
    if (cudaDeviceSynchronize() != cudaSuccess) {
      printf("lbo %5d sbo %5d kstep %4d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(cudaGetLastError()));
      return 1;
    }
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int ok = 0, q[4] = {0, 0, 0, 0};
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        const bool e = h[m * 128 + n] == want[m * 128 + n];
        ok += e;
        q[(m >> 6) * 2 + (n >> 6)] += e;
      }
    printf("row shift %2d  LBO %5d  SBO %5d  k-step %4d B : %5d / 16384 exact  (quadrants m<64,n<64: %4d  m<64,n>=64: %4d  m>=64,n<64: %4d  "
           "m>=64,n>=64: %4d)\n", shift, c[0], c[1], c[2], ok, q[0], q[1], q[2], q[3]);
  }
  return 0;
}
