// Probe of the tcgen05 path the VFE kernel's dense_2 stage uses (lisec_b200/csrc/umma.cuh): one CTA computes
//     D^T[64 ch][256 rows] = W^T[64 x 64] * X^T[64 x 256]        (kind::tf32, M=64, N=256, 8 k-steps of K=8)
// with both operands MN-major / 128-byte swizzle in shared memory, 3xTF32 operand splitting, accumulators in TMEM,
// and reads them back with tcgen05.ld.32x32b. Checks the result against a float64 host reference and reports the
// error of 1xTF32, 3xTF32 (small terms first / last) and of a float32 FMA chain, plus the MMA latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lisec_b200/csrc tools/umma_probe.cu -o tools/umma_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace lisec::umma;

// ---- MN-major / plain SWIZZLE_128B variants (hypotheses 0 and 1: they yield zeros for tf32) ----
// byte offset of element (mn, k) inside an operand whose k-atoms are k_atom_stride bytes apart
__device__ __host__ __forceinline__ uint32_t op_offset(int mn, int k, uint32_t k_atom_stride) {
  return (uint32_t)(k >> 3) * k_atom_stride + (uint32_t)(mn >> 5) * 1024u + (uint32_t)(k & 7) * 128u +
         (uint32_t)((((mn & 31) >> 2) ^ (k & 7)) << 4) + (uint32_t)(mn & 3) * 4u;
}

// Shared-memory matrix descriptor (64 bit): start address, leading / stride byte offsets (all >> 4), version 1
// (Blackwell), layout type 2 = SWIZZLE_128B. For an MN-major swizzled operand the "leading" offset is the distance
// between mn-atoms (1 KB here) and the "stride" offset the distance between k-atoms.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t mn_atom_stride,
                                                       uint32_t k_atom_stride) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((mn_atom_stride >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((k_atom_stride >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::tf32, float32 accumulation, both operands MN-major.
__device__ __host__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
  return (1u << 4)                      // D format: F32
         | (2u << 7) | (2u << 10)       // A, B format: TF32
         | (1u << 15) | (1u << 16)      // A, B major: MN
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}


constexpr int N = 256, K = 64, M = 64;
constexpr uint32_t X_KSTRIDE = (N / 32) * 1024;  // 8 KB between k-atoms of X
constexpr uint32_t W_KSTRIDE = (M / 32) * 1024;  // 2 KB between k-atoms of W^T
constexpr int X_BYTES = (K / 8) * X_KSTRIDE;     // 64 KB
constexpr int W_BYTES = (K / 8) * W_KSTRIDE;     // 16 KB
constexpr int SMEM = 2 * X_BYTES + 2 * W_BYTES + 1024 + 64;

// K-major, 128-byte swizzle (the layout every bf16 GEMM uses): row mn = 128 bytes = 32 tf32 of k, 8-row groups of 1 KB,
// 16-byte chunks XOR-swizzled with mn % 8; a k-block of 32 is one such slab, slabs `slab` bytes apart.
__device__ __host__ inline uint32_t kmaj_offset(int mn, int k, uint32_t slab) {
  return (uint32_t)(k >> 5) * slab + (uint32_t)mn * 128u + (uint32_t)((((k & 31) >> 2) ^ (mn & 7)) << 4) +
         (uint32_t)(k & 3) * 4u;
}
__device__ inline uint64_t probe_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;             // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// mode: 0 = 3xTF32, small terms first; 1 = 3xTF32, big term first; 2 = 1xTF32 (hi*hi only)
// swap: 1 = exchange the leading / stride byte offsets in both descriptors (layout hypothesis check)
__global__ void __launch_bounds__(128) probe(const float* __restrict__ X, const float* __restrict__ W,
                                             float* __restrict__ D, long long* __restrict__ cycles, int mode,
                                             int swap) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sXh = smem;
  unsigned char* sXl = sXh + X_BYTES;
  unsigned char* sWh = sXl + X_BYTES;
  unsigned char* sWl = sWh + W_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sWl + W_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    float hi, lo;
    tf32_split(X[i], hi, lo);
    const uint32_t o = swap == 2 ? kmaj_offset(n, k, N * 128) : op_offset(n, k, X_KSTRIDE);
    *reinterpret_cast<float*>(sXh + o) = hi;
    *reinterpret_cast<float*>(sXl + o) = lo;
  }
  for (int i = tid; i < K * M; i += blockDim.x) {
    const int k = i / M, m = i % M;  // W is (C_in, C_out) row-major, as Keras stores it
    float hi, lo;
    tf32_split(W[i], hi, lo);
    const uint32_t o = swap == 2 ? kmaj_offset(m, k, M * 128) : op_offset(m, k, W_KSTRIDE);
    *reinterpret_cast<float*>(sWh + o) = hi;
    *reinterpret_cast<float*>(sWl + o) = lo;
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init_fence();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  {  // TMEM store -> load round trip on columns 480..511 (outside the accumulator)
    uint32_t r[4];
    for (int j = 0; j < 4; ++j) r[j] = 1000u * (warp * 32 + lane) + j;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                     tbase + ((uint32_t)(warp * 32) << 16) + 480),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t q[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3])
                 : "r"(tbase + ((uint32_t)(warp * 32) << 16) + 480)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    int ok = 1;
    for (int j = 0; j < 4; ++j) ok &= (q[j] == r[j]);
    if (!__all_sync(0xffffffffu, ok) && lane == 0) cycles[1] = -1;
    if (tid == 0) cycles[2] = tbase;
  }
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = swap == 2 ? (make_idesc_tf32_mn(M, N) & ~((1u << 15) | (1u << 16))) : make_idesc_tf32_mn(M, N);
    auto desc = [&](unsigned char* base, int kb, uint32_t kstride) {
      if (swap == 2)  // K-major: k-block kb/4 is a slab of (kstride/1024*32) rows x 128 B, k-step inside it = +32 B
        return probe_desc_k_sw128(smem_u32(base) + (kb >> 2) * (kstride / 1024 * 32 * 128) + (kb & 3) * 32);
      const uint32_t addr = smem_u32(base) + kb * kstride;
      return swap ? make_desc_mn_sw128(addr, kstride, 1024) : make_desc_mn_sw128(addr, 1024, kstride);
    };
    t0 = clock64();
    uint32_t acc = 0;
    auto pass = [&](unsigned char* w, unsigned char* x) {
      for (int kb = 0; kb < K / 8; ++kb) {
        mma_tf32_ss(tbase, desc(w, kb, W_KSTRIDE), desc(x, kb, X_KSTRIDE), idesc, acc);
        acc = 1;
      }
    };
    if (mode == 0) {
      pass(sWh, sXl);
      pass(sWl, sXh);
      pass(sWh, sXh);
    } else if (mode == 1) {
      pass(sWh, sXh);
      pass(sWh, sXl);
      pass(sWl, sXh);
    } else {
      pass(sWh, sXh);
    }
    mma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  fence_after_sync();
  if (tid == 0) cycles[0] = clock64() - t0;

  // warp q reads TMEM lanes 32q..32q+31; for M = 64 lanes 0..15 of each quadrant hold channels 16q..16q+15
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];  // D is [128 tmem lanes][256]
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  std::vector<float> X(N * K), W(K * M);
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX; };
  for (auto& x : X) x = rnd() < 0.3f ? 0.f : rnd() * 3.f;        // post-ReLU activations
  for (auto& w : W) w = (rnd() - 0.5f) * 0.6f;
  std::vector<double> ref(M * N);
  std::vector<float> chain(M * N);
  double scale = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      float c = 0.f;
      for (int k = 0; k < K; ++k) {
        s += (double)W[k * M + m] * X[n * K + k];
        c = fmaf(W[k * M + m], X[n * K + k], c);
      }
      ref[m * N + n] = s;
      chain[m * N + n] = c;
      scale += s * s;
    }
  scale = std::sqrt(scale / (M * N));
  double e_chain = 0;
  for (int i = 0; i < M * N; ++i) e_chain = std::fmax(e_chain, std::fabs(chain[i] - ref[i]));
  printf("rms(ref) = %.4f; float32 FMA chain: max abs err %.3e (%.3e of rms)\n", scale, e_chain, e_chain / scale);

  float *dX, *dW, *dD;
  long long* dC;
  cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dW, W.size() * 4);
  cudaMalloc(&dD, 128 * N * 4);
  cudaMalloc(&dC, 64);
  cudaMemset(dC, 0, 64);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const char* names[3] = {"3xTF32 small-first", "3xTF32 big-first", "1xTF32"};
  for (int swap = 0; swap < 3; ++swap)
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemset(dD, 0xff, 128 * N * 4);
      probe<<<1, 128, SMEM>>>(dX, dW, dD, dC, mode, swap);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("swap=%d %s: CUDA error %s\n", swap, names[mode], cudaGetErrorString(e));
        return 1;
      }
      std::vector<float> D(128 * N);
      long long cyc = 0, info[3];
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(info, dC, 24, cudaMemcpyDeviceToHost);
      cyc = info[0];
      int nz = 0, nan = 0;
      for (float d : D) { nz += (d != 0.f); nan += (d != d); }
      printf("  [tmem base 0x%llx, st/ld round trip %s, %d nonzero / %d NaN of %d dumped words]\n", info[2],
             info[1] == 0 ? "ok" : "FAILED", nz, nan, (int)D.size());
      double err = 0;
      int bad = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          const int tl = (m % 16) + 32 * (m / 16);  // M = 64: channel m lives in TMEM lane (m % 16) + 32 (m / 16)
          const double d = std::fabs((double)D[tl * N + n] - ref[m * N + n]);
          if (!(d <= 1e-2 * scale)) ++bad;
          if (d == d) err = std::fmax(err, d);
        }
      printf("swap=%d %-20s: max abs err %.3e (%.3e of rms), %d of %d off by > 1%%, issue->done %lld cycles\n", swap,
             names[mode], err, err / scale, bad, M * N, cyc);
      if (mode == 0 && swap == 0) {
        printf("  D^T[0][0..3] = %.6f %.6f %.6f %.6f | ref %.6f %.6f %.6f %.6f\n", D[0], D[1], D[2], D[3], ref[0], ref[1],
               ref[2], ref[3]);
        printf("  tmem lane 16 (unused half of quadrant 0) col 0: %g\n", D[16 * N]);
      }
    }
  return 0;
}
