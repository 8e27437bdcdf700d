// Calibration probe (not part of the library): which inner-loop form reaches the FP32 pipe on sm_100a for the
// VFE's small dense layers? D[M,64] = relu(H[M,32] * W[32,64]), M rows, fp32.
//   U : thread per row, weights as __grid_constant__ (uniform-register operands), k-outer / j-inner
//   S : shared-memory register tiling, 256-row tile, 8x8 micro-tile per thread, operands via LDS.128
//   S4: same with 8 rows x 4 cols (two passes over N), the shape layer 2 would use
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_probe ffma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

struct W { float w[32][64]; };
#ifndef REP
#define REP 1
#endif

__global__ void __launch_bounds__(256, 2) k_uniform(const __grid_constant__ W w, const float* __restrict__ H,
                                                    float* __restrict__ D, int M) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += gridDim.x * blockDim.x) {
    float h[32];
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      float4 v = *reinterpret_cast<const float4*>(H + (size_t)r * 32 + k);
      h[k] = v.x; h[k + 1] = v.y; h[k + 2] = v.z; h[k + 3] = v.w;
    }
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
    for (int k = 0; k < 32; ++k)
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = fmaf(h[k], w.w[k][j], acc[j]);
#pragma unroll
    for (int k = 0; k < 32; ++k) h[k] *= 0.5f;
    }
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) m += fmaxf(acc[j], 0.f);
    D[r] = m;  // reduced output: the probe measures the math pipe, not the store path
  }
}

constexpr int TM = 256, PITCH = TM + 4;
// A tile k-major in smem: sA[k][row]; W in smem: sW[k][col]
template <int COLS_PER_THREAD>
__global__ void __launch_bounds__(256, 2) k_smem(const float* __restrict__ Wg, const float* __restrict__ H,
                                                 float* __restrict__ D, int M) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;            // [32][64]
  float* sA = smem + 32 * 64;  // [32][PITCH]
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * 64; i += 256) sW[i] = Wg[i];
  constexpr int CG = 64 / COLS_PER_THREAD;       // column groups
  constexpr int PASSES = (CG * 32) / 256;        // 8x8 -> 1 pass; 8x4 -> 2 passes
  for (int tile = blockIdx.x; tile * TM < M; tile += gridDim.x) {
    __syncthreads();
    // load + transpose the H tile: thread t owns row t
    {
      const float* src = H + ((size_t)tile * TM + tid) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        sA[(k + 0) * PITCH + tid] = v.x; sA[(k + 1) * PITCH + tid] = v.y;
        sA[(k + 2) * PITCH + tid] = v.z; sA[(k + 3) * PITCH + tid] = v.w;
      }
    }
    __syncthreads();
    float out = 0.f;
#pragma unroll
    for (int pass = 0; pass < PASSES; ++pass) {
      const int g = tid + pass * 256;
      const int tx = g % CG, ty = (g / CG) % 32;   // ty: row group (8 rows), tx: col group
      float acc[8][COLS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < COLS_PER_THREAD; ++j) acc[i][j] = 0.f;
#pragma unroll 1
      for (int rep = 0; rep < REP; ++rep)
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        float a[8], b[COLS_PER_THREAD];
        float4 a0 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8);
        float4 a1 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8 + 4);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
        for (int j = 0; j < COLS_PER_THREAD; j += 4) {
          float4 bv = *reinterpret_cast<const float4*>(sW + k * 64 + tx * COLS_PER_THREAD + j);
          b[j] = bv.x; b[j + 1] = bv.y; b[j + 2] = bv.z; b[j + 3] = bv.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < COLS_PER_THREAD; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < COLS_PER_THREAD; ++j) out += fmaxf(acc[i][j], 0.f);
    }
    D[(size_t)tile * TM + tid] = out;
  }
}

// ---- packed FFMA2 variants (fma.rn.f32x2: two IEEE fp32 FMAs per instruction) ------------------------------
__global__ void __launch_bounds__(256, 2) k_uniform2(const __grid_constant__ W w, const float* __restrict__ H,
                                                     float* __restrict__ D, int M) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < M; r += gridDim.x * blockDim.x) {
    float h[32];
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      float4 v = *reinterpret_cast<const float4*>(H + (size_t)r * 32 + k);
      h[k] = v.x; h[k + 1] = v.y; h[k + 2] = v.z; h[k + 3] = v.w;
    }
    float2 acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float2 hh = make_float2(h[k], h[k]);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __ffma2_rn(hh, make_float2(w.w[k][2 * j], w.w[k][2 * j + 1]), acc[j]);
    }
    float m = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) m += fmaxf(acc[j].x, 0.f) + fmaxf(acc[j].y, 0.f);
    D[r] = m;
  }
}

// pair along ROWS: acc2[rowpair][col] += (a_i, a_i+1) * (w, w); weights pre-duplicated in smem as float2
__global__ void __launch_bounds__(256, 2) k_smem2(const float* __restrict__ Wg, const float* __restrict__ H,
                                                  float* __restrict__ D, int M) {
  extern __shared__ __align__(16) float smem[];
  float2* sW2 = reinterpret_cast<float2*>(smem);  // [32][64] (w,w)
  float* sA = smem + 2 * 32 * 64;                 // [32][PITCH]
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * 64; i += 256) sW2[i] = make_float2(Wg[i], Wg[i]);
  const int tx = tid % 8, ty = tid / 8;
  for (int tile = blockIdx.x; tile * TM < M; tile += gridDim.x) {
    __syncthreads();
    {
      const float* src = H + ((size_t)tile * TM + tid) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        sA[(k + 0) * PITCH + tid] = v.x; sA[(k + 1) * PITCH + tid] = v.y;
        sA[(k + 2) * PITCH + tid] = v.z; sA[(k + 3) * PITCH + tid] = v.w;
      }
    }
    __syncthreads();
    float2 acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8 + 4);
      const float2 a[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                           make_float2(a1.z, a1.w)};
      float2 b[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float4 bv = *reinterpret_cast<const float4*>(sW2 + k * 64 + tx * 8 + j);
        b[j] = make_float2(bv.x, bv.y); b[j + 1] = make_float2(bv.z, bv.w);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = __ffma2_rn(a[i], b[j], acc[i][j]);
    }
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) out += fmaxf(acc[i][j].x, 0.f) + fmaxf(acc[i][j].y, 0.f);
    D[(size_t)tile * TM + tid] = out;
  }
}

// pair along COLUMNS: acc2[row][colpair] += (a_i, a_i) * (w_j, w_j+1); the (a,a) pair is built on the fly
__global__ void __launch_bounds__(256, 2) k_smem2c(const float* __restrict__ Wg, const float* __restrict__ H,
                                                   float* __restrict__ D, int M) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;
  float* sA = smem + 2 * 32 * 64;
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * 64; i += 256) sW[i] = Wg[i];
  const int tx = tid % 8, ty = tid / 8;
  for (int tile = blockIdx.x; tile * TM < M; tile += gridDim.x) {
    __syncthreads();
    {
      const float* src = H + ((size_t)tile * TM + tid) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        sA[(k + 0) * PITCH + tid] = v.x; sA[(k + 1) * PITCH + tid] = v.y;
        sA[(k + 2) * PITCH + tid] = v.z; sA[(k + 3) * PITCH + tid] = v.w;
      }
    }
    __syncthreads();
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8);
      const float4 a1 = *reinterpret_cast<const float4*>(sA + k * PITCH + ty * 8 + 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float4 b0 = *reinterpret_cast<const float4*>(sW + k * 64 + tx * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(sW + k * 64 + tx * 8 + 4);
      const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                           make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(aa, b[j], acc[i][j]);
      }
    }
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) out += fmaxf(acc[i][j].x, 0.f) + fmaxf(acc[i][j].y, 0.f);
    D[(size_t)tile * TM + tid] = out;
  }
}

// H: warp w owns output columns [8w, 8w+8) with those weights in uniform registers (all lanes use the same
// weights); lane owns rows {lane + 32 i} (VEC=false, 8 conflict-free LDS.32 per k) or rows [8 lane, 8 lane + 8)
// (VEC=true, 2 LDS.128 per k). Every warp streams the whole A tile from shared memory.
template <bool VEC>
__global__ void __launch_bounds__(256, 2) k_hybrid(const __grid_constant__ W w, const float* __restrict__ H,
                                                   float* __restrict__ D, int M) {
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;  // [32][PITCH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int tile = blockIdx.x; tile * TM < M; tile += gridDim.x) {
    __syncthreads();
    {
      const float* src = H + ((size_t)tile * TM + tid) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        sA[(k + 0) * PITCH + tid] = v.x; sA[(k + 1) * PITCH + tid] = v.y;
        sA[(k + 2) * PITCH + tid] = v.z; sA[(k + 3) * PITCH + tid] = v.w;
      }
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    // the column block is warp-uniform: switch so that each case indexes the constant bank statically
#define HYB_BODY(CW)                                                                                   \
    _Pragma("unroll 4") for (int k = 0; k < 32; ++k) {                                                 \
      float a[8];                                                                                      \
      if (VEC) {                                                                                       \
        const float4 a0 = *reinterpret_cast<const float4*>(sA + k * PITCH + lane * 8);                 \
        const float4 a1 = *reinterpret_cast<const float4*>(sA + k * PITCH + lane * 8 + 4);             \
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w; \
      } else {                                                                                         \
        _Pragma("unroll") for (int i = 0; i < 8; ++i) a[i] = sA[k * PITCH + lane + 32 * i];            \
      }                                                                                                \
      _Pragma("unroll") for (int i = 0; i < 8; ++i)                                                    \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w.w[k][(CW) * 8 + j], acc[i][j]); \
    }
    switch (warp) {
      case 0: HYB_BODY(0) break;
      case 1: HYB_BODY(1) break;
      case 2: HYB_BODY(2) break;
      case 3: HYB_BODY(3) break;
      case 4: HYB_BODY(4) break;
      case 5: HYB_BODY(5) break;
      case 6: HYB_BODY(6) break;
      default: HYB_BODY(7) break;
    }
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) out += fmaxf(acc[i][j], 0.f);
    D[(size_t)tile * TM + tid] = out;
  }
}

// Hc / Hs: the hybrid data flow with ONE loop body for all warps. MODE 0: weights by dynamically indexed constant
// load (w.w[k][8*warp + j]); MODE 1: weights from shared memory, warp-uniform address (broadcast LDS.128).
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_hybrid1(const __grid_constant__ W w, const float* __restrict__ Wg,
                                                    const float* __restrict__ H, float* __restrict__ D, int M) {
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;               // [32][PITCH]
  float* sW = smem + 32 * PITCH;  // [32][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * 64; i += 256) sW[i] = Wg[i];
  for (int tile = blockIdx.x; tile * TM < M; tile += gridDim.x) {
    __syncthreads();
    {
      const float* src = H + ((size_t)tile * TM + tid) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        float4 v = *reinterpret_cast<const float4*>(src + k);
        sA[(k + 0) * PITCH + tid] = v.x; sA[(k + 1) * PITCH + tid] = v.y;
        sA[(k + 2) * PITCH + tid] = v.z; sA[(k + 3) * PITCH + tid] = v.w;
      }
    }
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int rep = 0; rep < REP; ++rep)
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sA[k * PITCH + lane + 32 * i];
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = w.w[k][warp * 8 + j];
      } else {
        const float4 b0 = *reinterpret_cast<const float4*>(sW + k * 64 + warp * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(sW + k * 64 + warp * 8 + 4);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    float out = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) out += fmaxf(acc[i][j], 0.f);
    D[(size_t)tile * TM + tid] = out;
  }
}

int main() {
  const int M = 1 << 20;  // about the row count of an 8-sweep batch
  std::vector<float> hH((size_t)M * 32), hW(32 * 64);
  for (size_t i = 0; i < hH.size(); ++i) hH[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
  for (int i = 0; i < 32 * 64; ++i) hW[i] = (float)((i * 40503u) % 1000) / 1000.f - 0.5f;
  float *dH, *dD, *dW;
  cudaMalloc(&dH, hH.size() * 4); cudaMalloc(&dD, (size_t)M * 4); cudaMalloc(&dW, 32 * 64 * 4);
  cudaMemcpy(dH, hH.data(), hH.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, hW.data(), 32 * 64 * 4, cudaMemcpyHostToDevice);
  W w; for (int k = 0; k < 32; ++k) for (int j = 0; j < 64; ++j) w.w[k][j] = hW[k * 64 + j];
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = (32 * 64 + 32 * PITCH) * sizeof(float);
  cudaFuncSetAttribute(k_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double flop = 2.0 * M * 32 * 64 * REP;
  std::vector<float> r0(M), r1(M);
  const size_t smem2 = (2 * 32 * 64 + 32 * PITCH) * sizeof(float);
  cudaFuncSetAttribute(k_smem2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  cudaFuncSetAttribute(k_smem2c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  const size_t smemh = (32 * PITCH) * sizeof(float);
  const size_t smemh1 = (32 * PITCH + 32 * 64) * sizeof(float);
  for (int variant = 0; variant < 10; ++variant) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      if (variant == 0) k_uniform<<<2 * sms, 256>>>(w, dH, dD, M);
      if (variant == 1) k_smem<8><<<2 * sms, 256, smem>>>(dW, dH, dD, M);
      if (variant == 2) k_smem<4><<<2 * sms, 256, smem>>>(dW, dH, dD, M);
      if (variant == 3) k_uniform2<<<2 * sms, 256>>>(w, dH, dD, M);
      if (variant == 4) k_smem2<<<2 * sms, 256, smem2>>>(dW, dH, dD, M);
      if (variant == 5) k_smem2c<<<2 * sms, 256, smem2>>>(dW, dH, dD, M);
      if (variant == 6) k_hybrid<false><<<2 * sms, 256, smemh>>>(w, dH, dD, M);
      if (variant == 7) k_hybrid<true><<<2 * sms, 256, smemh>>>(w, dH, dD, M);
      if (variant == 8) k_hybrid1<0><<<2 * sms, 256, smemh1>>>(w, dW, dH, dD, M);
      if (variant == 9) k_hybrid1<1><<<2 * sms, 256, smemh1>>>(w, dW, dH, dD, M);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaMemcpy(variant == 0 ? r0.data() : r1.data(), dD, (size_t)M * 4, cudaMemcpyDeviceToHost);
    double maxdiff = 0; if (variant) for (int i = 0; i < M; ++i) { double d = fabs((double)r0[i] - r1[i]); if (d > maxdiff) maxdiff = d; }
    const char* names[10] = {"U  thread-per-row, uniform-register weights", "S  smem 8x8 register tile", "S4 smem 8x4 register tile x2", "U2 thread-per-row, uniform weights, FFMA2", "S2 smem 8x8, FFMA2 row pairs, dup W", "S2c smem 8x8, FFMA2 col pairs", "H  warp-uniform weights (UR), rows on lanes, LDS.32", "Hv warp-uniform weights (UR), rows on lanes, LDS.128", "Hc one body, weights by indexed constant load", "Hs one body, weights by smem broadcast LDS.128"};
    printf("%-46s %8.3f ms  %7.2f TFLOP/s  (H read %.0f GB/s)  maxdiff_vs_U %.3g  %s\n", names[variant], best,
           flop / best / 1e9, (double)M * 128 / best / 1e6, maxdiff, cudaGetErrorString(err));
  }
  return 0;
}
