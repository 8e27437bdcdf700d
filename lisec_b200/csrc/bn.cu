// Training-mode BatchNormalization for channels-last activations — forward statistics / apply and the backward pass
// (DESIGN.md §4e). Reference: every BatchNormalization() of createModel runs in training mode under model.fit
// (model_training.py:171, 194, 204, 299): y = gamma * (x - mean_B) / sqrt(var_B + 1e-3) + beta with the batch mean and the
// biased batch variance over all positions, moving statistics updated with momentum 0.99.
//
// All four kernels stream [P positions][C channels] bf16 tensors once: HBM-bound (2 B read per element for the
// statistics, 2 + 2 for apply, 6 + 2 for the backward pair). Reductions are deterministic: every block leaves its
// partial sums (double) in its own slot, a finalize kernel adds the slots in order — no floating-point atomics.
//   bn_stats_kernel        per block: sum x, sum x^2 per channel
//   bn_finalize_kernel     mean, biased variance -> a = gamma / sqrt(var + eps), b = beta - mean * a; moving statistics
//   bn_apply_kernel        y = x * a + b (optional ReLU), bf16
//   bn_bwd_reduce_kernel   g = dy * (y > 0 if relu): per block sum g, sum g * xhat per channel
//   bn_bwd_finalize_kernel dgamma = sum g xhat, dbeta = sum g; coefficients of the apply
//   bn_bwd_apply_kernel    dx = gamma * invstd * (g - mean(g) - xhat * mean(g xhat)), bf16
#include <cuda_bf16.h>

#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace lisec {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxC = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// the gradient tensor arriving at a BatchNormalization may be float32: its per-channel common mode is large against what
// the backward pass keeps, so a bf16 copy of it loses the part that matters
template <bool F32>
__device__ __forceinline__ void load_dy8(const void* dy, long long idx8, float (&g)[8]) {
  if (F32) {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(dy) + 2 * idx8), b = __ldcg(reinterpret_cast<const float4*>(dy) + 2 * idx8 + 1);
    g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
  } else {
    unpack8(__ldcg(reinterpret_cast<const uint4*>(dy) + idx8), g);
  }
}

// Two per-channel sums over the rows a block owns. Thread = (row lane, 8-channel chunk); a block's rows are
// blockIdx.x, blockIdx.x + gridDim.x, ... in groups of rows_per_iter. MODE 0: (x, x^2). MODE 1: (g, g * xhat) with
// g = dy masked by y > 0 (relu) and xhat = (x - mean) * invstd.
template <int MODE, bool DYF32 = false>
__global__ void __launch_bounds__(kBnThreads)
    bn_reduce_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ dy,
                     const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ invstd,
                     long long P, int C, int relu, double* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = C >> 3, rows_per_iter = kBnThreads / chunks;
  const int chunk = threadIdx.x % chunks, rl = threadIdx.x / chunks;
  // (MODE 1) mean / invstd were written by the kernels in front: through L2 into shared memory once per block, not 16
  // L2 loads per thread
  __shared__ float s_mean[MODE == 1 ? kBnMaxC : 1], s_invstd[MODE == 1 ? kBnMaxC : 1];
  if (MODE == 1) {
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
      s_mean[c] = __ldcg(mean + c);
      s_invstd[c] = __ldcg(invstd + c);
    }
    __syncthreads();
  }
  float s0[8], s1[8], mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s0[i] = s1[i] = 0.f;
    mu[i] = MODE == 1 ? s_mean[8 * chunk + i] : 0.f;
    is[i] = MODE == 1 ? s_invstd[8 * chunk + i] : 0.f;
  }
  double d0[8], d1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) d0[i] = d1[i] = 0.0;
  int since = 0;
  if (rl < rows_per_iter)
    for (long long r = (long long)blockIdx.x * rows_per_iter + rl; r < P; r += (long long)gridDim.x * rows_per_iter) {
      float xv[8];
      unpack8(*reinterpret_cast<const uint4*>(x + r * C + 8 * chunk), xv);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += xv[i];
          s1[i] = fmaf(xv[i], xv[i], s1[i]);
        }
      } else {
        float g[8], yv[8];
        load_dy8<DYF32>(dy, (r * C + 8 * chunk) >> 3, g);
        if (relu) unpack8(*reinterpret_cast<const uint4*>(y + r * C + 8 * chunk), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float gi = (relu && !(yv[i] > 0.f)) ? 0.f : g[i];
          s0[i] += gi;
          s1[i] = fmaf(gi, (xv[i] - mu[i]) * is[i], s1[i]);
        }
      }
      if (++since == 64) {  // float partial sums are promoted every 64 rows
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          d0[i] += (double)s0[i];
          d1[i] += (double)s1[i];
          s0[i] = s1[i] = 0.f;
        }
        since = 0;
      }
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    d0[i] += (double)s0[i];
    d1[i] += (double)s1[i];
  }
  __shared__ double sh[kBnThreads][2];
  for (int i = 0; i < 8; ++i) {
    sh[threadIdx.x][0] = d0[i];
    sh[threadIdx.x][1] = d1[i];
    __syncthreads();
    if (rl == 0 && threadIdx.x < chunks) {  // fixed order over the block's row lanes
      double a = 0.0, b = 0.0;
      for (int q = 0; q < rows_per_iter; ++q) {
        a += sh[q * chunks + chunk][0];
        b += sh[q * chunks + chunk][1];
      }
      partial[((long long)blockIdx.x * 2 + 0) * C + 8 * chunk + i] = a;
      partial[((long long)blockIdx.x * 2 + 1) * C + 8 * chunk + i] = b;
    }
    __syncthreads();
  }
}

// MODE 0: statistics -> (mean, invstd, a, b) and the moving statistics. MODE 1: (dgamma, dbeta, c_mean_g, c_mean_gx).
// Grid = C / 8 blocks of kBnFinThreads threads; a block owns 8 channels (64 contiguous bytes of every slot) and its
// thread (slice s, channel c) adds the partials of blocks s, s + S, s + 2S, ... (S = 128 slices, four independent running
// sums combined in a fixed order, so that the loads overlap instead of queueing behind one dependent chain of float64
// adds), then thread (0, c) adds the S slice sums in order. One thread per channel walking all ~1200 slots — the first
// version — took 100+ us per call, 40 % of a training step. The summation order is fixed: results are bit-reproducible.
constexpr int kBnFinThreads = 1024;
constexpr int kBnFinChannels = 8;
template <int MODE>
__global__ void __launch_bounds__(kBnFinThreads)
    bn_finalize_kernel(const double* __restrict__ partial, int blocks, long long P, int C, float eps, float momentum,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ o0,
                       float* __restrict__ o1, float* __restrict__ o2, float* __restrict__ o3,
                       float* __restrict__ moving_mean, float* __restrict__ moving_var, int unbiased_moving) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double sa[kBnFinThreads], sb[kBnFinThreads];
  constexpr int S = kBnFinThreads / kBnFinChannels;
  const int lc = threadIdx.x % kBnFinChannels, sl = threadIdx.x / kBnFinChannels;
  const int c = blockIdx.x * kBnFinChannels + lc;
  double a = 0.0, b = 0.0;
  {
    double a4[4] = {0.0, 0.0, 0.0, 0.0}, b4[4] = {0.0, 0.0, 0.0, 0.0};
    int g = sl;
    for (; g + 3 * S < blocks; g += 4 * S) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a4[u] += __ldcg(partial + ((long long)(g + u * S) * 2 + 0) * C + c);
        b4[u] += __ldcg(partial + ((long long)(g + u * S) * 2 + 1) * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u)  // at most three slots are left
      if (g + u * S < blocks) {
        a4[u] += __ldcg(partial + ((long long)(g + u * S) * 2 + 0) * C + c);
        b4[u] += __ldcg(partial + ((long long)(g + u * S) * 2 + 1) * C + c);
      }
    a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    b = (b4[0] + b4[1]) + (b4[2] + b4[3]);
  }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  if (sl != 0) return;
  for (int k = 1; k < S; ++k) {
    a += sa[k * kBnFinChannels + lc];
    b += sb[k * kBnFinChannels + lc];
  }
  if (MODE == 0) {
    const double mean = a / (double)P, var = fmax(b / (double)P - mean * mean, 0.0);
    const double invstd = 1.0 / sqrt(var + (double)eps);
    o0[c] = (float)mean;
    o1[c] = (float)invstd;
    const double sc = (double)gamma[c] * invstd;
    o2[c] = (float)sc;
    o3[c] = (float)((double)beta[c] - mean * sc);
    if (moving_mean) {
      moving_mean[c] = (float)((double)moving_mean[c] * momentum + mean * (1.0 - (double)momentum));
      // Keras's fused BatchNormalization (rank-4 inputs) feeds the Bessel-corrected variance to the moving average
      const double mv = (unbiased_moving && P > 1) ? var * (double)P / (double)(P - 1) : var;
      moving_var[c] = (float)((double)moving_var[c] * momentum + mv * (1.0 - (double)momentum));
    }
  } else if (MODE == 2) {
    o0[c] = (float)a;  // plain column sums: a convolution's bias gradient = sum over positions of dy
  } else {
    o0[c] = (float)b;                 // dgamma = sum g * xhat
    o1[c] = (float)a;                 // dbeta  = sum g
    o2[c] = (float)(a / (double)P);   // mean(g)
    o3[c] = (float)(b / (double)P);   // mean(g * xhat)
  }
}

__global__ void __launch_bounds__(kBnThreads)
    bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                    long long n8, int chunks, int relu, __nv_bfloat16* __restrict__ y) {
  pdl_launch_dependents();
  pdl_wait();
  // scale / shift were written by the finalize kernel in front of this one: fetched through L2 once per block into shared
  // memory (programmatic dependent launch: see common.cuh), read from there per element. (Holding them in registers
  // instead cost the occupancy this memory-bound kernel lives on: 4x slower.)
  __shared__ float s_a[kBnMaxC], s_b[kBnMaxC];
  for (int c = threadIdx.x; c < chunks * 8; c += blockDim.x) {
    s_a[c] = __ldcg(a + c);
    s_b[c] = __ldcg(b + c);
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int c0 = (int)(i % chunks) * 8;
    float v[8];
    unpack8(__ldcg(reinterpret_cast<const uint4*>(x) + i), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = fmaf(v[k], s_a[c0 + k], s_b[c0 + k]);
      if (relu) v[k] = fmaxf(v[k], 0.f);
    }
    reinterpret_cast<uint4*>(y)[i] = pack8(v);
  }
}

template <bool DYF32>
__global__ void __launch_bounds__(kBnThreads)
    bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ dy,
                        const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ mean_g, const float* __restrict__ mean_gx, long long n8, int chunks,
                        int relu, __nv_bfloat16* __restrict__ dx) {
  pdl_launch_dependents();
  pdl_wait();
  // the per-channel coefficients of this step (written by the kernels in front): through L2 into shared memory, once
  __shared__ float s_mu[kBnMaxC], s_is[kBnMaxC], s_gs[kBnMaxC], s_mg[kBnMaxC], s_mgx[kBnMaxC];
  for (int c = threadIdx.x; c < chunks * 8; c += blockDim.x) {
    const float is = __ldcg(invstd + c);
    s_mu[c] = __ldcg(mean + c);
    s_is[c] = is;
    s_gs[c] = __ldcg(gamma + c) * is;
    s_mg[c] = __ldcg(mean_g + c);
    s_mgx[c] = __ldcg(mean_gx + c);
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int c0 = (int)(i % chunks) * 8;
    float xv[8], g[8], yv[8], o[8];
    unpack8(__ldcg(reinterpret_cast<const uint4*>(x) + i), xv);
    load_dy8<DYF32>(dy, i, g);
    if (relu) unpack8(__ldcg(reinterpret_cast<const uint4*>(y) + i), yv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      const float gi = (relu && !(yv[k] > 0.f)) ? 0.f : g[k];
      const float xhat = (xv[k] - s_mu[c]) * s_is[c];
      o[k] = s_gs[c] * (gi - s_mg[c] - xhat * s_mgx[c]);
    }
    reinterpret_cast<uint4*>(dx)[i] = pack8(o);
  }
}

thread_local char g_bn_error[256] = "";
int32_t bn_fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_bn_error, sizeof(g_bn_error), fmt, ap);
  va_end(ap);
  return code;
}

int bn_blocks(long long P, int C) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int rows_per_iter = kBnThreads / (C / 8);
  long long b = (P + rows_per_iter - 1) / rows_per_iter;
  if (b > (long long)sms * 8) b = (long long)sms * 8;
  return (int)(b < 1 ? 1 : b);
}

bool bn_shape_ok(long long P, int C) { return P > 0 && C >= 8 && C <= kBnMaxC && C % 8 == 0 && kBnThreads % (C / 8) == 0; }

}  // namespace
}  // namespace lisec

using namespace lisec;

extern "C" {

const char* lisec_bn_last_error(void) { return g_bn_error; }

int64_t lisec_bn_workspace_bytes(int64_t positions, int32_t channels) {
  if (!bn_shape_ok(positions, channels)) return -1;
  return (int64_t)bn_blocks(positions, channels) * 2 * channels * (int64_t)sizeof(double);
}

int32_t lisec_bn_train_forward(const void* x, int64_t positions, int32_t channels, const float* gamma, const float* beta,
                               float eps, float momentum, float* moving_mean, float* moving_var, int32_t relu, void* y,
                               float* mean, float* invstd, float* scale, float* shift, void* workspace, void* stream) {
  if (!x || !gamma || !beta || !y || !mean || !invstd || !scale || !shift || !workspace)
    return bn_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (!bn_shape_ok(positions, channels))
    return bn_fail(LISEC_ERR_BAD_CONFIG, "channels: a multiple of 8 dividing 2048, at most %d", kBnMaxC);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = bn_blocks(positions, channels);
  double* part = static_cast<double*>(workspace);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  cudaError_t e = launch_pdl(bn_reduce_kernel<0>, dim3(blocks), dim3(kBnThreads), 0, st, xb, xb, xb, (const float*)mean,
                             (const float*)invstd, (long long)positions, (int)channels, 0, part);
  if (e == cudaSuccess)
    e = launch_pdl(bn_finalize_kernel<0>, dim3(channels / kBnFinChannels), dim3(kBnFinThreads), 0, st, (const double*)part, blocks, (long long)positions,
                   (int)channels, eps, momentum, gamma, beta, mean, invstd, scale, shift, moving_mean, moving_var,
                   (int)((relu >> 1) & 1));
  const long long n8 = positions * channels / 8;
  long long ab = (n8 + kBnThreads - 1) / kBnThreads;
  if (ab > 148 * 16) ab = 148 * 16;
  if (e == cudaSuccess)
    e = launch_pdl(bn_apply_kernel, dim3((unsigned)ab), dim3(kBnThreads), 0, st, xb, (const float*)scale,
                   (const float*)shift, n8, (int)(channels / 8), (int)(relu & 1), static_cast<__nv_bfloat16*>(y));
  if (e != cudaSuccess) return bn_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_channel_sums(const void* x, int64_t positions, int32_t channels, float* sums, void* workspace, void* stream) {
  if (!x || !sums || !workspace) return bn_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (!bn_shape_ok(positions, channels))
    return bn_fail(LISEC_ERR_BAD_CONFIG, "channels: a multiple of 8 dividing 2048, at most %d", kBnMaxC);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = bn_blocks(positions, channels);
  double* part = static_cast<double*>(workspace);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  cudaError_t e = launch_pdl(bn_reduce_kernel<0>, dim3(blocks), dim3(kBnThreads), 0, st, xb, xb, xb, (const float*)sums,
                             (const float*)sums, (long long)positions, (int)channels, 0, part);
  if (e == cudaSuccess)
    e = launch_pdl(bn_finalize_kernel<2>, dim3(channels / kBnFinChannels), dim3(kBnFinThreads), 0, st, (const double*)part, blocks, (long long)positions,
                   (int)channels, 0.f, 0.f, (const float*)sums, (const float*)sums, sums, sums, sums, sums,
                   (float*)nullptr, (float*)nullptr, 0);
  if (e != cudaSuccess) return bn_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

}  // extern "C"

static int32_t bn_backward_impl(const void* x, const void* dy, int dy_f32, const void* y, int64_t positions, int32_t channels,
                                const float* gamma, const float* mean, const float* invstd, int32_t relu, void* dx,
                                float* dgamma, float* dbeta, float* mean_g, float* mean_gx, void* workspace,
                                void* stream) {
  if (!x || !dy || (relu && !y) || !gamma || !mean || !invstd || !dx || !dgamma || !dbeta || !mean_g || !mean_gx || !workspace)
    return bn_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (!bn_shape_ok(positions, channels))
    return bn_fail(LISEC_ERR_BAD_CONFIG, "channels: a multiple of 8 dividing 2048, at most %d", kBnMaxC);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = bn_blocks(positions, channels);
  double* part = static_cast<double*>(workspace);
  const __nv_bfloat16 *xb = static_cast<const __nv_bfloat16*>(x), *yb = static_cast<const __nv_bfloat16*>(relu ? y : x);
  const void* dyb = dy;
  cudaError_t e = dy_f32 ? launch_pdl(bn_reduce_kernel<1, true>, dim3(blocks), dim3(kBnThreads), 0, st, xb, dyb, yb, mean, invstd,
                                      (long long)positions, (int)channels, (int)relu, part)
                         : launch_pdl(bn_reduce_kernel<1, false>, dim3(blocks), dim3(kBnThreads), 0, st, xb, dyb, yb, mean, invstd,
                                      (long long)positions, (int)channels, (int)relu, part);
  if (e == cudaSuccess)
    e = launch_pdl(bn_finalize_kernel<1>, dim3(channels / kBnFinChannels), dim3(kBnFinThreads), 0, st, (const double*)part, blocks, (long long)positions,
                   (int)channels, 0.f, 0.f, gamma, gamma, dgamma, dbeta, mean_g, mean_gx, (float*)nullptr, (float*)nullptr, 0);
  const long long n8 = positions * channels / 8;
  long long ab = (n8 + kBnThreads - 1) / kBnThreads;
  if (ab > 148 * 16) ab = 148 * 16;
  if (e == cudaSuccess)
    e = dy_f32 ? launch_pdl(bn_bwd_apply_kernel<true>, dim3((unsigned)ab), dim3(kBnThreads), 0, st, xb, dyb, yb, mean, invstd,
                            gamma, (const float*)mean_g, (const float*)mean_gx, n8, (int)(channels / 8), (int)relu,
                            static_cast<__nv_bfloat16*>(dx))
               : launch_pdl(bn_bwd_apply_kernel<false>, dim3((unsigned)ab), dim3(kBnThreads), 0, st, xb, dyb, yb, mean, invstd,
                            gamma, (const float*)mean_g, (const float*)mean_gx, n8, (int)(channels / 8), (int)relu,
                            static_cast<__nv_bfloat16*>(dx));
  if (e != cudaSuccess) return bn_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}


extern "C" {

int32_t lisec_bn_train_backward(const void* x, const void* dy, const void* y, int64_t positions, int32_t channels,
                                const float* gamma, const float* mean, const float* invstd, int32_t relu, void* dx,
                                float* dgamma, float* dbeta, float* mean_g, float* mean_gx, void* workspace,
                                void* stream) {
  return bn_backward_impl(x, dy, 0, y, positions, channels, gamma, mean, invstd, relu, dx, dgamma, dbeta, mean_g, mean_gx,
                          workspace, stream);
}

int32_t lisec_bn_train_backward_f32(const void* x, const float* dy, const void* y, int64_t positions, int32_t channels,
                                    const float* gamma, const float* mean, const float* invstd, int32_t relu, void* dx,
                                    float* dgamma, float* dbeta, float* mean_g, float* mean_gx, void* workspace,
                                    void* stream) {
  return bn_backward_impl(x, dy, 1, y, positions, channels, gamma, mean, invstd, relu, dx, dgamma, dbeta, mean_g, mean_gx,
                          workspace, stream);
}

}  // extern "C"
