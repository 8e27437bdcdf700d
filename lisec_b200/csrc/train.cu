// Training-step pieces that exist so far: the optimizer update and the loss head — NOT the
// backward pass (lisec_b200/compat.py: train() says so). Reference model_training.py:295-299:
//     sgd = optimizers.SGD(lr=0.01, decay=1e-6, momentum=0.9, nesterov=True);  model.compile(optimizer=sgd, loss=['mse', 'mse'])
//
// sgd_nesterov_kernel: one Keras SGD update (optimizer_v2, resource_apply_keras_momentum) over the FLAT parameter buffer
// (6 491 024 float32 for this model) after the gradient all-reduce:
//     g = grad * grad_scale          (grad_scale = 1 / world_size: the all-reduce sums)
//     step = lr_t * g;  accum = accum * momentum - step;  var = var + (accum * momentum - step)   [nesterov]
// evaluated operation by operation with round-to-nearest float32 intrinsics (no FMA contraction), i.e. bit for bit what
// oracle/train_oracle.py: sgd_nesterov_update computes in numpy float32. HBM-bound: 12 B read + 8 B written per parameter
// (130 MB per step: ~20 us at the measured copy bandwidth); 16-byte accesses, grid-stride over 148 x 8 CTAs.
//
// mse_loss_grad_kernel: loss=['mse','mse'] on one output: sum of (y - t)^2 into a double accumulator (scaled by 1/n on the
// host side of the ABI) and d loss / d y = 2 (y - t) / n, the tensor the backward pass will start from.
#include <cstdarg>
#include <cstdio>

#include <cuda_bf16.h>

#include "common.cuh"

namespace lisec {

namespace {

__device__ __forceinline__ float sgd_one(float& var, float& acc, float g, float scale, float lr_t, float mom, int nesterov) {
  const float step = __fmul_rn(lr_t, __fmul_rn(g, scale));
  acc = __fsub_rn(__fmul_rn(acc, mom), step);
  var = nesterov ? __fadd_rn(var, __fsub_rn(__fmul_rn(acc, mom), step)) : __fadd_rn(var, acc);
  return var;
}

__global__ void __launch_bounds__(256)
    sgd_nesterov_kernel(float* __restrict__ var, float* __restrict__ accum, const float* __restrict__ grad, long long n,
                        float grad_scale, float lr_t, float momentum, int nesterov) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(var)[i];
    float4 a = reinterpret_cast<float4*>(accum)[i];
    const float4 g = __ldcg(reinterpret_cast<const float4*>(grad) + i);
    sgd_one(v.x, a.x, g.x, grad_scale, lr_t, momentum, nesterov);
    sgd_one(v.y, a.y, g.y, grad_scale, lr_t, momentum, nesterov);
    sgd_one(v.z, a.z, g.z, grad_scale, lr_t, momentum, nesterov);
    sgd_one(v.w, a.w, g.w, grad_scale, lr_t, momentum, nesterov);
    reinterpret_cast<float4*>(var)[i] = v;
    reinterpret_cast<float4*>(accum)[i] = a;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = var[i], a = accum[i];
    sgd_one(v, a, grad[i], grad_scale, lr_t, momentum, nesterov);
    var[i] = v;
    accum[i] = a;
  }
}

__global__ void __launch_bounds__(256)
    mse_loss_grad_kernel(const float* __restrict__ y, const float* __restrict__ t, long long n, float* __restrict__ dy,
                         double* __restrict__ sum_sq) {
  pdl_launch_dependents();
  pdl_wait();
  const float two_over_n = (float)(2.0 / (double)n);
  double local = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = __fsub_rn(y[i], t[i]);
    local += (double)d * (double)d;
    if (dy) dy[i] = __fmul_rn(d, two_over_n);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += s[w];
    atomicAdd(sum_sq, tot);
  }
}

// dgrad operand: the data gradient of a stride-1 convolution is a convolution of dY with the kernel flipped in every
// spatial direction and its channel roles swapped: out[(kd-1-a, kh-1-b, kw-1-c)][ci][co] = w[(a, b, c)][co][ci]
// (both in the forward plans' [tap][N][C] layout, float32 master weights in, bf16 operand out).
__global__ void __launch_bounds__(256)
    flip_transpose_kernel(const float* __restrict__ w, int kd, int kh, int kw, int n_out, int c_in,
                          __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)kd * kh * kw * n_out * c_in;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the OUTPUT [tap'][ci][co] so that the writes are coalesced
  const int co = (int)(i % n_out);
  const int ci = (int)((i / n_out) % c_in);
  const int tp = (int)(i / ((long long)n_out * c_in));
  const int c = tp % kw, b = (tp / kw) % kh, a = tp / (kw * kh);
  const int tap = ((kd - 1 - a) * kh + (kh - 1 - b)) * kw + (kw - 1 - c);
  out[i] = __float2bfloat16(w[((long long)tap * n_out + co) * c_in + ci]);
}

// Every operand refresh of a training step in ONE launch: a table of (float32 source, bf16 destination, first element,
// shape, mode) entries — mode 0 a plain cast, mode 1 the flip + transpose above — walked by element index. After an
// optimizer step every layer's operand copies are re-derived from its master weights: ~90 launches of 2-5 us otherwise.
constexpr int kRefreshMax = 256;
__global__ void __launch_bounds__(256) refresh_operands_kernel(const lisec_refresh_entry* __restrict__ tab, int n, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ long long s_first[kRefreshMax + 1];
  for (int e = threadIdx.x; e < n; e += blockDim.x) s_first[e] = tab[e].first;
  if (threadIdx.x == 0) s_first[n] = total;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int lo = 0, hi = n;  // the entry with first[e] <= i < first[e + 1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_first[mid] <= i) lo = mid; else hi = mid;
    }
    const lisec_refresh_entry& t = tab[lo];
    const long long j = i - s_first[lo];
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(t.dst);
    if (t.mode == 0) {
      out[j] = __float2bfloat16(t.src[j]);
    } else {  // j indexes the OUTPUT [tap'][ci][co]
      const int co = (int)(j % t.out_c);
      const int ci = (int)((j / t.out_c) % t.in_c);
      const int tp = (int)(j / ((long long)t.out_c * t.in_c));
      const int c = tp % t.kw, b = (tp / t.kw) % t.kh, a = tp / (t.kw * t.kh);
      const int tap = ((t.kd - 1 - a) * t.kh + (t.kh - 1 - b)) * t.kw + (t.kw - 1 - c);
      out[j] = __float2bfloat16(t.src[((long long)tap * t.out_c + co) * t.in_c + ci]);
    }
  }
}

// float32 master weights -> the bf16 operand copy the plans read (after every optimizer step)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ w, long long n, __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __float2bfloat16(w[i]);
}

// ReLU backward for stages whose ReLU does not sit behind a BatchNormalization (the Dense of a Conv3D block):
// out = dy where y > 0, else 0 (bf16, 8 elements per thread)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, long long n8,
                                                       uint4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 g = __ldcg(dy + i);
    const uint4 v = __ldcg(y + i);
    const __nv_bfloat16* yv = reinterpret_cast<const __nv_bfloat16*>(&v);
    __nv_bfloat16* gv = reinterpret_cast<__nv_bfloat16*>(&g);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (!(__bfloat162float(yv[k]) > 0.f)) gv[k] = __float2bfloat16(0.f);
    out[i] = g;
  }
}

// the same from a float32 gradient (the gradient tensors between stages are float32): out bf16
__global__ void __launch_bounds__(256) relu_bwd_f32_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                           long long n, __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16(__bfloat162float(y[i]) > 0.f ? dy[i] : 0.f);
}

// Zero-dilation of a gradient tensor: the data gradient of a STRIDED convolution is the stride-1 data gradient of dy with
// (stride - 1) zeros between its positions (and k - 1 - pad zeros around it, which the plan's own padding provides).
// out [B, (D-1)*sd+1 + ed, (H-1)*s+1 + eh, (W-1)*s+1 + ew, C] is cleared by the caller once; the zeros never change.
__global__ void __launch_bounds__(256)
    dilate_kernel(const uint4* __restrict__ in, int B, int D, int H, int W, int c8, int sd, int s, int OD, int OH, int OW,
                  uint4* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * D * H * W * c8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % c8);
    long long r = i / c8;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    r /= H;
    const int d = (int)(r % D);
    const int b = (int)(r / D);
    out[((((long long)b * OD + (long long)d * sd) * OH + (long long)h * s) * OW + (long long)w * s) * c8 + c] = __ldcg(in + i);
  }
}

// float32 [P][c_in] -> bf16 [P][c_out] with zero padding of the channels (c_out >= c_in): the heads' 16-column gradient
// widened to the 64 channels the tensor-core operands are made of.
__global__ void __launch_bounds__(256)
    pad_channels_kernel(const float* __restrict__ in, long long P, int c_in, int c_out, __nv_bfloat16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = P * c_out, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % c_out);
    out[i] = __float2bfloat16(c < c_in ? in[(i / c_out) * c_in + c] : 0.f);
  }
}

// a += b (bf16, float32 add, 8 elements per thread): the gradients of a tensor with two consumers (an RPN block's output
// feeds the next block and its transposed convolution)
__global__ void __launch_bounds__(256) add_bf16_kernel(uint4* __restrict__ a, const uint4* __restrict__ b, long long n8) {
  pdl_launch_dependents();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 x = a[i];
    const uint4 y = __ldcg(b + i);
    __nv_bfloat162* xv = reinterpret_cast<__nv_bfloat162*>(&x);
    const __nv_bfloat162* yv = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = __bfloat1622float2(xv[k]), q = __bfloat1622float2(yv[k]);
      xv[k] = __floats2bfloat162_rn(p.x + q.x, p.y + q.y);
    }
    a[i] = x;
  }
}

thread_local char g_train_error[256] = "";

int32_t train_fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_train_error, sizeof(g_train_error), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace

}  // namespace lisec

using namespace lisec;

extern "C" {

const char* lisec_train_last_error(void) { return g_train_error; }

int32_t lisec_sgd_nesterov(float* var, float* accum, const float* grad, int64_t n, float grad_scale, float lr_t,
                           float momentum, int32_t nesterov, void* stream) {
  if (n < 0) return train_fail(LISEC_ERR_BAD_ARG, "negative size");
  if (n == 0) return LISEC_OK;
  if (!var || !accum || !grad) return train_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (((uintptr_t)var | (uintptr_t)accum | (uintptr_t)grad) & 15)
    return train_fail(LISEC_ERR_BAD_ARG, "buffers must be 16-byte aligned");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  if (blocks < 1) blocks = 1;
  cudaError_t e = launch_pdl(sgd_nesterov_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             var, accum, grad, (long long)n, grad_scale, lr_t, momentum, (int)nesterov);
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_weights_flip_transpose(const float* w, int32_t kd, int32_t kh, int32_t kw, int32_t out_c, int32_t in_c,
                                     void* out_bf16, void* stream) {
  if (!w || !out_bf16 || kd < 1 || kh < 1 || kw < 1 || out_c < 1 || in_c < 1) return train_fail(LISEC_ERR_BAD_ARG, "bad argument");
  const long long total = (long long)kd * kh * kw * out_c * in_c;
  cudaError_t e = launch_pdl(flip_transpose_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                             static_cast<cudaStream_t>(stream), w, (int)kd, (int)kh, (int)kw, (int)out_c, (int)in_c,
                             static_cast<__nv_bfloat16*>(out_bf16));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_refresh_operands(const lisec_refresh_entry* entries, int32_t n, int64_t total, void* stream) {
  if (!entries || n < 1 || n > kRefreshMax || total < 1) return train_fail(LISEC_ERR_BAD_ARG, "1..%d entries", kRefreshMax);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaError_t e = launch_pdl(refresh_operands_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             entries, (int)n, (long long)total);
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_relu_backward(const void* dy, const void* y, int64_t n, void* out, void* stream) {
  if (n < 0 || n % 8 || (n > 0 && (!dy || !y || !out))) return train_fail(LISEC_ERR_BAD_ARG, "n must be a multiple of 8");
  if (n == 0) return LISEC_OK;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaError_t e = launch_pdl(relu_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             static_cast<const uint4*>(dy), static_cast<const uint4*>(y), (long long)(n / 8),
                             static_cast<uint4*>(out));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_relu_backward_f32(const float* dy, const void* y, int64_t n, void* out, void* stream) {
  if (n < 0 || (n > 0 && (!dy || !y || !out))) return train_fail(LISEC_ERR_BAD_ARG, "bad argument");
  if (n == 0) return LISEC_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaError_t e = launch_pdl(relu_bwd_f32_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), dy,
                             static_cast<const __nv_bfloat16*>(y), (long long)n, static_cast<__nv_bfloat16*>(out));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_dilate(const void* in, int32_t batch, int32_t d, int32_t h, int32_t w, int32_t channels, int32_t stride_d,
                     int32_t stride_hw, int32_t out_d, int32_t out_h, int32_t out_w, void* out, void* stream) {
  if (!in || !out || channels % 8 || batch < 1 || d < 1 || h < 1 || w < 1 || stride_d < 1 || stride_hw < 1 ||
      out_d < (d - 1) * stride_d + 1 || out_h < (h - 1) * stride_hw + 1 || out_w < (w - 1) * stride_hw + 1)
    return train_fail(LISEC_ERR_BAD_ARG, "dilate: bad shape");
  const long long total = (long long)batch * d * h * w * (channels / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaError_t e = launch_pdl(dilate_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             static_cast<const uint4*>(in), (int)batch, (int)d, (int)h, (int)w, (int)(channels / 8),
                             (int)stride_d, (int)stride_hw, (int)out_d, (int)out_h, (int)out_w, static_cast<uint4*>(out));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_pad_channels_bf16(const float* in, int64_t positions, int32_t c_in, int32_t c_out, void* out_bf16, void* stream) {
  if (!in || !out_bf16 || positions < 1 || c_in < 1 || c_out < c_in) return train_fail(LISEC_ERR_BAD_ARG, "bad argument");
  long long blocks = (positions * c_out + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaError_t e = launch_pdl(pad_channels_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), in,
                             (long long)positions, (int)c_in, (int)c_out, static_cast<__nv_bfloat16*>(out_bf16));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_add_bf16(void* a, const void* b, int64_t n, void* stream) {
  if (n < 0 || n % 8 || (n > 0 && (!a || !b))) return train_fail(LISEC_ERR_BAD_ARG, "n must be a multiple of 8");
  if (n == 0) return LISEC_OK;
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaError_t e = launch_pdl(add_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             static_cast<uint4*>(a), static_cast<const uint4*>(b), (long long)(n / 8));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_cast_f32_to_bf16(const float* w, int64_t n, void* out_bf16, void* stream) {
  if (n < 0 || (n > 0 && (!w || !out_bf16))) return train_fail(LISEC_ERR_BAD_ARG, "bad argument");
  if (n == 0) return LISEC_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaError_t e = launch_pdl(cast_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), w,
                             (long long)n, static_cast<__nv_bfloat16*>(out_bf16));
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_mse_loss_grad(const float* y, const float* target, int64_t n, float* dy, double* sum_sq, void* stream) {
  if (n <= 0 || !y || !target || !sum_sq) return train_fail(LISEC_ERR_BAD_ARG, "null argument or empty tensor");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  cudaError_t e = launch_pdl(mse_loss_grad_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream),
                             y, target, (long long)n, dy, sum_sq);
  if (e != cudaSuccess) return train_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

}  // extern "C"
