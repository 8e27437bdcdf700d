// The VFE stack in TRAINING mode (model_training.py:229-235 under fit(), :299): forward with batch statistics and the
// backward pass, on ROWS WITH MULTIPLICITIES — the exact restatement of the dense graph that oracle/train_oracle.py:
// forward_train_rows() states and tests. The dense input holds three classes of rows:
//   kept points                      weight 1
//   one virtual pad row per non-full voxel, standing for its T - s identical zero rows      weight T - s
//   ONE empty row standing for the T * n_empty rows of all empty voxels                     weight T * n_empty
// Every BatchNormalization statistic is the weighted sum over these rows divided by M = n_cells * T; the per-voxel max
// runs over a voxel's rows; its gradient is split equally among tied COPIES (TensorFlow's reduce_max rule), so compact
// row r receives w_r * g / n_ties with n_ties = sum of the weights of the tied rows.
//
// The rows are the voxelizer's own VFE rows (row_start / row_voxel / row_xyz of the last lisec_voxelize(): kept rows and
// the pad row of every voxel, contiguous), plus one more row for the empty voxels at index n_rows. Everything is O(rows):
// ~115 k rows x 112 channels per 100 k-point sweep; one thread per row (or per voxel x 4 channels), float32 arithmetic,
// every reduction in float64 with a fixed summation order (bit-reproducible, no atomics).
#include <cuda_bf16.h>

#include <new>

#include "handle.cuh"
#include "vfe_math.cuh"

namespace lisec {

constexpr int kTrBlocks = 296;      // persistent blocks of the reduction kernels (2 per SM)
constexpr int kTrThreads = 256;

struct VfeTrainState {
  long long max_rows = 0, max_vox = 0;
  float* x0 = nullptr;        // [rows][6] input features
  float* w = nullptr;         // [rows] copies per compact row
  int* seg = nullptr;         // [rows] voxel of the row (n_voxels for the empty row)
  float* u[3] = {};           // [rows][C] dense outputs (pre-BN)
  float* hh[3] = {};          // [rows][C] layer outputs (post ReLU)
  float* pooled[3] = {};      // [voxels + 1][C] per-voxel max (row n_voxels: the empty voxels)
  float* bn[3] = {};          // [6][C]: a = gamma * inv_std, b = beta - mean * a, mean, inv_std, S1 / M, S2 / M
  float* gy = nullptr;        // [rows][64] gradient w.r.t. the BN output, then (in place) w.r.t. the dense output
  float* gdir = nullptr;      // [rows][32] gradient reaching a layer's output through the next layer's pointwise half
  float* ggath = nullptr;     // [rows][32] ... through the next layer's pooled half (summed per voxel by the pool backward)
  float* gpool = nullptr;     // [voxels + 1][64] gradient w.r.t. the last layer's per-voxel max
  double* partial = nullptr;  // [kTrBlocks][2 * 64] column-sum partials
  float* wpartial = nullptr;  // [kTrBlocks][64 * 64] weight-gradient partials
  float* bg = nullptr;        // [64] the empty voxels' output row (the grid's background)
  int n_sweeps = 0;
};

void free_vfe_train_state(VfeTrainState* s) {
  if (!s) return;
  void* ptrs[] = {s->x0, s->w, s->seg, s->u[0], s->u[1], s->u[2], s->hh[0], s->hh[1], s->hh[2], s->pooled[0], s->pooled[1],
                  s->pooled[2], s->bn[0], s->bn[1], s->bn[2], s->gy, s->gdir, s->ggath, s->gpool, s->partial, s->wpartial,
                  s->bg};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete s;
}

namespace {

constexpr int kC[3] = {16, 32, 64};

// ---- rows: centroid + features (model_training.py:134-141), weights, segments ------------------------------
template <typename PT>
__global__ void __launch_bounds__(256) tr_rows_kernel(const PT* __restrict__ row_xyz, const int* __restrict__ row_start,
                                                      const int* __restrict__ row_voxel,
                                                      const long long* __restrict__ totals, long long ncells, int T,
                                                      float* __restrict__ x0, float* __restrict__ w,
                                                      int* __restrict__ seg) {
  const long long V = totals[TOT_VOXELS], R = totals[TOT_ROWS];
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > V) return;
  if (v == V) {  // the empty voxels' row: zero input, T * n_empty copies
#pragma unroll
    for (int k = 0; k < 6; ++k) x0[R * 6 + k] = 0.f;
    w[R] = (float)((double)T * (double)(ncells - V));
    seg[R] = (int)V;
    return;
  }
  const int rs = row_start[v], re = row_start[v + 1];
  const bool pad = (row_voxel[re - 1] & kRowPadFlag) != 0;
  const int n = re - rs - (pad ? 1 : 0);
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int r = rs; r < rs + n; ++r) {  // np.mean: float64 adds in list order, one divide (:135)
    sx += (double)row_xyz[3 * (size_t)r];
    sy += (double)row_xyz[3 * (size_t)r + 1];
    sz += (double)row_xyz[3 * (size_t)r + 2];
  }
  const double dn = (double)n;
  const double cx = sx / dn, cy = sy / dn, cz = sz / dn;
  for (int r = rs; r < rs + n; ++r) {
    float f[6];
    point_features((double)row_xyz[3 * (size_t)r], (double)row_xyz[3 * (size_t)r + 1], (double)row_xyz[3 * (size_t)r + 2],
                   cx, cy, cz, f);
#pragma unroll
    for (int k = 0; k < 6; ++k) x0[(size_t)r * 6 + k] = f[k];
    w[r] = 1.f;
    seg[r] = (int)v;
  }
  if (pad) {
#pragma unroll
    for (int k = 0; k < 6; ++k) x0[(size_t)(re - 1) * 6 + k] = 0.f;
    w[re - 1] = (float)(T - n);
    seg[re - 1] = (int)v;
  }
}

// ---- dense forward: u[r] = in[r] * W, in = x0[r] (first layer) or [pooled_prev[seg[r]] | h_prev[r]] ---------
template <int CIN, int COUT, bool CONCAT>
__global__ void __launch_bounds__(kTrThreads) tr_dense_kernel(const float* __restrict__ x0,
                                                              const float* __restrict__ pooled_prev,
                                                              const float* __restrict__ h_prev,
                                                              const int* __restrict__ seg, const float* __restrict__ W,
                                                              const long long* __restrict__ totals,
                                                              float* __restrict__ u) {
  __shared__ __align__(16) float sW[CIN * COUT];
  for (int i = threadIdx.x; i < CIN * COUT; i += kTrThreads) sW[i] = W[i];
  __syncthreads();
  const long long R = totals[TOT_ROWS] + 1;
  const long long r = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (r >= R) return;
  float in[CIN];
  if (CONCAT) {
    constexpr int H = CIN / 2;
    const float* p = pooled_prev + (size_t)seg[r] * H;
    const float* q = h_prev + (size_t)r * H;
#pragma unroll
    for (int i = 0; i < H; i += 4) {  // 16-byte pieces of the two half rows
      const float4 p4 = *reinterpret_cast<const float4*>(p + i), q4 = *reinterpret_cast<const float4*>(q + i);
      in[i] = p4.x; in[i + 1] = p4.y; in[i + 2] = p4.z; in[i + 3] = p4.w;
      in[H + i] = q4.x; in[H + i + 1] = q4.y; in[H + i + 2] = q4.z; in[H + i + 3] = q4.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CIN; ++i) in[i] = x0[(size_t)r * CIN + i];
  }
  float* dst = u + (size_t)r * COUT;
  // eight outputs at a time: two 16-byte broadcast loads of the weights per eight FMAs (summation order over i unchanged)
#pragma unroll 1
  for (int c = 0; c < COUT; c += 8) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < CIN; ++i) {
      const float4 w0 = *reinterpret_cast<const float4*>(&sW[i * COUT + c]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sW[i * COUT + c + 4]);
      acc[0] = fmaf(in[i], w0.x, acc[0]);
      acc[1] = fmaf(in[i], w0.y, acc[1]);
      acc[2] = fmaf(in[i], w0.z, acc[2]);
      acc[3] = fmaf(in[i], w0.w, acc[3]);
      acc[4] = fmaf(in[i], w1.x, acc[4]);
      acc[5] = fmaf(in[i], w1.y, acc[5]);
      acc[6] = fmaf(in[i], w1.z, acc[6]);
      acc[7] = fmaf(in[i], w1.w, acc[7]);
    }
    *reinterpret_cast<float4*>(dst + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + c + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---- column sums over the rows, float64, fixed order ---------------------------------------------------------
// MODE 0: Sa = sum w u, Sb = sum w u^2 (batch statistics).   MODE 1: Sa = sum gy, Sb = sum gy * xhat (BN backward).
// MODE 2: Sa = sum a (plain column sum of a [n][C] array), Sb unused.
// thread = (channel, row slice); a block's partial lands in partial[block][2 * C]
template <int C, int MODE>
__global__ void __launch_bounds__(kTrThreads) tr_colsum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               const float* __restrict__ w, const float* __restrict__ bn,
                                                               const long long* __restrict__ n_ptr, long long n_extra,
                                                               long long n_fixed, double* __restrict__ partial) {
  constexpr int SL = kTrThreads / C;
  __shared__ double red[2][kTrThreads];
  const int c = threadIdx.x % C, sl = threadIdx.x / C;
  const long long n = n_ptr ? *n_ptr + n_extra : n_fixed;
  double sa = 0.0, sb = 0.0;
  float mean = 0.f, inv = 0.f;
  if (MODE == 1) {
    mean = bn[2 * C + c];
    inv = bn[3 * C + c];
  }
  for (long long r = (long long)blockIdx.x * SL + sl; r < n; r += (long long)gridDim.x * SL) {
    const float v = a[(size_t)r * C + c];
    if (MODE == 0) {
      const double wv = (double)w[r] * (double)v;
      sa += wv;
      sb += wv * (double)v;
    } else if (MODE == 1) {
      sa += (double)v;
      sb += (double)v * (double)((b[(size_t)r * C + c] - mean) * inv);
    } else {
      sa += (double)v;
    }
  }
  red[0][threadIdx.x] = sa;
  red[1][threadIdx.x] = sb;
  __syncthreads();
  if (sl == 0) {
    for (int s = 1; s < SL; ++s) {
      sa += red[0][s * C + c];
      sb += red[1][s * C + c];
    }
    partial[(size_t)blockIdx.x * 2 * C + c] = sa;
    partial[(size_t)blockIdx.x * 2 * C + C + c] = sb;
  }
}

// Column sums of the blocks' partials [nblocks][2][C] in a fixed order, by a block of kFinThreads threads that owns 8
// channels (blockIdx.x * 8 ..): thread (slice, channel) adds blocks slice, slice + 128, ..., thread (0, channel) then adds
// the 128 slice sums in order. (One thread per channel walking all 296 slots in a dependent chain of float64 adds took
// ~55 us per call, six calls per step.) Returns true for the threads that hold a channel's totals.
constexpr int kFinThreads = 1024, kFinChannels = 8;
template <int C>
__device__ __forceinline__ bool sum_partials(const double* __restrict__ partial, int nblocks, int& c, double& sa, double& sb) {
  __shared__ double s_a[kFinThreads], s_b[kFinThreads];
  constexpr int S = kFinThreads / kFinChannels;
  const int lc = threadIdx.x % kFinChannels, sl = threadIdx.x / kFinChannels;
  c = blockIdx.x * kFinChannels + lc;
  double a = 0.0, b = 0.0;
  for (int g = sl; g < nblocks; g += S) {
    a += partial[(size_t)g * 2 * C + c];
    b += partial[(size_t)g * 2 * C + C + c];
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (sl != 0) return false;
  for (int k = 1; k < S; ++k) {
    a += s_a[k * kFinChannels + lc];
    b += s_b[k * kFinChannels + lc];
  }
  sa = a;
  sb = b;
  return true;
}

// batch statistics -> folded BN, moving statistics (momentum 0.99; the biased batch variance: these 6-D inputs take
// Keras's non-fused BatchNormalization path)
template <int C>
__global__ void tr_stats_finalize_kernel(const double* __restrict__ partial, int nblocks, long long ncells, int T,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                         float momentum, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                         float* __restrict__ bn) {
  int c;
  double sa = 0.0, sb = 0.0;
  if (!sum_partials<C>(partial, nblocks, c, sa, sb)) return;
  const double M = (double)ncells * (double)T;
  const double mean = sa / M, var = fmax(sb / M - mean * mean, 0.0);
  const double inv = 1.0 / sqrt(var + (double)eps);
  const double a = (double)gamma[c] * inv;
  bn[c] = (float)a;
  bn[C + c] = (float)((double)beta[c] - mean * a);
  bn[2 * C + c] = (float)mean;
  bn[3 * C + c] = (float)inv;
  if (moving_mean) {
    moving_mean[c] = (float)((double)moving_mean[c] * momentum + mean * (1.0 - (double)momentum));
    moving_var[c] = (float)((double)moving_var[c] * momentum + var * (1.0 - (double)momentum));
  }
}

// BN backward sums -> dgamma = S2, dbeta = S1, and S1 / M, S2 / M for the data gradient
template <int C>
__global__ void tr_bn_bwd_finalize_kernel(const double* __restrict__ partial, int nblocks, long long ncells, int T,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ bn) {
  int c;
  double sa = 0.0, sb = 0.0;
  if (!sum_partials<C>(partial, nblocks, c, sa, sb)) return;
  const double M = (double)ncells * (double)T;
  dgamma[c] = (float)sb;
  dbeta[c] = (float)sa;
  bn[4 * C + c] = (float)(sa / M);
  bn[5 * C + c] = (float)(sb / M);
}

// ---- BN + ReLU + per-voxel max: thread = (voxel, 4 channels) ------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) tr_act_pool_kernel(const float* __restrict__ u, const float* __restrict__ bn,
                                                          const int* __restrict__ row_start,
                                                          const long long* __restrict__ totals, float* __restrict__ h,
                                                          float* __restrict__ pooled) {
  constexpr int Q = C / 4;
  const long long V = totals[TOT_VOXELS], R = totals[TOT_ROWS];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long v = idx / Q;
  const int q = (int)(idx % Q);
  if (v > V) return;
  const long long rs = v < V ? row_start[v] : R, re = v < V ? row_start[v + 1] : R + 1;
  const float4 a = *reinterpret_cast<const float4*>(bn + 4 * q);
  const float4 b = *reinterpret_cast<const float4*>(bn + C + 4 * q);
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (long long r = rs; r < re; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(u + (size_t)r * C + 4 * q);
    float4 y;
    y.x = fmaxf(fmaf(x.x, a.x, b.x), 0.f);
    y.y = fmaxf(fmaf(x.y, a.y, b.y), 0.f);
    y.z = fmaxf(fmaf(x.z, a.z, b.z), 0.f);
    y.w = fmaxf(fmaf(x.w, a.w, b.w), 0.f);
    *reinterpret_cast<float4*>(h + (size_t)r * C + 4 * q) = y;
    mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
  }
  *reinterpret_cast<float4*>(pooled + (size_t)v * C + 4 * q) = mx;
}

// ---- gradient arriving at the stack: d out[v] = dgrid[cell of v] ------------------------------------------
__global__ void __launch_bounds__(256) tr_gather_dout_kernel(const float* __restrict__ dgrid,
                                                             const int* __restrict__ voxel_cell,
                                                             const long long* __restrict__ totals,
                                                             float* __restrict__ gpool) {
  const long long V = totals[TOT_VOXELS];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long v = idx >> 4;
  const int q = (int)(idx & 15);
  if (v >= V) return;
  *reinterpret_cast<float4*>(gpool + (size_t)v * 64 + 4 * q) =
      *reinterpret_cast<const float4*>(dgrid + (size_t)voxel_cell[v] * 64 + 4 * q);
}
// the empty row's share: sum of dgrid over the EMPTY cells = (sum over all cells) - (sum over the occupied ones)
__global__ void __launch_bounds__(1024) tr_empty_dout_kernel(const double* __restrict__ part_all,
                                                             const double* __restrict__ part_occ, int nblocks,
                                                             const long long* __restrict__ totals, float* __restrict__ gpool) {
  // 8 channels per block x 128 slices of the partial slots, fixed order (as sum_partials, defined further down)
  __shared__ double s_d[1024];
  const int lc = threadIdx.x & 7, sl = threadIdx.x >> 3, c = blockIdx.x * 8 + lc;
  double d = 0.0;
  for (int b = sl; b < nblocks; b += 128) d += part_all[(size_t)b * 128 + c] - part_occ[(size_t)b * 128 + c];
  s_d[threadIdx.x] = d;
  __syncthreads();
  if (sl != 0) return;
  for (int k = 1; k < 128; ++k) d += s_d[k * 8 + lc];
  gpool[(size_t)totals[TOT_VOXELS] * 64 + c] = (float)d;
}

// ---- max + ReLU backward: thread = (voxel, 4 channels) -------------------------------------------------------
// gradient w.r.t. the per-voxel max: gpool[v] (last layer) or the sum over the voxel's rows of ggath (the next layer's
// pooled half); split equally among the tied copies; plus the next layer's pointwise half (gdir); through the ReLU.
template <int C, bool LAST>
__global__ void __launch_bounds__(256) tr_pool_bwd_kernel(const float* __restrict__ h, const float* __restrict__ pooled,
                                                          const float* __restrict__ w, const float* __restrict__ gpool,
                                                          const float* __restrict__ ggath, const float* __restrict__ gdir,
                                                          const int* __restrict__ row_start,
                                                          const long long* __restrict__ totals, float* __restrict__ gy) {
  constexpr int Q = C / 4;
  const long long V = totals[TOT_VOXELS], R = totals[TOT_ROWS];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long v = idx / Q;
  const int q = (int)(idx % Q);
  if (v > V) return;
  const long long rs = v < V ? row_start[v] : R, re = v < V ? row_start[v + 1] : R + 1;
  const float4 p = *reinterpret_cast<const float4*>(pooled + (size_t)v * C + 4 * q);
  float g[4], ties[4] = {0.f, 0.f, 0.f, 0.f};
  if (LAST) {
    const float4 t = *reinterpret_cast<const float4*>(gpool + (size_t)v * C + 4 * q);
    g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
  } else {
    g[0] = g[1] = g[2] = g[3] = 0.f;
  }
  for (long long r = rs; r < re; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(h + (size_t)r * C + 4 * q);
    const float wr = w[r];
    ties[0] += x.x == p.x ? wr : 0.f;
    ties[1] += x.y == p.y ? wr : 0.f;
    ties[2] += x.z == p.z ? wr : 0.f;
    ties[3] += x.w == p.w ? wr : 0.f;
    if (!LAST) {
      const float4 t = *reinterpret_cast<const float4*>(ggath + (size_t)r * C + 4 * q);
      g[0] += t.x; g[1] += t.y; g[2] += t.z; g[3] += t.w;
    }
  }
  const float pv[4] = {p.x, p.y, p.z, p.w};
  for (long long r = rs; r < re; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(h + (size_t)r * C + 4 * q);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float wr = w[r];
    float out[4];
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!LAST) d = *reinterpret_cast<const float4*>(gdir + (size_t)r * C + 4 * q);
    const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gh = dv[k] + (xv[k] == pv[k] ? wr * g[k] / ties[k] : 0.f);
      out[k] = xv[k] > 0.f ? gh : 0.f;
    }
    *reinterpret_cast<float4*>(gy + (size_t)r * C + 4 * q) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

// ---- BN + dense backward (data): thread = row ----------------------------------------------------------------
// gu = a * (gy - w * S1/M - w * xhat * S2/M) (weighted batch statistics), stored in place of gy;
// g_in = gu * W^T -> first half: ggath (pooled path), second half: gdir (pointwise path) of the previous layer
template <int CIN, int COUT, bool CONCAT>
__global__ void __launch_bounds__(kTrThreads) tr_dense_bwd_kernel(const float* __restrict__ u,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ bn,
                                                                  const float* __restrict__ W,
                                                                  const long long* __restrict__ totals,
                                                                  float* __restrict__ gy, float* __restrict__ ggath,
                                                                  float* __restrict__ gdir) {
  __shared__ __align__(16) float sWT[COUT * CIN];  // transposed: [c][i], so that four input channels are one 16-byte load
  __shared__ float sBn[6 * COUT];
  for (int e = threadIdx.x; e < CIN * COUT; e += kTrThreads) sWT[(e % COUT) * CIN + e / COUT] = W[e];
  for (int i = threadIdx.x; i < 6 * COUT; i += kTrThreads) sBn[i] = bn[i];
  __syncthreads();
  const long long R = totals[TOT_ROWS] + 1;
  const long long r = (long long)blockIdx.x * kTrThreads + threadIdx.x;
  if (r >= R) return;
  const float wr = w[r];
  float gin[CONCAT ? CIN : 1];
  if (CONCAT) {
#pragma unroll
    for (int i = 0; i < CIN; ++i) gin[i] = 0.f;
  }
  float* g = gy + (size_t)r * COUT;
  const float* ur = u + (size_t)r * COUT;
  // the row in 16-byte pieces (a thread's row is 4 * COUT contiguous bytes: scalar accesses would fetch every sector 8 times)
#pragma unroll 1
  for (int c = 0; c < COUT; c += 4) {
    const float4 u4 = *reinterpret_cast<const float4*>(ur + c);
    const float4 g4 = *reinterpret_cast<const float4*>(g + c);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float guv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xhat = (uu[k] - sBn[2 * COUT + c + k]) * sBn[3 * COUT + c + k];
      guv[k] = sBn[c + k] * (gg[k] - wr * sBn[4 * COUT + c + k] - wr * xhat * sBn[5 * COUT + c + k]);
    }
    *reinterpret_cast<float4*>(g + c) = make_float4(guv[0], guv[1], guv[2], guv[3]);
    if (CONCAT) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gu = guv[k];
#pragma unroll
        for (int i = 0; i < CIN; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&sWT[(c + k) * CIN + i]);
          gin[i] = fmaf(gu, w4.x, gin[i]);
          gin[i + 1] = fmaf(gu, w4.y, gin[i + 1]);
          gin[i + 2] = fmaf(gu, w4.z, gin[i + 2]);
          gin[i + 3] = fmaf(gu, w4.w, gin[i + 3]);
        }
      }
    }
  }
  if (CONCAT) {
    constexpr int H = CIN / 2;
#pragma unroll
    for (int i = 0; i < H; i += 4) {
      *reinterpret_cast<float4*>(ggath + (size_t)r * H + i) = make_float4(gin[i], gin[i + 1], gin[i + 2], gin[i + 3]);
      *reinterpret_cast<float4*>(gdir + (size_t)r * H + i) = make_float4(gin[H + i], gin[H + i + 1], gin[H + i + 2], gin[H + i + 3]);
    }
  }
}

// ---- dense backward (weights): dW[i][j] = sum_r in[r][i] * gu[r][j] ------------------------------------------
// persistent blocks; 32 rows at a time through shared memory; a thread owns entries tid, tid + 256, ... of the matrix
template <int CIN, int COUT, bool CONCAT>
__global__ void __launch_bounds__(kTrThreads) tr_wgrad_kernel(const float* __restrict__ x0,
                                                              const float* __restrict__ pooled_prev,
                                                              const float* __restrict__ h_prev,
                                                              const int* __restrict__ seg, const float* __restrict__ gu,
                                                              const long long* __restrict__ totals,
                                                              float* __restrict__ wpartial) {
  // a thread owns a T x T block of the matrix (T = 4 for 64 x 64, 2 for 32 x 32: 256 threads cover it), so that one row of
  // the chunk costs two vector loads from shared memory per T * T FMAs; the 6 x 16 first layer keeps one entry per thread.
  // Every entry still adds its rows in ascending order: the sums are bit-identical to the one-entry-per-thread version.
  constexpr int RB = 32, E = CIN * COUT;
  constexpr int T = E == 16 * kTrThreads ? 4 : (E == 4 * kTrThreads ? 2 : 1);
  constexpr int PADI = T > 1 ? 4 : 1;
  __shared__ __align__(16) float sIn[RB][CIN + PADI];
  __shared__ __align__(16) float sG[RB][COUT + PADI];
  const long long R = totals[TOT_ROWS] + 1;
  float acc[T][T];
#pragma unroll
  for (int a = 0; a < T; ++a)
#pragma unroll
    for (int b = 0; b < T; ++b) acc[a][b] = 0.f;
  const int i0 = T > 1 ? (threadIdx.x / (COUT / T)) * T : threadIdx.x / COUT;
  const int j0 = T > 1 ? (threadIdx.x % (COUT / T)) * T : threadIdx.x % COUT;
  const bool mine = T > 1 || threadIdx.x < E;
  for (long long r0 = (long long)blockIdx.x * RB; r0 < R; r0 += (long long)gridDim.x * RB) {
    __syncthreads();
    for (int t = threadIdx.x; t < RB * CIN; t += kTrThreads) {
      const int rr = t / CIN, i = t % CIN;
      const long long r = r0 + rr;
      float v = 0.f;
      if (r < R) {
        if (CONCAT) v = i < CIN / 2 ? pooled_prev[(size_t)seg[r] * (CIN / 2) + i] : h_prev[(size_t)r * (CIN / 2) + i - CIN / 2];
        else v = x0[(size_t)r * CIN + i];
      }
      sIn[rr][i] = v;
    }
    for (int t = threadIdx.x; t < RB * COUT; t += kTrThreads) {
      const int rr = t / COUT, j = t % COUT;
      const long long r = r0 + rr;
      sG[rr][j] = r < R ? gu[(size_t)r * COUT + j] : 0.f;
    }
    __syncthreads();
    if (mine) {
#pragma unroll 8
      for (int rr = 0; rr < RB; ++rr) {
        float xi[T], gj[T];
        if (T == 4) {
          const float4 x4 = *reinterpret_cast<const float4*>(&sIn[rr][i0]);
          const float4 g4 = *reinterpret_cast<const float4*>(&sG[rr][j0]);
          xi[0] = x4.x; xi[1] = x4.y; xi[T - 2] = x4.z; xi[T - 1] = x4.w;
          gj[0] = g4.x; gj[1] = g4.y; gj[T - 2] = g4.z; gj[T - 1] = g4.w;
        } else if (T == 2) {
          const float2 x2 = *reinterpret_cast<const float2*>(&sIn[rr][i0]);
          const float2 g2 = *reinterpret_cast<const float2*>(&sG[rr][j0]);
          xi[0] = x2.x; xi[T - 1] = x2.y;
          gj[0] = g2.x; gj[T - 1] = g2.y;
        } else {
          xi[0] = sIn[rr][i0];
          gj[0] = sG[rr][j0];
        }
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
          for (int b = 0; b < T; ++b) acc[a][b] = fmaf(xi[a], gj[b], acc[a][b]);
      }
    }
  }
  if (mine) {
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
      for (int b = 0; b < T; ++b) wpartial[(size_t)blockIdx.x * E + (size_t)(i0 + a) * COUT + j0 + b] = acc[a][b];
  }
}
__global__ void tr_wgrad_reduce_kernel(const float* __restrict__ wpartial, int nblocks, int E, float* __restrict__ dW) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double s4[4] = {0.0, 0.0, 0.0, 0.0};  // four loads in flight; fixed order
  int b = 0;
  for (; b + 3 < nblocks; b += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) s4[u] += (double)wpartial[(size_t)(b + u) * E + e];
  }
#pragma unroll
  for (int u = 0; u < 3; ++u)
    if (b + u < nblocks) s4[u] += (double)wpartial[(size_t)(b + u) * E + e];
  dW[e] = (float)((s4[0] + s4[1]) + (s4[2] + s4[3]));
}

__global__ void tr_copy_bg_kernel(const float* __restrict__ pooled, const long long* __restrict__ totals,
                                  float* __restrict__ bg) {
  if (threadIdx.x < 64) bg[threadIdx.x] = pooled[(size_t)totals[TOT_VOXELS] * 64 + threadIdx.x];
}

unsigned blocks_for(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

int ensure_state(lisec_handle* h) {
  if (h->train) return LISEC_OK;
  VfeTrainState* s = new (std::nothrow) VfeTrainState();
  if (!s) return fail(h, LISEC_ERR_CUDA, "out of host memory");
  h->train = s;
  s->max_vox = h->max_voxels + 1;
  s->max_rows = h->cfg.max_points + h->max_voxels + 1;
  const size_t Rm = (size_t)s->max_rows, Vm = (size_t)s->max_vox;
  LISEC_CUDA(h, dev_alloc(h, &s->x0, Rm * 6));
  LISEC_CUDA(h, dev_alloc(h, &s->w, Rm));
  LISEC_CUDA(h, dev_alloc(h, &s->seg, Rm));
  for (int l = 0; l < 3; ++l) {
    LISEC_CUDA(h, dev_alloc(h, &s->u[l], Rm * kC[l]));
    LISEC_CUDA(h, dev_alloc(h, &s->hh[l], Rm * kC[l]));
    LISEC_CUDA(h, dev_alloc(h, &s->pooled[l], Vm * kC[l]));
    LISEC_CUDA(h, dev_alloc(h, &s->bn[l], (size_t)6 * kC[l]));
  }
  LISEC_CUDA(h, dev_alloc(h, &s->gy, Rm * 64));
  LISEC_CUDA(h, dev_alloc(h, &s->gdir, Rm * 32));
  LISEC_CUDA(h, dev_alloc(h, &s->ggath, Rm * 32));
  LISEC_CUDA(h, dev_alloc(h, &s->gpool, Vm * 64));
  LISEC_CUDA(h, dev_alloc(h, &s->partial, (size_t)2 * kTrBlocks * 128));
  LISEC_CUDA(h, dev_alloc(h, &s->wpartial, (size_t)kTrBlocks * 64 * 64));
  LISEC_CUDA(h, dev_alloc(h, &s->bg, (size_t)64));
  return LISEC_OK;
}

template <int L>
cudaError_t forward_layer(lisec_handle* h, const lisec_vfe_train_params* p, long long ncells, cudaStream_t st) {
  VfeTrainState* s = h->train;
  constexpr int CIN = L == 0 ? 6 : (L == 1 ? 32 : 64), COUT = L == 0 ? 16 : (L == 1 ? 32 : 64);
  const long long* tot = h->ws.totals;
  tr_dense_kernel<CIN, COUT, (L > 0)><<<blocks_for(s->max_rows), kTrThreads, 0, st>>>(
      s->x0, L > 0 ? s->pooled[L - 1] : nullptr, L > 0 ? s->hh[L - 1] : nullptr, s->seg, p->dense_kernel[L], tot, s->u[L]);
  tr_colsum_kernel<COUT, 0><<<kTrBlocks, kTrThreads, 0, st>>>(s->u[L], nullptr, s->w, nullptr, tot + TOT_ROWS, 1, 0,
                                                             s->partial);
  tr_stats_finalize_kernel<COUT><<<COUT / kFinChannels, kFinThreads, 0, st>>>(s->partial, kTrBlocks, ncells, h->geom.T, p->bn_gamma[L], p->bn_beta[L],
                                                  p->bn_epsilon, p->bn_momentum, p->moving_mean[L], p->moving_var[L],
                                                  s->bn[L]);
  tr_act_pool_kernel<COUT><<<blocks_for(s->max_vox * (COUT / 4)), 256, 0, st>>>(s->u[L], s->bn[L], h->ws.row_start, tot,
                                                                              s->hh[L], s->pooled[L]);
  h->launches += 4;
  return cudaGetLastError();
}

template <int L>
cudaError_t backward_layer(lisec_handle* h, const lisec_vfe_train_params* p, const lisec_vfe_train_grads* g,
                           long long ncells, cudaStream_t st) {
  VfeTrainState* s = h->train;
  constexpr int CIN = L == 0 ? 6 : (L == 1 ? 32 : 64), COUT = L == 0 ? 16 : (L == 1 ? 32 : 64);
  const long long* tot = h->ws.totals;
  tr_pool_bwd_kernel<COUT, (L == 2)><<<blocks_for(s->max_vox * (COUT / 4)), 256, 0, st>>>(
      s->hh[L], s->pooled[L], s->w, s->gpool, s->ggath, s->gdir, h->ws.row_start, tot, s->gy);
  tr_colsum_kernel<COUT, 1><<<kTrBlocks, kTrThreads, 0, st>>>(s->gy, s->u[L], nullptr, s->bn[L], tot + TOT_ROWS, 1, 0,
                                                             s->partial);
  tr_bn_bwd_finalize_kernel<COUT><<<COUT / kFinChannels, kFinThreads, 0, st>>>(s->partial, kTrBlocks, ncells, h->geom.T, g->dgamma[L], g->dbeta[L],
                                                   s->bn[L]);
  // (ggath / gdir of THIS layer have been consumed by the pool backward above: the data gradient may overwrite them)
  tr_dense_bwd_kernel<CIN, COUT, (L > 0)><<<blocks_for(s->max_rows), kTrThreads, 0, st>>>(
      s->u[L], s->w, s->bn[L], p->dense_kernel[L], tot, s->gy, s->ggath, s->gdir);
  tr_wgrad_kernel<CIN, COUT, (L > 0)><<<kTrBlocks, kTrThreads, 0, st>>>(
      s->x0, L > 0 ? s->pooled[L - 1] : nullptr, L > 0 ? s->hh[L - 1] : nullptr, s->seg, s->gy, tot, s->wpartial);
  tr_wgrad_reduce_kernel<<<blocks_for(CIN * COUT), 256, 0, st>>>(s->wpartial, kTrBlocks, CIN * COUT, g->dkernel[L]);
  h->launches += 6;
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_grid_write(const Geom& g, int n_sweeps, int c3, int grid_dtype, const int* cell_voxel,
                              const float* voxel_feat, const float* c_empty, void* grid, int sm_count,
                              cudaStream_t st, int* launches);

}  // namespace lisec

using namespace lisec;

extern "C" {

int32_t lisec_vfe_train_forward(lisec_handle* h, const lisec_vfe_train_params* p, void* grid, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!p || !grid) return fail(h, LISEC_ERR_BAD_ARG, "params / grid is NULL");
  if (h->cfg.c1 != 16 || h->cfg.c2 != 32 || h->cfg.c3 != 64 || h->cfg.fcn_post_dense)
    return fail(h, LISEC_ERR_UNSUPPORTED, "training is built for the graph train() creates (createModel, model_training.py:"
                "222-257: widths 16, 32, 64, Dense-BN-ReLU); this handle holds (%d, %d, %d, post_dense = %d)", h->cfg.c1,
                h->cfg.c2, h->cfg.c3, h->cfg.fcn_post_dense);
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  for (int l = 0; l < 3; ++l)
    if (!p->dense_kernel[l] || !p->bn_gamma[l] || !p->bn_beta[l])
      return fail(h, LISEC_ERR_BAD_ARG, "training parameters of layer %d contain a NULL pointer", l);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  int rc = ensure_state(h);
  if (rc) return rc;
  h->launches = 0;
  VfeTrainState* s = h->train;
  s->n_sweeps = h->last_so.n;
  const long long ncells = (long long)h->last_so.n * h->geom.cells;
  if (h->last_dtype == LISEC_F32)
    tr_rows_kernel<float><<<blocks_for(s->max_vox), 256, 0, st>>>(static_cast<const float*>(h->ws.row_xyz), h->ws.row_start,
                                                               h->ws.row_voxel, h->ws.totals, ncells, h->geom.T, s->x0,
                                                               s->w, s->seg);
  else
    tr_rows_kernel<double><<<blocks_for(s->max_vox), 256, 0, st>>>(static_cast<const double*>(h->ws.row_xyz),
                                                                h->ws.row_start, h->ws.row_voxel, h->ws.totals, ncells,
                                                                h->geom.T, s->x0, s->w, s->seg);
  ++h->launches;
  LISEC_CUDA(h, cudaGetLastError());
  LISEC_CUDA(h, forward_layer<0>(h, p, ncells, st));
  LISEC_CUDA(h, forward_layer<1>(h, p, ncells, st));
  LISEC_CUDA(h, forward_layer<2>(h, p, ncells, st));
  // the voxel rows into the grid, the empty voxels' output (row n_voxels of the last pooled table, an address only the
  // device knows) as the background
  tr_copy_bg_kernel<<<1, 64, 0, st>>>(s->pooled[2], h->ws.totals, s->bg);
  ++h->launches;
  LISEC_CUDA(h, cudaGetLastError());
  LISEC_CUDA(h, launch_grid_write(h->geom, h->last_so.n, 64, h->cfg.grid_dtype, h->ws.cell_voxel, s->pooled[2], s->bg, grid,
                                  h->sm_count, st, &h->launches));
  return LISEC_OK;
}

int32_t lisec_vfe_train_backward(lisec_handle* h, const lisec_vfe_train_params* p, const float* dgrid,
                                 const lisec_vfe_train_grads* g, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!p || !dgrid || !g) return fail(h, LISEC_ERR_BAD_ARG, "params / dgrid / grads is NULL");
  if (!h->train || h->train->n_sweeps == 0) return fail(h, LISEC_ERR_STATE, "no lisec_vfe_train_forward() on this handle");
  for (int l = 0; l < 3; ++l)
    if (!g->dkernel[l] || !g->dgamma[l] || !g->dbeta[l] || !p->dense_kernel[l])
      return fail(h, LISEC_ERR_BAD_ARG, "gradient outputs of layer %d contain a NULL pointer", l);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  VfeTrainState* s = h->train;
  const long long ncells = (long long)s->n_sweeps * h->geom.cells;
  const long long* tot = h->ws.totals;
  // d out[v] = dgrid[cell of v]; the empty row gets the sum over the empty cells
  tr_gather_dout_kernel<<<blocks_for(s->max_vox * 16), 256, 0, st>>>(dgrid, h->ws.voxel_cell, tot, s->gpool);
  tr_colsum_kernel<64, 2><<<kTrBlocks, kTrThreads, 0, st>>>(dgrid, nullptr, nullptr, nullptr, nullptr, 0, ncells, s->partial);
  tr_colsum_kernel<64, 2><<<kTrBlocks, kTrThreads, 0, st>>>(s->gpool, nullptr, nullptr, nullptr, tot + TOT_VOXELS, 0, 0,
                                                           s->partial + (size_t)kTrBlocks * 128);
  tr_empty_dout_kernel<<<8, 1024, 0, st>>>(s->partial, s->partial + (size_t)kTrBlocks * 128, kTrBlocks, tot, s->gpool);
  h->launches += 4;
  LISEC_CUDA(h, cudaGetLastError());
  LISEC_CUDA(h, backward_layer<2>(h, p, g, ncells, st));
  LISEC_CUDA(h, backward_layer<1>(h, p, g, ncells, st));
  LISEC_CUDA(h, backward_layer<0>(h, p, g, ncells, st));
  return LISEC_OK;
}

/* Test aid: the training forward's per-voxel output rows [n_voxels + 1][64] (the last row = the empty voxels) and the
   batch statistics (mean, inverse standard deviation) of layer `layer`, copied to the host. Synchronous. */
int32_t lisec_vfe_train_read(lisec_handle* h, int32_t layer, float* out_rows, int64_t n_rows, float* mean, float* inv_std) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!h->train) return fail(h, LISEC_ERR_STATE, "no lisec_vfe_train_forward() on this handle");
  if (layer < 0 || layer > 2) return fail(h, LISEC_ERR_BAD_ARG, "layer = %d", layer);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  LISEC_CUDA(h, cudaDeviceSynchronize());
  const int C = kC[layer];
  if (out_rows)
    LISEC_CUDA(h, cudaMemcpy(out_rows, h->train->pooled[layer], sizeof(float) * (size_t)n_rows * C, cudaMemcpyDeviceToHost));
  if (mean) LISEC_CUDA(h, cudaMemcpy(mean, h->train->bn[layer] + 2 * C, sizeof(float) * C, cudaMemcpyDeviceToHost));
  if (inv_std) LISEC_CUDA(h, cudaMemcpy(inv_std, h->train->bn[layer] + 3 * C, sizeof(float) * C, cudaMemcpyDeviceToHost));
  return LISEC_OK;
}

}  // extern "C"
