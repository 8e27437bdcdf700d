// Stacked VFE on sm_100a: centroid augmentation, three pointwise linears (+BN+ReLU), two per-voxel max-pools with
// concat, and the final max over T — reference model_training.py:134-141 (features) and :155-186, 229-235 (layers).
//
// One thread owns one VFE row (a kept point, or the single virtual zero row that stands for all identical pad rows
// of a non-full voxel). A tile is a run of whole voxels holding at most kVfeThreads rows. The
// pointwise products run entirely in registers with the weights arriving as uniform-register / constant-bank
// operands (VfeParams is a __grid_constant__ parameter), so the inner loops are pure FFMA. Because
// Concatenate([pooled, pointwise]) feeds a bias-free Dense, the pooled half of the next layer's product is the
// same for every row of a voxel: it is computed once per voxel and added as the accumulator's initial value.
// Per-voxel max-pools go through shared memory (row-major, odd stride, conflict-free), not atomics.
#include "common.cuh"

namespace lisec {

int vfe_rows_per_tile(int T) { return kVfeThreads - T + 1; }

namespace {

constexpr int kHS = 65;                     // float stride of a row in sH (odd: conflict-free both ways)
constexpr int kPS = 66;                     // float stride of a row in sP (even: a row can also hold 32 doubles)
constexpr int kMaxVox = kVfeThreads / 2;    // every voxel has >= 2 rows unless it is full (then T rows)

template <typename PT>
__device__ __forceinline__ void load_point(const PT* __restrict__ pts, long long p, PT& x, PT& y, PT& z) {
  x = __ldg(pts + 3 * p);
  y = __ldg(pts + 3 * p + 1);
  z = __ldg(pts + 3 * p + 2);
}

// [x, y, z, x-cx, y-cy, z-cz]: subtraction in float64, one rounding to float32 (model_training.py:137-140 and the
// float32 cast at the Keras model input)
__device__ __forceinline__ void point_features(double x, double y, double z, double cx, double cy, double cz,
                                               float (&f)[6]) {
  f[0] = __double2float_rn(x);
  f[1] = __double2float_rn(y);
  f[2] = __double2float_rn(z);
  f[3] = __double2float_rn(x - cx);
  f[4] = __double2float_rn(y - cy);
  f[5] = __double2float_rn(z - cz);
}

template <int C>
__device__ __forceinline__ void pool_rows(const float* __restrict__ sH, const int* __restrict__ sRowOff, int nv,
                                          float* __restrict__ dst, int dst_stride) {
  for (int idx = threadIdx.x; idx < nv * C; idx += kVfeThreads) {
    const int lv = idx / C, ch = idx % C;
    const int rb = sRowOff[lv], re = sRowOff[lv + 1];
    float m = sH[rb * kHS + ch];
    for (int r = rb + 1; r < re; ++r) m = fmaxf(m, sH[r * kHS + ch]);
    dst[lv * dst_stride + ch] = m;
  }
}

struct VfeSmem {
  float* sH;      // [kVfeThreads][kHS]  layer outputs, one row per thread
  float* sP;      // [kMaxVox][kHS]      pooled vector, then (in place) its product with the pooled-half weights
  double* sCen;   // [kMaxVox][3]
  int* sRowOff;   // [kMaxVox + 1]
  int* sKept;     // [kMaxVox]
  int* sEstart;   // [kMaxVox]
};

template <typename PT>
constexpr size_t vfe_smem_bytes() {
  return sizeof(float) * (kVfeThreads * kHS + kMaxVox * kPS) + sizeof(double) * kMaxVox * 3 +
         sizeof(PT) * kVfeThreads * 3 + sizeof(int) * (3 * kMaxVox + 4);
}

template <typename PT>
__global__ void __launch_bounds__(kVfeThreads, 2)
    vfe_kernel(const PT* __restrict__ pts, const __grid_constant__ VfeParams P, int T,
               const int* __restrict__ tile_first, const int* __restrict__ voxel_start,
               const int* __restrict__ row_start, const int* __restrict__ list_sorted,
               const long long* __restrict__ n_tiles_ptr, float* __restrict__ voxel_feat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sCen = reinterpret_cast<double*>(smem_raw);
  PT* sPt = reinterpret_cast<PT*>(sCen + kMaxVox * 3);
  float* sH = reinterpret_cast<float*>(sPt + kVfeThreads * 3);
  float* sP = sH + kVfeThreads * kHS;  // 8-byte aligned: kVfeThreads * kHS is even
  // pooled half of dense_1 in float64: row lv of sP viewed as doubles, floats [2, 66) (the voxel's own thread has
  // read its pooled vector out of floats [0,16) before it writes these)
  auto sQ2 = [sP](int lv) { return reinterpret_cast<double*>(sP + lv * kPS + 2); };
  int* sRowOff = reinterpret_cast<int*>(sP + kMaxVox * kPS);
  int* sKept = sRowOff + kMaxVox + 1;
  int* sEstart = sKept + kMaxVox;

  const int tid = threadIdx.x;
  const int n_tiles = (int)*n_tiles_ptr;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int v0 = tile_first[t];
    const int nv = tile_first[t + 1] - v0;
    const int r0 = row_start[v0];
    if (tid <= nv) sRowOff[tid] = row_start[v0 + tid] - r0;
    if (tid < nv) {
      const int s = voxel_start[v0 + tid];
      const int c = voxel_start[v0 + tid + 1] - s;
      sKept[tid] = c < T ? c : T;
      sEstart[tid] = s;
    }
    __syncthreads();
    const int nrows = sRowOff[nv];
    const bool has_row = tid < nrows;
    int lv = 0;
    bool real = false;
    PT px = PT(0), py = PT(0), pz = PT(0);
    if (has_row) {
      int lo = 0, hi = nv;  // largest lv with sRowOff[lv] <= tid
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sRowOff[mid] <= tid) lo = mid; else hi = mid;
      }
      lv = lo;
      const int i = tid - sRowOff[lv];
      real = i < sKept[lv];
      if (real) load_point(pts, (long long)list_sorted[sEstart[lv] + i], px, py, pz);
      sPt[tid * 3] = px;
      sPt[tid * 3 + 1] = py;
      sPt[tid * 3 + 2] = pz;
    }
    __syncthreads();
    // centroid = np.mean(currPoints, axis=0) (model_training.py:135): float64, rows added in list order, one divide
    if (tid < nv) {
      const int kept = sKept[tid];
      double sx = 0.0, sy = 0.0, sz = 0.0;
      const PT* q = sPt + sRowOff[tid] * 3;
      for (int i = 0; i < kept; ++i) {
        sx += (double)q[3 * i];
        sy += (double)q[3 * i + 1];
        sz += (double)q[3 * i + 2];
      }
      const double inv_n = (double)(kept > 0 ? kept : 1);
      sCen[tid * 3] = sx / inv_n;
      sCen[tid * 3 + 1] = sy / inv_n;
      sCen[tid * 3 + 2] = sz / inv_n;
    }
    __syncthreads();

    // ---- VFE-1: Dense(6->16, no bias) + BN + ReLU (addVFELayer(in, 6, 32), :231 -> :155-166) ----
    float h1[16];
    if (has_row) {
      float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // pad row: six zeros (:141)
      if (real) point_features((double)px, (double)py, (double)pz, sCen[lv * 3], sCen[lv * 3 + 1], sCen[lv * 3 + 2], f);
#pragma unroll
      for (int j = 0; j < 16; ++j) h1[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k)  // k outer, j inner: the weights of one k are contiguous -> 128-bit constant loads
#pragma unroll
        for (int j = 0; j < 16; ++j) h1[j] = fmaf(f[k], P.w1[k][j], h1[j]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        h1[j] = fmaxf(fmaf(h1[j], P.a1[j], P.b1[j]), 0.f);
        sH[tid * kHS + j] = h1[j];
      }
    }
    __syncthreads();
    pool_rows<16>(sH, sRowOff, nv, sP, kPS);  // MaxPoolingVFELayer over T (:160); RepeatLayer is implicit
    __syncthreads();
    if (tid < nv) {  // pooled half of dense_1, once per voxel
      float pool[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) pool[k] = sP[tid * kPS + k];
      double acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma((double)pool[k], P.w2p[k][j], acc[j]);
#pragma unroll
      for (int j = 0; j < 32; ++j) sQ2(tid)[j] = acc[j];
    }
    __syncthreads();

    // ---- VFE-2: Dense(32->32) + BN + ReLU on concat[pooled, pointwise] (addVFELayer(., 32, 64), :232) ----
    float h2[32];
    if (has_row) {
      // float64 accumulation of the whole 32-term product (pooled half first), one rounding to float32
#pragma unroll
      for (int jc = 0; jc < 32; jc += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sQ2(lv)[jc + j];
#pragma unroll
        for (int k = 0; k < 16; ++k)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fma((double)h1[k], P.w2x[k][jc + j], acc[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          h2[jc + j] = fmaxf(fmaf(__double2float_rn(acc[j]), P.a2[jc + j], P.b2[jc + j]), 0.f);
          sH[tid * kHS + jc + j] = h2[jc + j];
        }
      }
    }
    __syncthreads();
    pool_rows<32>(sH, sRowOff, nv, sP, kPS);
    __syncthreads();
    if (tid < nv) {  // pooled half of dense_2
      float pool[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) pool[k] = sP[tid * kPS + k];
float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k)
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = fmaf(pool[k], P.w3p[k][j], acc[j]);
#pragma unroll
      for (int j = 0; j < 64; ++j) sP[tid * kPS + j] = acc[j];
    }
    __syncthreads();

    // ---- FCN: Dense(64->64) + BN + ReLU (addFCN(., 64, 64), :233) ----
    if (has_row) {
float acc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) acc[j] = sP[lv * kPS + j];
#pragma unroll
      for (int k = 0; k < 32; ++k)
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = fmaf(h2[k], P.w3x[k][j], acc[j]);
#pragma unroll
      for (int j = 0; j < 64; ++j) sH[tid * kHS + j] = fmaxf(fmaf(acc[j], P.a3[j], P.b3[j]), 0.f);
    }
    __syncthreads();
    // MaxPoolingVFELayer(combine=True) (:235): one C3 row per voxel, written coalesced
    pool_rows<64>(sH, sRowOff, nv, voxel_feat + (size_t)v0 * 64, 64);
    __syncthreads();
  }
}

// ---- export of the grouping in the reference's terms (tests, drop-in COO/dense emission) ------------------
template <typename PT>
__global__ void __launch_bounds__(256) export_kernel(const PT* __restrict__ pts, const __grid_constant__ SweepOffsets so,
                                                     const __grid_constant__ Geom g,
                                                     const int* __restrict__ voxel_cell,
                                                     const int* __restrict__ voxel_start,
                                                     const int* __restrict__ list_sorted,
                                                     const long long* __restrict__ totals, int32_t* __restrict__ coords,
                                                     int32_t* __restrict__ counts, int32_t* __restrict__ point_idx,
                                                     float* __restrict__ features, float* __restrict__ dense) {
  const int lane = threadIdx.x & 31;
  const long long n_voxels = totals[TOT_VOXELS];
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < n_voxels; v += warps) {
    const int cell = voxel_cell[v];
    const int sweep = cell / g.cells;
    const int rem = cell - sweep * g.cells;
    const int s = voxel_start[v];
    const int c = voxel_start[v + 1] - s;
    const int kept = c < g.T ? c : g.T;
    if (lane == 0) {
      if (coords) {
        coords[4 * v] = sweep;
        coords[4 * v + 1] = rem / (g.nx * g.ny);
        coords[4 * v + 2] = (rem / g.ny) % g.nx;
        coords[4 * v + 3] = rem % g.ny;
      }
      if (counts) counts[v] = c;
    }
    double cx = 0.0, cy = 0.0, cz = 0.0;
    if (features || dense) {
      for (int i = 0; i < kept; ++i) {  // same operation order as the VFE kernel: sequential float64 adds
        PT x, y, z;
        load_point(pts, (long long)list_sorted[s + i], x, y, z);
        cx += (double)x;
        cy += (double)y;
        cz += (double)z;
      }
      const double n = (double)(kept > 0 ? kept : 1);
      cx /= n;
      cy /= n;
      cz /= n;
    }
    for (int i = lane; i < g.T; i += 32) {
      const bool real = i < kept;
      const int p = real ? list_sorted[s + i] : -1;
      if (point_idx) point_idx[v * g.T + i] = real ? (int)(p - so.off[sweep]) : -1;
      if (features || dense) {
        float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (real) {
          PT x, y, z;
          load_point(pts, (long long)p, x, y, z);
          point_features((double)x, (double)y, (double)z, cx, cy, cz, f);
        }
        if (features) {
#pragma unroll
          for (int j = 0; j < 6; ++j) features[(v * g.T + i) * 6 + j] = f[j];
        }
        if (dense && real) {  // dense was zero-filled: sparse.to_dense(default_value=0.) (model_training.py:279)
#pragma unroll
          for (int j = 0; j < 6; ++j) dense[((long long)cell * g.T + i) * 6 + j] = f[j];
        }
      }
    }
  }
}

}  // namespace

cudaError_t launch_vfe(const void* pts, int pts_dtype, const Geom& g, const VfeParams& p, const int* tile_first,
                       const int* voxel_start, const int* row_start, const int* list_sorted,
                       const long long* n_tiles, float* voxel_feat, int sm_count, cudaStream_t st,
                       int* launches) {
  cudaError_t err;
  const int grid = 2 * sm_count;  // persistent: two resident CTAs per SM, tiles strided over them
  if (pts_dtype == LISEC_F32) {
    err = cudaFuncSetAttribute(vfe_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)vfe_smem_bytes<float>());
    if (err != cudaSuccess) return err;
    vfe_kernel<float><<<grid, kVfeThreads, vfe_smem_bytes<float>(), st>>>(
        static_cast<const float*>(pts), p, g.T, tile_first, voxel_start, row_start, list_sorted, n_tiles,
        voxel_feat);
  } else {
    err = cudaFuncSetAttribute(vfe_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)vfe_smem_bytes<double>());
    if (err != cudaSuccess) return err;
    vfe_kernel<double><<<grid, kVfeThreads, vfe_smem_bytes<double>(), st>>>(
        static_cast<const double*>(pts), p, g.T, tile_first, voxel_start, row_start, list_sorted, n_tiles,
        voxel_feat);
  }
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_export(const void* pts, int pts_dtype, const SweepOffsets& so, const Geom& g,
                          const Workspace& w, long long max_voxels, int32_t* coords, int32_t* counts,
                          int32_t* point_idx, float* features, float* dense, cudaStream_t st, int* launches) {
  long long blocks = (max_voxels * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (pts_dtype == LISEC_F32)
    export_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(pts), so, g, w.voxel_cell,
                                                           w.voxel_start, w.list_sorted, w.totals, coords, counts,
                                                           point_idx, features, dense);
  else
    export_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const double*>(pts), so, g, w.voxel_cell,
                                                            w.voxel_start, w.list_sorted, w.totals, coords, counts,
                                                            point_idx, features, dense);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace lisec
