// The VFE stack for EVERY graph the reference's .h5 may hold — parametric in the widths (C1, C2, C3) and in
// the FCN variant — in plain float32 FMAs, one thread per VFE row:
//   POST = false  addFCN as the code stands: Dense -> BatchNormalization -> ReLU                    (model_training.py:169-174)
//   POST = true   the line commented out at :172 switched on (the graph model.png shows):
//                 Dense -> BatchNormalization -> Dense(units, relu, no bias)
// around addVFELayer's max-pool / repeat / concat (:155-166) and MaxPoolingVFELayer(combine=True) (:235).
//
// The current createModel() widths (16, 32, 64, POST = false) have the tuned tensor-core kernel (vfe.cu); this kernel is
// what serves any other weight set (and, under LISEC_GENERIC_VFE=1, the current one too: the two kernels check each other
// in tests/test_gpu_arch.py). It reads the same tiles the grouping chain packs for vfe.cu (<= 128 rows and <= 64 whole
// voxels per tile; every non-full voxel's virtual pad row is a row of its tile with a zero input), so the pad rows need
// no special case: they run through the stack like any other row and join their voxel's max.
//
// Per tile, 128 threads: thread t owns row t. A Dense layer is CIN x COUT FMAs per thread with the input row in registers
// and the weights fetched as warp-uniform float4 loads (read-only path; the whole weight set, <= 157 KB, lives in L1/L2);
// the outputs go to the thread's own row of a shared-memory tile, where the per-voxel max (one thread per voxel x 4
// channels) finds them. Bound by the FP32 pipe: 39 264 MACs per row for (16, 64, 128, POST), 5 216 for (16, 32, 64).
#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

namespace {

constexpr int kGenThreads = 128;  // = rows per tile
constexpr int kGenVox = 64;       // voxels per tile

template <int C1, int C2, int C3, bool POST>
struct GenericLayout {
  // float offsets into the parameter block: per FCN  W [cin][cout] | a [cout] | b [cout] | D [cout][cout] (POST only)
  __host__ __device__ static constexpr int cin(int s) { return s == 0 ? 6 : (s == 1 ? 2 * C1 : 2 * C2); }
  __host__ __device__ static constexpr int cout(int s) { return s == 0 ? C1 : (s == 1 ? C2 : C3); }
  __host__ __device__ static constexpr int stage_floats(int s) {
    return cin(s) * cout(s) + 2 * cout(s) + (POST ? cout(s) * cout(s) : 0);
  }
  __host__ __device__ static constexpr int w(int s) {
    return s == 0 ? 0 : (s == 1 ? stage_floats(0) : stage_floats(0) + stage_floats(1));
  }
  __host__ __device__ static constexpr int a(int s) { return w(s) + cin(s) * cout(s); }
  __host__ __device__ static constexpr int b(int s) { return a(s) + cout(s); }
  __host__ __device__ static constexpr int d(int s) { return b(s) + cout(s); }
  static constexpr int total = stage_floats(0) + stage_floats(1) + stage_floats(2);
  static constexpr int kCH = C3 > C2 ? (C3 > C1 ? C3 : C1) : (C2 > C1 ? C2 : C1);  // widest row the tile holds
  static constexpr int kHS = kCH + 4;                                                // row stride: conflict-free float4 rows
  static constexpr int kCP = C2 > C1 ? C2 : C1;                                      // widest pooled row that is re-read
  static constexpr size_t smem = sizeof(float) * ((size_t)kGenThreads * kHS + (size_t)kGenVox * kCP) +
                                 sizeof(double) * 3 * kGenVox + sizeof(int) * (kGenVox + 4);
};

// out[c] = act(affine(sum_i in[i] * W[i][c])), eight columns at a time, into the thread's shared-memory row
template <int CIN, int COUT, bool AFFINE, bool RELU>
__device__ __forceinline__ void row_dense(const float (&in)[CIN], const float* __restrict__ W, const float* __restrict__ a,
                                          const float* __restrict__ b, float* __restrict__ dst) {
#pragma unroll 1
  for (int c = 0; c < COUT; c += 8) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < CIN; ++i) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + i * COUT + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + i * COUT + c + 4));
      acc[0] = fmaf(in[i], w0.x, acc[0]);
      acc[1] = fmaf(in[i], w0.y, acc[1]);
      acc[2] = fmaf(in[i], w0.z, acc[2]);
      acc[3] = fmaf(in[i], w0.w, acc[3]);
      acc[4] = fmaf(in[i], w1.x, acc[4]);
      acc[5] = fmaf(in[i], w1.y, acc[5]);
      acc[6] = fmaf(in[i], w1.z, acc[6]);
      acc[7] = fmaf(in[i], w1.w, acc[7]);
    }
    if (AFFINE) {  // BatchNormalization at inference, folded on the host: y = z * a + b
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(a + c)), a1 = __ldg(reinterpret_cast<const float4*>(a + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
      acc[0] = fmaf(acc[0], a0.x, b0.x);
      acc[1] = fmaf(acc[1], a0.y, b0.y);
      acc[2] = fmaf(acc[2], a0.z, b0.z);
      acc[3] = fmaf(acc[3], a0.w, b0.w);
      acc[4] = fmaf(acc[4], a1.x, b1.x);
      acc[5] = fmaf(acc[5], a1.y, b1.y);
      acc[6] = fmaf(acc[6], a1.z, b1.z);
      acc[7] = fmaf(acc[7], a1.w, b1.w);
    }
    if (RELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.f);
    }
    *reinterpret_cast<float4*>(dst + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + c + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// one FCN on the thread's row: Dense -> BN -> ReLU, or Dense -> BN -> Dense -> ReLU; result in hrow[0..COUT)
template <int CIN, int COUT, bool POST>
__device__ __forceinline__ void row_fcn(const float (&in)[CIN], const float* __restrict__ W, const float* __restrict__ a,
                                        const float* __restrict__ b, const float* __restrict__ D, float* __restrict__ hrow) {
  row_dense<CIN, COUT, true, !POST>(in, W, a, b, hrow);
  if (POST) {
    float u[COUT];
#pragma unroll
    for (int i = 0; i < COUT; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(hrow + i);
      u[i] = v.x;
      u[i + 1] = v.y;
      u[i + 2] = v.z;
      u[i + 3] = v.w;
    }
    row_dense<COUT, COUT, false, true>(u, D, nullptr, nullptr, hrow);
  }
}

// per-voxel max over the voxel's rows of the tile (kept rows and, for a non-full voxel, its pad row): MaxPoolingVFELayer
template <int C, int HS>
__device__ __forceinline__ void pool_rows(const float* __restrict__ H, const int* __restrict__ rs, int nvox, int tid,
                                          float* __restrict__ dst, int dst_stride) {
  constexpr int G = C / 4;
  for (int item = tid; item < nvox * G; item += kGenThreads) {
    const int v = item / G, g = item - v * G;
    const int r0 = rs[v], r1 = rs[v + 1];
    float4 m = *reinterpret_cast<const float4*>(H + (size_t)r0 * HS + 4 * g);
    for (int r = r0 + 1; r < r1; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(H + (size_t)r * HS + 4 * g);
      m.x = fmaxf(m.x, x.x);
      m.y = fmaxf(m.y, x.y);
      m.z = fmaxf(m.z, x.z);
      m.w = fmaxf(m.w, x.w);
    }
    *reinterpret_cast<float4*>(dst + (size_t)v * dst_stride + 4 * g) = m;
  }
}

// [pooled[voxel] | own row]: Concatenate([pooling, layer]) (:164-165) as the next Dense's input row
template <int C>
__device__ __forceinline__ void load_concat(const float* __restrict__ pooled, const float* __restrict__ hrow,
                                            float (&in)[2 * C]) {
#pragma unroll
  for (int i = 0; i < C; i += 4) {
    const float4 p = *reinterpret_cast<const float4*>(pooled + i);
    const float4 h = *reinterpret_cast<const float4*>(hrow + i);
    in[i] = p.x;
    in[i + 1] = p.y;
    in[i + 2] = p.z;
    in[i + 3] = p.w;
    in[C + i] = h.x;
    in[C + i + 1] = h.y;
    in[C + i + 2] = h.z;
    in[C + i + 3] = h.w;
  }
}

template <int C1, int C2, int C3, bool POST, typename PT>
__global__ void __launch_bounds__(kGenThreads) vfe_generic_kernel(const float* __restrict__ params, const VfeProblem prob,
                                                                  float* __restrict__ voxel_feat) {
  using L = GenericLayout<C1, C2, C3, POST>;
  constexpr int HS = L::kHS, CP = L::kCP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* H = reinterpret_cast<float*>(smem_raw);          // [128][HS] the tile's rows, current layer
  float* PL = H + (size_t)kGenThreads * HS;                // [64][CP] per-voxel max of the current layer
  double* cen = reinterpret_cast<double*>(PL + (size_t)kGenVox * CP);  // [64][3] centroids
  int* rs = reinterpret_cast<int*>(cen + 3 * kGenVox);    // [65] first row of each voxel, relative to the tile

  const int t = threadIdx.x;
  const PT* __restrict__ xyz = static_cast<const PT*>(prob.row_xyz);
  const long long n_chunks = *prob.n_chunks;
  float* hrow = H + (size_t)t * HS;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int nt = prob.chunk_ntiles[c];
    for (int j = 0; j < nt; ++j) {
      const int v0 = prob.tile_first[c * kChunkSlots + j], v1 = prob.tile_first[c * kChunkSlots + j + 1];
      const int r0 = prob.tile_row0[c * kChunkSlots + j], r1 = prob.tile_row0[c * kChunkSlots + j + 1];
      const int nvox = v1 - v0, nrows = r1 - r0;
      __syncthreads();  // the previous tile's readers are done with rs / cen / PL / H
      if (t <= nvox) rs[t] = prob.row_start[v0 + t] - r0;
      if (t < nvox) {  // np.mean(currPoints, axis=0): float64 adds in list order, one divide (model_training.py:135)
        const int a = prob.row_start[v0 + t], e = prob.row_start[v0 + t + 1];
        const int n = e - a - ((prob.row_voxel[e - 1] & kRowPadFlag) ? 1 : 0);
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int r = a; r < a + n; ++r) {
          sx += (double)xyz[3 * (size_t)r];
          sy += (double)xyz[3 * (size_t)r + 1];
          sz += (double)xyz[3 * (size_t)r + 2];
        }
        const double dn = (double)n;
        cen[3 * t] = sx / dn;
        cen[3 * t + 1] = sy / dn;
        cen[3 * t + 2] = sz / dn;
      }
      __syncthreads();
      const bool live = t < nrows;
      int lv = 0;
      // ---- addVFELayer(in, 6, 2*C1) ----
      if (live) {
        const int rv = prob.row_voxel[r0 + t];
        lv = (rv & ~kRowPadFlag) - v0;
        float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // the pad row: the dense input's zero rows (:141-142)
        if (!(rv & kRowPadFlag)) {
          const size_t r = (size_t)(r0 + t);
          point_features((double)xyz[3 * r], (double)xyz[3 * r + 1], (double)xyz[3 * r + 2], cen[3 * lv], cen[3 * lv + 1],
                         cen[3 * lv + 2], f);
        }
        row_fcn<6, C1, POST>(f, params + L::w(0), params + L::a(0), params + L::b(0), params + L::d(0), hrow);
      }
      __syncthreads();
      pool_rows<C1, HS>(H, rs, nvox, t, PL, CP);
      __syncthreads();
      // ---- addVFELayer(., 2*C1, 2*C2) ----
      if (live) {
        float in[2 * C1];
        load_concat<C1>(PL + (size_t)lv * CP, hrow, in);
        row_fcn<2 * C1, C2, POST>(in, params + L::w(1), params + L::a(1), params + L::b(1), params + L::d(1), hrow);
      }
      __syncthreads();
      pool_rows<C2, HS>(H, rs, nvox, t, PL, CP);
      __syncthreads();
      // ---- addFCN(., 2*C2, C3) + MaxPoolingVFELayer(combine=True) ----
      if (live) {
        float in[2 * C2];
        load_concat<C2>(PL + (size_t)lv * CP, hrow, in);
        row_fcn<2 * C2, C3, POST>(in, params + L::w(2), params + L::a(2), params + L::b(2), params + L::d(2), hrow);
      }
      __syncthreads();
      pool_rows<C3, HS>(H, rs, nvox, t, voxel_feat + (size_t)v0 * C3, C3);
    }
  }
}

template <int C1, int C2, int C3, bool POST>
cudaError_t launch_one(const float* params, const VfeProblem& prob, float* voxel_feat, int sm_count, cudaStream_t st) {
  using L = GenericLayout<C1, C2, C3, POST>;
  int per_sm = (int)((220 * 1024) / L::smem);
  if (per_sm > 6) per_sm = 6;
  if (per_sm < 1) per_sm = 1;
  const unsigned blocks = (unsigned)(sm_count * per_sm);
  if (prob.pts_dtype == LISEC_F32) {
    auto k = vfe_generic_kernel<C1, C2, C3, POST, float>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem);
    if (e != cudaSuccess) return e;
    k<<<blocks, kGenThreads, L::smem, st>>>(params, prob, voxel_feat);
  } else {
    auto k = vfe_generic_kernel<C1, C2, C3, POST, double>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem);
    if (e != cudaSuccess) return e;
    k<<<blocks, kGenThreads, L::smem, st>>>(params, prob, voxel_feat);
  }
  return cudaGetLastError();
}

template <int C1, int C2, int C3, bool POST>
void pack_one(const GenericVfeWeights& w, float* out) {
  using L = GenericLayout<C1, C2, C3, POST>;
  for (int s = 0; s < 3; ++s) {
    const int cin = L::cin(s), cout = L::cout(s);
    for (int i = 0; i < cin * cout; ++i) out[L::w(s) + i] = w.dense[s][i];
    for (int j = 0; j < cout; ++j) {
      out[L::a(s) + j] = w.a[s][j];
      out[L::b(s) + j] = w.b[s][j];
    }
    if (POST)
      for (int i = 0; i < cout * cout; ++i) out[L::d(s) + i] = w.post[s][i];
  }
}

}  // namespace

#define LISEC_GENERIC_DISPATCH(c1, c2, c3, post, CALL) \
  do {                                                  \
    if (c1 == 16 && c2 == 32 && c3 == 64) {             \
      if (post) { CALL(16, 32, 64, true); } else { CALL(16, 32, 64, false); } \
    } else if (c1 == 16 && c2 == 64 && c3 == 128) {     \
      if (post) { CALL(16, 64, 128, true); } else { CALL(16, 64, 128, false); } \
    }                                                   \
  } while (0)

bool vfe_generic_supports(int c1, int c2, int c3) {
  return (c1 == 16 && c2 == 32 && c3 == 64) || (c1 == 16 && c2 == 64 && c3 == 128);
}

size_t vfe_generic_param_floats(int c1, int c2, int c3, bool post) {
#define LISEC_CALL(A, B, C, P) return (size_t)GenericLayout<A, B, C, P>::total
  LISEC_GENERIC_DISPATCH(c1, c2, c3, post, LISEC_CALL);
#undef LISEC_CALL
  return 0;
}

void vfe_generic_pack(int c1, int c2, int c3, bool post, const GenericVfeWeights& w, float* out) {
#define LISEC_CALL(A, B, C, P) pack_one<A, B, C, P>(w, out); return
  LISEC_GENERIC_DISPATCH(c1, c2, c3, post, LISEC_CALL);
#undef LISEC_CALL
}

cudaError_t launch_vfe_generic(int c1, int c2, int c3, bool post, const float* params, const VfeProblem& prob,
                               float* voxel_feat, int sm_count, cudaStream_t st, int* launches) {
  ++*launches;
#define LISEC_CALL(A, B, C, P) return launch_one<A, B, C, P>(params, prob, voxel_feat, sm_count, st)
  LISEC_GENERIC_DISPATCH(c1, c2, c3, post, LISEC_CALL);
#undef LISEC_CALL
  return cudaErrorInvalidValue;
}

}  // namespace lisec
