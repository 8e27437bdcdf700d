// The VFE stack for EVERY graph the reference's .h5 may hold — parametric in the widths (C1, C2, C3) and in
// the FCN variant — in plain float32 FMAs:
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
// Per tile, 512 threads around one shared-memory tile [128 rows][<= 128 channels (+4)]. A Dense layer is a register-tiled
// GEMM on it: thread (row group, column group) owns 4 rows x COUT/16 columns, reads its rows' inputs as 16-byte pieces
// from the tile and the weights as 16-byte read-only loads (the whole weight set, <= 157 KB, lives in L1/L2), and writes
// the outputs back into the tile behind a barrier — at the column offset that leaves room for the pooled half, so that
// Concatenate([pooling, layer]) is a layout, not a copy. The per-voxel max (one thread per voxel x 4 channels) then
// writes the pooled half of every row of the voxel (RepeatLayer). Bound by the FP32 pipe: 39 264 MACs per row for
// (16, 64, 128, POST), 5 216 for (16, 32, 64). (First version: one thread per row with the input row in registers and one
// weight load per FMA — 12 TFLOP/s, bound by the L1 data pipe: profiles/vfe_generic_r2_summary.txt.)
#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

namespace {

constexpr int kGenThreads = 512;
constexpr int kGenRows = 128;  // rows per tile
constexpr int kGenVox = 64;    // voxels per tile

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
  static constexpr int kCW = 2 * C2 > C3 ? 2 * C2 : C3;  // widest row the tile holds: a concat input or the last output
  static constexpr int kXS = kCW + 4;                     // row stride in floats: 16-byte rows, conflict-free
  static constexpr size_t smem = sizeof(float) * (size_t)kGenRows * kXS + sizeof(double) * 3 * kGenVox +
                                 sizeof(int) * (kGenVox + 4 + kGenRows);
};

// T[row][out_col + c] = act(affine(sum_k T[row][in_col + k] * W[k][c])) for the tile's 128 rows, in place: every thread
// finishes reading its input rows before anybody writes (the barrier in the middle), and the outputs are visible to all
// when the function returns. A warp's 32 lanes are 32 column groups (16 for the 16-column first layer) and its rows are
// the same for every lane: the input loads are pure broadcasts, a weight load covers 32 x CT contiguous floats — the L1
// data pipe sees ~0.2 wavefronts per FMA instruction (the 4-row x 8-column tile of the previous version: ~0.6, and it
// was what bounded the kernel, profiles/vfe_generic_r2_summary.txt). Thread = RT rows x CT columns.
template <int CIN, int COUT, int XS, bool AFFINE, bool RELU>
__device__ __forceinline__ void tile_dense(float* __restrict__ T, int in_col, int out_col, const float* __restrict__ W,
                                           const float* __restrict__ a, const float* __restrict__ b, int tid) {
  constexpr int CG = COUT >= 64 ? 32 : 16;        // column groups = lanes that differ in their columns (narrow layers: 16, two row groups per warp)
  constexpr int CT = COUT / CG;                   // columns per thread: 4, 2 or 1
  constexpr int RT = kGenRows / (kGenThreads / CG);  // rows per thread: 8 (4 for the first layer)
  const int rg = tid / CG, cg = tid % CG;
  float acc[RT][CT];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[r][c] = 0.f;
  const float* x0 = T + (size_t)(RT * rg) * XS + in_col;
  const float* wc = W + cg * CT;
  auto fma_k = [&](const float (&xv)[RT], int k) {  // one input channel: CT weights, RT x CT FMAs
    float w[CT];
    if (CT == 4) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + (size_t)k * COUT));
      w[0] = w0.x; w[1 % CT] = w0.y; w[2 % CT] = w0.z; w[3 % CT] = w0.w;  // (% CT: in bounds in the other instantiations)
    } else if (CT == 2) {
      const float2 w0 = __ldg(reinterpret_cast<const float2*>(wc + (size_t)k * COUT));
      w[0] = w0.x; w[1 % CT] = w0.y;
    } else {
      w[0] = __ldg(wc + (size_t)k * COUT);
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < CT; ++c) acc[r][c] = fmaf(xv[r], w[c], acc[r][c]);
  };
  if (CIN % 4 == 0) {
#pragma unroll 2
    for (int k = 0; k < CIN; k += 4) {
      float4 x[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) x[r] = *reinterpret_cast<const float4*>(x0 + (size_t)r * XS + k);
      float xv[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) xv[r] = x[r].x;
      fma_k(xv, k);
#pragma unroll
      for (int r = 0; r < RT; ++r) xv[r] = x[r].y;
      fma_k(xv, k + 1);
#pragma unroll
      for (int r = 0; r < RT; ++r) xv[r] = x[r].z;
      fma_k(xv, k + 2);
#pragma unroll
      for (int r = 0; r < RT; ++r) xv[r] = x[r].w;
      fma_k(xv, k + 3);
    }
  } else {  // the first layer's six input features
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
      float xv[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) xv[r] = x0[(size_t)r * XS + k];
      fma_k(xv, k);
    }
  }
  __syncthreads();
  float* y0 = T + (size_t)(RT * rg) * XS + out_col + cg * CT;
  float sa[CT], sb[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) {  // BatchNormalization at inference, folded on the host: y = z * a + b
    sa[c] = AFFINE ? __ldg(a + cg * CT + c) : 1.f;
    sb[c] = AFFINE ? __ldg(b + cg * CT + c) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    float v[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      v[c] = AFFINE ? fmaf(acc[r][c], sa[c], sb[c]) : acc[r][c];
      if (RELU) v[c] = fmaxf(v[c], 0.f);
    }
    float* y = y0 + (size_t)r * XS;
    if (CT == 4) *reinterpret_cast<float4*>(y) = make_float4(v[0], v[1 % CT], v[2 % CT], v[3 % CT]);
    else if (CT == 2) *reinterpret_cast<float2*>(y) = make_float2(v[0], v[1 % CT]);
    else y[0] = v[0];
  }
  __syncthreads();
}

// one FCN on the tile: Dense -> BN -> ReLU, or Dense -> BN -> Dense -> ReLU; input at columns [in_col, in_col + CIN),
// result at [out_col, out_col + COUT)
template <int CIN, int COUT, int XS, bool POST>
__device__ __forceinline__ void tile_fcn(float* __restrict__ T, int in_col, int out_col, const float* __restrict__ W,
                                         const float* __restrict__ a, const float* __restrict__ b,
                                         const float* __restrict__ D, int tid) {
  tile_dense<CIN, COUT, XS, true, !POST>(T, in_col, out_col, W, a, b, tid);
  if (POST) tile_dense<COUT, COUT, XS, false, true>(T, out_col, out_col, D, nullptr, nullptr, tid);
}

// MaxPoolingVFELayer: per-voxel max over the voxel's rows of the tile (kept rows and, for a non-full voxel, its pad row)
// of columns [col, col + C). REPEAT: RepeatLayer + the pooled half of Concatenate — the max goes to columns [0, C) of
// every row of the voxel; otherwise it is the voxel's output row (MaxPoolingVFELayer(combine=True)) in `dst`.
template <int C, int XS, bool REPEAT>
__device__ __forceinline__ void pool_rows(float* __restrict__ T, int col, const int* __restrict__ rs, int nvox, int tid,
                                          float* __restrict__ dst) {
  constexpr int G = C / 4;
  for (int item = tid; item < nvox * G; item += kGenThreads) {
    const int v = item / G, g = item - v * G;
    const int r0 = rs[v], r1 = rs[v + 1];
    float4 m = *reinterpret_cast<const float4*>(T + (size_t)r0 * XS + col + 4 * g);
    for (int r = r0 + 1; r < r1; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(T + (size_t)r * XS + col + 4 * g);
      m.x = fmaxf(m.x, x.x);
      m.y = fmaxf(m.y, x.y);
      m.z = fmaxf(m.z, x.z);
      m.w = fmaxf(m.w, x.w);
    }
    if (REPEAT) {
      for (int r = r0; r < r1; ++r) *reinterpret_cast<float4*>(T + (size_t)r * XS + 4 * g) = m;
    } else {
      *reinterpret_cast<float4*>(dst + (size_t)v * C + 4 * g) = m;
    }
  }
}

template <int C1, int C2, int C3, bool POST, typename PT>
__global__ void __launch_bounds__(kGenThreads) vfe_generic_kernel(const float* __restrict__ params, const VfeProblem prob,
                                                                  float* __restrict__ voxel_feat) {
  using L = GenericLayout<C1, C2, C3, POST>;
  constexpr int XS = L::kXS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* T = reinterpret_cast<float*>(smem_raw);                          // [128][XS] the tile
  double* cen = reinterpret_cast<double*>(T + (size_t)kGenRows * XS);     // [64][3] centroids
  int* rs = reinterpret_cast<int*>(cen + 3 * kGenVox);                    // [65] first row of each voxel, tile-relative

  const int t = threadIdx.x;
  const PT* __restrict__ xyz = static_cast<const PT*>(prob.row_xyz);
  const long long n_chunks = __ldcg(prob.n_chunks);  // (written by scan_down of this call: through L2)
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int nt = prob.chunk_ntiles[c];
    for (int j = 0; j < nt; ++j) {
      const int v0 = prob.tile_first[c * kChunkSlots + j], v1 = prob.tile_first[c * kChunkSlots + j + 1];
      const int r0 = prob.tile_row0[c * kChunkSlots + j], r1 = prob.tile_row0[c * kChunkSlots + j + 1];
      const int nvox = v1 - v0, nrows = r1 - r0;
      __syncthreads();  // the previous tile's readers are done with rs / cen / T
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
      if (t < kGenRows) {  // the tile's input rows: columns [0, 6); rows past the tile's end stay zero (never pooled)
        float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // the pad row: the dense input's zero rows (:141-142)
        if (t < nrows) {
          const int rv = prob.row_voxel[r0 + t];
          if (!(rv & kRowPadFlag)) {
            const int lv = rv - v0;
            const size_t r = (size_t)(r0 + t);
            point_features((double)xyz[3 * r], (double)xyz[3 * r + 1], (double)xyz[3 * r + 2], cen[3 * lv], cen[3 * lv + 1],
                           cen[3 * lv + 2], f);
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) T[(size_t)t * XS + k] = f[k];
      }
      __syncthreads();
      // ---- addVFELayer(in, 6, 2*C1): the FCN's output at columns [C1, 2 C1), the pooled half in front of it ----
      tile_fcn<6, C1, XS, POST>(T, 0, C1, params + L::w(0), params + L::a(0), params + L::b(0), params + L::d(0), t);
      pool_rows<C1, XS, true>(T, C1, rs, nvox, t, nullptr);
      __syncthreads();
      // ---- addVFELayer(., 2*C1, 2*C2) ----
      tile_fcn<2 * C1, C2, XS, POST>(T, 0, C2, params + L::w(1), params + L::a(1), params + L::b(1), params + L::d(1), t);
      pool_rows<C2, XS, true>(T, C2, rs, nvox, t, nullptr);
      __syncthreads();
      // ---- addFCN(., 2*C2, C3) + MaxPoolingVFELayer(combine=True) ----
      tile_fcn<2 * C2, C3, XS, POST>(T, 0, 0, params + L::w(2), params + L::a(2), params + L::b(2), params + L::d(2), t);
      pool_rows<C3, XS, false>(T, 0, rs, nvox, t, voxel_feat + (size_t)v0 * C3);
    }
  }
}

template <typename K>
cudaError_t launch_kernel(K k, size_t smem, const float* params, const VfeProblem& prob, float* voxel_feat, int sm_count,
                          cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 1;  // resident CTAs per SM (registers: 512 threads x ~120 allow one for the wide graphs, two for the narrow)
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kGenThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  k<<<(unsigned)(sm_count * per_sm), kGenThreads, smem, st>>>(params, prob, voxel_feat);
  return cudaGetLastError();
}

template <int C1, int C2, int C3, bool POST>
cudaError_t launch_one(const float* params, const VfeProblem& prob, float* voxel_feat, int sm_count, cudaStream_t st) {
  using L = GenericLayout<C1, C2, C3, POST>;
  if (prob.pts_dtype == LISEC_F32)
    return launch_kernel(vfe_generic_kernel<C1, C2, C3, POST, float>, L::smem, params, prob, voxel_feat, sm_count, st);
  return launch_kernel(vfe_generic_kernel<C1, C2, C3, POST, double>, L::smem, params, prob, voxel_feat, sm_count, st);
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
