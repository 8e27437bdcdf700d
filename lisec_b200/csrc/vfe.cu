// Stacked VFE on sm_100a: centroid augmentation, three pointwise linears (+BN+ReLU), two per-voxel max-pools with
// concat, and the final max over T — reference model_training.py:134-141 (features) and :155-186, 229-235 (layers).
//
// Work unit: a tile = a run of whole voxels holding at most 256 VFE rows (a row is a kept point, or the single
// virtual zero row that stands for all identical pad rows of a non-full voxel). One 256-thread CTA
// per tile, persistent over tiles. Inside a tile every Dense is a small GEMM on on-chip data:
//
//     rows    H_next[256 x N] = H[256 x K] * Wx[K x N]  (+ Q[voxel(row)])       8x8 / 8x4 register tile per thread
//     voxels  Q[128 x N]      = Pool[128 x K] * Wp[K x N]                        4x8 / 4x4 register tile per thread
//
// with the activations k-major in shared memory (A operand: 16-byte loads of 4 consecutive rows) and the weights in
// shared memory ([K][N], 16-byte loads of 4 consecutive columns). This is the classic SIMT SGEMM inner loop:
// 4 LDS.128 feed 64 FFMA, measured at 54 TFLOP/s on B200 against a 58.7 TFLOP/s FP32-pipe peak
// (tools/ffma_probe.cu, tools/fp32_peak.cu); the earlier thread-per-row form topped out at 39.
// Because Concatenate([pooled, pointwise]) feeds a bias-free Dense (model_training.py:164-165, 184), the pooled half
// of the next product is the same for every row of a voxel: it is computed once per voxel (Q) and used as the
// accumulators' initial value. Thread mapping: lane = 8 consecutive tile rows, warp = column group, so a voxel's rows
// all sit in one warp and every max-pool is an in-register segmented max plus a 4-step segmented shuffle scan — no
// shared-memory pooling passes, no atomics (ncu on the previous version: 48 % of the time in smem pooling loops).
//
// Precision (parity bar 1e-5 against the float64 oracle): dense (6->16) acts on raw coordinates up to +-50 m and is
// accumulated in float64 (96 DFMA per row); dense_1's two halves cancel, so its 32-term sum is accumulated in
// float32 blocks of 4 (a fresh accumulator per block, then added) — measured worst case 4.7e-6 over seeds and clouds,
// the level of a CPU float32 forward; dense_2 is a plain float32 FMA chain.
#include <cuda_bf16.h>

#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

int vfe_rows_per_tile(int T) { return kVfeThreads - T + 1; }

namespace {

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&h);
}

constexpr int kRows = kVfeThreads;   // 256 rows per tile
constexpr int kVox = kVfeThreads / 2;  // 128 voxels per tile: a non-full voxel has >= 2 rows, a full one T >= 2
constexpr int PR = kRows + 4;        // float pitch of row-indexed k-major tiles (16-byte aligned rows, 4-bank skew)
constexpr int PV = kVox + 4;         // float pitch of voxel-indexed k-major tiles
constexpr int QS = 68;               // float stride of a voxel's row in sQ

// shared-memory map (bytes)
constexpr int OFF_W2P = 0;                         // [16][32]
constexpr int OFF_W2X = OFF_W2P + 16 * 32 * 4;     // [16][32]
constexpr int OFF_W3P = OFF_W2X + 16 * 32 * 4;     // [32][64]
constexpr int OFF_W3X = OFF_W3P + 32 * 64 * 4;     // [32][64]
constexpr int OFF_H1T = OFF_W3X + 32 * 64 * 4;     // [16][PR]   dense outputs of VFE-1, k-major
constexpr int OFF_P1T = OFF_H1T + 16 * PR * 4;     // [16][PV]   pooled VFE-1
constexpr int OFF_P2T = OFF_H1T;                   // [32][PV]   pooled VFE-2, reuses H1T+P1T (dead by then)
constexpr int OFF_H2T = OFF_P1T + 16 * PV * 4;     // [32][PR]   VFE-2 outputs; later one 32-channel half of FCN outputs
constexpr int OFF_Q = OFF_H2T + 32 * PR * 4;       // [kVox][QS] pooled-half products of the current layer
constexpr int OFF_ROWVOX = OFF_Q + kVox * QS * 4;  // uint8[kRows] local voxel of each tile row
constexpr int OFF_VOXCELL = OFF_ROWVOX + kRows;      // int[kVox] cell of each tile voxel (grid output modes)
constexpr int kSmemBytes = OFF_VOXCELL + kVox * 4;
static_assert(32 * PV * 4 <= 16 * PR * 4 + 16 * PV * 4, "P2T must fit in H1T+P1T");
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");

// ---- register-tile GEMM: acc[R][C] += A[k][row(r)] * W[k][col(c)], k = 0..K-1 -------------------------------
// The lane's rows come as R/4 float4 chunks at row0 + i*chunk_stride (consecutive lanes -> consecutive 16 bytes);
// the warp's columns come in groups of 4 at offsets coff[g] (same address for every lane: a broadcast load).
// BLOCK4: float32 accumulation in blocks of 4 k-steps (see the precision note above).
template <int R, int CG, int K, bool BLOCK4>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ sA, int pitch, int chunk_stride, int row0,
                                          const float* __restrict__ sW, int ldw, const int (&coff)[CG],
                                          float (&acc)[R][CG * 4]) {
  static_assert(R % 4 == 0 && K % 4 == 0, "tile shape");
  if (BLOCK4) {
#pragma unroll 1
    for (int kb = 0; kb < K; kb += 4) {
      float blk[R][CG * 4];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CG * 4; ++c) blk[r][c] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = kb + kk;
        float a[R], b[CG * 4];
#pragma unroll
        for (int i = 0; i < R / 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(sA + k * pitch + row0 + chunk_stride * i);
          a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < CG; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(sW + k * ldw + coff[g]);
          b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < CG * 4; ++c) blk[r][c] = fmaf(a[r], b[c], blk[r][c]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CG * 4; ++c) acc[r][c] += blk[r][c];
    }
  } else {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float a[R], b[CG * 4];
#pragma unroll
      for (int i = 0; i < R / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(sA + k * pitch + row0 + chunk_stride * i);
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(sW + k * ldw + coff[g]);
        b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CG * 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
  }
}

// barrier among the 8 compute warps only (the 9th warp of the fused kernel streams the grid background and never joins)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kVfeThreads) : "memory"); }

// position of tile row r inside a k-major row tile: lane l = r/8 owns rows 8l..8l+7 and loads them as two float4 at
// 4l and 128+4l, so both 16-byte loads of a warp cover 512 contiguous bytes (no bank conflicts)
__device__ __forceinline__ int row_pos(int r) { return ((r >> 3) << 2) + (r & 3) + ((r & 4) << 5); }

// ---- per-voxel max in registers ----------------------------------------------------------------------------
// A lane holds 8 consecutive tile rows, a warp all 256 of them, so every voxel (a run of consecutive rows) lives in
// one warp. Column-independent bookkeeping, computed once per tile:
struct PoolMeta {
  int v[8];          // local voxel of each of the lane's rows (255 = padding row past the tile's last row)
  unsigned bnd;      // bit r (1..7): row r starts a new voxel inside this lane
  int kh, kt;        // voxel of the first / last row
  bool cont;         // the first row's voxel continues from the previous lane
  bool emit_head;    // the first row's voxel ends inside this lane
  bool emit_tail;    // the last row's voxel starts and ends inside this lane (and is not the first row's voxel)
  bool take[4];      // segmented-scan schedule over lanes, distances 1,2,4,8 (a voxel spans at most 9 lanes for T<=64)
};

__device__ __forceinline__ PoolMeta make_pool_meta(const unsigned char* __restrict__ sRowVox, int lane) {
  PoolMeta m;
  const uint2 rv = *reinterpret_cast<const uint2*>(sRowVox + lane * 8);
  m.bnd = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    m.v[r] = ((r < 4 ? rv.x : rv.y) >> (8 * (r & 3))) & 0xff;
    if (r > 0 && m.v[r] != m.v[r - 1]) m.bnd |= 1u << r;
  }
  m.kh = m.v[0];
  m.kt = m.v[7];
  const bool whole = m.bnd == 0;
  const int prev_kt = __shfl_up_sync(0xffffffffu, m.kt, 1);
  const int next_kh = __shfl_down_sync(0xffffffffu, m.kh, 1);
  m.cont = lane > 0 && prev_kt == m.kh;
  const bool tail_cont = lane < 31 && next_kh == m.kt;
  m.emit_head = whole ? !tail_cont : true;
  m.emit_tail = !whole && !tail_cont;
  // inclusive segmented max-scan over lanes of the "last voxel of the lane" values; a lane extends the run of its
  // predecessor iff it is one whole voxel that continues from it
  bool flag = !(whole && m.cont);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const bool up = __shfl_up_sync(0xffffffffu, flag, 1 << s);
    m.take[s] = lane >= (1 << s) && !flag;
    if (m.take[s]) flag = up;
  }
  return m;
}

// val[r][c]: the lane's 8 rows x NC columns. emit(voxel, values[NC]) is called exactly once per voxel, by the lane
// in which the voxel ends, with the max over all of the voxel's rows for the lane's NC columns.
template <int NC, typename Emit>
__device__ __forceinline__ void pool_lane_rows(const float (&val)[8][NC], const PoolMeta& m, Emit emit) {
  float run[NC], head[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) run[c] = val[0][c];
  bool first = true;
#pragma unroll
  for (int r = 1; r < 8; ++r) {
    if (m.bnd & (1u << r)) {
      if (first) {
#pragma unroll
        for (int c = 0; c < NC; ++c) head[c] = run[c];
        first = false;
      } else {  // a voxel that starts and ends inside the lane
        emit(m.v[r - 1], run);
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] = val[r][c];
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) run[c] = fmaxf(run[c], val[r][c]);
    }
  }
  if (first) {
#pragma unroll
    for (int c = 0; c < NC; ++c) head[c] = run[c];
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float x = run[c];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float y = __shfl_up_sync(0xffffffffu, x, 1 << s);
      if (m.take[s]) x = fmaxf(x, y);
    }
    const float in = __shfl_up_sync(0xffffffffu, x, 1);
    if (m.cont) head[c] = fmaxf(in, head[c]);
  }
  if (m.emit_head) emit(m.kh, head);
  if (m.emit_tail) emit(m.kt, run);
}

// float64 mean of each voxel's kept points, added in list order with one divide — np.mean(currPoints, axis=0)
// (model_training.py:135) bit for bit. One thread per voxel; a pre-pass so that VFE tiles start without a barrier.
template <typename PT>
__global__ void __launch_bounds__(256) centroid_kernel(const PT* __restrict__ pts, int T,
                                                       const int* __restrict__ voxel_start,
                                                       const int* __restrict__ list_sorted,
                                                       const long long* __restrict__ totals,
                                                       double* __restrict__ centroid) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= totals[TOT_VOXELS]) return;
  const int s = voxel_start[v];
  const int c = voxel_start[v + 1] - s;
  const int kept = c < T ? c : T;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int i = 0; i < kept; ++i) {
    PT x, y, z;
    load_point(pts, (long long)list_sorted[s + i], x, y, z);
    sx += (double)x;
    sy += (double)y;
    sz += (double)z;
  }
  const double n = (double)(kept > 0 ? kept : 1);
  centroid[3 * v] = sx / n;
  centroid[3 * v + 1] = sy / n;
  centroid[3 * v + 2] = sz / n;
}

// ---- background writer (fused kernel, warp 8 of every CTA) --------------------------------------------------
// Streams c_empty into every EMPTY cell of the grid while the compute warps are busy on the FP32 pipe; occupied cells
// are written by the compute warps (their voxel rows), so every grid element is still written exactly once.
// 32 cells per step: one coalesced 128-byte read of the occupancy map, then 16-byte streaming stores.
template <typename GT>
__device__ __forceinline__ void background_writer(const int* __restrict__ cell_voxel, const float* __restrict__ c_empty,
                                                  GT* __restrict__ grid, long long ncells, int lane) {
  constexpr int LPC = sizeof(GT) == 4 ? 16 : 8;  // lanes per cell: 64 channels x sizeof(GT) / 16 bytes
  constexpr int CPS = 32 / LPC;                  // cells per store instruction
  const int sub = lane / LPC, chunk = lane % LPC;
  uint4 bg;
  if (sizeof(GT) == 4) {
    bg = reinterpret_cast<const uint4*>(c_empty)[chunk];
  } else {
    const float4 b0 = reinterpret_cast<const float4*>(c_empty)[2 * chunk];
    const float4 b1 = reinterpret_cast<const float4*>(c_empty)[2 * chunk + 1];
    bg = make_uint4(pack_bf16x2(b0.x, b0.y), pack_bf16x2(b0.z, b0.w), pack_bf16x2(b1.x, b1.y), pack_bf16x2(b1.z, b1.w));
  }
  const long long ngroups = (ncells + 31) >> 5;
  constexpr int U = 4;  // occupancy-map loads kept in flight
  for (long long g0 = blockIdx.x; g0 < ngroups; g0 += (long long)gridDim.x * U) {
    int occ[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long cell = ((g0 + (long long)u * gridDim.x) << 5) + lane;
      occ[u] = (g0 + (long long)u * gridDim.x < ngroups && cell < ncells) ? __ldg(cell_voxel + cell) : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long base = (g0 + (long long)u * gridDim.x) << 5;
#pragma unroll
      for (int i = 0; i < 32 / CPS; ++i) {
        const int vox = __shfl_sync(0xffffffffu, occ[u], CPS * i + sub);
        if (vox < 0) __stcs(reinterpret_cast<uint4*>(grid + (base + CPS * i + sub) * 64) + chunk, bg);
      }
    }
  }
}

// MODE 0: voxel rows to voxel_feat[V][64] (float32), no background.  MODE 1 / 2: rows straight into the float32 /
// bf16 dense grid at their cell, background by the 9th warp.
struct VfeOutput {
  float* voxel_feat;
  void* grid;
  const int* voxel_cell;  // voxel row -> cell (sweep * cells + (z*nx + x)*ny + y)
  const int* cell_voxel;  // occupancy map
  const float* c_empty;
  long long ncells;
};

template <typename PT, int MODE>
__global__ void __launch_bounds__(kVfeThreads + 32, 2)
    vfe_kernel(const PT* __restrict__ pts, const __grid_constant__ VfeSmall P, const float* __restrict__ wblob,
               const __grid_constant__ VfeProblem prob, const __grid_constant__ VfeOutput out) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (threadIdx.x >= kVfeThreads) {  // warp 8
    if (MODE == 1) background_writer(out.cell_voxel, out.c_empty, static_cast<float*>(out.grid), out.ncells, threadIdx.x & 31);
    if (MODE == 2) background_writer(out.cell_voxel, out.c_empty, static_cast<__nv_bfloat16*>(out.grid), out.ncells, threadIdx.x & 31);
    return;
  }
  float* sW2P = reinterpret_cast<float*>(smem + OFF_W2P);
  float* sW2X = reinterpret_cast<float*>(smem + OFF_W2X);
  float* sW3P = reinterpret_cast<float*>(smem + OFF_W3P);
  float* sW3X = reinterpret_cast<float*>(smem + OFF_W3X);
  float* sH1T = reinterpret_cast<float*>(smem + OFF_H1T);
  float* sP1T = reinterpret_cast<float*>(smem + OFF_P1T);
  float* sP2T = reinterpret_cast<float*>(smem + OFF_P2T);
  float* sH2T = reinterpret_cast<float*>(smem + OFF_H2T);
  float* sQ = reinterpret_cast<float*>(smem + OFF_Q);
  unsigned char* sRowVox = smem + OFF_ROWVOX;
  int* sVoxCell = reinterpret_cast<int*>(smem + OFF_VOXCELL);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // register tiles: lane = row group (8 tile rows / 4 voxel rows), warp = column group
  const int n_tiles = (int)*prob.n_tiles;
  if ((int)blockIdx.x >= n_tiles) return;

  // weights: [W2P | W2X | W3P | W3X] as laid out by the host, straight into the first 20 KB
  for (int i = tid; i < (OFF_H1T / 16); i += kVfeThreads)
    reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(wblob) + i);

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int v0 = prob.tile_first[t], v1 = prob.tile_first[t + 1];
    const int nv = v1 - v0;
    const int r0 = prob.row_start[v0];
    const int nrows = prob.row_start[v1] - r0;
    const bool has_row = tid < nrows;

    // ---- VFE-1: Dense(6->16, no bias) + BN + ReLU (addVFELayer(in, 6, 32), :231 -> :155-166); one row per thread ----
    {
      int p = -1, lv = 255;  // 255 = padding row past the tile's last row
      if (has_row) {
        p = prob.row_point[r0 + tid];
        lv = prob.row_voxel[r0 + tid] - v0;
      }
      float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // pad row: six zeros (:141)
      if (p >= 0) {
        PT px, py, pz;
        load_point(pts, (long long)p, px, py, pz);
        const double* c = prob.centroid + 3 * (size_t)(v0 + lv);
        point_features((double)px, (double)py, (double)pz, c[0], c[1], c[2], f);
      }
      double d[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) d[j] = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double fk = (double)f[k];
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = fma(fk, P.w1[k][j], d[j]);
      }
      const int pos = row_pos(tid);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        sH1T[j * PR + pos] = has_row ? fmaxf(fmaf(__double2float_rn(d[j]), P.a1[j], P.b1[j]), 0.f) : 0.f;
      sRowVox[tid] = (unsigned char)lv;
      if (MODE != 0 && tid < nv) sVoxCell[tid] = out.voxel_cell[v0 + tid];
    }
    compute_sync();
    const PoolMeta meta = make_pool_meta(sRowVox, lane);
    {  // MaxPoolingVFELayer over T (:160); RepeatLayer is implicit. Warp w pools channels 2w, 2w+1.
      float val[8][2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float* src = sH1T + (2 * warp + c) * PR;
        const float4 lo = *reinterpret_cast<const float4*>(src + 4 * lane);
        const float4 hi = *reinterpret_cast<const float4*>(src + 128 + 4 * lane);
        val[0][c] = lo.x; val[1][c] = lo.y; val[2][c] = lo.z; val[3][c] = lo.w;
        val[4][c] = hi.x; val[5][c] = hi.y; val[6][c] = hi.z; val[7][c] = hi.w;
      }
      pool_lane_rows<2>(val, meta, [&](int v, const float(&x)[2]) {
        if (v < nv) {
          sP1T[(2 * warp) * PV + v] = x[0];
          sP1T[(2 * warp + 1) * PV + v] = x[1];
        }
      });
    }
    compute_sync();

    // ---- VFE-2: Dense(32->32) + BN + ReLU on concat[pooled, pointwise] (addVFELayer(., 32, 64), :232) ----
    const int coff4[1] = {warp * 4};
    {  // pooled half, once per voxel: Q2[128 x 32] = P1[128 x 16] * W2p; 4x4 tile per thread
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      tile_gemm<4, 1, 16, true>(sP1T, PV, 0, lane * 4, sW2P, 32, coff4, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(sQ + (lane * 4 + r) * QS + warp * 4) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    compute_sync();
    {  // rows: 8x4 tile per thread, accumulators start at the voxel's pooled-half product
      float acc[8][4];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 q = *reinterpret_cast<const float4*>(sQ + (meta.v[r] & (kVox - 1)) * QS + warp * 4);
        acc[r][0] = q.x; acc[r][1] = q.y; acc[r][2] = q.z; acc[r][3] = q.w;
      }
      tile_gemm<8, 1, 16, true>(sH1T, PR, 128, lane * 4, sW2X, 32, coff4, acc);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float a = P.a2[warp * 4 + c], b = P.b2[warp * 4 + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r][c] = fmaxf(fmaf(acc[r][c], a, b), 0.f);
        float* dst = sH2T + (warp * 4 + c) * PR + 4 * lane;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        *reinterpret_cast<float4*>(dst + 128) = make_float4(acc[4][c], acc[5][c], acc[6][c], acc[7][c]);
      }
      compute_sync();  // every warp is done with sH1T (A operand) and sQ: sP2T may now overwrite sH1T/sP1T
      pool_lane_rows<4>(acc, meta, [&](int v, const float(&x)[4]) {
        if (v < nv) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sP2T[(warp * 4 + c) * PV + v] = x[c];
        }
      });
    }
    compute_sync();

    // ---- FCN: Dense(64->64) + BN + ReLU (addFCN(., 64, 64), :233), then MaxPoolingVFELayer(combine=True) (:235) ----
    const int coff8[2] = {warp * 4, 32 + warp * 4};  // this warp's 8 output channels
    {  // pooled half: Q3[128 x 64] = P2[128 x 32] * W3p; 4x8 tile per thread
      float acc[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      tile_gemm<4, 2, 32, false>(sP2T, PV, 0, lane * 4, sW3P, 64, coff8, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float* dst = sQ + (lane * 4 + r) * QS;
        *reinterpret_cast<float4*>(dst + coff8[0]) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(dst + coff8[1]) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
      }
    }
    compute_sync();
    {  // rows: 8x8 tile per thread; the per-voxel max goes straight from registers to the output row
      float out8[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float* q = sQ + (meta.v[r] & (kVox - 1)) * QS;
        const float4 q0 = *reinterpret_cast<const float4*>(q + coff8[0]);
        const float4 q1 = *reinterpret_cast<const float4*>(q + coff8[1]);
        out8[r][0] = q0.x; out8[r][1] = q0.y; out8[r][2] = q0.z; out8[r][3] = q0.w;
        out8[r][4] = q1.x; out8[r][5] = q1.y; out8[r][6] = q1.z; out8[r][7] = q1.w;
      }
      tile_gemm<8, 2, 32, false>(sH2T, PR, 128, lane * 4, sW3X, 64, coff8, out8);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = coff8[c >> 2] + (c & 3);
        const float a = P.a3[col], b = P.b3[col];
#pragma unroll
        for (int r = 0; r < 8; ++r) out8[r][c] = fmaxf(fmaf(out8[r][c], a, b), 0.f);
      }
      pool_lane_rows<8>(out8, meta, [&](int v, const float(&x)[8]) {
        if (v < nv) {  // two 16-byte (bf16: 8-byte) stores per voxel and warp; the 8 warps complete the row
          if (MODE == 0) {
            float* dst = out.voxel_feat + (size_t)(v0 + v) * 64;
            *reinterpret_cast<float4*>(dst + coff8[0]) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(dst + coff8[1]) = make_float4(x[4], x[5], x[6], x[7]);
          } else if (MODE == 1) {
            float* dst = static_cast<float*>(out.grid) + (size_t)sVoxCell[v] * 64;
            __stcs(reinterpret_cast<float4*>(dst + coff8[0]), make_float4(x[0], x[1], x[2], x[3]));
            __stcs(reinterpret_cast<float4*>(dst + coff8[1]), make_float4(x[4], x[5], x[6], x[7]));
          } else {
            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(out.grid) + (size_t)sVoxCell[v] * 64;
            *reinterpret_cast<uint2*>(dst + coff8[0]) = make_uint2(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]));
            *reinterpret_cast<uint2*>(dst + coff8[1]) = make_uint2(pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
          }
        }
      });
    }
    compute_sync();  // the next tile's setup overwrites sH1T and sRowVox
  }
}

}  // namespace

cudaError_t launch_centroids(const void* pts, int pts_dtype, const Geom& g, const Workspace& w, long long max_voxels,
                             cudaStream_t st, int* launches) {
  const unsigned blocks = (unsigned)((max_voxels + 255) / 256);  // threads past the device-side voxel count exit
  if (pts_dtype == LISEC_F32)
    centroid_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(pts), g.T, w.voxel_start, w.list_sorted,
                                                   w.totals, w.centroid);
  else
    centroid_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(pts), g.T, w.voxel_start,
                                                    w.list_sorted, w.totals, w.centroid);
  ++*launches;
  return cudaGetLastError();
}

template <typename PT, int MODE>
static cudaError_t launch_vfe_mode(const PT* pts, const VfeSmall& p, const float* wblob, const VfeProblem& prob,
                                   const VfeOutput& out, int sm_count, cudaStream_t st) {
  cudaError_t err = cudaFuncSetAttribute(vfe_kernel<PT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return err;
  // persistent: two resident CTAs per SM (8 compute warps + 1 background-writer warp each), tiles strided over them
  vfe_kernel<PT, MODE><<<2 * sm_count, kVfeThreads + 32, kSmemBytes, st>>>(pts, p, wblob, prob, out);
  return cudaGetLastError();
}

cudaError_t launch_vfe(const void* pts, int pts_dtype, const VfeSmall& p, const float* wblob, const VfeProblem& prob,
                       float* voxel_feat, int sm_count, cudaStream_t st, int* launches) {
  const VfeOutput out{voxel_feat, nullptr, nullptr, nullptr, nullptr, 0};
  ++*launches;
  if (pts_dtype == LISEC_F32)
    return launch_vfe_mode<float, 0>(static_cast<const float*>(pts), p, wblob, prob, out, sm_count, st);
  return launch_vfe_mode<double, 0>(static_cast<const double*>(pts), p, wblob, prob, out, sm_count, st);
}

cudaError_t launch_vfe_to_grid(const void* pts, int pts_dtype, const VfeSmall& p, const float* wblob,
                               const VfeProblem& prob, const Workspace& w, const Geom& g, int n_sweeps, int grid_dtype,
                               void* grid, int sm_count, cudaStream_t st, int* launches) {
  const VfeOutput out{nullptr, grid, w.voxel_cell, w.cell_voxel, w.c_empty, (long long)n_sweeps * g.cells};
  ++*launches;
  if (pts_dtype == LISEC_F32) {
    const float* q = static_cast<const float*>(pts);
    return grid_dtype == LISEC_F32 ? launch_vfe_mode<float, 1>(q, p, wblob, prob, out, sm_count, st)
                                   : launch_vfe_mode<float, 2>(q, p, wblob, prob, out, sm_count, st);
  }
  const double* q = static_cast<const double*>(pts);
  return grid_dtype == LISEC_F32 ? launch_vfe_mode<double, 1>(q, p, wblob, prob, out, sm_count, st)
                                 : launch_vfe_mode<double, 2>(q, p, wblob, prob, out, sm_count, st);
}

}  // namespace lisec
