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

#include <type_traits>

#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

int vfe_rows_per_tile(int T) { return kVfeThreads - T + 1; }

namespace {

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&h);
}

constexpr int kGroups = 2;  // tile groups per CTA
constexpr int kWriterThreads = 128;  // one warpgroup of background writers (register-trimmed with setmaxnreg)
#ifndef LISEC_WRITER_WARPS
#define LISEC_WRITER_WARPS 4
#endif
constexpr int kWriterWarps = LISEC_WRITER_WARPS;  // how many of its 4 warps actually write
constexpr int kCtaThreads = kGroups * kVfeThreads + kWriterThreads;
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
constexpr int OFF_FSTAGE = OFF_VOXCELL + kVox * 4;   // float[kRows][6] the NEXT tile's feature rows (cp.async prefetch)
constexpr int OFF_VSTAGE = OFF_FSTAGE + kRows * 6 * 4;  // int[kRows] the NEXT tile's row -> voxel
constexpr int kGroupBytes = OFF_VSTAGE + kRows * 4 - OFF_H1T;  // per tile group; the weights (first 20 KB) are shared
constexpr int OFF_BG = OFF_H1T + kGroups * kGroupBytes;  // 32 cells x 64 channels of c_empty: the TMA source tile
constexpr int kSmemBytes = OFF_BG + 32 * 64 * 4;
static_assert(32 * PV * 4 <= 16 * PR * 4 + 16 * PV * 4, "P2T must fit in H1T+P1T");
static_assert(kSmemBytes <= 232448, "one CTA per SM, 227 KB opt-in limit");
static_assert(kGroupBytes % 16 == 0, "group regions stay 16-byte aligned");

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

// One CTA per SM = two tile groups of 8 warps (each works on its own tile with its own named barrier) plus one
// warpgroup of background writers. 20 warps put 5 on every SM sub-partition, i.e. 96 registers each at launch; the
// writers then give theirs back (setmaxnreg.dec 32) and the four compute warpgroups grow to 112 (setmaxnreg.inc;
// the pool is per CTA, so the compute side can only take what the writers released: 4 x (112-96) = 96-32).
__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kVfeThreads) : "memory");
}
__device__ __forceinline__ void all_compute_sync() {
  asm volatile("bar.sync 3, %0;" ::"n"(kGroups * kVfeThreads) : "memory");
}

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

// MODE 0: voxel rows to voxel_feat[V][64] (float32), no background.  MODE 1 / 2: rows straight into the float32 /
// bf16 dense grid at their cell, background by the writer warpgroup.
struct VfeOutput {
  float* voxel_feat;
  void* grid;
  const int* voxel_cell;  // voxel row -> cell (sweep * cells + (z*nx + x)*ny + y)
  const int* cell_voxel;  // occupancy map
  const float* c_empty;
  long long ncells;
};

// ---- background writer (fused kernel, the CTA's last warpgroup) ----------------------------------------------
// c_empty goes into every EMPTY cell of the grid while the compute warps keep the FP32 pipe busy; occupied cells are
// written by the tiles' own voxel rows, so every grid element is still written exactly once.
// The data never touches the LSU: a 32-cell tile of replicated c_empty sits in shared memory and every run of
// consecutive empty cells is ONE TMA bulk store (cp.async.bulk shared -> global, SASS UBLKCP) issued by the lane of
// the run's first cell.
__device__ __forceinline__ void bulk_store(void* gdst, unsigned ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename GT>
__device__ __forceinline__ void background_writer(const int* __restrict__ cell_voxel, const float* __restrict__ c_empty,
                                                  GT* __restrict__ grid, long long ncells, unsigned char* sBg,
                                                  int wtid) {
  const int lane = wtid & 31, wwarp = wtid >> 5;
  // fill the tile: 32 cells x 64 channels of GT, every cell = c_empty (rounded once for bf16)
  for (int i = wtid; i < 32 * 64; i += kWriterThreads) {
    if (sizeof(GT) == 4) reinterpret_cast<float*>(sBg)[i] = c_empty[i & 63];
    else reinterpret_cast<__nv_bfloat16*>(sBg)[i] = __float2bfloat16_rn(c_empty[i & 63]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA
  asm volatile("bar.sync 4, %0;" ::"n"(kWriterThreads) : "memory");
  if (wwarp >= kWriterWarps) return;
  const unsigned src = (unsigned)__cvta_generic_to_shared(sBg);
  const int ngroups = (int)((ncells + 31) >> 5);
  const int stride = gridDim.x * kWriterWarps;
  constexpr int U = 4;  // 32-cell groups per step; the next step's occupancy words are already in flight
  auto load_occ = [&](int g0, int (&occ)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u * stride;
      const long long cell = ((long long)g << 5) + lane;
      occ[u] = (g < ngroups && cell < ncells) ? __ldg(cell_voxel + cell) : 0;  // 0 = "not empty": nothing to write
    }
  };
  int occ[U], nxt[U];
  int g0 = blockIdx.x * kWriterWarps + wwarp;
  load_occ(g0, occ);
  for (; g0 < ngroups; g0 += U * stride) {
    load_occ(g0 + U * stride, nxt);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned empty = __ballot_sync(0xffffffffu, occ[u] < 0);
      const bool starts = ((empty >> lane) & 1u) && !(lane > 0 && ((empty >> (lane - 1)) & 1u));
      if (starts) {
        const unsigned rest = ~(empty >> lane);  // zeros shift in on top, so rest == 0 only for lane 0 of a full group
        const int len = rest ? __ffs(rest) - 1 : 32;  // consecutive empty cells from this lane on
        GT* dst = grid + (((long long)(g0 + u * stride) << 5) + lane) * 64;
        bulk_store(dst, src, (unsigned)(len * 64 * sizeof(GT)));
      }
    }
    bulk_commit();
#pragma unroll
    for (int u = 0; u < U; ++u) occ[u] = nxt[u];
  }
  bulk_wait_all();  // the tile must outlive every read of it; also makes the writes complete before the warp retires
}

// Pre-pass, one thread per voxel: float64 mean of the voxel's kept points, added in list order with one divide —
// np.mean(currPoints, axis=0) (model_training.py:135) bit for bit — then the float32 feature rows
// [x,y,z,x-cx,y-cy,z-cz] (:137-140 + the Keras input cast) of its VFE rows, contiguous in row order, and six zeros for
// the virtual pad row (:141). The VFE kernel then starts every tile from one contiguous, prefetchable 24 B/row read
// instead of a chain of dependent gathers.
template <typename PT>
__global__ void __launch_bounds__(256) row_features_kernel(const PT* __restrict__ pts, int T,
                                                           const int* __restrict__ voxel_start,
                                                           const int* __restrict__ row_start,
                                                           const int* __restrict__ list_sorted,
                                                           const long long* __restrict__ totals,
                                                           float* __restrict__ row_feat) {
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
  const double cx = sx / n, cy = sy / n, cz = sz / n;
  float2* dst = reinterpret_cast<float2*>(row_feat + (size_t)row_start[v] * 6);
  for (int i = 0; i < kept; ++i) {
    PT x, y, z;
    load_point(pts, (long long)list_sorted[s + i], x, y, z);
    float f[6];
    point_features((double)x, (double)y, (double)z, cx, cy, cz, f);
    dst[3 * i] = make_float2(f[0], f[1]);
    dst[3 * i + 1] = make_float2(f[2], f[3]);
    dst[3 * i + 2] = make_float2(f[4], f[5]);
  }
  if (kept < T) {
    dst[3 * kept] = make_float2(0.f, 0.f);
    dst[3 * kept + 1] = make_float2(0.f, 0.f);
    dst[3 * kept + 2] = make_float2(0.f, 0.f);
  }
}

__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(kCtaThreads, 1)
    vfe_kernel(const __grid_constant__ VfeSmall P, const float* __restrict__ wblob,
               const __grid_constant__ VfeProblem prob, const __grid_constant__ VfeOutput out) {
  extern __shared__ __align__(16) unsigned char smem[];
  if (threadIdx.x >= kGroups * kVfeThreads) {  // the writer warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    const int wtid = threadIdx.x - kGroups * kVfeThreads;
    if (MODE == 1)
      background_writer(out.cell_voxel, out.c_empty, static_cast<float*>(out.grid), out.ncells, smem + OFF_BG, wtid);
    if (MODE == 2)
      background_writer(out.cell_voxel, out.c_empty, static_cast<__nv_bfloat16*>(out.grid), out.ncells, smem + OFF_BG,
                        wtid);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");  // 4 x (112 - 96) = 96 - 32: exactly what the writers released
  const int group = threadIdx.x / kVfeThreads;
  const int tid = threadIdx.x % kVfeThreads, lane = tid & 31, warp = tid >> 5;
  // register tiles: lane = row group (8 tile rows / 4 voxel rows), warp = column group
  unsigned char* gsm = smem + group * kGroupBytes;  // this group's activations; offsets below are group-relative
  float* sW2P = reinterpret_cast<float*>(smem + OFF_W2P);
  float* sW2X = reinterpret_cast<float*>(smem + OFF_W2X);
  float* sW3P = reinterpret_cast<float*>(smem + OFF_W3P);
  float* sW3X = reinterpret_cast<float*>(smem + OFF_W3X);
  float* sH1T = reinterpret_cast<float*>(gsm + OFF_H1T);
  float* sP1T = reinterpret_cast<float*>(gsm + OFF_P1T);
  float* sP2T = reinterpret_cast<float*>(gsm + OFF_P2T);
  float* sH2T = reinterpret_cast<float*>(gsm + OFF_H2T);
  float* sQ = reinterpret_cast<float*>(gsm + OFF_Q);
  unsigned char* sRowVox = gsm + OFF_ROWVOX;
  int* sVoxCell = reinterpret_cast<int*>(gsm + OFF_VOXCELL);
  float* sFeatStage = reinterpret_cast<float*>(gsm + OFF_FSTAGE);
  int* sVoxStage = reinterpret_cast<int*>(gsm + OFF_VSTAGE);

  const int n_tiles = (int)*prob.n_tiles;
  // weights: [W2P | W2X | W3P | W3X] as laid out by the host, straight into the first 20 KB (all compute threads)
  for (int i = threadIdx.x; i < (OFF_H1T / 16); i += kGroups * kVfeThreads)  // (writer threads left above)
    reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(wblob) + i);
  all_compute_sync();

  // tile header = (first voxel, first row) of the tile and of its successor; rows and voxels are contiguous
  struct Header { int v0, v1, r0, r1; };
  auto load_header = [&](int t) {
    Header h{0, 0, 0, 0};
    if (t < n_tiles) {
      h.v0 = __ldg(prob.tile_first + t);
      h.v1 = __ldg(prob.tile_first + t + 1);
      h.r0 = __ldg(prob.tile_row0 + t);
      h.r1 = __ldg(prob.tile_row0 + t + 1);
    }
    return h;
  };
  // asynchronous copy of a tile's feature rows (24 B each, contiguous) and row -> voxel words into the staging buffers
  auto prefetch_rows = [&](const Header& h) {
    const int nrows = h.r1 - h.r0;
    const float* src = prob.row_feat + (size_t)h.r0 * 6;
#pragma unroll
    for (int c = tid; c < kRows * 3; c += kVfeThreads)
      if (c < nrows * 3) cp_async8(sFeatStage + 2 * c, src + 2 * c);
    if (tid < nrows) cp_async4(sVoxStage + tid, prob.row_voxel + h.r0 + tid);
  };
  const int tstride = gridDim.x * kGroups;
  int t = blockIdx.x * kGroups + group;
  Header cur = load_header(t);
  if (t < n_tiles) prefetch_rows(cur);
  cp_async_commit();
  cp_async_wait_all();
  group_sync(group);

  for (; t < n_tiles; t += tstride) {
    const int v0 = cur.v0, nv = cur.v1 - cur.v0, nrows = cur.r1 - cur.r0;
    const bool has_row = tid < nrows;
    const Header nxt = load_header(t + tstride);  // consumed after the first barrier: its latency hides behind VFE-1

    // ---- VFE-1: Dense(6->16, no bias) + BN + ReLU (addVFELayer(in, 6, 32), :231 -> :155-166); one row per thread ----
    {
      float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int lv = 255;  // 255 = padding row past the tile's last row
      if (has_row) {
        const float2 f01 = *reinterpret_cast<const float2*>(sFeatStage + 6 * tid);
        const float2 f23 = *reinterpret_cast<const float2*>(sFeatStage + 6 * tid + 2);
        const float2 f45 = *reinterpret_cast<const float2*>(sFeatStage + 6 * tid + 4);
        f[0] = f01.x; f[1] = f01.y; f[2] = f23.x; f[3] = f23.y; f[4] = f45.x; f[5] = f45.y;
        lv = sVoxStage[tid] - v0;
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
    }
    group_sync(group);
    // the staging buffers are free again: start the next tile's rows (and this tile's voxel -> cell words, needed only
    // by the final store) on their way; they land while the GEMMs run
    if (t + tstride < n_tiles) prefetch_rows(nxt);
    if (MODE != 0 && tid < nv) cp_async4(sVoxCell + tid, out.voxel_cell + v0 + tid);
    cp_async_commit();
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
      pool_lane_rows<2>(val, make_pool_meta(sRowVox, lane), [&](int v, const float(&x)[2]) {
        if (v < nv) {
          sP1T[(2 * warp) * PV + v] = x[0];
          sP1T[(2 * warp + 1) * PV + v] = x[1];
        }
      });
    }
    group_sync(group);

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
    group_sync(group);
    {  // rows: 8x4 tile per thread, accumulators start at the voxel's pooled-half product
      float acc[8][4];
      {
        const uint2 rv = *reinterpret_cast<const uint2*>(sRowVox + lane * 8);  // local voxel of each of the 8 rows
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int v = ((r < 4 ? rv.x : rv.y) >> (8 * (r & 3))) & (kVox - 1);  // (padding rows read some valid row)
          const float4 q = *reinterpret_cast<const float4*>(sQ + v * QS + warp * 4);
          acc[r][0] = q.x; acc[r][1] = q.y; acc[r][2] = q.z; acc[r][3] = q.w;
        }
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
      group_sync(group);  // every warp is done with sH1T (A operand) and sQ: sP2T may now overwrite sH1T/sP1T
      pool_lane_rows<4>(acc, make_pool_meta(sRowVox, lane), [&](int v, const float(&x)[4]) {
        if (v < nv) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sP2T[(warp * 4 + c) * PV + v] = x[c];
        }
      });
    }
    group_sync(group);

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
    cp_async_wait_all();  // issued a whole tile ago; the barrier below publishes the staged rows and sVoxCell
    group_sync(group);
    {  // rows: 8x8 tile per thread; the per-voxel max goes straight from registers to the output row
      float out8[8][8];
      {
        const uint2 rv = *reinterpret_cast<const uint2*>(sRowVox + lane * 8);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float* q = sQ + (((r < 4 ? rv.x : rv.y) >> (8 * (r & 3))) & (kVox - 1)) * QS;
          const float4 q0 = *reinterpret_cast<const float4*>(q + coff8[0]);
          const float4 q1 = *reinterpret_cast<const float4*>(q + coff8[1]);
          out8[r][0] = q0.x; out8[r][1] = q0.y; out8[r][2] = q0.z; out8[r][3] = q0.w;
          out8[r][4] = q1.x; out8[r][5] = q1.y; out8[r][6] = q1.z; out8[r][7] = q1.w;
        }
      }
      tile_gemm<8, 2, 32, false>(sH2T, PR, 128, lane * 4, sW3X, 64, coff8, out8);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int col = coff8[c >> 2] + (c & 3);
        const float a = P.a3[col], b = P.b3[col];
#pragma unroll
        for (int r = 0; r < 8; ++r) out8[r][c] = fmaxf(fmaf(out8[r][c], a, b), 0.f);
      }
      pool_lane_rows<8>(out8, make_pool_meta(sRowVox, lane), [&](int v, const float(&x)[8]) {
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
    group_sync(group);  // the next tile's setup overwrites sH1T and sRowVox
    cur = nxt;
  }
}

}  // namespace

cudaError_t launch_row_features(const void* pts, int pts_dtype, const Geom& g, const Workspace& w, long long max_voxels,
                                cudaStream_t st, int* launches) {
  const unsigned blocks = (unsigned)((max_voxels + 255) / 256);  // threads past the device-side voxel count exit
  if (pts_dtype == LISEC_F32)
    row_features_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(pts), g.T, w.voxel_start, w.row_start,
                                                       w.list_sorted, w.totals, w.row_feat);
  else
    row_features_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(pts), g.T, w.voxel_start,
                                                        w.row_start, w.list_sorted, w.totals, w.row_feat);
  ++*launches;
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_vfe_mode(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const VfeOutput& out,
                                   int sm_count, cudaStream_t st) {
  cudaError_t err = cudaFuncSetAttribute(vfe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return err;
  // persistent: one CTA per SM (2 tile groups x 8 warps + 1 writer warpgroup), tiles strided over the tile groups
  vfe_kernel<MODE><<<sm_count, kCtaThreads, kSmemBytes, st>>>(p, wblob, prob, out);
  return cudaGetLastError();
}

cudaError_t launch_vfe(const VfeSmall& p, const float* wblob, const VfeProblem& prob, float* voxel_feat, int sm_count,
                       cudaStream_t st, int* launches) {
  const VfeOutput out{voxel_feat, nullptr, nullptr, nullptr, nullptr, 0};
  ++*launches;
  return launch_vfe_mode<0>(p, wblob, prob, out, sm_count, st);
}

cudaError_t launch_vfe_to_grid(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const Workspace& w,
                               const Geom& g, int n_sweeps, int grid_dtype, void* grid, int sm_count, cudaStream_t st,
                               int* launches) {
  const VfeOutput out{nullptr, grid, w.voxel_cell, w.cell_voxel, w.c_empty, (long long)n_sweeps * g.cells};
  ++*launches;
  return grid_dtype == LISEC_F32 ? launch_vfe_mode<1>(p, wblob, prob, out, sm_count, st)
                                 : launch_vfe_mode<2>(p, wblob, prob, out, sm_count, st);
}

}  // namespace lisec
