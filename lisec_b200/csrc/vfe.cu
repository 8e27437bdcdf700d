// Stacked VFE on sm_100a: centroid augmentation, three pointwise linears (+BN+ReLU), two per-voxel max-pools with
// concat, and the final max over T — reference model_training.py:134-141 (features) and :155-186, 229-235 (layers).
//
// Work unit: a tile = a run of whole voxels holding at most 256 VFE rows (a row is a kept point, or the single
// virtual zero row that stands for all identical pad rows of a non-full voxel). One persistent CTA per
// SM, warp-specialised into a three-stage pipeline over tiles:
//
//   WRITER (warps 0-2, fused modes)  streams c_empty into the empty cells with TMA bulk stores (cp.async.bulk, evict-first)
//   TENSOR (warp 3, one lane)        FCN Dense(64->64): D^T[64 ch x 256 rows] = W3^T * X^T as 24 tcgen05.mma (kind::tf32,
//                                    M=64, N=256, K=8; 3xTF32: Wh*Xl + Wl*Xh + Wh*Xh), 4 accumulators in TMEM,
//                                    completion signalled to an mbarrier by tcgen05.commit
//   BACK   (warps 4-7)               tcgen05.ld: a thread owns one output CHANNEL and walks the tile's rows in order, so
//                                    the final per-voxel max is a sequential in-register scan; BN + ReLU are applied once
//                                    per voxel (they are monotonic, so they commute with the max), and the row goes
//                                    straight to voxel_feat or to its grid cell
//   FRONT  (warps 8-23, FP32 pipe)   VFE-1 (6->16, compensated float32), max-pool, VFE-2 (32->32) as register-tiled SIMT
//                                    GEMMs with in-register max-pools; leaves the FCN input X = [pooled | pointwise]
//                                    (256 x 64), split into tf32 hi/lo parts, in shared memory in the tensor core's
//                                    operand layout
//
// Why the FCN alone goes to the tensor core: it is 75 % of the path's FLOPs, it follows the last ReLU (no cancellation
// in its sums), and 3xTF32 reproduces a float32 FMA chain (tools/umma_probe.cu: 1.1e-6 vs 1.0e-6 of rms). dense_1's
// two halves cancel, so it stays on the FP32 pipe with blocked accumulation, and dense (6->16) acts on raw coordinates
// up to +-50 m and is evaluated in compensated float32 (Dot2; see VFE-1 below). Parity bar 1e-5 against the float64 oracle.
//
// Because Concatenate([pooled, pointwise]) feeds a bias-free Dense (model_training.py:164-165, 184), the pooled half of
// dense_1's product is the same for every row of a voxel: it is computed once per voxel (Q) and used as the
// accumulators' initial value. For the FCN the pooled half is simply broadcast into X's first 32 channels.
// SIMT thread mapping: lane = 8 consecutive tile rows, warp = column group, so a voxel's rows all sit in one warp and
// the max-pools are an in-register segmented max plus a 4-step segmented shuffle scan.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "vfe_math.cuh"

namespace lisec {

int vfe_rows_per_tile(int T) { return kVfeThreads - T + 1; }

namespace {

__device__ unsigned long long* g_trace = nullptr;  // debug timeline (set_trace_vfe), normally null

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&h);
}

// Warp roles. The SM's arbiter favours high warp ids, so the stage that sets the pace (FRONT) sits on top.
constexpr int kWriterWarps = 3;             // warps 0-2: background writers (fused modes)
constexpr int kTensorWarp = 3;              // warp 3: lane 0 issues the MMAs
constexpr int kBackWarp0 = 4;               // warps 4-7: back stage, warp id % 4 = its TMEM lane quadrant
constexpr int kBackThreads = 128;
constexpr int kFrontWarp0 = 8;              // warps 8-23: front stage (16 warps on one 256-row tile: the stage is
constexpr int kFrontWarps = 16;             // latency-bound, so it gets the warps; each warp owns 2 of 32 channels)
constexpr int kFrontThreads = 32 * kFrontWarps;
constexpr int kCtaThreads = 32 * kFrontWarp0 + kFrontThreads;
constexpr int kSlots = 4;                   // accumulator buffers in flight (and TileInfo slots)
constexpr int kRows = kVfeThreads;     // 256 rows per tile
constexpr int kVox = kVfeThreads / 2;  // 128 voxels per tile: a non-full voxel has >= 2 rows, a full one T >= 2
constexpr int PR = kRows + 4;          // float pitch of row-indexed k-major tiles (16-byte aligned rows, 4-bank skew)
constexpr int PV = kVox + 4;           // float pitch of voxel-indexed k-major tiles
constexpr int QS = 34;                 // float stride of a voxel's 32-channel row in sQ / sP2: 8-byte accesses at
                                       // v * QS + 2 * warp fall on bank pair (v + warp) % 16 -> consecutive voxels spread
constexpr int kBgCells = 32;           // cells in the writers' TMA source tile

// tensor-core operands (K-major, 128-byte swizzle; see umma.cuh)
constexpr uint32_t kXSlab = kRows * 128;  // one 32-channel half of X: 256 rows x 128 B
constexpr uint32_t kWSlab = 64 * 128;     // one 32-channel half of W3^T: 64 rows x 128 B
constexpr int kTmemCols = 512;            // 4 accumulators: 2 column ranges of 256 (one column per tile row) x 2 lane halves

// per accumulator buffer: what the back stage needs to know about the tile
struct TileInfo {
  unsigned last_mask[8];  // bit r: tile row r is the last row of its voxel
  int nrows, nv, v0, pad;
  int voxcell[kVox];      // cell of each tile voxel (grid output modes)
};

// shared-memory map (bytes; the base is 1 KB-aligned)
constexpr int OFF_XH = 0;                           // X hi: 2 slabs
constexpr int OFF_XL = OFF_XH + 2 * kXSlab;         // X lo
constexpr int OFF_W3H = OFF_XL + 2 * kXSlab;        // W3^T hi: 2 slabs (pooled half, pointwise half)
constexpr int OFF_W3L = OFF_W3H + 2 * kWSlab;       // W3^T lo
constexpr int OFF_W2P = OFF_W3L + 2 * kWSlab;       // [16][32]
constexpr int OFF_W2X = OFF_W2P + 16 * 32 * 4;      // [16][32]
constexpr int OFF_H1T = OFF_W2X + 16 * 32 * 4;      // [16][PR]   dense outputs of VFE-1, k-major
constexpr int OFF_P1T = OFF_H1T + 16 * PR * 4;      // [16][PV]   pooled VFE-1
constexpr int OFF_Q = OFF_P1T + 16 * PV * 4;        // [kVox][QS] pooled-half products of dense_1
constexpr int OFF_P2 = OFF_Q;                       // [kVox][QS] pooled VFE-2: same place — warp w alone reads and
                                                    // writes columns 2w, 2w+1 of both, so a __syncwarp orders them
constexpr int OFF_ROWVOX = OFF_Q + kVox * QS * 4;   // uint8[kRows] local voxel of each tile row
constexpr int OFF_VOXCELL = OFF_ROWVOX + kRows;     // int[kVox] cell of each tile voxel, front stage's own copy
constexpr int OFF_FSTAGE = OFF_VOXCELL + kVox * 4;  // float[kRows][6] the NEXT tile's feature rows (cp.async prefetch)
constexpr int OFF_VSTAGE = OFF_FSTAGE + kRows * 6 * 4;  // int[kRows] the NEXT tile's row -> voxel
constexpr int OFF_INFO = OFF_VSTAGE + kRows * 4;    // TileInfo[kSlots]
constexpr int OFF_BAR = OFF_INFO + kSlots * (int)sizeof(TileInfo);  // mbarriers acc_full[4], acc_empty[4], x_full
constexpr int OFF_TMEM_SLOT = OFF_BAR + 8 * (2 * kSlots + 1);       // TMEM base address (written by tcgen05.alloc)
constexpr int OFF_BG = (OFF_TMEM_SLOT + 8 + 127) & ~127;            // kBgCells x 64 channels of c_empty: the TMA source tile
constexpr int kSmemBytes = OFF_BG + kBgCells * 64 * 4;  // dynamic shared memory starts 1 KB-aligned (no static smem)
static_assert(kSmemBytes <= 232448, "one CTA per SM, 227 KB opt-in limit");
static_assert(OFF_W3H % 1024 == 0 && OFF_XL % 1024 == 0, "operand slabs are 1 KB-aligned");
static_assert(sizeof(TileInfo) % 16 == 0 && OFF_INFO % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

// ---- register-tile GEMM: acc[R][C] += A[k][row(r)] * W[k][col(c)], k = 0..K-1 -------------------------------
// The lane's rows come as R/4 float4 chunks at row0 + i*chunk_stride (consecutive lanes -> consecutive 16 bytes);
// the warp's columns are NC consecutive ones at col0 (same address for every lane: a broadcast load).
// float32 accumulation in blocks of 4 k-steps (a fresh accumulator per block, then added): dense_1's two halves cancel.
// NC = 2 or 4 columns per thread at column offset col0 (one 8- or 16-byte broadcast load per k).
template <int R, int NC, int K>
__device__ __forceinline__ void tile_gemm_blocked(const float* __restrict__ sA, int pitch, int chunk_stride, int row0,
                                                  const float* __restrict__ sW, int ldw, int col0,
                                                  float (&acc)[R][NC]) {
  static_assert(R % 4 == 0 && K % 4 == 0 && (NC == 2 || NC == 4), "tile shape");
#pragma unroll 1
  for (int kb = 0; kb < K; kb += 4) {
    float blk[R][NC];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = kb + kk;
      float a[R], b[NC];
#pragma unroll
      for (int i = 0; i < R / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(sA + k * pitch + row0 + chunk_stride * i);
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
      }
      if (NC == 4) {
        const float4 v = *reinterpret_cast<const float4*>(sW + k * ldw + col0);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[NC - 1] = v.w;
      } else {
        const float2 v = *reinterpret_cast<const float2*>(sW + k * ldw + col0);
        b[0] = v.x; b[1] = v.y;
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c)  // the block's first term is a plain product (no zeroing pass; only the sign of a zero differs)
          blk[r][c] = kk == 0 ? __fmul_rn(a[r], b[c]) : fmaf(a[r], b[c], blk[r][c]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[r][c] += blk[r][c];
  }
}

__device__ __forceinline__ void front_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kFrontThreads) : "memory");
}

// position of tile row r inside a k-major row tile: lane l = r/8 owns rows 8l..8l+7 and loads them as two float4 at
// 4l and 128+4l, so both 16-byte loads of a warp cover 512 contiguous bytes (no bank conflicts)
__device__ __forceinline__ int row_pos(int r) { return ((r >> 3) << 2) + (r & 3) + ((r & 4) << 5); }

// Row r of the tile is row n(r) of the tensor-core operand X (and column n(r) of the accumulator): the 8 rows of lane
// l are rotated by l inside their group of 8, so that for a fixed register index the lanes of a quarter-warp hit eight
// different 16-byte chunks of the swizzled 128-byte lines (conflict-free STS.128). The back stage undoes the rotation
// with compile-time register indices.
__device__ __forceinline__ int x_row(int l, int i) { return 8 * l + ((i + l) & 7); }
__device__ __forceinline__ uint32_t x_offset(int slab, int n, int chunk) {
  return (uint32_t)slab * kXSlab + (uint32_t)n * 128u + (uint32_t)((chunk ^ (n & 7)) << 4);
}

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
// bf16 dense grid at their cell, background by the writer warps.
struct VfeOutput {
  float* voxel_feat;
  void* grid;
  const int* voxel_cell;  // voxel row -> cell (sweep * cells + (z*nx + x)*ny + y)
  const int* cell_voxel;  // occupancy map
  const float* c_empty;
  long long ncells;
  const int* warm;       // the per-cell count table: pulled back into L2 for the NEXT call's point pass (see the writer)
  long long* prof;  // debug (LISEC_TRACE=1): per-CTA cycle counters [kProfSlots], see Workspace::trace
};
constexpr int kProfSlots = 16;
// slots: 0 front total, 1 VFE-1, 2 pool1, 3 Q2, 4 dense_1, 5 wait X free, 6 X pointwise, 7 pool2, 8 X pooled,
//        9 wait accumulator free, 10 info + MMA issue, 11 back total, 12 back wait full, 13 back scan, 14 tiles
#ifndef LISEC_PROF
#define LISEC_PROF 0  // 1: compile the per-stage cycle counters in (make PROF=1); they cost registers in the hot loops
#endif
#if !LISEC_PROF
struct Prof {
  long long* dst;
  long long acc[kProfSlots];
  __device__ __forceinline__ void begin(long long*) { dst = nullptr; }
  __device__ __forceinline__ void lap(int) {}
  __device__ __forceinline__ void flush(int, int) {}
};
#else
struct Prof {
  long long* dst;
  long long t0, acc[kProfSlots];
  __device__ __forceinline__ void begin(long long* d) {
    dst = d;
    if (dst) {
#pragma unroll
      for (int i = 0; i < kProfSlots; ++i) acc[i] = 0;
      t0 = clock64();
    }
  }
  __device__ __forceinline__ void lap(int slot) {
    if (dst) {
      const long long t1 = clock64();
      acc[slot] += t1 - t0;
      t0 = t1;
    }
  }
  __device__ __forceinline__ void flush(int first, int last) {
    if (dst)
      for (int i = first; i <= last; ++i) dst[(size_t)blockIdx.x * kProfSlots + i] = acc[i];
  }
};
#endif

// ---- background writer (fused modes, warps 12-15) ------------------------------------------------------------
// c_empty goes into every EMPTY cell of the grid while the other warps compute; occupied cells are written by the
// back stage, so every grid element is still written exactly once. The data never touches the LSU: a 32-cell tile of
// replicated c_empty sits in shared memory and every run of consecutive empty cells is ONE TMA bulk store
// (cp.async.bulk shared -> global, SASS UBLKCP) issued by the lane of the run's first cell.
// The 1.2 GB background stream is written once and not read again by this path: evict-first in L2, so that it does
// not push out the row features, tile tables and cell maps the other warps are prefetching.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_store(void* gdst, unsigned ssrc, unsigned bytes, unsigned long long policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc),
               "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename GT>
__device__ __forceinline__ void background_writer(const int* __restrict__ cell_voxel, const float* __restrict__ c_empty,
                                                  GT* __restrict__ grid, long long ncells, unsigned char* sBg,
                                                  int wtid) {
  constexpr int kWarps = kWriterWarps;
  const int lane = wtid & 31, wwarp = wtid >> 5;
  // fill the tile: 32 cells x 64 channels of GT, every cell = c_empty (rounded once for bf16)
  for (int i = wtid; i < kBgCells * 64; i += 32 * kWriterWarps) {
    if (sizeof(GT) == 4) reinterpret_cast<float*>(sBg)[i] = c_empty[i & 63];
    else reinterpret_cast<__nv_bfloat16*>(sBg)[i] = __float2bfloat16_rn(c_empty[i & 63]);
  }
  umma::fence_async_smem();  // generic-proxy writes -> visible to the TMA
  asm volatile("bar.sync 2, %0;" ::"n"(32 * kWriterWarps) : "memory");
  const unsigned src = (unsigned)__cvta_generic_to_shared(sBg);
  const unsigned long long policy = l2_evict_first_policy();
  const int ngroups = (int)((ncells + 31) >> 5);
  const int stride = gridDim.x * kWarps;
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
  int g0 = blockIdx.x * kWarps + wwarp;
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
        bulk_store(dst, src, (unsigned)(len * 64 * sizeof(GT)), policy);
      }
    }
    bulk_commit();
#pragma unroll
    for (int u = 0; u < U; ++u) occ[u] = nxt[u];
  }
  bulk_wait_all();  // the tile must outlive every read of it; also makes the writes complete before the warp retires
}

// The grid stream has just pushed everything else out of L2, and the next call's point pass starts with ~0.5 M
// scattered atomics into the per-cell count table (ncu: it was bound by those read-modify-writes going to DRAM). The
// writers are done well before the other stages, so they pull the table (4 B per cell, all zeros at this point) back
// into L2, marked evict-last.
__device__ __forceinline__ void warm_count_table(const int* __restrict__ count, long long ncells, int wtid) {
  const long long lines = (ncells * 4 + 127) >> 7;
  const char* base = reinterpret_cast<const char*>(count);
  for (long long i = (long long)blockIdx.x * (32 * kWriterWarps) + wtid; i < lines;
       i += (long long)gridDim.x * (32 * kWriterWarps))
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + (i << 7)));
}

// Pre-pass, one thread per VFE row: float64 mean of the row's voxel's kept points, added in list order with one
// divide — np.mean(currPoints, axis=0) (model_training.py:135) bit for bit — then the float32 feature row
// [x,y,z,x-cx,y-cy,z-cz] (:137-140 + the Keras input cast), or six zeros for the virtual pad row (:141). Rows are
// contiguous in row order, so the VFE kernel starts every tile from one prefetchable 24 B/row read instead of a chain
// of dependent gathers. The rows of a voxel each redo its (<= T term) sum: those re-reads hit L1, the stores are
// coalesced, and no thread carries a whole saturated voxel alone.
template <typename PT>
__global__ void __launch_bounds__(256, 6) row_features_kernel(const PT* __restrict__ pts, int T,
                                                           const int* __restrict__ voxel_start,
                                                           const int* __restrict__ row_start,
                                                           const int* __restrict__ row_voxel,
                                                           const int* __restrict__ list_sorted,
                                                           const long long* __restrict__ totals,
                                                           float* __restrict__ row_feat) {
  pdl_launch_dependents();
  pdl_wait();
  timeline_stamp(g_trace, TL_ROWFEAT);
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= totals[TOT_ROWS]) return;
  const int v = row_voxel[r];
  const int s = voxel_start[v];
  const int c = voxel_start[v + 1] - s;
  const int kept = c < T ? c : T;
  const int mine = (int)(r - row_start[v]);  // == kept: the virtual pad row
  float2* dst = reinterpret_cast<float2*>(row_feat + (size_t)r * 6);
  if (mine >= kept) {
    dst[0] = make_float2(0.f, 0.f);
    dst[1] = make_float2(0.f, 0.f);
    dst[2] = make_float2(0.f, 0.f);
    return;
  }
  // The loads of a chunk are issued together; the additions stay in list order, one at a time.
  constexpr int CH = 4;
  double sx = 0.0, sy = 0.0, sz = 0.0, px = 0.0, py = 0.0, pz = 0.0;
  for (int i0 = 0; i0 < kept; i0 += CH) {
    int id[CH];
    PT x[CH], y[CH], z[CH];
#pragma unroll
    for (int u = 0; u < CH; ++u) id[u] = i0 + u < kept ? list_sorted[s + i0 + u] : -1;
#pragma unroll
    for (int u = 0; u < CH; ++u)
      if (id[u] >= 0) load_point(pts, (long long)id[u], x[u], y[u], z[u]);
#pragma unroll
    for (int u = 0; u < CH; ++u)
      if (id[u] >= 0) {
        sx += (double)x[u];
        sy += (double)y[u];
        sz += (double)z[u];
        if (i0 + u == mine) {
          px = (double)x[u];
          py = (double)y[u];
          pz = (double)z[u];
        }
      }
  }
  const double n = (double)kept;
  float f[6];
  point_features(px, py, pz, sx / n, sy / n, sz / n, f);
  dst[0] = make_float2(f[0], f[1]);
  dst[1] = make_float2(f[2], f[3]);
  dst[2] = make_float2(f[4], f[5]);
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
__device__ __forceinline__ void cp_async_wait_all_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// ---- TENSOR stage: the FCN's Dense(64->64) for one tile, issued by one thread (lane 0 of the tensor warp) ----
// D^T[ch][n] (+)= sum_k W3^T[ch][k] * X[n][k], k over [pooled 32 | pointwise 32]. 3xTF32, small terms first so that
// their sum is not rounded against the large one: Wh*Xl, Wl*Xh, then Wh*Xh.
__device__ __forceinline__ void issue_fcn_mma(uint32_t smem_base, uint32_t d_tmem) {
  constexpr uint32_t idesc = umma::make_idesc_tf32_k(64, kRows);
  const uint32_t w[3] = {smem_base + OFF_W3H, smem_base + OFF_W3L, smem_base + OFF_W3H};
  const uint32_t x[3] = {smem_base + OFF_XL, smem_base + OFF_XH, smem_base + OFF_XH};
  uint32_t acc = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {  // k-step kb: slab kb/4, 32 bytes per step inside the swizzled 128-byte rows
      const uint64_t a = umma::make_desc_k_sw128(w[s] + (kb >> 2) * kWSlab + (kb & 3) * 32);
      const uint64_t b = umma::make_desc_k_sw128(x[s] + (kb >> 2) * kXSlab + (kb & 3) * 32);
      umma::mma_tf32_ss(d_tmem, a, b, idesc, acc);
      acc = 1;
    }
}

// mbarrier addresses and the accumulator of pipeline slot s = (tile ordinal) & 3: even ordinals use TMEM lanes 0..15
// of every quadrant, odd ones lanes 16..31 (the two interleaved placements of an M=64 accumulator); the column range
// alternates every second tile.
__device__ __forceinline__ uint32_t bar_acc_full(uint32_t smem_base, int s) { return smem_base + OFF_BAR + 8 * s; }
__device__ __forceinline__ uint32_t bar_acc_empty(uint32_t smem_base, int s) {
  return smem_base + OFF_BAR + 8 * (kSlots + s);
}
__device__ __forceinline__ uint32_t bar_x_full(uint32_t smem_base) { return smem_base + OFF_BAR + 8 * (2 * kSlots); }
__device__ __forceinline__ uint32_t acc_addr(uint32_t tmem_base, int s) {
  return tmem_base + ((uint32_t)(16 * (s & 1)) << 16) + (uint32_t)(s >> 1) * kRows;
}

// The tensor warp's loop: X complete -> accumulator free -> 24 MMAs -> commit.
__device__ __forceinline__ void tensor_stage(uint32_t smem_base, uint32_t tmem_base, int my_tiles) {
  for (int it = 0; it < my_tiles; ++it) {
    const int s = it & (kSlots - 1);
    umma::mbar_wait(bar_x_full(smem_base), it & 1);
    if (it >= kSlots) umma::mbar_wait(bar_acc_empty(smem_base, s), ((it / kSlots) - 1) & 1);
    umma::fence_after_sync();
    issue_fcn_mma(smem_base, acc_addr(tmem_base, s));
    umma::mma_commit(bar_acc_full(smem_base, s));
    umma::mbar_arrive(bar_acc_full(smem_base, s));  // release: publishes the front stage's TileInfo to the back stage
  }
}

// ---- BACK stage: accumulators -> per-voxel max -> BN + ReLU -> output row ------------------------------------
// Warp q of the stage reads TMEM lanes 32q..32q+31; an M=64 accumulator keeps channel c in lane (c % 16) + 32 (c / 16)
// (+16 for the interleaved placement), so lanes 0..15 of the warp own channels 16q..16q+15 of an EVEN tile and lanes
// 16..31 the same channels of the following ODD tile: the warp scans a pair of tiles per pass, every lane busy.
// Columns are tile rows (rotated inside groups of 8, see x_row): a thread walks its tile's rows in order, so the per-
// voxel max is a sequential scan; voxel ends come from the tile's last-row bit mask (per lane: the two halves of the
// warp work on different tiles). y = relu(a*z + b) is monotonic in z; the host folds sign(a) into dense_2's output
// column (api.cu), so a >= 0 here for every channel and max_rows relu(a*z_r + b) = relu(a*max_r z_r + b): one running
// max per thread, BN + ReLU applied once per voxel.
template <int MODE>
__device__ __forceinline__ void back_stage(const VfeSmall& P, const VfeOutput& out, unsigned char* smem,
                                           uint32_t smem_base, int my_tiles, uint32_t tmem_base, int bwarp, int lane) {
  const int ch = 16 * bwarp + (lane & 15), hh = lane >> 4;
  const float a = P.a3[ch], b = P.b3[ch];  // a = |BN scale|, see above
  const uint32_t tlane = tmem_base + ((uint32_t)(32 * bwarp) << 16);
  Prof prof;
  prof.begin(bwarp == 0 && lane == 0 ? out.prof : nullptr);
  for (int p = 0; 2 * p < my_tiles; ++p) {
    const int it = 2 * p + hh;           // this lane's tile
    const bool valid = it < my_tiles;
    const int s_even = (2 * p) & (kSlots - 1), s = valid ? (it & (kSlots - 1)) : s_even;
    umma::mbar_wait(bar_acc_full(smem_base, s_even), ((2 * p) / kSlots) & 1);
    if (2 * p + 1 < my_tiles) umma::mbar_wait(bar_acc_full(smem_base, s_even + 1), ((2 * p + 1) / kSlots) & 1);
    umma::fence_after_sync();
    prof.lap(12);
    const TileInfo* info = reinterpret_cast<const TileInfo*>(smem + OFF_INFO) + s;
    const int nrows = valid ? info->nrows : 0;
    const int nrows_max = max(nrows, __shfl_xor_sync(0xffffffffu, nrows, 16));
    float mx = -INFINITY;
    int v = 0;
    int cell = MODE != 0 ? info->voxcell[0] : 0;
    float* row0 = MODE == 0 ? out.voxel_feat + (size_t)info->v0 * 64 + ch : nullptr;
    const uint32_t tcol = tlane + (uint32_t)((2 * p / 2) & 1) * kRows;
#pragma unroll 1
    for (int c0 = 0; c0 < nrows_max; c0 += 64) {  // two 32-column loads per step: the rotation pattern repeats every 64
      uint2 mask = *reinterpret_cast<const uint2*>(&info->last_mask[c0 >> 5]);
      if (!valid) mask = make_uint2(0u, 0u);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float x[32];
        umma::tmem_ld_32x32(tcol + c0 + 32 * half, x);
        const unsigned m = half ? mask.y : mask.x;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // tile row c0 + 32 half + 8 g + i sits in column 8 g + ((i + G) & 7), G = 4 half + g
            mx = fmaxf(mx, x[8 * g + ((i + 4 * half + g) & 7)]);
            if ((m >> (8 * g + i)) & 1u) {  // the voxel's last row (uniform over the 16 lanes that share the tile)
              const float y = fmaxf(fmaf(mx, a, b), 0.f);
              if (MODE == 0) row0[(size_t)v * 64] = y;
              else if (MODE == 1) __stcs(static_cast<float*>(out.grid) + (size_t)cell * 64 + ch, y);
              else static_cast<__nv_bfloat16*>(out.grid)[(size_t)cell * 64 + ch] = __float2bfloat16_rn(y);
              ++v;
              if (MODE != 0) cell = info->voxcell[v & (kVox - 1)];  // the next voxel's cell: in flight while its rows are scanned
              mx = -INFINITY;
            }
          }
      }
    }
    umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      umma::mbar_arrive(bar_acc_empty(smem_base, s_even));
      if (2 * p + 1 < my_tiles) umma::mbar_arrive(bar_acc_empty(smem_base, s_even + 1));
    }
    prof.lap(13);
  }
  if (prof.dst) prof.acc[11] = prof.acc[12] + prof.acc[13];
  prof.flush(11, 13);
}

template <int MODE>
__global__ void __launch_bounds__(kCtaThreads, 1)
    vfe_kernel(const __grid_constant__ VfeSmall P, const float* __restrict__ wblob,
               const __grid_constant__ VfeProblem prob, const __grid_constant__ VfeOutput out) {
  extern __shared__ __align__(1024) unsigned char smem[];  // the operand slabs need 1 KB alignment (128-byte swizzle)
  pdl_launch_dependents();  // (the set-up below touches nothing its predecessors write; pdl_wait() follows it)
  const int warp_in_cta = threadIdx.x >> 5;
  const uint32_t smem_base = umma::smem_u32(smem);
  if (smem_base & 1023u) __trap();
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_SLOT);

  // ---- one-time setup: weights, barriers, TMEM ----
  {
    // blob = [W2P | W2X] float32 row-major, then the W3^T hi and lo operand images (built by the host, api.cu)
    const float4* src = reinterpret_cast<const float4*>(wblob);
    float4* w2 = reinterpret_cast<float4*>(smem + OFF_W2P);
    float4* w3 = reinterpret_cast<float4*>(smem + OFF_W3H);
    constexpr int n2 = 2 * 16 * 32 / 4, n3 = 4 * (int)kWSlab / 16;
    for (int i = threadIdx.x; i < n2 + n3; i += kCtaThreads) {
      const float4 v = __ldg(src + i);
      if (i < n2) w2[i] = v;
      else w3[i - n2] = v;
    }
    umma::fence_async_smem();  // W3 is read by the tensor core
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      umma::mbar_init(bar_acc_full(smem_base, s), 2);  // tcgen05.commit + the issuing thread's own (release) arrive
      umma::mbar_init(bar_acc_empty(smem_base, s), kBackThreads / 32);
    }
    umma::mbar_init(bar_x_full(smem_base), kFrontThreads / 32);
    umma::mbar_init_fence();
  }
  if (warp_in_cta == kBackWarp0) umma::tmem_alloc<kTmemCols>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // from here on: the grouping, the row features and the occupancy map of this call
  if (MODE != 0) timeline_stamp(g_trace, TL_VFE);
  const int n_tiles = (int)*prob.n_tiles;
  // tiles are strided over the CTAs: this one owns ordinals it = 0 .. my_tiles-1, tile blockIdx.x + it * gridDim.x
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp_in_cta < kWriterWarps) {  // ---- WRITER ----
    if (MODE == 1)
      background_writer(out.cell_voxel, out.c_empty, static_cast<float*>(out.grid), out.ncells, smem + OFF_BG,
                        (int)threadIdx.x);
    if (MODE == 2)
      background_writer(out.cell_voxel, out.c_empty, static_cast<__nv_bfloat16*>(out.grid), out.ncells, smem + OFF_BG,
                        (int)threadIdx.x);
    if (MODE != 0 && out.warm) warm_count_table(out.warm, out.ncells, (int)threadIdx.x);
    return;
  }
  if (warp_in_cta == kTensorWarp) {  // ---- TENSOR ----
    if ((threadIdx.x & 31) == 0) tensor_stage(smem_base, tmem_base, my_tiles);
    return;
  }
  if (warp_in_cta < kFrontWarp0) {  // ---- BACK ----
    back_stage<MODE>(P, out, smem, smem_base, my_tiles, tmem_base, warp_in_cta - kBackWarp0, threadIdx.x & 31);
    asm volatile("bar.sync 3, %0;" ::"n"(kFrontThreads + kBackThreads) : "memory");
    if (warp_in_cta == kBackWarp0) umma::tmem_dealloc<kTmemCols>(tmem_base);
    return;
  }

  // ---- FRONT ----
  // 512 threads on one 256-row tile. GEMM / pooling mapping: lane = 8 consecutive tile rows (4 voxels for the
  // per-voxel product), warp w = channels 2w, 2w+1 of 32 (pool1: channel w of 16). VFE-1: thread = (row, half of the
  // 16 outputs).
  const int tid = threadIdx.x - 32 * kFrontWarp0, lane = tid & 31, warp = tid >> 5;
  float* sW2P = reinterpret_cast<float*>(smem + OFF_W2P);
  float* sW2X = reinterpret_cast<float*>(smem + OFF_W2X);
  float* sH1T = reinterpret_cast<float*>(smem + OFF_H1T);
  float* sP1T = reinterpret_cast<float*>(smem + OFF_P1T);
  float* sP2 = reinterpret_cast<float*>(smem + OFF_P2);
  float* sQ = reinterpret_cast<float*>(smem + OFF_Q);
  unsigned char* sRowVox = smem + OFF_ROWVOX;
  int* sVoxCell = reinterpret_cast<int*>(smem + OFF_VOXCELL);
  float* sFeatStage = reinterpret_cast<float*>(smem + OFF_FSTAGE);
  int* sVoxStage = reinterpret_cast<int*>(smem + OFF_VSTAGE);

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
    for (int c = tid; c < kRows * 3; c += kFrontThreads)
      if (c < nrows * 3) cp_async8(sFeatStage + 2 * c, src + 2 * c);
    if (tid < nrows) cp_async4(sVoxStage + tid, prob.row_voxel + h.r0 + tid);
  };
  const int tstride = gridDim.x;
  int t = blockIdx.x;
  Header cur = load_header(t);
  Header nxt = load_header(t + tstride);
  if (t < n_tiles) prefetch_rows(cur);
  cp_async_commit();
  cp_async_wait_all();
  front_sync();

  Prof prof;
  prof.begin(tid == 0 ? out.prof : nullptr);
  int it = 0;
  for (; t < n_tiles; t += tstride, ++it) {
    const int slot = it & (kSlots - 1);
    const int v0 = cur.v0, nv = cur.v1 - cur.v0, nrows = cur.r1 - cur.r0;
    const Header nxt2 = load_header(t + 2 * tstride);  // two tiles ahead: a whole tile to land

    // ---- VFE-1: Dense(6->16, no bias) + BN + ReLU (addVFELayer(in, 6, 32), :231 -> :155-166) ----
    // thread = (row, 8 of the 16 outputs). The sum runs over raw coordinates up to +-50 m and must come out as the
    // correctly rounded float64 result; float64 FMAs share their pipe with the tensor core (they stall for the whole
    // FCN of the previous tile), so it is done in float32 with exact error terms (Ogita-Rump-Oishi Dot2): TwoProduct
    // by FMA and TwoSum for the three large terms, a plain FMA chain for the three centroid offsets (|.| < 1 voxel).
    {
      const int row = tid & (kRows - 1), jh = tid >> 8;
      const bool has_row = row < nrows;
      float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int lv = 255;  // 255 = padding row past the tile's last row
      if (has_row) {
        const float2 f01 = *reinterpret_cast<const float2*>(sFeatStage + 6 * row);
        const float2 f23 = *reinterpret_cast<const float2*>(sFeatStage + 6 * row + 2);
        const float2 f45 = *reinterpret_cast<const float2*>(sFeatStage + 6 * row + 4);
        f[0] = f01.x; f[1] = f01.y; f[2] = f23.x; f[3] = f23.y; f[4] = f45.x; f[5] = f45.y;
        lv = sVoxStage[row] - v0;
      }
      const int pos = row_pos(row);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = 8 * jh + jj;
        const float w0 = P.w1f[0][j], w1 = P.w1f[1][j], w2 = P.w1f[2][j];
        const float p0 = __fmul_rn(f[0], w0), e0 = __fmaf_rn(f[0], w0, -p0);
        const float p1 = __fmul_rn(f[1], w1), e1 = __fmaf_rn(f[1], w1, -p1);
        const float p2 = __fmul_rn(f[2], w2), e2 = __fmaf_rn(f[2], w2, -p2);
        const float s1 = __fadd_rn(p0, p1), b1 = __fsub_rn(s1, p0);
        const float r1 = __fadd_rn(__fsub_rn(p0, __fsub_rn(s1, b1)), __fsub_rn(p1, b1));
        const float s2 = __fadd_rn(s1, p2), b2 = __fsub_rn(s2, s1);
        const float r2 = __fadd_rn(__fsub_rn(s1, __fsub_rn(s2, b2)), __fsub_rn(p2, b2));
        const float small = __fmaf_rn(f[3], P.w1f[3][j], __fmaf_rn(f[4], P.w1f[4][j], __fmul_rn(f[5], P.w1f[5][j])));
        const float corr = __fadd_rn(__fadd_rn(__fadd_rn(e0, e1), __fadd_rn(e2, r1)), __fadd_rn(r2, small));
        const float d = __fadd_rn(s2, corr);
        sH1T[j * PR + pos] = has_row ? fmaxf(fmaf(d, P.a1[j], P.b1[j]), 0.f) : 0.f;
      }
      if (jh == 0) sRowVox[row] = (unsigned char)lv;
    }
    front_sync();
    prof.lap(1);
    // the staging buffers are free again: start this tile's voxel -> cell words (needed by the back stage, group 1) and
    // the next tile's rows (needed when this tile is done, group 2) on their way; they land while the GEMMs run
    if (MODE != 0 && tid < nv) cp_async4(sVoxCell + tid, out.voxel_cell + v0 + tid);
    cp_async_commit();
    if (t + tstride < n_tiles) prefetch_rows(nxt);
    cp_async_commit();
    const PoolMeta meta = make_pool_meta(sRowVox, lane);
    {  // MaxPoolingVFELayer over T (:160); RepeatLayer is implicit. Warp w pools channel w.
      float val[8][1];
      const float* src = sH1T + warp * PR;
      const float4 lo = *reinterpret_cast<const float4*>(src + 4 * lane);
      const float4 hi = *reinterpret_cast<const float4*>(src + 128 + 4 * lane);
      val[0][0] = lo.x; val[1][0] = lo.y; val[2][0] = lo.z; val[3][0] = lo.w;
      val[4][0] = hi.x; val[5][0] = hi.y; val[6][0] = hi.z; val[7][0] = hi.w;
      pool_lane_rows<1>(val, meta, [&](int v, const float(&x)[1]) {
        if (v < nv) sP1T[warp * PV + v] = x[0];
      });
    }
    front_sync();
    prof.lap(2);

    // ---- VFE-2: Dense(32->32) + BN + ReLU on concat[pooled, pointwise] (addVFELayer(., 32, 64), :232) ----
    {  // pooled half, once per voxel: Q2[128 x 32] = P1[128 x 16] * W2p. A thread takes voxels lane, lane + 32, lane + 64,
       // lane + 96 (not 4 consecutive ones: its float2 stores into sQ are then conflict-free) x 2 channels.
      float acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < 16; kb += 4) {
        float blk[4][2];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float* a = sP1T + (kb + kk) * PV + lane;
          const float2 b = *reinterpret_cast<const float2*>(sW2P + (kb + kk) * 32 + 2 * warp);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float av = a[32 * r];
            blk[r][0] = kk == 0 ? __fmul_rn(av, b.x) : fmaf(av, b.x, blk[r][0]);
            blk[r][1] = kk == 0 ? __fmul_rn(av, b.y) : fmaf(av, b.y, blk[r][1]);
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][0] += blk[r][0];
          acc[r][1] += blk[r][1];
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float2*>(sQ + (lane + 32 * r) * QS + 2 * warp) = make_float2(acc[r][0], acc[r][1]);
    }
    __syncwarp();  // columns 2 warp, 2 warp + 1 of sQ are this warp's own: no block barrier
    prof.lap(3);
    float h2[8][2];  // rows 8 lane .. 8 lane + 7, channels 2 warp, 2 warp + 1 of the VFE-2 pointwise output
    {  // rows: 8x2 tile per thread, accumulators start at the voxel's pooled-half product
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int v = meta.v[r] & (kVox - 1);  // (padding rows read some valid row)
        const float2 q = *reinterpret_cast<const float2*>(sQ + v * QS + 2 * warp);
        h2[r][0] = q.x; h2[r][1] = q.y;
      }
      tile_gemm_blocked<8, 2, 16>(sH1T, PR, 128, lane * 4, sW2X, 32, 2 * warp, h2);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float sa = P.a2[2 * warp + c], sb = P.b2[2 * warp + c];
#pragma unroll
        for (int r = 0; r < 8; ++r) h2[r][c] = fmaxf(fmaf(h2[r][c], sa, sb), 0.f);
      }
    }
    prof.lap(4);
    // X is single-buffered: the previous tile's MMAs must have finished reading it
    if (it > 0) umma::mbar_wait(bar_acc_full(smem_base, (it - 1) & (kSlots - 1)), ((it - 1) / kSlots) & 1);
    prof.lap(5);
    // pointwise half of the FCN input: X[n][32 + 2 warp ..] = h2, split into tf32 hi / lo
    const uint32_t xsub = (uint32_t)(warp & 1) * 8;  // which half of the 16-byte chunk warp / 2
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float2 hi, lo;
      umma::tf32_split(h2[r][0], hi.x, lo.x);
      umma::tf32_split(h2[r][1], hi.y, lo.y);
      const uint32_t off = x_offset(1, x_row(lane, r), warp >> 1) + xsub;
      *reinterpret_cast<float2*>(smem + OFF_XH + off) = hi;
      *reinterpret_cast<float2*>(smem + OFF_XL + off) = lo;
    }
    __syncwarp();  // every lane has taken its accumulator seeds out of sQ: sP2 (same columns) may overwrite them
    prof.lap(6);
    pool_lane_rows<2>(h2, meta, [&](int v, const float(&x)[2]) {
      if (v < nv) *reinterpret_cast<float2*>(sP2 + v * QS + 2 * warp) = make_float2(x[0], x[1]);
    });
    __syncwarp();
    prof.lap(7);

    // ---- FCN input, pooled half: Concatenate([pooled, pointwise]) (:164-165) = the voxel's pooled row, repeated ----
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float2 p = *reinterpret_cast<const float2*>(sP2 + (meta.v[r] & (kVox - 1)) * QS + 2 * warp);
      float2 hi, lo;
      umma::tf32_split(p.x, hi.x, lo.x);
      umma::tf32_split(p.y, hi.y, lo.y);
      const uint32_t off = x_offset(0, x_row(lane, r), warp >> 1) + xsub;
      *reinterpret_cast<float2*>(smem + OFF_XH + off) = hi;
      *reinterpret_cast<float2*>(smem + OFF_XL + off) = lo;
    }
    prof.lap(8);
    // the slot's TileInfo is free once the back stage has drained the slot's previous use
    if (it >= kSlots) umma::mbar_wait(bar_acc_empty(smem_base, slot), ((it / kSlots) - 1) & 1);
    prof.lap(9);
    cp_async_wait_all_but_last();  // this thread's own sVoxCell word (group 1); the rows may still be in flight
    if (tid < kRows) {
      TileInfo* info = reinterpret_cast<TileInfo*>(smem + OFF_INFO) + slot;
      const int my = sRowVox[tid], next = tid + 1 < kRows ? sRowVox[tid + 1] : 255;
      const unsigned last = __ballot_sync(0xffffffffu, tid < nrows && (tid + 1 == nrows || my != next));
      if (lane == 0) info->last_mask[warp] = last;
      if (MODE != 0 && tid < nv) info->voxcell[tid] = sVoxCell[tid];
      if (tid == 0) {
        info->nrows = nrows;
        info->nv = nv;
        info->v0 = v0;
      }
    }
    umma::fence_async_smem();  // X (generic-proxy writes) -> visible to the tensor core
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(bar_x_full(smem_base));  // 16 warps -> the tensor warp issues this tile's MMAs
    cp_async_wait_all();  // the next tile's rows (group 2), issued a whole tile ago
    front_sync();  // publishes them; also: the next tile's VFE-1 overwrites sH1T and sRowVox, its Q2 overwrites sQ
    prof.lap(10);
    cur = nxt;
    nxt = nxt2;
  }
  if (prof.dst) {
    prof.acc[0] = 0;
    for (int i = 1; i <= 10; ++i) prof.acc[0] += prof.acc[i];
    prof.acc[14] = it;
    prof.flush(0, 10);
    prof.flush(14, 14);
  }
  asm volatile("bar.sync 3, %0;" ::"n"(kFrontThreads + kBackThreads) : "memory");
  if (MODE != 0 && tid == 0) {  // (threadIdx.x != 0 here: stamp by hand)
    unsigned long long* tr = g_trace;
    if (tr) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      atomicMax(tr + (size_t)(kTimelineRow0 + TL_VFE_END) * kTraceSlots + 1, now);
    }
  }
}

}  // namespace

cudaError_t set_trace_vfe(unsigned long long* trace) { return cudaMemcpyToSymbol(g_trace, &trace, sizeof(trace)); }


cudaError_t launch_row_features(const void* pts, int pts_dtype, const Geom& g, const Workspace& w, long long max_rows,
                                cudaStream_t st, int* launches) {
  const unsigned blocks = (unsigned)((max_rows + 255) / 256) + 1;  // threads past the device-side row count exit
  cudaError_t err;
  if (pts_dtype == LISEC_F32)
    err = launch_pdl(row_features_kernel<float>, blocks, 256, 0, st, static_cast<const float*>(pts), g.T,
                     (const int*)w.voxel_start, (const int*)w.row_start, (const int*)w.row_voxel,
                     (const int*)w.list_sorted, (const long long*)w.totals, w.row_feat);
  else
    err = launch_pdl(row_features_kernel<double>, blocks, 256, 0, st, static_cast<const double*>(pts), g.T,
                     (const int*)w.voxel_start, (const int*)w.row_start, (const int*)w.row_voxel,
                     (const int*)w.list_sorted, (const long long*)w.totals, w.row_feat);
  ++*launches;
  return err;
}

template <int MODE>
static cudaError_t launch_vfe_mode(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const VfeOutput& out,
                                   int sm_count, cudaStream_t st) {
  cudaError_t err = cudaFuncSetAttribute(vfe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return err;
  // persistent: one CTA per SM (3 writer warps, 1 tensor warp, 4 back warps, 16 front warps), tiles strided over the CTAs
  return launch_pdl(vfe_kernel<MODE>, sm_count, kCtaThreads, kSmemBytes, st, p, wblob, prob, out);
}

cudaError_t launch_vfe(const VfeSmall& p, const float* wblob, const VfeProblem& prob, float* voxel_feat, int sm_count,
                       cudaStream_t st, int* launches, long long* prof) {
  const VfeOutput out{voxel_feat, nullptr, nullptr, nullptr, nullptr, 0, nullptr, prof};
  ++*launches;
  return launch_vfe_mode<0>(p, wblob, prob, out, sm_count, st);
}

cudaError_t launch_vfe_to_grid(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const Workspace& w,
                               const Geom& g, int n_sweeps, int grid_dtype, void* grid, int sm_count, cudaStream_t st,
                               int* launches) {
  const VfeOutput out{nullptr, grid, w.voxel_cell, w.cell_voxel, w.c_empty, (long long)n_sweeps * g.cells, w.count,
                      reinterpret_cast<long long*>(w.trace)};
  ++*launches;
  return grid_dtype == LISEC_F32 ? launch_vfe_mode<1>(p, wblob, prob, out, sm_count, st)
                                 : launch_vfe_mode<2>(p, wblob, prob, out, sm_count, st);
}

}  // namespace lisec
