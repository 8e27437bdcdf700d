// Stacked VFE on sm_100a: centroid augmentation, three pointwise linears (+BN+ReLU), two per-voxel max-pools with
// concat, and the final max over T — reference model_training.py:128-141 (centroid, features) and :155-186, 229-235
// (layers).
//
// Work unit: a tile = a run of whole voxels holding at most 128 VFE rows (a row is a kept point, or the single
// virtual zero row that stands for all identical pad rows of a non-full voxel) = one M = 128 block of
// the tensor core. Tiles are packed greedily inside chunks of <= 512 rows (tile_plan_kernel, voxelize.cu); chunks are
// strided over the CTAs. One persistent CTA per SM, warp-specialised:
//
//   WRITER (warps 0-2, fused modes)  streams c_empty into the empty cells with TMA bulk stores (cp.async.bulk, evict-first)
//   TENSOR (warp 3, one lane)        issues both GEMMs of every tile with tcgen05.mma kind::tf32 as 3xTF32
//                                    (Wh*Xl + Wl*Xh + Wh*Xh), accumulators in TMEM, completion via tcgen05.commit:
//                                      VFE-2  D2[128 rows x 64] = X1[rows x 32] * W2B   (M=128, N=64, K=8 x 12): the pooled
//                                             half's product lands in columns 0..31, the pointwise half's in 32..63 —
//                                             two accumulators, added in float32 by the front stage (the halves cancel)
//                                      FCN    D3^T[64 ch x 128 rows] = W3^T * X2^T        (M=64, N=128, K=8 x 24)
//   BACK   (warps 4-7)               tcgen05.ld of D3^T: a thread owns one output CHANNEL and walks the tile's rows in
//                                    order, so the final per-voxel max is a sequential in-register scan; BN + ReLU once
//                                    per voxel (monotonic, so they commute with the max); the row goes straight to
//                                    voxel_feat or to its grid cell
//   FRONT  (warps 8-23)              two TEAMS of 8 warps, each on its own tile (team g: tile ordinals g, g+2, ...), so
//                                    that one team's latencies and barriers are covered by the other's arithmetic;
//                                    2 threads per tile row (TMEM lane = row, half of the channels each):
//                                      F1  float64 centroid of the row's voxel in list order (np.mean, :135), features,
//                                          VFE-1 Dense(6->16)+BN+ReLU on the FP32 pipe, per-voxel max through shared
//                                          memory, X1 = [pooled | pointwise] split into tf32 hi/lo in operand layout
//                                      F2  tcgen05.ld of D2, the two halves added, BN+ReLU, per-voxel max, X2 likewise
//
// VFE-1 without float64 products per output: Dense(6->16) acts on raw coordinates up to +-50 m and must come out close
// to the correctly rounded sum. Each coordinate is split exactly into a coarse part o (top 11 mantissa bits) and the
// rest l (|l| < 2^-10 |x|): sum_i o_i w_i is evaluated in float64 (3 DFMA per output, kept as a float32 hi + lo pair),
// the six small terms l_i w_i + (x_i - c_i) w_{3+i} are a float32 FMA chain, and z = hi + (lo + small).
// tools/vfe_numerics.py emulates this arithmetic (and the 3xTF32 GEMMs with truncating accumulation) on the CPU:
// worst 6.7e-6 against the float64 oracle under the parity metric, bar 1e-5.
#include <cuda_bf16.h>

#include <cstdio>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "vfe_math.cuh"

namespace lisec {

int vfe_rows_per_chunk(int T) { return kVfeChunkRows - T + 1; }  // a chunk: whole voxels, at most kVfeChunkRows rows

namespace {

__device__ unsigned long long* g_trace = nullptr;  // debug timeline (set_trace_vfe), normally null

// Warp roles. The SM's arbiter favours high warp ids, so the stage that carries the arithmetic (FRONT) sits on top.
constexpr int kWriterWarps = 3;             // warps 0-2: background writers (fused modes)
constexpr int kTensorWarp = 3;              // warp 3: lane 0 issues the MMAs
constexpr int kBackWarp0 = 4;               // warps 4-7: back stage, warp id % 4 = its TMEM lane quadrant
constexpr int kBackThreads = 128;
constexpr int kFrontWarp0 = 8;              // warps 8-23: front stage, two teams of 8 warps; inside a team
constexpr int kFrontWarps = 16;             // warp % 4 = TMEM lane quadrant = row block, warp / 4 = which half of the channels
constexpr int kTeamWarps = 8;
constexpr int kTeamThreads = 32 * kTeamWarps;
constexpr int kFrontThreads = 32 * kFrontWarps;
constexpr int kCtaThreads = 32 * kFrontWarp0 + kFrontThreads;
constexpr int kSlots = 4;                   // FCN accumulator buffers in flight (and TileInfo slots)
constexpr int kRows = kVfeThreads;          // 128 rows per tile
constexpr int kVox = kVfeThreads / 2;       // 64 voxels per tile: a non-full voxel has >= 2 rows, a full one T >= 2
constexpr int kBgCells = 16;                // cells in the writers' TMA source tile
constexpr int kMetaSlots = 2;               // per team: the tile in work and the one being prefetched

// tensor-core operands (K-major, 128-byte swizzle; see umma.cuh)
constexpr uint32_t kXSlab = kRows * 128;  // 128 rows x 32 channels
constexpr uint32_t kWSlab = 64 * 128;     // 64 rows x 32 channels
// TMEM columns: FCN accumulators [0, 256) (2 column ranges x the 2 lane placements of an M=64 accumulator), D2 x 2
constexpr int kTmemCols = 512;
constexpr uint32_t kD2Col0 = 2 * kRows;

// per FCN accumulator buffer: what the back stage needs to know about the tile
struct TileInfo {
  unsigned last_mask[kRows / 32];  // bit r: tile row r is the last row of its voxel
  int nrows, nv, v0, pad;
  int voxcell[kVox];               // cell of each tile voxel (grid output modes)
};
// per tile in flight in the front stage (cp.async landing buffers)
// (16-byte cp.async.cg chunks from 16-byte-aligned global addresses: every array starts up to 3 elements before the
// tile's first entry, hence the + 4)
struct TileMeta {
  int rowvox[kRows + 4];     // row -> voxel row (| kRowPadFlag)
  int vrs[kVox + 4 + 4];     // voxel -> first VFE row (absolute), nv + 1 entries
  int voxcell[kVox + 4];     // voxel -> cell
};
struct FrontParams {  // the front stage's share of VfeSmall, in shared memory (indexed by the thread's channel quarter)
  double w1d[3][16];
  float w1f[6][16];
  float a1[16], b1[16], a2[32], b2[32];
};

// shared-memory map (bytes; the base is 1 KB-aligned)
constexpr int OFF_X2H = 0;                          // X2 hi: 2 slabs (pooled, pointwise)
constexpr int OFF_X2L = OFF_X2H + 2 * kXSlab;       // X2 lo
constexpr int OFF_X1 = OFF_X2L + 2 * kXSlab;        // X1[team]: hi slab, lo slab each
constexpr int OFF_W3H = OFF_X1 + 4 * kXSlab;        // W3^T hi: 2 slabs   | the weight blob, one contiguous copy
constexpr int OFF_W3L = OFF_W3H + 2 * kWSlab;       // W3^T lo            |
constexpr int OFF_W2H = OFF_W3L + 2 * kWSlab;       // W2B hi: 1 slab     |
constexpr int OFF_W2L = OFF_W2H + kWSlab;           // W2B lo             |
// per team: layer outputs for the max-pools (sH2 [kRows][32] floats, 16-byte chunks XOR-swizzled with the row; sH1
// [kRows][16] lives in the same place, the team's stages being sequential), the tile's points, the landing buffers
constexpr int kHBytes = kRows * 32 * 4;
constexpr int kXyzBytes = kRows * 3 * 8 + 16;
constexpr int kHdrBytes = 4 * 16;  // ring of 4 tile headers (v0, v1, r0, r1), written by one thread of the team
constexpr int kTeamBytes = kHBytes + kXyzBytes + kMetaSlots * (int)sizeof(TileMeta) + kHdrBytes;
constexpr int OFF_TEAM = OFF_W2L + kWSlab;          // [2] x { H | PT xyz[kRows][3] | TileMeta[kMetaSlots] }
constexpr int OFF_PAR = OFF_TEAM + 2 * kTeamBytes;
constexpr int OFF_INFO = OFF_PAR + (int)sizeof(FrontParams);
constexpr int OFF_BAR = OFF_INFO + kSlots * (int)sizeof(TileInfo);  // acc_full[4], acc_empty[4], x1_full[team], d2_full[team], x2_full
constexpr int kNumBars = 2 * kSlots + 5;
constexpr int OFF_TMEM_SLOT = OFF_BAR + 8 * kNumBars;  // TMEM base address (written by tcgen05.alloc)
constexpr int OFF_BG = (OFF_TMEM_SLOT + 8 + 127) & ~127;  // kBgCells x 64 channels of c_empty: the TMA source tile
constexpr int kSmemBytes = OFF_BG + kBgCells * 64 * 4;    // dynamic shared memory starts 1 KB-aligned (no static smem)
static_assert(kSmemBytes <= 232448, "one CTA per SM, 227 KB opt-in limit");
static_assert(OFF_W3H % 1024 == 0 && OFF_X2L % 1024 == 0 && OFF_X1 % 1024 == 0 && OFF_W2H % 1024 == 0, "operand slabs are 1 KB-aligned");
static_assert(sizeof(TileInfo) % 16 == 0 && sizeof(TileMeta) % 16 == 0 && sizeof(FrontParams) % 16 == 0, "alignment");
static_assert(OFF_TEAM % 16 == 0 && kTeamBytes % 16 == 0 && OFF_PAR % 16 == 0 && OFF_INFO % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
static_assert((2 * kVfeW3ImageFloats + 2 * kVfeW2ImageFloats) * 4 == OFF_TEAM - OFF_W3H, "blob = the four operand images");

__device__ __forceinline__ void team_sync(int team) {  // named barriers 4, 5 (0 = CTA, 2 = writers, 3 = front + back)
  asm volatile("bar.sync %0, %1;" ::"r"(4 + team), "n"(kTeamThreads) : "memory");
}
// float offset of the 16-byte chunk c of row r in sH2 (8 chunks per row) / sH1 (4 chunks per row, two rows per 128 B)
__device__ __forceinline__ int h2_off(int r, int c) { return r * 32 + ((c ^ (r & 7)) << 2); }
__device__ __forceinline__ int h1_off(int r, int c) { return r * 16 + ((c ^ ((r >> 1) & 3)) << 2); }

// byte offset of the 16-byte chunk `chunk` (4 channels) of operand row n inside a slab
__device__ __forceinline__ uint32_t x_chunk(int n, int chunk) {
  return (uint32_t)n * 128u + (uint32_t)((chunk ^ (n & 7)) << 4);
}
// x (4 channels) -> tf32 hi / lo parts, stored as one 16-byte chunk each
__device__ __forceinline__ void store_split4(unsigned char* hi_slab, unsigned char* lo_slab, uint32_t off, const float (&x)[4]) {
  float4 hi, lo;
  umma::tf32_split(x[0], hi.x, lo.x);
  umma::tf32_split(x[1], hi.y, lo.y);
  umma::tf32_split(x[2], hi.z, lo.z);
  umma::tf32_split(x[3], hi.w, lo.w);
  *reinterpret_cast<float4*>(hi_slab + off) = hi;
  *reinterpret_cast<float4*>(lo_slab + off) = lo;
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
  int first_group;       // 32-cell groups below this one already hold the background (grid_fill_kernel, scatter.cu)
  int* writer_claim;     // [2] the writers' batch counter and finished-warp count, zero between launches (background_writer)
};

// ---- background writer (fused modes, warps 0-2) --------------------------------------------------------------
// c_empty goes into every EMPTY cell of the grid while the other warps compute; occupied cells are written by the
// back stage, so every grid element is still written exactly once. The data never touches the LSU: a 16-cell tile of
// replicated c_empty sits in shared memory and every run of consecutive empty cells (<= 32: one occupancy word per
// lane) is one or two TMA bulk stores (cp.async.bulk shared -> global, SASS UBLKCP) issued by the lane of the run's
// first cell.
// The 1.2 GB background stream is written once and not read again by this path: evict-first in L2, so that it does
// not push out the row tables and cell maps the other warps are prefetching.
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
                                                  GT* __restrict__ grid, long long ncells, int first_group,
                                                  int* __restrict__ claim, unsigned char* sBg, int wtid) {
  constexpr int kWarps = kWriterWarps;
  const int lane = wtid & 31, wwarp = wtid >> 5;
  // fill the tile: kBgCells cells x 64 channels of GT, every cell = c_empty (rounded once for bf16)
  for (int i = wtid; i < kBgCells * 64; i += 32 * kWarps) {
    if (sizeof(GT) == 4) reinterpret_cast<float*>(sBg)[i] = c_empty[i & 63];
    else reinterpret_cast<__nv_bfloat16*>(sBg)[i] = __float2bfloat16_rn(c_empty[i & 63]);
  }
  umma::fence_async_smem();  // generic-proxy writes -> visible to the TMA
  asm volatile("bar.sync 2, %0;" ::"n"(32 * kWriterWarps) : "memory");
  const unsigned src = (unsigned)__cvta_generic_to_shared(sBg);
  const unsigned long long policy = l2_evict_first_policy();
  const int ngroups = (int)((ncells + 31) >> 5);
  // Batches of U consecutive 32-cell groups (32 KB of float32 grid) are CLAIMED, not assigned: the writer warps of the
  // 148 CTAs do not run at one speed beside the VFE stages (equal static shares finished 86 us apart, tools/timeline.py),
  // and an SM whose writers are late is also where the VFE pipeline is late. claim[0] = next batch, claim[1] = writer
  // warps that have finished; the last one zeroes both for the next launch. Lane 0 claims two iterations ahead (an
  // iteration is ~3 us, the atomic's round trip ~1-3), the next batch's occupancy words are in flight one ahead.
  constexpr int U = 4;
  const int nbatches = (ngroups - first_group + U - 1) / U;
  auto load_occ = [&](int b, int (&occ)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = first_group + b * U + u;
      const long long cell = ((long long)g << 5) + lane;
      occ[u] = (b < nbatches && g < ngroups && cell < ncells) ? __ldcg(cell_voxel + cell) : 0;  // 0 = "not empty": nothing to write
    }
  };
  auto claim_one = [&]() { return lane == 0 ? atomicAdd(claim, 1) : 0; };
  int b_cur = __shfl_sync(0xffffffffu, claim_one(), 0);
  int pending = claim_one();  // (lane 0's register; first read one iteration later)
  int occ[U], nxt[U];
  load_occ(b_cur, occ);
  while (b_cur < nbatches) {
    const int b_nxt = __shfl_sync(0xffffffffu, pending, 0);
    pending = claim_one();
    load_occ(b_nxt, nxt);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned empty = __ballot_sync(0xffffffffu, occ[u] < 0);
      const bool starts = ((empty >> lane) & 1u) && !(lane > 0 && ((empty >> (lane - 1)) & 1u));
      if (starts) {
        const unsigned rest = ~(empty >> lane);  // zeros shift in on top, so rest == 0 only for lane 0 of a full group
        const int len = rest ? __ffs(rest) - 1 : 32;  // consecutive empty cells from this lane on
        GT* dst = grid + (((long long)(first_group + b_cur * U + u) << 5) + lane) * 64;
        const int first = len < kBgCells ? len : kBgCells;  // the source tile holds kBgCells cells
        bulk_store(dst, src, (unsigned)(first * 64 * sizeof(GT)), policy);
        if (len > kBgCells) bulk_store(dst + kBgCells * 64, src, (unsigned)((len - kBgCells) * 64 * sizeof(GT)), policy);
      }
    }
    bulk_commit();
#pragma unroll
    for (int u = 0; u < U; ++u) occ[u] = nxt[u];
    b_cur = b_nxt;
  }
  bulk_wait_all();  // the tile must outlive every read of it; also makes the writes complete before the warp retires
  if (lane == 0) {
    if (pending < 0) __trap();  // (waits for the claim still in flight: none may land after the counter's reset)
    __threadfence();
    if (atomicAdd(claim + 1, 1) == (int)gridDim.x * kWarps - 1) {
      claim[0] = 0;
      claim[1] = 0;
      __threadfence();
    }
  }
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

// 16 bytes, L2 only (.cg): the sources are tables written by this kernel's predecessors under programmatic dependent
// launch, and a load through L1 (.ca, or a plain ld) can return a line of the PREVIOUS call's table (tools/check_tables.py)
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}
// copy elements [first, first + n) of a global array of 4- or 8-byte elements into sdst such that element `first` lands
// at sdst[first & (16 / size - 1)]: whole 16-byte chunks from 16-byte-aligned addresses, chunk i by thread t0 + i
template <typename T>
__device__ __forceinline__ void cp_async_range(T* sdst, const T* gbase, size_t first, int n, int t, int nthreads) {
  constexpr int per = 16 / (int)sizeof(T);
  const size_t c0 = first / per;
  const int chunks = (int)((first + n + per - 1) / per - c0);
  for (int i = t; i < chunks; i += nthreads) cp_async16(sdst + per * i, gbase + per * (c0 + i));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- mbarrier addresses --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bar_acc_full(uint32_t smem_base, int s) { return smem_base + OFF_BAR + 8 * s; }
__device__ __forceinline__ uint32_t bar_acc_empty(uint32_t smem_base, int s) { return smem_base + OFF_BAR + 8 * (kSlots + s); }
__device__ __forceinline__ uint32_t bar_x1_full(uint32_t smem_base, int b) { return smem_base + OFF_BAR + 8 * (2 * kSlots + b); }
__device__ __forceinline__ uint32_t bar_d2_full(uint32_t smem_base, int b) { return smem_base + OFF_BAR + 8 * (2 * kSlots + 2 + b); }
__device__ __forceinline__ uint32_t bar_x2_full(uint32_t smem_base) { return smem_base + OFF_BAR + 8 * (2 * kSlots + 4); }
// FCN accumulator of pipeline slot s = (tile ordinal) & 3: even ordinals use TMEM lanes 0..15 of every quadrant, odd
// ones lanes 16..31 (the two interleaved placements of an M=64 accumulator); the column range alternates every second tile
__device__ __forceinline__ uint32_t acc_addr(uint32_t tmem_base, int s) {
  return tmem_base + ((uint32_t)(16 * (s & 1)) << 16) + (uint32_t)(s >> 1) * kRows;
}

// ---- this CTA's tiles: chunks blockIdx.x, blockIdx.x + gridDim.x, ..., the tiles of each in order --------------
// The tables are written by this kernel's predecessors in the stream, which under programmatic dependent launch may
// still be running when this kernel starts: they are read with ld.global.cg AFTER griddepcontrol.wait. Never __ldg here:
// a non-coherent load may be hoisted above the wait or served from a stale line (seen: tile headers of one problem
// with the row tables of the next).
__device__ __forceinline__ int count_my_tiles(const VfeProblem& prob, int n_chunks) {
  int n = 0;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) n += __ldcg(prob.chunk_ntiles + c);
  return n;
}
struct TileCursor {  // walks every `step`-th tile of the CTA's sequence
  int c, j, n, n_next;  // chunk, tile inside it, its tile count; the NEXT chunk's tile count, loaded a chunk ahead so
                        // that crossing into it never waits for an L2 round trip (2-3 us beside the grid stream)
  __device__ __forceinline__ int load_n(const VfeProblem& prob, int n_chunks, int chunk) const {
    return chunk < n_chunks ? __ldcg(prob.chunk_ntiles + chunk) : 0;
  }
  __device__ __forceinline__ void settle(const VfeProblem& prob, int n_chunks) {
    while (c < n_chunks && j >= n) {
      j -= n;
      c += gridDim.x;
      n = n_next;
      n_next = load_n(prob, n_chunks, c + gridDim.x);
    }
  }
  __device__ __forceinline__ void init(const VfeProblem& prob, int n_chunks, int first) {
    c = blockIdx.x;
    n = load_n(prob, n_chunks, c);
    n_next = load_n(prob, n_chunks, c + gridDim.x);
    j = first;
    settle(prob, n_chunks);
  }
  __device__ __forceinline__ void advance(const VfeProblem& prob, int n_chunks, int step) {
    j += step;
    settle(prob, n_chunks);
  }
};

// ---- TENSOR stage: both GEMMs, issued by one thread (lane 0 of the tensor warp) ------------------------------
// 3xTF32, small terms first so that their sum is not rounded against the large one.
// VFE-2: D2[row][n] (+)= sum_k X1[row][k] * W2B[n][k], k over [pooled 16 | pointwise 16].
__device__ __forceinline__ void issue_vfe2_mma(uint32_t smem_base, int b, uint32_t d_tmem) {
  constexpr uint32_t idesc = umma::make_idesc_tf32_k(kRows, 64);
  const uint32_t xh = smem_base + OFF_X1 + (uint32_t)b * 2 * kXSlab, xl = xh + kXSlab;
  const uint32_t x[3] = {xl, xh, xh};
  const uint32_t w[3] = {smem_base + OFF_W2H, smem_base + OFF_W2L, smem_base + OFF_W2H};
  uint32_t acc = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {  // k-step kb: 32 bytes per step inside the swizzled 128-byte rows
      const uint64_t a = umma::make_desc_k_sw128(x[s] + kb * 32);
      const uint64_t bd = umma::make_desc_k_sw128(w[s] + kb * 32);
      umma::mma_tf32_ss(d_tmem, a, bd, idesc, acc);
      acc = 1;
    }
}
// FCN: D3^T[ch][n] (+)= sum_k W3^T[ch][k] * X2[n][k], k over [pooled 32 | pointwise 32].
__device__ __forceinline__ void issue_fcn_mma(uint32_t smem_base, uint32_t d_tmem) {
  constexpr uint32_t idesc = umma::make_idesc_tf32_k(64, kRows);
  const uint32_t w[3] = {smem_base + OFF_W3H, smem_base + OFF_W3L, smem_base + OFF_W3H};
  const uint32_t x[3] = {smem_base + OFF_X2L, smem_base + OFF_X2H, smem_base + OFF_X2H};
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

// The tensor thread serves three queues: X1 of team 0, X1 of team 1 (VFE-2 GEMMs, any order) and X2 (FCN GEMMs, in
// tile-ordinal order: the teams take turns on the single X2 buffer). It polls them without blocking on any.
__device__ __forceinline__ void tensor_stage(uint32_t smem_base, uint32_t tmem_base, int my_tiles) {
  const int n_team[2] = {(my_tiles + 1) >> 1, my_tiles >> 1};
  int k1[2] = {0, 0};  // next tile (team-local index) whose X1 is awaited
  int j = 0;           // next tile ordinal whose X2 is awaited
  unsigned idle = 0;
  while (k1[0] < n_team[0] || k1[1] < n_team[1] || j < my_tiles) {
    bool progress = false;
#pragma unroll
    for (int g = 0; g < 2; ++g)
      // all 8 warps of the team arrive after their tcgen05.ld of the previous D2[g] — which this GEMM overwrites — is done
      if (k1[g] < n_team[g] && umma::mbar_test_wait(bar_x1_full(smem_base, g), k1[g] & 1)) {
        umma::fence_after_sync();
        issue_vfe2_mma(smem_base, g, tmem_base + kD2Col0 + 64u * g);
        umma::mma_commit(bar_d2_full(smem_base, g));
        ++k1[g];
        progress = true;
      }
    if (j < my_tiles && umma::mbar_test_wait(bar_x2_full(smem_base), j & 1)) {
      const int s = j & (kSlots - 1);
      if (j < kSlots || umma::mbar_test_wait(bar_acc_empty(smem_base, s), ((j / kSlots) - 1) & 1)) {
        umma::fence_after_sync();
        issue_fcn_mma(smem_base, acc_addr(tmem_base, s));
        umma::mma_commit(bar_acc_full(smem_base, s));
        umma::mbar_arrive(bar_acc_full(smem_base, s));  // release: publishes the front stage's TileInfo to the back stage
        ++j;
        progress = true;
      }
    }
    if (progress) idle = 0;
    else {
      __nanosleep(32);
      if (++idle > (1u << 26)) __trap();  // a protocol bug must surface as an error, not as a hung GPU
    }
  }
}

// ---- BACK stage: accumulators -> per-voxel max -> BN + ReLU -> output row ------------------------------------
// Warp q of the stage reads TMEM lanes 32q..32q+31; an M=64 accumulator keeps channel c in lane (c % 16) + 32 (c / 16)
// (+16 for the interleaved placement), so lanes 0..15 of the warp own channels 16q..16q+15 of an EVEN tile and lanes
// 16..31 the same channels of the following ODD tile: the warp scans a pair of tiles per pass, every lane busy.
// Columns are tile rows: a thread walks its tile's rows in order, so the per-voxel max is a sequential scan; voxel ends
// come from the tile's last-row bit mask (per lane: the two halves of the warp work on different tiles).
// y = relu(a*z + b) is monotonic in z; the host folds sign(a) into dense_2's output column (api.cu), so a >= 0 here for
// every channel and max_rows relu(a*z_r + b) = relu(a*max_r z_r + b): one running max per thread, BN + ReLU once per voxel.
template <int MODE>
__device__ __forceinline__ void back_stage(const VfeSmall& P, const VfeOutput& out, unsigned char* smem,
                                           uint32_t smem_base, int my_tiles, uint32_t tmem_base, int bwarp, int lane) {
  const int ch = 16 * bwarp + (lane & 15), hh = lane >> 4;
  const float a = P.a3[ch], b = P.b3[ch];  // a = |BN scale|, see above
  const uint32_t tlane = tmem_base + ((uint32_t)(32 * bwarp) << 16);
  for (int p = 0; 2 * p < my_tiles; ++p) {
    const int it = 2 * p + hh;           // this lane's tile
    const bool valid = it < my_tiles;
    const int s_even = (2 * p) & (kSlots - 1), s = valid ? (it & (kSlots - 1)) : s_even;
    umma::mbar_wait(bar_acc_full(smem_base, s_even), ((2 * p) / kSlots) & 1);
    if (2 * p + 1 < my_tiles) umma::mbar_wait(bar_acc_full(smem_base, s_even + 1), ((2 * p + 1) / kSlots) & 1);
    umma::fence_after_sync();
    const TileInfo* info = reinterpret_cast<const TileInfo*>(smem + OFF_INFO) + s;
    const int nrows = valid ? info->nrows : 0;
    const int nrows_max = max(nrows, __shfl_xor_sync(0xffffffffu, nrows, 16));
    float mx = -INFINITY;
    int v = 0;
    int cell = MODE != 0 ? info->voxcell[0] : 0;
    float* row0 = MODE == 0 ? out.voxel_feat + (size_t)info->v0 * 64 + ch : nullptr;
    const uint32_t tcol = tlane + (uint32_t)(p & 1) * kRows;
#pragma unroll 1
    for (int c0 = 0; c0 < nrows_max; c0 += 32) {
      unsigned m = valid ? info->last_mask[c0 >> 5] : 0u;
      float x[32];
      umma::tmem_ld_32x32(tcol + c0, x);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx = fmaxf(mx, x[i]);
        if ((m >> i) & 1u) {  // the voxel's last row (uniform over the 16 lanes that share the tile)
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
    umma::fence_before_sync();
    __syncwarp();
    if (lane == 0) {
      umma::mbar_arrive(bar_acc_empty(smem_base, s_even));
      if (2 * p + 1 < my_tiles) umma::mbar_arrive(bar_acc_empty(smem_base, s_even + 1));
    }
  }
}

template <int MODE, typename PT>
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
    const float4* src = reinterpret_cast<const float4*>(wblob);  // [W3 hi | W3 lo | W2B hi | W2B lo] operand images (api.cu)
    float4* dst = reinterpret_cast<float4*>(smem + OFF_W3H);
    for (int i = threadIdx.x; i < kVfeBlobFloats / 4; i += kCtaThreads) dst[i] = __ldg(src + i);
    FrontParams* fp = reinterpret_cast<FrontParams*>(smem + OFF_PAR);
    for (int i = threadIdx.x; i < 96; i += kCtaThreads) {
      fp->w1f[i / 16][i % 16] = P.w1f[i / 16][i % 16];
      if (i < 48) fp->w1d[i / 16][i % 16] = P.w1d[i / 16][i % 16];
      if (i < 16) {
        fp->a1[i] = P.a1[i];
        fp->b1[i] = P.b1[i];
      }
      if (i < 32) {
        fp->a2[i] = P.a2[i];
        fp->b2[i] = P.b2[i];
      }
    }
    umma::fence_async_smem();  // the operand images are read by the tensor core
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      umma::mbar_init(bar_acc_full(smem_base, s), 2);  // tcgen05.commit + the issuing thread's own (release) arrive
      umma::mbar_init(bar_acc_empty(smem_base, s), kBackThreads / 32);
    }
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(bar_x1_full(smem_base, b), kTeamWarps);
      umma::mbar_init(bar_d2_full(smem_base, b), 1);
    }
    umma::mbar_init(bar_x2_full(smem_base), kTeamWarps);
    umma::mbar_init_fence();
  }
  if (warp_in_cta == kBackWarp0) umma::tmem_alloc<kTmemCols>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // from here on: the grouping, the row tables and the occupancy map of this call
  if (MODE != 0) timeline_stamp(g_trace, TL_VFE);
  const int n_chunks = (int)__ldcg(prob.n_chunks);  // (written by scan_down of this call: through L2, like every table)
  // chunks are strided over the CTAs; this one owns tile ordinals 0 .. my_tiles-1 = the tiles of its chunks in order
  const int my_tiles = count_my_tiles(prob, n_chunks);

  if (warp_in_cta < kWriterWarps) {  // ---- WRITER ----
    if (MODE == 1)
      background_writer(out.cell_voxel, out.c_empty, static_cast<float*>(out.grid), out.ncells, out.first_group,
                        out.writer_claim, smem + OFF_BG, (int)threadIdx.x);
    if (MODE == 2)
      background_writer(out.cell_voxel, out.c_empty, static_cast<__nv_bfloat16*>(out.grid), out.ncells,
                        out.first_group, out.writer_claim, smem + OFF_BG, (int)threadIdx.x);
    if (MODE != 0 && threadIdx.x == 0) {  // debug timeline: when this CTA's background was done
      unsigned long long* tr = g_trace;
      if (tr) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        atomicMin(tr + (size_t)(kTimelineRow0 + TL_WRITER_END) * kTraceSlots, now);  // earliest CTA
        atomicMax(tr + (size_t)(kTimelineRow0 + TL_WRITER_END) * kTraceSlots + 1, now);
      }
    }
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
  // Two teams of 256 threads, each on its own 128-row tile: thread = (row, part); row = TMEM lane, part = which half of
  // the channels. Team g owns tile ordinals i = 2k + g of this CTA.
  const int tid = threadIdx.x - 32 * kFrontWarp0, team = tid >> 8, ttid = tid & (kTeamThreads - 1);
  const int lane = ttid & 31, tw = ttid >> 5;
  const int row = 32 * (tw & 3) + lane, part = tw >> 2;
  unsigned char* team_base = smem + OFF_TEAM + team * kTeamBytes;
  float* sH = reinterpret_cast<float*>(team_base);
  PT* sXYZbuf = reinterpret_cast<PT*>(team_base + kHBytes);
  TileMeta* metas = reinterpret_cast<TileMeta*>(team_base + kHBytes + kXyzBytes);
  int4* hdr_ring = reinterpret_cast<int4*>(team_base + kHBytes + kXyzBytes + kMetaSlots * sizeof(TileMeta));
  const FrontParams* fp = reinterpret_cast<const FrontParams*>(smem + OFF_PAR);
  const PT* g_xyz = static_cast<const PT*>(prob.row_xyz);
  const uint32_t tlane = tmem_base + ((uint32_t)(32 * (tw & 3)) << 16);
  unsigned char* x1h = smem + OFF_X1 + (size_t)team * 2 * kXSlab;
  unsigned char* x1l = x1h + kXSlab;
  const int n_mine = (my_tiles + 1 - team) >> 1;  // tiles of this team

  // tile header = (first voxel, first row) of the tile and of its successor; rows and voxels are contiguous
  // tile header = (first voxel, end voxel, first row, end row); rows and voxels of a tile are contiguous. ONE thread of
  // the team walks the CTA's tile sequence (ordinals team, team + 2, ...) two tiles ahead and leaves the headers in a
  // ring in shared memory; the other 255 threads carry no cursor state (registers are what this kernel is short of).
  struct Header { int v0, v1, r0, r1; };
  const bool walker = ttid == kTeamThreads - 1;
  TileCursor cursor;
  if (walker) cursor.init(prob, n_chunks, team);
  // walker only: header of the team's tile k -> ring slot k & 3, by an ASYNCHRONOUS 16-byte copy of the tile's record
  // (tile_hdr, written by the tile plan). With four loads and a store the walker waited a whole L2 round trip here at the
  // top of every tile — 2-3 us beside the grid stream — and its team for it at the next barrier; the copy lands with the
  // team's other cp.async traffic (same commit group, waited for at the end of the iteration, a tile before it is read).
  auto publish_header = [&](int k) {
    if (cursor.c < n_chunks) {
      cp_async16(&hdr_ring[k & 3], reinterpret_cast<const int4*>(prob.tile_hdr) + (cursor.c * kChunkSlots + cursor.j));
      cursor.advance(prob, n_chunks, 2);
    } else {
      hdr_ring[k & 3] = make_int4(0, 0, 0, 0);
    }
  };
  auto read_header = [&](int k) {
    const int4 h = hdr_ring[k & 3];
    return Header{h.x, h.y, h.z, h.w};
  };
  // asynchronous copy of a tile's points (contiguous in row order) and row / voxel tables into the landing buffers
  auto prefetch = [&](const Header& h, int k) {
    TileMeta* m = metas + (k & 1);
    const int nrows = h.r1 - h.r0, nv = h.v1 - h.v0;
    // warps 0-5: the points; warp 6: row -> voxel; warp 7: the voxel tables
    if (ttid < 192) cp_async_range(sXYZbuf, g_xyz, 3 * (size_t)h.r0, 3 * nrows, ttid, 192);
    else if (ttid < 224) cp_async_range(m->rowvox, prob.row_voxel, (size_t)h.r0, nrows, ttid - 192, 32);
    else {
      cp_async_range(m->vrs, prob.row_start, (size_t)h.v0, nv + 1, ttid - 224, 32);
      if (MODE != 0) cp_async_range(m->voxcell, out.voxel_cell, (size_t)h.v0, nv, ttid - 224, 32);
    }
  };

  if (walker) {
    publish_header(0);
    publish_header(1);
    cp_async_commit();
    cp_async_wait_all();  // the first two are needed at once
  }
  team_sync(team);
  if (n_mine > 0) prefetch(read_header(0), 0);
  cp_async_commit();
  cp_async_wait_all();
  team_sync(team);

  for (int k = 0; k < n_mine; ++k) {
    const int i = 2 * k + team;  // tile ordinal of this CTA
    const Header h = read_header(k);
    if (walker) publish_header(k + 2);  // (ring slot (k + 2) & 3 was last read two iterations ago)
    const TileMeta& mt = metas[k & 1];
    const int nrows = h.r1 - h.r0, nv = h.v1 - h.v0;
    // the landing buffers start at the 16-byte chunk that holds the tile's first element
    struct View { const int *rowvox, *vrs, *voxcell; } m;
    m.rowvox = mt.rowvox + (h.r0 & 3);
    m.vrs = mt.vrs + (h.v0 & 3);
    m.voxcell = mt.voxcell + (h.v0 & 3);
    const PT* sXYZ = sXYZbuf + (3 * (size_t)h.r0) % (16 / sizeof(PT));
#ifdef LISEC_DEBUG_CHECKS
    if (ttid == 0 && (nrows <= 0 || nrows > kRows || nv <= 0 || nv > kVox))
      printf("BAD HEADER cta %d team %d k %d/%d: v %d..%d r %d..%d chunks %d my_tiles %d\n", (int)blockIdx.x, team, k, n_mine,
             h.v0, h.v1, h.r0, h.r1, n_chunks, my_tiles);
    if (row < nrows) {
      const int rv = m.rowvox[row] & ~kRowPadFlag;
      if (rv < h.v0 || rv >= h.v1) printf("BAD ROWVOX cta %d team %d k %d row %d rv %d v %d..%d\n", (int)blockIdx.x, team, k, row, rv, h.v0, h.v1);
      else {
        const int rs = m.vrs[rv - h.v0] - h.r0, re = m.vrs[rv - h.v0 + 1] - h.r0;
        if (rs < 0 || re > nrows || rs >= re || row < rs || row >= re)
          printf("BAD VRS cta %d team %d k %d row %d lv %d rs %d re %d nrows %d\n", (int)blockIdx.x, team, k, row, rv - h.v0, rs, re, nrows);
      }
    }
#endif

    // ---- F1: centroid, features, VFE-1, max-pool, X1 ------------------------------------------------------------
    {
      // addVFELayer(in, 6, 32) (:231 -> :155-166): Dense(6->16, no bias) + BN + ReLU; this thread: 8 of the 16 outputs
      float hv[8];
      const float4 a1a = *reinterpret_cast<const float4*>(fp->a1 + 8 * part), a1b = *reinterpret_cast<const float4*>(fp->a1 + 8 * part + 4);
      const float4 b1a = *reinterpret_cast<const float4*>(fp->b1 + 8 * part), b1b = *reinterpret_cast<const float4*>(fp->b1 + 8 * part + 4);
      const float av[8] = {a1a.x, a1a.y, a1a.z, a1a.w, a1b.x, a1b.y, a1b.z, a1b.w};
      const float bv[8] = {b1a.x, b1a.y, b1a.z, b1a.w, b1b.x, b1b.y, b1b.z, b1b.w};
      const int rv = row < nrows ? m.rowvox[row] : kRowPadFlag;
      if (rv & kRowPadFlag) {
        // the virtual pad row: a zero input row (:141) -> relu(b); rows past the tile's end get the same (never read)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[j] = fmaxf(bv[j], 0.f);
      } else {
        const int lv = rv - h.v0;
        const int rs = m.vrs[lv] - h.r0, re = m.vrs[lv + 1] - h.r0;
        const int n = re - rs - ((m.rowvox[re - 1] & kRowPadFlag) ? 1 : 0);  // kept points of the voxel
        const PT x = sXYZ[3 * row], y = sXYZ[3 * row + 1], z = sXYZ[3 * row + 2];
        // np.mean(currPoints, axis=0) (:135): float64 adds in list order from the additive identity, one divide
        double cx = (double)x, cy = (double)y, cz = (double)z;
        if (n > 1) {
          double sx = 0.0, sy = 0.0, sz = 0.0;
          for (int r = rs; r < rs + n; ++r) {
            sx += (double)sXYZ[3 * r];
            sy += (double)sXYZ[3 * r + 1];
            sz += (double)sXYZ[3 * r + 2];
          }
          const double dn = (double)n;
          cx = sx / dn;
          cy = sy / dn;
          cz = sz / dn;
        }
        float f[6];  // [x, y, z, x - cx, y - cy, z - cz] (:137-140), float32 as the Keras input cast leaves them
        point_features((double)x, (double)y, (double)z, cx, cy, cz, f);
        // exact split of the raw coordinates: coarse part (top 11 mantissa bits) + rest
        float l[3];
        double od[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float o = __uint_as_float(__float_as_uint(f[c]) & 0xffffe000u);
          l[c] = __fsub_rn(f[c], o);
          od[c] = (double)o;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * part + j;
          const double big = fma(od[2], fp->w1d[2][c], fma(od[1], fp->w1d[1][c], od[0] * fp->w1d[0][c]));
          const float bh = __double2float_rn(big), bl = __double2float_rn(big - (double)bh);
          float sm = __fmul_rn(f[5], fp->w1f[5][c]);
          sm = __fmaf_rn(f[4], fp->w1f[4][c], sm);
          sm = __fmaf_rn(f[3], fp->w1f[3][c], sm);
          sm = __fmaf_rn(l[2], fp->w1f[2][c], sm);
          sm = __fmaf_rn(l[1], fp->w1f[1][c], sm);
          sm = __fmaf_rn(l[0], fp->w1f[0][c], sm);
          const float d = __fadd_rn(bh, __fadd_rn(bl, sm));
          hv[j] = fmaxf(fmaf(d, av[j], bv[j]), 0.f);
        }
      }
      const float lo4[4] = {hv[0], hv[1], hv[2], hv[3]}, hi4[4] = {hv[4], hv[5], hv[6], hv[7]};
      *reinterpret_cast<float4*>(sH + h1_off(row, 2 * part)) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      *reinterpret_cast<float4*>(sH + h1_off(row, 2 * part + 1)) = make_float4(hv[4], hv[5], hv[6], hv[7]);
      // pointwise half of the VFE-2 input: X1[row][16 + 8 part ..]. The team's previous VFE-2 GEMM has long finished
      // reading X1 (its result was consumed in F2 of the previous tile).
      store_split4(x1h, x1l, x_chunk(row, 4 + 2 * part), lo4);
      store_split4(x1h, x1l, x_chunk(row, 5 + 2 * part), hi4);
    }
    team_sync(team);
    // the point buffer and the other landing buffer are free: the team's next tile is on its way while this one is worked on
    if (k + 1 < n_mine) prefetch(read_header(k + 1), k + 1);
    cp_async_commit();
    {  // MaxPoolingVFELayer over T (:160) + RepeatLayer + the pooled half of Concatenate (:164-165): item = (voxel, 4 channels)
      const int v = ttid >> 2, c = ttid & 3;
      if (v < nv) {
        const int rs = m.vrs[v] - h.r0, re = m.vrs[v + 1] - h.r0;
        float4 mx = *reinterpret_cast<const float4*>(sH + h1_off(rs, c));
        for (int r = rs + 1; r < re; ++r) {
          const float4 t = *reinterpret_cast<const float4*>(sH + h1_off(r, c));
          mx.x = fmaxf(mx.x, t.x); mx.y = fmaxf(mx.y, t.y); mx.z = fmaxf(mx.z, t.z); mx.w = fmaxf(mx.w, t.w);
        }
        float4 hi, lo;
        umma::tf32_split(mx.x, hi.x, lo.x);
        umma::tf32_split(mx.y, hi.y, lo.y);
        umma::tf32_split(mx.z, hi.z, lo.z);
        umma::tf32_split(mx.w, hi.w, lo.w);
        for (int r = rs; r < re; ++r) {
          const uint32_t off = x_chunk(r, c);
          *reinterpret_cast<float4*>(x1h + off) = hi;
          *reinterpret_cast<float4*>(x1l + off) = lo;
        }
      }
    }
    umma::fence_before_sync();  // this warp's tcgen05.ld of the previous D2[team] -> before the GEMM that overwrites it
    umma::fence_async_smem();   // X1 (generic-proxy writes) -> visible to the tensor core
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(bar_x1_full(smem_base, team));

    // ---- F2: D2 -> VFE-2 output, max-pool, X2, TileInfo ----------------------------------------------------------
    const int slot = i & (kSlots - 1);
    // (the GEMM was issued after all 8 warps of the team had arrived, i.e. after every pool-1 item had read sH as sH1:
    // sH may be overwritten as sH2 from here on)
    umma::mbar_wait(bar_d2_full(smem_base, team), k & 1);
    umma::fence_after_sync();
    {
      // addVFELayer(., 32, 64) (:232): Dense(32->32) + BN + ReLU; this thread: channels 16 part .. 16 part + 15
      float hv[16];
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        float q[8], xw[8];
        const uint32_t d2 = tlane + kD2Col0 + 64u * team + 16u * part + 8u * g8;
        umma::tmem_ld_2x8(d2, d2 + 32u, q, xw);
        const float* ap = fp->a2 + 16 * part + 8 * g8;
        const float* bp = fp->b2 + 16 * part + 8 * g8;
        const float4 a2a = *reinterpret_cast<const float4*>(ap), a2b = *reinterpret_cast<const float4*>(ap + 4);
        const float4 b2a = *reinterpret_cast<const float4*>(bp), b2b = *reinterpret_cast<const float4*>(bp + 4);
        const float av[8] = {a2a.x, a2a.y, a2a.z, a2a.w, a2b.x, a2b.y, a2b.z, a2b.w};
        const float bv[8] = {b2a.x, b2a.y, b2a.z, b2a.w, b2b.x, b2b.y, b2b.z, b2b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[8 * g8 + j] = fmaxf(fmaf(__fadd_rn(q[j], xw[j]), av[j], bv[j]), 0.f);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(sH + h2_off(row, 4 * part + c)) = make_float4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
      // X2 is shared by the teams, which take turns: the previous tile's FCN (the other team's) must have finished reading it
      if (i > 0) umma::mbar_wait(bar_acc_full(smem_base, (i - 1) & (kSlots - 1)), ((i - 1) / kSlots) & 1);
      // pointwise half of the FCN input: X2[row][32 + 16 part ..] (slab 1)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v4[4] = {hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]};
        store_split4(smem + OFF_X2H + kXSlab, smem + OFF_X2L + kXSlab, x_chunk(row, 4 * part + c), v4);
      }
    }
    team_sync(team);
#pragma unroll 1
    for (int it = ttid; it < 8 * kVox; it += kTeamThreads) {  // pooled half (slab 0): item = (voxel, 4 channels), two per thread
      const int v = it >> 3, c = it & 7;
      if (v < nv) {
        const int rs = m.vrs[v] - h.r0, re = m.vrs[v + 1] - h.r0;
        float4 mx = *reinterpret_cast<const float4*>(sH + h2_off(rs, c));
        for (int r = rs + 1; r < re; ++r) {
          const float4 t = *reinterpret_cast<const float4*>(sH + h2_off(r, c));
          mx.x = fmaxf(mx.x, t.x); mx.y = fmaxf(mx.y, t.y); mx.z = fmaxf(mx.z, t.z); mx.w = fmaxf(mx.w, t.w);
        }
        float4 hi, lo;
        umma::tf32_split(mx.x, hi.x, lo.x);
        umma::tf32_split(mx.y, hi.y, lo.y);
        umma::tf32_split(mx.z, hi.z, lo.z);
        umma::tf32_split(mx.w, hi.w, lo.w);
        for (int r = rs; r < re; ++r) {
          const uint32_t off = x_chunk(r, c);
          *reinterpret_cast<float4*>(smem + OFF_X2H + off) = hi;
          *reinterpret_cast<float4*>(smem + OFF_X2L + off) = lo;
        }
      }
    }
    // the slot's TileInfo is free once the back stage has drained the slot's previous use
    if (i >= kSlots) umma::mbar_wait(bar_acc_empty(smem_base, slot), ((i / kSlots) - 1) & 1);
    if (ttid < kRows) {  // team warps 0-3: ttid == row
      TileInfo* info = reinterpret_cast<TileInfo*>(smem + OFF_INFO) + slot;
      const int my = m.rowvox[ttid] & ~kRowPadFlag;
      const int next = ttid + 1 < nrows ? (m.rowvox[ttid + 1] & ~kRowPadFlag) : -1;
      const unsigned last = __ballot_sync(0xffffffffu, ttid < nrows && my != next);
      if (lane == 0) info->last_mask[tw] = last;
      if (MODE != 0 && ttid < nv) info->voxcell[ttid] = m.voxcell[ttid];
      if (ttid == 0) {
        info->nrows = nrows;
        info->nv = nv;
        info->v0 = h.v0;
      }
    }
    umma::fence_async_smem();  // X2 (generic-proxy writes) -> visible to the tensor core
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(bar_x2_full(smem_base));  // 8 warps -> the tensor thread issues this tile's FCN
    cp_async_wait_all();
    team_sync(team);  // publishes the landing buffers and the walker's header; sH may be overwritten
  }
  asm volatile("bar.sync 3, %0;" ::"n"(kFrontThreads + kBackThreads) : "memory");
  if (MODE != 0 && tid == 0) {  // (threadIdx.x != 0 here: stamp by hand)
    unsigned long long* tr = g_trace;
    if (tr) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      atomicMin(tr + (size_t)(kTimelineRow0 + TL_VFE_END) * kTraceSlots, now);  // earliest CTA
      atomicMax(tr + (size_t)(kTimelineRow0 + TL_VFE_END) * kTraceSlots + 1, now);
    }
  }
}

}  // namespace

cudaError_t set_trace_vfe(unsigned long long* trace) { return cudaMemcpyToSymbol(g_trace, &trace, sizeof(trace)); }

template <int MODE, typename PT>
static cudaError_t launch_vfe_mode(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const VfeOutput& out,
                                   int sm_count, cudaStream_t st) {
  cudaError_t err = cudaFuncSetAttribute(vfe_kernel<MODE, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (err != cudaSuccess) return err;
  // persistent: one CTA per SM (3 writer warps, 1 tensor warp, 4 back warps, 16 front warps), tiles strided over the CTAs
  return launch_pdl(vfe_kernel<MODE, PT>, sm_count, kCtaThreads, kSmemBytes, st, p, wblob, prob, out);
}
template <int MODE>
static cudaError_t launch_vfe_dtype(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const VfeOutput& out,
                                    int sm_count, cudaStream_t st) {
  return prob.pts_dtype == LISEC_F32 ? launch_vfe_mode<MODE, float>(p, wblob, prob, out, sm_count, st)
                                     : launch_vfe_mode<MODE, double>(p, wblob, prob, out, sm_count, st);
}

cudaError_t launch_vfe(const VfeSmall& p, const float* wblob, const VfeProblem& prob, float* voxel_feat, int sm_count,
                       cudaStream_t st, int* launches, long long*) {
  const VfeOutput out{voxel_feat, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr};
  ++*launches;
  return launch_vfe_dtype<0>(p, wblob, prob, out, sm_count, st);
}

cudaError_t launch_vfe_to_grid(const VfeSmall& p, const float* wblob, const VfeProblem& prob, const Workspace& w,
                               const Geom& g, int n_sweeps, int grid_dtype, void* grid, int first_group, int sm_count,
                               cudaStream_t st, int* launches) {
  const VfeOutput out{nullptr, grid, w.voxel_cell, w.cell_voxel, w.c_empty, (long long)n_sweeps * g.cells, w.count,
                      first_group, w.writer_claim};
  ++*launches;
  return grid_dtype == LISEC_F32 ? launch_vfe_dtype<1>(p, wblob, prob, out, sm_count, st)
                                 : launch_vfe_dtype<2>(p, wblob, prob, out, sm_count, st);
}

}  // namespace lisec
