// Point -> voxel grouping on sm_100a.
//
// Replaces reference model_training.py:103-126 (get_voxel + the dict-building loop of VFE_preprocessing) and the
// sampling rule at :131-132 with its deterministic contract (first T indices in point order).
//
// Pipeline (all launches on the caller's stream, no host round trip):
//   point_pass   12 B/point read, float4-vectorised; voxel key in float64; strict range test; warp-aggregated
//                atomicAdd histogram into the per-cell count table; writes cell_of_point.
//   cell_scan    two kernels (block totals; prefix of earlier blocks + scan): three fused exclusive scans over the cell table (occupied -> voxel row, count -> CSR offset,
//                kept+pad -> VFE row offset); writes the occupancy map cell_voxel. Voxel rows therefore come out
//                in ascending (sweep, z, x, y) cell order, independent of thread scheduling.
//   fill_pass    warp-aggregated atomicSub slot claim drains the count table back to zero (so the next call
//                needs no memset) and writes the CSR payload in arrival order.
//   order_pass   rank-by-counting inside each voxel segment, early exit at T: entry p lands at position
//                #{q in voxel : q < p}. This is what makes the slot assignment deterministic in point order
//                whatever order the atomics resolved in. Also marks the VFE tile boundaries and the row -> voxel table.
#include <cstdlib>

#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

namespace {

__device__ unsigned long long* g_trace = nullptr;  // debug timeline (set_trace_voxelize), normally null

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- voxel key (model_training.py:103-107, 117-122) ------------------------------------------------------
// status: 0 kept, 1 out of range, 2 non-finite (the reference would raise in math.floor; we drop and count).
__device__ __forceinline__ int cell_of(double x, double y, double z, const Geom& g, int& status) {
  if (!(isfinite(x) && isfinite(y) && isfinite(z))) {
    status = 2;
    return -1;
  }
  double kx, ky, kz;
  if (g.exact_inv) {  // power-of-two voxel sizes: the product is exact, so it equals the quotient bit for bit
    kx = floor(x * g.inv[0]);
    ky = floor(y * g.inv[1]);
    kz = floor(z * g.inv[2]);
  } else {
    kx = floor(x / g.size[0]);
    ky = floor(y / g.size[1]);
    kz = floor(z / g.size[2]);
  }
  // strict on both sides (:118-120); comparisons stay in float64 so huge coordinates cannot wrap an int
  const bool keep = (kx > -(double)g.maxx) && (kx < (double)g.maxx) && (ky > -(double)g.maxy) &&
                    (ky < (double)g.maxy) && (kz > 0.0) && (kz < (double)g.maxz);
  if (!keep) {
    status = 1;
    return -1;
  }
  status = 0;
  // fixedKey = (kx + maxVoxelX, ky + maxVoxelY, kz) (:122); dense layout is (z, x, y) (:148, :151-152)
  return ((int)kz * g.nx + ((int)kx + g.maxx)) * g.ny + ((int)ky + g.maxy);
}

template <typename PT>
struct Vec4Load;
template <>
struct Vec4Load<float> {
  // 4 points = 12 floats = three 16-byte loads. L2 loads (.cg): the points are rewritten between calls (a caller's buffer,
  // the alternating staging buffers of the host entry point) and this kernel is launched under programmatic dependent
  // launch, where a line still sitting in an SM's L1 from an earlier call has been seen to be served again;
  // the points are read once, so L1 has nothing to give here anyway.
  static __device__ __forceinline__ void load(const float* base, long long g, float (&v)[12]) {
    const float4* p = reinterpret_cast<const float4*>(base) + 3 * g;
    float4 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
  }
};
template <>
struct Vec4Load<double> {
  static __device__ __forceinline__ void load(const double* base, long long g, double (&v)[12]) {
    const double2* p = reinterpret_cast<const double2*>(base) + 6 * g;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double2 a = __ldcg(p + i);
      v[2 * i] = a.x;
      v[2 * i + 1] = a.y;
    }
  }
};

// ---- K1: point pass ---------------------------------------------------------------------------------------
template <typename PT>
__global__ void __launch_bounds__(256) point_pass_kernel(const PT* __restrict__ pts, long long n_total,
                                                         const __grid_constant__ SweepOffsets so,
                                                         const __grid_constant__ Geom g,
                                                         int* __restrict__ cell_of_point, int* __restrict__ count,
                                                         unsigned long long* __restrict__ totals) {
  __shared__ int s_drop[2];
  pdl_launch_dependents();
  if (threadIdx.x == 0) s_drop[0] = s_drop[1] = 0;
  __syncthreads();
  pdl_wait();
  timeline_stamp(g_trace, TL_POINT);
  const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p0 = grp * 4;
  PT v[12];
  const bool full = p0 + 3 < n_total;
  if (full) {
    Vec4Load<PT>::load(pts, grp, v);
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      long long idx = p0 * 3 + i;
      v[i] = (idx < n_total * 3) ? pts[idx] : PT(0);
    }
  }
  // sweep of the first point; later points of the group advance it (groups may straddle sweeps)
  int s = 0;
  if (p0 < n_total) {
    while (s + 1 < so.n && p0 >= so.off[s + 1]) ++s;
  }
  int cell[4];
  int n_oor = 0, n_nonf = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long p = p0 + j;
    cell[j] = -1;
    if (p < n_total) {
      while (s + 1 < so.n && p >= so.off[s + 1]) ++s;
      int status;
      int c = cell_of((double)v[3 * j], (double)v[3 * j + 1], (double)v[3 * j + 2], g, status);
      n_oor += (status == 1);
      n_nonf += (status == 2);
      cell[j] = (c >= 0) ? s * g.cells + c : -1;
    }
  }
  if (full) {
    reinterpret_cast<int4*>(cell_of_point)[grp] = make_int4(cell[0], cell[1], cell[2], cell[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p0 + j < n_total) cell_of_point[p0 + j] = cell[j];
  }
  // warp-aggregated histogram: lanes that hit the same cell elect one leader that adds the group size
  const int lane = lane_id();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned active = __ballot_sync(0xffffffffu, cell[j] >= 0);
    if (cell[j] >= 0) {
      const unsigned peers = __match_any_sync(active, cell[j]);
      if (lane == __ffs(peers) - 1) atomicAdd(&count[cell[j]], __popc(peers));
    }
  }
  // dropped-point statistics: reduced over the warp, then over the block in shared memory — ONE global atomic per block
  // and counter. (One per warp was 6 250 read-modify-writes of the same L2 line per 8-sweep call, which the L2 slice that
  // owns the line takes one after the other; the kernel is not complete until they have drained.)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_oor += __shfl_xor_sync(0xffffffffu, n_oor, o);
    n_nonf += __shfl_xor_sync(0xffffffffu, n_nonf, o);
  }
  if (lane == 0) {
    if (n_oor) atomicAdd(&s_drop[0], n_oor);
    if (n_nonf) atomicAdd(&s_drop[1], n_nonf);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_drop[0]) atomicAdd(&totals[TOT_ACC_OUT_OF_RANGE], (unsigned long long)s_drop[0]);
    if (s_drop[1]) atomicAdd(&totals[TOT_ACC_NONFINITE], (unsigned long long)s_drop[1]);
  }
}

// ---- K2: cell-table scans ---------------------------------------------------------------------------------
struct Tri {
  int v, e, r;  // voxels, CSR entries, VFE rows
};
__device__ __forceinline__ Tri tri_add(Tri a, Tri b) { return Tri{a.v + b.v, a.e + b.e, a.r + b.r}; }
__device__ __forceinline__ Tri tri_of_count(int c, int T) {
  // rows = kept points + one pad row when the voxel is not full (all pad rows of a voxel are identical, so one
  // virtual zero row reproduces the unmasked reference exactly)
  return c > 0 ? Tri{1, c, (c < T ? c + 1 : T)} : Tri{0, 0, 0};
}
__device__ __forceinline__ Tri tri_shfl_up(Tri a, int d) {
  return Tri{__shfl_up_sync(0xffffffffu, a.v, d), __shfl_up_sync(0xffffffffu, a.e, d),
             __shfl_up_sync(0xffffffffu, a.r, d)};
}

// exclusive scan of one Tri per thread over a 256-thread block; *total = block sum
__device__ __forceinline__ Tri block_exclusive(Tri x, Tri* total, Tri* smem /*[8+1]*/) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  Tri inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Tri y = tri_shfl_up(inc, d);
    if (lane >= d) inc = tri_add(inc, y);
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    Tri w = lane < (kScanThreads / 32) ? smem[lane] : Tri{0, 0, 0};
    Tri winc = w;
#pragma unroll
    for (int d = 1; d < kScanThreads / 32; d <<= 1) {
      Tri y = tri_shfl_up(winc, d);
      if (lane >= d) winc = tri_add(winc, y);
    }
    if (lane < kScanThreads / 32) smem[lane] = Tri{winc.v - w.v, winc.e - w.e, winc.r - w.r};
    if (lane == kScanThreads / 32 - 1) smem[kScanThreads / 32] = winc;
  }
  __syncthreads();
  const Tri base = smem[warp];
  *total = smem[kScanThreads / 32];
  return Tri{base.v + inc.v - x.v, base.e + inc.e - x.e, base.r + inc.r - x.r};
}

// sum of one Tri per thread over the block (every thread gets it): the hardware warp reduction, then the 8 warp sums
__device__ __forceinline__ Tri block_sum(Tri x, Tri* smem /*[8+1]*/) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const Tri w{__reduce_add_sync(0xffffffffu, x.v), __reduce_add_sync(0xffffffffu, x.e), __reduce_add_sync(0xffffffffu, x.r)};
  if (lane == 0) smem[warp] = w;
  __syncthreads();
  if (warp == 0) {
    const Tri y = lane < kScanThreads / 32 ? smem[lane] : Tri{0, 0, 0};
    const Tri all{__reduce_add_sync(0xffffffffu, y.v), __reduce_add_sync(0xffffffffu, y.e), __reduce_add_sync(0xffffffffu, y.r)};
    if (lane == 0) smem[kScanThreads / 32] = all;
  }
  __syncthreads();
  return smem[kScanThreads / 32];
}

__device__ __forceinline__ void load_counts(const int* __restrict__ count, long long base, long long ncells,
                                            int (&c)[kScanItems]) {
  // (L2 loads: the table is written by the kernels in front under programmatic dependent launch)
  if (base + kScanItems <= ncells) {
    const int4* p = reinterpret_cast<const int4*>(count + base);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; ++q) {
      const int4 a = __ldcg(p + q);
      c[4 * q] = a.x; c[4 * q + 1] = a.y; c[4 * q + 2] = a.z; c[4 * q + 3] = a.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) c[i] = (base + i < ncells) ? __ldcg(count + base + i) : 0;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const int* __restrict__ count, long long ncells,
                                                                   int T, int nblocks, int group_offset,
                                                                   int* __restrict__ block_sums) {
  __shared__ Tri smem[kScanThreads / 32 + 1];
  pdl_launch_dependents();
  pdl_wait();
  timeline_stamp(g_trace, TL_SCAN_REDUCE);
  const long long base = ((long long)blockIdx.x * kScanThreads + threadIdx.x) * kScanItems;
  int c[kScanItems];
  load_counts(count, base, ncells, c);
  Tri t{0, 0, 0};
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) t = tri_add(t, tri_of_count(c[i], T));
  Tri total;
  block_exclusive(t, &total, smem);
  // Two-level totals: the block's own (sums[b]) and, by three fire-and-forget atomics, the running total of its GROUP of
  // kScanGroup consecutive blocks (grp[b / kScanGroup]; zeroed again by the order pass). A scan_down block then needs the
  // groups before its own plus the blocks before it inside its group: one load per thread, one round trip. (Every block
  // used to add up the totals of ALL blocks before it: 3 M loads = 50 MB of L2 traffic per 8-sweep call, up to three
  // dependent round trips in front of every block's own scan. A last-block-done spine inside this kernel was measured
  // too: it costs this kernel more (+6.8 us: a fence and a ticket per block, one block's serial tail) than it saves there.)
  if (threadIdx.x == 0) {
    int4* sums = reinterpret_cast<int4*>(block_sums);
    sums[blockIdx.x] = make_int4(total.v, total.e, total.r, 0);  // one 16-byte slot per block
    if (total.v) {
      int* grp = reinterpret_cast<int*>(sums + group_offset + (blockIdx.x / kScanGroup));
      atomicAdd(grp, total.v);
      atomicAdd(grp + 1, total.e);
      atomicAdd(grp + 2, total.r);
    }
  }
}

// Second pass. There is no separate "spine" kernel: a block gets its exclusive prefix from scan_reduce's two-level
// totals, and the last block, which thereby holds the grand totals, writes them and the sentinels.
__global__ void __launch_bounds__(kScanThreads) scan_down_kernel(const int* __restrict__ count, long long ncells,
                                                                 int T, int cells_per_sweep, int nblocks, int n_sweeps,
                                                                 int rows_per_chunk, int* __restrict__ chunk_first,
                                                                 const int* __restrict__ block_sums, int group_offset,
                                                                 long long* __restrict__ totals,
                                                                 int* __restrict__ cell_voxel,
                                                                 int* __restrict__ voxel_cell,
                                                                 int* __restrict__ voxel_start,
                                                                 int* __restrict__ row_start,
                                                                 int* __restrict__ sweep_voxel_start) {
  __shared__ Tri smem[kScanThreads / 32 + 1];
  pdl_launch_dependents();
  pdl_wait();
  timeline_stamp(g_trace, TL_SCAN_DOWN);
  const long long base = ((long long)blockIdx.x * kScanThreads + threadIdx.x) * kScanItems;
  int c[kScanItems];
  load_counts(count, base, ncells, c);
  Tri before{0, 0, 0};
  {
    // the sum over all earlier blocks = the groups before this block's group + the blocks before it inside the group:
    // one 16-byte L2 load per thread (the sums were written by scan_reduce under programmatic dependent launch)
    const int4* sums = reinterpret_cast<const int4*>(block_sums);
    const int4* grp = sums + group_offset;
    const int g = (int)blockIdx.x / kScanGroup, in_group = (int)blockIdx.x % kScanGroup;
    for (int j = threadIdx.x; j < g + in_group; j += kScanThreads) {
      const int4 a = __ldcg(j < g ? grp + j : sums + g * kScanGroup + (j - g));
      before = tri_add(before, Tri{a.x, a.y, a.z});
    }
  }
  const Tri prefix = block_sum(before, smem);  // the sum over all earlier blocks
  __syncthreads();                             // smem is reused below
  Tri t{0, 0, 0};
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) t = tri_add(t, tri_of_count(c[i], T));
  Tri total;
  Tri ex = block_exclusive(t, &total, smem);
  Tri run = tri_add(prefix, ex);
  if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) {
    const Tri carry = tri_add(prefix, total);
    totals[TOT_VOXELS] = carry.v;
    totals[TOT_ENTRIES] = carry.e;
    totals[TOT_ROWS] = carry.r;
    totals[TOT_CHUNKS] = carry.r > 0 ? (carry.r - 1) / rows_per_chunk + 1 : 0;
    // the point pass's drop counters (complete: that kernel has finished) become this call's figures and are left zero
    // for the next call — a call begins with a kernel, not with a memset
    totals[TOT_NONFINITE] = __ldcg(totals + TOT_ACC_NONFINITE);
    totals[TOT_OUT_OF_RANGE] = __ldcg(totals + TOT_ACC_OUT_OF_RANGE);
    totals[TOT_ACC_NONFINITE] = 0;
    totals[TOT_ACC_OUT_OF_RANGE] = 0;
    voxel_start[carry.v] = carry.e;
    row_start[carry.v] = carry.r;
    sweep_voxel_start[n_sweeps] = carry.v;
  }
  int cv[kScanItems];
  // cell ids fit 32 bits (checked in lisec_create); one division per thread finds the position inside the sweep
  const int cell0 = (int)base;
  int in_sweep = base < ncells ? cell0 % cells_per_sweep : 1;
  int sweep = base < ncells ? cell0 / cells_per_sweep : 0;
  // Most warps see nothing but empty cells (7.7 % of the cells of a Lyft-shaped sweep are occupied, in clusters): unless
  // some lane holds an occupied cell or the first cell of a sweep, the warp only has the map's "-1"s to write.
  const bool work = t.v > 0 || (base < ncells && (in_sweep == 0 || in_sweep + kScanItems > cells_per_sweep));
  if (!__any_sync(0xffffffffu, work)) {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) cv[i] = -1;
  } else {
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int cell = cell0 + i;
    cv[i] = -1;
    if (base + i < ncells) {
      if (in_sweep == 0) sweep_voxel_start[sweep] = run.v;
      if (c[i] > 0) {
        cv[i] = run.v;
        voxel_cell[run.v] = cell;
        voxel_start[run.v] = run.e;
        row_start[run.v] = run.r;
        // VFE chunks: chunk(v) = row_start[v] / rows_per_chunk. A voxel has at most T rows < rows_per_chunk, so the first
        // voxel of chunk t starts less than T rows into it: only those voxels (~7 %) compete for chunk_first[t]
        // (preset to a huge value). The chunk count comes from the row total (last block): it may name one trailing
        // chunk that no voxel starts in, which the tile plan leaves with zero tiles.
        const int t = run.r / rows_per_chunk;
        if (run.r - t * rows_per_chunk < T) atomicMin(chunk_first + t, run.v);
        run = tri_add(run, tri_of_count(c[i], T));
      }
    }
    if (++in_sweep == cells_per_sweep) {
      in_sweep = 0;
      ++sweep;
    }
  }
  }
  if (base + kScanItems <= ncells) {
    int4* p = reinterpret_cast<int4*>(cell_voxel + base);
#pragma unroll
    for (int q = 0; q < kScanItems / 4; ++q) p[q] = make_int4(cv[4 * q], cv[4 * q + 1], cv[4 * q + 2], cv[4 * q + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
      if (base + i < ncells) cell_voxel[base + i] = cv[i];
  }
}

// ---- K3: fill pass ----------------------------------------------------------------------------------------
// Tile plan (blocks past the fill blocks of fill_pass_kernel): one warp per chunk packs the chunk's voxels greedily
// into tiles of at most kVfeThreads rows (a tile = whole voxels): the next tile ends at the last voxel whose rows still
// fit. A voxel has >= 2 rows, so a tile has <= kVfeThreads / 2 voxels: two candidates per lane cover the search. Tiles
// ~93 % full instead of the 73 % a fixed row stride with its worst-case reserve (T - 1 rows) gives. The chunk's row
// offsets arrive in one round trip; the walk stays in shared memory. It needs only what scan_down left (row_start, the
// chunk marks), so it rides in the fill pass's launch: as a kernel of its own at the end of the chain it cost 13 us per
// step (launch, drain, and the VFE kernel's start-up behind it).
constexpr int kPlanVox = kVfeChunkRows / 2 + 4;  // row offsets of a chunk: <= kVfeChunkRows / 2 voxels + the end entry
__device__ __forceinline__ void plan_chunk_tiles(long long c, int lane, int* rs, const int* __restrict__ chunk_first,
                                                 const int* __restrict__ row_start,
                                                 const long long* __restrict__ totals, int* __restrict__ tile_first,
                                                 int* __restrict__ tile_row0, int* __restrict__ chunk_ntiles,
                                                 int* __restrict__ tile_hdr) {
  // (everything read here was written by scan_down under programmatic dependent launch: ld.global.cg, never through L1
  // — a plain load here returned lines of the PREVIOUS call's tables: tools/check_tables.py)
  const long long n_chunks = __ldcg(totals + TOT_CHUNKS);
  if (c >= n_chunks) return;
  const int V = (int)__ldcg(totals + TOT_VOXELS);
  const int v0 = min(__ldcg(chunk_first + c), V);  // (a trailing chunk no voxel starts in keeps the preset: empty)
  const int v_end = c + 1 < n_chunks ? min(__ldcg(chunk_first + c + 1), V) : V;
  const int nv = v_end - v0;
  for (int i = lane; i <= nv; i += 32) rs[i] = __ldcg(row_start + v0 + i);
  __syncwarp();
  int b = 0, j = 0;
  int* tf = tile_first + c * kChunkSlots;
  int* tr = tile_row0 + c * kChunkSlots;
  while (b < nv && j < kChunkSlots - 1) {
    const int base = rs[b];
    const int e1 = b + 1 + lane, e2 = e1 + 32;
    const bool ok1 = e1 <= nv && rs[e1] - base <= kVfeThreads;
    const bool ok2 = e2 <= nv && rs[e2] - base <= kVfeThreads;
    const int fit = __popc(__ballot_sync(0xffffffffu, ok1)) + __popc(__ballot_sync(0xffffffffu, ok2));  // monotone: a count
    if (lane == 0) {
      tf[j] = v0 + b;
      tr[j] = base;
      // the tile as one 16-byte record (first voxel, end voxel, first row, end row) for the VFE kernel's walkers
      reinterpret_cast<int4*>(tile_hdr)[c * kChunkSlots + j] = make_int4(v0 + b, v0 + b + fit, base, rs[b + fit]);
    }
    b += fit;  // fit >= 1: one voxel always fits
    ++j;
  }
  if (lane == 0) {
    tf[j] = v_end;
    tr[j] = rs[nv];
    chunk_ntiles[c] = j;
  }
}

__global__ void __launch_bounds__(256) fill_pass_kernel(const int* __restrict__ cell_of_point, long long n_total,
                                                        const int* __restrict__ cell_voxel,
                                                        const int* __restrict__ voxel_start,
                                                        int* __restrict__ count, int* __restrict__ list_unsorted,
                                                        int* __restrict__ entry_voxel, unsigned fill_blocks,
                                                        const int* __restrict__ chunk_first,
                                                        const int* __restrict__ row_start,
                                                        const long long* __restrict__ totals,
                                                        int* __restrict__ tile_first, int* __restrict__ tile_row0,
                                                        int* __restrict__ chunk_ntiles, int* __restrict__ tile_hdr) {
  __shared__ int s_rs[8][kPlanVox];
  pdl_launch_dependents();
  pdl_wait();
  timeline_stamp(g_trace, TL_FILL);
  if (blockIdx.x >= fill_blocks) {  // the tile plan's blocks: 8 chunks each
    const int wib = threadIdx.x >> 5;
    plan_chunk_tiles((long long)(blockIdx.x - fill_blocks) * 8 + wib, lane_id(), s_rs[wib], chunk_first, row_start, totals,
                     tile_first, tile_row0, chunk_ntiles, tile_hdr);
    return;
  }
  const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long p0 = grp * 4;
  int cell[4] = {-1, -1, -1, -1};
  if (p0 + 3 < n_total) {
    int4 c = __ldcg(reinterpret_cast<const int4*>(cell_of_point) + grp);  // (written by the point pass of this call: through L2)
    cell[0] = c.x; cell[1] = c.y; cell[2] = c.z; cell[3] = c.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (p0 + j < n_total) cell[j] = __ldcg(cell_of_point + p0 + j);
  }
  const int lane = lane_id();
  // the chain cell -> voxel -> segment start -> slot is three dependent L2 round trips: run the four points of the
  // thread through each stage together so the round trips overlap
  int v[4], start[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = cell[j] >= 0 ? __ldcg(cell_voxel + cell[j]) : -1;  // (written by the predecessor kernel: no __ldg under PDL)
#pragma unroll
  for (int j = 0; j < 4; ++j) start[j] = cell[j] >= 0 ? __ldcg(voxel_start + v[j]) : 0;
  unsigned peers[4];
  int old[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned active = __ballot_sync(0xffffffffu, cell[j] >= 0);
    peers[j] = 0;
    old[j] = 0;
    if (cell[j] >= 0) {
      peers[j] = __match_any_sync(active, cell[j]);
      if (lane == __ffs(peers[j]) - 1) old[j] = atomicSub(&count[cell[j]], __popc(peers[j]));  // drains the table to zero
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (cell[j] >= 0) {
      const int leader = __ffs(peers[j]) - 1;
      const int npeers = __popc(peers[j]);
      const int o = __shfl_sync(peers[j], old[j], leader);
      const int rank = __popc(peers[j] & ((1u << lane) - 1u));
      const int slot = o - npeers + rank;
      list_unsorted[start[j] + slot] = (int)(p0 + j);
      entry_voxel[start[j] + slot] = v[j];
    }
  }
}

// ---- K4: order pass (+ VFE tile boundaries, + the VFE row tables) -----------------------------------------
template <typename PT>
__global__ void __launch_bounds__(256) order_pass_kernel(const PT* __restrict__ pts,
                                                         const int* __restrict__ list_unsorted,
                                                         const int* __restrict__ entry_voxel,
                                                         const int* __restrict__ voxel_start,
                                                         const int* __restrict__ row_start, int T,
                                                         const long long* __restrict__ totals,
                                                         int* __restrict__ list_sorted,
                                                         int* __restrict__ row_voxel, PT* __restrict__ row_xyz,
                                                         int* __restrict__ chunk_first, long long chunk_cap,
                                                         int* __restrict__ scan_groups, int n_group_ints) {
  pdl_launch_dependents();
  pdl_wait();
  timeline_stamp(g_trace, TL_ORDER);
  // The totals are fetched by ONE thread per block: with a load per warp, 25 000 warps asked the same L2 slice for the
  // same 32 bytes, one after the other, before any of them could start. (Predecessor-written: ld.global.cg)
  __shared__ long long s_tot[2];
  if (threadIdx.x == 0) {
    s_tot[0] = __ldcg(totals + TOT_ENTRIES);
    s_tot[1] = __ldcg(totals + TOT_CHUNKS);
  }
  __syncthreads();
  const long long n_entries = s_tot[0];
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  {
    // The chunk marks (atomicMin targets of scan_down) have been consumed by the tile plan in the fill pass's launch:
    // put the preset back for the next call, so that no call starts with a memset.
    const long long used = min(s_tot[1] + 2, chunk_cap);
    for (long long i = e; i < used; i += (long long)gridDim.x * blockDim.x) chunk_first[i] = kChunkFirstPreset;
    // likewise the scans' group totals (accumulated by atomics in scan_reduce, read by scan_down)
    for (long long i = e; i < n_group_ints; i += (long long)gridDim.x * blockDim.x) scan_groups[i] = 0;
  }
  if (e >= n_entries) return;
  const int v = __ldcg(entry_voxel + e);
  const int s = __ldcg(voxel_start + v);
  const int n = __ldcg(voxel_start + v + 1) - s;
  const int row0 = __ldcg(row_start + v);
  const int p = __ldcg(list_unsorted + e);
  // the point itself, fetched while the rank is counted: the VFE kernel reads its rows' coordinates contiguously
  PT px, py, pz;
  load_point(pts, (long long)p, px, py, pz);
  int rank = 0;
  if (n > 1) {
    // chunks of 8 independent loads, then the early exit: not among the first T in point order = dropped (:131)
    for (int i0 = 0; i0 < n && rank < T; i0 += 8) {
      int q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = i0 + u < n ? __ldcg(list_unsorted + s + i0 + u) : 0x7fffffff;
#pragma unroll
      for (int u = 0; u < 8; ++u) rank += (q[u] < p);
    }
  }
  if (rank < T) {
    list_sorted[s + rank] = p;
    // VFE row tables: row_start[v] + rank is this point's row; a non-full voxel gets one virtual pad row after its points
    const int row = row0 + rank;
    row_voxel[row] = v;
    if (rank == 0 && n < T) row_voxel[row + n] = v | kRowPadFlag;
    PT* dst = row_xyz + 3 * (long long)row;
    dst[0] = px;
    dst[1] = py;
    dst[2] = pz;
  }
}

}  // namespace

// ---- launchers --------------------------------------------------------------------------------------------
cudaError_t set_trace_voxelize(unsigned long long* trace) {
  return cudaMemcpyToSymbol(g_trace, &trace, sizeof(trace));
}

cudaError_t launch_point_pass(const void* pts, int pts_dtype, long long n_total, const SweepOffsets& so,
                              const Geom& g, Workspace& w, long long chunk_cap, cudaStream_t st, int* launches) {
  // No memsets here: totals[] is rewritten by scan_down (which also zeroes the drop counters this kernel adds to), and
  // the chunk marks (atomicMin targets of scan_down, preset kChunkFirstPreset) are put back by the order pass; both are
  // initialised by lisec_create and by do_voxelize's recovery path (api.cu). A call is a chain of kernels only, so the
  // point pass's programmatic launch overlaps the previous call's last kernel.
  cudaError_t err = cudaSuccess;
  (void)chunk_cap;
  if (n_total == 0) return cudaSuccess;
  const long long groups = (n_total + 3) / 4;
  const unsigned blocks = (unsigned)((groups + 255) / 256);
  auto* tot = reinterpret_cast<unsigned long long*>(w.totals);
  if (pts_dtype == LISEC_F32)
    err = launch_pdl(point_pass_kernel<float>, blocks, 256, 0, st, static_cast<const float*>(pts), n_total, so, g,
                     w.cell_of_point, w.count, tot);
  else
    err = launch_pdl(point_pass_kernel<double>, blocks, 256, 0, st, static_cast<const double*>(pts), n_total, so, g,
                     w.cell_of_point, w.count, tot);
  ++*launches;
  return err;
}

cudaError_t launch_cell_scan(const SweepOffsets& so, const Geom& g, Workspace& w, int scan_blocks_cap,
                             int rows_per_chunk, cudaStream_t st, int* launches) {
  const long long ncells = (long long)so.n * g.cells;
  const int nblocks = (int)((ncells + kScanTile - 1) / kScanTile);
  if (nblocks > scan_blocks_cap) return cudaErrorInvalidValue;
  // block_sums = [cap] block totals | [cap / kScanGroup + 1] group totals (16-byte slots; zero between calls)
  cudaError_t err = launch_pdl(scan_reduce_kernel, nblocks, kScanThreads, 0, st, (const int*)w.count, ncells, g.T,
                               nblocks, scan_blocks_cap, w.block_sums);
  if (err == cudaSuccess)
    err = launch_pdl(scan_down_kernel, nblocks, kScanThreads, 0, st, (const int*)w.count, ncells, g.T, g.cells, nblocks,
                     so.n, rows_per_chunk, w.chunk_first, (const int*)w.block_sums, scan_blocks_cap, w.totals, w.cell_voxel,
                     w.voxel_cell, w.voxel_start, w.row_start, w.sweep_voxel_start);
  *launches += 2;
  return err;
}

cudaError_t launch_fill_and_order(const void* pts, int pts_dtype, long long n_total, const Geom& g, long long max_chunks,
                                  int scan_blocks_cap, Workspace& w, cudaStream_t st, int* launches) {
  if (n_total == 0) return cudaSuccess;
  const long long groups = (n_total + 3) / 4;
  const unsigned fill_blocks = (unsigned)((groups + 255) / 256);
  const unsigned plan_blocks = (unsigned)((max_chunks + 7) / 8);  // the tile plan rides in the same launch (8 chunks per block)
  cudaError_t err = launch_pdl(fill_pass_kernel, fill_blocks + plan_blocks, 256, 0, st, (const int*)w.cell_of_point, n_total,
                               (const int*)w.cell_voxel, (const int*)w.voxel_start, w.count, w.list_unsorted,
                               w.entry_voxel, fill_blocks, (const int*)w.chunk_first, (const int*)w.row_start,
                               (const long long*)w.totals, w.tile_first, w.tile_row0, w.chunk_ntiles, w.tile_hdr);
  // entries <= points; threads beyond the device-side totals exit
  if (err == cudaSuccess) {
    const unsigned blocks = (unsigned)((n_total + 255) / 256);
    if (pts_dtype == LISEC_F32)
      err = launch_pdl(order_pass_kernel<float>, blocks, 256, 0, st, static_cast<const float*>(pts),
                       (const int*)w.list_unsorted, (const int*)w.entry_voxel, (const int*)w.voxel_start,
                       (const int*)w.row_start, g.T, (const long long*)w.totals, w.list_sorted, w.row_voxel,
                       static_cast<float*>(w.row_xyz), w.chunk_first, max_chunks + 2,
                       w.block_sums + 4 * (size_t)scan_blocks_cap, 4 * (scan_blocks_cap / kScanGroup + 1));
    else
      err = launch_pdl(order_pass_kernel<double>, blocks, 256, 0, st, static_cast<const double*>(pts),
                       (const int*)w.list_unsorted, (const int*)w.entry_voxel, (const int*)w.voxel_start,
                       (const int*)w.row_start, g.T, (const long long*)w.totals, w.list_sorted, w.row_voxel,
                       static_cast<double*>(w.row_xyz), w.chunk_first, max_chunks + 2,
                       w.block_sums + 4 * (size_t)scan_blocks_cap, 4 * (scan_blocks_cap / kScanGroup + 1));
  }
  *launches += 2;
  return err;
}

}  // namespace lisec
