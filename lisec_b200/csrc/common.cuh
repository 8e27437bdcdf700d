// Internal declarations shared by the kernels and the C ABI of liblisec_b200.so. Not installed.
#pragma once
#include <cstdlib>

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lisec_b200.h"

namespace lisec {

// ---- geometry: the positional arguments of VFE_preprocessing (reference model_training.py:112) ----------
struct Geom {
  double size[3];  // xSize, ySize, zSize
  double inv[3];   // 1/size when that is exact (size is a power of two), else 0
  int exact_inv;   // 1: key = floor(p * inv) is bit-identical to floor(p / size)
  int maxx, maxy, maxz;  // maxVoxelX, maxVoxelY, maxVoxelZ
  int nx, ny, nz;        // 2*maxx, 2*maxy, maxz
  int T;                 // sampleSize
  int cells;             // nz*nx*ny cells per sweep
};

// sweep_offsets travels by value so the point pass needs no extra device read.
struct SweepOffsets {
  long long off[LISEC_MAX_SWEEPS + 1];
  int n;
};

// Folded VFE parameters, primary architecture (model_training.py:231-233): 6->16 | 32->32 | 64->64.
// VfeSmall travels as a __grid_constant__ kernel parameter (every CTA copies the front stage's part into shared
// memory once); the tensor-core operands travel as one device blob of four images in the K-major 128-byte-swizzled
// layout of umma.cuh, each a tf32 hi part and a tf32 lo part (3xTF32):
//   W3^T[c_out 64][c_in 64]  dense_2, A operand of the FCN GEMM, 2 slabs of 32 input channels:
//                            slab 0 = kernel rows that multiply the POOLED half (Concatenate([pooling, layer]), :164-165)
//   W2B[n 64][k 32]          dense_1, B operand of the VFE-2 GEMM, block-diagonal so that the two halves of the sum land
//                            in SEPARATE accumulator columns: n < 32 = pooled half (k < 16), n >= 32 = pointwise half
struct VfeSmall {
  float w1f[6][16];      // dense (6,16)
  double w1d[3][16];     // its first three rows in float64 (the coarse part of the coordinates, see vfe.cu VFE-1)
  float a1[16], b1[16];  // BN folded: y = x*a + b, a = gamma*rsqrt(var+eps), b = beta - mean*a
  float a2[32], b2[32];
  float a3[64], b3[64];
};
constexpr int kVfeW3ImageFloats = 64 * 64;  // one hi or lo image of W3^T
constexpr int kVfeW2ImageFloats = 64 * 32;  // one hi or lo image of W2B
constexpr int kVfeBlobFloats = 2 * kVfeW3ImageFloats + 2 * kVfeW2ImageFloats;  // [W3 hi | W3 lo | W2B hi | W2B lo]

// totals[] slots (device, long long)
enum {
  TOT_VOXELS = 0,
  TOT_ENTRIES = 1,  // in-range points
  TOT_ROWS = 2,     // kept rows + one pad row per non-full voxel
  TOT_CHUNKS = 3,   // VFE chunks (runs of whole voxels with at most kVfeChunkRows rows; tiles are packed inside a chunk)
  TOT_NONFINITE = 4,
  TOT_OUT_OF_RANGE = 5,
  TOT_ACC_NONFINITE = 6,     // the point pass's running drop counters: moved to the two slots above and zeroed again by
  TOT_ACC_OUT_OF_RANGE = 7,  // scan_down, so that no call starts with a memset (zero between calls)
  TOT_COUNT = 8
};

constexpr int kVfeThreads = 128;  // rows per VFE tile (one M = 128 accumulator block)
constexpr int kVfeChunkRows = 4 * kVfeThreads;  // rows per VFE chunk
constexpr int kChunkFirstPreset = 0x7f7f7f7f;  // chunk_first[] between calls: above any voxel index (atomicMin target)
constexpr int kChunkSlots = 12;  // tile-table entries per chunk (<= 9 tiles + the end sentinel)
constexpr int kRowPadFlag = 1 << 30;  // row_voxel[] bit: the row is its voxel's virtual pad row
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;     // cells per thread in the cell-table scans (16 measured slower: scan_down 20.5 -> 23.3 us)
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kScanGroup = 32;    // scan blocks per group of the two-level block totals

struct Workspace {
  // cell tables, [max_sweeps * cells]
  int* count = nullptr;       // points per cell; built by the point pass, drained back to 0 by the fill pass
  int* cell_voxel = nullptr;  // occupancy map: voxel row of the cell, or -1
  // per point, [max_points]
  int* cell_of_point = nullptr;
  int* list_unsorted = nullptr;  // CSR payload in arrival order
  int* list_sorted = nullptr;    // CSR payload, first min(count,T) of each segment ascending
  int* entry_voxel = nullptr;    // voxel row of each CSR entry
  // per voxel, [max_voxels + 1]
  int* voxel_cell = nullptr;
  int* voxel_start = nullptr;  // CSR offsets into list_*
  int* row_start = nullptr;    // offsets in VFE rows (kept + pad)
  int* chunk_first = nullptr;  // first voxel of each VFE chunk, [max_chunks + 2] (atomicMin marks of scan_down)
  int* chunk_row0 = nullptr;   // (unused)
  int* chunk_ntiles = nullptr; // tiles packed into each chunk
  int* tile_first = nullptr;   // [max_chunks][kChunkSlots] first voxel of each tile of a chunk, then the end sentinel
  // per VFE row, [max_points + max_voxels]: what the VFE kernel needs to start a tile with one coalesced read
  int* row_voxel = nullptr;    // voxel row the VFE row belongs to (| kRowPadFlag for the virtual pad row)
  void* row_xyz = nullptr;     // [rows][3] the point of every VFE row in the input dtype (unwritten for pad rows)
  int* tile_hdr = nullptr;     // [max_chunks][kChunkSlots][4] (first voxel, end voxel, first row, end row) of each tile
  int* tile_row0 = nullptr;    // [max_chunks][kChunkSlots] first VFE row of each tile
  int* writer_claim = nullptr; // [2] the fused kernel's background writers: next batch of cells, finished warps (vfe.cu)
  int* block_sums = nullptr;   // [cap][4] (voxels, entries, rows, -) of each scan block | [cap / kScanGroup + 1][4] of each group of blocks
  int* sweep_voxel_start = nullptr;  // [max_sweeps + 1]
  long long* totals = nullptr;       // [TOT_COUNT]
  float* voxel_feat = nullptr;       // [max_voxels, c3] for the fused entry point
  float* c_empty = nullptr;          // [c3]
  float* vfe_w = nullptr;            // [kVfeBlobFloats] weight blob (see VfeSmall)
  void* staging = nullptr;           // host->device landing buffer for *_host entry points
  int* empty_desc = nullptr;         // 8 ints describing the one-voxel problem that yields c_empty
  // Debug only (LISEC_TRACE=1 at lisec_create, else nullptr): per-CTA cycle counters of the VFE kernel's pipeline
  // stages, long long [kTraceCtas][16] (slots listed in vfe.cu); read back with lisec_debug_trace().
  unsigned long long* trace = nullptr;
};
constexpr int kTraceCtas = 256, kTraceSlots = 16;

// ---- launchers (each returns the cudaError_t of its launches) --------------------------------------------
cudaError_t launch_point_pass(const void* pts, int pts_dtype, long long n_total, const SweepOffsets& so,
                              const Geom& g, Workspace& w, long long chunk_cap, cudaStream_t st, int* launches);
cudaError_t launch_cell_scan(const SweepOffsets& so, const Geom& g, Workspace& w, int scan_blocks_cap,
                             int rows_per_chunk, cudaStream_t st, int* launches);
cudaError_t launch_fill_and_order(const void* pts, int pts_dtype, long long n_total, const Geom& g, long long max_chunks,
                                  int scan_blocks_cap, Workspace& w, cudaStream_t st, int* launches);
cudaError_t launch_export(const void* pts, int pts_dtype, const SweepOffsets& so, const Geom& g,
                          const Workspace& w, long long max_voxels, int32_t* coords, int32_t* counts,
                          int32_t* point_idx, float* features, float* dense, cudaStream_t st, int* launches);
// The VFE problem description (device pointers): tiles -> rows -> (point, voxel); see Workspace.
struct VfeProblem {
  const int* tile_first;   // [n_chunks][kChunkSlots] first voxel of each tile of a chunk (entry ntiles = the end)
  const int* tile_row0;    // [n_chunks][kChunkSlots] first VFE row
  const int* chunk_ntiles; // [n_chunks]
  const int* row_voxel;    // [rows] (| kRowPadFlag)
  const void* row_xyz;     // [rows][3] float32 or float64 (pts_dtype)
  const int* row_start;    // [voxels + 1] first VFE row of each voxel
  const long long* n_chunks;
  int pts_dtype;
  const int* tile_hdr;     // [n_chunks][kChunkSlots][4] the same tiles as (first voxel, end voxel, first row, end row): ONE
                           // 16-byte-aligned record per tile, which vfe_kernel's walkers fetch with an asynchronous copy
};
cudaError_t launch_vfe(const VfeSmall& p, const float* wblob, const VfeProblem& prob,
                       float* voxel_feat, int sm_count, cudaStream_t st, int* launches, long long* prof = nullptr);
// Fused VFE + dense grid: voxel rows go straight to their cells, a 9th warp per CTA streams c_empty into empty cells.
cudaError_t launch_vfe_to_grid(const VfeSmall& p, const float* wblob,
                               const VfeProblem& prob, const Workspace& w, const Geom& g, int n_sweeps, int grid_dtype,
                               void* grid, int first_group, int sm_count, cudaStream_t st, int* launches);
// c_empty into cells [0, ncells) regardless of occupancy (the blind prefix, see scatter.cu)
cudaError_t launch_grid_fill(int grid_dtype, const float* c_empty, void* grid, long long ncells, int sm_count,
                             cudaStream_t st);
cudaError_t launch_grid_write(const Geom& g, int n_sweeps, int c3, int grid_dtype, const int* cell_voxel,
                              const float* voxel_feat, const float* c_empty, void* grid, int sm_count,
                              cudaStream_t st, int* launches);

// The generic VFE kernel (vfe_generic.cu): any supported (c1, c2, c3) and either FCN variant, float32 FMAs. `post` = the
// FCN's second Dense (model_training.py:172, commented-out variant); a / b = BatchNormalization folded to y = z * a + b.
struct GenericVfeWeights {
  const float* dense[3];
  const float* a[3];
  const float* b[3];
  const float* post[3];
};
bool vfe_generic_supports(int c1, int c2, int c3);
size_t vfe_generic_param_floats(int c1, int c2, int c3, bool post);
void vfe_generic_pack(int c1, int c2, int c3, bool post, const GenericVfeWeights& w, float* out);
cudaError_t launch_vfe_generic(int c1, int c2, int c3, bool post, const float* params, const VfeProblem& prob,
                               float* voxel_feat, int sm_count, cudaStream_t st, int* launches);

int vfe_rows_per_chunk(int T);
cudaError_t set_trace_voxelize(unsigned long long* trace);
cudaError_t set_trace_vfe(unsigned long long* trace);

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// The path is a chain of short kernels on one stream. Each is launched with programmatic stream serialization: its
// CTAs may be scheduled as soon as every CTA of its predecessor has executed pdl_launch_dependents() (first statement
// of every kernel), and pdl_wait() then blocks until the predecessor has completed and its writes are visible. The
// launch latency and any set-up a kernel does before pdl_wait() overlap the predecessor's tail.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Debug timeline (LISEC_TRACE=1): every kernel stamps the moment its pdl_wait() returned — i.e. the moment its
// predecessor completed — into rows 200+k of the trace buffer ([0] = earliest CTA, [1] = latest stamp). Kernel ids:
enum { TL_POINT = 0, TL_SCAN_REDUCE, TL_SCAN_DOWN, TL_FILL, TL_ORDER, TL_WRITER_END, TL_VFE, TL_VFE_END, TL_COUNT };
constexpr int kTimelineRow0 = 200;
__device__ __forceinline__ void timeline_stamp(unsigned long long* trace, int k) {
  if (trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    atomicMin(trace + (size_t)(kTimelineRow0 + k) * kTraceSlots, t);
    atomicMax(trace + (size_t)(kTimelineRow0 + k) * kTraceSlots + 1, t);
  }
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  // LISEC_NO_PDL=1: plain stream order (every kernel starts after its predecessor has completed and flushed) — the
  // reference behaviour that tests/test_gpu_pdl.py compares the overlapped launches against, bit for bit
  static const int allow = [] {
    const char* e = getenv("LISEC_NO_PDL");
    return (e && e[0] == '1') ? 0 : 1;
  }();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = allow;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

}  // namespace lisec
