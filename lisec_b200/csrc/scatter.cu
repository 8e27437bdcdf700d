// Dense voxel-grid writer on sm_100a: the tensor MaxPoolingVFELayer(combine=True) hands to the first Conv3D
// (reference model_training.py:235-236), shape [N, nz, nx, ny, C3].
//
// The reference reaches it by densifying the INPUT (sparse.to_dense, model_training.py:279 / Predict.py:29-30) and
// running the VFE on all nz*nx*ny*T slots. Here the VFE ran on occupied voxels only, and this kernel writes every
// grid element exactly once: the voxel's feature row where the occupancy map says so, c_empty elsewhere (the
// unmasked network's output for an all-zero voxel is a non-zero constant). No memset + scatter:
// that would write the occupied rows twice.
//
// HBM-bound: C3 x 4 B (f32) / C3 x 2 B (bf16) stored per cell against 4 B of occupancy map read. A warp takes 32
// consecutive cells: one coalesced 128 B map load, then 16-byte streaming stores, 512 contiguous bytes per
// store instruction. C3 = 64 (the current createModel) or 128 (the graph model.png shows).
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace lisec {

namespace {

__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&h);
}

// 16 bytes of a cell's row: 4 float32 channels, or 8 channels rounded once from the float32 result to bf16
template <typename GT>
__device__ __forceinline__ float4 row_piece(const float* __restrict__ row, int piece) {
  if (sizeof(GT) == 4) return __ldg(reinterpret_cast<const float4*>(row) + piece);
  const float4 f0 = __ldg(reinterpret_cast<const float4*>(row) + 2 * piece);
  const float4 f1 = __ldg(reinterpret_cast<const float4*>(row) + 2 * piece + 1);
  return make_float4(__uint_as_float(pack_bf16x2(f0.x, f0.y)), __uint_as_float(pack_bf16x2(f0.z, f0.w)),
                     __uint_as_float(pack_bf16x2(f1.x, f1.y)), __uint_as_float(pack_bf16x2(f1.z, f1.w)));
}

// C channels of GT per cell: kLanes = C * sizeof(GT) / 16 lanes x 16 B cover one cell (8: bf16 C = 64; 16: f32 C = 64 or
// bf16 C = 128; 32: f32 C = 128), so a warp stores 32 / kLanes cells = 512 contiguous bytes per instruction.
template <int C, typename GT>
__global__ void __launch_bounds__(256) grid_write_kernel(const int* __restrict__ cell_voxel,
                                                         const float* __restrict__ voxel_feat,
                                                         const float* __restrict__ c_empty, GT* __restrict__ grid,
                                                         long long ncells) {
  constexpr int kLanes = C * (int)sizeof(GT) / 16;
  constexpr int kCellsPerStore = 32 / kLanes;
  const int lane = threadIdx.x & 31;
  const int sub = lane / kLanes;
  const int piece = lane % kLanes;
  const float4 bg = row_piece<GT>(c_empty, piece);
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long ngroups = (ncells + 31) >> 5;
  for (long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; grp < ngroups; grp += warps) {
    const long long base = grp << 5;
    const int my = (base + lane < ncells) ? __ldg(cell_voxel + base + lane) : -1;
#pragma unroll 8
    for (int i = 0; i < kLanes; ++i) {
      const int vox = __shfl_sync(0xffffffffu, my, kCellsPerStore * i + sub);
      const float4 val = vox >= 0 ? row_piece<GT>(voxel_feat + (size_t)vox * C, piece) : bg;
      const long long cell = base + kCellsPerStore * i + sub;
      if (cell < ncells) st_stream(reinterpret_cast<float4*>(grid + cell * C) + piece, val);
    }
  }
}

}  // namespace

// The background of cells [0, ncells) without looking at the occupancy map: every cell gets c_empty. The front end
// runs this for a PREFIX of the grid on a side stream while the grouping chain (which leaves HBM idle) still decides
// which cells are occupied; the fused kernel then writes the occupied cells of that prefix over it and streams the
// rest of the background itself. 16-byte streaming stores, a warp covers 512 contiguous bytes per instruction.
template <typename GT>
__global__ void __launch_bounds__(256) grid_fill_kernel(const float* __restrict__ c_empty, GT* __restrict__ grid,
                                                        long long ncells) {
  constexpr int kLanesPerCell = 64 * (int)sizeof(GT) / 16;  // 16 (f32) or 8 (bf16)
  const int piece = threadIdx.x % kLanesPerCell;
  float4 val;
  if (sizeof(GT) == 4) {
    val = reinterpret_cast<const float4*>(c_empty)[piece];
  } else {
    unsigned p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = pack_bf16x2(c_empty[8 * piece + 2 * i], c_empty[8 * piece + 2 * i + 1]);
    val = make_float4(__uint_as_float(p[0]), __uint_as_float(p[1]), __uint_as_float(p[2]), __uint_as_float(p[3]));
  }
  const long long n16 = ncells * kLanesPerCell;  // 16-byte pieces; piece index i belongs to lane i % kLanesPerCell
  float4* dst = reinterpret_cast<float4*>(grid);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    st_stream(dst + i, val);
}

cudaError_t launch_grid_fill(int grid_dtype, const float* c_empty, void* grid, long long ncells, int sm_count,
                             cudaStream_t st) {
  if (ncells <= 0) return cudaSuccess;
  static const int per_sm = [] {  // experiment switch: fill CTAs per SM (default 4 x 256 threads: half an SM's thread slots)
    const char* e = getenv("LISEC_FILL_CTAS_PER_SM");
    const int v = e ? atoi(e) : 4;
    return v >= 1 && v <= 8 ? v : 4;
  }();
  const unsigned blocks = (unsigned)sm_count * per_sm;
  if (grid_dtype == LISEC_F32)
    grid_fill_kernel<float><<<blocks, 256, 0, st>>>(c_empty, static_cast<float*>(grid), ncells);
  else
    grid_fill_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(c_empty, static_cast<__nv_bfloat16*>(grid), ncells);
  return cudaGetLastError();
}

template <int C>
static void grid_write_launch(int grid_dtype, unsigned blocks, cudaStream_t st, const int* cell_voxel,
                              const float* voxel_feat, const float* c_empty, void* grid, long long ncells) {
  if (grid_dtype == LISEC_F32)
    grid_write_kernel<C, float><<<blocks, 256, 0, st>>>(cell_voxel, voxel_feat, c_empty, static_cast<float*>(grid), ncells);
  else
    grid_write_kernel<C, __nv_bfloat16><<<blocks, 256, 0, st>>>(cell_voxel, voxel_feat, c_empty,
                                                               static_cast<__nv_bfloat16*>(grid), ncells);
}

cudaError_t launch_grid_write(const Geom& g, int n_sweeps, int c3, int grid_dtype, const int* cell_voxel,
                              const float* voxel_feat, const float* c_empty, void* grid, int sm_count,
                              cudaStream_t st, int* launches) {
  if (c3 != 64 && c3 != 128) return cudaErrorInvalidValue;
  const long long ncells = (long long)n_sweeps * g.cells;
  if (ncells == 0) return cudaSuccess;
  long long blocks = ((ncells + 31) / 32 + 7) / 8;  // 8 warps per block, one 32-cell group per warp
  const long long cap = (long long)sm_count * 8 * 4; // grid-stride beyond 4 waves of 8 resident CTAs per SM
  if (blocks > cap) blocks = cap;
  if (c3 == 64)
    grid_write_launch<64>(grid_dtype, (unsigned)blocks, st, cell_voxel, voxel_feat, c_empty, grid, ncells);
  else
    grid_write_launch<128>(grid_dtype, (unsigned)blocks, st, cell_voxel, voxel_feat, c_empty, grid, ncells);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace lisec
