// Export of the grouping in the reference's terms: coords (z,x,y), counts, ordered first-T point lists, the float32
// feature rows of model_training.py:134-141, and (tests / tiny grids) the dense [N,nz,nx,ny,T,6] model input of
// model_training.py:143-152 + sparse.to_dense. Not on the product path.
#include "common.cuh"
#include "vfe_math.cuh"

namespace lisec {

namespace {

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
