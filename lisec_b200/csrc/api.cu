// C ABI of liblisec_b200.so (see include/lisec_b200.h). Host-side only: argument checks, workspace ownership,
// BN folding, launch sequencing. No torch types, no exceptions across the boundary, no CPU compute path.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "handle.cuh"
#include "umma.cuh"

namespace {

bool is_pow2_double(double s) {
  int e;
  return s > 0 && std::frexp(s, &e) == 0.5;
}

void free_workspace(Workspace& w) {
  void* ptrs[] = {w.count, w.cell_voxel, w.cell_of_point, w.list_unsorted, w.list_sorted, w.entry_voxel,
                  w.voxel_cell, w.voxel_start, w.row_start, w.tile_first, w.tile_row0, w.chunk_first, w.chunk_row0,
                  w.chunk_ntiles, w.row_voxel, w.row_xyz,
                  w.block_sums, w.sweep_voxel_start,
                  w.totals, w.voxel_feat, w.c_empty, w.vfe_w, w.staging, w.empty_desc, w.trace, w.writer_claim, w.tile_hdr};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  w = Workspace();
}

int check_offsets(lisec_handle* h, const int64_t* off, int n_sweeps, SweepOffsets* so) {
  if (!off) return fail(h, LISEC_ERR_BAD_ARG, "sweep_offsets is NULL");
  if (n_sweeps < 1) return fail(h, LISEC_ERR_BAD_ARG, "n_sweeps = %d, need >= 1", n_sweeps);
  if (n_sweeps > h->cfg.max_sweeps)
    return fail(h, LISEC_ERR_CAPACITY, "n_sweeps = %d exceeds the handle's max_sweeps = %d", n_sweeps,
                h->cfg.max_sweeps);
  if (off[0] != 0) return fail(h, LISEC_ERR_BAD_ARG, "sweep_offsets[0] = %lld, need 0", (long long)off[0]);
  for (int s = 0; s < n_sweeps; ++s)
    if (off[s + 1] < off[s]) return fail(h, LISEC_ERR_BAD_ARG, "sweep_offsets decreases at %d", s + 1);
  if (off[n_sweeps] > h->cfg.max_points)
    return fail(h, LISEC_ERR_CAPACITY, "%lld points exceed the handle's max_points = %lld",
                (long long)off[n_sweeps], (long long)h->cfg.max_points);
  so->n = n_sweeps;
  for (int s = 0; s <= n_sweeps; ++s) so->off[s] = off[s];
  for (int s = n_sweeps + 1; s <= LISEC_MAX_SWEEPS; ++s) so->off[s] = off[n_sweeps];
  return LISEC_OK;
}

int do_voxelize(lisec_handle* h, const void* points, int dtype, const SweepOffsets& so, cudaStream_t st) {
  const long long n_total = so.off[so.n];
  if (h->count_dirty) {  // first call, or a call that failed half way: the state the kernels otherwise leave behind
    LISEC_CUDA(h, cudaMemsetAsync(h->ws.count, 0, sizeof(int) * (size_t)h->ncells_cap, st));
    LISEC_CUDA(h, cudaMemsetAsync(h->ws.totals, 0, sizeof(long long) * TOT_COUNT, st));
    LISEC_CUDA(h, cudaMemsetAsync(h->ws.chunk_first, 0x7f, sizeof(int) * ((size_t)h->max_chunks + 2), st));
    LISEC_CUDA(h, cudaMemsetAsync(h->ws.block_sums + (size_t)4 * h->scan_blocks_cap, 0,
                                  sizeof(int) * 4 * ((size_t)h->scan_blocks_cap / kScanGroup + 1), st));  // the scans' group totals
    h->count_dirty = false;
  }
  h->voxelized = false;
  h->count_dirty = true;  // until the fill pass has been enqueued
  LISEC_CUDA(h, launch_point_pass(points, dtype, n_total, so, h->geom, h->ws, h->max_chunks, st, &h->launches));
  LISEC_CUDA(h, launch_cell_scan(so, h->geom, h->ws, h->scan_blocks_cap, h->rows_per_chunk, st, &h->launches));
  LISEC_CUDA(h, launch_fill_and_order(points, dtype, n_total, h->geom, h->max_chunks, h->scan_blocks_cap, h->ws, st, &h->launches));
  h->count_dirty = false;
  h->last_points = points;
  h->last_dtype = dtype;
  h->last_so = so;
  h->voxelized = true;
  return LISEC_OK;
}

VfeProblem vfe_problem(const lisec_handle* h) {
  return VfeProblem{h->ws.tile_first, h->ws.tile_row0, h->ws.chunk_ntiles, h->ws.row_voxel, h->ws.row_xyz,
                    h->ws.row_start,  h->ws.totals + TOT_CHUNKS, h->last_dtype, h->ws.tile_hdr};
}

VfeProblem empty_problem(const lisec_handle* h) {
  const int* d = h->ws.empty_desc;
  return VfeProblem{d, d + 2, d + 1, d + 8, d + 16, d + 4, reinterpret_cast<const long long*>(d + 12), LISEC_F64, d + 24};
}

// vfe_generic.cu's parameter block: the Keras kernels as they are, BatchNormalization folded to y = z * a + b in float32
// (Keras inference, model_training.py:171), the FCNs' second Dense where the graph has one. c_empty = the same kernel on
// the one-voxel problem that holds nothing but a pad row.
int set_generic_weights(lisec_handle* h, const lisec_vfe_weights* w, cudaStream_t st) {
  const lisec_config& c = h->cfg;
  const bool post = c.fcn_post_dense != 0;
  const int C[3] = {c.c1, c.c2, c.c3};
  float a[3][128], b[3][128];
  GenericVfeWeights g;
  for (int l = 0; l < 3; ++l) {
    if (post && !w->post_dense_kernel[l])
      return fail(h, LISEC_ERR_BAD_ARG, "fcn_post_dense = 1 but post_dense_kernel[%d] is NULL", l);
    for (int j = 0; j < C[l]; ++j) {
      a[l][j] = (1.0f / std::sqrt(w->bn_var[l][j] + w->bn_epsilon)) * w->bn_gamma[l][j];
      b[l][j] = w->bn_beta[l][j] - w->bn_mean[l][j] * a[l][j];
    }
    g.dense[l] = w->dense_kernel[l];
    g.a[l] = a[l];
    g.b[l] = b[l];
    g.post[l] = post ? w->post_dense_kernel[l] : nullptr;
  }
  const size_t n = vfe_generic_param_floats(c.c1, c.c2, c.c3, post);
  float* host = new (std::nothrow) float[n];
  if (!host) return fail(h, LISEC_ERR_CUDA, "out of host memory");
  vfe_generic_pack(c.c1, c.c2, c.c3, post, g, host);
  cudaError_t e = cudaMemcpyAsync(h->generic_params, host, sizeof(float) * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  delete[] host;
  LISEC_CUDA(h, e);
  h->weights_set = true;
  LISEC_CUDA(h, launch_vfe_generic(c.c1, c.c2, c.c3, post, h->generic_params, empty_problem(h), h->ws.c_empty,
                                   h->sm_count, st, &h->launches));
  LISEC_CUDA(h, cudaStreamSynchronize(st));
  return LISEC_OK;
}

int do_vfe(lisec_handle* h, float* voxel_feat, cudaStream_t st) {
  const VfeProblem prob = vfe_problem(h);
  if (h->generic) {
    LISEC_CUDA(h, launch_vfe_generic(h->cfg.c1, h->cfg.c2, h->cfg.c3, h->cfg.fcn_post_dense != 0, h->generic_params, prob,
                                     voxel_feat, h->sm_count, st, &h->launches));
    return LISEC_OK;
  }
  LISEC_CUDA(h, launch_vfe(h->params, h->ws.vfe_w, prob, voxel_feat, h->sm_count, st, &h->launches,
                           reinterpret_cast<long long*>(h->ws.trace)));
  return LISEC_OK;
}

int check_points(lisec_handle* h, const void* points, int dtype, long long n_total, bool device) {
  if (dtype != LISEC_F32 && dtype != LISEC_F64)
    return fail(h, LISEC_ERR_BAD_ARG, "points_dtype = %d, need LISEC_F32 or LISEC_F64", dtype);
  if (n_total > 0 && !points) return fail(h, LISEC_ERR_BAD_ARG, "points is NULL");
  if (device && (reinterpret_cast<uintptr_t>(points) & 15))
    return fail(h, LISEC_ERR_BAD_ARG, "points must be 16-byte aligned");
  return LISEC_OK;
}

}  // namespace

extern "C" {

int32_t lisec_abi_version(void) { return LISEC_ABI_VERSION; }

int32_t lisec_create(const lisec_config* cfg, lisec_handle** out) {
  if (!cfg || !out) return LISEC_ERR_BAD_ARG;
  *out = nullptr;
  lisec_handle* h = new (std::nothrow) lisec_handle();
  if (!h) return LISEC_ERR_CUDA;
  h->err[0] = 0;
  h->cfg = *cfg;
  *out = h;  // returned even on failure so the caller can read lisec_last_error(); lisec_destroy() frees it

  const lisec_config& c = h->cfg;
  if (!(c.voxel_x > 0 && c.voxel_y > 0 && c.voxel_z > 0) || !std::isfinite(c.voxel_x) ||
      !std::isfinite(c.voxel_y) || !std::isfinite(c.voxel_z))
    return fail(h, LISEC_ERR_BAD_CONFIG, "voxel sizes must be finite and > 0");
  if (c.max_voxel_x < 1 || c.max_voxel_y < 1 || c.max_voxel_z < 1)
    return fail(h, LISEC_ERR_BAD_CONFIG, "max_voxel_{x,y,z} must be >= 1");
  if (c.sample_size < 2 || c.sample_size > 64)
    return fail(h, LISEC_ERR_BAD_CONFIG, "sample_size = %d, supported range is 2..64", c.sample_size);
  if (!vfe_generic_supports(c.c1, c.c2, c.c3))
    return fail(h, LISEC_ERR_UNSUPPORTED,
                "VFE widths (%d,%d,%d): built are (16,32,64) = createModel as it stands and (16,64,128) = model.png's",
                c.c1, c.c2, c.c3);
  if (c.fcn_post_dense != 0 && c.fcn_post_dense != 1)
    return fail(h, LISEC_ERR_BAD_CONFIG, "fcn_post_dense = %d, need 0 or 1", c.fcn_post_dense);
  h->generic = !(c.c1 == 16 && c.c2 == 32 && c.c3 == 64 && c.fcn_post_dense == 0);
  if (const char* e = getenv("LISEC_GENERIC_VFE"); e && e[0] == '1') h->generic = true;  // cross-check switch
  if (c.grid_dtype != LISEC_F32 && c.grid_dtype != LISEC_BF16)
    return fail(h, LISEC_ERR_BAD_CONFIG, "grid_dtype must be LISEC_F32 or LISEC_BF16");
  if (c.max_sweeps < 1 || c.max_sweeps > LISEC_MAX_SWEEPS)
    return fail(h, LISEC_ERR_BAD_CONFIG, "max_sweeps = %d, supported range is 1..%d", c.max_sweeps,
                LISEC_MAX_SWEEPS);
  if (c.max_points < 1 || c.max_points > 2000000000LL)
    return fail(h, LISEC_ERR_BAD_CONFIG, "max_points out of range");

  Geom& g = h->geom;
  g.size[0] = c.voxel_x; g.size[1] = c.voxel_y; g.size[2] = c.voxel_z;
  g.exact_inv = is_pow2_double(c.voxel_x) && is_pow2_double(c.voxel_y) && is_pow2_double(c.voxel_z);
  for (int i = 0; i < 3; ++i) g.inv[i] = g.exact_inv ? 1.0 / g.size[i] : 0.0;
  g.maxx = c.max_voxel_x; g.maxy = c.max_voxel_y; g.maxz = c.max_voxel_z;
  g.nx = 2 * c.max_voxel_x; g.ny = 2 * c.max_voxel_y; g.nz = c.max_voxel_z;
  g.T = c.sample_size;
  const long long cells = (long long)g.nz * g.nx * g.ny;
  if (cells * c.max_sweeps > 2000000000LL)
    return fail(h, LISEC_ERR_BAD_CONFIG, "grid of %lld cells x %d sweeps exceeds int32 indexing", cells,
                c.max_sweeps);
  g.cells = (int)cells;

  int ndev = 0;
  LISEC_CUDA(h, cudaGetDeviceCount(&ndev));
  if (c.device < 0 || c.device >= ndev)
    return fail(h, LISEC_ERR_CUDA, "device %d not available (%d CUDA devices)", c.device, ndev);
  LISEC_CUDA(h, cudaSetDevice(c.device));
  int major = 0;
  LISEC_CUDA(h, cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, c.device));
  if (major != 10)
    return fail(h, LISEC_ERR_CUDA, "device %d has compute capability %d.x; this library is built for sm_100a only",
                c.device, major);
  LISEC_CUDA(h, cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, c.device));
  if (const char* e = getenv("LISEC_VFE_CTAS")) {  // experiment knob: persistent CTAs of the VFE kernel (default: one per SM)
    const int v = atoi(e);
    if (v > 0 && v < h->sm_count) h->sm_count = v;
  }

  h->rows_per_chunk = vfe_rows_per_chunk(g.T);
  h->ncells_cap = cells * c.max_sweeps;
  h->max_voxels = c.max_points < h->ncells_cap ? c.max_points : h->ncells_cap;
  h->max_chunks = (c.max_points + h->max_voxels) / h->rows_per_chunk + 2;
  h->scan_blocks_cap = (int)((h->ncells_cap + kScanTile - 1) / kScanTile);

  Workspace& w = h->ws;
  const size_t P = (size_t)c.max_points + 4, V = (size_t)h->max_voxels + 1;
  LISEC_CUDA(h, dev_alloc(h, &w.count, (size_t)h->ncells_cap + kScanItems));
  LISEC_CUDA(h, dev_alloc(h, &w.cell_voxel, (size_t)h->ncells_cap + kScanItems));
  LISEC_CUDA(h, dev_alloc(h, &w.cell_of_point, P));
  LISEC_CUDA(h, dev_alloc(h, &w.list_unsorted, P));
  LISEC_CUDA(h, dev_alloc(h, &w.list_sorted, P));
  LISEC_CUDA(h, dev_alloc(h, &w.entry_voxel, P));
  LISEC_CUDA(h, dev_alloc(h, &w.voxel_cell, V));
  LISEC_CUDA(h, dev_alloc(h, &w.voxel_start, V + 1));
  LISEC_CUDA(h, dev_alloc(h, &w.row_start, V + 1));
  LISEC_CUDA(h, dev_alloc(h, &w.chunk_first, (size_t)h->max_chunks + 2));
  LISEC_CUDA(h, dev_alloc(h, &w.chunk_row0, (size_t)h->max_chunks + 2));
  LISEC_CUDA(h, dev_alloc(h, &w.chunk_ntiles, (size_t)h->max_chunks + 2));
  LISEC_CUDA(h, dev_alloc(h, &w.tile_first, ((size_t)h->max_chunks + 2) * kChunkSlots));
  LISEC_CUDA(h, dev_alloc(h, &w.tile_row0, ((size_t)h->max_chunks + 2) * kChunkSlots));
  LISEC_CUDA(h, dev_alloc(h, &w.tile_hdr, ((size_t)h->max_chunks + 2) * kChunkSlots * 4));
  LISEC_CUDA(h, dev_alloc(h, &w.row_voxel, P + V));
  LISEC_CUDA(h, dev_alloc(h, reinterpret_cast<unsigned char**>(&w.row_xyz), 3 * sizeof(double) * (P + V)));
  // (v, e, r, -) per scan block and per group of kScanGroup blocks (the groups: accumulated by atomics, zero between calls)
  LISEC_CUDA(h, dev_alloc(h, &w.block_sums, (size_t)8 * h->scan_blocks_cap + 4));
  LISEC_CUDA(h, cudaMemset(w.block_sums, 0, sizeof(int) * ((size_t)8 * h->scan_blocks_cap + 4)));
  LISEC_CUDA(h, dev_alloc(h, &w.sweep_voxel_start, (size_t)c.max_sweeps + 2));
  LISEC_CUDA(h, dev_alloc(h, &w.totals, (size_t)TOT_COUNT));
  LISEC_CUDA(h, dev_alloc(h, &w.writer_claim, (size_t)4));
  LISEC_CUDA(h, cudaMemset(w.writer_claim, 0, sizeof(int) * 4));
  LISEC_CUDA(h, dev_alloc(h, &w.voxel_feat, V * (size_t)c.c3));
  LISEC_CUDA(h, dev_alloc(h, &w.c_empty, (size_t)c.c3));
  LISEC_CUDA(h, dev_alloc(h, &w.vfe_w, (size_t)kVfeBlobFloats));
  if (h->generic)
    LISEC_CUDA(h, dev_alloc(h, &h->generic_params, vfe_generic_param_floats(c.c1, c.c2, c.c3, c.fcn_post_dense != 0)));
  LISEC_CUDA(h, dev_alloc(h, reinterpret_cast<unsigned char**>(&w.staging), P * 3 * sizeof(double)));
  LISEC_CUDA(h, dev_alloc(h, reinterpret_cast<unsigned char**>(&h->staging2), P * 3 * sizeof(double)));
  LISEC_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    LISEC_CUDA(h, cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming));
    LISEC_CUDA(h, cudaEventCreateWithFlags(&h->ev_free[b], cudaEventDisableTiming));
  }
  if (const char* t = std::getenv("LISEC_TRACE"); t && t[0] == '1') {
    LISEC_CUDA(h, dev_alloc(h, &w.trace, (size_t)kTraceCtas * kTraceSlots));
    LISEC_CUDA(h, cudaMemset(w.trace, 0, sizeof(unsigned long long) * kTraceCtas * kTraceSlots));
    LISEC_CUDA(h, cudaMemset(w.trace + (size_t)kTimelineRow0 * kTraceSlots, 0xff, sizeof(unsigned long long)));
    for (int k = 0; k < TL_COUNT; ++k)  // [0] = min stamp starts at ~0, [1] = max stamp starts at 0
      LISEC_CUDA(h, cudaMemset(w.trace + (size_t)(kTimelineRow0 + k) * kTraceSlots, 0xff, sizeof(unsigned long long)));
    LISEC_CUDA(h, set_trace_voxelize(w.trace));
    LISEC_CUDA(h, set_trace_vfe(w.trace));
  }
  for (int b = 0; b < 2; ++b) LISEC_CUDA(h, cudaEventCreate(&h->ev_kernel[b]));
  LISEC_CUDA(h, cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) LISEC_CUDA(h, cudaEventCreateWithFlags(&h->ev_side[b], cudaEventDisableTiming));
  if (const char* e = getenv("LISEC_BLIND_FRACTION")) {
    const double v = atof(e);
    if (v >= 0.0 && v <= 0.9) h->blind_fraction = v;
  }
  LISEC_CUDA(h, dev_alloc(h, &w.empty_desc, (size_t)32));
  // the one-voxel problem whose VFE output is c_empty: 0 kept points, 1 pad row, 1 chunk of 1 tile. Layout (ints; the
  // arrays the kernel copies with 16-byte cp.async start on 16-byte boundaries):
  //   [0..1] tile_first {0,1} ([1] doubles as chunk_ntiles {1}) | [2..3] tile_row0 {0,1} | [4..5] row_start {0,1} |
  //   [8] row_voxel {0 | pad flag} | [12..13] n_chunks (int64) 1 | [16..21] row_xyz: 3 doubles (unused: a pad row) |
  //   [24..27] tile_hdr {0, 1, 0, 1}
  int desc[32] = {0};
  desc[1] = 1;
  desc[3] = 1;
  desc[5] = 1;
  desc[25] = 1;
  desc[27] = 1;
  desc[8] = kRowPadFlag;
  const long long one = 1;
  std::memcpy(&desc[12], &one, sizeof(one));
  LISEC_CUDA(h, cudaMemcpy(w.empty_desc, desc, sizeof(desc), cudaMemcpyHostToDevice));
  LISEC_CUDA(h, cudaMemset(w.totals, 0, sizeof(long long) * TOT_COUNT));
  return LISEC_OK;
}

void lisec_destroy(lisec_handle* h) {
  if (!h) return;
  if (h->sm_count > 0) cudaSetDevice(h->cfg.device);
  free_workspace(h->ws);
  if (h->train) free_vfe_train_state(h->train);
  if (h->generic_params) cudaFree(h->generic_params);
  if (h->staging2) cudaFree(h->staging2);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  for (int b = 0; b < 2; ++b)
    if (h->ev_side[b]) cudaEventDestroy(h->ev_side[b]);
  for (int b = 0; b < 2; ++b)
    if (h->ev_kernel[b]) cudaEventDestroy(h->ev_kernel[b]);
  for (int b = 0; b < 2; ++b) {
    if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
    if (h->ev_free[b]) cudaEventDestroy(h->ev_free[b]);
  }
  delete h;
}

const char* lisec_last_error(const lisec_handle* h) { return h ? h->err : "null handle"; }

int64_t lisec_workspace_bytes(const lisec_handle* h) { return h ? h->workspace_bytes : 0; }

int32_t lisec_last_launch_count(const lisec_handle* h) { return h ? h->launches : 0; }

int32_t lisec_set_vfe_weights(lisec_handle* h, const lisec_vfe_weights* w, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!w) return fail(h, LISEC_ERR_BAD_ARG, "weights is NULL");
  for (int l = 0; l < 3; ++l)
    if (!w->dense_kernel[l] || !w->bn_gamma[l] || !w->bn_beta[l] || !w->bn_mean[l] || !w->bn_var[l])
      return fail(h, LISEC_ERR_BAD_ARG, "weights for layer %d contain a NULL pointer", l);
  if (!(w->bn_epsilon >= 0.f)) return fail(h, LISEC_ERR_BAD_ARG, "bn_epsilon must be >= 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  if (h->generic) return set_generic_weights(h, w, st);
  VfeSmall& p = h->params;
  const float* k0 = w->dense_kernel[0];
  for (int k = 0; k < 6; ++k)
    for (int j = 0; j < 16; ++j) {
      p.w1f[k][j] = k0[k * 16 + j];
      if (k < 3) p.w1d[k][j] = (double)k0[k * 16 + j];
    }
  // BatchNormalization at inference (Keras defaults, model_training.py:171): y = x*a + b
  float* A[3] = {p.a1, p.a2, p.a3};
  float* B[3] = {p.b1, p.b2, p.b3};
  const int C[3] = {16, 32, 64};
  for (int l = 0; l < 3; ++l)
    for (int j = 0; j < C[l]; ++j) {
      const float a = (1.0f / std::sqrt(w->bn_var[l][j] + w->bn_epsilon)) * w->bn_gamma[l][j];
      A[l][j] = a;
      B[l][j] = w->bn_beta[l][j] - w->bn_mean[l][j] * a;
    }
  // The back stage takes the per-voxel max of the FCN's raw sums and applies BN + ReLU once: relu(a*z + b) is monotonic
  // in z, increasing for a >= 0 and decreasing for a < 0. Folding sign(a) into dense_2's output column makes it
  // increasing for every channel: relu(|a| * max(s*z) + b). The kernel then tracks one running max per channel.
  float sgn3[64];
  for (int m = 0; m < 64; ++m) {
    sgn3[m] = p.a3[m] < 0.f ? -1.f : 1.f;
    p.a3[m] = std::fabs(p.a3[m]);
  }
  // blob = [W3^T hi | W3^T lo | W2B hi | W2B lo]: tensor-core operand images (K-major, 128-byte swizzle, umma.cuh). Both
  // dense layers run as 3xTF32: each weight is split into hi = rn_tf32(w) and lo = rn_tf32(w - hi). A Keras kernel is
  // (C_in, C_out) row-major with the pooled half's rows first (Concatenate([pooling, layer]), :164-165).
  float* blob = h->wblob;
  std::memset(blob, 0, sizeof(float) * kVfeBlobFloats);
  {
    auto tf32_rn = [](float x) {  // cvt.rna.tf32.f32: round to nearest, ties away, 10 explicit mantissa bits
      uint32_t u;
      std::memcpy(&u, &x, 4);
      if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
      float r;
      std::memcpy(&r, &u, 4);
      return r;
    };
    unsigned char* hi = reinterpret_cast<unsigned char*>(blob);
    unsigned char* lo = hi + kVfeW3ImageFloats * sizeof(float);
    const float* k2 = w->dense_kernel[2];
    for (int k = 0; k < 64; ++k)
      for (int m = 0; m < 64; ++m) {  // A operand of the FCN: W3^T[c_out][c_in]
        const float v = k2[k * 64 + m] * sgn3[m];  // exact sign flip
        const float vh = tf32_rn(v), vl = tf32_rn(v - vh);
        const uint32_t off = umma::kmajor_offset(m, k, 64 * 128);
        std::memcpy(hi + off, &vh, 4);
        std::memcpy(lo + off, &vl, 4);
      }
    // B operand of the VFE-2 GEMM: W2B[n][k], n < 32: column n of the pooled half (k < 16), n >= 32: column n - 32 of
    // the pointwise half (k >= 16); the other entries are zero, so the two halves' sums stay in separate accumulators
    hi = reinterpret_cast<unsigned char*>(blob + 2 * kVfeW3ImageFloats);
    lo = hi + kVfeW2ImageFloats * sizeof(float);
    const float* k1 = w->dense_kernel[1];
    for (int k = 0; k < 32; ++k)
      for (int c = 0; c < 32; ++c) {
        const float v = k1[k * 32 + c];
        const float vh = tf32_rn(v), vl = tf32_rn(v - vh);
        const uint32_t off = umma::kmajor_offset(k < 16 ? c : 32 + c, k, 0);
        std::memcpy(hi + off, &vh, 4);
        std::memcpy(lo + off, &vl, 4);
      }
  }
  h->weights_set = true;
  // c_empty: the same kernel, run on one voxel that holds nothing but the pad row
  LISEC_CUDA(h, cudaMemcpyAsync(h->ws.vfe_w, blob, sizeof(float) * kVfeBlobFloats, cudaMemcpyHostToDevice, st));
  LISEC_CUDA(h, launch_vfe(p, h->ws.vfe_w, empty_problem(h), h->ws.c_empty, h->sm_count, st, &h->launches));
  LISEC_CUDA(h, cudaStreamSynchronize(st));
  return LISEC_OK;
}

int32_t lisec_get_c_empty(lisec_handle* h, float* out) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!out) return fail(h, LISEC_ERR_BAD_ARG, "c_empty_host is NULL");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  LISEC_CUDA(h, cudaMemcpy(out, h->ws.c_empty, sizeof(float) * h->cfg.c3, cudaMemcpyDeviceToHost));
  return LISEC_OK;
}

int32_t lisec_voxelize(lisec_handle* h, const void* points, int32_t dtype, const int64_t* sweep_offsets,
                       int32_t n_sweeps, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  SweepOffsets so;
  int rc = check_offsets(h, sweep_offsets, n_sweeps, &so);
  if (rc) return rc;
  rc = check_points(h, points, dtype, so.off[so.n], true);
  if (rc) return rc;
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  return do_voxelize(h, points, dtype, so, static_cast<cudaStream_t>(stream));
}

int32_t lisec_voxel_counts(lisec_handle* h, int32_t* per_sweep, int64_t* n_voxels, int64_t* n_in_range,
                           int64_t* n_oor, int64_t* n_nonfinite, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  long long tot[TOT_COUNT];
  int svs[LISEC_MAX_SWEEPS + 1];
  LISEC_CUDA(h, cudaMemcpyAsync(tot, h->ws.totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
  LISEC_CUDA(h, cudaMemcpyAsync(svs, h->ws.sweep_voxel_start, sizeof(int) * (h->last_so.n + 1),
                                cudaMemcpyDeviceToHost, st));
  LISEC_CUDA(h, cudaStreamSynchronize(st));
  if (per_sweep)
    for (int s = 0; s < h->last_so.n; ++s) per_sweep[s] = svs[s + 1] - svs[s];
  if (n_voxels) *n_voxels = tot[TOT_VOXELS];
  if (n_in_range) *n_in_range = tot[TOT_ENTRIES];
  if (n_oor) *n_oor = tot[TOT_OUT_OF_RANGE];
  if (n_nonfinite) *n_nonfinite = tot[TOT_NONFINITE];
  return LISEC_OK;
}

int32_t lisec_voxels_export(lisec_handle* h, int32_t* coords, int32_t* counts, int32_t* point_idx,
                            float* features, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  LISEC_CUDA(h, launch_export(h->last_points, h->last_dtype, h->last_so, h->geom, h->ws, h->max_voxels, coords,
                              counts, point_idx, features, nullptr, static_cast<cudaStream_t>(stream),
                              &h->launches));
  return LISEC_OK;
}

int32_t lisec_emit_dense_input(lisec_handle* h, float* dense, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!dense) return fail(h, LISEC_ERR_BAD_ARG, "dense is NULL");
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  const size_t bytes = sizeof(float) * (size_t)h->last_so.n * h->geom.cells * h->geom.T * 6;
  LISEC_CUDA(h, cudaMemsetAsync(dense, 0, bytes, st));
  LISEC_CUDA(h, launch_export(h->last_points, h->last_dtype, h->last_so, h->geom, h->ws, h->max_voxels, nullptr,
                              nullptr, nullptr, nullptr, dense, st, &h->launches));
  return LISEC_OK;
}

int32_t lisec_vfe_forward(lisec_handle* h, float* voxel_feat, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!voxel_feat) return fail(h, LISEC_ERR_BAD_ARG, "voxel_feat is NULL");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  return do_vfe(h, voxel_feat, static_cast<cudaStream_t>(stream));
}

int32_t lisec_scatter_dense(lisec_handle* h, const float* voxel_feat, void* grid, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!voxel_feat || !grid) return fail(h, LISEC_ERR_BAD_ARG, "voxel_feat / grid is NULL");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  if (reinterpret_cast<uintptr_t>(grid) & 15) return fail(h, LISEC_ERR_BAD_ARG, "grid must be 16-byte aligned");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  LISEC_CUDA(h, launch_grid_write(h->geom, h->last_so.n, h->cfg.c3, h->cfg.grid_dtype, h->ws.cell_voxel,
                                  voxel_feat, h->ws.c_empty, grid, h->sm_count,
                                  static_cast<cudaStream_t>(stream), &h->launches));
  return LISEC_OK;
}

static int fused_stage(lisec_handle* h, void* grid, int first_group, cudaStream_t st) {
  const VfeProblem prob = vfe_problem(h);
  if (h->generic) {  // two kernels: the float32 VFE into the handle's voxel rows, then every grid cell written once
    LISEC_CUDA(h, cudaEventRecord(h->ev_kernel[0], st));
    int rc = do_vfe(h, h->ws.voxel_feat, st);
    if (rc) return rc;
    LISEC_CUDA(h, launch_grid_write(h->geom, h->last_so.n, h->cfg.c3, h->cfg.grid_dtype, h->ws.cell_voxel,
                                    h->ws.voxel_feat, h->ws.c_empty, grid, h->sm_count, st, &h->launches));
    LISEC_CUDA(h, cudaEventRecord(h->ev_kernel[1], st));
    h->kernel_timed = true;
    return LISEC_OK;
  }
  // one kernel: VFE (FP32 pipe + tensor core), voxel rows and the c_empty background written to the grid concurrently
  LISEC_CUDA(h, cudaEventRecord(h->ev_kernel[0], st));
  LISEC_CUDA(h, launch_vfe_to_grid(h->params, h->ws.vfe_w, prob, h->ws, h->geom, h->last_so.n, h->cfg.grid_dtype, grid,
                                   first_group, h->sm_count, st, &h->launches));
  LISEC_CUDA(h, cudaEventRecord(h->ev_kernel[1], st));
  h->kernel_timed = true;
  return LISEC_OK;
}

static int frontend(lisec_handle* h, const void* dev_points, int dtype, const SweepOffsets& so, void* grid,
                    cudaStream_t st) {
  // The grouping chain (~75 us of latency-bound kernels) leaves HBM idle, and the background does not depend on it for
  // cells that are simply overwritten later: a side stream fills a prefix of the grid with c_empty meanwhile; the fused
  // kernel starts after both, writes the prefix's occupied cells over it and streams the rest of the background itself.
  const long long ngroups = ((long long)so.n * h->geom.cells + 31) >> 5;
  const int first_group = h->generic ? 0 : (int)(h->blind_fraction * (double)ngroups);
  if (first_group > 0) {
    LISEC_CUDA(h, cudaEventRecord(h->ev_side[0], st));  // everything queued before this call (the grid's last readers)
    LISEC_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev_side[0], 0));
    LISEC_CUDA(h, launch_grid_fill(h->cfg.grid_dtype, h->ws.c_empty, grid, (long long)first_group << 5, h->sm_count,
                                   h->side_stream));
    LISEC_CUDA(h, cudaEventRecord(h->ev_side[1], h->side_stream));
    ++h->launches;
  }
  int rc = do_voxelize(h, dev_points, dtype, so, st);
  if (rc) return rc;
  if (first_group > 0) LISEC_CUDA(h, cudaStreamWaitEvent(st, h->ev_side[1], 0));
  return fused_stage(h, grid, first_group, st);
}

int32_t lisec_vfe_scatter_fused(lisec_handle* h, void* grid, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!grid) return fail(h, LISEC_ERR_BAD_ARG, "grid is NULL");
  if (reinterpret_cast<uintptr_t>(grid) & 15) return fail(h, LISEC_ERR_BAD_ARG, "grid must be 16-byte aligned");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  return fused_stage(h, grid, 0, static_cast<cudaStream_t>(stream));
}

int32_t lisec_frontend_forward(lisec_handle* h, const void* points, int32_t dtype, const int64_t* sweep_offsets,
                               int32_t n_sweeps, void* grid, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!grid) return fail(h, LISEC_ERR_BAD_ARG, "grid is NULL");
  if (reinterpret_cast<uintptr_t>(grid) & 15) return fail(h, LISEC_ERR_BAD_ARG, "grid must be 16-byte aligned");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  SweepOffsets so;
  int rc = check_offsets(h, sweep_offsets, n_sweeps, &so);
  if (rc) return rc;
  rc = check_points(h, points, dtype, so.off[so.n], true);
  if (rc) return rc;
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  return frontend(h, points, dtype, so, grid, static_cast<cudaStream_t>(stream));
}

int32_t lisec_frontend_forward_host(lisec_handle* h, const void* points_host, int32_t dtype,
                                    const int64_t* sweep_offsets, int32_t n_sweeps, void* grid, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!grid) return fail(h, LISEC_ERR_BAD_ARG, "grid is NULL");
  if (reinterpret_cast<uintptr_t>(grid) & 15) return fail(h, LISEC_ERR_BAD_ARG, "grid must be 16-byte aligned");
  if (!h->weights_set) return fail(h, LISEC_ERR_STATE, "lisec_set_vfe_weights() has not been called");
  SweepOffsets so;
  int rc = check_offsets(h, sweep_offsets, n_sweeps, &so);
  if (rc) return rc;
  rc = check_points(h, points_host, dtype, so.off[so.n], false);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  h->launches = 0;
  const size_t bytes = (size_t)so.off[so.n] * 3 * (dtype == LISEC_F64 ? sizeof(double) : sizeof(float));
  const int b = h->staging_idx;
  h->staging_idx ^= 1;
  void* staging = b ? h->staging2 : h->ws.staging;
  // copy stream: wait until the kernels of two calls ago are done with this buffer, then copy
  LISEC_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_free[b], 0));
  // In pieces: the end-to-end step is bound by this copy (9.6 MB per 8 sweeps beside a kernel that saturates HBM), and on
  // this pool's hosts the rate of a pinned copy is not monotonic in its size (tools/pcie_probe.py). Measured end to end
  // (tools/e2e_probe.py, two boxes): one copy 0.443 / 0.390 ms per step, 3.2 MB pieces 0.380 / 0.372 ms, 1.6 MB 0.375,
  // 2.4 MB and 4.8 MB no better or worse than one copy.
  static const size_t piece = [] {
    const char* e = getenv("LISEC_H2D_PIECE_BYTES");
    const long long v = e ? atoll(e) : 3200000LL;
    return (size_t)(v >= 65536 ? v : 3200000LL) & ~(size_t)255;
  }();
  const size_t step_bytes = bytes > (size_t)24 << 20 ? bytes : piece;  // large copies run at the link rate as they are
  for (size_t off = 0; off < bytes; off += step_bytes) {
    const size_t nb = bytes - off < step_bytes ? bytes - off : step_bytes;
    LISEC_CUDA(h, cudaMemcpyAsync(static_cast<unsigned char*>(staging) + off, static_cast<const unsigned char*>(points_host) + off,
                                  nb, cudaMemcpyHostToDevice, h->copy_stream));
  }
  LISEC_CUDA(h, cudaEventRecord(h->ev_copied[b], h->copy_stream));
  // compute stream: wait for the copy, run the path, release the buffer
  LISEC_CUDA(h, cudaStreamWaitEvent(st, h->ev_copied[b], 0));
  rc = frontend(h, staging, dtype, so, grid, st);
  if (rc) return rc;
  LISEC_CUDA(h, cudaEventRecord(h->ev_free[b], st));
  return LISEC_OK;
}

int32_t lisec_last_fused_kernel_ms(lisec_handle* h, float* ms) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!ms) return fail(h, LISEC_ERR_BAD_ARG, "ms is NULL");
  if (!h->kernel_timed) return fail(h, LISEC_ERR_STATE, "no fused call has run on this handle");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  LISEC_CUDA(h, cudaEventSynchronize(h->ev_kernel[1]));
  LISEC_CUDA(h, cudaEventElapsedTime(ms, h->ev_kernel[0], h->ev_kernel[1]));
  return LISEC_OK;
}

int32_t lisec_debug_trace(lisec_handle* h, int64_t* out, int64_t n) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!h->ws.trace) return fail(h, LISEC_ERR_STATE, "tracing is off: set LISEC_TRACE=1 before lisec_create()");
  if (!out || n < (int64_t)kTraceCtas * kTraceSlots) return fail(h, LISEC_ERR_BAD_ARG, "out too small");
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  LISEC_CUDA(h, cudaDeviceSynchronize());
  LISEC_CUDA(h, cudaMemcpy(out, h->ws.trace, sizeof(int64_t) * kTraceCtas * kTraceSlots, cudaMemcpyDeviceToHost));
  for (int k = 0; k < TL_COUNT; ++k) {  // re-arm the timeline rows: min stamp = ~0, max stamp = 0
    unsigned long long* row = h->ws.trace + (size_t)(kTimelineRow0 + k) * kTraceSlots;
    LISEC_CUDA(h, cudaMemset(row, 0xff, sizeof(unsigned long long)));
    LISEC_CUDA(h, cudaMemset(row + 1, 0, sizeof(unsigned long long)));
  }
  return LISEC_OK;
}

// Debug / test aid: copy one of the grouping tables to the host (synchronous). which: 0 row_start, 1 row_voxel,
// 2 tile_first, 3 tile_row0, 4 chunk_ntiles, 5 chunk_first, 6 voxel_cell.
int32_t lisec_debug_table(lisec_handle* h, int32_t which, int32_t* out, int64_t n) {
  if (!h || !out) return LISEC_ERR_BAD_ARG;
  const int* src[] = {h->ws.row_start, h->ws.row_voxel, h->ws.tile_first, h->ws.tile_row0, h->ws.chunk_ntiles,
                      h->ws.chunk_first, h->ws.voxel_cell};
  if (which < 0 || which > 6) return fail(h, LISEC_ERR_BAD_ARG, "which = %d", which);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  LISEC_CUDA(h, cudaDeviceSynchronize());
  LISEC_CUDA(h, cudaMemcpy(out, src[which], sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
  return LISEC_OK;
}

// Device pointers of what a consumer needs to read the front end's SPARSE output in place (the first Conv3D's gather
// plans, lisec_conv_plan_set_gather): the occupancy map of the last lisec_voxelize() and c_empty. Stable for the handle's life.
int32_t lisec_workspace_pointers(lisec_handle* h, const int32_t** cell_voxel, const float** c_empty, int64_t* max_voxels) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (cell_voxel) *cell_voxel = h->ws.cell_voxel;
  if (c_empty) *c_empty = h->ws.c_empty;
  if (max_voxels) *max_voxels = h->max_voxels;
  return LISEC_OK;
}

int32_t lisec_voxel_counts_async(lisec_handle* h, void* pinned_out, int64_t pinned_bytes, void* stream) {
  if (!h) return LISEC_ERR_BAD_ARG;
  if (!pinned_out) return fail(h, LISEC_ERR_BAD_ARG, "pinned_out is NULL");
  if (!h->voxelized) return fail(h, LISEC_ERR_STATE, "no lisec_voxelize() result on this handle");
  const int64_t need = (int64_t)sizeof(long long) * TOT_COUNT + (int64_t)sizeof(int) * (h->last_so.n + 1);
  if (pinned_bytes < need) return fail(h, LISEC_ERR_BAD_ARG, "pinned_out holds %lld bytes, need %lld", (long long)pinned_bytes, (long long)need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LISEC_CUDA(h, cudaSetDevice(h->cfg.device));
  unsigned char* out = static_cast<unsigned char*>(pinned_out);
  LISEC_CUDA(h, cudaMemcpyAsync(out, h->ws.totals, sizeof(long long) * TOT_COUNT, cudaMemcpyDeviceToHost, st));
  LISEC_CUDA(h, cudaMemcpyAsync(out + sizeof(long long) * TOT_COUNT, h->ws.sweep_voxel_start,
                                sizeof(int) * (h->last_so.n + 1), cudaMemcpyDeviceToHost, st));
  return LISEC_OK;
}

}  // extern "C"
