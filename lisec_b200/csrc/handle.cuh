// The handle behind the C ABI and the helpers every ABI source file shares (api.cu, vfe_train.cu). Not installed.
#pragma once

#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace lisec {
struct VfeTrainState;  // vfe_train.cu: activations kept between the training forward and backward
}

using namespace lisec;

struct lisec_handle {
  lisec_config cfg;
  Geom geom;
  Workspace ws;
  VfeSmall params;
  float wblob[kVfeBlobFloats];
  int sm_count = 0;
  int rows_per_chunk = 0;
  long long max_voxels = 0;
  long long max_chunks = 0;
  long long ncells_cap = 0;
  int scan_blocks_cap = 0;
  int64_t workspace_bytes = 0;
  bool weights_set = false;
  bool voxelized = false;
  bool count_dirty = true;  // count table must be zero before a point pass; the fill pass leaves it zero
  // last lisec_voxelize() inputs (export / VFE gather from them)
  const void* last_points = nullptr;
  int last_dtype = LISEC_F32;
  SweepOffsets last_so;
  int launches = 0;
  // host-input pipeline: copies go on their own stream into alternating staging buffers, so the H2D copy of call i+1
  // overlaps the kernels of call i
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};  // staging[b] holds the new points
  cudaEvent_t ev_free[2] = {nullptr, nullptr};    // the kernels that read staging[b] have been enqueued and finished
  void* staging2 = nullptr;                       // second staging buffer (the first is ws.staging)
  int staging_idx = 0;
  // CUDA events around the fused VFE + grid kernel of the last fused call (bench.py's live roofline figure)
  cudaEvent_t ev_kernel[2] = {nullptr, nullptr};
  bool kernel_timed = false;
  // the blind background prefix (an experiment, off by default: LISEC_BLIND_FRACTION): a side stream fills the first
  // `blind_fraction` of the grid while the grouping chain runs. Measured: the chain makes no progress beside the fill
  // (0.357 -> 0.401 / 0.426 / 0.457 ms per step at 0.2 / 0.33 / 0.45), so nothing is gained.
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_side[2] = {nullptr, nullptr};
  double blind_fraction = 0.0;
  lisec::VfeTrainState* train = nullptr;  // allocated by the first lisec_vfe_train_forward()
  // the float32 VFE kernel for every graph but the current createModel's (vfe_generic.cu); also under LISEC_GENERIC_VFE=1
  bool generic = false;
  float* generic_params = nullptr;  // device parameter block (GenericLayout)
  char err[512];
};

inline int fail(lisec_handle* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

inline int cuda_fail(lisec_handle* h, cudaError_t e, const char* what) {
  return fail(h, LISEC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define LISEC_CUDA(h, call)                                 \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);  \
  } while (0)


template <typename T>
inline cudaError_t dev_alloc(lisec_handle* h, T** p, size_t n) {
  size_t bytes = n * sizeof(T);
  bytes = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
  if (e == cudaSuccess) h->workspace_bytes += (int64_t)bytes;
  return e;
}

namespace lisec {
void free_vfe_train_state(VfeTrainState* s);
}
