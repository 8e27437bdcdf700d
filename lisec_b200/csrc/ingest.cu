// Lidar ingest pre-pass on sm_100a: what combine_lidar_data + rotate_points do to the raw sensor
// files before the voxelizer sees a point (reference model_training.py:65-98):
//
//     rawPoints = np.fromfile(path, float32).reshape(-1, 5)[:, :3]        :87-90   (x, y, z, intensity, ring)
//     points    = np.dot(Quaternion(rotation).rotation_matrix, rawPoints.T).T   :65-69, :93
//     points    = points + np.array(translation)                               :94
//     allPoints = np.concatenate(...)                                          :96
//
// One launch takes the concatenated records of up to kIngestMaxSegments (sensor, sweep) segments, each with its own
// 3x3 float64 rotation matrix and translation (kernel parameters: no device table to upload), and writes the float64
// (n, 3) points lisec_voxelize()/lisec_frontend_forward() consume as LISEC_F64. The matrix itself is built on the host
// from the quaternion (3 sensors per sweep: not device work; lisec_b200/ingest.py).
//
// Arithmetic, bit for bit what numpy does: float32 -> float64 widening (exact), then per output coordinate
//     acc = R[i][0]*x;  acc = fma(R[i][1], y, acc);  acc = fma(R[i][2], z, acc);  out = acc + t[i]
// which is the k-ascending FMA chain OpenBLAS's dgemm micro-kernels evaluate for the (3,3)x(3,n) product (pinned by
// tests/test_oracle.py against np.dot itself), followed by the float64 add of :94.
//
// HBM-bound, 20 B read + 24 B written per point. A block takes 256 consecutive points: the 1280 record floats arrive
// by coalesced 4-byte loads into shared memory (stride-5 reads of it are conflict-free), the 768 doubles leave by
// coalesced 8-byte stores.
#include <cstdio>

#include "common.cuh"

namespace lisec {

namespace {

constexpr int kIngestThreads = 256;
constexpr int kIngestMaxSegments = 24;  // 8 sweeps x 3 sensors per launch; more segments = more launches

struct IngestSegments {
  long long off[kIngestMaxSegments + 1];  // point offsets relative to the launch's first point
  double rot[kIngestMaxSegments][9];      // row-major rotation matrix
  double trans[kIngestMaxSegments][3];
  int n;
};

__global__ void __launch_bounds__(kIngestThreads)
    ingest_kernel(const float* __restrict__ rec, int rec_floats, long long n_points,
                  const __grid_constant__ IngestSegments seg, double* __restrict__ out) {
  pdl_launch_dependents();
  __shared__ float s_in[kIngestThreads * 8];
  __shared__ double s_out[kIngestThreads * 3];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * kIngestThreads;
  const int here = (int)min((long long)kIngestThreads, n_points - base);
  pdl_wait();
  const float* src = rec + base * rec_floats;
  for (int i = tid; i < here * rec_floats; i += kIngestThreads) s_in[i] = __ldcg(src + i);
  __syncthreads();
  if (tid < here) {
    const long long p = base + tid;
    int s = 0;
    while (s + 1 < seg.n && base >= seg.off[s + 1]) ++s;  // uniform over the block: the segment of its first point
    while (s + 1 < seg.n && p >= seg.off[s + 1]) ++s;     // a block rarely straddles a file boundary
    const double x = (double)s_in[tid * rec_floats], y = (double)s_in[tid * rec_floats + 1],
                 z = (double)s_in[tid * rec_floats + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double acc = __dmul_rn(seg.rot[s][3 * i], x);
      acc = __fma_rn(seg.rot[s][3 * i + 1], y, acc);
      acc = __fma_rn(seg.rot[s][3 * i + 2], z, acc);
      s_out[tid * 3 + i] = __dadd_rn(acc, seg.trans[s][i]);
    }
  }
  __syncthreads();
  double* dst = out + base * 3;
  for (int i = tid; i < here * 3; i += kIngestThreads) dst[i] = s_out[i];
}

thread_local char g_ingest_error[256] = "";

int32_t ingest_fail(int32_t code, const char* msg) {
  snprintf(g_ingest_error, sizeof(g_ingest_error), "%s", msg);
  return code;
}

}  // namespace

}  // namespace lisec

using namespace lisec;

extern "C" {

const char* lisec_ingest_last_error(void) { return g_ingest_error; }

int32_t lisec_ingest_lidar(const float* records, int32_t record_floats, const int64_t* segment_offsets,
                           const lisec_sensor_pose* poses, int32_t n_segments, double* points, void* stream,
                           int32_t* launches_out) {
  if (launches_out) *launches_out = 0;
  if (!segment_offsets || !poses || n_segments < 0) return ingest_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (record_floats < 3 || record_floats > 8)
    return ingest_fail(LISEC_ERR_BAD_ARG, "record_floats must be 3..8 (the Lyft .bin files hold 5 float32 per point)");
  for (int s = 0; s < n_segments; ++s)
    if (segment_offsets[s + 1] < segment_offsets[s] || segment_offsets[0] != 0)
      return ingest_fail(LISEC_ERR_BAD_ARG, "segment_offsets must start at 0 and be non-decreasing");
  if (n_segments == 0 || segment_offsets[n_segments] == 0) return LISEC_OK;
  if (!records || !points) return ingest_fail(LISEC_ERR_BAD_ARG, "null device pointer");
  int launches = 0;
  for (int s0 = 0; s0 < n_segments; s0 += kIngestMaxSegments) {
    IngestSegments seg;
    seg.n = n_segments - s0 < kIngestMaxSegments ? n_segments - s0 : kIngestMaxSegments;
    const long long first = segment_offsets[s0];
    for (int s = 0; s <= seg.n; ++s) seg.off[s] = segment_offsets[s0 + s] - first;
    for (int s = 0; s < seg.n; ++s) {
      for (int k = 0; k < 9; ++k) seg.rot[s][k] = poses[s0 + s].rotation[k];
      for (int k = 0; k < 3; ++k) seg.trans[s][k] = poses[s0 + s].translation[k];
    }
    const long long n = seg.off[seg.n];
    if (n == 0) continue;
    const long long blocks = (n + kIngestThreads - 1) / kIngestThreads;
    cudaError_t e = launch_pdl(ingest_kernel, dim3((unsigned)blocks), dim3(kIngestThreads), 0,
                               static_cast<cudaStream_t>(stream), records + first * record_floats, (int)record_floats,
                               n, seg, points + first * 3);
    if (e != cudaSuccess) return ingest_fail(LISEC_ERR_CUDA, cudaGetErrorString(e));
    ++launches;
  }
  if (launches_out) *launches_out = launches;
  return LISEC_OK;
}

}  // extern "C"
