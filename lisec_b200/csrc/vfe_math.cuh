// Device helpers shared by the VFE kernel and the export kernel, so that both produce bit-identical feature rows.
#pragma once

#include "common.cuh"

namespace lisec {

// One point (three PT, 3 * sizeof(PT)-byte stride) in two L2 loads instead of three: the pair that is 2 * sizeof(PT)
// aligned as one vector load. L2 loads (.cg): the caller rewrites its point buffer between calls (common.cuh).
template <typename PT>
__device__ __forceinline__ void load_point(const PT* __restrict__ pts, long long p, PT& x, PT& y, PT& z);
template <>
__device__ __forceinline__ void load_point<float>(const float* __restrict__ pts, long long p, float& x, float& y, float& z) {
  const float* q = pts + 3 * p;
  if ((p & 1) == 0) {  // 12 p bytes from a 16-byte aligned base: 8-byte aligned for even p
    const float2 a = __ldcg(reinterpret_cast<const float2*>(q));
    x = a.x; y = a.y; z = __ldcg(q + 2);
  } else {
    x = __ldcg(q);
    const float2 a = __ldcg(reinterpret_cast<const float2*>(q + 1));
    y = a.x; z = a.y;
  }
}
template <>
__device__ __forceinline__ void load_point<double>(const double* __restrict__ pts, long long p, double& x, double& y, double& z) {
  const double* q = pts + 3 * p;
  if ((p & 1) == 0) {  // 24 p bytes: 16-byte aligned for even p
    const double2 a = __ldcg(reinterpret_cast<const double2*>(q));
    x = a.x; y = a.y; z = __ldcg(q + 2);
  } else {
    x = __ldcg(q);
    const double2 a = __ldcg(reinterpret_cast<const double2*>(q + 1));
    y = a.x; z = a.y;
  }
}

// [x, y, z, x-cx, y-cy, z-cz]: subtraction in float64, one rounding to float32 (model_training.py:137-140 and the
// float32 cast at the Keras model input)
__device__ __forceinline__ void point_features(double x, double y, double z, double cx, double cy, double cz,
                                               float (&f)[6]) {
  f[0] = __double2float_rn(x);
  f[1] = __double2float_rn(y);
  f[2] = __double2float_rn(z);
  f[3] = __double2float_rn(x - cx);
  f[4] = __double2float_rn(y - cy);
  f[5] = __double2float_rn(z - cz);
}

}  // namespace lisec
