// Device helpers shared by the VFE kernel and the export kernel, so that both produce bit-identical feature rows.
#pragma once

#include "common.cuh"

namespace lisec {

template <typename PT>
__device__ __forceinline__ void load_point(const PT* __restrict__ pts, long long p, PT& x, PT& y, PT& z) {
  x = __ldcg(pts + 3 * p);  // (L2 loads: the caller rewrites its point buffer between calls)
  y = __ldcg(pts + 3 * p + 1);
  z = __ldcg(pts + 3 * p + 2);
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
