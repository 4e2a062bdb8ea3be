// Visibility predicate pieces shared by the standalone map kernel (kernels_vis.cu), the
// pre-step kernel (kernels_env.cu) and the fused step kernel (kernels_ukf.cu).
#pragma once
#include "pc_common.cuh"

namespace pc {

// Horizon term of one agent with the reference's below-surface semantics:
// |r| >= RE -> sqrt(|r|^2 - RE^2); RE - |r| <= tol -> 0 (clamped to the surface);
// deeper -> NaN (acos(>1) in the reference: the pair is never visible).
__device__ __forceinline__ double horizon_q(double x, double y, double z, double RE, double tol) {
  const double rm = sqrt(fma(z, z, fma(y, y, x * x)));
  if (rm >= RE) return sqrt((rm - RE) * (rm + RE));
  if (RE - rm <= tol) return 0.0;
  return __longlong_as_double(0x7ff8000000000000LL);
}

// visible  <=>  r_s . r_t > RE^2 - q_s q_t      (see kernels_vis.cu header)
__device__ __forceinline__ bool sees(const double4& s, double tx, double ty, double tz, double tq,
                                     double RE2) {
  return fma(s.z, tz, fma(s.y, ty, s.x * tx)) > fma(-s.w, tq, RE2);
}

}  // namespace pc
