// Exact box arithmetic shared by boxes.cu and targets.cu (both compiled with -fmad=false).
#pragma once
#include <cuda_runtime.h>

namespace dt {

// ---------------------------------------------------------------- IoU ------
// cython_bbox.pyx:34-56, one (box, query) pair, one frame.
__device__ __forceinline__ float iou_pair_ref(const float* __restrict__ b,
                                              const float* __restrict__ q) {
  // box_area (query): fp64 product rounded to fp32 (:35-38)
  const float qarea = __double2float_rn(
      __dmul_rn(__dadd_rn((double)__fsub_rn(q[2], q[0]), 1.0),
                __dadd_rn((double)__fsub_rn(q[3], q[1]), 1.0)));
  const float iw = __double2float_rn(
      __dadd_rn((double)__fsub_rn(fminf(b[2], q[2]), fmaxf(b[0], q[0])), 1.0));
  if (!(iw > 0.f)) return 0.f;
  const float ih = __double2float_rn(
      __dadd_rn((double)__fsub_rn(fminf(b[3], q[3]), fmaxf(b[1], q[1])), 1.0));
  if (!(ih > 0.f)) return 0.f;
  const float inter = __fmul_rn(iw, ih);
  const double barea = __dmul_rn(__dadd_rn((double)__fsub_rn(b[2], b[0]), 1.0),
                                 __dadd_rn((double)__fsub_rn(b[3], b[1]), 1.0));
  const float ua = __double2float_rn(
      __dsub_rn(__dadd_rn(barea, (double)qarea), (double)inter));
  return __fdiv_rn(inter, ua);
}


}  // namespace dt
