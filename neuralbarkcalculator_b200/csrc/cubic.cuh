// torch's upsample_bicubic2d index / weight arithmetic (A = -0.75, align_corners=False) in explicitly rounded f32
// operations, shared by the forward (head.cu) and its adjoint (train_kernels2.cu).
#pragma once
#include "common.cuh"

namespace nbc {

// torch cubic convolution coefficients, A = -0.75, evaluated with separately rounded f32 operations
__device__ __forceinline__ float cc1(float x) {  // |x| <= 1 : ((A+2)x - (A+3)) x x + 1
  const float a = __fsub_rn(__fmul_rn(1.25f, x), 2.25f);
  return __fadd_rn(__fmul_rn(__fmul_rn(a, x), x), 1.f);
}
__device__ __forceinline__ float cc2(float x) {  // 1 < |x| < 2 : ((A x - 5A) x + 8A) x - 4A
  const float a = __fsub_rn(__fmul_rn(-0.75f, x), -3.75f);
  const float b = __fadd_rn(__fmul_rn(a, x), -6.f);
  return __fsub_rn(__fmul_rn(b, x), -3.f);
}
__device__ __forceinline__ void cubic_taps(int dst, float scale, int in_size, int (&idx)[4], float (&wt)[4]) {
  const float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  const float fl = floorf(src);
  const float t = __fsub_rn(src, fl);
  const int i0 = (int)fl;
  const float omt = __fsub_rn(1.f, t);
  wt[0] = cc2(__fadd_rn(t, 1.f));
  wt[1] = cc1(t);
  wt[2] = cc1(omt);
  wt[3] = cc2(__fadd_rn(omt, 1.f));
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), in_size - 1);
}
}  // namespace nbc
