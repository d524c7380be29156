// "One warp owns one row" helpers for the row kernels over d <= 1024 columns (LayerNorm fwd/bwd, RecAvg pooling):
// lane l owns the float8 chunks k = l + 32*i (i < NC), i.e. columns 8k .. 8k+7.  No shared memory and no block
// barrier: row statistics are warp-shuffle reductions, and one Philox call covers a whole chunk (common.cuh).
// Compared with the CTA-per-row-tile kernels (rowtile.cuh: one float4 per thread, two block reductions per tile)
// this cuts the instructions issued per row ~3x -- ncu showed those kernels issue-bound, not memory-bound.
#pragma once
#include "rowtile.cuh"

__device__ __forceinline__ void load8(const float* __restrict__ row, int k, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(row) + 2 * k);
  const float4 b = __ldg(reinterpret_cast<const float4*>(row) + 2 * k + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void store8(float* __restrict__ row, int k, const float (&v)[8]) {
  reinterpret_cast<float4*>(row)[2 * k] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(row)[2 * k + 1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void zero8(float (&o)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
}
// number of float8 chunks per lane for a row of d columns (d % 8 == 0, d <= 1024), 0 if unsupported
static inline int rowwarp_nc(int d) {
  if (d <= 0 || (d & 7) || d > 1024) return 0;
  return ((d >> 3) + 31) / 32;
}
