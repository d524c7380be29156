// "CTA owns TT rows, each thread owns NCH float4 column groups" helpers shared
// by the RecAvg pooling kernels and the LayerNorm-over-d kernels.
#pragma once
#include "common.cuh"

// Sum TT per-thread partials over the whole CTA; every thread gets all TT sums.
// red: >= 32*TT floats of shared memory.
template <int TT>
__device__ __forceinline__ void block_sum_multi(float (&v)[TT], float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int t = 0; t < TT; ++t) v[t] = warp_sum(v[t]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < TT; ++t) red[w * TT + t] = v[t];
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ++ww) s += red[ww * TT + t];
    v[t] = s;
  }
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float f4_sum(const float4& a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void f4_fma(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ void f4_add(float4& acc, const float4& v) {
  acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
}

// dropout keep-scales for the 4 consecutive elements whose flat index / 4 == idx4 (half of a Philox call, common.cuh)
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint32_t site, uint64_t idx4, uint32_t thr, float inv_keep) {
  if (thr == 0u) return make_float4(1.f, 1.f, 1.f, 1.f);
  const uint64_t c = idx4 >> 1;
  const Philox4 r = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), site, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t w0 = (idx4 & 1) ? r.z : r.x, w1 = (idx4 & 1) ? r.w : r.y;
  return make_float4((w0 & 0xFFFFu) >= thr ? inv_keep : 0.f, (w0 >> 16) >= thr ? inv_keep : 0.f,
                     (w1 & 0xFFFFu) >= thr ? inv_keep : 0.f, (w1 >> 16) >= thr ? inv_keep : 0.f);
}
__host__ __device__ __forceinline__ float inv_keep_from_thr(uint32_t thr) {
  // realised drop rate = thr / 2^16 ; 1/(1-rate)
  return thr == 0u ? 1.f : (float)(1.0 / (1.0 - (double)thr / 65536.0));
}
