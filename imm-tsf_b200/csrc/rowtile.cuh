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

// ---- packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE operations per issued instruction, each lane rounded like the
// scalar op).  The row kernels are bound by instruction issue, so their elementwise LayerNorm / dropout arithmetic runs
// on float4 halves.
__device__ __forceinline__ float2 lo2(const float4& a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ float2 hi2(const float4& a) { return make_float2(a.z, a.w); }
__device__ __forceinline__ float4 cat4(const float2& l, const float2& h) { return make_float4(l.x, l.y, h.x, h.y); }
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
// a * s
__device__ __forceinline__ float4 f4_muls(const float4& a, float s) { return cat4(__fmul2_rn(lo2(a), bc2(s)), __fmul2_rn(hi2(a), bc2(s))); }
// a + s
__device__ __forceinline__ float4 f4_adds(const float4& a, float s) { return cat4(__fadd2_rn(lo2(a), bc2(s)), __fadd2_rn(hi2(a), bc2(s))); }
// a * b
__device__ __forceinline__ float4 f4_mul(const float4& a, const float4& b) { return cat4(__fmul2_rn(lo2(a), lo2(b)), __fmul2_rn(hi2(a), hi2(b))); }
// a * b + c
__device__ __forceinline__ float4 f4_fma3(const float4& a, const float4& b, const float4& c) {
  return cat4(__ffma2_rn(lo2(a), lo2(b), lo2(c)), __ffma2_rn(hi2(a), hi2(b), hi2(c)));
}
// a * s + c (s scalar)
__device__ __forceinline__ float4 f4_fmas(const float4& a, float s, const float4& c) {
  return cat4(__ffma2_rn(lo2(a), bc2(s), lo2(c)), __ffma2_rn(hi2(a), bc2(s), hi2(c)));
}
// a * s + t (s, t scalars)
__device__ __forceinline__ float4 f4_fmass(const float4& a, float s, float t) {
  return cat4(__ffma2_rn(lo2(a), bc2(s), bc2(t)), __ffma2_rn(hi2(a), bc2(s), bc2(t)));
}
// acc2 += lo(a) + hi(a)  (pairwise partial sums of a row; the caller adds acc2.x + acc2.y at the end)
__device__ __forceinline__ void f2_acc_sum(float2& acc2, const float4& a) { acc2 = __fadd2_rn(acc2, __fadd2_rn(lo2(a), hi2(a))); }
// acc2 += lo(a)*lo(b) + hi(a)*hi(b)
__device__ __forceinline__ void f2_acc_dot(float2& acc2, const float4& a, const float4& b) {
  acc2 = __ffma2_rn(lo2(a), lo2(b), acc2);
  acc2 = __ffma2_rn(hi2(a), hi2(b), acc2);
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
