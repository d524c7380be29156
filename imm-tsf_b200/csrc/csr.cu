// K1: padded [B,N,d_m] notes -> ragged CSR layout.
// Replaces the content mask `(V.abs().sum(dim=2) > 0)` of
// fusions/TTF_RecAvg.py:69 and fusions/TTF_T2V_XAttn.py:107, the NaN guard
// (:75 / :116) and `M_txt = note_mask.any(dim=1)` (:110 / :124).
//
// sum_k |v_k| > 0  <=>  some v_k != 0 (no NaN present; a sum of non-negative
// floats cannot underflow to 0 and +inf > 0), so the mask is computed as
// "any element non-zero" -- bit-exact with the reference on NaN-free input,
// and NaN input raises on both sides.
//
// HBM traffic: pass 1 reads B*N*d_m*4 bytes once (float4, one warp per row);
// pass 3 re-reads the valid rows (L2-resident at Time-IMM sizes) and writes
// sumN*d_m*4 bytes compacted.
#include "common.cuh"
#include "../../include/immtsf.h"

// one warp per (b,n) row
__global__ void csr_mask_kernel(const float* __restrict__ notes, int rows_total, int d_m,
                                uint8_t* __restrict__ note_mask, int32_t* __restrict__ flags) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows_total) return;
  const float* row = notes + (size_t)warp * d_m;
  bool nz = false, bad = false;
  if ((d_m & 3) == 0 && ((uintptr_t)row & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int i = lane; i < (d_m >> 2); i += 32) {
      const float4 v = __ldg(r4 + i);
      nz |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
      bad |= isnan(v.x) | isnan(v.y) | isnan(v.z) | isnan(v.w);
    }
  } else {
    for (int i = lane; i < d_m; i += 32) {
      const float v = __ldg(row + i);
      nz |= (v != 0.f);
      bad |= isnan(v);
    }
  }
  nz = __any_sync(0xffffffffu, nz);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    // a NaN row has sum == NaN and `NaN > 0` is False in the reference
    note_mask[warp] = (nz && !bad) ? 1 : 0;
    if (bad) flags[IMMTSF_FLAG_V] = 1;
  }
}

// single CTA: per-sample counts -> exclusive scan -> offsets, rows, seg, tau_flat, m_txt
__global__ void csr_scan_kernel(const uint8_t* __restrict__ note_mask, const float* __restrict__ tau, int B, int N,
                                int32_t* __restrict__ offsets, int32_t* __restrict__ rows, int32_t* __restrict__ seg,
                                float* __restrict__ tau_flat, uint8_t* __restrict__ m_txt, int M_alloc) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) { s_carry = 0; offsets[0] = 0; }
  __syncthreads();
  // phase 1: counts + block scan over samples, blockDim samples at a time
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int cnt = 0;
    if (b < B) {
      const uint8_t* mrow = note_mask + (size_t)b * N;
      for (int n = 0; n < N; ++n) cnt += mrow[n];
      m_txt[b] = cnt > 0 ? 1 : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
      int v = lane < nw ? s_warp[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      s_warp[lane] = v;  // inclusive over warps
    }
    __syncthreads();
    const int warp_excl = w == 0 ? 0 : s_warp[w - 1];
    const int carry = s_carry;
    if (b < B) offsets[b + 1] = carry + warp_excl + incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + warp_excl + incl;
    __syncthreads();
  }
  // rows / seg / tau_flat / emb_flat are filled by csr_gather_kernel (whole grid); here only the pad of tau_flat
  __syncthreads();
  const int total = s_carry;
  int end = (total + 127) / 128 * 128;
  if (end > M_alloc) end = M_alloc;
  for (int i = total + threadIdx.x; i < end; i += blockDim.x) tau_flat[i] = 0.f;
}

// Source-driven compaction, one warp per padded row (b, n): a valid row finds its destination
// p = offsets[b] + #valid rows before it in its sample (ballot/popcount over the sample's mask bytes), copies its
// embedding with 128-bit accesses and writes rows[p], seg[p], tau_flat[p].  Warps beyond B*N zero the pad rows
// [sumN, roundup(sumN, 128)) of emb_flat.
__global__ void __launch_bounds__(256) csr_gather_kernel(const float* __restrict__ notes, const float* __restrict__ tau,
                                                         const uint8_t* __restrict__ note_mask,
                                                         const int32_t* __restrict__ offsets, int B, int N, int d_m,
                                                         int32_t* __restrict__ rows, int32_t* __restrict__ seg,
                                                         float* __restrict__ tau_flat, float* __restrict__ emb_flat,
                                                         int ld_emb, float* __restrict__ emb_lo, int ld_lo, int M_alloc) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int total_src = B * N;
  const int total = offsets[B];
  int end = (total + 127) / 128 * 128;
  if (end > M_alloc) end = M_alloc;
  const int npad = end - total;
  for (int i = gw; i < total_src + npad; i += nw) {
    float* dst;
    const float* src = nullptr;
    if (i < total_src) {
      if (!note_mask[i]) continue;
      const int b = i / N, n = i - b * N;
      const uint8_t* mrow = note_mask + (size_t)b * N;
      int before = 0;
      for (int n0 = 0; n0 < n; n0 += 32) {
        const bool v = n0 + lane < n && mrow[n0 + lane];
        before += __popc(__ballot_sync(0xffffffffu, v));
      }
      const int p = offsets[b] + before;
      if (lane == 0) { rows[p] = i; seg[p] = b; tau_flat[p] = tau[i]; }
      dst = emb_flat + (size_t)p * ld_emb;
      src = notes + (size_t)i * d_m;
    } else {
      dst = emb_flat + (size_t)(total + (i - total_src)) * ld_emb;
    }
    if (emb_flat == nullptr) continue;
    // optional second output: the row's tcgen05 lo operand x - trunc_tf32(x) (the consumer's 3xTF32 product needs it)
    float* dlo = emb_lo != nullptr ? emb_lo + (size_t)(dst - emb_flat) / ld_emb * ld_lo : nullptr;
    if ((d_m & 3) == 0 && ((uintptr_t)dst & 15) == 0 && (src == nullptr || ((uintptr_t)src & 15) == 0) &&
        (dlo == nullptr || ((uintptr_t)dlo & 15) == 0)) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      float4* l4 = reinterpret_cast<float4*>(dlo);
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int k = lane; k < (d_m >> 2); k += 32) {
        const float4 v = src ? __ldg(s4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        d4[k] = v;
        if (dlo != nullptr)
          l4[k] = make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                              v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
      }
    } else {
      for (int k = lane; k < d_m; k += 32) {
        const float v = src ? __ldg(src + k) : 0.f;
        dst[k] = v;
        if (dlo != nullptr) dlo[k] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      }
    }
  }
}

extern "C" int immtsf_csr_build(const float* notes, const float* tau, int B, int N, int d_m, uint8_t* note_mask,
                                int32_t* offsets, int32_t* rows, int32_t* seg, float* emb_flat, float* tau_flat,
                                uint8_t* m_txt, int32_t* flags, int M_alloc, void* stream) {
  return immtsf_csr_build_ex(notes, tau, B, N, d_m, note_mask, offsets, rows, seg, emb_flat, d_m, nullptr, 0, tau_flat, m_txt, flags,
                             M_alloc, stream);
}

extern "C" int immtsf_csr_build_ex(const float* notes, const float* tau, int B, int N, int d_m, uint8_t* note_mask,
                                   int32_t* offsets, int32_t* rows, int32_t* seg, float* emb_flat, int ld_emb, float* emb_lo,
                                   int ld_lo, float* tau_flat, uint8_t* m_txt, int32_t* flags, int M_alloc, void* stream) {
  IMMTSF_REQUIRE(ld_emb >= d_m && (emb_lo == nullptr || ld_lo >= d_m), "csr_build: leading dimension smaller than d_model");
  IMMTSF_REQUIRE(B >= 0 && N >= 0 && d_m >= 0, "csr_build: negative size");
  IMMTSF_REQUIRE(offsets && m_txt && flags, "csr_build: null output");
  IMMTSF_REQUIRE(M_alloc >= B * N, "csr_build: M_alloc (%d) < B*N (%d)", M_alloc, B * N);
  cudaStream_t st = (cudaStream_t)stream;
  const int total_rows = B * N;
  if (total_rows > 0 && d_m > 0) {
    IMMTSF_REQUIRE(notes && tau && note_mask && rows && seg && tau_flat, "csr_build: null pointer");
    csr_mask_kernel<<<ceil_div(total_rows, 8), 256, 0, st>>>(notes, total_rows, d_m, note_mask, flags);
    IMMTSF_CHECK_LAUNCH("csr_mask");
  } else if (total_rows > 0) {
    cudaMemsetAsync(note_mask, 0, total_rows, st);
  }
  csr_scan_kernel<<<1, 1024, 0, st>>>(note_mask, tau, B, N, offsets, rows, seg, tau_flat, m_txt, M_alloc);
  IMMTSF_CHECK_LAUNCH("csr_scan");
  if (total_rows > 0) {
    int grid = ceil_div(total_rows + 128, 8);
    if (grid > 148 * 16) grid = 148 * 16;
    csr_gather_kernel<<<grid, 256, 0, st>>>(notes, tau, note_mask, offsets, B, N, d_m, rows, seg, tau_flat,
                                            d_m > 0 ? emb_flat : nullptr, ld_emb, d_m > 0 ? emb_lo : nullptr, ld_lo, M_alloc);
    IMMTSF_CHECK_LAUNCH("csr_gather");
  }
  return IMMTSF_OK;
}
