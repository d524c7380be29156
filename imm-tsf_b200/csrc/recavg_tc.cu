// K2, large N x T corner: the helper kernels that put TTF_RecAvg's pooling (fusions/TTF_RecAvg.py:94-102) on the tcgen05
// GEMM.  With N_i notes and T query times per sample the pooling is a dense [T x N_i] . [N_i x d] product per sample
// (TTF_RecAvg.py:100, the einsum) and its backward two more (dV' = W^T dE_raw, and the log-sigma sensitivity); beyond
// N, T ~ 64 that is tensor-core work, not a streaming kernel.  The products run on immtsf_gemm_batched
// (3xTF32, fp32-exact to ~1e-6); this file supplies what surrounds them:
//   immtsf_recavg_weights : Wn[b,t,n] = w_nt / max(sum_n w_nt, 1e-6)  (and Cn = c_nt / den, c_nt = w_nt 2 (delta/sigma)^2,
//                           csum[b,t] = sum_n Cn) as dense [B, T, Np] operands, zero beyond the sample's N_i notes;
//                           the producers also write the operands' lo parts (x - trunc_tf32(x)): no separate split pass
//   immtsf_csr_to_padded  : V' rows of the ragged layout -> [B, Np, d], zero rows beyond N_i (uniform batch strides for the
//                           4-D tensor maps of the batched product)
//   immtsf_padded_to_csr  : the inverse, for dV'
//   immtsf_recavg_dls     : dlog_sigma = sum_{t,j} dE_raw[t,j] (R[t,j] - csum_t E_raw[t,j]) in double, R = Cn V'
// Derivation of the last line: E_raw_t = sum_n Wn_nt V'_n, dWn_nt/dlog_sigma = Cn_nt - Wn_nt csum_t, hence
// dE_raw_t/dlog_sigma = R_t - csum_t E_raw_t.
#include "common.cuh"
#include "../../include/immtsf.h"

namespace {

// the lo operand of the 3xTF32 products: x - trunc_tf32(x) (what immtsf_split_lo writes), made by the producer of x
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// one warp per (sample, query time) row of the weight matrices
__global__ void __launch_bounds__(256) recavg_weights_kernel(const float* __restrict__ tau, const int32_t* __restrict__ offsets,
                                                             const float* __restrict__ t_hat, int t_bstride,
                                                             const float* __restrict__ log_sigma, int B, int T, int Np,
                                                             float* __restrict__ Wn, float* __restrict__ Cn,
                                                             float* __restrict__ Wn_lo, float* __restrict__ Cn_lo,
                                                             float* __restrict__ wsum, float* __restrict__ csum) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  const int nb = offsets[b], cnt = offsets[b + 1] - nb;
  const float inv_sigma = 1.f / expf(__ldg(log_sigma));  // (the forward kernels' expression)
  const float th = t_hat[(size_t)b * t_bstride + t];
  float ws = 0.f;
  for (int n = lane; n < cnt; n += 32) {
    const float r = fmaxf(th - __ldg(tau + nb + n), 0.f) * inv_sigma;
    ws += expf(-(r * r));
  }
  ws = warp_sum(ws);
  const float inv_den = 1.f / fmaxf(ws, 1e-6f);
  float cs = 0.f;
  float* wrow = Wn + row * Np;
  float* crow = Cn ? Cn + row * Np : nullptr;
  for (int n = lane; n < Np; n += 32) {
    float w = 0.f, c = 0.f;
    if (n < cnt) {
      const float r = fmaxf(th - __ldg(tau + nb + n), 0.f) * inv_sigma;
      w = expf(-(r * r)) * inv_den;
      c = w * 2.f * r * r;
    }
    wrow[n] = w;
    if (Wn_lo) Wn_lo[row * Np + n] = tf32_lo(w);
    if (crow) crow[n] = c;
    if (Cn_lo) Cn_lo[row * Np + n] = tf32_lo(c);
    cs += c;
  }
  if (lane == 0 && wsum) wsum[row] = ws;
  if (csum) {
    cs = warp_sum(cs);
    if (lane == 0) csum[row] = cs;
  }
}

// dst[b, n, :] = n < N_b ? src[offsets[b] + n, :] : 0      (float4 columns; grid-stride over B * Np rows)
__global__ void __launch_bounds__(256) csr_to_padded_kernel(const float* __restrict__ src, int lds, const int32_t* __restrict__ offsets,
                                                            int B, int Np, int d4, float* __restrict__ dst, float* __restrict__ dst_lo) {
  const long total = (long)B * Np * d4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long rowp = i / d4;
    const int c = (int)(i % d4), b = (int)(rowp / Np), n = (int)(rowp % Np);
    const int nb = offsets[b], cnt = offsets[b + 1] - nb;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < cnt) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(nb + n) * lds) + c);
    reinterpret_cast<float4*>(dst)[i] = v;
    if (dst_lo) reinterpret_cast<float4*>(dst_lo)[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

// dst[offsets[b] + n, :] = src[b, n, :] for n < N_b
__global__ void __launch_bounds__(256) padded_to_csr_kernel(const float* __restrict__ src, const int32_t* __restrict__ offsets, int B,
                                                            int Np, int d4, float* __restrict__ dst, int ldd) {
  const long total = (long)B * Np * d4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long rowp = i / d4;
    const int c = (int)(i % d4), b = (int)(rowp / Np), n = (int)(rowp % Np);
    const int nb = offsets[b], cnt = offsets[b + 1] - nb;
    if (n < cnt) reinterpret_cast<float4*>(dst + (size_t)(nb + n) * ldd)[c] = __ldg(reinterpret_cast<const float4*>(src) + i);
  }
}

// out += sum_{r, j} dE[r, j] * (R[r, j] - csum[r] * E[r, j]); one warp per row, double from the warp level up, one
// atomicAdd(double) per CTA (the order of those adds is the only non-determinism, at 1e-16 relative)
__global__ void __launch_bounds__(256) recavg_dls_kernel(const float* __restrict__ dE, const float* __restrict__ Rm,
                                                         const float* __restrict__ E, const float* __restrict__ csum, long rows,
                                                         int d4, double* __restrict__ out) {
  __shared__ double s_red[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc = 0.0;
  for (long r = (long)blockIdx.x * 8 + w; r < rows; r += (long)gridDim.x * 8) {
    const float cs = csum[r];
    const float4* a = reinterpret_cast<const float4*>(dE) + r * d4;
    const float4* rr = reinterpret_cast<const float4*>(Rm) + r * d4;
    const float4* e = reinterpret_cast<const float4*>(E) + r * d4;
    float part = 0.f;
    for (int c = lane; c < d4; c += 32) {
      const float4 x = __ldg(a + c), y = __ldg(rr + c), z = __ldg(e + c);
      part = fmaf(x.x, fmaf(-cs, z.x, y.x), part);
      part = fmaf(x.y, fmaf(-cs, z.y, y.y), part);
      part = fmaf(x.z, fmaf(-cs, z.z, y.z), part);
      part = fmaf(x.w, fmaf(-cs, z.w, y.w), part);
    }
    acc += (double)part;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_red[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_red[i];
    atomicAdd(out, t);
  }
}

}  // namespace

extern "C" int immtsf_recavg_weights(const float* tau_flat, const int32_t* offsets, const float* t_hat, int t_hat_bstride,
                                     const float* log_sigma, int B, int T, int Np, float* Wn, float* Cn, float* Wn_lo,
                                     float* Cn_lo, float* wsum, float* csum, void* stream) {
  if (B <= 0 || T <= 0 || Np <= 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(tau_flat && offsets && t_hat && log_sigma && Wn, "recavg_weights: null pointer");
  IMMTSF_REQUIRE((Cn == nullptr) == (csum == nullptr), "recavg_weights: Cn and csum go together");
  const long rows = (long)B * T;
  recavg_weights_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(tau_flat, offsets, t_hat, t_hat_bstride, log_sigma,
                                                                                     B, T, Np, Wn, Cn, Wn_lo, Cn_lo, wsum, csum);
  IMMTSF_CHECK_LAUNCH("recavg_weights");
  return IMMTSF_OK;
}

extern "C" int immtsf_csr_to_padded(const float* src, int lds, const int32_t* offsets, int B, int Np, int d, float* dst,
                                    float* dst_lo, void* stream) {
  if (B <= 0 || Np <= 0 || d <= 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(src && offsets && dst, "csr_to_padded: null pointer");
  IMMTSF_REQUIRE((d & 3) == 0 && (lds & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)dst_lo & 15) == 0,
                 "csr_to_padded: d and lds must be multiples of 4, pointers 16B aligned");
  const long total = (long)B * Np * (d >> 2);
  const int grid = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
  csr_to_padded_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, offsets, B, Np, d >> 2, dst, dst_lo);
  IMMTSF_CHECK_LAUNCH("csr_to_padded");
  return IMMTSF_OK;
}

extern "C" int immtsf_padded_to_csr(const float* src, const int32_t* offsets, int B, int Np, int d, float* dst, int ldd,
                                    void* stream) {
  if (B <= 0 || Np <= 0 || d <= 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(src && offsets && dst, "padded_to_csr: null pointer");
  IMMTSF_REQUIRE((d & 3) == 0 && (ldd & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0,
                 "padded_to_csr: d and ldd must be multiples of 4, pointers 16B aligned");
  const long total = (long)B * Np * (d >> 2);
  const int grid = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
  padded_to_csr_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, offsets, B, Np, d >> 2, dst, ldd);
  IMMTSF_CHECK_LAUNCH("padded_to_csr");
  return IMMTSF_OK;
}

extern "C" int immtsf_recavg_dls(const float* dE_raw, const float* R, const float* E_raw, const float* csum, long rows, int d,
                                 double* dlog_sigma, void* stream) {
  if (rows <= 0 || d <= 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dE_raw && R && E_raw && csum && dlog_sigma, "recavg_dls: null pointer");
  IMMTSF_REQUIRE((d & 3) == 0 && ((uintptr_t)dE_raw & 15) == 0 && ((uintptr_t)R & 15) == 0 && ((uintptr_t)E_raw & 15) == 0,
                 "recavg_dls: d must be a multiple of 4, pointers 16B aligned");
  const long want = (rows + 7) / 8;
  const int grid = (int)(want < 148L * 8 ? want : 148L * 8);
  recavg_dls_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dE_raw, R, E_raw, csum, rows, d >> 2, dlog_sigma);
  IMMTSF_CHECK_LAUNCH("recavg_dls");
  return IMMTSF_OK;
}
