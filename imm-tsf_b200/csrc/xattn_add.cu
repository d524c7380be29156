// K6: MMF_XAttn_Add core and tail (fusions/MMF_XAttn_Add.py:73-102).
//
// Core = what nn.MultiheadAttention does between its in- and out-projections
// (those are immtsf_gemm calls): per (sample, head) S = (q * hd^-1/2) k^T over
// the T x T query/key grid, softmax, dropout on the weights, O = P v.  The
// key-padding mask of the reference is all-or-nothing per sample (:73): a
// sample without text has every key masked, its softmax is NaN and the
// reference overwrites the output with zeros (:79-80) -- here such samples
// simply produce zeros (and zero gradients).
// Forward: one CTA per (b, h, tile of 8 query rows); scores live in shared
// memory; probabilities are saved for backward.  Backward: one CTA per (b,h)
// walking the query tiles so dk/dv accumulate without atomics.
//
// Tail: delta = dropout(LayerNorm_C(delta_y)) (zeroed without text),
//       Y_out = (Y + kappa * delta) / (1 + kappa).
#include "rowtile.cuh"
#include "../../include/immtsf.h"

constexpr int XQ = 8;  // query rows per tile

// T <= 32 fast path (xattn_small.cu)
int immtsf_xattn_small_ok(int T, int H, int d);
int immtsf_xattn_small_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const uint8_t* m_txt, int B,
                           int T, int H, int d, uint32_t drop_thr, uint64_t seed, float* o, int ldo, float* probs,
                           cudaStream_t st);
int immtsf_xattn_small_bwd(const float* d_o, int lddo, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                           const float* probs, const uint8_t* m_txt, int B, int T, int H, int d, uint32_t drop_thr,
                           uint64_t seed, float* dq, int lddq, float* dk, int lddk, float* dv, int lddv, cudaStream_t st);

struct XArgs {
  const float* q; int ldq; const float* k; int ldk; const float* v; int ldv;
  const uint8_t* m_txt; int B, T, H, d, hd; uint32_t thr; SeedArg seed; float scale;
  float* o; int ldo; float* probs;
  const float* d_o; int lddo; float* dq; int lddq; float* dk; int lddk; float* dv; int lddv;
};

__device__ __forceinline__ float warp_dot2(const float* __restrict__ a, const float* __restrict__ b, int n, int lane) {
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = fmaf(a[j], b[j], s);
  return warp_sum(s);
}

// smem: s_s [XQ][T]
__global__ void __launch_bounds__(256) xattn_core_fwd_kernel(const XArgs a) {
  extern __shared__ float smem[];
  const int T = a.T, hd = a.hd, H = a.H;
  float* s_s = smem;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int i0 = blockIdx.y * XQ;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int hd4 = hd >> 2;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const size_t rbase = (size_t)b * T;
  if (a.m_txt[b] == 0) {
    for (int ii = 0; ii < XQ && i0 + ii < T; ++ii) {
      for (int c = threadIdx.x; c < hd; c += blockDim.x) a.o[(rbase + i0 + ii) * a.ldo + h * hd + c] = 0.f;
      if (a.probs)
        for (int j = threadIdx.x; j < T; j += blockDim.x) a.probs[(((size_t)b * H + h) * T + i0 + ii) * T + j] = 0.f;
    }
    return;
  }
  // scores
  for (int p = w; p < XQ * T; p += nw) {
    const int ii = p / T, j = p % T, i = i0 + ii;
    float s = 0.f;
    if (i < T) s = a.scale * warp_dot2(a.q + (rbase + i) * a.ldq + h * hd, a.k + (rbase + j) * a.ldk + h * hd, hd, lane);
    if (lane == 0) s_s[ii * T + j] = s;
  }
  __syncthreads();
  // softmax rows; keep p in smem scaled by the dropout keep mask, save clean p
  for (int ii = w; ii < XQ; ii += nw) {
    const int i = i0 + ii;
    if (i >= T) continue;
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, s_s[ii * T + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(s_s[ii * T + j] - mx);
      s_s[ii * T + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < T; j += 32) {
      const float p = s_s[ii * T + j] / sum;
      const size_t pidx = (((size_t)b * H + h) * T + i) * T + j;
      if (a.probs) a.probs[pidx] = p;
      s_s[ii * T + j] = p * dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
    }
  }
  __syncthreads();
  // O = P~ v
  for (int c4 = threadIdx.x; c4 < hd4; c4 += blockDim.x) {
    float4 acc[XQ];
#pragma unroll
    for (int ii = 0; ii < XQ; ++ii) acc[ii] = f4_zero();
    for (int j = 0; j < T; ++j) {
      const float4 vv = __ldg(reinterpret_cast<const float4*>(a.v + (rbase + j) * a.ldv + h * hd) + c4);
#pragma unroll
      for (int ii = 0; ii < XQ; ++ii) f4_fma(acc[ii], s_s[ii * T + j], vv);
    }
#pragma unroll
    for (int ii = 0; ii < XQ; ++ii)
      if (i0 + ii < T) reinterpret_cast<float4*>(a.o + (rbase + i0 + ii) * a.ldo + h * hd)[c4] = acc[ii];
  }
}

// smem: s_dp [XQ][T] (dP~ then dS) | s_pt [XQ][T] (P~)
__global__ void __launch_bounds__(256) xattn_core_bwd_kernel(const XArgs a) {
  extern __shared__ float smem[];
  const int T = a.T, hd = a.hd, H = a.H;
  float* s_ds = smem;
  float* s_pt = smem + (size_t)XQ * T;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int hd4 = hd >> 2;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const size_t rbase = (size_t)b * T;
  if (a.m_txt[b] == 0) {
    for (int i = 0; i < T; ++i)
      for (int c = threadIdx.x; c < hd; c += blockDim.x) {
        a.dq[(rbase + i) * a.lddq + h * hd + c] = 0.f;
        a.dk[(rbase + i) * a.lddk + h * hd + c] = 0.f;
        a.dv[(rbase + i) * a.lddv + h * hd + c] = 0.f;
      }
    return;
  }
  for (int i0 = 0; i0 < T; i0 += XQ) {
    __syncthreads();
    // dP~[i][j] = dO_i . v_j
    for (int p = w; p < XQ * T; p += nw) {
      const int ii = p / T, j = p % T, i = i0 + ii;
      float s = 0.f;
      if (i < T) s = warp_dot2(a.d_o + (rbase + i) * a.lddo + h * hd, a.v + (rbase + j) * a.ldv + h * hd, hd, lane);
      if (lane == 0) s_ds[ii * T + j] = s;
    }
    __syncthreads();
    for (int ii = w; ii < XQ; ii += nw) {
      const int i = i0 + ii;
      if (i >= T) {
        for (int j = lane; j < T; j += 32) { s_ds[ii * T + j] = 0.f; s_pt[ii * T + j] = 0.f; }
        continue;
      }
      float D = 0.f;
      for (int j = lane; j < T; j += 32) {
        const size_t pidx = (((size_t)b * H + h) * T + i) * T + j;
        const float ks = dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
        const float p = a.probs[pidx];
        const float dp = s_ds[ii * T + j] * ks;
        s_ds[ii * T + j] = dp;
        s_pt[ii * T + j] = p * ks;
        D = fmaf(p, dp, D);
      }
      D = warp_sum(D);
      for (int j = lane; j < T; j += 32) {
        const size_t pidx = (((size_t)b * H + h) * T + i) * T + j;
        const float p = a.probs[pidx];
        s_ds[ii * T + j] = a.scale * p * (s_ds[ii * T + j] - D);  // dS with the q-scale folded in
      }
    }
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < hd4; c4 += blockDim.x) {
      float4 go[XQ], qq[XQ], dq[XQ];
#pragma unroll
      for (int ii = 0; ii < XQ; ++ii) {
        const bool ok = i0 + ii < T;
        go[ii] = ok ? __ldg(reinterpret_cast<const float4*>(a.d_o + (rbase + i0 + ii) * a.lddo + h * hd) + c4) : f4_zero();
        qq[ii] = ok ? __ldg(reinterpret_cast<const float4*>(a.q + (rbase + i0 + ii) * a.ldq + h * hd) + c4) : f4_zero();
        dq[ii] = f4_zero();
      }
      for (int j = 0; j < T; ++j) {
        const float4 kk = __ldg(reinterpret_cast<const float4*>(a.k + (rbase + j) * a.ldk + h * hd) + c4);
        float4 dvj = f4_zero(), dkj = f4_zero();
#pragma unroll
        for (int ii = 0; ii < XQ; ++ii) {
          const float pt = s_pt[ii * T + j], ds = s_ds[ii * T + j];
          f4_fma(dvj, pt, go[ii]);
          f4_fma(dkj, ds, qq[ii]);
          f4_fma(dq[ii], ds, kk);
        }
        float4* pdv = reinterpret_cast<float4*>(a.dv + (rbase + j) * a.lddv + h * hd) + c4;
        float4* pdk = reinterpret_cast<float4*>(a.dk + (rbase + j) * a.lddk + h * hd) + c4;
        if (i0 == 0) { *pdv = dvj; *pdk = dkj; }
        else { float4 x = *pdv; f4_add(x, dvj); *pdv = x; float4 y = *pdk; f4_add(y, dkj); *pdk = y; }
      }
#pragma unroll
      for (int ii = 0; ii < XQ; ++ii)
        if (i0 + ii < T) reinterpret_cast<float4*>(a.dq + (rbase + i0 + ii) * a.lddq + h * hd)[c4] = dq[ii];
    }
  }
}

static int xattn_check(const char* name, int B, int T, int H, int d) {
  (void)B;
  IMMTSF_REQUIRE(H >= 1 && d % H == 0 && ((d / H) & 3) == 0, "%s: head_dim = d/H must be a multiple of 4 (d=%d H=%d)", name, d, H);
  IMMTSF_REQUIRE(T >= 1 && (size_t)2 * XQ * T * sizeof(float) <= 200 * 1024, "%s: T=%d too large for this kernel (max 3200)", name, T);
  return IMMTSF_OK;
}
static inline bool al16(const void* p, int ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

extern "C" int immtsf_xattn_core_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                     const uint8_t* m_txt, int B, int T, int H, int d, uint32_t drop_thr,
                                     uint64_t seed, float* o, int ldo, float* probs, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(q && k && v && m_txt && o, "xattn_core_fwd: null pointer");
  int rc = xattn_check("xattn_core_fwd", B, T, H, d);
  if (rc) return rc;
  IMMTSF_REQUIRE(al16(v, ldv) && al16(o, ldo), "xattn_core_fwd: v/o must be 16B aligned with ld %% 4 == 0");
  if (immtsf_xattn_small_ok(T, H, d) && al16(q, ldq) && al16(k, ldk))
    return immtsf_xattn_small_fwd(q, ldq, k, ldk, v, ldv, m_txt, B, T, H, d, drop_thr, seed, o, ldo, probs, (cudaStream_t)stream);
  XArgs a = {};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.d = d;
  a.hd = d / H; a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.o = o; a.ldo = ldo; a.probs = probs;
  const size_t smem = (size_t)XQ * T * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(xattn_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(B * H, ceil_div(T, XQ));
  xattn_core_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_core_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_xattn_core_bwd(const float* d_o, int lddo, const float* q, int ldq, const float* k, int ldk,
                                     const float* v, int ldv, const float* probs, const uint8_t* m_txt, int B, int T,
                                     int H, int d, uint32_t drop_thr, uint64_t seed, float* dq, int lddq, float* dk,
                                     int lddk, float* dv, int lddv, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(d_o && q && k && v && probs && m_txt && dq && dk && dv, "xattn_core_bwd: null pointer");
  int rc = xattn_check("xattn_core_bwd", B, T, H, d);
  if (rc) return rc;
  IMMTSF_REQUIRE(al16(d_o, lddo) && al16(q, ldq) && al16(k, ldk) && al16(dq, lddq) && al16(dk, lddk) && al16(dv, lddv),
                 "xattn_core_bwd: operands must be 16B aligned with ld %% 4 == 0");
  if (immtsf_xattn_small_ok(T, H, d) && al16(v, ldv))
    return immtsf_xattn_small_bwd(d_o, lddo, q, ldq, k, ldk, v, ldv, probs, m_txt, B, T, H, d, drop_thr, seed, dq, lddq, dk, lddk,
                                  dv, lddv, (cudaStream_t)stream);
  XArgs a = {};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.d = d;
  a.hd = d / H; a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.probs = const_cast<float*>(probs);
  a.d_o = d_o; a.lddo = lddo; a.dq = dq; a.lddq = lddq; a.dk = dk; a.lddk = lddk; a.dv = dv; a.lddv = lddv;
  const size_t smem = (size_t)2 * XQ * T * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(xattn_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  xattn_core_bwd_kernel<<<B * H, 256, smem, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_core_bwd");
  return IMMTSF_OK;
}

// ------------------------------------------------------------------ tail
// one warp per (b,t) row; C arbitrary (lanes stride over C)
__global__ void __launch_bounds__(256) xattn_tail_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ delta_y,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const uint8_t* __restrict__ m_txt, int B, int T, int C, float eps,
                                                             float kappa, uint32_t thr, SeedArg seed_, float* __restrict__ Y_out,
                                                             int32_t* __restrict__ flags) {
  const uint64_t seed = resolve_seed(seed_);
  const int lane = threadIdx.x & 31;
  const int rows = B * T;
  const float inv_keep = inv_keep_from_thr(thr);
  bool bad = false;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += (gridDim.x * blockDim.x) >> 5) {
    const float* dl = delta_y + (size_t)row * C;
    float s = 0.f;
    for (int j = lane; j < C; j += 32) { s += dl[j]; bad |= isnan(dl[j]); }
    const float mu = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int j = lane; j < C; j += 32) v += (dl[j] - mu) * (dl[j] - mu);
    const float rs = 1.f / sqrtf(warp_sum(v) / (float)C + eps);
    const bool has_txt = m_txt[row / T] != 0;
    for (int j = lane; j < C; j += 32) {
      float dd = 0.f;
      if (has_txt)
        dd = ((dl[j] - mu) * rs * gamma[j] + beta[j]) * dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, (uint64_t)row * C + j, thr, inv_keep);
      const float o = (Y[(size_t)row * C + j] + kappa * dd) / (1.f + kappa);
      Y_out[(size_t)row * C + j] = o;
      bad |= isnan(o);
    }
  }
  if (flags != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) flags[IMMTSF_FLAG_OUT] = 1;
}

__global__ void __launch_bounds__(256) xattn_tail_bwd_kernel(const float* __restrict__ dY_out, const float* __restrict__ delta_y,
                                                             const float* __restrict__ gamma, const uint8_t* __restrict__ m_txt,
                                                             int B, int T, int C, float eps, float kappa, uint32_t thr,
                                                             SeedArg seed_, float* __restrict__ d_delta_y,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const uint64_t seed = resolve_seed(seed_);
  const int lane = threadIdx.x & 31;
  const int rows = B * T;
  const float inv_keep = inv_keep_from_thr(thr);
  const float kfac = kappa / (1.f + kappa);
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f};  // C <= 128: lane owns j = lane + 32u
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += (gridDim.x * blockDim.x) >> 5) {
    const float* dl = delta_y + (size_t)row * C;
    const bool has_txt = m_txt[row / T] != 0;
    if (!has_txt) {
      for (int j = lane; j < C; j += 32) d_delta_y[(size_t)row * C + j] = 0.f;
      continue;
    }
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += dl[j];
    const float mu = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int j = lane; j < C; j += 32) v += (dl[j] - mu) * (dl[j] - mu);
    const float rs = 1.f / sqrtf(warp_sum(v) / (float)C + eps);
    float p1 = 0.f, p2 = 0.f;
    float dn_[4], x_[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + u * 32;
      dn_[u] = 0.f; x_[u] = 0.f;
      if (j < C) {
        const float x = (dl[j] - mu) * rs;
        const float dn = dY_out[(size_t)row * C + j] * kfac * dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, (uint64_t)row * C + j, thr, inv_keep);
        dgam[u] = fmaf(dn, x, dgam[u]);
        dbet[u] += dn;
        const float g = dn * gamma[j];
        dn_[u] = g; x_[u] = x;
        p1 += g;
        p2 += g * x;
      }
    }
    const float m1 = warp_sum(p1) / (float)C, m2 = warp_sum(p2) / (float)C;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = lane + u * 32;
      if (j < C) d_delta_y[(size_t)row * C + j] = rs * (dn_[u] - m1 - x_[u] * m2);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = lane + u * 32;
    if (j < C) { atomicAdd(dgamma + j, dgam[u]); atomicAdd(dbeta + j, dbet[u]); }
  }
}

extern "C" int immtsf_xattn_tail_fwd(const float* Y, const float* delta_y, const float* gamma, const float* beta,
                                     const uint8_t* m_txt, int B, int T, int C, float eps, float kappa,
                                     uint32_t drop_thr, uint64_t seed, float* Y_out, int32_t* flags, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(Y && delta_y && gamma && beta && m_txt && Y_out && C >= 1, "xattn_tail_fwd: bad args");
  int grid = ceil_div(B * T, 8);
  if (grid > 148 * 8) grid = 148 * 8;
  xattn_tail_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, delta_y, gamma, beta, m_txt, B, T, C, eps, kappa, drop_thr, make_seed(seed), Y_out, flags);
  IMMTSF_CHECK_LAUNCH("xattn_tail_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_xattn_tail_bwd(const float* dY_out, const float* delta_y, const float* gamma, const uint8_t* m_txt,
                                     int B, int T, int C, float eps, float kappa, uint32_t drop_thr, uint64_t seed,
                                     float* d_delta_y, float* dgamma, float* dbeta, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dY_out && delta_y && gamma && m_txt && d_delta_y && dgamma && dbeta && C >= 1 && C <= 128, "xattn_tail_bwd: C must be in [1,128]");
  int grid = ceil_div(B * T, 8);
  if (grid > 148 * 4) grid = 148 * 4;
  xattn_tail_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dY_out, delta_y, gamma, m_txt, B, T, C, eps, kappa, drop_thr, make_seed(seed), d_delta_y, dgamma, dbeta);
  IMMTSF_CHECK_LAUNCH("xattn_tail_bwd");
  return IMMTSF_OK;
}


// ------------------------------------------------------------------ large-T path: softmax rows around batched GEMMs
// For T > 32 the T x T contractions run on the tensor cores (immtsf_gemm_batched: S = Q K^T, O = P~ V, and the
// four backward products); what is left are these two row kernels over the [B, H, T, Tp] score buffers
// (Tp = T rounded up to 4 floats; dropout indices ignore the padding: same masks as the small-T kernels).
//   fwd: P = softmax(scale * S) written in place (saved for backward), P~ = P * keep / (1-p) -> Pt
//   bwd: dP~ = dO V^T comes in through dS; dS <- scale * P * (dP~ * ks - D), D = sum_j P dP~ ks; Pt <- P * ks
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(float* __restrict__ S, float* __restrict__ Pt,
                                                               const uint8_t* __restrict__ m_txt, int B, int H, int T, int Tp,
                                                               float scale, uint32_t thr, SeedArg seed_) {
  const uint64_t seed = resolve_seed(seed_);
  const float inv_keep = inv_keep_from_thr(thr);
  const int lane = threadIdx.x & 31;
  const long rows = (long)B * H * T;
  for (long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += ((long)gridDim.x * blockDim.x) >> 5) {
    const int b = (int)(row / ((long)H * T));
    float* srow = S + row * Tp;
    float* prow = Pt + row * Tp;
    if (m_txt[b] == 0) {
      for (int j = lane; j < Tp; j += 32) { srow[j] = 0.f; prow[j] = 0.f; }
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, scale * srow[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) sum += expf(scale * srow[j] - mx);
    sum = warp_sum(sum);
    for (int j = lane; j < Tp; j += 32) {
      float p = 0.f, pt = 0.f;
      if (j < T) {
        p = expf(scale * srow[j] - mx) / sum;
        pt = p * dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, (uint64_t)row * T + j, thr, inv_keep);
      }
      srow[j] = p;
      prow[j] = pt;
    }
  }
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(float* __restrict__ dS, const float* __restrict__ P,
                                                               float* __restrict__ Pt, const uint8_t* __restrict__ m_txt, int B,
                                                               int H, int T, int Tp, float scale, uint32_t thr, SeedArg seed_) {
  const uint64_t seed = resolve_seed(seed_);
  const float inv_keep = inv_keep_from_thr(thr);
  const int lane = threadIdx.x & 31;
  const long rows = (long)B * H * T;
  for (long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += ((long)gridDim.x * blockDim.x) >> 5) {
    const int b = (int)(row / ((long)H * T));
    float* drow = dS + row * Tp;
    float* trow = Pt + row * Tp;
    const float* prow = P + row * Tp;
    if (m_txt[b] == 0) {
      for (int j = lane; j < Tp; j += 32) { drow[j] = 0.f; trow[j] = 0.f; }
      continue;
    }
    float D = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float ks = dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, (uint64_t)row * T + j, thr, inv_keep);
      D = fmaf(prow[j], drow[j] * ks, D);
    }
    D = warp_sum(D);
    for (int j = lane; j < Tp; j += 32) {
      float ds = 0.f, pt = 0.f;
      if (j < T) {
        const float ks = dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, (uint64_t)row * T + j, thr, inv_keep);
        const float p = prow[j];
        ds = scale * p * (drow[j] * ks - D);
        pt = p * ks;
      }
      drow[j] = ds;
      trow[j] = pt;
    }
  }
}

extern "C" int immtsf_softmax_rows_fwd(float* S, float* Pt, const uint8_t* m_txt, int B, int H, int T, int Tp, float scale,
                                       uint32_t drop_thr, uint64_t seed, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(S && Pt && m_txt && H >= 1 && Tp >= T, "softmax_rows_fwd: bad args");
  const long rows = (long)B * H * T;
  int grid = (int)((rows + 7) / 8 < 148 * 16 ? (rows + 7) / 8 : 148 * 16);
  softmax_rows_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, Pt, m_txt, B, H, T, Tp, scale, drop_thr, make_seed(seed));
  IMMTSF_CHECK_LAUNCH("softmax_rows_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_softmax_rows_bwd(float* dS, const float* P, float* Pt, const uint8_t* m_txt, int B, int H, int T, int Tp,
                                       float scale, uint32_t drop_thr, uint64_t seed, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dS && P && Pt && m_txt && H >= 1 && Tp >= T, "softmax_rows_bwd: bad args");
  const long rows = (long)B * H * T;
  int grid = (int)((rows + 7) / 8 < 148 * 16 ? (rows + 7) / 8 : 148 * 16);
  softmax_rows_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dS, P, Pt, m_txt, B, H, T, Tp, scale, drop_thr, make_seed(seed));
  IMMTSF_CHECK_LAUNCH("softmax_rows_bwd");
  return IMMTSF_OK;
}
