// K3 pieces of TTF_T2V_XAttn: Time2Vec features and the single-learned-query
// attention over each ragged note segment.
//
// The reference (fusions/TTF_T2V_XAttn.py:143-166) broadcasts ONE learned query
// to every (sample, query time), copies K/V T_f times and re-projects them
// inside nn.MultiheadAttention.  Here K/V projections run once per note
// (immtsf_gemm over sumN rows) and this kernel does, per sample and head:
//   s_n = q_h . K_{n,h};  p = softmax_n(s);  o_{t,h} = sum_n drop_t(p_n) V_{n,h}
// In eval / dropout 0 the output does not depend on t and is produced once
// per sample (R = B rows); with attention dropout every (t,h,n) weight gets
// its own Philox keep bit and R = B*T rows.
#include "tile32.cuh"
#include "rowwarp.cuh"
#include "../../include/immtsf.h"

// ------------------------------------------------------------------ Time2Vec
// out[n, 0] = w_lin*tau + b_lin ; out[n, k] = sin(w_per[k-1]*tau + b_per[k-1])
__global__ void time2vec_fwd_kernel(const float* __restrict__ tau, const float* __restrict__ w_lin,
                                    const float* __restrict__ b_lin, const float* __restrict__ w_per,
                                    const float* __restrict__ b_per, int d_tau, float* __restrict__ out, int ld,
                                    float* __restrict__ out_lo, int ld_lo, const int32_t* __restrict__ m_dev, int M_alloc) {
  const int m = ragged_rows(M_alloc, m_dev);
  int end = (m + 127) / 128 * 128;
  if (end > M_alloc) end = M_alloc;
  for (int n = blockIdx.x; n < end; n += gridDim.x) {
    float* o = out + (size_t)n * ld;
    float* l = out_lo != nullptr ? out_lo + (size_t)n * ld_lo : nullptr;
    if (n < m) {
      const float t = tau[n];
      for (int k = threadIdx.x; k < d_tau; k += blockDim.x) {
        const float v = (k == 0) ? fmaf(w_lin[0], t, b_lin[0]) : sinf(fmaf(w_per[k - 1], t, b_per[k - 1]));
        o[k] = v;
        if (l != nullptr) l[k] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);  // tcgen05 operand split
      }
    } else {
      for (int k = threadIdx.x; k < d_tau; k += blockDim.x) {
        o[k] = 0.f;
        if (l != nullptr) l[k] = 0.f;
      }
    }
  }
}

// grid.x over feature tiles of 32, grid.y over row chunks; 32x8 threads
__global__ void time2vec_bwd_kernel(const float* __restrict__ dphi, int ld, const float* __restrict__ tau,
                                    const float* __restrict__ w_per, const float* __restrict__ b_per, int d_tau,
                                    float* __restrict__ dw_lin, float* __restrict__ db_lin, float* __restrict__ dw_per,
                                    float* __restrict__ db_per, const int32_t* __restrict__ m_dev, int M_alloc) {
  __shared__ float rw[8][33], rb[8][33];
  const int m = ragged_rows(M_alloc, m_dev);
  const int k = blockIdx.x * 32 + threadIdx.x;
  float sw = 0.f, sb = 0.f;
  if (k < d_tau) {
    const float wk = k == 0 ? 0.f : w_per[k - 1], bk = k == 0 ? 0.f : b_per[k - 1];
    for (int n = blockIdx.y * 8 + threadIdx.y; n < m; n += gridDim.y * 8) {
      const float t = tau[n];
      const float g = dphi[(size_t)n * ld + k];
      const float dpre = k == 0 ? g : g * cosf(fmaf(wk, t, bk));
      sw = fmaf(dpre, t, sw);
      sb += dpre;
    }
  }
  rw[threadIdx.y][threadIdx.x] = sw;
  rb[threadIdx.y][threadIdx.x] = sb;
  __syncthreads();
  if (threadIdx.y == 0 && k < d_tau) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += rw[i][threadIdx.x]; c += rb[i][threadIdx.x]; }
    if (k == 0) { atomicAdd(dw_lin, a); atomicAdd(db_lin, c); }
    else { atomicAdd(dw_per + k - 1, a); atomicAdd(db_per + k - 1, c); }
  }
}

extern "C" int immtsf_time2vec_fwd(const float* tau_flat, const float* w_lin, const float* b_lin, const float* w_per,
                                   const float* b_per, int d_tau, float* out, int ld, float* out_lo, int ld_lo,
                                   const int32_t* m_dev, int M_alloc, void* stream) {
  if (M_alloc == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(tau_flat && w_lin && b_lin && w_per && b_per && out && m_dev, "time2vec_fwd: null pointer");
  IMMTSF_REQUIRE(d_tau > 1 && ld >= d_tau, "time2vec_fwd: d_tau must be > 1 (TTF_T2V_XAttn.py:14) and ld >= d_tau");
  IMMTSF_REQUIRE(out_lo == nullptr || ld_lo >= d_tau, "time2vec_fwd: ld_lo < d_tau");
  int grid = M_alloc < 148 * 8 ? M_alloc : 148 * 8;
  time2vec_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(tau_flat, w_lin, b_lin, w_per, b_per, d_tau, out, ld, out_lo, ld_lo,
                                                              m_dev, M_alloc);
  IMMTSF_CHECK_LAUNCH("time2vec_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_time2vec_bwd(const float* dphi, int ld, const float* tau_flat, const float* w_per,
                                   const float* b_per, int d_tau, float* dw_lin, float* db_lin, float* dw_per,
                                   float* db_per, const int32_t* m_dev, int M_alloc, void* stream) {
  if (M_alloc == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dphi && tau_flat && w_per && b_per && dw_lin && db_lin && dw_per && db_per && m_dev, "time2vec_bwd: null pointer");
  int gy = ceil_div(M_alloc, 64);
  if (gy > 64) gy = 64;
  dim3 grid(ceil_div(d_tau, 32), gy);
  time2vec_bwd_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(dphi, ld, tau_flat, w_per, b_per, d_tau, dw_lin, db_lin,
                                                                       dw_per, db_per, m_dev, M_alloc);
  IMMTSF_CHECK_LAUNCH("time2vec_bwd");
  return IMMTSF_OK;
}

// ------------------------------------------------------------- segment attention
struct SegArgs {
  const float* q; const float* KVp; const int32_t* offsets;
  int B, T, H, d, hd, N_max, per_query, TT; uint32_t thr; SeedArg seed;
  float* attn_cat; float* probs;
  const float* dO; float* dKVp; float* dq_partial;
};

// warp-cooperative dot of two length-n vectors (coalesced scalar loads)
__device__ __forceinline__ float warp_dot(const float* __restrict__ a, const float* __restrict__ b, int n, int lane) {
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s = fmaf(a[j], b[j], s);
  return warp_sum(s);
}

// smem: s_p [H][N_max] | s_pt [TT][H][N_max]
__global__ void __launch_bounds__(256) segattn_fwd_kernel(const SegArgs a) {
  extern __shared__ float smem[];
  const int H = a.H, d = a.d, hd = a.hd, NM = a.N_max, TT = a.TT;
  float* s_p = smem;
  float* s_pt = smem + (size_t)H * NM;
  const int b = blockIdx.x;
  const int nb = a.offsets[b], ne = a.offsets[b + 1], nn = ne - nb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ld = 2 * d, d4 = d >> 2;
  const int Teff = a.per_query ? a.T : 1;
  const float inv_keep = inv_keep_from_thr(a.thr);
  if (nn == 0) {  // no notes: attention output is defined as 0 (reference zeroes it, :171-173)
    for (int t = 0; t < Teff; ++t)
      for (int c = threadIdx.x; c < d; c += blockDim.x) a.attn_cat[((size_t)b * Teff + t) * d + c] = 0.f;
    return;
  }
  // 1) scores
  for (int i = w; i < nn * H; i += nw) {
    const int n = i / H, h = i % H;
    const float s = warp_dot(a.q + h * hd, a.KVp + (size_t)(nb + n) * ld + h * hd, hd, lane);
    if (lane == 0) s_p[h * NM + n] = s;
  }
  __syncthreads();
  // 2) softmax over the segment, one warp per head
  for (int h = w; h < H; h += nw) {
    float mx = -INFINITY;
    for (int n = lane; n < nn; n += 32) mx = fmaxf(mx, s_p[h * NM + n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < nn; n += 32) {
      const float e = expf(s_p[h * NM + n] - mx);
      s_p[h * NM + n] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int n = lane; n < nn; n += 32) {
      const float p = s_p[h * NM + n] / sum;
      s_p[h * NM + n] = p;
      if (a.probs) a.probs[(size_t)(nb + n) * H + h] = p;
    }
  }
  __syncthreads();
  // 3) weighted sums, TT query times at a time
  for (int t0 = 0; t0 < Teff; t0 += TT) {
    __syncthreads();
    for (int i = threadIdx.x; i < TT * H * nn; i += blockDim.x) {
      const int tt = i / (H * nn), r = i % (H * nn), h = r / nn, n = r % nn;
      const int t = t0 + tt;
      float p = 0.f;
      if (t < Teff) {
        p = s_p[h * NM + n];
        if (a.per_query)
          p *= dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_TTF_ATTN, (((uint64_t)b * a.T + t) * H + h) * NM + n, a.thr, inv_keep);
      }
      s_pt[((size_t)tt * H + h) * NM + n] = p;
    }
    __syncthreads();
    for (int col4 = threadIdx.x; col4 < d4; col4 += blockDim.x) {
      const int h = (col4 * 4) / hd;
      float4 acc[4];
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) acc[tt] = f4_zero();
      for (int n = 0; n < nn; ++n) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.KVp + (size_t)(nb + n) * ld + d) + col4);
#pragma unroll
        for (int tt = 0; tt < 4; ++tt)
          if (tt < TT) f4_fma(acc[tt], s_pt[((size_t)tt * H + h) * NM + n], v);
      }
#pragma unroll
      for (int tt = 0; tt < 4; ++tt)
        if (tt < TT && t0 + tt < Teff)
          reinterpret_cast<float4*>(a.attn_cat + ((size_t)b * Teff + t0 + tt) * d)[col4] = acc[tt];
    }
  }
}

// ------------------------------------------------------------- one head, train mode: attention + residual + LayerNorm + dropout
// TTF_T2V_XAttn.py:143-179 for ONE head with attention dropout (every (sample, query) row distinct), in one launch:
// softmax over the sample's notes, dropout on the weights, pooling of the value rows, + out-projection bias + learned query,
// LayerNorm, dropout.  Round 1 ran this as segattn_fwd (one CTA per sample, four query rows per pass: six passes over the
// value rows, 30 us at cfg2) followed by ln_fwd (17 us): 256 CTAs on 148 SMs, latency-bound.  Here a CTA owns EIGHT query
// rows of a sample (grid = T/8 x B = 768 CTAs at cfg2): the scores are recomputed per tile (N_i dot products), the value
// rows are staged in shared memory once per CTA, and warp w owns query row w -- its pooled row never leaves registers
// before the LayerNorm statistics (warp shuffles), so the [B*T, d] attention output is written once (saved for backward)
// and never read back.  Same summation order and lane ownership as the two kernels it replaces: bit-identical results.
constexpr int SL_TQ = 12;   // query rows per CTA (T 24: two tiles per sample = 512 CTAs, one wave at 4 CTAs per SM)
constexpr int SL_RW = 3;    // query rows per warp: every staged value chunk read from shared memory feeds all of them (ncu on the
                            // one-row-per-warp version: short-scoreboard 5.1 per issue -- bound by shared-memory loads)
constexpr int SL_WARPS = SL_TQ / SL_RW;
constexpr int SL_NS = 16;   // value rows staged per pass

struct SegLnArgs {
  const float* q; const float* KVp; const int32_t* offsets;
  int B, T, d, N_max; uint32_t thr; SeedArg seed; float eps;
  const float* xbias; const float* res; const float* gamma; const float* beta;
  float* attn_cat; float* probs; float* y; float* mean; float* rstd;
};

// smem: s_v [SL_NS][d] | s_p [N_max] | s_pt [SL_TQ][N_max]
template <int NC>
__global__ void __launch_bounds__(SL_WARPS * 32) segattn_ln_fwd_kernel(const SegLnArgs a) {
  extern __shared__ __align__(16) float sl_smem[];
  const int d = a.d, d8 = d >> 3, d4 = d >> 2, NM = a.N_max, ld = 2 * d;
  float* s_v = sl_smem;
  float* s_p = s_v + (size_t)SL_NS * d;
  float* s_pt = s_p + NM;
  const int b = blockIdx.y, t0 = blockIdx.x * SL_TQ;
  const int nb = a.offsets[b], nn = a.offsets[b + 1] - nb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)d;
  const uint64_t seed = resolve_seed(a.seed);
  float acc[SL_RW][NC][8];
#pragma unroll
  for (int q = 0; q < SL_RW; ++q)
#pragma unroll
    for (int i = 0; i < NC; ++i) zero8(acc[q][i]);
  if (nn > 0) {
    // the first SL_NS value rows start their way into shared memory now (cp.async) and land while the scores, the softmax and
    // the dropout weights are computed
    {
      const int cnt0 = min(SL_NS, nn);
      for (int i = threadIdx.x; i < cnt0 * d4; i += blockDim.x) {
        const int r = i / d4, c = i % d4;
        const unsigned int dst = (unsigned int)__cvta_generic_to_shared(reinterpret_cast<float4*>(s_v) + (size_t)r * d4 + c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const float4*>(a.KVp + (size_t)(nb + r) * ld + d) + c) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {
      float4 qv[(NC * 2)];  // this lane's part of the query vector: float4 chunks lane, lane + 32, ...
#pragma unroll
      for (int j = 0; j < NC * 2; ++j) qv[j] = lane + 32 * j < d4 ? __ldg(reinterpret_cast<const float4*>(a.q) + lane + 32 * j) : f4_zero();
      for (int n = w; n < nn; n += SL_WARPS) {
        const float4* kp = reinterpret_cast<const float4*>(a.KVp + (size_t)(nb + n) * ld);
        float sc = 0.f;
#pragma unroll
        for (int j = 0; j < NC * 2; ++j) {
          if (lane + 32 * j < d4) {
            const float4 kv = __ldg(kp + lane + 32 * j);
            sc = fmaf(qv[j].x, kv.x, fmaf(qv[j].y, kv.y, fmaf(qv[j].z, kv.z, fmaf(qv[j].w, kv.w, sc))));
          }
        }
        sc = warp_sum(sc);
        if (lane == 0) s_p[n] = sc;
      }
    }
    __syncthreads();
    if (w == 0) {
      float mx = -INFINITY;
      for (int n = lane; n < nn; n += 32) mx = fmaxf(mx, s_p[n]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int n = lane; n < nn; n += 32) {
        const float e = expf(s_p[n] - mx);
        s_p[n] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      for (int n = lane; n < nn; n += 32) {
        const float p = s_p[n] / sum;
        s_p[n] = p;
        if (a.probs && blockIdx.x == 0) a.probs[nb + n] = p;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SL_TQ * nn; i += blockDim.x) {
      const int tt = i / nn, n = i % nn;
      float p = 0.f;
      if (t0 + tt < a.T)
        p = s_p[n] * dropout_scale(seed, IMMTSF_SITE_TTF_ATTN, ((uint64_t)b * a.T + t0 + tt) * NM + n, a.thr, inv_keep);
      s_pt[tt * NM + n] = p;
    }
    // pooling: value rows through shared memory, SL_NS at a time; warp w accumulates query rows t0 + SL_RW w ..
    for (int n0 = 0; n0 < nn; n0 += SL_NS) {
      const int cnt = min(SL_NS, nn - n0);
      __syncthreads();  // (also orders the s_pt writes above before their first use)
      if (n0 > 0) {
        for (int i = threadIdx.x; i < cnt * d4; i += blockDim.x) {
          const int r = i / d4, c = i % d4;
          reinterpret_cast<float4*>(s_v)[(size_t)r * d4 + c] = __ldg(reinterpret_cast<const float4*>(a.KVp + (size_t)(nb + n0 + r) * ld + d) + c);
        }
      } else {
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      __syncthreads();
      for (int r = 0; r < cnt; ++r) {
        float p[SL_RW];
#pragma unroll
        for (int q = 0; q < SL_RW; ++q) p[q] = s_pt[(w * SL_RW + q) * NM + n0 + r];  // (0 for rows past T)
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int k = lane + 32 * i;
          if (k < d8) {
            const float4 v0 = reinterpret_cast<const float4*>(s_v + (size_t)r * d)[2 * k];
            const float4 v1 = reinterpret_cast<const float4*>(s_v + (size_t)r * d)[2 * k + 1];
#pragma unroll
            for (int q = 0; q < SL_RW; ++q) {
              acc[q][i][0] = fmaf(p[q], v0.x, acc[q][i][0]); acc[q][i][1] = fmaf(p[q], v0.y, acc[q][i][1]);
              acc[q][i][2] = fmaf(p[q], v0.z, acc[q][i][2]); acc[q][i][3] = fmaf(p[q], v0.w, acc[q][i][3]);
              acc[q][i][4] = fmaf(p[q], v1.x, acc[q][i][4]); acc[q][i][5] = fmaf(p[q], v1.y, acc[q][i][5]);
              acc[q][i][6] = fmaf(p[q], v1.z, acc[q][i][6]); acc[q][i][7] = fmaf(p[q], v1.w, acc[q][i][7]);
            }
          }
        }
      }
    }
  }
  // the pooled rows (saved: LayerNorm backward needs them), then + bias + learned query, LayerNorm, dropout
#pragma unroll
  for (int q = 0; q < SL_RW; ++q) {
    const int t = t0 + w * SL_RW + q;
    if (t >= a.T) break;
    const size_t rowi = (size_t)b * a.T + t;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        store8(a.attn_cat + rowi * d, k, acc[q][i]);
        if (nn > 0 && a.xbias) { float tb[8]; load8(a.xbias, k, tb);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[q][i][e] += tb[e]; }
        if (a.res) { float tr[8]; load8(a.res, k, tr);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[q][i][e] += tr[e]; }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += acc[q][i][e];
      }
    }
    const float mu = warp_sum(s) * inv_d;
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (lane + 32 * i < d8)
#pragma unroll
        for (int e = 0; e < 8; ++e) qv = fmaf(acc[q][i][e] - mu, acc[q][i][e] - mu, qv);
    const float rs = 1.f / sqrtf(warp_sum(qv) * inv_d + a.eps);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float g[8], be[8], ks[8], yv[8];
        load8(a.gamma, k, g);
        load8(a.beta, k, be);
        dropout_scale8(seed, IMMTSF_SITE_TTF_DROPOUT, (uint64_t)rowi * d8 + k, a.thr, inv_keep, ks);
#pragma unroll
        for (int e = 0; e < 8; ++e) yv[e] = ((acc[q][i][e] - mu) * rs * g[e] + be[e]) * ks[e];
        store8(a.y + rowi * d, k, yv);
      }
    }
    if (lane == 0) {
      if (a.mean) a.mean[rowi] = mu;
      if (a.rstd) a.rstd[rowi] = rs;
    }
  }
}

extern "C" int immtsf_segattn_ln_ok(int d, int N_max) {
  return rowwarp_nc(d) > 0 && ((size_t)SL_NS * d + (size_t)(SL_TQ + 1) * N_max) * sizeof(float) <= 160 * 1024;
}

extern "C" int immtsf_segattn_ln_fwd(const float* q, const float* KVp, const int32_t* offsets, int B, int T, int d, int N_max,
                                     uint32_t drop_thr, uint64_t seed, const float* xbias, const float* res, const float* gamma,
                                     const float* beta, float eps, float* attn_cat, float* probs, float* y, float* mean,
                                     float* rstd, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(q && KVp && offsets && gamma && beta && attn_cat && y, "segattn_ln_fwd: null pointer");
  IMMTSF_REQUIRE(immtsf_segattn_ln_ok(d, N_max) && N_max >= 1, "segattn_ln_fwd: needs d %% 8 == 0, d <= 1024 (d=%d N_max=%d)", d, N_max);
  IMMTSF_REQUIRE(B <= 65535, "segattn_ln_fwd: B > 65535");
  for (const void* p : {(const void*)q, (const void*)KVp, (const void*)xbias, (const void*)res, (const void*)gamma, (const void*)beta,
                        (const void*)attn_cat, (const void*)y})
    IMMTSF_REQUIRE(((uintptr_t)p & 15) == 0, "segattn_ln_fwd: buffers must be 16B aligned");
  SegLnArgs a = {};
  a.q = q; a.KVp = KVp; a.offsets = offsets; a.B = B; a.T = T; a.d = d; a.N_max = N_max; a.thr = drop_thr; a.seed = make_seed(seed);
  a.eps = eps; a.xbias = xbias; a.res = res; a.gamma = gamma; a.beta = beta; a.attn_cat = attn_cat; a.probs = probs; a.y = y;
  a.mean = mean; a.rstd = rstd;
  const size_t smem = ((size_t)SL_NS * d + (size_t)(SL_TQ + 1) * N_max) * sizeof(float);
  const dim3 grid(ceil_div(T, SL_TQ), B);
  const int nc = rowwarp_nc(d);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(segattn_ln_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(segattn_ln_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(segattn_ln_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(segattn_ln_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (nc == 1) segattn_ln_fwd_kernel<1><<<grid, SL_WARPS * 32, smem, st>>>(a);
  else if (nc == 2) segattn_ln_fwd_kernel<2><<<grid, SL_WARPS * 32, smem, st>>>(a);
  else if (nc == 3) segattn_ln_fwd_kernel<3><<<grid, SL_WARPS * 32, smem, st>>>(a);
  else segattn_ln_fwd_kernel<4><<<grid, SL_WARPS * 32, smem, st>>>(a);
  IMMTSF_CHECK_LAUNCH("segattn_ln_fwd");
  return IMMTSF_OK;
}

// Backward.  smem: tile32 staging | s_out [32*32] | s_p [H][NM] | s_ds [H][NM] | s_dp [TT][H][NM]
//  a) dp~[t][h][n] = dO[t, head h] . V[n, head h]: 32 x 32 register-tiled X Y^T (tile32.cuh), dO and V staged
//     through shared memory once per tile instead of one warp-shuffle dot product per (t, n) pair;
//  b) softmax backward, accumulated over the query rows of the tile;
//  c) dV[n] = sum_t p~[t][n] dO[t]: register accumulators over 8 notes, dO rows streamed 4 at a time -- a
//     single pass (no read-modify-write of dKVp) whenever all T query rows fit one tile (T <= TT);
//  d) dK[n] = ds[n] q ; dq_b = sum_n ds[n] K[n].
__global__ void __launch_bounds__(256) segattn_bwd_kernel(const SegArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, d = a.d, hd = a.hd, NM = a.N_max, TT = a.TT;
  float* s_x = smem;
  float* s_y = s_x + XS_T * XS_LD;
  float* s_part = s_y + XS_T * XS_LD;
  float* s_out = s_part + 8 * XS_T * XS_T;
  float* s_p = s_out + XS_T * XS_T;
  float* s_ds = s_p + (size_t)H * NM;
  float* s_dp = s_ds + (size_t)H * NM;
  const int b = blockIdx.x;
  const int nb = a.offsets[b], ne = a.offsets[b + 1], nn = ne - nb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ld = 2 * d, d4 = d >> 2;
  const int Teff = a.per_query ? a.T : 1;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  if (nn == 0) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) a.dq_partial[(size_t)b * d + c] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < H * nn; i += blockDim.x) {
    const int h = i / nn, n = i % nn;
    s_p[h * NM + n] = a.probs[(size_t)(nb + n) * H + h];
    s_ds[h * NM + n] = 0.f;
  }
  for (int t0 = 0; t0 < Teff; t0 += TT) {
    const int tcnt = min(TT, Teff - t0);
    // a) dp~ for this tile of query rows
    for (int h = 0; h < H; ++h)
      for (int n0 = 0; n0 < nn; n0 += XS_T) {
        const int ncnt = min(XS_T, nn - n0);
        tile_xyt(a.dO + ((size_t)b * Teff + t0) * d + h * hd, d, a.KVp + (size_t)(nb + n0) * ld + d + h * hd, ld, tcnt, ncnt, hd,
                 s_x, s_y, s_part, s_out);
        for (int i = threadIdx.x; i < tcnt * ncnt; i += blockDim.x) {
          const int tt = i / ncnt, j = i % ncnt;
          s_dp[((size_t)tt * H + h) * NM + n0 + j] = s_out[tt * XS_T + j];
        }
      }
    __syncthreads();
    // b) softmax backward (accumulated over t) ; s_dp <- dropped probabilities p~
    for (int i = w; i < tcnt * H; i += nw) {
      const int tt = i / H, h = i % H, t = t0 + tt;
      float* dprow = s_dp + ((size_t)tt * H + h) * NM;
      float D = 0.f;
      for (int n = lane; n < nn; n += 32) {
        float ks = 1.f;
        if (a.per_query) ks = dropout_scale(seed, IMMTSF_SITE_TTF_ATTN, (((uint64_t)b * a.T + t) * H + h) * NM + n, a.thr, inv_keep);
        const float dp = dprow[n] * ks;
        dprow[n] = dp;
        D = fmaf(s_p[h * NM + n], dp, D);
      }
      D = warp_sum(D);
      for (int n = lane; n < nn; n += 32) {
        float ks = 1.f;
        if (a.per_query) ks = dropout_scale(seed, IMMTSF_SITE_TTF_ATTN, (((uint64_t)b * a.T + t) * H + h) * NM + n, a.thr, inv_keep);
        const float p = s_p[h * NM + n];
        atomicAdd(&s_ds[h * NM + n], p * (dprow[n] - D));  // rows t of one head are spread over warps
        dprow[n] = p * ks;
      }
    }
    __syncthreads();
    // c) dV[n] (+)= sum_tt p~[tt][h][n] dO[t]
    for (int col4 = threadIdx.x; col4 < d4; col4 += blockDim.x) {
      const int h = (col4 * 4) / hd;
      const float4* gop = reinterpret_cast<const float4*>(a.dO + ((size_t)b * Teff + t0) * d) + col4;
      for (int n0 = 0; n0 < nn; n0 += 8) {
        float4 acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = f4_zero();
        for (int tq = 0; tq < tcnt; tq += 4) {
          float4 go[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) go[v] = tq + v < tcnt ? __ldg(gop + (size_t)(tq + v) * d4) : f4_zero();
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (tq + v < tcnt) {
              const float* pr = s_dp + ((size_t)(tq + v) * H + h) * NM + n0;
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (n0 + u < nn) f4_fma(acc[u], pr[u], go[v]);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (n0 + u < nn) {
            float4* dst = reinterpret_cast<float4*>(a.dKVp + (size_t)(nb + n0 + u) * ld + d) + col4;
            if (t0 == 0) *dst = acc[u];
            else { float4 o = *dst; f4_add(o, acc[u]); *dst = o; }
          }
        }
      }
    }
    __syncthreads();
  }
  // d) dK[n] = ds[h][n] q ; dq_b = sum_n ds[h][n] K[n]
  for (int col4 = threadIdx.x; col4 < d4; col4 += blockDim.x) {
    const int h = (col4 * 4) / hd;
    const float4 qv = __ldg(reinterpret_cast<const float4*>(a.q) + col4);
    float4 dq = f4_zero();
    for (int n0 = 0; n0 < nn; n0 += 4) {
      float4 kv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        kv[u] = n0 + u < nn ? __ldg(reinterpret_cast<const float4*>(a.KVp + (size_t)(nb + n0 + u) * ld) + col4) : f4_zero();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (n0 + u < nn) {
          const float ds = s_ds[h * NM + n0 + u];
          f4_fma(dq, ds, kv[u]);
          float4 dk;
          dk.x = ds * qv.x; dk.y = ds * qv.y; dk.z = ds * qv.z; dk.w = ds * qv.w;
          reinterpret_cast<float4*>(a.dKVp + (size_t)(nb + n0 + u) * ld)[col4] = dk;
        }
      }
    }
    reinterpret_cast<float4*>(a.dq_partial + (size_t)b * d)[col4] = dq;
  }
}

static int seg_setup(SegArgs& a, int extra_planes, size_t& smem) {
  // planes of [H][N_max] floats: extra_planes fixed + TT for the per-t tile
  const size_t plane = (size_t)a.H * a.N_max * sizeof(float);
  int TT = 4;
  while (TT > 1 && (extra_planes + TT) * plane > 200 * 1024) TT >>= 1;
  if ((extra_planes + TT) * plane > 200 * 1024) return -1;
  if (!a.per_query) TT = 1;
  a.TT = TT;
  smem = (extra_planes + TT) * plane;
  return 0;
}

extern "C" int immtsf_segattn_fwd(const float* q, const float* KVp, const int32_t* offsets, int B, int T, int H, int d,
                                  int N_max, int per_query, uint32_t drop_thr, uint64_t seed, float* attn_cat,
                                  float* probs, void* stream) {
  if (B == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(q && KVp && offsets && attn_cat, "segattn_fwd: null pointer");
  IMMTSF_REQUIRE(H >= 1 && d % H == 0 && ((d / H) & 3) == 0, "segattn_fwd: head_dim = d/H must be a multiple of 4 (d=%d H=%d)", d, H);
  IMMTSF_REQUIRE(((uintptr_t)KVp & 15) == 0 && ((uintptr_t)attn_cat & 15) == 0, "segattn_fwd: buffers must be 16B aligned");
  IMMTSF_REQUIRE(N_max >= 1 && T >= 1, "segattn_fwd: N_max and T must be >= 1");
  SegArgs a = {};
  a.q = q; a.KVp = KVp; a.offsets = offsets; a.B = B; a.T = T; a.H = H; a.d = d; a.hd = d / H; a.N_max = N_max;
  a.per_query = per_query ? 1 : 0; a.thr = per_query ? drop_thr : 0u; a.seed = make_seed(seed); a.attn_cat = attn_cat; a.probs = probs;
  size_t smem;
  if (seg_setup(a, 1, smem)) { immtsf_set_error("segattn_fwd: H*N_max=%d too large for shared memory", H * N_max); return IMMTSF_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024) cudaFuncSetAttribute(segattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  segattn_fwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("segattn_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_segattn_bwd(const float* d_attn_cat, const float* q, const float* KVp, const float* probs,
                                  const int32_t* offsets, int B, int T, int H, int d, int N_max, int per_query,
                                  uint32_t drop_thr, uint64_t seed, float* dKVp, float* dq_partial, void* stream) {
  if (B == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(d_attn_cat && q && KVp && probs && offsets && dKVp && dq_partial, "segattn_bwd: null pointer");
  IMMTSF_REQUIRE(H >= 1 && d % H == 0 && ((d / H) & 3) == 0, "segattn_bwd: head_dim = d/H must be a multiple of 4 (d=%d H=%d)", d, H);
  IMMTSF_REQUIRE(((uintptr_t)KVp & 15) == 0 && ((uintptr_t)dKVp & 15) == 0 && ((uintptr_t)d_attn_cat & 15) == 0 &&
                     ((uintptr_t)q & 15) == 0 && ((uintptr_t)dq_partial & 15) == 0, "segattn_bwd: buffers must be 16B aligned");
  SegArgs a = {};
  a.q = q; a.KVp = KVp; a.offsets = offsets; a.B = B; a.T = T; a.H = H; a.d = d; a.hd = d / H; a.N_max = N_max;
  a.per_query = per_query ? 1 : 0; a.thr = per_query ? drop_thr : 0u; a.seed = make_seed(seed); a.probs = const_cast<float*>(probs);
  a.dO = d_attn_cat; a.dKVp = dKVp; a.dq_partial = dq_partial;
  // fixed part: tile32 staging + s_out + s_p + s_ds ; then as many query rows (<= 32) of [H][N_max] planes as fit
  const size_t plane = (size_t)H * N_max * sizeof(float);
  const size_t fixed = (size_t)(XS_TILE_FLOATS + XS_T * XS_T) * sizeof(float) + 2 * plane;
  int TT = 32;
  while (TT > 1 && fixed + TT * plane > 200 * 1024) TT >>= 1;
  if (fixed + TT * plane > 200 * 1024) { immtsf_set_error("segattn_bwd: H*N_max=%d too large for shared memory", H * N_max); return IMMTSF_ERR_UNSUPPORTED; }
  if (!a.per_query) TT = 1;
  a.TT = TT;
  const size_t smem = fixed + TT * plane;
  IMMTSF_REQUIRE((d & 3) == 0, "segattn_bwd: d must be a multiple of 4");
  static size_t smem_set = 0;
  if (smem > smem_set) { cudaFuncSetAttribute(segattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); smem_set = smem; }
  segattn_bwd_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("segattn_bwd");
  return IMMTSF_OK;
}
