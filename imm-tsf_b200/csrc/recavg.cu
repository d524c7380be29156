// K2: TTF_RecAvg pooling, forward and backward.
//
// Forward replaces fusions/TTF_RecAvg.py:94-106:
//   delta = clamp_min(t_hat - tau, 0); w = exp(-(delta/sigma)^2) (masked notes
//   simply do not exist in the ragged layout); E_raw = sum_n w V'_n /
//   clamp_min(sum_n w, 1e-6); E_drop = dropout(LayerNorm(E_raw)).
// One CTA per (sample, tile of TT query times).  Each thread owns NCH float4
// column groups of the d-wide row and TT accumulators per group; every V'
// row of the segment is streamed once per query tile with 128-bit loads
// (re-reads across tiles hit L1/L2: a segment is N_i * d * 4 B ~ 48 KB), the
// recency weights are computed once per (note, t) into shared memory, the
// LayerNorm statistics are block reductions over the register tile, and the
// dropout mask is Philox on the flat [B,T,d] index.
//
// Algorithmic HBM bytes per sample (fp32): 4*(N_i*d [V'] + N_i [tau] + T
// [t_hat] + T*d [E_drop out] (+ T*d E_raw when training)).
//
// Backward (autograd of the same lines): LayerNorm backward per row, then
//   dS_t = dE_raw_t / den_t,  dV'_n = sum_t w_nt dS_t,
//   dlog_sigma = sum_{n,t} (dS_t . V'_n + dwsum_t) * w_nt * 2 (delta/sigma)^2
// where the (n,t) dot products are never formed: sum_n c_nt (dS_t . V'_n) =
// dS_t . (sum_n c_nt V'_n), a second pooled vector accumulated in the same
// pass.  d(den) uses the closed form sum_j dE_raw_j E_raw_j = s2*eps*rstd^2
// (LayerNorm is scale invariant up to eps).
#include <stdlib.h>
#include "rowwarp.cuh"
#include "../../include/immtsf.h"

struct PoolArgs {
  const float* Vp; int ldv;
  const float* tau; const int32_t* offsets;
  const float* t_hat; int t_bstride;
  const float* log_sigma; const float* gamma; const float* beta;
  int B, T, d; float eps; uint32_t thr; SeedArg seed;
  float* E_drop; float* E_raw; float* mean; float* rstd; float* wsum;
  // backward only
  const float* dE_drop; float* dVp; int lddv; float* dgamma; float* dbeta; double* dlog_sigma; float* dS; int N_max;
};

#ifndef IMMTSF_RECAVG_FUSED_BWD_DEFAULT
#define IMMTSF_RECAVG_FUSED_BWD_DEFAULT 1  // one-launch backward (the whole GPU suite is green with it; =0: two-kernel path)
#endif
constexpr int POOL_NB = 32;  // notes per shared-memory weight block

// ncu (B 2048, N<=16, T 24, d 768) showed the first version of this kernel issue-bound, not memory-bound (47 % issue
// utilisation with 18 warps/SM, long-scoreboard stalls ~1): hence the register cap (more resident warps), the
// reciprocal instead of 32 IEEE divisions per thread, and only two rows of V' in flight.
template <int NCH>
__global__ void __launch_bounds__(256, NCH == 1 ? 4 : 2) recavg_pool_fwd_kernel(const PoolArgs a) {
  constexpr int TT = 8 / NCH;
  __shared__ float s_w[POOL_NB][TT];
  __shared__ float s_red[32 * TT];
  __shared__ float s_th[TT];
  const int b = blockIdx.y, t0 = blockIdx.x * TT;  // the tiles of one sample are neighbours: they share its V' rows in L2
  const int nb = a.offsets[b], ne = a.offsets[b + 1];
  const float sigma = expf(__ldg(a.log_sigma));
  const int d4 = a.d >> 2;
  if (threadIdx.x < TT)
    s_th[threadIdx.x] = (t0 + threadIdx.x < a.T) ? a.t_hat[(size_t)b * a.t_bstride + t0 + threadIdx.x] : 0.f;

  float4 acc[TT][NCH];
  float wsum[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    wsum[t] = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[t][c] = f4_zero();
  }
  for (int n0 = nb; n0 < ne; n0 += POOL_NB) {
    __syncthreads();
    for (int i = threadIdx.x; i < POOL_NB * TT; i += blockDim.x) {
      const int nn = i / TT, t = i % TT, n = n0 + nn;
      float w = 0.f;
      if (n < ne) {
        const float delta = fmaxf(s_th[t] - __ldg(a.tau + n), 0.f);
        const float r = delta / sigma;
        w = expf(-(r * r));
      }
      s_w[nn][t] = w;
    }
    __syncthreads();
    const int cnt = min(POOL_NB, ne - n0);
    for (int nq = 0; nq < cnt; nq += 2) {  // 2 rows in flight per thread
      float4 v[2][NCH];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col4 = threadIdx.x + c * blockDim.x;
          v[u][c] = (nq + u < cnt && col4 < d4)
                        ? __ldg(reinterpret_cast<const float4*>(a.Vp + (size_t)(n0 + nq + u) * a.ldv) + col4) : f4_zero();
        }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (nq + u < cnt) {
#pragma unroll
          for (int t = 0; t < TT; ++t) {
            const float w = s_w[nq + u][t];
            wsum[t] += w;
#pragma unroll
            for (int c = 0; c < NCH; ++c) f4_fma(acc[t][c], w, v[u][c]);
          }
        }
      }
    }
  }
  // E_raw = E_wsum / clamp_min(denom, 1e-6)
  float s1[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const float inv_den = 1.f / fmaxf(wsum[t], 1e-6f);  // x * (1/den) is within 1 ulp of x / den
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      acc[t][c].x *= inv_den; acc[t][c].y *= inv_den; acc[t][c].z *= inv_den; acc[t][c].w *= inv_den;
      s += f4_sum(acc[t][c]);  // columns >= d are exactly 0
    }
    s1[t] = s;
  }
  block_sum_multi<TT>(s1, s_red);
  float s2[TT];
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const float mu = s1[t] / (float)a.d;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col4 = threadIdx.x + c * blockDim.x;
      if (col4 < d4) {
        const float dx = acc[t][c].x - mu, dy = acc[t][c].y - mu, dz = acc[t][c].z - mu, dw = acc[t][c].w - mu;
        s += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    s2[t] = s;
  }
  block_sum_multi<TT>(s2, s_red);
  const float inv_keep = inv_keep_from_thr(a.thr);
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const int tt = t0 + t;
    if (tt >= a.T) continue;
    const float mu = s1[t] / (float)a.d;
    const float rs = 1.f / sqrtf(s2[t] / (float)a.d + a.eps);
    const size_t rowi = (size_t)b * a.T + tt;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col4 = threadIdx.x + c * blockDim.x;
      if (col4 >= d4) continue;
      const float4 x = acc[t][c];
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + col4);
      const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta) + col4);
      const float4 ks = dropout_scale4(resolve_seed(a.seed), IMMTSF_SITE_TTF_DROPOUT, rowi * d4 + col4, a.thr, inv_keep);
      float4 y;
      y.x = ((x.x - mu) * rs * g.x + be.x) * ks.x;
      y.y = ((x.y - mu) * rs * g.y + be.y) * ks.y;
      y.z = ((x.z - mu) * rs * g.z + be.z) * ks.z;
      y.w = ((x.w - mu) * rs * g.w + be.w) * ks.w;
      reinterpret_cast<float4*>(a.E_drop + rowi * a.d)[col4] = y;
      if (a.E_raw) reinterpret_cast<float4*>(a.E_raw + rowi * a.d)[col4] = x;
    }
    if (threadIdx.x == 0) {
      if (a.mean) a.mean[rowi] = mu;
      if (a.rstd) a.rstd[rowi] = rs;
      if (a.wsum) a.wsum[rowi] = wsum[t];
    }
  }
}

// Backward, phase 1: one CTA per (sample, tile of TT query rows): LayerNorm backward of the tile's rows gives
// dS_t = dE_raw_t / den_t (written to the dS scratch) and the scalar d(den_t) (written behind it).  No note is
// touched here: every term that contracts over notes is formed in phase 2.
template <int NCH>
__global__ void __launch_bounds__(256, NCH == 1 ? 3 : 2) recavg_bwd_rows_kernel(const PoolArgs a) {
  constexpr int TT = 8 / NCH;
  __shared__ float s_red[32 * TT];
  const int b = blockIdx.y, t0 = blockIdx.x * TT;
  const int d4 = a.d >> 2;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const float inv_d = 1.f / (float)a.d;
  const uint64_t seed = resolve_seed(a.seed);
  float* dwsum_out = a.dS + (size_t)a.B * a.T * a.d;  // [B*T] scalars behind the dS rows

  float4 g[TT][NCH], xh[TT][NCH], dgam[NCH], dbet[NCH];
  float s1[TT], s2[TT], rs[TT], den[TT], dwsum[TT];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { dgam[c] = f4_zero(); dbet[c] = f4_zero(); }
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    const int tt = t0 + t;
    const bool ok = tt < a.T;
    const size_t rowi = (size_t)b * a.T + (ok ? tt : 0);
    const float mu = ok ? a.mean[rowi] : 0.f;
    rs[t] = ok ? a.rstd[rowi] : 0.f;
    const float ws = ok ? a.wsum[rowi] : 1.f;
    den[t] = fmaxf(ws, 1e-6f);
    dwsum[t] = (ok && ws >= 1e-6f) ? 1.f : 0.f;  // clamp_min passes gradient where wsum >= 1e-6
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col4 = threadIdx.x + c * blockDim.x;
      g[t][c] = f4_zero();
      xh[t][c] = f4_zero();
      if (ok && col4 < d4) {
        float4 dy = __ldg(reinterpret_cast<const float4*>(a.dE_drop + rowi * a.d) + col4);
        const float4 ks = dropout_scale4(seed, IMMTSF_SITE_TTF_DROPOUT, rowi * d4 + col4, a.thr, inv_keep);
        dy.x *= ks.x; dy.y *= ks.y; dy.z *= ks.z; dy.w *= ks.w;
        const float4 x = __ldg(reinterpret_cast<const float4*>(a.E_raw + rowi * a.d) + col4);
        float4 h;
        h.x = (x.x - mu) * rs[t]; h.y = (x.y - mu) * rs[t]; h.z = (x.z - mu) * rs[t]; h.w = (x.w - mu) * rs[t];
        dgam[c].x = fmaf(dy.x, h.x, dgam[c].x); dgam[c].y = fmaf(dy.y, h.y, dgam[c].y);
        dgam[c].z = fmaf(dy.z, h.z, dgam[c].z); dgam[c].w = fmaf(dy.w, h.w, dgam[c].w);
        f4_add(dbet[c], dy);
        const float4 ga = __ldg(reinterpret_cast<const float4*>(a.gamma) + col4);
        float4 gg;
        gg.x = dy.x * ga.x; gg.y = dy.y * ga.y; gg.z = dy.z * ga.z; gg.w = dy.w * ga.w;
        g[t][c] = gg;
        xh[t][c] = h;
        p1 += f4_sum(gg);
        p2 += f4_dot(gg, h);
      }
    }
    s1[t] = p1;
    s2[t] = p2;
  }
  block_sum_multi<TT>(s1, s_red);
  block_sum_multi<TT>(s2, s_red);
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    if (t0 + t >= a.T) continue;
    const float m1 = s1[t] * inv_d, m2 = s2[t] * inv_d;
    const float sc = rs[t] / den[t];
    const size_t rowi = (size_t)b * a.T + t0 + t;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col4 = threadIdx.x + c * blockDim.x;
      if (col4 >= d4) continue;
      // dE_raw = rstd * (g - mean(g) - xhat * mean(g*xhat));  dS = dE_raw / den
      float4 o;
      o.x = sc * (g[t][c].x - m1 - xh[t][c].x * m2);
      o.y = sc * (g[t][c].y - m1 - xh[t][c].y * m2);
      o.z = sc * (g[t][c].z - m1 - xh[t][c].z * m2);
      o.w = sc * (g[t][c].w - m1 - xh[t][c].w * m2);
      reinterpret_cast<float4*>(a.dS + rowi * a.d)[col4] = o;
    }
    // d(den) = -sum_j dE_raw_j E_raw_j / den = -(s2 * eps * rstd^2) / den
    if (threadIdx.x == 0) dwsum_out[rowi] = dwsum[t] * (-(s2[t] * a.eps * rs[t] * rs[t]) / den[t]);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col4 = threadIdx.x + c * blockDim.x;
    if (col4 < d4) {
      float* pg = a.dgamma + col4 * 4;
      float* pb = a.dbeta + col4 * 4;
      atomicAdd(pg + 0, dgam[c].x); atomicAdd(pg + 1, dgam[c].y); atomicAdd(pg + 2, dgam[c].z); atomicAdd(pg + 3, dgam[c].w);
      atomicAdd(pb + 0, dbet[c].x); atomicAdd(pb + 1, dbet[c].y); atomicAdd(pb + 2, dbet[c].z); atomicAdd(pb + 3, dbet[c].w);
    }
  }
}

// Backward, phase 2: one CTA per (sample, tile of NT notes).  With w_nt the recency weight and
// c_nt = dw_nt/dlog_sigma = w_nt * 2 (delta/sigma)^2, both contractions over the query rows have the same shape:
//   dV'_n = sum_t w_nt dS_t                      (written once, no read-modify-write)
//   Q_n   = sum_t c_nt dS_t,  dlog_sigma += Q_n . V'_n + sum_t c_nt d(den_t)
// so dS rows are streamed once (4 in flight) into two register accumulators per note.
// mbarrier + 1-D bulk asynchronous copy (TMA) helpers shared by the staged kernels below
__device__ __forceinline__ uint32_t rs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rs_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rs_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "RS_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra RS_WAIT_DONE;\n\t"
      "bra RS_WAIT_LOOP;\n\t"
      "RS_WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void rs_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void f4_fma_s(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

constexpr int POOL_TB = 32;  // query rows per shared-memory weight block
template <int NCH>
__global__ void __launch_bounds__(256, NCH == 1 ? 2 : 1) recavg_bwd_notes_kernel(const PoolArgs a) {
  constexpr int NT = 8 / NCH;  // notes per CTA
  __shared__ float s_w[POOL_TB][NT];
  __shared__ float s_c[POOL_TB][NT];
  __shared__ float s_dw[POOL_TB];
  __shared__ double s_redd[8];
  const int b = blockIdx.y;
  const int nb = a.offsets[b], ne = a.offsets[b + 1];
  const int n0 = nb + blockIdx.x * NT;
  if (n0 >= ne) return;
  const int ncnt = min(NT, ne - n0);
  const float sigma = expf(__ldg(a.log_sigma));
  const int d4 = a.d >> 2;
  const float* dwsum_in = a.dS + (size_t)a.B * a.T * a.d;
  float4 accw[NT][NCH], accc[NT][NCH];
#pragma unroll
  for (int u = 0; u < NT; ++u)
#pragma unroll
    for (int c = 0; c < NCH; ++c) { accw[u][c] = f4_zero(); accc[u][c] = f4_zero(); }
  float sc_term = 0.f;  // sum_t c_nt d(den_t), lane u < NT of warp 0 owns note u
  for (int t0 = 0; t0 < a.T; t0 += POOL_TB) {
    __syncthreads();
    for (int i = threadIdx.x; i < POOL_TB * NT; i += blockDim.x) {
      const int tt = i / NT, u = i % NT;
      float w = 0.f, cc = 0.f;
      if (t0 + tt < a.T && u < ncnt) {
        const float delta = fmaxf(a.t_hat[(size_t)b * a.t_bstride + t0 + tt] - __ldg(a.tau + n0 + u), 0.f);
        const float r = delta / sigma;
        w = expf(-(r * r));
        cc = w * 2.f * r * r;
      }
      s_w[tt][u] = w;
      s_c[tt][u] = cc;
    }
    if (threadIdx.x < POOL_TB) s_dw[threadIdx.x] = t0 + threadIdx.x < a.T ? dwsum_in[(size_t)b * a.T + t0 + threadIdx.x] : 0.f;
    __syncthreads();
    const int tcnt = min(POOL_TB, a.T - t0);
    if (threadIdx.x < NT)
      for (int tt = 0; tt < tcnt; ++tt) sc_term = fmaf(s_c[tt][threadIdx.x], s_dw[tt], sc_term);
    constexpr int PF = NCH == 1 ? 8 : 4;  // dS rows in flight per thread: the loop is bound by L2 latency otherwise
    for (int tq = 0; tq < tcnt; tq += PF) {
      float4 g[PF][NCH];
#pragma unroll
      for (int v = 0; v < PF; ++v)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int col4 = threadIdx.x + c * blockDim.x;
          g[v][c] = (tq + v < tcnt && col4 < d4)
                        ? __ldg(reinterpret_cast<const float4*>(a.dS + ((size_t)b * a.T + t0 + tq + v) * a.d) + col4) : f4_zero();
        }
      if (NT == 8 && ncnt <= 4) {  // half-empty tile (CTA-uniform): skip the empty note slots
#pragma unroll
        for (int v = 0; v < PF; ++v) {
          if (tq + v < tcnt) {
#pragma unroll
            for (int u = 0; u < NT / 2; ++u) {
              const float w = s_w[tq + v][u], cc = s_c[tq + v][u];
#pragma unroll
              for (int c = 0; c < NCH; ++c) { f4_fma(accw[u][c], w, g[v][c]); f4_fma(accc[u][c], cc, g[v][c]); }
            }
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < PF; ++v) {
          if (tq + v < tcnt) {
#pragma unroll
            for (int u = 0; u < NT; ++u) {
              const float w = s_w[tq + v][u], cc = s_c[tq + v][u];
#pragma unroll
              for (int c = 0; c < NCH; ++c) { f4_fma(accw[u][c], w, g[v][c]); f4_fma(accc[u][c], cc, g[v][c]); }
            }
          }
        }
      }
    }
  }
  // The terms Q_n . V'_n cancel almost completely across notes and columns (dS_t is orthogonal to the pooled row),
  // so this one scalar is summed in double from the thread level up to the global accumulator.
  double dls = 0.0;
#pragma unroll
  for (int u = 0; u < NT; ++u) {
    if (u < ncnt) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        if (col4 < d4) {
          reinterpret_cast<float4*>(a.dVp + (size_t)(n0 + u) * a.lddv)[col4] = accw[u][c];
          const float4 v = __ldg(reinterpret_cast<const float4*>(a.Vp + (size_t)(n0 + u) * a.ldv) + col4);
          const float4 q = accc[u][c];
          dls += (double)q.x * v.x + (double)q.y * v.y + (double)q.z * v.z + (double)q.w * v.w;
        }
      }
    }
  }
  if (threadIdx.x < NT) dls += (double)sc_term;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dls += __shfl_xor_sync(0xffffffffu, dls, o);
  if ((threadIdx.x & 31) == 0) s_redd[threadIdx.x >> 5] = dls;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += s_redd[w];
    atomicAdd(a.dlog_sigma, t);
  }
}

// Notes kernel with the sample's dS rows staged by one bulk asynchronous copy per tile of TB query rows (they are
// contiguous: [T][d] per sample) instead of PF-deep dependent loads through L1/L2 -- the register version was latency-bound
// (long-scoreboard 4.0 per issue at 18 % warps active, profiles/r1_ncu_recavg_v4_summary.txt).  d <= 1024 (one float4
// column per thread), 8 notes per CTA.  The weights of the tile are computed while the copy is in flight.
__global__ void __launch_bounds__(256, 2) recavg_bwd_notes_s_kernel(const PoolArgs a, int TB) {
  constexpr int NT = 8;
  extern __shared__ __align__(128) float s_g[];  // [TB][d]
  __shared__ __align__(16) float s_w[POOL_TB][NT];
  __shared__ __align__(16) float s_c[POOL_TB][NT];
  __shared__ float s_dw[POOL_TB];
  __shared__ double s_redd[8];
  __shared__ __align__(8) unsigned long long s_bar;
  const int b = blockIdx.y;
  const int nb = a.offsets[b], ne = a.offsets[b + 1];
  const int n0 = nb + blockIdx.x * NT;
  if (n0 >= ne) return;
  const int ncnt = min(NT, ne - n0);
  const float sigma = expf(__ldg(a.log_sigma));
  const int d4 = a.d >> 2;
  const float* dwsum_in = a.dS + (size_t)a.B * a.T * a.d;
  const uint32_t bar = rs_smem_u32(&s_bar), sg = rs_smem_u32(s_g);
  if (threadIdx.x == 0) rs_mbar_init(bar, 1);
  float4 accw[NT], accc[NT];
#pragma unroll
  for (int u = 0; u < NT; ++u) { accw[u] = f4_zero(); accc[u] = f4_zero(); }
  float sc_term = 0.f;  // sum_t c_nt d(den_t), thread u < NT owns note u
  uint32_t phase = 0;
  const bool half = ncnt <= 4;  // half-empty tile (CTA-uniform): skip the empty note slots
  for (int t0 = 0; t0 < a.T; t0 += TB) {
    const int tcnt = min(TB, a.T - t0);
    __syncthreads();  // the previous tile is consumed (and the barrier initialised, first time round)
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)tcnt * (uint32_t)a.d * 4u;
      rs_mbar_expect_tx(bar, bytes);
      rs_bulk_g2s(sg, a.dS + ((size_t)b * a.T + t0) * a.d, bytes, bar);
    }
    for (int i = threadIdx.x; i < TB * NT; i += blockDim.x) {
      const int tt = i / NT, u = i % NT;
      float w = 0.f, cc = 0.f;
      if (tt < tcnt && u < ncnt) {
        const float delta = fmaxf(a.t_hat[(size_t)b * a.t_bstride + t0 + tt] - __ldg(a.tau + n0 + u), 0.f);
        const float r = delta / sigma;
        w = expf(-(r * r));
        cc = w * 2.f * r * r;
      }
      s_w[tt][u] = w;
      s_c[tt][u] = cc;
    }
    if (threadIdx.x < TB) s_dw[threadIdx.x] = threadIdx.x < tcnt ? dwsum_in[(size_t)b * a.T + t0 + threadIdx.x] : 0.f;
    __syncthreads();
    if (threadIdx.x < NT)
      for (int tt = 0; tt < tcnt; ++tt) sc_term = fmaf(s_c[tt][threadIdx.x], s_dw[tt], sc_term);
    rs_mbar_wait(bar, phase);
    phase ^= 1u;
    if ((int)threadIdx.x < d4) {
      const float4* gp = reinterpret_cast<const float4*>(s_g) + threadIdx.x;
      for (int tt = 0; tt < tcnt; ++tt) {
        const float4 g = gp[(size_t)tt * d4];
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[tt][0]), c0 = *reinterpret_cast<const float4*>(&s_c[tt][0]);
        f4_fma(accw[0], w0.x, g); f4_fma(accc[0], c0.x, g);
        f4_fma(accw[1], w0.y, g); f4_fma(accc[1], c0.y, g);
        f4_fma(accw[2], w0.z, g); f4_fma(accc[2], c0.z, g);
        f4_fma(accw[3], w0.w, g); f4_fma(accc[3], c0.w, g);
        if (!half) {
          const float4 w1 = *reinterpret_cast<const float4*>(&s_w[tt][4]), c1 = *reinterpret_cast<const float4*>(&s_c[tt][4]);
          f4_fma(accw[4], w1.x, g); f4_fma(accc[4], c1.x, g);
          f4_fma(accw[5], w1.y, g); f4_fma(accc[5], c1.y, g);
          f4_fma(accw[6], w1.z, g); f4_fma(accc[6], c1.z, g);
          f4_fma(accw[7], w1.w, g); f4_fma(accc[7], c1.w, g);
        }
      }
    }
  }
  // (same epilogue as recavg_bwd_notes_kernel: dV' rows out, the scalar d log sigma summed in double)
  double dls = 0.0;
  if ((int)threadIdx.x < d4) {
#pragma unroll
    for (int h0 = 0; h0 < NT; h0 += 4) {  // four V' rows requested before the first store (which may alias, as far as the compiler knows)
      float4 vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        vv[u] = h0 + u < ncnt ? __ldg(reinterpret_cast<const float4*>(a.Vp + (size_t)(n0 + h0 + u) * a.ldv) + threadIdx.x) : f4_zero();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 q = accc[h0 + u], v = vv[u];  // empty slots: q == v == 0
        dls += (double)q.x * v.x + (double)q.y * v.y + (double)q.z * v.z + (double)q.w * v.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (h0 + u < ncnt) reinterpret_cast<float4*>(a.dVp + (size_t)(n0 + h0 + u) * a.lddv)[threadIdx.x] = accw[h0 + u];
    }
  }
  if (threadIdx.x < NT) dls += (double)sc_term;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dls += __shfl_xor_sync(0xffffffffu, dls, o);
  if ((threadIdx.x & 31) == 0) s_redd[threadIdx.x >> 5] = dls;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += s_redd[w];
    atomicAdd(a.dlog_sigma, t);
  }
}

// ------------------------------------------------------------------ warp-per-row variants (d % 8 == 0, d <= 1024)
// Forward: a CTA is 8 warps = 8 consecutive query times of one sample; each warp pools the sample's notes for its own
// query time (the 8 warps read the same V' rows back to back, so 7 of 8 reads hit L1), weights are computed by the
// lanes (one note per lane) and broadcast with shuffles.  No shared memory, no block barrier.
template <int NC, int TPW>
__global__ void __launch_bounds__(256) recavg_pool_fwd_w_kernel(const PoolArgs a) {
  // warp w of the CTA owns the TPW query times t0 + w, t0 + w + 8, ...: every V' chunk it loads feeds TPW
  // accumulators (the first version, one query time per warp, was bound by L1 bandwidth: 74 % l1tex throughput)
  const int b = blockIdx.y, tb = blockIdx.x * (8 * TPW) + (threadIdx.x >> 5);
  if (tb >= a.T) return;
  const int lane = threadIdx.x & 31, d8 = a.d >> 3;
  const int nb = a.offsets[b], ne = a.offsets[b + 1];
  const float inv_sigma = 1.f / expf(__ldg(a.log_sigma));
  float th[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) th[q] = tb + 8 * q < a.T ? a.t_hat[(size_t)b * a.t_bstride + tb + 8 * q] : 0.f;
  float acc[TPW][NC][8];
  float wsum_l[TPW];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    wsum_l[q] = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) zero8(acc[q][i]);
  }
  for (int n0 = nb; n0 < ne; n0 += 32) {
    float wl[TPW];
    const float tn = n0 + lane < ne ? __ldg(a.tau + n0 + lane) : 0.f;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float r = fmaxf(th[q] - tn, 0.f) * inv_sigma;
      wl[q] = n0 + lane < ne ? expf(-(r * r)) : 0.f;
      wsum_l[q] += wl[q];
    }
    const int cnt = min(32, ne - n0);
    for (int j = 0; j < cnt; ++j) {
      const float* r0 = a.Vp + (size_t)(n0 + j) * a.ldv;
      float v0[NC][8];
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        if (lane + 32 * i < d8) load8(r0, lane + 32 * i, v0[i]);
        else zero8(v0[i]);
      }
#pragma unroll
      for (int q = 0; q < TPW; ++q) {
        const float w0 = __shfl_sync(0xffffffffu, wl[q], j);
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[q][i][e] = fmaf(w0, v0[i][e], acc[q][i][e]);
      }
    }
  }
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  const float inv_d = 1.f / (float)a.d;
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int t = tb + 8 * q;
    if (t >= a.T) break;
    const float wsum = warp_sum(wsum_l[q]);
    const float inv_den = 1.f / fmaxf(wsum, 1e-6f);  // E_raw = E_wsum / clamp_min(denom, 1e-6)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc[q][i][e] *= inv_den; s += acc[q][i][e]; }  // chunks beyond d are exactly 0
    const float mu = warp_sum(s) * inv_d;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (lane + 32 * i < d8)
#pragma unroll
        for (int e = 0; e < 8; ++e) qq = fmaf(acc[q][i][e] - mu, acc[q][i][e] - mu, qq);
    const float rs = 1.f / sqrtf(warp_sum(qq) * inv_d + a.eps);
    const size_t rowi = (size_t)b * a.T + t;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float g[8], be[8], ks[8], y[8];
        load8(a.gamma, k, g);
        load8(a.beta, k, be);
        dropout_scale8(seed, IMMTSF_SITE_TTF_DROPOUT, rowi * d8 + k, a.thr, inv_keep, ks);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ((acc[q][i][e] - mu) * rs * g[e] + be[e]) * ks[e];
        store8(a.E_drop + rowi * a.d, k, y);
        if (a.E_raw) store8(a.E_raw + rowi * a.d, k, acc[q][i]);
      }
    }
    if (lane == 0) {
      if (a.mean) a.mean[rowi] = mu;
      if (a.rstd) a.rstd[rowi] = rs;
      if (a.wsum) a.wsum[rowi] = wsum;
    }
  }
}

// ------------------------------------------------------------------ forward with TMA-staged segments
// Same warp-per-query-time arithmetic as recavg_pool_fwd_w_kernel, but the sample's V' rows reach the SM ONCE, as a bulk
// asynchronous copy (cp.async.bulk, completion on an mbarrier) into shared memory, instead of one dependent L1/L2 round trip
// per note inside the pooling loop (ncu on the register version: long-scoreboard stalls 11.6 per issue at 41 % issue
// utilisation -- latency-bound on exactly those loads).  The recency weights are computed while the copy is in flight.
// Bank conflicts: a lane owns float8 chunks (one Philox call per chunk), i.e. 32-byte strides between lanes, which would be
// 2-way conflicts for LDS.128.  Lanes 4-7 of every quarter warp therefore read the SECOND float4 of their chunk first
// (p = 1) and keep their halves swapped in registers until the epilogue; every LDS.128 wavefront then covers 8 distinct
// 16-byte bank groups.
// smem: s_v [RS][d] (RS <= 32 rows per stage).  grid (ceil(T / (8*TPW)), B), 256 threads.
// (Round 2 A/B-ed three other epilogues on a B200 -- Philox round keys in uniform registers, packed FFMA2 / FMUL2 / FADD2
// arithmetic, both -- which execute 11-20 % fewer instructions and were all 3-11 % SLOWER than this one
// (profiles/r2_ab_recavg_fwd_bwd.txt); what did help is the mul.wide.u32 in common.cuh's Philox: 106.5 -> 99.4 us.)
template <int NC, int TPW, int MINB, bool FULL>
__global__ void __launch_bounds__(256, MINB) recavg_pool_fwd_s_kernel(const PoolArgs a, int RS) {
  extern __shared__ __align__(128) float s_v[];
  __shared__ __align__(8) unsigned long long s_bar;
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int tb = blockIdx.x * (8 * TPW) + (threadIdx.x >> 5);
  const bool active = tb < a.T;  // inactive warps still take part in the barriers
  const int d = a.d, d8 = d >> 3;
  const int nb = a.offsets[b], ne = a.offsets[b + 1];
  const uint32_t bar = rs_smem_u32(&s_bar), sv = rs_smem_u32(s_v);
  if (threadIdx.x == 0) rs_mbar_init(bar, 1);
  __syncthreads();
  const float inv_sigma = 1.f / expf(__ldg(a.log_sigma));
  const int p = (lane >> 2) & 1;
  float th[TPW], wsum_l[TPW];
  float4 accA[TPW][NC], accB[TPW][NC];
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    th[q] = tb + 8 * q < a.T ? a.t_hat[(size_t)b * a.t_bstride + tb + 8 * q] : 0.f;
    wsum_l[q] = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) { accA[q][i] = f4_zero(); accB[q][i] = f4_zero(); }
  }
  uint32_t phase = 0;
  for (int n0 = nb; n0 < ne; n0 += RS) {
    const int cnt = min(RS, ne - n0);
    if (n0 != nb) __syncthreads();  // every warp is done reading the previous stage
    if (threadIdx.x == 0) {
      const uint32_t row_bytes = (uint32_t)d * 4u;
      rs_mbar_expect_tx(bar, (uint32_t)cnt * row_bytes);
      if (a.ldv == d) {
        rs_bulk_g2s(sv, a.Vp + (size_t)n0 * a.ldv, (uint32_t)cnt * row_bytes, bar);
      } else {
        for (int j = 0; j < cnt; ++j) rs_bulk_g2s(sv + (uint32_t)j * row_bytes, a.Vp + (size_t)(n0 + j) * a.ldv, row_bytes, bar);
      }
    }
    // recency weights of the stage's notes (lane j <-> note n0 + j) while the copy is in flight
    float wl[TPW];
    const float tn = lane < cnt ? __ldg(a.tau + n0 + lane) : 0.f;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float r = fmaxf(th[q] - tn, 0.f) * inv_sigma;
      wl[q] = lane < cnt ? expf(-(r * r)) : 0.f;
      wsum_l[q] += wl[q];
    }
    rs_mbar_wait(bar, phase);
    phase ^= 1u;
    if (active) {
      for (int j = 0; j < cnt; ++j) {
        const float4* row = reinterpret_cast<const float4*>(s_v + (size_t)j * d);
        float4 vA[NC], vB[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const int k = lane + 32 * i;
          if (FULL || k < d8) { vA[i] = row[2 * k + p]; vB[i] = row[2 * k + 1 - p]; }
          else { vA[i] = f4_zero(); vB[i] = f4_zero(); }
        }
#pragma unroll
        for (int q = 0; q < TPW; ++q) {
          const float w0 = __shfl_sync(0xffffffffu, wl[q], j);
#pragma unroll
          for (int i = 0; i < NC; ++i) { f4_fma_s(accA[q][i], w0, vA[i]); f4_fma_s(accB[q][i], w0, vB[i]); }
        }
      }
    }
  }
  if (!active) return;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  const float inv_d = 1.f / (float)a.d;
#pragma unroll
  for (int q = 0; q < TPW; ++q) {
    const int t = tb + 8 * q;
    if (t >= a.T) break;
    const float wsum = warp_sum(wsum_l[q]);
    const float inv_den = 1.f / fmaxf(wsum, 1e-6f);  // E_raw = E_wsum / clamp_min(denom, 1e-6)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {  // chunks beyond d are exactly 0
      float4& A = accA[q][i];
      float4& Bv = accB[q][i];
      A.x *= inv_den; A.y *= inv_den; A.z *= inv_den; A.w *= inv_den;
      Bv.x *= inv_den; Bv.y *= inv_den; Bv.z *= inv_den; Bv.w *= inv_den;
      s += (A.x + A.y) + (A.z + A.w) + (Bv.x + Bv.y) + (Bv.z + Bv.w);
    }
    const float mu = warp_sum(s) * inv_d;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (FULL || lane + 32 * i < d8) {
        const float4 A = accA[q][i], Bv = accB[q][i];
        qq = fmaf(A.x - mu, A.x - mu, qq); qq = fmaf(A.y - mu, A.y - mu, qq); qq = fmaf(A.z - mu, A.z - mu, qq); qq = fmaf(A.w - mu, A.w - mu, qq);
        qq = fmaf(Bv.x - mu, Bv.x - mu, qq); qq = fmaf(Bv.y - mu, Bv.y - mu, qq); qq = fmaf(Bv.z - mu, Bv.z - mu, qq); qq = fmaf(Bv.w - mu, Bv.w - mu, qq);
      }
    const float rs = 1.f / sqrtf(warp_sum(qq) * inv_d + a.eps);
    const size_t rowi = (size_t)b * a.T + t;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (FULL || k < d8) {
        float ks[8];
        dropout_scale8(seed, IMMTSF_SITE_TTF_DROPOUT, rowi * d8 + k, a.thr, inv_keep, ks);
        const int oA = 2 * k + p, oB = 2 * k + 1 - p;  // float4 index of each half within the row
        const float4 gA = __ldg(reinterpret_cast<const float4*>(a.gamma) + oA), gB = __ldg(reinterpret_cast<const float4*>(a.gamma) + oB);
        const float4 bA = __ldg(reinterpret_cast<const float4*>(a.beta) + oA), bB = __ldg(reinterpret_cast<const float4*>(a.beta) + oB);
        const float4 A = accA[q][i], Bv = accB[q][i];
        float4 kA, kB, yA, yB;
        kA.x = p ? ks[4] : ks[0]; kA.y = p ? ks[5] : ks[1]; kA.z = p ? ks[6] : ks[2]; kA.w = p ? ks[7] : ks[3];
        kB.x = p ? ks[0] : ks[4]; kB.y = p ? ks[1] : ks[5]; kB.z = p ? ks[2] : ks[6]; kB.w = p ? ks[3] : ks[7];
        yA.x = ((A.x - mu) * rs * gA.x + bA.x) * kA.x; yA.y = ((A.y - mu) * rs * gA.y + bA.y) * kA.y;
        yA.z = ((A.z - mu) * rs * gA.z + bA.z) * kA.z; yA.w = ((A.w - mu) * rs * gA.w + bA.w) * kA.w;
        yB.x = ((Bv.x - mu) * rs * gB.x + bB.x) * kB.x; yB.y = ((Bv.y - mu) * rs * gB.y + bB.y) * kB.y;
        yB.z = ((Bv.z - mu) * rs * gB.z + bB.z) * kB.z; yB.w = ((Bv.w - mu) * rs * gB.w + bB.w) * kB.w;
        float4* eo = reinterpret_cast<float4*>(a.E_drop + rowi * a.d);
        eo[oA] = yA;
        eo[oB] = yB;
        if (a.E_raw) {
          float4* er = reinterpret_cast<float4*>(a.E_raw + rowi * a.d);
          er[oA] = A;
          er[oB] = Bv;
        }
      }
    }
    if (lane == 0) {
      if (a.mean) a.mean[rowi] = mu;
      if (a.rstd) a.rstd[rowi] = rs;
      if (a.wsum) a.wsum[rowi] = wsum;
    }
  }
}

template <int NC, int TPW, int MINB, bool FULL>
static void launch_fwd_s2(const PoolArgs& a, dim3 grid, int RS, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem + 1024 > 48 * 1024 && smem > smem_set) {  // (the kernel also has 128 B of static shared memory)
    cudaFuncSetAttribute(recavg_pool_fwd_s_kernel<NC, TPW, MINB, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  recavg_pool_fwd_s_kernel<NC, TPW, MINB, FULL><<<grid, 256, smem, st>>>(a, RS);
}
// tpw: query times per warp.  3: capped at 128 registers, 2 CTAs per SM (MINB = 2); 1: capped at 80 registers, 3 CTAs per SM (MINB = 3).
template <int NC>
static void launch_fwd_s(const PoolArgs& a, int tpw, int T, int B, int RS, size_t smem, cudaStream_t st) {
  const bool full = (a.d >> 3) == 32 * NC;
  if (tpw != 3) tpw = 1;
  dim3 grid(ceil_div(T, 8 * tpw), B);
  if (tpw == 3) { if (full) launch_fwd_s2<NC, 3, 2, true>(a, grid, RS, smem, st); else launch_fwd_s2<NC, 3, 2, false>(a, grid, RS, smem, st); }
  else { if (full) launch_fwd_s2<NC, 1, 3, true>(a, grid, RS, smem, st); else launch_fwd_s2<NC, 1, 3, false>(a, grid, RS, smem, st); }
}

// Backward phase 1 (LayerNorm backward of the pooled rows -> dS, d(den)), one warp per (sample, query time) row.
template <int NC>
__global__ void __launch_bounds__(128) recavg_bwd_rows_w_kernel(const PoolArgs a) {
  __shared__ float s_acc[2 * 1024];  // dgamma | dbeta of this CTA
  const int d8 = a.d >> 3, lane = threadIdx.x & 31;
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)a.d;
  const uint64_t seed = resolve_seed(a.seed);
  float* dwsum_out = a.dS + (size_t)a.B * a.T * a.d;
  for (int i = threadIdx.x; i < 2 * a.d; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float dgam[NC][8], dbet[NC][8];
#pragma unroll
  for (int i = 0; i < NC; ++i) { zero8(dgam[i]); zero8(dbet[i]); }
  const int R = a.B * a.T;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < R; r += nw) {
    const float mu = a.mean[r], rs = a.rstd[r], ws = a.wsum[r];
    const float den = fmaxf(ws, 1e-6f);
    float g[NC][8], h[NC][8];
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      zero8(g[i]); zero8(h[i]);
      if (k < d8) {
        float dy[8], ks[8], x[8], ga[8];
        load8(a.dE_drop + (size_t)r * a.d, k, dy);
        load8(a.E_raw + (size_t)r * a.d, k, x);
        load8(a.gamma, k, ga);
        dropout_scale8(seed, IMMTSF_SITE_TTF_DROPOUT, (uint64_t)r * d8 + k, a.thr, inv_keep, ks);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dye = dy[e] * ks[e];
          const float he = (x[e] - mu) * rs;
          dgam[i][e] = fmaf(dye, he, dgam[i][e]);
          dbet[i][e] += dye;
          const float gg = dye * ga[e];
          g[i][e] = gg;
          h[i][e] = he;
          p1 += gg;
          p2 = fmaf(gg, he, p2);
        }
      }
    }
    const float s2 = warp_sum(p2);
    const float m1 = warp_sum(p1) * inv_d, m2 = s2 * inv_d;
    const float sc = rs / den;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float o[8];  // dE_raw = rstd * (g - mean(g) - xhat * mean(g*xhat));  dS = dE_raw / den
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sc * (g[i][e] - m1 - h[i][e] * m2);
        store8(a.dS + (size_t)r * a.d, k, o);
      }
    }
    // d(den) = -sum_j dE_raw_j E_raw_j / den = -(s2 * eps * rstd^2) / den ; clamp_min passes gradient where wsum >= 1e-6
    if (lane == 0) dwsum_out[r] = ws >= 1e-6f ? -(s2 * a.eps * rs * rs) / den : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int k = lane + 32 * i;
    if (k < d8)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_acc[8 * k + e], dgam[i][e]);
        atomicAdd(&s_acc[a.d + 8 * k + e], dbet[i][e]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    atomicAdd(a.dgamma + i, s_acc[i]);
    atomicAdd(a.dbeta + i, s_acc[a.d + i]);
  }
}

// The same rows kernel with a two-stage shared-memory ring per WARP: lane 0 posts the bulk asynchronous copies of the NEXT
// row's dE_drop and E_raw (2 x d x 4 bytes, completion on the warp's mbarrier of that stage) before the warp works on the
// current row, so two rows per warp are in flight instead of one (the register version: 3.06 TB/s of DRAM traffic at 12
// resident warps/SM, long-scoreboard 4.0 per issue -- memory-level parallelism, not bandwidth, was the limit).
// smem (dynamic): [4 warps][2 stages][2 rows][d].
__device__ __forceinline__ void lds8(const float* row, int k, float (&o)[8]) {
  const float4 a = reinterpret_cast<const float4*>(row)[2 * k];
  const float4 b = reinterpret_cast<const float4*>(row)[2 * k + 1];
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
template <int NC>
__global__ void __launch_bounds__(128) recavg_bwd_rows_s_kernel(const PoolArgs a) {
  extern __shared__ __align__(128) float s_ring[];
  __shared__ float s_acc[2 * 1024];  // dgamma | dbeta of this CTA
  __shared__ __align__(8) unsigned long long s_bar[4][2];
  const int d = a.d, d8 = d >> 3, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)d;
  const uint64_t seed = resolve_seed(a.seed);
  float* dwsum_out = a.dS + (size_t)a.B * a.T * d;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) s_acc[i] = 0.f;
  if (lane == 0) { rs_mbar_init(rs_smem_u32(&s_bar[w][0]), 1); rs_mbar_init(rs_smem_u32(&s_bar[w][1]), 1); }
  __syncthreads();
  float dgam[NC][8], dbet[NC][8];
#pragma unroll
  for (int i = 0; i < NC; ++i) { zero8(dgam[i]); zero8(dbet[i]); }
  const int R = a.B * a.T;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  float* ring = s_ring + (size_t)w * 4 * d;  // stage s: dy at ring + s*2*d, x at ring + s*2*d + d
  const uint32_t row_bytes = (uint32_t)d * 4u;
  if (lane == 0 && gw < R) {
    const uint32_t bar = rs_smem_u32(&s_bar[w][0]);
    rs_mbar_expect_tx(bar, 2u * row_bytes);
    rs_bulk_g2s(rs_smem_u32(ring), a.dE_drop + (size_t)gw * d, row_bytes, bar);
    rs_bulk_g2s(rs_smem_u32(ring + d), a.E_raw + (size_t)gw * d, row_bytes, bar);
  }
  int it = 0;
  for (int r = gw; r < R; r += nw, ++it) {
    const int st = it & 1;
    __syncwarp();  // every lane is done reading stage st^1 (the previous row)
    if (lane == 0 && r + nw < R) {
      const uint32_t bar = rs_smem_u32(&s_bar[w][st ^ 1]);
      float* nxt = ring + (size_t)(st ^ 1) * 2 * d;
      rs_mbar_expect_tx(bar, 2u * row_bytes);
      rs_bulk_g2s(rs_smem_u32(nxt), a.dE_drop + (size_t)(r + nw) * d, row_bytes, bar);
      rs_bulk_g2s(rs_smem_u32(nxt + d), a.E_raw + (size_t)(r + nw) * d, row_bytes, bar);
    }
    const float mu = a.mean[r], rs = a.rstd[r], ws = a.wsum[r];
    const float den = fmaxf(ws, 1e-6f);
    rs_mbar_wait(rs_smem_u32(&s_bar[w][st]), (uint32_t)((it >> 1) & 1));
    const float* sdy = ring + (size_t)st * 2 * d;
    const float* sx = sdy + d;
    float g[NC][8], h[NC][8];
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      zero8(g[i]); zero8(h[i]);
      if (k < d8) {
        float dy[8], ks[8], x[8], ga[8];
        lds8(sdy, k, dy);
        lds8(sx, k, x);
        load8(a.gamma, k, ga);
        dropout_scale8(seed, IMMTSF_SITE_TTF_DROPOUT, (uint64_t)r * d8 + k, a.thr, inv_keep, ks);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dye = dy[e] * ks[e];
          const float he = (x[e] - mu) * rs;
          dgam[i][e] = fmaf(dye, he, dgam[i][e]);
          dbet[i][e] += dye;
          const float gg = dye * ga[e];
          g[i][e] = gg;
          h[i][e] = he;
          p1 += gg;
          p2 = fmaf(gg, he, p2);
        }
      }
    }
    const float s2 = warp_sum(p2);
    const float m1 = warp_sum(p1) * inv_d, m2 = s2 * inv_d;
    const float sc = rs / den;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float o[8];  // dE_raw = rstd * (g - mean(g) - xhat * mean(g*xhat));  dS = dE_raw / den
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sc * (g[i][e] - m1 - h[i][e] * m2);
        store8(a.dS + (size_t)r * d, k, o);
      }
    }
    if (lane == 0) dwsum_out[r] = ws >= 1e-6f ? -(s2 * a.eps * rs * rs) / den : 0.f;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int k = lane + 32 * i;
    if (k < d8)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_acc[8 * k + e], dgam[i][e]);
        atomicAdd(&s_acc[d + 8 * k + e], dbet[i][e]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(a.dgamma + i, s_acc[i]);
    atomicAdd(a.dbeta + i, s_acc[d + i]);
  }
}

template <int NC>
static void launch_rows_s(const PoolArgs& a, int want, cudaStream_t st) {
  const size_t smem = (size_t)4 * 4 * a.d * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(recavg_bwd_rows_s_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * 1024 * (int)sizeof(float));
    attr = true;
  }
  recavg_bwd_rows_s_kernel<NC><<<resident_grid((const void*)recavg_bwd_rows_s_kernel<NC>, 128, smem, want, 4), 128, smem, st>>>(a);
}

// ------------------------------------------------------------------ backward in ONE launch (short prediction windows)
// The two-kernel backward writes dS [B*T, d] to HBM and reads it back once per tile of 8 notes (B 2048, N <= 16, T 24, d 768:
// 151 MB written + up to 302 MB read against 405 MB of algorithmic traffic).  Here a persistent CTA owns one sample at a time and
// dS never leaves shared memory:
//   1. per-row bulk asynchronous copies bring the sample's dE_drop rows [T][d] into s_g (one mbarrier per row: a row warp starts
//      when ITS row has landed); each warp streams its E_raw rows (t = w, w + 8, ...) through a private one-row buffer;
//   2. rows phase, warp per query row, two passes over shared memory (so that g and x^ need no registers across the warp
//      reductions): s_g row <- dy*keep*gamma, then s_g row <- dS_t; d(den_t) -> s_dw[t];
//   3. the warps' dgamma / dbeta partials of the sample are exchanged through the (now idle) E_raw buffers and summed by
//      column-owner threads into shared-memory accumulators that live for the whole kernel;
//   4. note phase on tensor cores (below).
// Round 1's version of this kernel (note phase on CUDA cores, one tile barrier, Philox with its key schedule per call) was
// bound by instruction issue at 16 resident warps (profiles/r1_ncu_recavg_fused_bwd_summary.txt: 41 k warp instructions per
// sample, 27 % of them the note phase's FFMAs, 18 % Philox, 37 % of the shared-memory wavefronts 2-way bank conflicts); this
// one executes 26 k (profiles/r2_ncu_recavg_bwd_mma_summary.txt).  What changed:
//   * the note phase is ONE small matrix product per pass of 8 notes, [w ; c] (16 x T) times dS (T x d), on mma.sync
//     m16n8k8 TF32 with the 3xTF32 split (lo*hi + hi*lo + hi*hi, fp32 accumulate: fp32-exact to ~1e-6 like the tcgen05
//     GEMM, and T <= 32 keeps the accumulation chain short).  Rows 0-7 of the A operand are the recency weights w_nt of the
//     pass's notes, rows 8-15 their log-sigma sensitivities c_nt, so accumulator registers c0,c1 of a lane are dV'_n and c2,c3
//     are Q_n of the SAME note and columns: the Q_n . V'_n contraction needs no exchange.  A lane's A fragment is (note g =
//     lane/4, times 8ks + lane%4 and + 4): every lane evaluates its own 2 x T/8 weights -- no s_w / s_c arrays and no CTA
//     barrier inside the note phase.  ~9 instructions per (8 columns x 8 times) tile instead of 64 FFMA + 5 LDS.128.
//   * dS rows are padded to d + 8 floats: the B-fragment loads (4 times x 8 columns per instruction) hit 32 distinct banks.
//   * float8 lane ownership with swapped halves in the rows phase (lanes 4-7 of a quarter warp read the second float4 of
//     their chunk first, as in recavg_pool_fwd_s_kernel): LDS.128 / STS.128 without bank conflicts.
//   * one mbarrier per dE_drop row, so a row warp starts when ITS row has landed.  (L2 prefetches of the sample's later E_raw rows
//     and of the next sample's tiles, cp.async.bulk.prefetch.L2, measured neutral to 5 % slower and were removed.)
// Requires T <= POOL_TB, d % 8 == 0, d <= 1024; dynamic shared memory T * (d + 8) * 4 + 8 * d * 4 bytes.
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo exactly, hi = x with the 13 low mantissa bits cleared (what a TF32 operand keeps)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void f4_to8(const float4& lo, const float4& hi, float (&o)[8]) {
  o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = lo.w; o[4] = hi.x; o[5] = hi.y; o[6] = hi.z; o[7] = hi.w;
}
// one k-step (8 query times) of a warp's 8-column tiles: acc[j] += [w ; c] (16 x 8) x dS (8 x 8), 3xTF32.
// r0 / r1: the lane's B-fragment rows (times tl and tl + 4 of the k-step) at the warp's first tile.
template <int NTW, bool GUARD>
__device__ __forceinline__ void note_tiles(float (&acc)[NTW][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           const float* r0, const float* r1, int nlive) {
#pragma unroll
  for (int j = 0; j < NTW; ++j) {
    if (!GUARD || j < nlive) {  // (warp-uniform)
      uint32_t bh0, bl0, bh1, bl1;
      tf32_split(r0[8 * j], bh0, bl0);
      tf32_split(r1[8 * j], bh1, bl1);
      mma_tf32_m16n8k8(acc[j], al, bh0, bh1);
      mma_tf32_m16n8k8(acc[j], ah, bl0, bl1);
      mma_tf32_m16n8k8(acc[j], ah, bh0, bh1);
    }
  }
}
template <int NC, bool FULL>
__global__ void __launch_bounds__(256, 2) recavg_bwd_mma_kernel(const PoolArgs a) {
  constexpr int NTW = 4 * NC;  // 8-column tiles per warp in the note phase (8 warps x NTW x 8 >= 256 * NC >= d)
  extern __shared__ __align__(128) float s_dyn[];  // s_g [T][d + 8] | s_x [8 warps][d] | s_col [2][d]
  __shared__ float s_dw[POOL_TB];
  __shared__ float s_th[POOL_TB];
  __shared__ double s_redd[8];
  __shared__ __align__(8) unsigned long long s_barr[POOL_TB];  // one per dE_drop row
  __shared__ __align__(8) unsigned long long s_barx[8];        // one per warp (its E_raw row)
  const int d = a.d, d8 = d >> 3, d4 = d >> 2, T = a.T, ldg = d + 8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = (lane >> 2) & 1;           // rows phase: which half of its float8 chunk the lane reads first
  const int g = lane >> 2, tl = lane & 3;  // note phase: fragment coordinates
  float* s_g = s_dyn;
  float* s_x = s_dyn + (size_t)T * ldg;
  float* s_xw = s_x + (size_t)w * d;
  float4* s_col4 = reinterpret_cast<float4*>(s_x + (size_t)8 * d);  // dgamma [d] | dbeta [d] of this CTA, over all its samples
  const uint32_t barx = rs_smem_u32(&s_barx[w]);
  if (threadIdx.x < POOL_TB) rs_mbar_init(rs_smem_u32(&s_barr[threadIdx.x]), 1);
  if (lane == 0) rs_mbar_init(barx, 1);
  for (int i = threadIdx.x; i < 2 * d4; i += blockDim.x) s_col4[i] = f4_zero();
  __syncthreads();
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)d;
  const PhiloxKeys pkey = philox_keys(resolve_seed(a.seed));  // uniform registers: operands of the rows phase's LOP3s
  const float inv_sigma = 1.f / expf(__ldg(a.log_sigma));  // (the forward's expression)
  const uint32_t row_bytes = (uint32_t)d * 4u;
  const int per = (d8 + 7) >> 3, nt0 = w * per;  // this warp's 8-column tiles: nt0 .. nt0 + nlive - 1 (per <= NTW)
  const int nlive = max(0, min(per, d8 - nt0));
  const bool fullw = nlive == NTW;
  double dls = 0.0;
  uint32_t phg = 0, phx = 0;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const int nb = a.offsets[b], ne = a.offsets[b + 1];
    // generic-proxy writes of the previous sample (dS rows, gradient partials) are ordered before the bulk copies below
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (w == 0) {
      if (lane < T) {  // one padded row per lane, each on its own barrier: the row warps start as soon as THEIR row has landed
        const uint32_t bar = rs_smem_u32(&s_barr[lane]);
        rs_mbar_expect_tx(bar, row_bytes);
        rs_bulk_g2s(rs_smem_u32(s_g + (size_t)lane * ldg), a.dE_drop + ((size_t)b * T + lane) * d, row_bytes, bar);
      }
    }
    if (w < T && lane == 0) {
      rs_mbar_expect_tx(barx, row_bytes);
      rs_bulk_g2s(rs_smem_u32(s_xw), a.E_raw + ((size_t)b * T + w) * d, row_bytes, barx);
    }
    // read by the note phase (after the rows phase's barriers)
    if (w == 1 && lane < T) s_th[lane] = a.t_hat[(size_t)b * a.t_bstride + lane];
    float4 dgam[NC][2], dbet[NC][2];  // [chunk][half], halves in the lane's order
#pragma unroll
    for (int i = 0; i < NC; ++i) { dgam[i][0] = f4_zero(); dgam[i][1] = f4_zero(); dbet[i][0] = f4_zero(); dbet[i][1] = f4_zero(); }
    float mu_n = 0.f, rs_n = 0.f, ws_n = 0.f;
    if (w < T) { const size_t r0 = (size_t)b * T + w; mu_n = a.mean[r0]; rs_n = a.rstd[r0]; ws_n = a.wsum[r0]; }
    for (int t = w; t < T; t += 8) {
      const size_t r = (size_t)b * T + t;
      const float mu = mu_n, rs = rs_n, ws = ws_n;
      if (t + 8 < T) { mu_n = a.mean[r + 8]; rs_n = a.rstd[r + 8]; ws_n = a.wsum[r + 8]; }  // consumed one row later
      const float den = fmaxf(ws, 1e-6f);
      float4* sg4 = reinterpret_cast<float4*>(s_g + (size_t)t * ldg);
      const float4* sx4 = reinterpret_cast<const float4*>(s_xw);
      const float4* ga4 = reinterpret_cast<const float4*>(a.gamma);
      rs_mbar_wait(rs_smem_u32(&s_barr[t]), phg);
      rs_mbar_wait(barx, phx);
      phx ^= 1u;
      const float nmr = -mu * rs;
      const uint64_t idx8_0 = (uint64_t)r * d8 + lane;  // Philox counter of the lane's first chunk of this row
      float2 p1 = make_float2(0.f, 0.f), p2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int k = lane + 32 * i;
        if (FULL || k < d8) {
          const int fA = 2 * k + p, fB = 2 * k + 1 - p;
          float ks[8];
          dropout_scale8_sw(pkey, IMMTSF_SITE_TTF_DROPOUT, idx8_0 + (uint64_t)(32 * i), a.thr, inv_keep, p, ks);
          // x^ = x * rstd - mu * rstd;  dy~ = dy * keep;  dgamma += dy~ x^;  dbeta += dy~;  g = dy~ gamma;  p1 += g;  p2 += g x^
          const float4 dyA = f4_mul(sg4[fA], make_float4(ks[0], ks[1], ks[2], ks[3]));
          const float4 dyB = f4_mul(sg4[fB], make_float4(ks[4], ks[5], ks[6], ks[7]));
          const float4 hA = f4_fmass(sx4[fA], rs, nmr), hB = f4_fmass(sx4[fB], rs, nmr);
          dgam[i][0] = f4_fma3(dyA, hA, dgam[i][0]);
          dgam[i][1] = f4_fma3(dyB, hB, dgam[i][1]);
          dbet[i][0] = cat4(__fadd2_rn(lo2(dbet[i][0]), lo2(dyA)), __fadd2_rn(hi2(dbet[i][0]), hi2(dyA)));
          dbet[i][1] = cat4(__fadd2_rn(lo2(dbet[i][1]), lo2(dyB)), __fadd2_rn(hi2(dbet[i][1]), hi2(dyB)));
          const float4 gA = f4_mul(dyA, __ldg(ga4 + fA)), gB = f4_mul(dyB, __ldg(ga4 + fB));
          f2_acc_sum(p1, gA);
          f2_acc_sum(p1, gB);
          f2_acc_dot(p2, gA, hA);
          f2_acc_dot(p2, gB, hB);
          sg4[fA] = gA;  // re-read below by this lane only
          sg4[fB] = gB;
        }
      }
      const float s2 = warp_sum(p2.x + p2.y);
      const float m1 = warp_sum(p1.x + p1.y) * inv_d, m2 = s2 * inv_d;
      const float sc = rs / den;
      // dE_raw = rstd * (g - mean(g) - xhat * mean(g*xhat)), dS = dE_raw / den:  dS = sc*g + kx*x + k0
      const float kx = -sc * m2 * rs, k0 = -sc * (m1 + m2 * nmr);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const int k = lane + 32 * i;
        if (FULL || k < d8) {
          const int fA = 2 * k + p, fB = 2 * k + 1 - p;
          sg4[fA] = f4_fmas(sg4[fA], sc, f4_fmass(sx4[fA], kx, k0));
          sg4[fB] = f4_fmas(sg4[fB], sc, f4_fmass(sx4[fB], kx, k0));
        }
      }
      // d(den) = -sum_j dE_raw_j E_raw_j / den = -(s2 * eps * rstd^2) / den ; clamp_min passes gradient where wsum >= 1e-6
      if (lane == 0) s_dw[t] = ws >= 1e-6f ? -(s2 * a.eps * rs * rs) / den : 0.f;
      __syncwarp();  // every lane is done with the E_raw row
      if (t + 8 < T && lane == 0) {
        rs_mbar_expect_tx(barx, row_bytes);
        rs_bulk_g2s(rs_smem_u32(s_xw), a.E_raw + (r + 8) * d, row_bytes, barx);
      }
    }
    phg ^= 1u;
    // the warps' dgamma, then dbeta, partials of this sample -> column owners (through the idle E_raw buffers)
    __syncwarp();
    float4* sxw4 = reinterpret_cast<float4*>(s_xw);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (FULL || k < d8) {
        sxw4[2 * k + p] = dgam[i][0];
        sxw4[2 * k + 1 - p] = dgam[i][1];
      }
    }
    __syncthreads();  // also: dS, d(den) and t_hat of every query row are in shared memory
    if ((int)threadIdx.x < d4) {
      float4 c = s_col4[threadIdx.x];
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) f4_add(c, reinterpret_cast<const float4*>(s_x + (size_t)ww * d)[threadIdx.x]);
      s_col4[threadIdx.x] = c;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (FULL || k < d8) {
        sxw4[2 * k + p] = dbet[i][0];
        sxw4[2 * k + 1 - p] = dbet[i][1];
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < d4) {
      float4 c = s_col4[d4 + threadIdx.x];
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) f4_add(c, reinterpret_cast<const float4*>(s_x + (size_t)ww * d)[threadIdx.x]);
      s_col4[d4 + threadIdx.x] = c;
    }
    // note phase (reads s_g, s_dw and s_th only; no barrier until the next sample's)
    for (int n0 = nb; n0 < ne; n0 += 8) {
      const bool nv = g < ne - n0;  // this lane's note exists
      const float tn = nv ? __ldg(a.tau + n0 + g) : 0.f;
      // lane (g, tl) ends up with dV'_n in acc[j][0..1] and Q_n in acc[j][2..3] of note n0 + g at columns 8 (nt0 + j) + 2 tl, + 1:
      // its V' values are requested now and used after the products
      const float* vrow = a.Vp + (size_t)(n0 + (nv ? g : 0)) * a.ldv + 8 * nt0 + 2 * tl;
      float2 vv[NTW];
#pragma unroll
      for (int j = 0; j < NTW; ++j)
        vv[j] = (nv && j < nlive) ? __ldg(reinterpret_cast<const float2*>(vrow + 8 * j)) : make_float2(0.f, 0.f);
      float acc[NTW][4];
#pragma unroll
      for (int j = 0; j < NTW; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
      float scl = 0.f;  // sum_t c_nt d(den_t) over this lane's (note, times)
#pragma unroll
      for (int ks = 0; ks < POOL_TB / 8; ++ks) {
        if (8 * ks < T) {
          const int t0 = 8 * ks + tl, t1 = t0 + 4;
          const int t0c = min(t0, T - 1), t1c = min(t1, T - 1);  // rows past T: weight 0 times a finite dS row
          float w0 = 0.f, c0 = 0.f, w1 = 0.f, c1 = 0.f;
          if (nv && t0 < T) {
            const float rr = fmaxf(s_th[t0] - tn, 0.f) * inv_sigma;
            w0 = expf(-(rr * rr));
            c0 = w0 * 2.f * rr * rr;
          }
          if (nv && t1 < T) {
            const float rr = fmaxf(s_th[t1] - tn, 0.f) * inv_sigma;
            w1 = expf(-(rr * rr));
            c1 = w1 * 2.f * rr * rr;
          }
          scl = fmaf(c0, s_dw[t0c], fmaf(c1, s_dw[t1c], scl));
          uint32_t ah[4], al[4];
          tf32_split(w0, ah[0], al[0]);
          tf32_split(c0, ah[1], al[1]);
          tf32_split(w1, ah[2], al[2]);
          tf32_split(c1, ah[3], al[3]);
          const float* r0 = s_g + (size_t)t0c * ldg + 8 * nt0 + g;
          const float* r1 = s_g + (size_t)t1c * ldg + 8 * nt0 + g;
          if (fullw) note_tiles<NTW, false>(acc, ah, al, r0, r1, NTW);
          else note_tiles<NTW, true>(acc, ah, al, r0, r1, nlive);
        }
      }
      if (w == 0) dls += (double)scl;  // (every warp holds the same A fragments)
      if (nv) {
        float* drow = a.dVp + (size_t)(n0 + g) * a.lddv + 8 * nt0 + 2 * tl;
        float qv = 0.f;
#pragma unroll
        for (int j = 0; j < NTW; ++j)
          if (j < nlive) {
            qv = fmaf(acc[j][2], vv[j].x, fmaf(acc[j][3], vv[j].y, qv));
            *reinterpret_cast<float2*>(drow + 8 * j) = make_float2(acc[j][0], acc[j][1]);
          }
        dls += (double)qv;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) {
    const float v = reinterpret_cast<const float*>(s_col4)[i];
    atomicAdd(i < d ? a.dgamma + i : a.dbeta + (i - d), v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dls += __shfl_xor_sync(0xffffffffu, dls, o);
  if (lane == 0) s_redd[w] = dls;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int ww = 0; ww < 8; ++ww) t += s_redd[ww];
    atomicAdd(a.dlog_sigma, t);
  }
}

static inline size_t bwd_mma_smem(int T, int d) { return ((size_t)T * (d + 8) + (size_t)10 * d) * sizeof(float); }
template <int NC, bool FULL>
static void launch_bwd_mma2(const PoolArgs& a, cudaStream_t st) {
  const size_t smem = bwd_mma_smem(a.T, a.d);
  static size_t smem_set = 0;
  if (smem + 4096 > 48 * 1024 && smem > smem_set) {
    cudaFuncSetAttribute(recavg_bwd_mma_kernel<NC, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  recavg_bwd_mma_kernel<NC, FULL><<<resident_grid((const void*)recavg_bwd_mma_kernel<NC, FULL>, 256, smem, a.B, 2), 256, smem, st>>>(a);
}
template <int NC>
static void launch_bwd_mma(const PoolArgs& a, cudaStream_t st) {
  if ((a.d >> 3) == 32 * NC) launch_bwd_mma2<NC, true>(a, st);
  else launch_bwd_mma2<NC, false>(a, st);
}

static int pool_geometry(int d, int& nch, int& threads) {
  if (d <= 0 || (d & 3)) return -1;
  const int d4 = d >> 2;
  if (d4 <= 256) nch = 1;
  else if (d4 <= 512) nch = 2;
  else if (d4 <= 1024) nch = 4;
  else return -1;
  threads = ((ceil_div(d4, nch) + 31) / 32) * 32;
  return 0;
}

extern "C" int immtsf_recavg_pool_fwd(const float* Vp, int ldv, const float* tau_flat, const int32_t* offsets,
                                      const float* t_hat, int t_hat_bstride, const float* log_sigma,
                                      const float* gamma, const float* beta, int B, int T, int d, int N_max, float eps,
                                      uint32_t drop_thr, uint64_t seed, float* E_drop, float* E_raw, float* mean,
                                      float* rstd, float* wsum, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(Vp && tau_flat && offsets && t_hat && log_sigma && gamma && beta && E_drop, "recavg_pool_fwd: null pointer");
  int nch, threads;
  IMMTSF_REQUIRE(pool_geometry(d, nch, threads) == 0, "recavg_pool_fwd: d=%d must be a multiple of 4 and <= 4096", d);
  IMMTSF_REQUIRE((ldv & 3) == 0 && ((uintptr_t)Vp & 15) == 0, "recavg_pool_fwd: Vp must be 16B aligned with ldv %% 4 == 0");
  PoolArgs a = {};
  a.Vp = Vp; a.ldv = ldv; a.tau = tau_flat; a.offsets = offsets; a.t_hat = t_hat; a.t_bstride = t_hat_bstride;
  a.log_sigma = log_sigma; a.gamma = gamma; a.beta = beta; a.B = B; a.T = T; a.d = d; a.eps = eps;
  a.thr = drop_thr; a.seed = make_seed(seed); a.E_drop = E_drop; a.E_raw = E_raw; a.mean = mean; a.rstd = rstd; a.wsum = wsum;
  cudaStream_t st = (cudaStream_t)stream;
  // Short segments (Time-IMM: a handful of notes per window): one warp per query time, the 8 warps of a CTA share
  // the segment through L1.  Long segments: the CTA-tile kernel streams each V' row once per 8 query times.
  const int nc = N_max <= 32 ? rowwarp_nc(d) : 0;
  if (nc > 0 && ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0 && ((uintptr_t)E_drop & 15) == 0 &&
      (E_raw == nullptr || ((uintptr_t)E_raw & 15) == 0)) {
    // query times per warp: 3 when the row fits comfortably in registers (d <= 768) and T is long enough
    const int tpw = (nc <= 3 && T > 16) ? 3 : (T > 8 && nc <= 3 ? 2 : 1);
    dim3 gridw(ceil_div(T, 8 * tpw), B);
    // default: the TMA-staged kernel (IMMTSF_RECAVG_TMA=0 keeps the register version for A/B runs)
    static const int use_tma = []() { const char* e = getenv("IMMTSF_RECAVG_TMA"); return !(e && e[0] == '0'); }();
    const int RS = N_max < 16 ? N_max : 16;
    const size_t smem_s = (size_t)RS * d * sizeof(float);
    if (use_tma && ((uintptr_t)Vp & 15) == 0 && (ldv & 3) == 0 && (d & 7) == 0) {
      int tpw_s = (nc <= 3 && T > 16) ? 3 : 1;  // measured (profiles/r1_sweep_hbm_v4.json): T 24: 3 > 1 > 2; T 16: 1 > 2 > 3
      static const int tpw_env = []() { const char* e = getenv("IMMTSF_RECAVG_TPW"); return e ? atoi(e) : 0; }();
      if ((tpw_env == 1 || tpw_env == 3) && nc <= 3) tpw_s = tpw_env;
      if (nc == 1) launch_fwd_s<1>(a, tpw_s, T, B, RS, smem_s, st);
      else if (nc == 2) launch_fwd_s<2>(a, tpw_s, T, B, RS, smem_s, st);
      else if (nc == 3) launch_fwd_s<3>(a, tpw_s, T, B, RS, smem_s, st);
      else launch_fwd_s<4>(a, 1, T, B, RS, smem_s, st);
      IMMTSF_CHECK_LAUNCH("recavg_pool_fwd_s");
      return IMMTSF_OK;
    }
#define FWD_W(NCV)                                                                    \
  do {                                                                                \
    if (tpw == 3) recavg_pool_fwd_w_kernel<NCV, 3><<<gridw, 256, 0, st>>>(a);         \
    else if (tpw == 2) recavg_pool_fwd_w_kernel<NCV, 2><<<gridw, 256, 0, st>>>(a);    \
    else recavg_pool_fwd_w_kernel<NCV, 1><<<gridw, 256, 0, st>>>(a);                  \
  } while (0)
    if (nc == 1) FWD_W(1);
    else if (nc == 2) FWD_W(2);
    else if (nc == 3) FWD_W(3);
    else recavg_pool_fwd_w_kernel<4, 1><<<gridw, 256, 0, st>>>(a);
#undef FWD_W
    IMMTSF_CHECK_LAUNCH("recavg_pool_fwd_w");
    return IMMTSF_OK;
  }
  const int TT = 8 / nch;
  dim3 grid(ceil_div(T, TT), B);
  if (nch == 1) recavg_pool_fwd_kernel<1><<<grid, threads, 0, st>>>(a);
  else if (nch == 2) recavg_pool_fwd_kernel<2><<<grid, threads, 0, st>>>(a);
  else recavg_pool_fwd_kernel<4><<<grid, threads, 0, st>>>(a);
  IMMTSF_CHECK_LAUNCH("recavg_pool_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_recavg_pool_bwd(const float* dE_drop, const float* E_raw, const float* mean, const float* rstd,
                                      const float* wsum, const float* Vp, int ldv, const float* tau_flat,
                                      const int32_t* offsets, const float* t_hat, int t_hat_bstride,
                                      const float* log_sigma, const float* gamma, int B, int T, int d, int N_max,
                                      uint32_t drop_thr, uint64_t seed, float* dS, float* dVp, int lddv, float* dgamma,
                                      float* dbeta, double* dlog_sigma, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dE_drop && E_raw && mean && rstd && wsum && Vp && tau_flat && offsets && t_hat && log_sigma && gamma &&
                     dS && dVp && dgamma && dbeta && dlog_sigma, "recavg_pool_bwd: null pointer");
  int nch, threads;
  IMMTSF_REQUIRE(pool_geometry(d, nch, threads) == 0, "recavg_pool_bwd: d=%d must be a multiple of 4 and <= 4096", d);
  IMMTSF_REQUIRE((ldv & 3) == 0 && (lddv & 3) == 0 && ((uintptr_t)Vp & 15) == 0 && ((uintptr_t)dVp & 15) == 0 &&
                     ((uintptr_t)dS & 15) == 0, "recavg_pool_bwd: Vp/dVp/dS must be 16B aligned with ld %% 4 == 0");
  IMMTSF_REQUIRE(N_max >= 1, "recavg_pool_bwd: N_max must be >= 1");
  PoolArgs a = {};
  a.Vp = Vp; a.ldv = ldv; a.tau = tau_flat; a.offsets = offsets; a.t_hat = t_hat; a.t_bstride = t_hat_bstride;
  a.log_sigma = log_sigma; a.gamma = gamma; a.B = B; a.T = T; a.d = d; a.eps = 1e-5f;
  a.thr = drop_thr; a.seed = make_seed(seed); a.E_raw = const_cast<float*>(E_raw); a.mean = const_cast<float*>(mean);
  a.rstd = const_cast<float*>(rstd); a.wsum = const_cast<float*>(wsum);
  a.dE_drop = dE_drop; a.dVp = dVp; a.lddv = lddv; a.dgamma = dgamma; a.dbeta = dbeta; a.dlog_sigma = dlog_sigma;
  a.dS = dS; a.N_max = N_max;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = rowwarp_nc(d);
  // Short prediction windows and segments (N_max <= 32: at N <= 64, T 28 the two-kernel path measured 272 us against 339 us,
  // profiles/r1_sweep_hbm_v6_fused.json): one launch, dS stays in shared memory (IMMTSF_RECAVG_FUSED_BWD=0 keeps the
  // two-kernel path).
  const char* fused_env = getenv("IMMTSF_RECAVG_FUSED_BWD");  // read per call: tests A/B the two paths inside one process
  const char* tma_env = getenv("IMMTSF_RECAVG_TMA");  // =0 means "no bulk-copy kernels at all": the fused kernel is one
  const int fused = (tma_env && tma_env[0] == '0') ? 0 : (fused_env ? atoi(fused_env) : IMMTSF_RECAVG_FUSED_BWD_DEFAULT);
  if (fused && nc > 0 && T <= POOL_TB && N_max <= 32 && bwd_mma_smem(T, d) <= 108 * 1024 && ((uintptr_t)gamma & 15) == 0 &&
      ((uintptr_t)dE_drop & 15) == 0 && ((uintptr_t)E_raw & 15) == 0) {
    if (nc == 1) launch_bwd_mma<1>(a, st);
    else if (nc == 2) launch_bwd_mma<2>(a, st);
    else if (nc == 3) launch_bwd_mma<3>(a, st);
    else launch_bwd_mma<4>(a, st);
    IMMTSF_CHECK_LAUNCH("recavg_bwd_mma");
    return IMMTSF_OK;
  }
  if (nc > 0 && ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)dE_drop & 15) == 0 && ((uintptr_t)E_raw & 15) == 0) {
    const int want = ceil_div(B * T, 4);
    static const int rows_tma = []() { const char* e = getenv("IMMTSF_RECAVG_TMA"); return !(e && e[0] == '0'); }();
    if (rows_tma) {
      if (nc == 1) launch_rows_s<1>(a, want, st);
      else if (nc == 2) launch_rows_s<2>(a, want, st);
      else if (nc == 3) launch_rows_s<3>(a, want, st);
      else launch_rows_s<4>(a, want, st);
    } else {
#define ROWS_W(NCV) recavg_bwd_rows_w_kernel<NCV><<<resident_grid((const void*)recavg_bwd_rows_w_kernel<NCV>, 128, 0, want, 4), 128, 0, st>>>(a)
    if (nc == 1) ROWS_W(1);
    else if (nc == 2) ROWS_W(2);
    else if (nc == 3) ROWS_W(3);
    else ROWS_W(4);
#undef ROWS_W
    }
    IMMTSF_CHECK_LAUNCH("recavg_bwd_rows_w");
  } else {
    const int TT = 8 / nch;
    dim3 grid1(ceil_div(T, TT), B);
    if (nch == 1) recavg_bwd_rows_kernel<1><<<grid1, threads, 0, st>>>(a);
    else if (nch == 2) recavg_bwd_rows_kernel<2><<<grid1, threads, 0, st>>>(a);
    else recavg_bwd_rows_kernel<4><<<grid1, threads, 0, st>>>(a);
    IMMTSF_CHECK_LAUNCH("recavg_bwd_rows");
  }
  dim3 grid2(ceil_div(N_max, 8 / nch), B);
  static const int notes_tma = []() { const char* e = getenv("IMMTSF_RECAVG_TMA"); return !(e && e[0] == '0'); }();
  if (nch == 1 && notes_tma) {
    int TB = T < POOL_TB ? T : POOL_TB;
    while (TB > 1 && (size_t)TB * d * sizeof(float) > 96 * 1024) --TB;
    const size_t smem = (size_t)TB * d * sizeof(float);
    static size_t smem_set = 0;
    if (smem + 4096 > 48 * 1024 && smem > smem_set) {
      cudaFuncSetAttribute(recavg_bwd_notes_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      smem_set = smem;
    }
    recavg_bwd_notes_s_kernel<<<grid2, threads, smem, st>>>(a, TB);
    IMMTSF_CHECK_LAUNCH("recavg_bwd_notes_s");
    return IMMTSF_OK;
  }
  if (nch == 1) recavg_bwd_notes_kernel<1><<<grid2, threads, 0, st>>>(a);
  else if (nch == 2) recavg_bwd_notes_kernel<2><<<grid2, threads, 0, st>>>(a);
  else recavg_bwd_notes_kernel<4><<<grid2, threads, 0, st>>>(a);
  IMMTSF_CHECK_LAUNCH("recavg_bwd_notes");
  return IMMTSF_OK;
}
