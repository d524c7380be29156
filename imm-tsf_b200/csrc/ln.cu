// Residual + LayerNorm over d + dropout, forward and backward
// (fusions/TTF_T2V_XAttn.py:171-179: zero the attention output of no-note
// samples, add the learned query, LayerNorm, dropout).
//   z = (valid[row / rows_per_sample] ? x : 0) + res ;  y = dropout(LN(z))
// CTA owns TT rows at a time (grid-stride over row tiles), threads own float4
// column groups, statistics are block reductions over the register tile.
// Backward accumulates dgamma / dbeta / dres per CTA in registers across its
// row tiles and issues one atomicAdd per owned column at the end.
#include "rowwarp.cuh"
#include "../../include/immtsf.h"

struct LnArgs {
  const float* x; int ldx; const float* xbias; const float* res; const uint8_t* valid; int rps;
  const float* gamma; const float* beta; int R, d; float eps; uint32_t thr; SeedArg seed; uint32_t site;
  float* y; float* mean; float* rstd;
  const float* dy; float* dx; float* dres; float* dgamma; float* dbeta;
};

template <int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const LnArgs a) {
  constexpr int TT = 8 / NCH;
  __shared__ float s_red[32 * TT];
  const int d4 = a.d >> 2;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const int ntiles = (a.R + TT - 1) / TT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TT;
    float4 z[TT][NCH];
    float s1[TT], s2[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int r = r0 + t;
      const bool ok = r < a.R;
      const bool v = ok && (a.valid == nullptr || a.valid[r / a.rps] != 0);
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        float4 q = f4_zero();
        if (ok && col4 < d4) {
          if (v) {
            q = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.ldx) + col4);
            if (a.xbias) f4_add(q, __ldg(reinterpret_cast<const float4*>(a.xbias) + col4));
          }
          if (a.res) f4_add(q, __ldg(reinterpret_cast<const float4*>(a.res) + col4));
        }
        z[t][c] = q;
        p += f4_sum(q);
      }
      s1[t] = p;
    }
    block_sum_multi<TT>(s1, s_red);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const float mu = s1[t] / (float)a.d;
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        if (col4 < d4) {
          const float dx = z[t][c].x - mu, dy = z[t][c].y - mu, dz = z[t][c].z - mu, dw = z[t][c].w - mu;
          p += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
      }
      s2[t] = p;
    }
    block_sum_multi<TT>(s2, s_red);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int r = r0 + t;
      if (r >= a.R) continue;
      const float mu = s1[t] / (float)a.d;
      const float rs = 1.f / sqrtf(s2[t] / (float)a.d + a.eps);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        if (col4 >= d4) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + col4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(a.beta) + col4);
        const float4 ks = dropout_scale4(resolve_seed(a.seed), a.site, (uint64_t)r * d4 + col4, a.thr, inv_keep);
        float4 y;
        y.x = ((z[t][c].x - mu) * rs * g.x + be.x) * ks.x;
        y.y = ((z[t][c].y - mu) * rs * g.y + be.y) * ks.y;
        y.z = ((z[t][c].z - mu) * rs * g.z + be.z) * ks.z;
        y.w = ((z[t][c].w - mu) * rs * g.w + be.w) * ks.w;
        reinterpret_cast<float4*>(a.y + (size_t)r * a.d)[col4] = y;
      }
      if (threadIdx.x == 0) {
        if (a.mean) a.mean[r] = mu;
        if (a.rstd) a.rstd[r] = rs;
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnArgs a) {
  constexpr int TT = 8 / NCH;
  __shared__ float s_red[32 * TT];
  const int d4 = a.d >> 2;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const float inv_d = 1.f / (float)a.d;
  const int ntiles = (a.R + TT - 1) / TT;
  float4 dgam[NCH], dbet[NCH], dres[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) { dgam[c] = f4_zero(); dbet[c] = f4_zero(); dres[c] = f4_zero(); }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int r0 = tile * TT;
    float4 g[TT][NCH], xh[TT][NCH];
    float s1[TT], s2[TT], rs[TT];
    bool vrow[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const int r = r0 + t;
      const bool ok = r < a.R;
      vrow[t] = ok && (a.valid == nullptr || a.valid[r / a.rps] != 0);
      const float mu = ok ? a.mean[r] : 0.f;
      rs[t] = ok ? a.rstd[r] : 0.f;
      float p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        g[t][c] = f4_zero();
        xh[t][c] = f4_zero();
        if (ok && col4 < d4) {
          float4 dy = __ldg(reinterpret_cast<const float4*>(a.dy + (size_t)r * a.d) + col4);
          const float4 ks = dropout_scale4(resolve_seed(a.seed), a.site, (uint64_t)r * d4 + col4, a.thr, inv_keep);
          dy.x *= ks.x; dy.y *= ks.y; dy.z *= ks.z; dy.w *= ks.w;
          float4 q = f4_zero();
          if (vrow[t]) {
            q = __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.ldx) + col4);
            if (a.xbias) f4_add(q, __ldg(reinterpret_cast<const float4*>(a.xbias) + col4));
          }
          if (a.res) f4_add(q, __ldg(reinterpret_cast<const float4*>(a.res) + col4));
          float4 h;
          h.x = (q.x - mu) * rs[t]; h.y = (q.y - mu) * rs[t]; h.z = (q.z - mu) * rs[t]; h.w = (q.w - mu) * rs[t];
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
      const int r = r0 + t;
      if (r >= a.R) continue;
      const float m1 = s1[t] * inv_d, m2 = s2[t] * inv_d;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col4 = threadIdx.x + c * blockDim.x;
        if (col4 >= d4) continue;
        float4 o;
        o.x = rs[t] * (g[t][c].x - m1 - xh[t][c].x * m2);
        o.y = rs[t] * (g[t][c].y - m1 - xh[t][c].y * m2);
        o.z = rs[t] * (g[t][c].z - m1 - xh[t][c].z * m2);
        o.w = rs[t] * (g[t][c].w - m1 - xh[t][c].w * m2);
        f4_add(dres[c], o);  // dz flows to the residual for every row
        reinterpret_cast<float4*>(a.dx + (size_t)r * a.d)[col4] = vrow[t] ? o : f4_zero();
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col4 = threadIdx.x + c * blockDim.x;
    if (col4 < d4) {
      float* pg = a.dgamma + col4 * 4;
      float* pb = a.dbeta + col4 * 4;
      atomicAdd(pg + 0, dgam[c].x); atomicAdd(pg + 1, dgam[c].y); atomicAdd(pg + 2, dgam[c].z); atomicAdd(pg + 3, dgam[c].w);
      atomicAdd(pb + 0, dbet[c].x); atomicAdd(pb + 1, dbet[c].y); atomicAdd(pb + 2, dbet[c].z); atomicAdd(pb + 3, dbet[c].w);
      if (a.dres) {
        float* pr = a.dres + col4 * 4;
        atomicAdd(pr + 0, dres[c].x); atomicAdd(pr + 1, dres[c].y); atomicAdd(pr + 2, dres[c].z); atomicAdd(pr + 3, dres[c].w);
      }
    }
  }
}

// ------------------------------------------------------------------ warp-per-row variants (d % 8 == 0, d <= 1024)
template <int NC>
__global__ void __launch_bounds__(256) ln_fwd_w_kernel(const LnArgs a) {
  const int d8 = a.d >> 3, lane = threadIdx.x & 31;
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)a.d;
  const uint64_t seed = resolve_seed(a.seed);
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < a.R; r += nw) {
    const bool v = a.valid == nullptr || a.valid[r / a.rps] != 0;
    float z[NC][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      zero8(z[i]);
      if (k < d8) {
        if (v) {
          load8(a.x + (size_t)r * a.ldx, k, z[i]);
          if (a.xbias) { float t[8]; load8(a.xbias, k, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) z[i][e] += t[e]; }
        }
        if (a.res) { float t[8]; load8(a.res, k, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) z[i][e] += t[e]; }
#pragma unroll
        for (int e = 0; e < 8; ++e) s += z[i][e];
      }
    }
    const float mu = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (lane + 32 * i < d8)
#pragma unroll
        for (int e = 0; e < 8; ++e) q = fmaf(z[i][e] - mu, z[i][e] - mu, q);
    const float rs = 1.f / sqrtf(warp_sum(q) * inv_d + a.eps);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float g[8], be[8], ks[8], y[8];
        load8(a.gamma, k, g);
        load8(a.beta, k, be);
        dropout_scale8(seed, a.site, (uint64_t)r * d8 + k, a.thr, inv_keep, ks);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = ((z[i][e] - mu) * rs * g[e] + be[e]) * ks[e];
        store8(a.y + (size_t)r * a.d, k, y);
      }
    }
    if (lane == 0) {
      if (a.mean) a.mean[r] = mu;
      if (a.rstd) a.rstd[r] = rs;
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(128) ln_bwd_w_kernel(const LnArgs a) {
  __shared__ float s_acc[3 * 1024];  // dgamma | dbeta | dres of this CTA
  const int d8 = a.d >> 3, lane = threadIdx.x & 31;
  const float inv_keep = inv_keep_from_thr(a.thr), inv_d = 1.f / (float)a.d;
  const uint64_t seed = resolve_seed(a.seed);
  for (int i = threadIdx.x; i < 3 * a.d; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float ga[NC][8], dgam[NC][8], dbet[NC][8], dres[NC][8];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    zero8(ga[i]); zero8(dgam[i]); zero8(dbet[i]); zero8(dres[i]);
    if (lane + 32 * i < d8) load8(a.gamma, lane + 32 * i, ga[i]);
  }
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw; r < a.R; r += nw) {
    const bool v = a.valid == nullptr || a.valid[r / a.rps] != 0;
    const float mu = a.mean[r], rs = a.rstd[r];
    float g[NC][8], h[NC][8];
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      zero8(g[i]); zero8(h[i]);
      if (k < d8) {
        float dy[8], ks[8], q[8];
        load8(a.dy + (size_t)r * a.d, k, dy);
        dropout_scale8(seed, a.site, (uint64_t)r * d8 + k, a.thr, inv_keep, ks);
        zero8(q);
        if (v) {
          load8(a.x + (size_t)r * a.ldx, k, q);
          if (a.xbias) { float t[8]; load8(a.xbias, k, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) q[e] += t[e]; }
        }
        if (a.res) { float t[8]; load8(a.res, k, t);
#pragma unroll
          for (int e = 0; e < 8; ++e) q[e] += t[e]; }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dye = dy[e] * ks[e];
          const float he = (q[e] - mu) * rs;
          dgam[i][e] = fmaf(dye, he, dgam[i][e]);
          dbet[i][e] += dye;
          const float gg = dye * ga[i][e];
          g[i][e] = gg;
          h[i][e] = he;
          p1 += gg;
          p2 = fmaf(gg, he, p2);
        }
      }
    }
    const float m1 = warp_sum(p1) * inv_d, m2 = warp_sum(p2) * inv_d;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int k = lane + 32 * i;
      if (k < d8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = rs * (g[i][e] - m1 - h[i][e] * m2);
          dres[i][e] += o[e];  // dz flows to the residual for every row
          if (!v) o[e] = 0.f;
        }
        store8(a.dx + (size_t)r * a.d, k, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int k = lane + 32 * i;
    if (k < d8)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_acc[8 * k + e], dgam[i][e]);
        atomicAdd(&s_acc[a.d + 8 * k + e], dbet[i][e]);
        atomicAdd(&s_acc[2 * a.d + 8 * k + e], dres[i][e]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
    atomicAdd(a.dgamma + i, s_acc[i]);
    atomicAdd(a.dbeta + i, s_acc[a.d + i]);
    if (a.dres) atomicAdd(a.dres + i, s_acc[2 * a.d + i]);
  }
}

static int ln_geometry(int d, int& nch, int& threads) {
  if (d <= 0 || (d & 3)) return -1;
  const int d4 = d >> 2;
  if (d4 <= 256) nch = 1;
  else if (d4 <= 512) nch = 2;
  else if (d4 <= 1024) nch = 4;
  else return -1;
  threads = ((ceil_div(d4, nch) + 31) / 32) * 32;
  return 0;
}

extern "C" int immtsf_ln_fwd(const float* x, int ldx, const float* xbias, const float* res, const uint8_t* valid, int rows_per_sample,
                             const float* gamma, const float* beta, int R, int d, float eps, uint32_t drop_thr,
                             uint64_t seed, uint32_t site, float* y, float* mean, float* rstd, void* stream) {
  if (R == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(x && gamma && beta && y, "ln_fwd: null pointer");
  IMMTSF_REQUIRE(valid == nullptr || rows_per_sample > 0, "ln_fwd: rows_per_sample must be > 0");
  int nch, threads;
  IMMTSF_REQUIRE(ln_geometry(d, nch, threads) == 0, "ln_fwd: d=%d must be a multiple of 4 and <= 4096", d);
  IMMTSF_REQUIRE((ldx & 3) == 0 && ((uintptr_t)x & 15) == 0, "ln_fwd: x must be 16B aligned with ldx %% 4 == 0");
  LnArgs a = {};
  a.x = x; a.ldx = ldx; a.xbias = xbias; a.res = res; a.valid = valid; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
  a.gamma = gamma; a.beta = beta; a.R = R; a.d = d; a.eps = eps; a.thr = drop_thr; a.seed = make_seed(seed); a.site = site;
  a.y = y; a.mean = mean; a.rstd = rstd;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = rowwarp_nc(d);
  if (nc > 0 && (res == nullptr || ((uintptr_t)res & 15) == 0) && (xbias == nullptr || ((uintptr_t)xbias & 15) == 0) &&
      ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)beta & 15) == 0 && ((uintptr_t)y & 15) == 0) {
    int gridw = ceil_div(R, 8);
    if (gridw > 148 * 8) gridw = 148 * 8;
    if (nc == 1) ln_fwd_w_kernel<1><<<gridw, 256, 0, st>>>(a);
    else if (nc == 2) ln_fwd_w_kernel<2><<<gridw, 256, 0, st>>>(a);
    else if (nc == 3) ln_fwd_w_kernel<3><<<gridw, 256, 0, st>>>(a);
    else ln_fwd_w_kernel<4><<<gridw, 256, 0, st>>>(a);
    IMMTSF_CHECK_LAUNCH("ln_fwd_w");
    return IMMTSF_OK;
  }
  const int TT = 8 / nch;
  int grid = ceil_div(R, TT);
  if (grid > 148 * 8) grid = 148 * 8;
  if (nch == 1) ln_fwd_kernel<1><<<grid, threads, 0, st>>>(a);
  else if (nch == 2) ln_fwd_kernel<2><<<grid, threads, 0, st>>>(a);
  else ln_fwd_kernel<4><<<grid, threads, 0, st>>>(a);
  IMMTSF_CHECK_LAUNCH("ln_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_ln_bwd(const float* dy, const float* x, int ldx, const float* xbias, const float* res, const uint8_t* valid,
                             int rows_per_sample, const float* gamma, const float* mean, const float* rstd, int R,
                             int d, uint32_t drop_thr, uint64_t seed, uint32_t site, float* dx, float* dres,
                             float* dgamma, float* dbeta, void* stream) {
  if (R == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "ln_bwd: null pointer");
  int nch, threads;
  IMMTSF_REQUIRE(ln_geometry(d, nch, threads) == 0, "ln_bwd: d=%d must be a multiple of 4 and <= 4096", d);
  IMMTSF_REQUIRE((ldx & 3) == 0 && ((uintptr_t)x & 15) == 0, "ln_bwd: x must be 16B aligned with ldx %% 4 == 0");
  LnArgs a = {};
  a.x = x; a.ldx = ldx; a.xbias = xbias; a.res = res; a.valid = valid; a.rps = rows_per_sample > 0 ? rows_per_sample : 1;
  a.gamma = gamma; a.R = R; a.d = d; a.thr = drop_thr; a.seed = make_seed(seed); a.site = site;
  a.mean = const_cast<float*>(mean); a.rstd = const_cast<float*>(rstd);
  a.dy = dy; a.dx = dx; a.dres = dres; a.dgamma = dgamma; a.dbeta = dbeta;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = rowwarp_nc(d);
  if (nc > 0 && (res == nullptr || ((uintptr_t)res & 15) == 0) && (xbias == nullptr || ((uintptr_t)xbias & 15) == 0) &&
      ((uintptr_t)gamma & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0) {
    const int want = ceil_div(R, 4);
#define LNB_W(NCV) ln_bwd_w_kernel<NCV><<<resident_grid((const void*)ln_bwd_w_kernel<NCV>, 128, 0, want, 4), 128, 0, st>>>(a)
    if (nc == 1) LNB_W(1);
    else if (nc == 2) LNB_W(2);
    else if (nc == 3) LNB_W(3);
    else LNB_W(4);
#undef LNB_W
    IMMTSF_CHECK_LAUNCH("ln_bwd_w");
    return IMMTSF_OK;
  }
  const int TT = 8 / nch;
  int grid = ceil_div(R, TT);
  if (grid > 148 * 3) grid = 148 * 3;
  if (nch == 1) ln_bwd_kernel<1><<<grid, threads, 0, st>>>(a);
  else if (nch == 2) ln_bwd_kernel<2><<<grid, threads, 0, st>>>(a);
  else ln_bwd_kernel<4><<<grid, threads, 0, st>>>(a);
  IMMTSF_CHECK_LAUNCH("ln_bwd");
  return IMMTSF_OK;
}
