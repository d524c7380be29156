// Per-(note, query) Time2Vec attention over each ragged note segment -- SURVEY.md 8f row f3, the semantics of
// fusions/TTF_T2V_XAttn_old.py:119-143 (Time2Vec of the clamped lag max(t_hat - tau, 0) of EVERY (note, query) pair
// enters the keys and values, so the attention is a dense [T_f x N_i] problem per sample).
//
// The reference builds [V ; phi] for B*T*N pairs, runs KV_proj on all of them and lets nn.MultiheadAttention project
// K and V again.  With X_nt = A_n + W_phi phi_nt (A_n = W_a V'_n + b_kv, once per note):
//   score_{h,n,t} = a_{n,h} + g_h . phi_nt          a = A U^T, u_h = W_k[h]^T q_h, g_h = W_phi^T u_h
//   Z_{t,h} = sum_n P~ A_n,  Phi_{t,h} = sum_n P~ phi_nt,  sp_{t,h} = sum_n P~     (P~ = dropout(softmax_n(score)))
// and the head output is W_v[h] (Z + W_phi Phi) + sp b_v[h] -- GEMMs on B*T rows done by the caller.  No vector of
// width d exists per (note, query) pair; sin() is evaluated on the fly (twice in forward, once in backward).
//
// Layout: A [M_alloc, d] (lda), a_sc / da [M_alloc, H], g [H, d_tau]; output rows r = (b*T + t)*H + h:
// Z [B*T*H, d], Phi [B*T*H, d_tau], sp [B*T*H]; probs[(h*T + t) * M_alloc + row] (softmax before dropout).
// Forward: grid (ceil(T/TQ), B); backward: one CTA per sample walking its query tiles (dA / da / the Time2Vec partials
// are owned by that CTA: no atomics, deterministic).
#include "rowtile.cuh"
#include "../../include/immtsf.h"

constexpr int PQ_MAXH = 8;

struct PQArgs {
  const float* A; int lda; const float* a_sc; const float* g; const float* tau; const int32_t* offsets;
  const float* t_hat; int t_bstride;
  const float* w_lin; const float* b_lin; const float* w_per; const float* b_per;
  int B, T, H, d, dt, NM, TQ; size_t M_alloc; uint32_t thr; SeedArg seed;
  float* Z; float* Phi; float* sp; float* probs;
  const float* dZ; const float* dPhi; const float* dsp; const float* probs_in;
  float* dA; int lddA; float* da; float* dpart; float* pt_g; float* ds_g;
};

__device__ __forceinline__ float pq_phi(int k, float dl, float wl, float bl, const float* __restrict__ w_per,
                                        const float* __restrict__ b_per) {
  return k == 0 ? fmaf(wl, dl, bl) : sinf(fmaf(__ldg(w_per + k - 1), dl, __ldg(b_per + k - 1)));
}

// smem: s_dl [TQ][NM] | s_p [TQ*H][NM].  H is a template parameter: with a run-time head count the guarded 8-way
// unrolled head loops were 2/3 of the executed instructions (ncu source page, profiles/r1_ncu_t2vq_v1_summary.txt).
template <int H>
__global__ void __launch_bounds__(256) t2vq_fwd_kernel(const PQArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = a.d, dt = a.dt, NM = a.NM, TQ = a.TQ, T = a.T;
  float* s_dl = smem;
  float* s_p = smem + (size_t)TQ * NM;
  const int b = blockIdx.y, t0 = blockIdx.x * TQ;
  const int tcnt = min(TQ, T - t0), rows = tcnt * H;
  const int nb = a.offsets[b], nn = a.offsets[b + 1] - nb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t row0 = ((size_t)b * T + t0) * H;
  if (nn == 0) {  // no notes: the pooled quantities are 0 (the caller's LayerNorm masks the sample, reference :146-150)
    for (int i = threadIdx.x; i < rows * d; i += blockDim.x) a.Z[row0 * d + i] = 0.f;
    for (int i = threadIdx.x; i < rows * dt; i += blockDim.x) a.Phi[row0 * dt + i] = 0.f;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) a.sp[row0 + i] = 0.f;
    return;
  }
  const float wl = a.w_lin[0], bl = a.b_lin[0];
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  // 0) lags
  for (int i = threadIdx.x; i < tcnt * nn; i += blockDim.x) {
    const int tt = i / nn, n = i % nn;
    s_dl[tt * NM + n] = fmaxf(a.t_hat[(size_t)b * a.t_bstride + t0 + tt] - a.tau[nb + n], 0.f);
  }
  __syncthreads();
  // 1) scores: one warp per (query time, note), lanes over the Time2Vec units, all heads at once
  for (int i = w; i < tcnt * nn; i += nw) {
    const int tt = i / nn, n = i % nn;
    const float dl = s_dl[tt * NM + n];
    float acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
    for (int k = lane; k < dt; k += 32) {
      const float ph = pq_phi(k, dl, wl, bl, a.w_per, a.b_per);
#pragma unroll
      for (int h = 0; h < H; ++h)
        acc[h] = fmaf(__ldg(a.g + (size_t)h * dt + k), ph, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      if (h < H) {
        const float s = warp_sum(acc[h]);
        if (lane == 0) s_p[(tt * H + h) * NM + n] = s + a.a_sc[(size_t)(nb + n) * H + h];
      }
    }
  }
  __syncthreads();
  // 2) softmax over the segment + dropout, one warp per (query time, head) row
  for (int r = w; r < rows; r += nw) {
    const int tt = r / H, h = r % H, t = t0 + tt;
    float* pr = s_p + (size_t)r * NM;
    float mx = -INFINITY;
    for (int n = lane; n < nn; n += 32) mx = fmaxf(mx, pr[n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < nn; n += 32) {
      const float e = expf(pr[n] - mx);
      pr[n] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float spt = 0.f;
    for (int n = lane; n < nn; n += 32) {
      const float p = pr[n] / sum;
      if (a.probs) a.probs[((size_t)h * T + t) * a.M_alloc + nb + n] = p;
      const float pt = p * dropout_scale(seed, IMMTSF_SITE_TTF_ATTN, (((uint64_t)b * T + t) * H + h) * NM + n, a.thr, inv_keep);
      pr[n] = pt;
      spt += pt;
    }
    spt = warp_sum(spt);
    if (lane == 0) a.sp[row0 + r] = spt;
  }
  __syncthreads();
  // 3) Z rows: thread per float4 column, 8 rows of the tile at a time
  const int d4 = d >> 2;
  for (int col4 = threadIdx.x; col4 < d4; col4 += blockDim.x) {
    for (int r0 = 0; r0 < rows; r0 += 8) {
      float4 acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = f4_zero();
      for (int n = 0; n < nn; ++n) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.A + (size_t)(nb + n) * a.lda) + col4);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (r0 + u < rows) f4_fma(acc[u], s_p[(size_t)(r0 + u) * NM + n], v);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < rows) reinterpret_cast<float4*>(a.Z + (row0 + r0 + u) * d)[col4] = acc[u];
    }
  }
  // 4) Phi rows: thread per (query time, unit), all heads
  for (int i = threadIdx.x; i < tcnt * dt; i += blockDim.x) {
    const int tt = i / dt, k = i % dt;
    float acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
    for (int n = 0; n < nn; ++n) {
      const float ph = pq_phi(k, s_dl[tt * NM + n], wl, bl, a.w_per, a.b_per);
#pragma unroll
      for (int h = 0; h < H; ++h)
        acc[h] = fmaf(s_p[(size_t)(tt * H + h) * NM + n], ph, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < H; ++h)
      a.Phi[(row0 + tt * H + h) * dt + k] = acc[h];
  }
}

// Backward, kernel 1 of 2 -- one CTA per (query tile, sample), like the forward.  smem: s_dl [TQ][NM] | s_pt [TQ*H][NM] |
// s_ds [TQ*H][NM].  Per tile: dP~ (dot products + Time2Vec part), dropout + softmax backward, the Time2Vec parameter
// partials / dg of the tile (dpart row b*ntiles + tile, every entry owned by one thread), and P~ / dS written to global
// scratch in the layout of probs for kernel 2.  (The first version walked the tiles of a sample inside ONE CTA to keep dA in
// place: 0.58 waves, 34 % issue utilisation -- profiles/r1_ncu_t2vq_v1_summary.txt.)
template <int H>
__global__ void __launch_bounds__(256) t2vq_bwd_tile_kernel(const PQArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = a.d, dt = a.dt, NM = a.NM, TQ = a.TQ, T = a.T;
  float* s_dl = smem;
  float* s_pt = s_dl + (size_t)TQ * NM;
  float* s_ds = s_pt + (size_t)TQ * H * NM;
  const int b = blockIdx.y, t0 = blockIdx.x * TQ;
  const int tcnt = min(TQ, T - t0), rows = tcnt * H;
  const int nb = a.offsets[b], nn = a.offsets[b + 1] - nb;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* part = a.dpart + ((size_t)b * gridDim.x + blockIdx.x) * (2 + H) * dt;
  if (nn == 0) {
    for (int i = threadIdx.x; i < (2 + H) * dt; i += blockDim.x) part[i] = 0.f;
    return;
  }
  const float wl = a.w_lin[0], bl = a.b_lin[0];
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  const int d4 = d >> 2;
  const size_t row0 = ((size_t)b * T + t0) * H;
  // 0) lags
  for (int i = threadIdx.x; i < tcnt * nn; i += blockDim.x) {
    const int tt = i / nn, n = i % nn;
    s_dl[tt * NM + n] = fmaxf(a.t_hat[(size_t)b * a.t_bstride + t0 + tt] - a.tau[nb + n], 0.f);
  }
  __syncthreads();
  // a) dP~[t,h,n] = dZ[t,h] . A[n] + dPhi[t,h] . phi[n,t] + dsp[t,h]: one warp per (query time, note)
  for (int i = w; i < tcnt * nn; i += nw) {
    const int tt = i / nn, n = i % nn;
    const float dl = s_dl[tt * NM + n];
    float acc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) acc[h] = 0.f;
    const float4* ar = reinterpret_cast<const float4*>(a.A + (size_t)(nb + n) * a.lda);
    for (int c4 = lane; c4 < d4; c4 += 32) {
      const float4 av = __ldg(ar + c4);
#pragma unroll
      for (int h = 0; h < H; ++h)
        acc[h] += f4_dot(av, __ldg(reinterpret_cast<const float4*>(a.dZ + (row0 + tt * H + h) * d) + c4));
    }
    for (int k = lane; k < dt; k += 32) {
      const float ph = pq_phi(k, dl, wl, bl, a.w_per, a.b_per);
#pragma unroll
      for (int h = 0; h < H; ++h)
        acc[h] = fmaf(__ldg(a.dPhi + (row0 + tt * H + h) * dt + k), ph, acc[h]);
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float s = warp_sum(acc[h]);
      if (lane == 0) s_ds[(size_t)(tt * H + h) * NM + n] = s + a.dsp[row0 + tt * H + h];
    }
  }
  __syncthreads();
  // b) dropout + softmax backward per row: s_ds <- dS, s_pt <- P~ ; both also to global for kernel 2
  for (int r = w; r < rows; r += nw) {
    const int tt = r / H, h = r % H, t = t0 + tt;
    float* dsr = s_ds + (size_t)r * NM;
    float* ptr = s_pt + (size_t)r * NM;
    const size_t g0 = ((size_t)h * T + t) * a.M_alloc + nb;
    const float* pg = a.probs_in + g0;
    float D = 0.f;
    for (int n = lane; n < nn; n += 32) {
      const float ks = dropout_scale(seed, IMMTSF_SITE_TTF_ATTN, (((uint64_t)b * T + t) * H + h) * NM + n, a.thr, inv_keep);
      const float p = pg[n];
      const float dp = dsr[n] * ks;
      dsr[n] = dp;
      const float pt = p * ks;
      ptr[n] = pt;
      a.pt_g[g0 + n] = pt;
      D = fmaf(p, dp, D);
    }
    D = warp_sum(D);
    for (int n = lane; n < nn; n += 32) {
      const float ds = pg[n] * (dsr[n] - D);
      dsr[n] = ds;
      a.ds_g[g0 + n] = ds;
    }
  }
  __syncthreads();
  // d) Time2Vec parameter partials and dg of this tile: thread per unit k
  for (int k = threadIdx.x; k < dt; k += blockDim.x) {
    const float wk = k == 0 ? wl : __ldg(a.w_per + k - 1), bk = k == 0 ? bl : __ldg(a.b_per + k - 1);
    float sw = 0.f, sb = 0.f, sg[H], gk[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      sg[h] = 0.f;
      gk[h] = __ldg(a.g + (size_t)h * dt + k);
    }
    for (int tt = 0; tt < tcnt; ++tt) {
      float dph[H];
#pragma unroll
      for (int h = 0; h < H; ++h) dph[h] = __ldg(a.dPhi + (row0 + tt * H + h) * dt + k);
      for (int n = 0; n < nn; ++n) {
        const float dl = s_dl[tt * NM + n];
        const float arg = fmaf(wk, dl, bk);
        float ph = arg, cs = 1.f;
        if (k != 0) sincosf(arg, &ph, &cs);
        float dphi = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float ds = s_ds[(size_t)(tt * H + h) * NM + n];
          dphi = fmaf(s_pt[(size_t)(tt * H + h) * NM + n], dph[h], dphi);
          dphi = fmaf(ds, gk[h], dphi);
          sg[h] = fmaf(ds, ph, sg[h]);
        }
        const float dpre = dphi * cs;
        sw = fmaf(dpre, dl, sw);
        sb += dpre;
      }
    }
    part[k] = sw;
    part[dt + k] = sb;
#pragma unroll
    for (int h = 0; h < H; ++h) part[(size_t)(2 + h) * dt + k] = sg[h];
  }
}

// Backward, kernel 2 of 2 -- dA[n] = sum over the sample's (t, h) rows of P~ dZ and da[n, h] = sum_t dS, from the global
// P~ / dS of kernel 1.  grid (ceil(d/256), B), 256 threads: thread (c = tid % 64, slot = tid / 64) owns float4 column
// blockIdx.x*64 + c of notes slot*4 .. slot*4+3 (+16 per pass); the sample's rows are staged RT at a time.  smem: [RT][NM].
template <int H>
__global__ void __launch_bounds__(256) t2vq_bwd_notes_kernel(const PQArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int d = a.d, NM = a.NM, T = a.T, RT = a.TQ;  // TQ carries the row-tile height here
  const int b = blockIdx.y;
  const int nb = a.offsets[b], nn = a.offsets[b + 1] - nb;
  if (nn == 0) return;
  const int R = T * H, d4 = d >> 2;
  const int c = threadIdx.x & 63, slot = threadIdx.x >> 6;
  const int col4 = blockIdx.x * 64 + c;
  if (blockIdx.x == 0) {  // da: sum over query times in order (deterministic)
    for (int i = threadIdx.x; i < H * nn; i += blockDim.x) {
      const int h = i / nn, n = i % nn;
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc += a.ds_g[((size_t)h * T + t) * a.M_alloc + nb + n];
      a.da[(size_t)(nb + n) * H + h] = acc;
    }
  }
  // row tiles: all threads stage, owners accumulate
  const int npass = (nn + 15) / 16;
  for (int pass = 0; pass < npass; ++pass) {
    const int n0 = pass * 16 + slot * 4;
    float4 acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] = f4_zero();
    for (int r0 = 0; r0 < R; r0 += RT) {
      const int rc = min(RT, R - r0);
      __syncthreads();
      for (int i = threadIdx.x; i < rc * nn; i += blockDim.x) {
        const int rr = i / nn, n = i % nn, r = r0 + rr, t = r / H, h = r % H;
        smem[rr * NM + n] = a.pt_g[((size_t)h * T + t) * a.M_alloc + nb + n];
      }
      __syncthreads();
      if (col4 < d4 && n0 < nn) {
        const float4* gz = reinterpret_cast<const float4*>(a.dZ + ((size_t)b * R + r0) * d) + col4;
        for (int rr = 0; rr < rc; ++rr) {
          const float4 go = __ldg(gz + (size_t)rr * d4);
          const float* pr = smem + rr * NM + n0;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (n0 + u < nn) f4_fma(acc[u], pr[u], go);
        }
      }
    }
    if (col4 < d4 && n0 < nn) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (n0 + u < nn) reinterpret_cast<float4*>(a.dA + (size_t)(nb + n0 + u) * a.lddA)[col4] = acc[u];
    }
  }
}

template <int H>
static void pq_launch_fwd(const PQArgs& a, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaFuncSetAttribute(t2vq_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  t2vq_fwd_kernel<H><<<dim3(ceil_div(a.T, a.TQ), a.B), 256, smem, st>>>(a);
}
template <int H>
static void pq_launch_bwd(PQArgs a, size_t smem, int RT, size_t smem2, cudaStream_t st) {
  static size_t smem_set = 0, smem2_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaFuncSetAttribute(t2vq_bwd_tile_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set = smem;
  }
  if (smem2 > 48 * 1024 && smem2 > smem2_set) {
    cudaFuncSetAttribute(t2vq_bwd_notes_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    smem2_set = smem2;
  }
  t2vq_bwd_tile_kernel<H><<<dim3(ceil_div(a.T, a.TQ), a.B), 256, smem, st>>>(a);
  immtsf_count_launch();
  a.TQ = RT;
  t2vq_bwd_notes_kernel<H><<<dim3(ceil_div(a.d, 256), a.B), 256, smem2, st>>>(a);
}

// query times per tile of the backward tile kernel (and so the number of dpart rows per sample)
static int pq_bwd_tq(int H, int N_max) {
  int TQ = 8;
  const size_t plane = (size_t)N_max * sizeof(float);
  while (TQ > 1 && (size_t)TQ * (1 + 2 * H) * plane > 160 * 1024) TQ >>= 1;
  return TQ;
}

static int pq_common(PQArgs& a, const char* who, const float* A, int lda, const float* g, const float* tau_flat,
                     const int32_t* offsets, const float* t_hat, int t_bstride, const float* w_lin, const float* b_lin,
                     const float* w_per, const float* b_per, int B, int T, int H, int d, int d_tau, int N_max, int M_alloc,
                     uint32_t drop_thr, uint64_t seed) {
  IMMTSF_REQUIRE(A && g && tau_flat && offsets && t_hat && w_lin && b_lin && w_per && b_per, "%s: null pointer", who);
  IMMTSF_REQUIRE(H >= 1 && H <= PQ_MAXH, "%s: 1 <= n_heads <= %d supported (got %d)", who, PQ_MAXH, H);
  IMMTSF_REQUIRE(d >= 4 && (d & 3) == 0 && (lda & 3) == 0 && lda >= d, "%s: d and lda must be multiples of 4 (d=%d lda=%d)", who, d, lda);
  IMMTSF_REQUIRE(d_tau > 1, "%s: d_tau must be > 1 (TTF_T2V_XAttn_old.py:14)", who);
  IMMTSF_REQUIRE(((uintptr_t)A & 15) == 0, "%s: A must be 16B aligned", who);
  IMMTSF_REQUIRE(N_max >= 1 && T >= 1 && M_alloc >= 1 && t_bstride >= 0, "%s: N_max, T, M_alloc must be >= 1", who);
  a.A = A; a.lda = lda; a.g = g; a.tau = tau_flat; a.offsets = offsets; a.t_hat = t_hat; a.t_bstride = t_bstride;
  a.w_lin = w_lin; a.b_lin = b_lin; a.w_per = w_per; a.b_per = b_per;
  a.B = B; a.T = T; a.H = H; a.d = d; a.dt = d_tau; a.NM = N_max; a.M_alloc = (size_t)M_alloc; a.thr = drop_thr; a.seed = make_seed(seed);
  return IMMTSF_OK;
}

extern "C" int immtsf_t2vq_attn_fwd(const float* A, int lda, const float* a_sc, const float* g, const float* tau_flat,
                                    const int32_t* offsets, const float* t_hat, int t_hat_bstride, const float* w_lin,
                                    const float* b_lin, const float* w_per, const float* b_per, int B, int T, int H, int d,
                                    int d_tau, int N_max, int M_alloc, uint32_t drop_thr, uint64_t seed, float* Z, float* Phi,
                                    float* sp, float* probs, void* stream) {
  if (B == 0) return IMMTSF_OK;
  PQArgs a = {};
  const int rc = pq_common(a, "t2vq_attn_fwd", A, lda, g, tau_flat, offsets, t_hat, t_hat_bstride, w_lin, b_lin, w_per, b_per, B, T, H, d,
                           d_tau, N_max, M_alloc, drop_thr, seed);
  if (rc != IMMTSF_OK) return rc;
  IMMTSF_REQUIRE(a_sc && Z && Phi && sp, "t2vq_attn_fwd: null pointer");
  IMMTSF_REQUIRE(((uintptr_t)Z & 15) == 0, "t2vq_attn_fwd: Z must be 16B aligned");
  IMMTSF_REQUIRE(B <= 65535, "t2vq_attn_fwd: B <= 65535");
  a.a_sc = a_sc; a.Z = Z; a.Phi = Phi; a.sp = sp; a.probs = probs;
  int TQ = 8;
  const size_t plane = (size_t)N_max * sizeof(float);
  while (TQ > 1 && (size_t)TQ * (1 + H) * plane > 160 * 1024) TQ >>= 1;
  const size_t smem = (size_t)TQ * (1 + H) * plane;
  if (smem > 200 * 1024) { immtsf_set_error("t2vq_attn_fwd: H*N_max=%d too large for shared memory", H * N_max); return IMMTSF_ERR_UNSUPPORTED; }
  a.TQ = TQ;
  switch (H) {
    case 1: pq_launch_fwd<1>(a, smem, (cudaStream_t)stream); break;
    case 2: pq_launch_fwd<2>(a, smem, (cudaStream_t)stream); break;
    case 3: pq_launch_fwd<3>(a, smem, (cudaStream_t)stream); break;
    case 4: pq_launch_fwd<4>(a, smem, (cudaStream_t)stream); break;
    case 5: pq_launch_fwd<5>(a, smem, (cudaStream_t)stream); break;
    case 6: pq_launch_fwd<6>(a, smem, (cudaStream_t)stream); break;
    case 7: pq_launch_fwd<7>(a, smem, (cudaStream_t)stream); break;
    default: pq_launch_fwd<8>(a, smem, (cudaStream_t)stream); break;
  }
  IMMTSF_CHECK_LAUNCH("t2vq_attn_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_t2vq_bwd_tiles(int T, int H, int N_max) {
  if (T < 1 || H < 1 || N_max < 1) return 0;
  return ceil_div(T, pq_bwd_tq(H, N_max));
}

extern "C" size_t immtsf_t2vq_bwd_workspace_bytes(int T, int H, int M_alloc) {
  return (size_t)2 * (size_t)H * (size_t)T * (size_t)M_alloc * sizeof(float);
}

extern "C" int immtsf_t2vq_attn_bwd(const float* dZ, const float* dPhi, const float* dsp, const float* A, int lda, const float* g,
                                    const float* probs, const float* tau_flat, const int32_t* offsets, const float* t_hat,
                                    int t_hat_bstride, const float* w_lin, const float* b_lin, const float* w_per,
                                    const float* b_per, int B, int T, int H, int d, int d_tau, int N_max, int M_alloc,
                                    uint32_t drop_thr, uint64_t seed, float* dA, int lddA, float* da, float* dpart,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0) return IMMTSF_OK;
  PQArgs a = {};
  const int rc = pq_common(a, "t2vq_attn_bwd", A, lda, g, tau_flat, offsets, t_hat, t_hat_bstride, w_lin, b_lin, w_per, b_per, B, T, H, d,
                           d_tau, N_max, M_alloc, drop_thr, seed);
  if (rc != IMMTSF_OK) return rc;
  IMMTSF_REQUIRE(dZ && dPhi && dsp && probs && dA && da && dpart && workspace, "t2vq_attn_bwd: null pointer");
  IMMTSF_REQUIRE(((uintptr_t)dZ & 15) == 0 && ((uintptr_t)dA & 15) == 0 && (lddA & 3) == 0 && lddA >= d,
                 "t2vq_attn_bwd: dZ / dA must be 16B aligned, lddA a multiple of 4");
  IMMTSF_REQUIRE(workspace_bytes >= immtsf_t2vq_bwd_workspace_bytes(T, H, M_alloc), "t2vq_attn_bwd: workspace too small");
  IMMTSF_REQUIRE(B <= 65535, "t2vq_attn_bwd: B <= 65535");
  a.dZ = dZ; a.dPhi = dPhi; a.dsp = dsp; a.probs_in = probs; a.dA = dA; a.lddA = lddA; a.da = da; a.dpart = dpart;
  a.pt_g = (float*)workspace;
  a.ds_g = a.pt_g + (size_t)H * T * M_alloc;
  const size_t plane = (size_t)N_max * sizeof(float);
  a.TQ = pq_bwd_tq(H, N_max);
  const size_t smem = (size_t)a.TQ * (1 + 2 * H) * plane;
  int RT = 32;
  while (RT > 1 && (size_t)RT * plane > 96 * 1024) RT >>= 1;
  const size_t smem2 = (size_t)RT * plane;
  if (smem > 200 * 1024 || smem2 > 200 * 1024) { immtsf_set_error("t2vq_attn_bwd: H*N_max=%d too large for shared memory", H * N_max); return IMMTSF_ERR_UNSUPPORTED; }
  switch (H) {
    case 1: pq_launch_bwd<1>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 2: pq_launch_bwd<2>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 3: pq_launch_bwd<3>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 4: pq_launch_bwd<4>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 5: pq_launch_bwd<5>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 6: pq_launch_bwd<6>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    case 7: pq_launch_bwd<7>(a, smem, RT, smem2, (cudaStream_t)stream); break;
    default: pq_launch_bwd<8>(a, smem, RT, smem2, (cudaStream_t)stream); break;
  }
  IMMTSF_CHECK_LAUNCH("t2vq_attn_bwd");
  return IMMTSF_OK;
}
