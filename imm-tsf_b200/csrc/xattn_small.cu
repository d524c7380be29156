// K6 fast path: MMF_XAttn_Add attention core for T <= 32 query/key times
// (fusions/MMF_XAttn_Add.py:73-80 + the softmax/dropout/bmm inside
// nn.MultiheadAttention).  Time-IMM prediction windows are 7..28 steps, so the
// whole T x T score matrix of one (sample, head) lives in registers/shared
// memory of ONE CTA and q, k, v, o are each read or written exactly once:
//
//   S = scale * Q K^T : Q and K are staged through shared memory in chunks of
//       XS_DC columns; each of the 8 warps owns an XS_DC/8-wide slice of the
//       contraction and a full T x T partial (lane = 4 x 8 register tile),
//       partials are summed across warps through shared memory.
//   softmax over keys, dropout on the weights (Philox, same indices as the
//       general kernel), probabilities saved for backward.
//   O = P~ V          : one thread per float4 column, 16 query rows of
//       accumulators at a time, V streamed with 128-bit loads.
// Backward mirrors it: dP = dO V^T with the same routine, softmax backward in
// shared memory, then dQ = dS K, dK = dS^T Q, dV = P~^T dO as three streamed
// passes with register accumulators -- no read-modify-write of global memory.
//
// HBM-bound: algorithmic bytes per (sample, head) 4*4*T*hd forward (q,k,v in,
// o out) and 4*7*T*hd backward, against 4*T*T*hd FLOP -- intensity T/4 FLOP/B.
#include "tile32.cuh"
#include "../../include/immtsf.h"

namespace {

constexpr int XS_RH = 16;    // rows of float4 accumulators per pass

struct XsArgs {
  const float* q; int ldq; const float* k; int ldk; const float* v; int ldv;
  const uint8_t* m_txt; int B, T, H, hd; uint32_t thr; SeedArg seed; float scale;
  float* o; int ldo; float* probs;
  const float* d_o; int lddo; float* dq; int lddq; float* dk; int lddk; float* dv; int lddv;
};

// out[i][:] = sum_j W[i][j] * V[j][:]  for i in [0,T) -- W in shared memory (row stride XS_T), V/out global.
// transposed = true uses W[j][i] instead (dK = dS^T Q, dV = P~^T dO).
template <bool TRANSPOSED>
__device__ __forceinline__ void tile_wv(const float* s_w, const float* __restrict__ V, int ldv, float* __restrict__ out, int ldo,
                                        int T, int hd) {
  const int hd4 = hd >> 2;
  for (int c4 = threadIdx.x; c4 < hd4; c4 += blockDim.x) {
    for (int i0 = 0; i0 < T; i0 += XS_RH) {
      float4 acc[XS_RH];
#pragma unroll
      for (int a = 0; a < XS_RH; ++a) acc[a] = f4_zero();
      for (int j0 = 0; j0 < T; j0 += 4) {  // 4 rows of V in flight per thread (global-load latency, not bandwidth, binds here)
        float4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          vv[u] = j0 + u < T ? __ldg(reinterpret_cast<const float4*>(V + (size_t)(j0 + u) * ldv) + c4) : f4_zero();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, XS_T - 1);
#pragma unroll
          for (int a = 0; a < XS_RH; ++a) {
            const float wgt = TRANSPOSED ? s_w[j * XS_T + i0 + a] : s_w[(i0 + a) * XS_T + j];
            f4_fma(acc[a], wgt, vv[u]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < XS_RH; ++a)
        if (i0 + a < T) reinterpret_cast<float4*>(out + (size_t)(i0 + a) * ldo)[c4] = acc[a];
    }
  }
}

// Same contraction over the output rows [i_begin, i_end) with RH rows per pass and PF rows of V in flight.  RH >= T
// streams V exactly once (the fixed 16-row version computes 32 rows for the Time-IMM windows of 17..24 steps).
template <bool TRANSPOSED, int RH, int PF>
__device__ __forceinline__ void tile_wv_t(const float* s_w, const float* __restrict__ V, int ldv, float* __restrict__ out, int ldo,
                                          int T, int hd, int i_begin, int i_end) {
  const int hd4 = hd >> 2;
  for (int c4 = threadIdx.x; c4 < hd4; c4 += blockDim.x) {
    for (int i0 = i_begin; i0 < i_end; i0 += RH) {
      float4 acc[RH];
#pragma unroll
      for (int a = 0; a < RH; ++a) acc[a] = f4_zero();
      for (int j0 = 0; j0 < T; j0 += PF) {
        float4 vv[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u)
          vv[u] = j0 + u < T ? __ldg(reinterpret_cast<const float4*>(V + (size_t)(j0 + u) * ldv) + c4) : f4_zero();
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          const int j = min(j0 + u, XS_T - 1);
#pragma unroll
          for (int a = 0; a < RH; ++a) {
            const float wgt = TRANSPOSED ? s_w[j * XS_T + min(i0 + a, XS_T - 1)] : s_w[min(i0 + a, XS_T - 1) * XS_T + j];
            f4_fma(acc[a], wgt, vv[u]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < RH; ++a)
        if (i0 + a < i_end) reinterpret_cast<float4*>(out + (size_t)(i0 + a) * ldo)[c4] = acc[a];
    }
  }
}
// 256-thread kernels (128-register cap at 2 CTAs/SM): 16 + 8 rows for 17..24 query times (no padded rows)
template <bool TRANSPOSED>
__device__ __forceinline__ void tile_wv_auto(const float* s_w, const float* __restrict__ V, int ldv, float* __restrict__ out, int ldo,
                                             int T, int hd) {
  if (T <= 8) tile_wv_t<TRANSPOSED, 8, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
  else if (T <= 16 || T > 24) tile_wv_t<TRANSPOSED, 16, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
  else {
    tile_wv_t<TRANSPOSED, 16, 4>(s_w, V, ldv, out, ldo, T, hd, 0, 16);
    tile_wv_t<TRANSPOSED, 8, 4>(s_w, V, ldv, out, ldo, T, hd, 16, T);
  }
}
// 192-thread forward kernel (170-register cap at 2 CTAs/SM): one pass over V up to 24 query times
__device__ __forceinline__ void tile_wv_wide(const float* s_w, const float* __restrict__ V, int ldv, float* __restrict__ out, int ldo,
                                             int T, int hd) {
  if (T <= 8) tile_wv_t<false, 8, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
  else if (T <= 16) tile_wv_t<false, 16, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
  else if (T <= 24) tile_wv_t<false, 24, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
  else tile_wv_t<false, 16, 4>(s_w, V, ldv, out, ldo, T, hd, 0, T);
}

constexpr size_t XS_SMEM = (size_t)(2 * XS_T * XS_LD + 8 * XS_T * XS_T + 2 * XS_T * XS_T) * sizeof(float);

__global__ void __launch_bounds__(256, 2) xattn_small_fwd_kernel(const XsArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_x = smem;
  float* s_y = s_x + XS_T * XS_LD;
  float* s_part = s_y + XS_T * XS_LD;
  float* s_s = s_part + 8 * XS_T * XS_T;  // scores -> dropped probabilities
  const int T = a.T, hd = a.hd, H = a.H;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t rbase = (size_t)b * T;
  float* o = a.o + rbase * a.ldo + h * hd;
  if (a.m_txt[b] == 0) {  // every key masked: the reference overwrites the NaN output with zeros (:79-80)
    for (int i = 0; i < T; ++i)
      for (int c = threadIdx.x; c < hd; c += blockDim.x) o[(size_t)i * a.ldo + c] = 0.f;
    if (a.probs)
      for (int i = threadIdx.x; i < T * T; i += blockDim.x) a.probs[((size_t)b * H + h) * T * T + i] = 0.f;
    return;
  }
  tile_xyt(a.q + rbase * a.ldq + h * hd, a.ldq, a.k + rbase * a.ldk + h * hd, a.ldk, T, T, hd, s_x, s_y, s_part, s_s);
  const float inv_keep = inv_keep_from_thr(a.thr);
  for (int i = w; i < T; i += 8) {  // softmax of row i, then dropout on the weights
    const float sv = lane < T ? a.scale * s_s[i * XS_T + lane] : -INFINITY;
    const float mx = warp_max(sv);
    const float e = lane < T ? expf(sv - mx) : 0.f;
    const float sum = warp_sum(e);
    if (lane < T) {
      const float p = e / sum;
      const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
      if (a.probs) a.probs[pidx] = p;
      s_s[i * XS_T + lane] = p * dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
    } else {
      s_s[i * XS_T + lane] = 0.f;
    }
  }
  for (int i = T * XS_T + threadIdx.x; i < XS_T * XS_T; i += blockDim.x) s_s[i] = 0.f;  // rows >= T
  __syncthreads();
  tile_wv_auto<false>(s_s, a.v + rbase * a.ldv + h * hd, a.ldv, o, a.ldo, T, hd);
}

__global__ void __launch_bounds__(256, 2) xattn_small_bwd_kernel(const XsArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_x = smem;
  float* s_y = s_x + XS_T * XS_LD;
  float* s_part = s_y + XS_T * XS_LD;
  float* s_ds = s_part + 8 * XS_T * XS_T;  // dP -> dS (with the q scale folded in)
  float* s_pt = s_ds + XS_T * XS_T;        // P~ = P * keep / (1-p)
  const int T = a.T, hd = a.hd, H = a.H;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t rbase = (size_t)b * T;
  float* dq = a.dq + rbase * a.lddq + h * hd;
  float* dk = a.dk + rbase * a.lddk + h * hd;
  float* dv = a.dv + rbase * a.lddv + h * hd;
  if (a.m_txt[b] == 0) {
    for (int i = 0; i < T; ++i)
      for (int c = threadIdx.x; c < hd; c += blockDim.x) {
        dq[(size_t)i * a.lddq + c] = 0.f;
        dk[(size_t)i * a.lddk + c] = 0.f;
        dv[(size_t)i * a.lddv + c] = 0.f;
      }
    return;
  }
  const float* go = a.d_o + rbase * a.lddo + h * hd;
  const float* qp = a.q + rbase * a.ldq + h * hd;
  const float* kp = a.k + rbase * a.ldk + h * hd;
  const float* vp = a.v + rbase * a.ldv + h * hd;
  tile_xyt(go, a.lddo, vp, a.ldv, T, T, hd, s_x, s_y, s_part, s_ds);  // dP~[i][j] = dO_i . v_j
  const float inv_keep = inv_keep_from_thr(a.thr);
  for (int i = w; i < XS_T; i += 8) {
    float p = 0.f, dp = 0.f, ks = 0.f;
    if (i < T && lane < T) {
      const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
      ks = dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
      p = a.probs[pidx];
      dp = s_ds[i * XS_T + lane] * ks;
    }
    const float D = warp_sum(p * dp);
    s_ds[i * XS_T + lane] = a.scale * p * (dp - D);
    s_pt[i * XS_T + lane] = p * ks;
  }
  __syncthreads();
  tile_wv_auto<false>(s_ds, kp, a.ldk, dq, a.lddq, T, hd);  // dQ = dS K
  tile_wv_auto<true>(s_ds, qp, a.ldq, dk, a.lddk, T, hd);   // dK = dS^T Q
  tile_wv_auto<true>(s_pt, go, a.lddo, dv, a.lddv, T, hd);  // dV = P~^T dO
}

// ---------------------------------------------------------------- rank-(C+1) query path
// In MMF_XAttn_Add the queries are a projection of the C-channel series (fusions/MMF_XAttn_Add.py:68 + in_proj_q):
// q_i = W y_i + b with y_i in R^C, so the score matrix has rank <= C+1:
//   q_i . k_j = [y_i ; 1] . kq_j ,   kq_j = [W^T k_j ; b . k_j]  in R^(C+1)
// kq is one skinny product over the key rows; q [B*T, d] is never formed, written or read.  Backward (dS as in
// the kernel above, the q scale folded in):
//   Z_j  = sum_i dS_ij [y_i ; 1]          ->  dk_j = [W | b] Z_j,   d[W | b] = sum_j k_j Z_j^T   (host: two skinny products)
//   dy_i = sum_j dS_ij kq_j[:C]           (the query-side gradient into Y_ts)
// so dq [B*T, d] does not exist either and the dQ / dK passes over K and Q disappear.
constexpr int XL_CMAX = 32;              // C + 1 <= 32
constexpr int XL_LD = XL_CMAX + 1;
struct XlArgs {
  const float* y; int ldy; const float* kq; int ldkq; const float* v; int ldv;
  const uint8_t* m_txt; int B, T, H, hd, C; uint32_t thr; SeedArg seed; float scale;
  float* o; int ldo; float* probs;
  const float* d_o; int lddo; float* dv; int lddv; float* z; int ldz; float* dyh;
};

// s_ya[i][c] = [y_i ; 1], s_kq[j][c] = kq_j of head h, c <= C; rows >= T are zero
__device__ __forceinline__ void xl_stage(const XlArgs& a, size_t rbase, int h, float* s_ya, float* s_kq) {
  const int C1 = a.C + 1;
  for (int i = threadIdx.x; i < XS_T * C1; i += blockDim.x) {
    const int r = i / C1, c = i % C1;
    float yv = 0.f, kv = 0.f;
    if (r < a.T) {
      yv = c < a.C ? __ldg(a.y + (rbase + r) * a.ldy + c) : 1.f;
      kv = __ldg(a.kq + (rbase + r) * a.ldkq + h * C1 + c);
    }
    s_ya[r * XL_LD + c] = yv;
    s_kq[r * XL_LD + c] = kv;
  }
}

__global__ void __launch_bounds__(192, 2) xattn_lr_fwd_kernel(const XlArgs a) {
  __shared__ float s_ya[XS_T * XL_LD], s_kq[XS_T * XL_LD];
  __shared__ __align__(16) float s_s[XS_T * XS_T];
  const int T = a.T, hd = a.hd, H = a.H, C1 = a.C + 1;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t rbase = (size_t)b * T;
  float* o = a.o + rbase * a.ldo + h * hd;
  if (a.m_txt[b] == 0) {  // every key masked: the reference overwrites the NaN output with zeros (:79-80)
    for (int i = 0; i < T; ++i)
      for (int c = threadIdx.x; c < hd; c += blockDim.x) o[(size_t)i * a.ldo + c] = 0.f;
    if (a.probs)
      for (int i = threadIdx.x; i < T * T; i += blockDim.x) a.probs[((size_t)b * H + h) * T * T + i] = 0.f;
    return;
  }
  xl_stage(a, rbase, h, s_ya, s_kq);
  __syncthreads();
  const float inv_keep = inv_keep_from_thr(a.thr);
  for (int i = w; i < XS_T; i += (int)(blockDim.x >> 5)) {  // scores of row i (lane = key), softmax, dropout on the weights
    float p_out = 0.f;
    if (i < T) {
      float sv = -INFINITY;
      if (lane < T) {
        float acc = 0.f;
        for (int c = 0; c < C1; ++c) acc = fmaf(s_ya[i * XL_LD + c], s_kq[lane * XL_LD + c], acc);
        sv = a.scale * acc;
      }
      const float mx = warp_max(sv);
      const float e = lane < T ? expf(sv - mx) : 0.f;
      const float sum = warp_sum(e);
      if (lane < T) {
        const float p = e / sum;
        const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
        if (a.probs) a.probs[pidx] = p;
        p_out = p * dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
      }
    }
    s_s[i * XS_T + lane] = p_out;
  }
  __syncthreads();
  tile_wv_wide(s_s, a.v + rbase * a.ldv + h * hd, a.ldv, o, a.ldo, T, hd);
}

__global__ void __launch_bounds__(256, 2) xattn_lr_bwd_kernel(const XlArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* s_x = smem;
  float* s_y = s_x + XS_T * XS_LD;
  float* s_part = s_y + XS_T * XS_LD;
  float* s_ds = s_part + 8 * XS_T * XS_T;
  float* s_pt = s_ds + XS_T * XS_T;
  float* s_ya = s_x;                       // the staging rows are free once tile_xyt has returned
  float* s_kq = s_x + XS_T * XL_LD;
  const int T = a.T, hd = a.hd, H = a.H, C = a.C, C1 = a.C + 1;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t rbase = (size_t)b * T;
  float* dv = a.dv + rbase * a.lddv + h * hd;
  float* z = a.z + rbase * a.ldz + h * C1;
  float* dyh = a.dyh + ((size_t)h * a.B * T + rbase) * C;
  if (a.m_txt[b] == 0) {
    for (int i = 0; i < T; ++i)
      for (int c = threadIdx.x; c < hd; c += blockDim.x) dv[(size_t)i * a.lddv + c] = 0.f;
    for (int i = threadIdx.x; i < T * C1; i += blockDim.x) z[(size_t)(i / C1) * a.ldz + i % C1] = 0.f;
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) dyh[i] = 0.f;
    return;
  }
  const float* go = a.d_o + rbase * a.lddo + h * hd;
  const float* vp = a.v + rbase * a.ldv + h * hd;
  tile_xyt(go, a.lddo, vp, a.ldv, T, T, hd, s_x, s_y, s_part, s_ds);  // dP~[i][j] = dO_i . v_j
  const float inv_keep = inv_keep_from_thr(a.thr);
  for (int i = w; i < XS_T; i += 8) {
    float p = 0.f, dp = 0.f, ks = 0.f;
    if (i < T && lane < T) {
      const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
      ks = dropout_scale(resolve_seed(a.seed), IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
      p = a.probs[pidx];
      dp = s_ds[i * XS_T + lane] * ks;
    }
    const float D = warp_sum(p * dp);
    s_ds[i * XS_T + lane] = a.scale * p * (dp - D);
    s_pt[i * XS_T + lane] = p * ks;
  }
  xl_stage(a, rbase, h, s_ya, s_kq);  // (tile_xyt ended with a barrier: s_x is no longer read)
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * C1; idx += blockDim.x) {  // Z_j = sum_i dS_ij [y_i ; 1]
    const int j = idx / C1, c = idx % C1;
    float acc = 0.f;
    for (int i = 0; i < T; ++i) acc = fmaf(s_ds[i * XS_T + j], s_ya[i * XL_LD + c], acc);
    z[(size_t)j * a.ldz + c] = acc;
  }
  for (int idx = threadIdx.x; idx < T * C; idx += blockDim.x) {  // dy_i = sum_j dS_ij kq_j[:C]
    const int i = idx / C, c = idx % C;
    float acc = 0.f;
    for (int j = 0; j < T; ++j) acc = fmaf(s_ds[i * XS_T + j], s_kq[j * XL_LD + c], acc);
    dyh[(size_t)i * C + c] = acc;
  }
  tile_wv_auto<true>(s_pt, go, a.lddo, dv, a.lddv, T, hd);  // dV = P~^T dO
}

inline bool al16(const void* p, int ld) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0; }

}  // namespace

// 1 if the small-T kernels apply to this problem
int immtsf_xattn_small_ok(int T, int H, int d) { return T >= 1 && T <= XS_T && H >= 1 && d % H == 0 && ((d / H) & 3) == 0; }

int immtsf_xattn_small_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, const uint8_t* m_txt, int B,
                           int T, int H, int d, uint32_t drop_thr, uint64_t seed, float* o, int ldo, float* probs,
                           cudaStream_t st) {
  IMMTSF_REQUIRE(al16(q, ldq) && al16(k, ldk) && al16(v, ldv) && al16(o, ldo), "xattn_core_fwd: operands must be 16B aligned with ld %% 4 == 0");
  XsArgs a = {};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.hd = d / H;
  a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.o = o; a.ldo = ldo; a.probs = probs;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(xattn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM);
    cudaFuncSetAttribute(xattn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM);
    attr = true;
  }
  xattn_small_fwd_kernel<<<B * H, 256, XS_SMEM, st>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_small_fwd");
  return IMMTSF_OK;
}

int immtsf_xattn_small_bwd(const float* d_o, int lddo, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                           const float* probs, const uint8_t* m_txt, int B, int T, int H, int d, uint32_t drop_thr,
                           uint64_t seed, float* dq, int lddq, float* dk, int lddk, float* dv, int lddv, cudaStream_t st) {
  IMMTSF_REQUIRE(al16(d_o, lddo) && al16(q, ldq) && al16(k, ldk) && al16(v, ldv) && al16(dq, lddq) && al16(dk, lddk) && al16(dv, lddv),
                 "xattn_core_bwd: operands must be 16B aligned with ld %% 4 == 0");
  XsArgs a = {};
  a.q = q; a.ldq = ldq; a.k = k; a.ldk = ldk; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.hd = d / H;
  a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.probs = const_cast<float*>(probs);
  a.d_o = d_o; a.lddo = lddo; a.dq = dq; a.lddq = lddq; a.dk = dk; a.lddk = lddk; a.dv = dv; a.lddv = lddv;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(xattn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM);
    cudaFuncSetAttribute(xattn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM);
    attr = true;
  }
  xattn_small_bwd_kernel<<<B * H, 256, XS_SMEM, st>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_small_bwd");
  return IMMTSF_OK;
}

// ---- rank-(C+1) query path (see the kernels): y [B*T, C], kq [B*T, H*(C+1)] (per head [W_h^T k ; b_h . k]) --------
extern "C" int immtsf_xattn_lowrank_ok(int T, int H, int d, int C) {
  return immtsf_xattn_small_ok(T, H, d) && C >= 1 && C + 1 <= XL_CMAX;
}

extern "C" int immtsf_xattn_lowrank_fwd(const float* y, int ldy, const float* kq, int ldkq, const float* v, int ldv,
                                        const uint8_t* m_txt, int B, int T, int H, int d, int C, uint32_t drop_thr, uint64_t seed,
                                        float* o, int ldo, float* probs, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(y && kq && v && m_txt && o, "xattn_lowrank_fwd: null pointer");
  IMMTSF_REQUIRE(immtsf_xattn_lowrank_ok(T, H, d, C), "xattn_lowrank_fwd: needs T <= 32, C + 1 <= 32, head_dim %% 4 == 0 (T=%d H=%d d=%d C=%d)", T, H, d, C);
  IMMTSF_REQUIRE(al16(v, ldv) && al16(o, ldo) && ldy >= C && ldkq >= H * (C + 1), "xattn_lowrank_fwd: v/o must be 16B aligned with ld %% 4 == 0");
  XlArgs a = {};
  a.y = y; a.ldy = ldy; a.kq = kq; a.ldkq = ldkq; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.hd = d / H; a.C = C;
  a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.o = o; a.ldo = ldo; a.probs = probs;
  xattn_lr_fwd_kernel<<<B * H, 192, 0, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_lowrank_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_xattn_lowrank_bwd(const float* d_o, int lddo, const float* y, int ldy, const float* kq, int ldkq, const float* v,
                                        int ldv, const float* probs, const uint8_t* m_txt, int B, int T, int H, int d, int C,
                                        uint32_t drop_thr, uint64_t seed, float* dv, int lddv, float* z, int ldz, float* dyh,
                                        void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(d_o && y && kq && v && probs && m_txt && dv && z && dyh, "xattn_lowrank_bwd: null pointer");
  IMMTSF_REQUIRE(immtsf_xattn_lowrank_ok(T, H, d, C), "xattn_lowrank_bwd: needs T <= 32, C + 1 <= 32, head_dim %% 4 == 0 (T=%d H=%d d=%d C=%d)", T, H, d, C);
  IMMTSF_REQUIRE(al16(d_o, lddo) && al16(v, ldv) && al16(dv, lddv) && ldy >= C && ldkq >= H * (C + 1) && ldz >= H * (C + 1),
                 "xattn_lowrank_bwd: d_o/v/dv must be 16B aligned with ld %% 4 == 0");
  XlArgs a = {};
  a.y = y; a.ldy = ldy; a.kq = kq; a.ldkq = ldkq; a.v = v; a.ldv = ldv; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.hd = d / H; a.C = C;
  a.thr = drop_thr; a.seed = make_seed(seed); a.scale = (float)sqrt(1.0 / (double)(d / H)); a.probs = const_cast<float*>(probs);
  a.d_o = d_o; a.lddo = lddo; a.dv = dv; a.lddv = lddv; a.z = z; a.ldz = ldz; a.dyh = dyh;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(xattn_lr_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XS_SMEM);
    attr = true;
  }
  xattn_lr_bwd_kernel<<<B * H, 256, XS_SMEM, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_lowrank_bwd");
  return IMMTSF_OK;
}
