// MMF_XAttn_Add as a rank-(2C+1) form in E_txt (fusions/MMF_XAttn_Add.py:65-103 + nn.MultiheadAttention).
//
// The reference computes, per sample and head h (hd = d/H columns):
//   q_i = in_q (W_Q y_i) + b_q          y_i in R^C  (the forecaster's output channels)
//   k_j = in_k (W_K e_j) + b_k ,  v_j = in_v (W_V e_j) + b_v        e_j in R^d_txt
//   delta_i = W_r (W_o sum_j softmax_j(q_i . k_j / sqrt(hd)) v_j + b_o) + b_r   in R^C
// Both ends of the attention are C-dimensional, so with Wq_aug = [in_q W_Q | b_q] (d x (C+1)) and Wo_f = W_r W_o (C x d)
//   q_i . k_j   = [y_i ; 1] . kq_j ,   kq_j = A e_j + a0 ,  A  = Wq_aug^T in_k W_K  ((C+1) x d_txt),  a0 = Wq_aug^T b_k
//   W_r W_o v_j = vo_j            ,   vo_j = G e_j + g0 ,  G  = Wo_f in_v W_V     ( C    x d_txt),  g0 = Wo_f b_v
// (per head: the rows / columns of head h).  The only pass over the wide data is ONE skinny product
//   R = [kq | vo] = E_txt [A ; G]^T + [a0 ; g0]      [B*T, H*(2C+1)]
// and the attention itself runs on T x (2C+1) numbers per (sample, head): no q, k, v, o, dq, dk, dv of width d is
// ever formed, and the d x d weight matrices only meet in skinny weight-space products.  Backward returns
//   dR = [Z | U],  Z_j = sum_i dS_ij [y_i ; 1],  U_j = sum_i P~_ij d_delta_i   (then dE = dR [A ; G], d[A ; G] = dR^T E)
//   dY_i = sum_h sum_j dS_ij kq_j[:C]
// The kernels below are that tiny attention: one CTA per sample, all heads, T <= 32, H*(C+1) <= 64.
#include "common.cuh"
#include "rowtile.cuh"
#include "../../include/immtsf.h"

namespace {

constexpr int XR_T = 32;        // max query / key times
constexpr int XR_W = 64;        // max H * (C + 1)
constexpr int XR_LD = XR_W + 1;

struct XrArgs {
  const float* y; int ldy;      // [B*T, C]
  const float* r; int ldr;      // [B*T, H*(C+1) + H*C] = [kq | vo]
  const float* bo;              // [C] folded output bias W_r b_o + b_r
  const uint8_t* m_txt; int B, T, H, C; float scale; uint32_t thr; SeedArg seed;
  float* delta_y;               // [B*T, C]
  float* probs;                 // [B, H, T, T] (nullable in forward)
  const float* d_delta;         // [B*T, C]
  float* dr; int lddr;          // [B*T, H*(2C+1)]
  float* dy;                    // [B*T, C]
};

__global__ void __launch_bounds__(128) xattn_rank_fwd_kernel(const XrArgs a) {
  __shared__ float s_y[XR_T][33];       // [y_i ; 1]
  __shared__ float s_kq[XR_T][XR_LD];   // all heads
  __shared__ float s_vo[XR_T][XR_LD];
  __shared__ float s_p[XR_T][33];       // P~ of the current head
  __shared__ float s_out[XR_T][33];     // delta accumulated over heads
  const int T = a.T, H = a.H, C = a.C, C1 = C + 1, n1 = H * C1;
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t rbase = (size_t)b * T;
  if (a.m_txt[b] == 0) {  // every key masked: the attention output is 0 (:79-80), so delta = the folded bias; the tail zeroes it
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) a.delta_y[rbase * C + i] = __ldg(a.bo + i % C);
    if (a.probs)
      for (int i = threadIdx.x; i < H * T * T; i += blockDim.x) a.probs[(size_t)b * H * T * T + i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < T * C1; i += blockDim.x) {
    const int r = i / C1, c = i % C1;
    s_y[r][c] = c < C ? __ldg(a.y + (rbase + r) * a.ldy + c) : 1.f;
  }
  for (int i = threadIdx.x; i < T * n1; i += blockDim.x) s_kq[i / n1][i % n1] = __ldg(a.r + (rbase + i / n1) * a.ldr + i % n1);
  for (int i = threadIdx.x; i < T * H * C; i += blockDim.x)
    s_vo[i / (H * C)][i % (H * C)] = __ldg(a.r + (rbase + i / (H * C)) * a.ldr + n1 + i % (H * C));
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) s_out[i / C][i % C] = __ldg(a.bo + i % C);
  __syncthreads();
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  for (int h = 0; h < H; ++h) {
    for (int i = w; i < T; i += nw) {  // row i: scores (lane = key), softmax, dropout on the weights
      float sv = -INFINITY;
      if (lane < T) {
        float acc = 0.f;
        for (int c = 0; c < C1; ++c) acc = fmaf(s_y[i][c], s_kq[lane][h * C1 + c], acc);
        sv = a.scale * acc;
      }
      const float mx = warp_max(sv);
      const float e = lane < T ? expf(sv - mx) : 0.f;
      const float sum = warp_sum(e);
      float pt = 0.f;
      if (lane < T) {
        const float p = e / sum;
        const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
        if (a.probs) a.probs[pidx] = p;
        pt = p * dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
      }
      s_p[i][lane] = pt;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * C; idx += blockDim.x) {  // delta_i += sum_j P~_ij vo_j
      const int i = idx / C, c = idx % C;
      float acc = s_out[i][c];
      for (int j = 0; j < T; ++j) acc = fmaf(s_p[i][j], s_vo[j][h * C + c], acc);
      s_out[i][c] = acc;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) a.delta_y[rbase * C + i] = s_out[i / C][i % C];
}

__global__ void __launch_bounds__(128) xattn_rank_bwd_kernel(const XrArgs a) {
  __shared__ float s_y[XR_T][33];
  __shared__ float s_kq[XR_T][XR_LD];
  __shared__ float s_vo[XR_T][XR_LD];
  __shared__ float s_dd[XR_T][33];      // d_delta rows
  __shared__ float s_ds[XR_T][33];      // dS of the current head (q scale folded in)
  __shared__ float s_pt[XR_T][33];      // P~ of the current head
  __shared__ float s_dy[XR_T][33];      // query-side gradient, summed over heads
  const int T = a.T, H = a.H, C = a.C, C1 = C + 1, n1 = H * C1;
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t rbase = (size_t)b * T;
  const int nr = n1 + H * C;
  if (a.m_txt[b] == 0) {
    for (int i = threadIdx.x; i < T * nr; i += blockDim.x) a.dr[(rbase + i / nr) * a.lddr + i % nr] = 0.f;
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) a.dy[rbase * C + i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < T * C1; i += blockDim.x) {
    const int r = i / C1, c = i % C1;
    s_y[r][c] = c < C ? __ldg(a.y + (rbase + r) * a.ldy + c) : 1.f;
  }
  for (int i = threadIdx.x; i < T * n1; i += blockDim.x) s_kq[i / n1][i % n1] = __ldg(a.r + (rbase + i / n1) * a.ldr + i % n1);
  for (int i = threadIdx.x; i < T * H * C; i += blockDim.x)
    s_vo[i / (H * C)][i % (H * C)] = __ldg(a.r + (rbase + i / (H * C)) * a.ldr + n1 + i % (H * C));
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    s_dd[i / C][i % C] = __ldg(a.d_delta + rbase * C + i);
    s_dy[i / C][i % C] = 0.f;
  }
  __syncthreads();
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  for (int h = 0; h < H; ++h) {
    for (int i = w; i < T; i += nw) {  // softmax backward of row i (lane = key)
      float p = 0.f, dp = 0.f, ks = 0.f;
      if (lane < T) {
        const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
        ks = dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
        p = a.probs[pidx];
        float acc = 0.f;  // dP~_ij = d_delta_i . vo_j
        for (int c = 0; c < C; ++c) acc = fmaf(s_dd[i][c], s_vo[lane][h * C + c], acc);
        dp = acc * ks;
      }
      const float D = warp_sum(p * dp);
      s_ds[i][lane] = a.scale * p * (dp - D);
      s_pt[i][lane] = p * ks;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * C1; idx += blockDim.x) {  // Z_j = sum_i dS_ij [y_i ; 1]
      const int j = idx / C1, c = idx % C1;
      float acc = 0.f;
      for (int i = 0; i < T; ++i) acc = fmaf(s_ds[i][j], s_y[i][c], acc);
      a.dr[(rbase + j) * a.lddr + h * C1 + c] = acc;
    }
    for (int idx = threadIdx.x; idx < T * C; idx += blockDim.x) {
      const int r = idx / C, c = idx % C;
      float u = 0.f;  // U_j = sum_i P~_ij d_delta_i   (r = j)
      for (int i = 0; i < T; ++i) u = fmaf(s_pt[i][r], s_dd[i][c], u);
      a.dr[(rbase + r) * a.lddr + n1 + h * C + c] = u;
      float g = s_dy[r][c];  // dy_i += sum_j dS_ij kq_j[c]   (r = i; each (r, c) is owned by one thread)
      for (int j = 0; j < T; ++j) g = fmaf(s_ds[r][j], s_kq[j][h * C1 + c], g);
      s_dy[r][c] = g;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) a.dy[rbase * C + i] = s_dy[i / C][i % C];
}

}  // namespace

extern "C" int immtsf_xattn_rank_ok(int T, int H, int d, int C) {
  return T >= 1 && T <= XR_T && H >= 1 && d % H == 0 && C >= 1 && C <= 31 && H * (C + 1) <= XR_W;
}

extern "C" int immtsf_xattn_rank_fwd(const float* y, int ldy, const float* r, int ldr, const float* bo, const uint8_t* m_txt,
                                     int B, int T, int H, int d, int C, uint32_t drop_thr, uint64_t seed, float* delta_y,
                                     float* probs, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(y && r && bo && m_txt && delta_y, "xattn_rank_fwd: null pointer");
  IMMTSF_REQUIRE(immtsf_xattn_rank_ok(T, H, d, C), "xattn_rank_fwd: needs T <= 32, C <= 31, H*(C+1) <= 64 (T=%d H=%d d=%d C=%d)", T, H, d, C);
  IMMTSF_REQUIRE(ldy >= C && ldr >= H * (2 * C + 1), "xattn_rank_fwd: leading dimension too small");
  XrArgs a = {};
  a.y = y; a.ldy = ldy; a.r = r; a.ldr = ldr; a.bo = bo; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.C = C;
  a.scale = (float)sqrt(1.0 / (double)(d / H)); a.thr = drop_thr; a.seed = make_seed(seed); a.delta_y = delta_y; a.probs = probs;
  xattn_rank_fwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_rank_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_xattn_rank_bwd(const float* d_delta, const float* y, int ldy, const float* r, int ldr, const float* probs,
                                     const uint8_t* m_txt, int B, int T, int H, int d, int C, uint32_t drop_thr, uint64_t seed,
                                     float* dr, int lddr, float* dy, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(d_delta && y && r && probs && m_txt && dr && dy, "xattn_rank_bwd: null pointer");
  IMMTSF_REQUIRE(immtsf_xattn_rank_ok(T, H, d, C), "xattn_rank_bwd: needs T <= 32, C <= 31, H*(C+1) <= 64 (T=%d H=%d d=%d C=%d)", T, H, d, C);
  IMMTSF_REQUIRE(ldy >= C && ldr >= H * (2 * C + 1) && lddr >= H * (2 * C + 1), "xattn_rank_bwd: leading dimension too small");
  XrArgs a = {};
  a.y = y; a.ldy = ldy; a.r = r; a.ldr = ldr; a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.C = C;
  a.scale = (float)sqrt(1.0 / (double)(d / H)); a.thr = drop_thr; a.seed = make_seed(seed); a.probs = const_cast<float*>(probs);
  a.d_delta = d_delta; a.dr = dr; a.lddr = lddr; a.dy = dy;
  xattn_rank_bwd_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_rank_bwd");
  return IMMTSF_OK;
}
