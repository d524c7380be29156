// MMF_XAttn_Add in its rank-(2C+1) form (csrc/xattn_rank.cu, fusions/MMF_XAttn_Add.py:65-103), data half in ONE launch
// per direction.  Round 1 ran it as R = E Wr^T (skinny product) -> tiny attention -> tail, and in backward as tail ->
// column sum -> attention -> axpby -> dWr = dR^T E (two launches) -> column sum (two launches) -> dE = dR Wr: 3 + 11
// launches that each stream the [B*T, d] text tensor at 13-20 % of HBM or are pure launch latency (cfg2 timeline: 35 us
// forward, 100 us backward on the critical path).  Here one CTA owns one sample:
//   forward : R rows from the sample's T rows of E (Wr staged in shared memory, warp per row), the T x (2C+1) attention per
//             head, then LayerNorm_C / dropout / kappa blend (:83-102) -> Y_out.  E is read once; R, delta, P are saved.
//   backward: tail backward -> d_delta, attention backward -> dR (stays in shared memory), dY; then, thread per float4
//             column of E, dE_r = dR_r Wr (written once) and the CTA's partial of dWr = dR^T E (E read once); per-CTA
//             partials of dWr, dbr, d(bo), dgamma, dbeta are added in CTA order by a second small launch (deterministic,
//             no atomics, no zero-fill).
// Eligibility: T <= 32, H (C+1) <= 64, C <= 31 (xattn_rank), nr = H (2C+1) <= 16, d_e % 4 == 0, d_e <= 1024.
#include "common.cuh"
#include "rowtile.cuh"
#include "../../include/immtsf.h"

namespace {

constexpr int XF_T = 32;
constexpr int XF_NR = 16;   // max H * (2C + 1)
constexpr int XF_THREADS = 256;

// floats per CTA in the partials buffer: dWr [nr][de] | dbr [nr] | dbo [C] | dgamma [C] | dbeta [C], rounded up to 16 bytes
__host__ __device__ inline size_t part_stride(int nr, int de, int C) { return (((size_t)nr * de + nr + 3 * C) + 3) & ~(size_t)3; }

struct XfArgs {
  const float* e; int lde; int de;   // [B*T, de] text-side tensor (E_txt, or dropout(LN(.)) when the TTF projection is folded)
  const float* wr; int ldwr;         // [nr, de]
  const float* br;                   // [nr]
  const float* y; int ldy;           // [B*T, C]
  const float* bo;                   // [C]
  const float* gamma; const float* beta;
  const uint8_t* m_txt;
  int B, T, H, C; float scale, eps, kappa; uint32_t thr; SeedArg seed;
  float* r; int ldr;                 // [B*T, nr] (saved)
  float* delta_y;                    // [B*T, C]  (saved)
  float* probs;                      // [B, H, T, T] (saved; nullable in forward)
  float* y_out;                      // [B*T, C]
  int32_t* flags;
  // backward
  const float* dy_out;               // [B*T, C]
  float* de_out; int ldde;           // [B*T, de]
  float* dy;                         // [B*T, C]
  float* partial;                    // [grid][nr*de + nr + 3C]
};

// R rows of one sample into shared memory: warp per row, lanes stride float4 chunks of E; Wr comes from shared memory.
template <int NR>
__device__ __forceinline__ bool rows_times_wr(const XfArgs& a, const float* __restrict__ s_wr, size_t rbase, int T, int nr,
                                              float (*s_r)[XF_NR + 1]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5, de4 = a.de >> 2;
  bool bad = false;
  // two rows per pass: both rows' loads are requested before either row's arithmetic (ncu: 4.5 long-scoreboard stalls per
  // issue with one row in flight), and every Wr chunk read from shared memory feeds both rows
  for (int t = w; t < T; t += 2 * nw) {
    const bool two = t + nw < T;
    float acc0[NR], acc1[NR];
#pragma unroll
    for (int c = 0; c < NR; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
    const float4* ep0 = reinterpret_cast<const float4*>(a.e + (rbase + t) * a.lde);
    const float4* ep1 = reinterpret_cast<const float4*>(a.e + (rbase + (two ? t + nw : t)) * a.lde);
    for (int k4 = lane; k4 < de4; k4 += 32) {
      const float4 e0 = __ldg(ep0 + k4);
      const float4 e1 = __ldg(ep1 + k4);
      bad |= (e0.x != e0.x) | (e0.y != e0.y) | (e0.z != e0.z) | (e0.w != e0.w) | (e1.x != e1.x) | (e1.y != e1.y) | (e1.z != e1.z) | (e1.w != e1.w);
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        if (c < nr) {
          const float4 wv = *reinterpret_cast<const float4*>(s_wr + (size_t)c * a.de + 4 * k4);
          acc0[c] = fmaf(e0.x, wv.x, fmaf(e0.y, wv.y, fmaf(e0.z, wv.z, fmaf(e0.w, wv.w, acc0[c]))));
          acc1[c] = fmaf(e1.x, wv.x, fmaf(e1.y, wv.y, fmaf(e1.z, wv.z, fmaf(e1.w, wv.w, acc1[c]))));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NR; ++c) {
      if (c < nr) {
        const float bc = __ldg(a.br + c);
        const float v0 = warp_sum(acc0[c]) + bc, v1 = warp_sum(acc1[c]) + bc;
        if (lane == 0) {
          s_r[t][c] = v0;
          a.r[(rbase + t) * a.ldr + c] = v0;
          if (two) {
            s_r[t + nw][c] = v1;
            a.r[(rbase + t + nw) * a.ldr + c] = v1;
          }
        }
      }
    }
  }
  return bad;
}

template <int NR>
__global__ void __launch_bounds__(XF_THREADS, 2) xattn_rank_fused_fwd_kernel(const XfArgs a) {
  extern __shared__ __align__(16) float s_wr[];  // [nr][de]
  __shared__ float s_y[XF_T][33];                // [y_i ; 1]
  __shared__ float s_r[XF_T][XF_NR + 1];         // [kq | vo] of the sample
  __shared__ float s_p[XF_T][33];                // P~ of the current head
  __shared__ float s_out[XF_T][33];              // delta accumulated over heads
  const int T = a.T, H = a.H, C = a.C, C1 = C + 1, n1 = H * C1, nr = n1 + H * C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t rbase = (size_t)b * T;
  const bool has_txt = a.m_txt[b] != 0;
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  bool bad_e = false, bad_out = false;
  if (has_txt) {
    for (int i = threadIdx.x; i < nr * (a.de >> 2); i += blockDim.x) {
      const int c = i / (a.de >> 2), k4 = i % (a.de >> 2);
      reinterpret_cast<float4*>(s_wr)[i] = __ldg(reinterpret_cast<const float4*>(a.wr + (size_t)c * a.ldwr) + k4);
    }
    for (int i = threadIdx.x; i < T * C1; i += blockDim.x) {
      const int r = i / C1, c = i % C1;
      s_y[r][c] = c < C ? __ldg(a.y + (rbase + r) * a.ldy + c) : 1.f;
    }
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) s_out[i / C][i % C] = __ldg(a.bo + i % C);
    __syncthreads();
    bad_e = rows_times_wr<NR>(a, s_wr, rbase, T, nr, s_r);
    __syncthreads();
    for (int h = 0; h < H; ++h) {
      for (int i = w; i < T; i += nw) {  // row i: scores (lane = key), softmax, dropout on the weights
        float sv = -INFINITY;
        if (lane < T) {
          float acc = 0.f;
          for (int c = 0; c < C1; ++c) acc = fmaf(s_y[i][c], s_r[lane][h * C1 + c], acc);
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
        for (int j = 0; j < T; ++j) acc = fmaf(s_p[i][j], s_r[j][n1 + h * C + c], acc);
        s_out[i][c] = acc;
      }
      __syncthreads();
    }
  } else {
    // every key masked: the attention output is 0 (:79-80), delta = the folded bias, and the tail zeroes it (:98-99)
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) s_out[i / C][i % C] = __ldg(a.bo + i % C);
    for (int i = threadIdx.x; i < T * nr; i += blockDim.x) a.r[(rbase + i / nr) * a.ldr + i % nr] = 0.f;
    if (a.probs)
      for (int i = threadIdx.x; i < H * T * T; i += blockDim.x) a.probs[(size_t)b * H * T * T + i] = 0.f;
    __syncthreads();
  }
  // tail (:83-102): LayerNorm over C, dropout, zero without text, kappa blend.  Warp per row, lane = channel (C <= 31).
  for (int t = w; t < T; t += nw) {
    const float x = lane < C ? s_out[t][lane] : 0.f;
    const float mu = warp_sum(x) / (float)C;
    const float dv = lane < C ? x - mu : 0.f;
    const float rs = 1.f / sqrtf(warp_sum(dv * dv) / (float)C + a.eps);
    if (lane < C) {
      const size_t idx = (rbase + t) * C + lane;
      a.delta_y[idx] = x;
      bad_out |= x != x;
      float dd = 0.f;
      if (has_txt) dd = (dv * rs * __ldg(a.gamma + lane) + __ldg(a.beta + lane)) * dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, idx, a.thr, inv_keep);
      const float o = (__ldg(a.y + (rbase + t) * a.ldy + lane) + a.kappa * dd) / (1.f + a.kappa);
      a.y_out[idx] = o;
      bad_out |= o != o;
    }
  }
  if (a.flags != nullptr) {
    if (__any_sync(0xffffffffu, bad_e) && lane == 0) a.flags[IMMTSF_FLAG_E] = 1;
    if (__any_sync(0xffffffffu, bad_out) && lane == 0) a.flags[IMMTSF_FLAG_OUT] = 1;
  }
}

template <int NR>
__global__ void __launch_bounds__(XF_THREADS, 2) xattn_rank_fused_bwd_kernel(const XfArgs a) {
  extern __shared__ __align__(16) float s_wr[];  // [nr][de]
  __shared__ float s_y[XF_T][33];
  __shared__ float s_r[XF_T][XF_NR + 1];
  __shared__ float s_dd[XF_T][33];      // d_delta rows
  __shared__ float s_ds[XF_T][33];      // dS of the current head (q scale folded in)
  __shared__ float s_pt[XF_T][33];      // P~ of the current head
  __shared__ float s_dy[XF_T][33];      // query-side gradient, summed over heads
  __shared__ float s_dr[XF_T][XF_NR + 1];
  __shared__ float s_small[XF_NR + 3 * 32];  // CTA partials: dbr [nr] | dbo [C] | dgamma [C] | dbeta [C]
  const int T = a.T, H = a.H, C = a.C, C1 = C + 1, n1 = H * C1, nr = n1 + H * C;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int de4 = a.de >> 2;
  const bool col = (int)threadIdx.x < de4;  // this thread owns float4 column threadIdx.x of E / dE / Wr
  const float inv_keep = inv_keep_from_thr(a.thr);
  const uint64_t seed = resolve_seed(a.seed);
  const float kfac = a.kappa / (1.f + a.kappa), pass = 1.f / (1.f + a.kappa);
  float4 acc[NR];  // the thread's column of this CTA's partial dWr
#pragma unroll
  for (int c = 0; c < NR; ++c) acc[c] = f4_zero();
  for (int i = threadIdx.x; i < nr * de4; i += blockDim.x)
    reinterpret_cast<float4*>(s_wr)[i] = __ldg(reinterpret_cast<const float4*>(a.wr + (size_t)(i / de4) * a.ldwr) + i % de4);
  float dgam = 0.f, dbet = 0.f, dbo = 0.f;  // lane = channel, summed over this warp's rows
  float dbr = 0.f;                          // thread c < nr
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const size_t rbase = (size_t)b * T;
    const bool has_txt = a.m_txt[b] != 0;
    __syncthreads();  // the previous sample's shared arrays are no longer read
    if (!has_txt) {  // delta is zeroed (:98-99): nothing flows to the text side; Y passes straight through the blend
      for (int i = threadIdx.x; i < T * C; i += blockDim.x) a.dy[rbase * C + i] = pass * __ldg(a.dy_out + rbase * C + i);
      if (col)
        for (int t = 0; t < T; ++t) reinterpret_cast<float4*>(a.de_out + (rbase + t) * a.ldde)[threadIdx.x] = f4_zero();
      continue;
    }
    for (int i = threadIdx.x; i < T * C1; i += blockDim.x) {
      const int r = i / C1, c = i % C1;
      s_y[r][c] = c < C ? __ldg(a.y + (rbase + r) * a.ldy + c) : 1.f;
    }
    for (int i = threadIdx.x; i < T * nr; i += blockDim.x) s_r[i / nr][i % nr] = __ldg(a.r + (rbase + i / nr) * a.ldr + i % nr);
    for (int i = threadIdx.x; i < T * C; i += blockDim.x) s_dy[i / C][i % C] = 0.f;
    // tail backward: warp per row, lane = channel
    for (int t = w; t < T; t += nw) {
      const size_t idx = (rbase + t) * C + lane;
      const float x0 = lane < C ? __ldg(a.delta_y + idx) : 0.f;
      const float mu = warp_sum(x0) / (float)C;
      const float dv = lane < C ? x0 - mu : 0.f;
      const float rs = 1.f / sqrtf(warp_sum(dv * dv) / (float)C + a.eps);
      float g = 0.f, x = 0.f;
      if (lane < C) {
        x = dv * rs;
        const float dn = __ldg(a.dy_out + idx) * kfac * dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, idx, a.thr, inv_keep);
        dgam = fmaf(dn, x, dgam);
        dbet += dn;
        g = dn * __ldg(a.gamma + lane);
      }
      const float m1 = warp_sum(g) / (float)C, m2 = warp_sum(g * x) / (float)C;
      if (lane < C) {
        const float dd = rs * (g - m1 - x * m2);
        s_dd[t][lane] = dd;
        dbo += dd;
      }
    }
    __syncthreads();
    for (int h = 0; h < H; ++h) {
      for (int i = w; i < T; i += nw) {  // softmax backward of row i (lane = key)
        float p = 0.f, dp = 0.f, ks = 0.f;
        if (lane < T) {
          const size_t pidx = (((size_t)b * H + h) * T + i) * T + lane;
          ks = dropout_scale(seed, IMMTSF_SITE_MMF_ATTN, pidx, a.thr, inv_keep);
          p = a.probs[pidx];
          float s = 0.f;  // dP~_ij = d_delta_i . vo_j
          for (int c = 0; c < C; ++c) s = fmaf(s_dd[i][c], s_r[lane][n1 + h * C + c], s);
          dp = s * ks;
        }
        const float D = warp_sum(p * dp);
        s_ds[i][lane] = a.scale * p * (dp - D);
        s_pt[i][lane] = p * ks;
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < T * C1; idx += blockDim.x) {  // Z_j = sum_i dS_ij [y_i ; 1]
        const int j = idx / C1, c = idx % C1;
        float s = 0.f;
        for (int i = 0; i < T; ++i) s = fmaf(s_ds[i][j], s_y[i][c], s);
        s_dr[j][h * C1 + c] = s;
      }
      for (int idx = threadIdx.x; idx < T * C; idx += blockDim.x) {
        const int r = idx / C, c = idx % C;
        float u = 0.f;  // U_j = sum_i P~_ij d_delta_i   (r = j)
        for (int i = 0; i < T; ++i) u = fmaf(s_pt[i][r], s_dd[i][c], u);
        s_dr[r][n1 + h * C + c] = u;
        float g = s_dy[r][c];  // dy_i += sum_j dS_ij kq_j[c]   (r = i; each (r, c) is owned by one thread)
        for (int j = 0; j < T; ++j) g = fmaf(s_ds[r][j], s_r[j][h * C1 + c], g);
        s_dy[r][c] = g;
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < T * C; i += blockDim.x)
      a.dy[rbase * C + i] = s_dy[i / C][i % C] + pass * __ldg(a.dy_out + rbase * C + i);
    if ((int)threadIdx.x < nr)
      for (int t = 0; t < T; ++t) dbr += s_dr[t][threadIdx.x];
    if (col) {  // dE_t = dR_t Wr ;  dWr += dR^T E
      for (int t = 0; t < T; ++t) {
        const float4 ev = __ldg(reinterpret_cast<const float4*>(a.e + (rbase + t) * a.lde) + threadIdx.x);
        float4 de = f4_zero();
#pragma unroll
        for (int c = 0; c < NR; ++c) {
          if (c < nr) {
            const float drc = s_dr[t][c];
            f4_fma(de, drc, *reinterpret_cast<const float4*>(s_wr + (size_t)c * a.de + 4 * threadIdx.x));
            f4_fma(acc[c], drc, ev);
          }
        }
        reinterpret_cast<float4*>(a.de_out + (rbase + t) * a.ldde)[threadIdx.x] = de;
      }
    }
  }
  // per-CTA partials
  float* part = a.partial + (size_t)blockIdx.x * part_stride(nr, a.de, C);
  if (col) {
#pragma unroll
    for (int c = 0; c < NR; ++c)
      if (c < nr) reinterpret_cast<float4*>(part + (size_t)c * a.de)[threadIdx.x] = acc[c];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < XF_NR + 96; i += blockDim.x) s_small[i] = 0.f;
  __syncthreads();
  if ((int)threadIdx.x < nr) s_small[threadIdx.x] = dbr;
  if (lane < C) {
    atomicAdd(&s_small[XF_NR + lane], dbo);
    atomicAdd(&s_small[XF_NR + 32 + lane], dgam);
    atomicAdd(&s_small[XF_NR + 64 + lane], dbet);
  }
  __syncthreads();
  float* ps = part + (size_t)nr * a.de;
  for (int i = threadIdx.x; i < nr; i += blockDim.x) ps[i] = s_small[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    ps[nr + i] = s_small[XF_NR + i];
    ps[nr + C + i] = s_small[XF_NR + 32 + i];
    ps[nr + 2 * C + i] = s_small[XF_NR + 64 + i];
  }
}

// out[i] = sum over the CTAs' partials in CTA order (deterministic).  blockDim (32, 8): 8 partial chains per output.
__global__ void __launch_bounds__(256) xattn_rank_partials_reduce_kernel(const float* __restrict__ partial, int nparts, int len,
                                                                         size_t stride, int n_wr, float* __restrict__ dwr,
                                                                         float* __restrict__ small) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < len)
    for (int p = threadIdx.y; p < nparts; p += 8) s += partial[(size_t)p * stride + i];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < len) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
    if (i < n_wr) dwr[i] = t;
    else small[i - n_wr] = t;
  }
}

bool fused_ok(int T, int H, int d, int C, int de) {
  return immtsf_xattn_rank_ok(T, H, d, C) && H * (2 * C + 1) <= XF_NR && de % 4 == 0 && de >= 4 && de <= 4 * XF_THREADS &&
         (size_t)H * (2 * C + 1) * de * sizeof(float) <= 96 * 1024;
}

int fused_grid(int B) { return B < 2 * 148 ? B : 2 * 148; }

}  // namespace

extern "C" int immtsf_xattn_rank_fused_ok(int T, int H, int d, int C, int de) { return fused_ok(T, H, d, C, de) ? 1 : 0; }

extern "C" size_t immtsf_xattn_rank_fused_bwd_workspace_bytes(int B, int H, int C, int de) {
  return (size_t)fused_grid(B) * part_stride(H * (2 * C + 1), de, C) * sizeof(float);
}

extern "C" int immtsf_xattn_rank_fused_fwd(const float* e, int lde, int de, const float* wr, int ldwr, const float* br,
                                           const float* y, int ldy, const float* bo, const float* gamma, const float* beta,
                                           const uint8_t* m_txt, int B, int T, int H, int d, int C, float eps, float kappa,
                                           uint32_t drop_thr, uint64_t seed, float* r, int ldr, float* delta_y, float* probs,
                                           float* y_out, int32_t* flags, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(e && wr && br && y && bo && gamma && beta && m_txt && r && delta_y && y_out, "xattn_rank_fused_fwd: null pointer");
  IMMTSF_REQUIRE(fused_ok(T, H, d, C, de), "xattn_rank_fused_fwd: unsupported shape (T=%d H=%d d=%d C=%d de=%d)", T, H, d, C, de);
  const int nr = H * (2 * C + 1);
  IMMTSF_REQUIRE(lde >= de && (lde & 3) == 0 && ((uintptr_t)e & 15) == 0 && ldwr >= de && (ldwr & 3) == 0 && ((uintptr_t)wr & 15) == 0 &&
                     ldy >= C && ldr >= nr,
                 "xattn_rank_fused_fwd: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  XfArgs a = {};
  a.e = e; a.lde = lde; a.de = de; a.wr = wr; a.ldwr = ldwr; a.br = br; a.y = y; a.ldy = ldy; a.bo = bo; a.gamma = gamma; a.beta = beta;
  a.m_txt = m_txt; a.B = B; a.T = T; a.H = H; a.C = C; a.scale = (float)sqrt(1.0 / (double)(d / H)); a.eps = eps; a.kappa = kappa;
  a.thr = drop_thr; a.seed = make_seed(seed); a.r = r; a.ldr = ldr; a.delta_y = delta_y; a.probs = probs; a.y_out = y_out; a.flags = flags;
  const size_t smem = (size_t)nr * de * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(xattn_rank_fused_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(xattn_rank_fused_fwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(xattn_rank_fused_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  if (nr <= 8) xattn_rank_fused_fwd_kernel<8><<<B, XF_THREADS, smem, st>>>(a);
  else if (nr <= 12) xattn_rank_fused_fwd_kernel<12><<<B, XF_THREADS, smem, st>>>(a);
  else xattn_rank_fused_fwd_kernel<16><<<B, XF_THREADS, smem, st>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_rank_fused_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_xattn_rank_fused_bwd(const float* dy_out, const float* delta_y, const float* gamma, const float* y, int ldy,
                                           const float* r, int ldr, const float* probs, const uint8_t* m_txt, const float* e,
                                           int lde, int de, const float* wr, int ldwr, int B, int T, int H, int d, int C,
                                           float eps, float kappa, uint32_t drop_thr, uint64_t seed, float* de_out, int ldde,
                                           float* dy, float* dwr, float* small, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dy_out && delta_y && gamma && y && r && probs && m_txt && e && wr && de_out && dy && dwr && small && workspace,
                 "xattn_rank_fused_bwd: null pointer");
  IMMTSF_REQUIRE(fused_ok(T, H, d, C, de), "xattn_rank_fused_bwd: unsupported shape (T=%d H=%d d=%d C=%d de=%d)", T, H, d, C, de);
  const int nr = H * (2 * C + 1);
  IMMTSF_REQUIRE(lde >= de && (lde & 3) == 0 && ((uintptr_t)e & 15) == 0 && ldwr >= de && (ldwr & 3) == 0 && ((uintptr_t)wr & 15) == 0 &&
                     ldde >= de && (ldde & 3) == 0 && ((uintptr_t)de_out & 15) == 0 && ldy >= C && ldr >= nr,
                 "xattn_rank_fused_bwd: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  IMMTSF_REQUIRE(workspace_bytes >= immtsf_xattn_rank_fused_bwd_workspace_bytes(B, H, C, de) && ((uintptr_t)workspace & 15) == 0,
                 "xattn_rank_fused_bwd: workspace too small or misaligned");
  XfArgs a = {};
  a.e = e; a.lde = lde; a.de = de; a.wr = wr; a.ldwr = ldwr; a.y = y; a.ldy = ldy; a.gamma = gamma; a.m_txt = m_txt;
  a.B = B; a.T = T; a.H = H; a.C = C; a.scale = (float)sqrt(1.0 / (double)(d / H)); a.eps = eps; a.kappa = kappa; a.thr = drop_thr;
  a.seed = make_seed(seed); a.r = const_cast<float*>(r); a.ldr = ldr; a.delta_y = const_cast<float*>(delta_y);
  a.probs = const_cast<float*>(probs); a.dy_out = dy_out; a.de_out = de_out; a.ldde = ldde; a.dy = dy; a.partial = (float*)workspace;
  const int grid = fused_grid(B);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)nr * de * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(xattn_rank_fused_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(xattn_rank_fused_bwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(xattn_rank_fused_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  if (nr <= 8) xattn_rank_fused_bwd_kernel<8><<<grid, XF_THREADS, smem, st>>>(a);
  else if (nr <= 12) xattn_rank_fused_bwd_kernel<12><<<grid, XF_THREADS, smem, st>>>(a);
  else xattn_rank_fused_bwd_kernel<16><<<grid, XF_THREADS, smem, st>>>(a);
  IMMTSF_CHECK_LAUNCH("xattn_rank_fused_bwd");
  const int len = nr * de + nr + 3 * C;
  xattn_rank_partials_reduce_kernel<<<ceil_div(len, 32), dim3(32, 8), 0, st>>>((const float*)workspace, grid, len, part_stride(nr, de, C), nr * de, dwr, small);
  IMMTSF_CHECK_LAUNCH("xattn_rank_partials_reduce");
  return IMMTSF_OK;
}
