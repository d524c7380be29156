// K5: MMF_GR_Add (fusions/MMF_GR_Add.py:43-60).
//
// The input-side affine maps of the GRU (W_ih x + b_ih) and of the gate
// (W_g x + b_g) are ONE dense projection G4 = [Y;E] [W_ih;W_g]^T + [b_ih;b_g]
// done by immtsf_gemm, so E_txt is read once instead of twice.  What is left
// is (1) the T-step scan, one warp per sample with W_hh resident in shared
// memory (row stride C+1: conflict-free for both the forward mat-vec and the
// transposed one in backward), latency- not bandwidth-bound; (2) a per-row
// tail: residual_head (C x C), LayerNorm over C, dropout, sigmoid gate and
// the blend  Y_out = g*Y + (1-g)*(Y + delta).
//
// PyTorch GRU cell (gate order r,z,n):
//   r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r*gh_n),
//   h' = (1-z)*n + z*h,  gh = W_hh h + b_hh,  h_0 = 0.
// Backward returns the pre-activation gradients so that every weight gradient
// is a plain immtsf_gemm / immtsf_colsum over rows:
//   dG4[:, :3C] = [da_r, da_z, da_n]   (-> dW_ih, db_ih, dX)
//   dGh         = [da_r, da_z, da_n*r] (-> dW_hh = dGh^T h_prev, db_hh)
//   d_delta     = d(residual_head out) (-> dW_r = d_delta^T h, db_r)
#include "common.cuh"
#include "../../include/immtsf.h"

constexpr int GR_WARPS = 8;

// dynamic smem: W [rows][C+1] | per-warp scratch
template <int UN>
__global__ void __launch_bounds__(GR_WARPS * 32) gru_scan_fwd_kernel(const float* __restrict__ G4, const float* __restrict__ w_hh,
                                                                     const float* __restrict__ b_hh, int B, int T, int C,
                                                                     float* __restrict__ h_all, float* __restrict__ h_prev) {
  extern __shared__ float smem[];
  const int ldw = C + 1;
  float* s_w = smem;                       // [3C][C+1]
  float* s_b = s_w + 3 * C * ldw;          // [3C]
  float* s_h = s_b + 3 * C;                // [GR_WARPS][C]
  for (int i = threadIdx.x; i < 3 * C * C; i += blockDim.x) s_w[(i / C) * ldw + (i % C)] = w_hh[i];
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_b[i] = b_hh[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * GR_WARPS + w;
  if (b >= B) return;
  float* hs = s_h + w * C;
  float h[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) h[u] = 0.f;
  for (int j = lane; j < C; j += 32) hs[j] = 0.f;
  __syncwarp();
  const int ldg = 4 * C;
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)b * T + t;
    const float* gi = G4 + row * ldg;
    float hn_[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      hn_[u] = 0.f;
      if (j < C) {
        float gr = s_b[j], gz = s_b[C + j], gn = s_b[2 * C + j];
        const float* wr = s_w + (size_t)j * ldw;
        const float* wz = s_w + (size_t)(C + j) * ldw;
        const float* wn = s_w + (size_t)(2 * C + j) * ldw;
        for (int k = 0; k < C; ++k) {
          const float hk = hs[k];
          gr = fmaf(wr[k], hk, gr);
          gz = fmaf(wz[k], hk, gz);
          gn = fmaf(wn[k], hk, gn);
        }
        const float r = sigmoidf_(gi[j] + gr);
        const float z = sigmoidf_(gi[C + j] + gz);
        const float n = tanhf(gi[2 * C + j] + r * gn);
        h_prev[row * C + j] = h[u];
        hn_[u] = (1.f - z) * n + z * h[u];
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) {
        h[u] = hn_[u];
        hs[j] = h[u];
        h_all[row * C + j] = h[u];
      }
    }
    __syncwarp();
  }
}

template <int UN>
__global__ void __launch_bounds__(GR_WARPS * 32) gru_scan_bwd_kernel(const float* __restrict__ G4, const float* __restrict__ h_prev,
                                                                     const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                                                                     const float* __restrict__ dh_out, int B, int T, int C,
                                                                     float* __restrict__ dG4, float* __restrict__ dGh) {
  extern __shared__ float smem[];
  const int ldw = C + 1;
  float* s_w = smem;                       // [3C][C+1]
  float* s_b = s_w + 3 * C * ldw;          // [3C]
  float* s_h = s_b + 3 * C;                // [GR_WARPS][C]   h_{t-1}
  float* s_g = s_h + GR_WARPS * C;         // [GR_WARPS][3C]  dGh of this step
  for (int i = threadIdx.x; i < 3 * C * C; i += blockDim.x) s_w[(i / C) * ldw + (i % C)] = w_hh[i];
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_b[i] = b_hh[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * GR_WARPS + w;
  if (b >= B) return;
  float* hs = s_h + w * C;
  float* gs = s_g + w * 3 * C;
  float dh[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) dh[u] = 0.f;
  const int ldg = 4 * C;
  for (int t = T - 1; t >= 0; --t) {
    const size_t row = (size_t)b * T + t;
    const float* gi = G4 + row * ldg;
    for (int j = lane; j < C; j += 32) hs[j] = h_prev[row * C + j];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) {
        float gr = s_b[j], gz = s_b[C + j], gn = s_b[2 * C + j];
        const float* wr = s_w + (size_t)j * ldw;
        const float* wz = s_w + (size_t)(C + j) * ldw;
        const float* wn = s_w + (size_t)(2 * C + j) * ldw;
        for (int k = 0; k < C; ++k) {
          const float hk = hs[k];
          gr = fmaf(wr[k], hk, gr);
          gz = fmaf(wz[k], hk, gz);
          gn = fmaf(wn[k], hk, gn);
        }
        const float r = sigmoidf_(gi[j] + gr);
        const float z = sigmoidf_(gi[C + j] + gz);
        const float n = tanhf(gi[2 * C + j] + r * gn);
        const float hp = hs[j];
        const float dht = dh[u] + dh_out[row * C + j];
        const float dn = dht * (1.f - z);
        const float dz = dht * (hp - n);
        const float da_n = dn * (1.f - n * n);
        const float da_r = da_n * gn * r * (1.f - r);
        const float da_z = dz * z * (1.f - z);
        const float dhn = da_n * r;
        dh[u] = dht * z;  // direct path to h_{t-1}; the W_hh^T part is added below
        dG4[row * ldg + j] = da_r;
        dG4[row * ldg + C + j] = da_z;
        dG4[row * ldg + 2 * C + j] = da_n;
        dGh[row * 3 * C + j] = da_r;
        dGh[row * 3 * C + C + j] = da_z;
        dGh[row * 3 * C + 2 * C + j] = dhn;
        gs[j] = da_r;
        gs[C + j] = da_z;
        gs[2 * C + j] = dhn;
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int k = lane + u * 32;
      if (k < C) {
        float s = 0.f;
        for (int i = 0; i < 3 * C; ++i) s = fmaf(s_w[(size_t)i * ldw + k], gs[i], s);
        dh[u] += s;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ wide recurrence (C > 32)
// One warp per sample is a serial chain of 3C*C/32 FMAs per lane per step (8.6 us per step at C = 96).  For many
// channels the recurrence gets a whole CTA per sample: 3C threads, thread j keeps row j of W_hh in REGISTERS, the
// hidden state lives in shared memory (broadcast float4 reads), two barriers per step.  Forward also stores the
// gate activations (r, z, n, W_hn h + b_hn) so that backward needs only the transposed product, for which thread
// (g, k) keeps column k of gate block g in registers.
template <int CMAX>
__global__ void __launch_bounds__(3 * CMAX) gru_scan_fwd_wide_kernel(const float* __restrict__ G4, const float* __restrict__ w_hh,
                                                                      const float* __restrict__ b_hh, int B, int T, int C,
                                                                      float* __restrict__ h_all, float* __restrict__ h_prev,
                                                                      float* __restrict__ gates) {
  __shared__ __align__(16) float s_h[CMAX];
  __shared__ float s_gh[3 * CMAX];
  const int b = blockIdx.x, j = threadIdx.x;
  const bool act = j < 3 * C;
  float w[CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) w[k] = (act && k < C) ? __ldg(w_hh + (size_t)j * C + k) : 0.f;
  const float bj = act ? __ldg(b_hh + j) : 0.f;
  if (j < CMAX) s_h[j] = 0.f;
  __syncthreads();
  const int ldg = 4 * C;
  float h = 0.f;
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)b * T + t;
    float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f;
    if (j < C) {  // issued before the product so that their latency hides behind it
      gi_r = __ldg(G4 + row * ldg + j);
      gi_z = __ldg(G4 + row * ldg + C + j);
      gi_n = __ldg(G4 + row * ldg + 2 * C + j);
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(&s_h[k]);
      a0 = fmaf(w[k], hv.x, a0); a1 = fmaf(w[k + 1], hv.y, a1); a2 = fmaf(w[k + 2], hv.z, a2); a3 = fmaf(w[k + 3], hv.w, a3);
    }
    if (act) s_gh[j] = bj + ((a0 + a1) + (a2 + a3));
    __syncthreads();
    if (j < C) {
      const float gn = s_gh[2 * C + j];
      const float r = sigmoidf_(gi_r + s_gh[j]);
      const float z = sigmoidf_(gi_z + s_gh[C + j]);
      const float n = tanhf(gi_n + r * gn);
      h_prev[row * C + j] = h;
      h = (1.f - z) * n + z * h;
      h_all[row * C + j] = h;
      s_h[j] = h;
      float* gp = gates + row * ldg;
      gp[j] = r; gp[C + j] = z; gp[2 * C + j] = n; gp[3 * C + j] = gn;
    }
    __syncthreads();
  }
}

template <int CMAX>
__global__ void __launch_bounds__(3 * CMAX) gru_scan_bwd_wide_kernel(const float* __restrict__ gates, const float* __restrict__ h_prev,
                                                                      const float* __restrict__ w_hh, const float* __restrict__ dh_out,
                                                                      int B, int T, int C, float* __restrict__ dG4,
                                                                      float* __restrict__ dGh) {
  __shared__ float s_g[3 * CMAX];     // da_r, da_z, da_n * r of this step
  __shared__ float s_part[3 * CMAX];  // per gate block: (W_g^T dgate_g)[k]
  const int b = blockIdx.x, tid = threadIdx.x;
  const bool act = tid < 3 * C;
  const int g = act ? tid / C : 0, k = act ? tid % C : 0;
  float wt[CMAX];  // column k of gate block g
#pragma unroll
  for (int jj = 0; jj < CMAX; ++jj) wt[jj] = (act && jj < C) ? __ldg(w_hh + ((size_t)g * C + jj) * C + k) : 0.f;
  const int ldg = 4 * C;
  float dh = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    const size_t row = (size_t)b * T + t;
    float dh_direct = 0.f;
    if (tid < C) {
      const int j = tid;
      const float* gp = gates + row * ldg;
      const float r = gp[j], z = gp[C + j], n = gp[2 * C + j], gn = gp[3 * C + j];
      const float hp = h_prev[row * C + j];
      const float dht = dh + dh_out[row * C + j];
      const float dn = dht * (1.f - z);
      const float dz = dht * (hp - n);
      const float da_n = dn * (1.f - n * n);
      const float da_r = da_n * gn * r * (1.f - r);
      const float da_z = dz * z * (1.f - z);
      const float dhn = da_n * r;
      dh_direct = dht * z;
      dG4[row * ldg + j] = da_r;
      dG4[row * ldg + C + j] = da_z;
      dG4[row * ldg + 2 * C + j] = da_n;
      dGh[row * 3 * C + j] = da_r;
      dGh[row * 3 * C + C + j] = da_z;
      dGh[row * 3 * C + 2 * C + j] = dhn;
      s_g[j] = da_r;
      s_g[C + j] = da_z;
      s_g[2 * C + j] = dhn;
    }
    __syncthreads();
    if (act) {
      const float* sg = s_g + g * C;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int jj = 0; jj < CMAX; jj += 2) {
        a0 = fmaf(wt[jj], jj < C ? sg[jj] : 0.f, a0);
        a1 = fmaf(wt[jj + 1], jj + 1 < C ? sg[jj + 1] : 0.f, a1);
      }
      s_part[tid] = a0 + a1;
    }
    __syncthreads();
    if (tid < C) dh = dh_direct + s_part[tid] + s_part[C + tid] + s_part[2 * C + tid];
  }
}

// ------------------------------------------------------------------ tail
// one warp per (b,t) row, grid-stride over rows
template <int UN>
__global__ void __launch_bounds__(GR_WARPS * 32) gr_tail_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ G4,
                                                                    const float* __restrict__ h_all, const float* __restrict__ w_r,
                                                                    const float* __restrict__ b_r, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, const uint8_t* __restrict__ m_txt,
                                                                    int B, int T, int C, float eps, uint32_t thr, SeedArg seed_,
                                                                    float* __restrict__ Y_out, int32_t* __restrict__ flags) {
  extern __shared__ float smem[];
  const uint64_t seed = resolve_seed(seed_);
  const int ldw = C + 1;
  float* s_w = smem;                  // [C][C+1]
  float* s_h = s_w + C * ldw;         // [GR_WARPS][C]
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_w[(i / C) * ldw + (i % C)] = w_r[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* hs = s_h + w * C;
  const float inv_keep = thr == 0u ? 1.f : (float)(1.0 / (1.0 - (double)thr / 65536.0));
  const int rows = B * T;
  bool bad = false;
  for (int row = blockIdx.x * GR_WARPS + w; row < rows; row += gridDim.x * GR_WARPS) {
    const int b = row / T;
    __syncwarp();
    for (int j = lane; j < C; j += 32) hs[j] = h_all[(size_t)row * C + j];
    __syncwarp();
    float dl[UN];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      dl[u] = 0.f;
      if (j < C) {
        float acc = b_r[j];
        const float* wr = s_w + (size_t)j * ldw;
        for (int k = 0; k < C; ++k) acc = fmaf(wr[k], hs[k], acc);
        dl[u] = acc;
        s += acc;
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float v = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) v += (dl[u] - mu) * (dl[u] - mu);
    }
    const float rs = 1.f / sqrtf(warp_sum(v) / (float)C + eps);
    const bool has_txt = m_txt[b] != 0;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) {
        const float dn = (dl[u] - mu) * rs * gamma[j] + beta[j];
        const float dd = dn * dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, (uint64_t)row * C + j, thr, inv_keep);
        const float g = has_txt ? sigmoidf_(G4[(size_t)row * 4 * C + 3 * C + j]) : 1.f;
        const float y = Y[(size_t)row * C + j];
        const float o = g * y + (1.f - g) * (y + dd);
        Y_out[(size_t)row * C + j] = o;
        bad |= isnan(o);
      }
    }
  }
  if (flags != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) flags[IMMTSF_FLAG_OUT] = 1;
}

template <int UN>
__global__ void __launch_bounds__(GR_WARPS * 32) gr_tail_bwd_kernel(const float* __restrict__ dY_out, const float* __restrict__ G4,
                                                                    const float* __restrict__ h_all, const float* __restrict__ w_r,
                                                                    const float* __restrict__ b_r, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, const uint8_t* __restrict__ m_txt,
                                                                    int B, int T, int C, float eps, uint32_t thr, SeedArg seed_,
                                                                    float* __restrict__ dG4, float* __restrict__ d_delta,
                                                                    float* __restrict__ dh_out, float* __restrict__ dgamma,
                                                                    float* __restrict__ dbeta) {
  extern __shared__ float smem[];
  const uint64_t seed = resolve_seed(seed_);
  const int ldw = C + 1;
  float* s_w = smem;                          // [C][C+1]
  float* s_h = s_w + C * ldw;                 // [GR_WARPS][C]
  float* s_d = s_h + GR_WARPS * C;            // [GR_WARPS][C]
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_w[(i / C) * ldw + (i % C)] = w_r[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* hs = s_h + w * C;
  float* ds = s_d + w * C;
  const float inv_keep = thr == 0u ? 1.f : (float)(1.0 / (1.0 - (double)thr / 65536.0));
  const int rows = B * T;
  float dgam[UN], dbet[UN];
#pragma unroll
  for (int u = 0; u < UN; ++u) { dgam[u] = 0.f; dbet[u] = 0.f; }
  for (int row = blockIdx.x * GR_WARPS + w; row < rows; row += gridDim.x * GR_WARPS) {
    const int b = row / T;
    __syncwarp();
    for (int j = lane; j < C; j += 32) hs[j] = h_all[(size_t)row * C + j];
    __syncwarp();
    float dl[UN];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      dl[u] = 0.f;
      if (j < C) {
        float acc = b_r[j];
        const float* wr = s_w + (size_t)j * ldw;
        for (int k = 0; k < C; ++k) acc = fmaf(wr[k], hs[k], acc);
        dl[u] = acc;
        s += acc;
      }
    }
    const float mu = warp_sum(s) / (float)C;
    float v = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) v += (dl[u] - mu) * (dl[u] - mu);
    }
    const float rs = 1.f / sqrtf(warp_sum(v) / (float)C + eps);
    const bool has_txt = m_txt[b] != 0;
    float gg[UN], xh[UN];
    float p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      gg[u] = 0.f; xh[u] = 0.f;
      if (j < C) {
        const float dyo = dY_out[(size_t)row * C + j];
        const float x = (dl[u] - mu) * rs;
        const float ks = dropout_scale(seed, IMMTSF_SITE_MMF_DROPOUT, (uint64_t)row * C + j, thr, inv_keep);
        const float dn = x * gamma[j] + beta[j];
        const float dd = dn * ks;
        float g = 1.f, dlogit = 0.f;
        if (has_txt) {
          g = sigmoidf_(G4[(size_t)row * 4 * C + 3 * C + j]);
          dlogit = -dd * dyo * g * (1.f - g);
        }
        dG4[(size_t)row * 4 * C + 3 * C + j] = dlogit;
        const float d_dn = (1.f - g) * dyo * ks;
        dgam[u] = fmaf(d_dn, x, dgam[u]);
        dbet[u] += d_dn;
        gg[u] = d_dn * gamma[j];
        xh[u] = x;
        p1 += gg[u];
        p2 += gg[u] * x;
      }
    }
    const float m1 = warp_sum(p1) / (float)C, m2 = warp_sum(p2) / (float)C;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = lane + u * 32;
      if (j < C) {
        const float dd = rs * (gg[u] - m1 - xh[u] * m2);
        d_delta[(size_t)row * C + j] = dd;
        ds[j] = dd;
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int k = lane + u * 32;
      if (k < C) {
        float acc = 0.f;
        for (int j = 0; j < C; ++j) acc = fmaf(s_w[(size_t)j * ldw + k], ds[j], acc);
        dh_out[(size_t)row * C + k] = acc;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    const int j = lane + u * 32;
    if (j < C) {
      atomicAdd(dgamma + j, dgam[u]);
      atomicAdd(dbeta + j, dbet[u]);
    }
  }
}

// ---------------------------------------------------------------- launchers
#define GR_DISPATCH_UN(C, ...)                                 \
  do {                                                         \
    const int un__ = ((C) + 31) / 32;                          \
    if (un__ == 1) { constexpr int UN = 1; __VA_ARGS__; }      \
    else if (un__ == 2) { constexpr int UN = 2; __VA_ARGS__; } \
    else if (un__ == 3) { constexpr int UN = 3; __VA_ARGS__; } \
    else { constexpr int UN = 4; __VA_ARGS__; }                \
  } while (0)

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      immtsf_set_error("%s: cannot reserve %zu B of shared memory: %s", name, bytes, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return IMMTSF_ERR_UNSUPPORTED;
    }
  }
  return IMMTSF_OK;
}

extern "C" int immtsf_gru_scan_fwd(const float* G4, const float* w_hh, const float* b_hh, int B, int T, int C,
                                   float* h_all, float* h_prev, float* gates, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(G4 && w_hh && b_hh && h_all && h_prev, "gru_scan_fwd: null pointer");
  IMMTSF_REQUIRE(C >= 1 && C <= 128, "gru_scan_fwd: C=%d must be in [1,128]", C);
  if (gates != nullptr && C > 32) {  // wide recurrence: one CTA per sample
    cudaStream_t stw = (cudaStream_t)stream;
    if (C <= 64) gru_scan_fwd_wide_kernel<64><<<B, 192, 0, stw>>>(G4, w_hh, b_hh, B, T, C, h_all, h_prev, gates);
    else if (C <= 96) gru_scan_fwd_wide_kernel<96><<<B, 288, 0, stw>>>(G4, w_hh, b_hh, B, T, C, h_all, h_prev, gates);
    else gru_scan_fwd_wide_kernel<128><<<B, 384, 0, stw>>>(G4, w_hh, b_hh, B, T, C, h_all, h_prev, gates);
    IMMTSF_CHECK_LAUNCH("gru_scan_fwd_wide");
    return IMMTSF_OK;
  }
  const size_t smem = sizeof(float) * ((size_t)3 * C * (C + 1) + 3 * C + (size_t)GR_WARPS * C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(B, GR_WARPS);
  GR_DISPATCH_UN(C, {
    int rc = set_smem(gru_scan_fwd_kernel<UN>, smem, "gru_scan_fwd");
    if (rc) return rc;
    gru_scan_fwd_kernel<UN><<<grid, GR_WARPS * 32, smem, st>>>(G4, w_hh, b_hh, B, T, C, h_all, h_prev);
  });
  IMMTSF_CHECK_LAUNCH("gru_scan_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_gru_scan_bwd(const float* G4, const float* h_prev, const float* w_hh, const float* b_hh,
                                   const float* dh_out, const float* gates, int B, int T, int C, float* dG4, float* dGh,
                                   void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(G4 && h_prev && w_hh && b_hh && dh_out && dG4 && dGh, "gru_scan_bwd: null pointer");
  IMMTSF_REQUIRE(C >= 1 && C <= 128, "gru_scan_bwd: C=%d must be in [1,128]", C);
  if (gates != nullptr && C > 32) {
    cudaStream_t stw = (cudaStream_t)stream;
    if (C <= 64) gru_scan_bwd_wide_kernel<64><<<B, 192, 0, stw>>>(gates, h_prev, w_hh, dh_out, B, T, C, dG4, dGh);
    else if (C <= 96) gru_scan_bwd_wide_kernel<96><<<B, 288, 0, stw>>>(gates, h_prev, w_hh, dh_out, B, T, C, dG4, dGh);
    else gru_scan_bwd_wide_kernel<128><<<B, 384, 0, stw>>>(gates, h_prev, w_hh, dh_out, B, T, C, dG4, dGh);
    IMMTSF_CHECK_LAUNCH("gru_scan_bwd_wide");
    return IMMTSF_OK;
  }
  const size_t smem = sizeof(float) * ((size_t)3 * C * (C + 1) + 3 * C + (size_t)GR_WARPS * C + (size_t)GR_WARPS * 3 * C);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(B, GR_WARPS);
  GR_DISPATCH_UN(C, {
    int rc = set_smem(gru_scan_bwd_kernel<UN>, smem, "gru_scan_bwd");
    if (rc) return rc;
    gru_scan_bwd_kernel<UN><<<grid, GR_WARPS * 32, smem, st>>>(G4, h_prev, w_hh, b_hh, dh_out, B, T, C, dG4, dGh);
  });
  IMMTSF_CHECK_LAUNCH("gru_scan_bwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_gr_tail_fwd(const float* Y, const float* G4, const float* h_all, const float* w_r,
                                  const float* b_r, const float* gamma, const float* beta, const uint8_t* m_txt,
                                  int B, int T, int C, float eps, uint32_t drop_thr, uint64_t seed, float* Y_out,
                                  int32_t* flags, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(Y && G4 && h_all && w_r && b_r && gamma && beta && m_txt && Y_out, "gr_tail_fwd: null pointer");
  IMMTSF_REQUIRE(C >= 1 && C <= 128, "gr_tail_fwd: C=%d must be in [1,128]", C);
  const size_t smem = sizeof(float) * ((size_t)C * (C + 1) + (size_t)GR_WARPS * C);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(B * T, GR_WARPS);
  if (grid > 148 * 8) grid = 148 * 8;
  GR_DISPATCH_UN(C, {
    int rc = set_smem(gr_tail_fwd_kernel<UN>, smem, "gr_tail_fwd");
    if (rc) return rc;
    gr_tail_fwd_kernel<UN><<<grid, GR_WARPS * 32, smem, st>>>(Y, G4, h_all, w_r, b_r, gamma, beta, m_txt, B, T, C, eps,
                                                                drop_thr, make_seed(seed), Y_out, flags);
  });
  IMMTSF_CHECK_LAUNCH("gr_tail_fwd");
  return IMMTSF_OK;
}

extern "C" int immtsf_gr_tail_bwd(const float* dY_out, const float* G4, const float* h_all, const float* w_r,
                                  const float* b_r, const float* gamma, const float* beta, const uint8_t* m_txt,
                                  int B, int T, int C, float eps, uint32_t drop_thr, uint64_t seed, float* dG4,
                                  float* d_delta, float* dh_out, float* dgamma, float* dbeta, void* stream) {
  if (B == 0 || T == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(dY_out && G4 && h_all && w_r && b_r && gamma && beta && m_txt && dG4 && d_delta && dh_out && dgamma && dbeta,
                 "gr_tail_bwd: null pointer");
  IMMTSF_REQUIRE(C >= 1 && C <= 128, "gr_tail_bwd: C=%d must be in [1,128]", C);
  const size_t smem = sizeof(float) * ((size_t)C * (C + 1) + (size_t)2 * GR_WARPS * C);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(B * T, GR_WARPS);
  if (grid > 148 * 4) grid = 148 * 4;
  GR_DISPATCH_UN(C, {
    int rc = set_smem(gr_tail_bwd_kernel<UN>, smem, "gr_tail_bwd");
    if (rc) return rc;
    gr_tail_bwd_kernel<UN><<<grid, GR_WARPS * 32, smem, st>>>(dY_out, G4, h_all, w_r, b_r, gamma, beta, m_txt, B, T, C, eps,
                                                                drop_thr, make_seed(seed), dG4, d_delta, dh_out, dgamma, dbeta);
  });
  IMMTSF_CHECK_LAUNCH("gr_tail_bwd");
  return IMMTSF_OK;
}
