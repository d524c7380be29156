// Masked-MSE training loss: the step right after the fusion path.
//
// Reference: lib/evaluation.py:17-69 compute_error(truth, pred, mask, "MSE", "mean") called at :107-113 --
//   error = (truth - pred)^2 * mask;  err_c = sum over (sample, time);  cnt_c = sum of mask
//   loss  = sum_c err_c / (cnt_c + 1e-8) / count_nonzero(cnt)
// plus the B-iteration host loop of :128-132 that raises when a sample's mask is all zero.  The reference runs
// repeat / sub / pow / mul / two reshape-sums / div / count_nonzero / sum / div and B host syncs; here:
//   partial   one pass over pred / truth / mask: per-CTA partial sums [grid][2C] in a fixed order, then the LAST CTA
//             (ticket counter) adds the partials in CTA order -> err[C], cnt[C]: deterministic, one launch;
//             a sample whose mask is all zero sets a device flag (read by the caller when it wants the ValueError)
//   finalize  loss and the per-variable gradient scale 1 / ((cnt_c + 1e-8) n_avail) from err (local) and cnt (global:
//             under batch sharding the caller all-reduces the C counts in between, immtsf/dp.py)
//   bwd       dpred = gloss * 2 (pred - truth) mask * scale_c
// HBM-bound: 12 B per element forward, 16 B backward.
#include "common.cuh"
#include "../../include/immtsf.h"

namespace {

constexpr int LS_CMAX = 128;  // channels per pass (C <= 128 covers the path: MIMIC-shaped C = 96)

__global__ void __launch_bounds__(256) masked_mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                                                                 const float* __restrict__ mask, long rows, int C, int T,
                                                                 float* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                                 float* __restrict__ err_cnt, int32_t* __restrict__ empty_flag) {
  __shared__ float s_e[8][LS_CMAX], s_c[8][LS_CMAX];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // warp w of CTA b owns the samples b*8 + w, b*8 + w + 8*grid, ...: lanes stride over the T*C elements of a sample
  float e[LS_CMAX / 32], c[LS_CMAX / 32];
#pragma unroll
  for (int k = 0; k < LS_CMAX / 32; ++k) { e[k] = 0.f; c[k] = 0.f; }
  const long B = rows / T;
  for (long b = (long)blockIdx.x * 8 + w; b < B; b += (long)gridDim.x * 8) {
    float any = 0.f;
    for (int t = 0; t < T; ++t) {
      const long base = (b * T + t) * C;
#pragma unroll
      for (int k = 0; k < LS_CMAX / 32; ++k) {
        const int ch = lane + 32 * k;
        if (ch < C) {
          const float m = mask[base + ch], df = truth[base + ch] - pred[base + ch];
          e[k] = fmaf(df * df, m, e[k]);
          c[k] += m;
          any += m;
        }
      }
    }
    any = warp_sum(any);
    if (lane == 0 && any == 0.f && empty_flag != nullptr) *empty_flag = 1;  // lib/evaluation.py:128-132
  }
#pragma unroll
  for (int k = 0; k < LS_CMAX / 32; ++k) { s_e[w][lane + 32 * k] = e[k]; s_c[w][lane + 32 * k] = c[k]; }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float se = 0.f, sc = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { se += s_e[ww][ch]; sc += s_c[ww][ch]; }
    partial[(size_t)blockIdx.x * 2 * C + ch] = se;
    partial[(size_t)blockIdx.x * 2 * C + C + ch] = sc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {  // fixed CTA order: the result does not depend on scheduling
    float s = 0.f;
    for (unsigned int g = 0; g < gridDim.x; ++g) s += partial[(size_t)g * 2 * C + i];
    err_cnt[i] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch
}

__global__ void masked_mse_finalize_kernel(const float* __restrict__ err, const float* __restrict__ cnt, int C,
                                           float* __restrict__ loss, float* __restrict__ scale) {
  __shared__ float red[32];
  float s = 0.f, n = 0.f;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    s += err[ch] / (cnt[ch] + 1e-8f);
    n += cnt[ch] != 0.f ? 1.f : 0.f;
  }
  s = block_sum(s, red);
  n = block_sum(n, red);
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) scale[ch] = 1.f / ((cnt[ch] + 1e-8f) * n);
  if (threadIdx.x == 0) *loss = s / n;  // n == 0 (no observation at all) gives NaN, like the reference's 0 / 0
}

__global__ void __launch_bounds__(256) masked_mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                                                             const float* __restrict__ mask, size_t n, int C,
                                                             const float* __restrict__ scale, const float* __restrict__ gloss,
                                                             float* __restrict__ dpred) {
  const float g2 = 2.f * (gloss != nullptr ? *gloss : 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dpred[i] = g2 * (pred[i] - truth[i]) * mask[i] * scale[i % C];
}

}  // namespace

extern "C" size_t immtsf_masked_mse_workspace_bytes(int C) { return (size_t)(148 * 2) * 2 * C * sizeof(float) + 256; }

extern "C" int immtsf_masked_mse_partial(const float* pred, const float* truth, const float* mask, long rows, int T, int C,
                                         float* err_cnt, int32_t* empty_flag, unsigned int* ticket, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  IMMTSF_REQUIRE(pred && truth && mask && err_cnt && ticket, "masked_mse_partial: null pointer");
  IMMTSF_REQUIRE(C >= 1 && C <= LS_CMAX && T >= 1 && rows >= 0 && rows % T == 0, "masked_mse_partial: need 1 <= C <= 128 and rows %% T == 0");
  IMMTSF_REQUIRE(workspace != nullptr && workspace_bytes >= immtsf_masked_mse_workspace_bytes(C), "masked_mse_partial: workspace too small");
  const long B = rows / T;
  int grid = (int)((B + 7) / 8);
  if (grid > 148 * 2) grid = 148 * 2;
  if (grid < 1) grid = 1;
  float* partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  masked_mse_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, truth, mask, rows, C, T, partial, ticket, err_cnt, empty_flag);
  IMMTSF_CHECK_LAUNCH("masked_mse_partial");
  return IMMTSF_OK;
}

extern "C" int immtsf_masked_mse_finalize(const float* err, const float* cnt, int C, float* loss, float* scale, void* stream) {
  IMMTSF_REQUIRE(err && cnt && loss && scale && C >= 1, "masked_mse_finalize: bad args");
  masked_mse_finalize_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(err, cnt, C, loss, scale);
  IMMTSF_CHECK_LAUNCH("masked_mse_finalize");
  return IMMTSF_OK;
}

extern "C" int immtsf_masked_mse_bwd(const float* pred, const float* truth, const float* mask, long rows, int C,
                                     const float* scale, const float* gloss, float* dpred, void* stream) {
  if (rows == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(pred && truth && mask && scale && dpred && C >= 1, "masked_mse_bwd: bad args");
  const size_t n = (size_t)rows * C;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  masked_mse_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, truth, mask, n, C, scale, gloss, dpred);
  IMMTSF_CHECK_LAUNCH("masked_mse_bwd");
  return IMMTSF_OK;
}
