// Text-embedding store kernels: the reference keeps, per record, a Python list of
// (rel_time, embedding row) tuples loaded from the per-record .pt file (lib/parse_datasets.py:132-147), filters that
// list with a Python comprehension for every chunk window (:204-209) and pads / stacks the selected rows again for every
// batch (multimodal_collate, :786-819).  Here all embedding rows of all records live once in HBM
// (emb_all [sumN, d_m], rel_all [sumN], entity_offsets [E+1]); the window filter runs once per dataset as two kernels
// (count, then an order-preserving fill after an exclusive scan) and produces a CSR over chunks (chunk_offsets,
// chunk_rows, chunk_tau); a batch is one gather launch from the resident store straight into the ragged layout the
// fusion kernels consume.  Integer work (counts, offsets, selected rows) is bit-exact against the reference's filter;
// tau = fp32(double(rel) - st) reproduces the reference's double subtraction followed by torch.tensor(..., float32).
#include "common.cuh"
#include "../../include/immtsf.h"

// in-window test of the reference: st <= t < hist_end, evaluated in double (Python floats)
__device__ __forceinline__ bool in_window(float rel, double st, double he) {
  const double t = (double)rel;
  return st <= t && t < he;
}

// one warp per chunk
__global__ void window_count_kernel(const float* __restrict__ rel_all, const int32_t* __restrict__ eo,
                                    const int32_t* __restrict__ ent, const double* __restrict__ st,
                                    const double* __restrict__ he, int n, int32_t* __restrict__ counts) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int e = ent[i];
  const int j0 = eo[e], j1 = eo[e + 1];
  const double s = st[i], h = he[i];
  int c = 0;
  for (int j = j0 + lane; j < j1; j += 32) c += in_window(rel_all[j], s, h) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[i] = c;
}

// offsets[0] = 0, offsets[i+1] = sum_{j<=i} counts[j]; single CTA of 1024 threads, carry across chunks of 1024
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ offsets,
                                                            int32_t* __restrict__ overflow) {
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_carry = 0; offsets[0] = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    long long v = i < n ? (long long)counts[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) s_w[w] = v;
    __syncthreads();
    if (w == 0) {
      long long x = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += u;
      }
      s_w[lane] = x;
    }
    __syncthreads();
    const long long incl = v + (w > 0 ? s_w[w - 1] : 0) + s_carry;
    if (i < n) {
      if (incl > 0x7fffffffLL) *overflow = 1;
      offsets[i + 1] = (int32_t)incl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
}

// one warp per chunk: order-preserving (file order, as the reference's list comprehension) compaction
__global__ void window_fill_kernel(const float* __restrict__ rel_all, const int32_t* __restrict__ eo,
                                   const int32_t* __restrict__ ent, const double* __restrict__ st,
                                   const double* __restrict__ he, int n, const int32_t* __restrict__ chunk_offsets,
                                   int32_t* __restrict__ chunk_rows, float* __restrict__ chunk_tau) {
  const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const int e = ent[i];
  const int j0 = eo[e], j1 = eo[e + 1];
  const double s = st[i], h = he[i];
  int base = chunk_offsets[i];
  for (int jb = j0; jb < j1; jb += 32) {
    const int j = jb + lane;
    const float rel = j < j1 ? rel_all[j] : 0.f;
    const bool in = j < j1 && in_window(rel, s, h);
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      chunk_rows[pos] = j;
      chunk_tau[pos] = (float)((double)rel - s);  // (t - st) in double, then float32 (parse_datasets.py:205, :786-790)
    }
    base += __popc(m);
  }
}

// grid (ceil(N_max / 8), B), 8 warps: warp w copies note blockIdx.x*8 + w of sample blockIdx.y
__global__ void __launch_bounds__(256) batch_gather_kernel(const float* __restrict__ emb_all, int ld, int d_m,
                                                           const int32_t* __restrict__ chunk_offsets,
                                                           const int32_t* __restrict__ chunk_rows,
                                                           const float* __restrict__ chunk_tau,
                                                           const int32_t* __restrict__ chunk_ids,
                                                           const int32_t* __restrict__ offsets, float* __restrict__ emb_flat,
                                                           int ld_out, float* __restrict__ tau_flat, int vec) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int c = chunk_ids[b];
  const int c0 = chunk_offsets[c], cnt = chunk_offsets[c + 1] - c0;
  if (r >= cnt) return;
  const int src = chunk_rows[c0 + r], dst = offsets[b] + r;
  const float* s = emb_all + (size_t)src * ld;
  float* o = emb_flat + (size_t)dst * ld_out;
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int k = lane; k < (d_m >> 2); k += 32) o4[k] = __ldg(s4 + k);
  } else {
    for (int k = lane; k < d_m; k += 32) o[k] = s[k];
  }
  if (lane == 0) tau_flat[dst] = chunk_tau[c0 + r];
}

extern "C" int immtsf_window_count(const float* rel_all, const int32_t* entity_offsets, const int32_t* ent, const double* st,
                                   const double* hist_end, int n, int32_t* counts, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(rel_all && entity_offsets && ent && st && hist_end && counts, "window_count: null pointer");
  window_count_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(rel_all, entity_offsets, ent, st, hist_end, n, counts);
  IMMTSF_CHECK_LAUNCH("window_count");
  return IMMTSF_OK;
}

extern "C" int immtsf_exclusive_scan_i32(const int32_t* counts, int n, int32_t* offsets, int32_t* overflow_flag, void* stream) {
  IMMTSF_REQUIRE(offsets && overflow_flag && (counts || n == 0), "exclusive_scan: null pointer");
  IMMTSF_REQUIRE(n >= 0, "exclusive_scan: n must be >= 0");
  exclusive_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(counts, n, offsets, overflow_flag);
  IMMTSF_CHECK_LAUNCH("exclusive_scan");
  return IMMTSF_OK;
}

extern "C" int immtsf_window_fill(const float* rel_all, const int32_t* entity_offsets, const int32_t* ent, const double* st,
                                  const double* hist_end, int n, const int32_t* chunk_offsets, int32_t* chunk_rows,
                                  float* chunk_tau, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(rel_all && entity_offsets && ent && st && hist_end && chunk_offsets && chunk_rows && chunk_tau, "window_fill: null pointer");
  window_fill_kernel<<<ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(rel_all, entity_offsets, ent, st, hist_end, n, chunk_offsets,
                                                                     chunk_rows, chunk_tau);
  IMMTSF_CHECK_LAUNCH("window_fill");
  return IMMTSF_OK;
}

extern "C" int immtsf_batch_gather(const float* emb_all, int ld, int d_m, const int32_t* chunk_offsets, const int32_t* chunk_rows,
                                   const float* chunk_tau, const int32_t* chunk_ids, const int32_t* offsets, int B, int N_max,
                                   float* emb_flat, int ld_out, float* tau_flat, void* stream) {
  if (B == 0 || N_max == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(emb_all && chunk_offsets && chunk_rows && chunk_tau && chunk_ids && offsets && emb_flat && tau_flat, "batch_gather: null pointer");
  IMMTSF_REQUIRE(d_m >= 1 && ld >= d_m && ld_out >= d_m, "batch_gather: leading dimensions must cover d_model");
  IMMTSF_REQUIRE(B <= 65535, "batch_gather: B <= 65535");
  const int vec = ((uintptr_t)emb_all & 15) == 0 && ((uintptr_t)emb_flat & 15) == 0 && (ld & 3) == 0 && (ld_out & 3) == 0 && (d_m & 3) == 0;
  batch_gather_kernel<<<dim3(ceil_div(N_max, 8), B), 256, 0, (cudaStream_t)stream>>>(emb_all, ld, d_m, chunk_offsets, chunk_rows, chunk_tau,
                                                                                   chunk_ids, offsets, emb_flat, ld_out, tau_flat, vec);
  IMMTSF_CHECK_LAUNCH("batch_gather");
  return IMMTSF_OK;
}
