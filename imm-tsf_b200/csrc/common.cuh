// Shared device/host helpers for the immtsf sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define IMMTSF_OK 0
#define IMMTSF_ERR_ARG (-1)
#define IMMTSF_ERR_UNSUPPORTED (-2)
#define IMMTSF_ERR_LAUNCH (-3)
#define IMMTSF_ERR_ARCH (-4)

void immtsf_set_error(const char* fmt, ...);
void immtsf_count_launch();  // host-side counter of kernel launches (bench.py's gpu_launches)

#define IMMTSF_REQUIRE(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      immtsf_set_error(__VA_ARGS__);         \
      return IMMTSF_ERR_ARG;                 \
    }                                        \
  } while (0)

#define IMMTSF_CHECK_LAUNCH(name)                                             \
  do {                                                                        \
    immtsf_count_launch();                                                    \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      immtsf_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return IMMTSF_ERR_LAUNCH;                                               \
    }                                                                         \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Grid size of a persistent (grid-stride) kernel: never more CTAs than can be resident at once -- a partial second wave
// makes the launch take two full CTA lifetimes (recavg_bwd_rows_w<3>: 159 registers = 3 CTAs/SM resident, but 4 per SM
// were launched: 1.33 waves in ncu).  The occupancy query is cached per kernel.
static inline int resident_grid(const void* kernel, int threads, size_t smem, int want, int max_per_sm) {
  static const void* keys[64];
  static int vals[64];
  static int n = 0;
  int per_sm = 0;
  for (int i = 0; i < n; ++i)
    if (keys[i] == kernel) per_sm = vals[i];
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (n < 64) { keys[n] = kernel; vals[n] = per_sm; ++n; }
  }
  if (per_sm > max_per_sm) per_sm = max_per_sm;
  const int cap = 148 * per_sm;
  return want < cap ? want : cap;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG for dropout: the mask is a pure function of
// (seed, site, element index), so backward regenerates it instead of storing it
// and the tests can rebuild the same mask on the host (tests/philox_ref.py).
struct Philox4 {
  uint32_t x, y, z, w;
};
__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  // one IMAD.WIDE.U32 (the C++ form below compiles to IMAD.WIDE plus an add of a zero high word per product: 20 wasted
  // instructions per Philox call, ~7 % of the instructions of the RecAvg backward's rows phase)
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
#else
  const uint64_t p = (uint64_t)a * (uint64_t)b;
#endif
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
// Dropout decision for flat element index `idx` at dropout site `site`: one Philox call serves EIGHT elements
// (16 random bits each): call counter = (idx>>3 lo, idx>>3 hi, site, 0), key = seed; element idx uses the 16-bit
// field (idx & 1) of word (idx>>1) & 3.  An element is dropped iff field < thr, thr = floor(p * 2^16); the
// realised rate thr / 2^16 differs from p by < 1.6e-5 and inv_keep uses the realised rate, so E[mask] = 1 exactly.
__device__ __forceinline__ uint32_t dropout_field(uint64_t seed, uint32_t site, uint64_t idx) {
  const uint64_t c = idx >> 3;
  const Philox4 r = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), site, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t k = (uint32_t)(idx >> 1) & 3u;
  const uint32_t w = k == 0 ? r.x : (k == 1 ? r.y : (k == 2 ? r.z : r.w));
  return (idx & 1) ? (w >> 16) : (w & 0xFFFFu);
}
// keep-scale for one element: 0 if dropped else 1/(1-p).
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t site, uint64_t idx, uint32_t thr, float inv_keep) {
  if (thr == 0u) return 1.f;
  return dropout_field(seed, site, idx) >= thr ? inv_keep : 0.f;
}
// keep-scales of the 8 consecutive elements idx8*8 .. idx8*8+7 (one Philox call)
__host__ __device__ __forceinline__ void dropout_scale8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thr, float inv_keep,
                                                        float (&o)[8]) {
  if (thr == 0u) {
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 1.f;
    return;
  }
  const Philox4 r = philox4x32_10((uint32_t)idx8, (uint32_t)(idx8 >> 32), site, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = (w[k] & 0xFFFFu) >= thr ? inv_keep : 0.f;
    o[2 * k + 1] = (w[k] >> 16) >= thr ? inv_keep : 0.f;
  }
}

// Round keys of Philox4x32-10 for one seed (the Weyl sequence of common.cuh's philox4x32_10), computed once per kernel:
// the key schedule is 20 of the ~85 instructions of a call when it is redone per call.
struct PhiloxKeys { uint32_t k0[10], k1[10]; };
__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  return k;
}
// keep-scales of chunk idx8 with the halves in the lane's order: o[0..3] = the float4 the lane reads first (p = 1: the second
// one).  Same mask as dropout_scale8 (one Philox4x32-10 call, 16-bit fields against thr); thr == 0 keeps everything
// (every field >= 0) without a branch.
__host__ __device__ __forceinline__ void dropout_scale8_sw(const PhiloxKeys& key, uint32_t site, uint64_t idx8, uint32_t thr, float inv_keep, int p,
                                                  float (&o)[8]) {
  uint32_t c0 = (uint32_t)idx8, c1 = (uint32_t)(idx8 >> 32), c2 = site, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    const uint32_t n0 = hi1 ^ c1 ^ key.k0[r], n2 = hi0 ^ c3 ^ key.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  const uint32_t w[4] = {p ? c2 : c0, p ? c3 : c1, p ? c0 : c2, p ? c1 : c3};
  const uint32_t thr_hi = thr << 16;  // (w >> 16) >= thr  <=>  w >= thr << 16 (thr <= 0xFFFF)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[2 * k] = (w[k] << 16) >= thr_hi ? inv_keep : 0.f;
    o[2 * k + 1] = w[k] >= thr_hi ? inv_keep : 0.f;
  }
}

// Dropout seed as the kernels see it: a host value plus an optional device-resident offset.  The offset lets a
// CUDA graph that captured one training step draw fresh masks on every replay (immtsf_set_seed_offset_ptr).
struct SeedArg {
  uint64_t base;
  const uint64_t* dev;
};
__device__ __forceinline__ uint64_t resolve_seed(const SeedArg& s) { return s.dev != nullptr ? s.base + *s.dev : s.base; }
const uint64_t* immtsf_seed_dev();
static inline SeedArg make_seed(uint64_t v) {
  SeedArg s;
  s.base = v;
  s.dev = immtsf_seed_dev();
  return s;
}

#define IMMTSF_SITE_TTF_DROPOUT 1u
#define IMMTSF_SITE_TTF_ATTN 2u
#define IMMTSF_SITE_MMF_DROPOUT 3u
#define IMMTSF_SITE_MMF_ATTN 4u

// NaN flag slots (int32[4], set with atomicOr-free plain stores of 1).
#define IMMTSF_FLAG_V 0
#define IMMTSF_FLAG_Y 1
#define IMMTSF_FLAG_E 2
#define IMMTSF_FLAG_OUT 3

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ragged row count: min(M, *m_dev) when m_dev != nullptr
__device__ __forceinline__ int ragged_rows(int M, const int32_t* __restrict__ m_dev) {
  if (m_dev == nullptr) return M;
  const int v = *m_dev;
  return v < M ? v : M;
}
