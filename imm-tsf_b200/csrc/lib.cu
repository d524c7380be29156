// Library-level entry points: version, error string, device capability.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "../../include/immtsf.h"

static thread_local char g_err[512] = "";

void immtsf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void immtsf_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long immtsf_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// device-resident dropout seed offset, added to every seed argument when set.  Process-wide (one process per
// GPU): forward runs on the caller's thread, backward on autograd's, and both must see the same pointer.
static const uint64_t* g_seed_dev = nullptr;
const uint64_t* immtsf_seed_dev() { return g_seed_dev; }
extern "C" int immtsf_set_seed_offset_ptr(const uint64_t* dev_ptr) {
  g_seed_dev = dev_ptr;
  return IMMTSF_OK;
}
__global__ void seed_advance_kernel(uint64_t* p, uint64_t inc) { *p += inc; }
extern "C" int immtsf_seed_advance(uint64_t* dev_ptr, uint64_t inc, void* stream) {
  IMMTSF_REQUIRE(dev_ptr != nullptr, "seed_advance: null pointer");
  seed_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_ptr, inc);
  IMMTSF_CHECK_LAUNCH("seed_advance");
  return IMMTSF_OK;
}

extern "C" int immtsf_version(void) { return IMMTSF_ABI_VERSION; }
extern "C" const char* immtsf_last_error_string(void) { return g_err; }

extern "C" int immtsf_device_supported(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    immtsf_set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return IMMTSF_ERR_ARCH;
  }
  return p.major == 10 ? 1 : 0;
}

// ----------------------------------------------------------------- helpers
__global__ void axpby_kernel(const float* __restrict__ x, float alpha, float* __restrict__ y, int acc, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = acc ? fmaf(alpha, x[i], y[i]) : alpha * x[i];
}
extern "C" int immtsf_axpby(const float* x, float alpha, float* y, int accumulate, size_t n, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(x && y, "axpby: null pointer");
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  axpby_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, alpha, y, accumulate, n);
  IMMTSF_CHECK_LAUNCH("axpby");
  return IMMTSF_OK;
}

__global__ void group_sum_rows_kernel(const float* __restrict__ x, int ldx, int R, int T, int d, float* __restrict__ out) {
  const int r = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  float s = 0.f;
  const float* p = x + (size_t)r * T * ldx + c;
  for (int t = 0; t < T; ++t) s += p[(size_t)t * ldx];
  out[(size_t)r * d + c] = s;
}
extern "C" int immtsf_group_sum_rows(const float* x, int ldx, int R, int T, int d, float* out, void* stream) {
  if (R == 0 || d == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(x && out && T > 0, "group_sum_rows: bad args");
  dim3 grid(ceil_div(d, 256), R);
  group_sum_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, R, T, d, out);
  IMMTSF_CHECK_LAUNCH("group_sum_rows");
  return IMMTSF_OK;
}

__global__ void nan_check_kernel(const float* __restrict__ x, size_t n, int32_t* flags, int slot) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < n; i += stride) bad |= isnan(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) flags[slot] = 1;
}
extern "C" int immtsf_nan_check(const float* x, size_t n, int32_t* flags, int slot, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(x && flags && slot >= 0 && slot < 4, "nan_check: bad args");
  int grid = (int)((n + 1023) / 1024);
  if (grid > 148 * 8) grid = 148 * 8;
  nan_check_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, flags, slot);
  IMMTSF_CHECK_LAUNCH("nan_check");
  return IMMTSF_OK;
}

__global__ void zero_pad_rows_kernel(float* X, int ld, int ncols, const int32_t* m_dev, int M_alloc) {
  const int m = ragged_rows(M_alloc, m_dev);
  int end = (m + 127) / 128 * 128;
  if (end > M_alloc) end = M_alloc;
  const int r = m + blockIdx.x;
  if (r >= end) return;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) X[(size_t)r * ld + c] = 0.f;
}
extern "C" int immtsf_zero_pad_rows(float* X, int ld, int ncols, const int32_t* m_dev, int M_alloc, void* stream) {
  IMMTSF_REQUIRE(X && m_dev, "zero_pad_rows: null pointer");
  zero_pad_rows_kernel<<<128, 256, 0, (cudaStream_t)stream>>>(X, ld, ncols, m_dev, M_alloc);
  IMMTSF_CHECK_LAUNCH("zero_pad_rows");
  return IMMTSF_OK;
}
