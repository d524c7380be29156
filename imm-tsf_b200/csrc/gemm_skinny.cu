// Skinny projections: one of M, N, K is at most 16 (the channel count C, or a
// rank-C fold of two weights).  The path has many of them -- MMF_XAttn_Add's
// proj_q (K = C) and residual_head (N = C), MMF_GR_Add's recurrent weights,
// every bias gradient (a column sum = a product with a ones vector) -- and a
// 64x64 FFMA tile wastes >90 % of its lanes on them.  They are pure HBM
// streams (the wide operand is read or written exactly once), so each shape
// class gets a kernel whose memory access is one coalesced pass:
//
//   small-K   C[M,N] = A'[M,K<=16] B'[K,N]        write-bound: a thread keeps its
//             K x 4 slice of B' in registers and streams rows of C (float4).
//   small-N   C[R,n<=16] = A[R,K] B'[K,n]          read-bound: a warp streams 4 rows
//             of A (float4 along k), B' comes from L1, warp-shuffle reduction.
//   tall-T    C[s<=16, N] = sum_k S'[k,s] W[k,N]   (weight / bias gradients: the
//             contraction runs over the rows).  CTAs split the rows, partial
//             sums go to the workspace and a second launch adds them in a fixed
//             order (deterministic -- no atomics).  S' == nullptr means a ones
//             vector: that is immtsf_colsum.
//
// Operands are addressed through (row stride, column stride) pairs so that
// every transposition case maps onto one of the three kernels.
// Algorithmic HBM bytes: 4 * (M*N + M*K + K*N), dominated by the wide operand.
#include "common.cuh"
#include "../../include/immtsf.h"

namespace {

struct SkArgs {
  int M, N, K;
  float alpha, beta;
  const float* A; long a_rs, a_cs;   // A'(m,k) = A[m*a_rs + k*a_cs]
  const float* B; long b_rs, b_cs;   // B'(k,n) = B[k*b_rs + n*b_cs]
  float* C; long c_rs, c_cs;         // C(m,n)  = C[m*c_rs + n*c_cs]
  const float* bias;                 // indexed by n (or by m when bias_on_m)
  int bias_on_m;
  const int32_t* ragged; int ragged_dim;
};

// ------------------------------------------------------------------ small-K
// requires c_cs == 1, C 16B-aligned rows (c_rs % 4 == 0); grid (n tiles of blockDim*4, row chunks)
template <int KMAX>
__global__ void __launch_bounds__(256) gemm_smallk_kernel(const SkArgs g) {
  int Meff = g.M;
  if (g.ragged_dim == 1) Meff = ragged_rows(g.M, g.ragged);
  int Mtouch = (Meff + 127) / 128 * 128;
  if (Mtouch > g.M) Mtouch = g.M;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (n >= g.N) return;
  const int nv = min(4, g.N - n);
  float b[KMAX][4];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) b[k][e] = (k < g.K && e < nv) ? __ldg(g.B + k * g.b_rs + (long)(n + e) * g.b_cs) : 0.f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.bias != nullptr)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < nv) bs[e] = __ldg(g.bias + n + e);
  // whole float4 groups of the A row may be read when the row (incl. its padding up to a multiple of 4) is in bounds
  const bool avec = KMAX >= 4 && g.a_cs == 1 && (g.a_rs & 3) == 0 && (((uintptr_t)g.A) & 15) == 0 && g.a_rs >= ((g.K + 3) & ~3);
  for (int m = blockIdx.y; m < Mtouch; m += gridDim.y) {
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = m < Meff;
    if (live) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (avec) {  // A rows are contiguous and 16B aligned: KMAX/4 broadcast float4 loads instead of KMAX scalar ones
#pragma unroll
        for (int k4 = 0; k4 < KMAX / 4; ++k4) {
          if (k4 * 4 < g.K) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(g.A + (long)m * g.a_rs) + k4);
            const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[e] = fmaf(a4[kk], b[k4 * 4 + kk][e], acc[e]);  // b[k] = 0 for k >= K
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < g.K) {
            const float a = __ldg(g.A + (long)m * g.a_rs + k * g.a_cs);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = fmaf(a, b[k][e], acc[e]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = g.alpha * acc[e] + bs[e];
    }
    float* cp = g.C + (long)m * g.c_rs + n;
    if (nv == 4) {
      float4 ov = make_float4(o[0], o[1], o[2], o[3]);
      if (g.beta != 0.f && live) {
        const float4 c = *reinterpret_cast<const float4*>(cp);
        ov.x = fmaf(g.beta, c.x, ov.x); ov.y = fmaf(g.beta, c.y, ov.y);
        ov.z = fmaf(g.beta, c.z, ov.z); ov.w = fmaf(g.beta, c.w, ov.w);
      }
      *reinterpret_cast<float4*>(cp) = ov;
    } else {
      for (int e = 0; e < nv; ++e) cp[e] = (g.beta != 0.f && live) ? fmaf(g.beta, cp[e], o[e]) : o[e];
    }
  }
}

// ------------------------------------------------------------------ small-N
// A rows are k-contiguous (a_cs == 1), 16B aligned (a_rs % 4 == 0, K % 4 == 0).
// One warp owns SN_R consecutive rows; lanes stride float4 chunks of k.
// SMEMB: B' is staged once per CTA as s_b[n][K] (k contiguous), so a lane fetches its four k's of column n with ONE
// conflict-free LDS.128 instead of four scalar global loads (NMAX = 16: 64 -> 16 load instructions per 128 FMA).
template <int NMAX, int SN_R, bool SMEMB = false>
__global__ void __launch_bounds__(256) gemm_smalln_kernel(const SkArgs g) {
  extern __shared__ __align__(16) float s_b[];
  if (SMEMB) {
    for (int i = threadIdx.x; i < g.N * g.K; i += blockDim.x) {
      const int n = i / g.K, k = i % g.K;
      s_b[i] = __ldg(g.B + (long)k * g.b_rs + (long)n * g.b_cs);
    }
    __syncthreads();
  }
  int Meff = g.M;
  if (g.ragged_dim == 1) Meff = ragged_rows(g.M, g.ragged);
  int Mtouch = (Meff + 127) / 128 * 128;
  if (Mtouch > g.M) Mtouch = g.M;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int K4 = g.K >> 2;
  for (int r0 = warp * SN_R; r0 < Mtouch; r0 += nwarps * SN_R) {
    float acc[SN_R][NMAX];
#pragma unroll
    for (int r = 0; r < SN_R; ++r)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) acc[r][n] = 0.f;
    if (r0 < Meff) {
      for (int k4 = lane; k4 < K4; k4 += 32) {
        float4 a[SN_R];
#pragma unroll
        for (int r = 0; r < SN_R; ++r)
          a[r] = (r0 + r < Meff) ? __ldg(reinterpret_cast<const float4*>(g.A + (long)(r0 + r) * g.a_rs) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* bp = g.B + (long)(k4 * 4) * g.b_rs;
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
          if (n < g.N) {
            float b0, b1, b2, b3;
            if (SMEMB) {
              const float4 bv = *reinterpret_cast<const float4*>(s_b + n * g.K + k4 * 4);
              b0 = bv.x; b1 = bv.y; b2 = bv.z; b3 = bv.w;
            } else {
              b0 = __ldg(bp + n * g.b_cs); b1 = __ldg(bp + g.b_rs + n * g.b_cs);
              b2 = __ldg(bp + 2 * g.b_rs + n * g.b_cs); b3 = __ldg(bp + 3 * g.b_rs + n * g.b_cs);
            }
#pragma unroll
            for (int r = 0; r < SN_R; ++r)
              acc[r][n] = fmaf(a[r].x, b0, fmaf(a[r].y, b1, fmaf(a[r].z, b2, fmaf(a[r].w, b3, acc[r][n]))));
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SN_R; ++r)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) acc[r][n] = warp_sum(acc[r][n]);
    // lane l writes element (r = l / NMAX, n = l % NMAX)
#pragma unroll
    for (int r = 0; r < SN_R; ++r)
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        if (lane == ((r * NMAX + n) & 31) && n < g.N && r0 + r < Mtouch) {
          const int m = r0 + r;
          float* cp = g.C + (long)m * g.c_rs + (long)n * g.c_cs;
          float x = 0.f;
          if (m < Meff) {
            x = g.alpha * acc[r][n];
            if (g.bias != nullptr) x += __ldg(g.bias + (g.bias_on_m ? m : n));
            if (g.beta != 0.f) x = fmaf(g.beta, *cp, x);
          }
          *cp = x;
        }
      }
  }
}

// ------------------------------------------------------------------ tall-T (split rows)
// partial[split][s][n] = sum_{k in split} S'(k,s) * W[k*w_rs + n]    (W n-contiguous, 16B aligned)
// grid (n tiles of 128 floats, splits); 256 threads = 8 warps, warp w takes rows k0+w, k0+w+8, ...
constexpr int TT_WARPS = 8;
template <int SMAX>
__global__ void __launch_bounds__(256) gemm_tallt_partial_kernel(const float* __restrict__ S, long s_ks, long s_ss, int Sn,
                                                                 const float* __restrict__ W, long w_rs, int N, int K,
                                                                 const int32_t* __restrict__ ragged, float* __restrict__ partial,
                                                                 int ldp) {
  extern __shared__ float4 s_red[];  // [TT_WARPS][SMAX][32]
  const int Keff = ragged_rows(K, ragged);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + lane) * 4;
  const int nsplit = gridDim.y;
  const int per = (Keff + nsplit - 1) / nsplit;
  const int k0 = blockIdx.y * per, k1 = min(Keff, k0 + per);
  float4 acc[SMAX];
#pragma unroll
  for (int s = 0; s < SMAX; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool svec = S != nullptr && s_ss == 1 && (s_ks & 3) == 0 && (((uintptr_t)S) & 15) == 0 && s_ks >= ((Sn + 3) & ~3);
  if (n < N) {
    const bool full = n + 3 < N;
    for (int kb = k0 + w; kb < k1; kb += 4 * TT_WARPS) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // 4 independent rows in flight
        const int k = kb + u * TT_WARPS;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < k1) {
          const float* wp = W + (long)k * w_rs + n;
          if (full) v[u] = __ldg(reinterpret_cast<const float4*>(wp));
          else {
            v[u].x = wp[0]; v[u].y = n + 1 < N ? wp[1] : 0.f; v[u].z = n + 2 < N ? wp[2] : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kb + u * TT_WARPS;
        if (k < k1) {
          if (SMAX >= 4 && svec) {  // S' rows contiguous, 16B aligned and padded: SMAX/4 float4 loads per row
#pragma unroll
            for (int s4 = 0; s4 < SMAX / 4; ++s4) {
              if (s4 * 4 < Sn) {
                const float4 cv = __ldg(reinterpret_cast<const float4*>(S + (long)k * s_ks) + s4);
                const float c4[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int s = s4 * 4 + e;  // rows s >= Sn accumulate padding values that are never written out
                  acc[s].x = fmaf(c4[e], v[u].x, acc[s].x); acc[s].y = fmaf(c4[e], v[u].y, acc[s].y);
                  acc[s].z = fmaf(c4[e], v[u].z, acc[s].z); acc[s].w = fmaf(c4[e], v[u].w, acc[s].w);
                }
              }
            }
          } else {
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
              if (s < Sn) {
                const float c = S != nullptr ? __ldg(S + (long)k * s_ks + s * s_ss) : 1.f;
                acc[s].x = fmaf(c, v[u].x, acc[s].x); acc[s].y = fmaf(c, v[u].y, acc[s].y);
                acc[s].z = fmaf(c, v[u].z, acc[s].z); acc[s].w = fmaf(c, v[u].w, acc[s].w);
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < SMAX; ++s) s_red[(w * SMAX + s) * 32 + lane] = acc[s];
  __syncthreads();
  // warp w reduces rows s = w, w+8, ...
  for (int s = w; s < Sn; s += TT_WARPS) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ww = 0; ww < TT_WARPS; ++ww) {
      const float4 v = s_red[(ww * SMAX + s) * 32 + lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    if (n < N) *reinterpret_cast<float4*>(partial + ((long)blockIdx.y * Sn + s) * ldp + n) = t;
  }
}

// C(s,n) = alpha * sum_split partial[split][s][n] + beta*C + bias.  Block = 32 outputs x 8 split lanes: the
// splits of one output are summed by 8 threads in a fixed interleaved order, then across the 8 in shared memory.
__global__ void __launch_bounds__(256) gemm_tallt_reduce_kernel(const float* __restrict__ partial, int nsplit, int Sn, int N,
                                                                int ldp, float alpha, float beta, const float* __restrict__ bias,
                                                                int bias_on_s, float* __restrict__ C, long c_ss, long c_ns) {
  __shared__ float red[8][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool ok = i < Sn * N;
  const int s = ok ? i / N : 0, n = ok ? i % N : 0;
  float t = 0.f;
  if (ok)
    for (int z = threadIdx.y; z < nsplit; z += 8) t += partial[((long)z * Sn + s) * ldp + n];
  red[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
    float x = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) x += red[y][threadIdx.x];
    x *= alpha;
    if (bias != nullptr) x += bias[bias_on_s ? s : n];
    float* cp = C + s * c_ss + n * c_ns;
    if (beta != 0.f) x = fmaf(beta, *cp, x);
    *cp = x;
  }
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int launch_smalln(const SkArgs& g, cudaStream_t st) {
  // many rows and a B' that fits shared memory: stage it (persistent-style grid, every CTA stages B' once)
  const size_t bbytes = (size_t)g.N * g.K * sizeof(float);
  // (staging costs N*K*4 bytes per CTA: at a few hundred rows the staging IS the kernel -- 28.8 us at 768 x 16 x 772 on the cfg1
  // timeline -- so few-row products take the plain variant below with one row per warp)
  if (g.M >= 2048 && g.N > 4 && bbytes <= 96 * 1024) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(gemm_smalln_kernel<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(gemm_smalln_kernel<16, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(gemm_smalln_kernel<32, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    const int R = g.N <= 16 ? 2 : 1;
    int grid = ceil_div(ceil_div(g.M, R), 8);
    const int cap = 148 * (bbytes <= 32 * 1024 ? 4 : 2);
    if (grid > cap) grid = cap;
    if (g.N <= 8) gemm_smalln_kernel<8, 2, true><<<grid, 256, bbytes, st>>>(g);
    else if (g.N <= 16) gemm_smalln_kernel<16, 2, true><<<grid, 256, bbytes, st>>>(g);
    else gemm_smalln_kernel<32, 1, true><<<grid, 256, bbytes, st>>>(g);
    IMMTSF_CHECK_LAUNCH("gemm_smalln");
    return IMMTSF_OK;
  }
  const bool few = g.M < 148 * 16 * 4;  // few rows: one or two rows per warp so that every SM has warps to hide latency
  const int R = g.N <= 8 ? (few ? (g.M < 148 * 16 * 2 ? 1 : 2) : 4) : (g.N <= 16 && !few ? 2 : 1);
  const int warps = ceil_div(g.M, R);
  int grid = ceil_div(warps, 8);
  if (grid > 148 * 8) grid = 148 * 8;
  if (g.N <= 4 && R == 4) gemm_smalln_kernel<4, 4><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 4 && R == 2) gemm_smalln_kernel<4, 2><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 4) gemm_smalln_kernel<4, 1><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 8 && R == 4) gemm_smalln_kernel<8, 4><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 8 && R == 2) gemm_smalln_kernel<8, 2><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 8) gemm_smalln_kernel<8, 1><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 16 && R == 2) gemm_smalln_kernel<16, 2><<<grid, 256, 0, st>>>(g);
  else if (g.N <= 16) gemm_smalln_kernel<16, 1><<<grid, 256, 0, st>>>(g);
  else gemm_smalln_kernel<32, 1><<<grid, 256, 0, st>>>(g);
  IMMTSF_CHECK_LAUNCH("gemm_smalln");
  return IMMTSF_OK;
}

int launch_tallt(const float* S, long s_ks, long s_ss, int Sn, const float* W, long w_rs, int N, int K, const int32_t* ragged,
                 float alpha, float beta, const float* bias, int bias_on_s, float* C, long c_ss, long c_ns, void* workspace,
                 size_t workspace_bytes, cudaStream_t st) {
  const int ntiles = ceil_div(N, 128);
  int nsplit = ceil_div(2 * 148, ntiles);
  if (nsplit > ceil_div(K, 16)) nsplit = ceil_div(K, 16);
  const int ldp = (N + 3) / 4 * 4;
  const int cap = (int)(((size_t)8 << 20) / ((size_t)Sn * ldp * sizeof(float)));  // keep the partials L2-resident
  if (nsplit > cap) nsplit = cap;
  if (nsplit < 1) nsplit = 1;
  const size_t need = (size_t)nsplit * Sn * ldp * sizeof(float) + 256;
  if (workspace == nullptr || workspace_bytes < need) return 1;  // caller falls back
  float* partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  dim3 grid(ntiles, nsplit);
#define TALLT(SM)                                                                                                        \
  gemm_tallt_partial_kernel<SM><<<grid, 256, TT_WARPS * SM * 32 * sizeof(float4), st>>>(S, s_ks, s_ss, Sn, W, w_rs, N, K, \
                                                                                         ragged, partial, ldp)
  if (Sn <= 1) TALLT(1);
  else if (Sn <= 4) TALLT(4);
  else if (Sn <= 8) TALLT(8);
  else {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(gemm_tallt_partial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(TT_WARPS * 16 * 32 * sizeof(float4)));
      attr = true;
    }
    TALLT(16);
  }
#undef TALLT
  IMMTSF_CHECK_LAUNCH("gemm_tallt_partial");
  gemm_tallt_reduce_kernel<<<ceil_div(Sn * N, 32), dim3(32, 8), 0, st>>>(partial, nsplit, Sn, N, ldp, alpha, beta, bias, bias_on_s, C,
                                                                  c_ss, c_ns);
  IMMTSF_CHECK_LAUNCH("gemm_tallt_reduce");
  return 0;
}

}  // namespace

size_t immtsf_gemm_skinny_workspace(int M, int N, int K) {
  // tall-T partials are capped at 8 MiB (or one split of 16 x wide floats)
  const size_t wide = (size_t)(M > N ? M : N);
  const size_t one = 16 * ((wide + 3) / 4 * 4) * sizeof(float);
  (void)K;
  return (one > ((size_t)8 << 20) ? one : ((size_t)8 << 20)) + 256;
}

// Returns IMMTSF_OK when a skinny kernel was launched, 1 when the shape is not skinny (caller continues with
// the general backends), negative on error.
int immtsf_gemm_skinny(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                       int ldb, float beta, float* C, int ldc, const float* bias, const int32_t* ragged, int ragged_dim,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (K < 1) return 1;
  SkArgs g;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
  g.A = A; g.a_rs = transA ? 1 : lda; g.a_cs = transA ? lda : 1;
  g.B = B; g.b_rs = transB ? 1 : ldb; g.b_cs = transB ? ldb : 1;
  g.C = C; g.c_rs = ldc; g.c_cs = 1; g.bias = bias; g.bias_on_m = 0; g.ragged = ragged; g.ragged_dim = ragged_dim;

  // ---- small-K: rank-K update streamed over the rows of C
  if (K <= 16 && N >= 32 && ragged_dim != 2 && al16(C) && (ldc & 3) == 0) {
    const int threads = min(256, ((N + 3) / 4 + 31) / 32 * 32);
    dim3 grid(ceil_div((N + 3) / 4, threads), min(M, max(1, 148 * 8 / ceil_div((N + 3) / 4, threads))));
    if (K <= 4) gemm_smallk_kernel<4><<<grid, threads, 0, st>>>(g);
    else if (K <= 8) gemm_smallk_kernel<8><<<grid, threads, 0, st>>>(g);
    else gemm_smallk_kernel<16><<<grid, threads, 0, st>>>(g);
    IMMTSF_CHECK_LAUNCH("gemm_smallk");
    return IMMTSF_OK;
  }
  // ---- tall-T: contraction over rows, one side at most 32 wide (chunks of 16)
  const int32_t* kr = ragged_dim == 2 ? ragged : nullptr;
  if (transA && !transB && ragged_dim != 1 && (M <= 32 || N <= 32)) {
    if (M <= 32 && M <= N && al16(B) && (ldb & 3) == 0) {  // C[s=m, n]: S' = A [K][M], W = B [K][N]
      for (int s0 = 0; s0 < M; s0 += 16) {
        const int rc = launch_tallt(A + s0, lda, 1, min(16, M - s0), B, ldb, N, K, kr, alpha, beta, bias, 0,
                                    C + (long)s0 * ldc, ldc, 1, workspace, workspace_bytes, st);
        if (rc != 0) return rc;
      }
      return IMMTSF_OK;
    }
    if (N <= 32 && al16(A) && (lda & 3) == 0) {  // C[m=wide, n=s]: S' = B [K][N], W = A [K][M]
      for (int s0 = 0; s0 < N; s0 += 16) {
        const int rc = launch_tallt(B + s0, ldb, 1, min(16, N - s0), A, lda, M, K, kr, alpha, beta,
                                    bias ? bias + s0 : nullptr, 1, C + s0, 1, ldc, workspace, workspace_bytes, st);
        if (rc != 0) return rc;
      }
      return IMMTSF_OK;
    }
    return 1;
  }
  // C[M<=16, N] = A[M,K] B[K,N] with B n-contiguous: same kernel, S'(k,s) = A[s*lda + k]
  if (!transA && !transB && M <= 16 && N > 16 && ragged_dim == 0 && al16(B) && (ldb & 3) == 0)
    return launch_tallt(A, 1, lda, M, B, ldb, N, K, nullptr, alpha, beta, bias, 0, C, ldc, 1, workspace, workspace_bytes, st);

  // ---- small-N: rows of A streamed once
  if (!transA && N <= 32 && ragged_dim != 2 && (K & 3) == 0 && al16(A) && (lda & 3) == 0) return launch_smalln(g, st);
  // C[M<=16, N] with B stored [N][K]: compute C^T[N, M] = B A'^T with the small-N kernel
  if (transB && M <= 32 && N > 32 && ragged_dim == 0 && (K & 3) == 0 && al16(B) && (ldb & 3) == 0) {
    SkArgs t = g;
    t.M = N; t.N = M;
    t.A = B; t.a_rs = ldb; t.a_cs = 1;
    t.B = A; t.b_rs = g.a_cs; t.b_cs = g.a_rs;  // B2'(k, s) = A'(s, k)
    t.c_rs = 1; t.c_cs = ldc;                    // C2(n, s) = C[s*ldc + n]
    t.bias_on_m = 1;                             // bias is indexed by the original n = row of C2
    return launch_smalln(t, st);
  }
  return 1;
}

// ------------------------------------------------------------------ colsum
// out[n] = beta*out[n] + sum_m X[m,n]  ==  tall-T with an implicit ones vector.
__global__ void colsum_serial_kernel(const float* __restrict__ X, int M, int N, int ldx, float* __restrict__ out, float beta,
                                     const int32_t* __restrict__ ragged) {
  __shared__ float red[8][33];
  const int m_eff = ragged_rows(M, ragged);
  const int n = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < m_eff; m += 8) s += X[(size_t)m * ldx + n];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[n] = (beta != 0.f ? beta * out[n] : 0.f) + t;
  }
}

extern "C" int immtsf_colsum(const float* X, int M, int N, int ldx, float* out, float beta, const int32_t* ragged,
                             void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(X && out, "colsum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (M >= 64 && al16(X) && (ldx & 3) == 0) {
    const int rc = launch_tallt(nullptr, 0, 0, 1, X, ldx, N, M, ragged, 1.f, beta, nullptr, 0, out, 0, 1, workspace, workspace_bytes, st);
    if (rc <= 0) return rc;
  }
  // small or unaligned input, or no workspace: one CTA per 32 columns
  colsum_serial_kernel<<<ceil_div(N, 32), dim3(32, 8), 0, st>>>(X, M, N, ldx, out, beta, ragged);
  IMMTSF_CHECK_LAUNCH("colsum");
  return IMMTSF_OK;
}
