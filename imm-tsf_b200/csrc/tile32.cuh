// 32 x 32 register-tiled X Y^T over a long contraction, shared by the small-T attention core (xattn_small.cu)
// and the segment-attention backward (t2v_segattn.cu).
#pragma once
#include "rowtile.cuh"

constexpr int XS_T = 32;     // max rows of X / Y per call
constexpr int XS_DC = 256;   // contraction chunk staged in shared memory
constexpr int XS_LD = XS_DC + 4;
// shared memory the routine needs: s_x, s_y [XS_T][XS_LD] each, s_part [8][XS_T*XS_T], s_out [XS_T*XS_T]
constexpr int XS_TILE_FLOATS = 2 * XS_T * XS_LD + 8 * XS_T * XS_T;

// S[i][j] = sum_c X[i][c] * Y[j][c] over c in [0, hd): result (unscaled) in s_out[XS_T*XS_T] (row stride XS_T).
// X: nx <= 32 rows, Y: ny <= 32 rows (global, 16B-aligned rows).  Needs blockDim.x == 256.  s_x, s_y: [XS_T][XS_LD] staging; s_part: [8][XS_T*XS_T].
__device__ __forceinline__ void tile_xyt(const float* __restrict__ X, int ldx, const float* __restrict__ Y, int ldy, int nx, int ny, int hd,
                                         float* s_x, float* s_y, float* s_part, float* s_out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ib = lane >> 2, jb = lane & 3;  // rows ib*4 + a (a < 4), cols b*4 + jb (b < 8): conflict-free LDS.128
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  for (int c0 = 0; c0 < hd; c0 += XS_DC) {
    const int cw = min(XS_DC, hd - c0), cw4 = cw >> 2;
    __syncthreads();  // previous chunk consumed
#pragma unroll 4
    for (int i = threadIdx.x; i < XS_T * (XS_DC / 4); i += blockDim.x) {
      const int r = i / (XS_DC / 4), c4 = i % (XS_DC / 4);
      float4 xv = f4_zero(), yv = f4_zero();
      if (c4 < cw4) {
        if (r < nx) xv = __ldg(reinterpret_cast<const float4*>(X + (size_t)r * ldx + c0) + c4);
        if (r < ny) yv = __ldg(reinterpret_cast<const float4*>(Y + (size_t)r * ldy + c0) + c4);
      }
      *reinterpret_cast<float4*>(s_x + r * XS_LD + c4 * 4) = xv;
      *reinterpret_cast<float4*>(s_y + r * XS_LD + c4 * 4) = yv;
    }
    __syncthreads();
    const int sl0 = w * (XS_DC / 8);
#pragma unroll 4
    for (int c = sl0; c < sl0 + XS_DC / 8; c += 4) {
      float4 xa[4], yb[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = *reinterpret_cast<const float4*>(s_x + (ib * 4 + a) * XS_LD + c);
#pragma unroll
      for (int b = 0; b < 8; ++b) yb[b] = *reinterpret_cast<const float4*>(s_y + (b * 4 + jb) * XS_LD + c);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b)
          acc[a][b] = fmaf(xa[a].x, yb[b].x, fmaf(xa[a].y, yb[b].y, fmaf(xa[a].z, yb[b].z, fmaf(xa[a].w, yb[b].w, acc[a][b]))));
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) s_part[w * (XS_T * XS_T) + (ib * 4 + a) * XS_T + b * 4 + jb] = acc[a][b];
  __syncthreads();
  for (int i = threadIdx.x; i < XS_T * XS_T; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += s_part[ww * (XS_T * XS_T) + i];
    s_out[i] = s;
  }
  __syncthreads();
}

