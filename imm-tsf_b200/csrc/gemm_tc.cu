// tcgen05 / TMEM / TMA GEMM backend with fp32 accuracy (3xTF32).
//
//   C[M,N] = alpha * op(A) op(B) + beta * C + bias
//
// fp32 parity (1e-5) rules out a single TF32 pass (10-bit mantissa, ~1e-3).
// Every operand x is split in global memory into x_hi = x with the low 13
// mantissa bits cleared (exactly representable in TF32) and x_lo = x - x_hi
// (exact in fp32, |x_lo| <= 2^-11 |x|), and the tensor core accumulates
//   A_lo*B_hi + A_hi*B_lo + A_hi*B_hi      (small terms first)
// in fp32 in TMEM; the dropped A_lo*B_lo term is ~2^-22 relative.
//
// Kernel (one CTA per 128 x 128 output tile, 256 threads, warp-specialised):
//   warp 0   TMA producer: per 32-wide k-block four 128B-swizzled tiles
//            (A_hi, A_lo, B_hi, B_lo; 64 KiB) into a 3-stage smem ring,
//            completion on mbarriers (cp.async.bulk.tensor.2d).
//   warp 1   MMA issuer: one lane issues 12 tcgen05.mma.kind::tf32
//            (3 products x 4 k-steps of 8) per k-block, accumulator =
//            128 lanes x 128 fp32 columns of TMEM; tcgen05.commit frees the
//            smem stage and finally signals the epilogue.
//   warp 2   TMEM allocation / deallocation.
//   warps 4-7 epilogue: the tensor core's fp32 accumulator adds with truncation,
//            so a long K chain drifts (measured 6e-6 at K=768).  The K loop is
//            therefore cut in chunks of KC k-blocks (K=128): each chunk is
//            accumulated in one of two TMEM buffers, drained with tcgen05.ld
//            (lane == output row) and promoted into fp32 registers with
//            round-to-nearest adds while the next chunk's MMAs run.  The two
//            small cross products (lo*hi, hi*lo) go to their own TMEM
//            accumulator so that they do not add truncation steps to the
//            large hi*hi sum (16 instead of 48 truncating adds per chunk).  Then
//            alpha / bias / beta, ragged-row zeroing, 128-bit stores.
// Operands may be K-major or MN-major in memory (all four transposition
// cases): the UMMA shared-memory descriptor and instruction descriptor carry
// the major-ness.  K-major tiles use the 128B swizzle; MN-major fp32/tf32
// tiles must use the "128B swizzle with 32B atoms" layout (UMMA layout type 1,
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) and are fetched as 32x32 boxes.
//
// Roofline: tensor pipe.  12 MMAs of 128x128x8 per k-block = 768 tensor
// cycles for 2*128*128*32 useful FLOP -> 3xTF32 ceiling = 1/3 of the TF32
// peak (= 1/6 of the bf16 peak used as roofline.peak in bench.py).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BKT = 32;
constexpr int TILE_A = BM * BKT * 4;  // 16 KiB
constexpr int KC = 4;                 // k-blocks per TMEM accumulation chunk (K = 128)
constexpr int TMEM_COLS = 512;        // both variants use all of it

// Two tile shapes.  The kernel is bound by L2->SM operand bandwidth (~42 B/clk/SM chip-wide, ncu: 49 % tensor-pipe
// activity with 128x128 tiles = 64 KiB per k-block per CTA), so the wide tile trades TMEM for bytes per FLOP:
//   BN = 128: 3 stages x 64 KiB, two TMEM accumulators per buffer (hi*hi and the cross terms), 4 epilogue warps.
//   BN = 256: 2 stages x 96 KiB (0.75x the operand bytes per FLOP), ONE accumulator per buffer (512 TMEM columns
//             = 2 buffers x 256), 8 epilogue warps (each thread promotes 128 columns of its row in registers;
//             setmaxnreg moves registers from the 4 control warps to the 8 epilogue warps).
template <int BN_>
struct Cfg {
  static constexpr int BN = BN_;
  static constexpr int STAGES = BN_ == 128 ? 3 : 2;
  static constexpr int NACC = BN_ == 128 ? 2 : 1;
  static constexpr int TILE_B = BN_ * BKT * 4;
  static constexpr int STAGE_BYTES = 2 * TILE_A + 2 * TILE_B;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int EPI_WARPS = 4 * (BN_ / 128);
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 2-D tile of a plain matrix, or of batch (b1, b2) of a 4-D [batch1][batch2][rows][cols] view
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int batched, int b2,
                                         int b1) {
  if (batched) tma_load_4d(dst, map, bar, c0, c1, b2, b1);
  else tma_load_2d(dst, map, bar, c0, c1);
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, version 1 (Blackwell).
//   K-major  (layout 2, SWIZZLE_128B): rows of 128 B (32 fp32 of K); 8-row groups SBO = 1024 B apart; LBO unused.
//   MN-major (layout 1, SWIZZLE_128B_BASE32B -- the only one legal for 32-bit MN-major operands): 32x32 boxes of
//             4096 B; k-rows of 128 B (32 fp32 of M/N); 4-row k-groups SBO = 512 B apart; blocks of 32 along M/N
//             LBO = 4096 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(mn_major ? (4096u >> 4) : 1u) << 16;
  d |= (uint64_t)(mn_major ? (512u >> 4) : (1024u >> 4)) << 32;
  d |= (uint64_t)1 << 46;                     // version
  d |= (uint64_t)(mn_major ? 1u : 2u) << 61;  // layout type
  return d;
}

struct TcArgs {
  float* C;
  int ldc;
  int M, N, K;
  float alpha, beta;
  const float* bias;
  const int32_t* ragged;
  int ragged_dim;
  float* partial;  // split-K partial tiles [splitk][M][ldp]
  int ldp;
  // batched mode (attention contractions): blockIdx.z = b1 * batch2 + b2, operands are 4-D tensor maps,
  // C(b1, b2) = C + b1 * c_s1 + b2 * c_s2; no split-K, no ragged bounds
  int batched, batch2;
  long c_s1, c_s2;
  // optional second output: lo = C - trunc_tf32(C), so that a consumer product does not need a split pass over C
  float* C_lo;
  int ldc_lo;
  // diagnostics (immtsf_gemm_trace): per-CTA clock64 stamps of the pair kernel's phases, 8 slots per CTA
  long long* trace;
};
#define TC_STAMP(slot)                                                                                             \
  do {                                                                                                             \
    if (g.trace != nullptr)                                                                                        \
      g.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (slot)] = clock64(); \
  } while (0)


// Epilogue write-out.  After tcgen05.ld a thread owns one output ROW (128 consecutive columns of it), so direct stores
// would touch 32 different 128 B lines per instruction (measured: 16 k cycles to store a 128 x 256 tile, a quarter of
// the CTA's lifetime at K = 768).  Each epilogue warp therefore stages its 32 rows x 128 columns in shared memory
// (the operand ring is idle once the last accumulator chunk is complete; row stride 132 floats keeps both the
// row-per-lane writes and the row-per-instruction reads conflict-free) and then writes whole 512 B row segments.
constexpr int EPI_LD = 132;
constexpr int EPI_WARP_BYTES = 32 * EPI_LD * 4;  // 16.5 KiB per epilogue warp

// rows [row0, row0+32) x columns [ncol0, ncol0+128) of this warp; `out` points at (row0, ncol0) of C or of the
// split-K partial tile.  partial: raw sums, rows < Mfull, columns < ncols (ld of the partial).  Otherwise
// alpha / bias / beta, zeros for ragged pad rows (row >= Mlive), columns < ncols.
__device__ __forceinline__ void epilogue_store(const float (&acc)[128], uint32_t stage, int lane, float* out, size_t ld,
                                               int row0, int ncol0, int Mfull, int Mlive, int ncols, bool partial, float alpha,
                                               float beta, const float* __restrict__ bias, float* out_lo = nullptr,
                                               size_t ld_lo = 0) {
  const uint32_t mine = stage + (uint32_t)(lane * EPI_LD * 4);
#pragma unroll
  for (int j4 = 0; j4 < 32; ++j4)
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mine + j4 * 16), "f"(acc[j4 * 4]), "f"(acc[j4 * 4 + 1]),
                 "f"(acc[j4 * 4 + 2]), "f"(acc[j4 * 4 + 3]) : "memory");
  __syncwarp();
  const int n = ncol0 + lane * 4;
  if (n >= ncols) return;
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (!partial && bias != nullptr) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n + e < ncols) b[e] = __ldg(bias + n + e);
  }
  const bool full4 = n + 3 < ncols;
  const int nrows = min(32, Mfull - row0);
  const bool use_beta = !partial && beta != 0.f;
  for (int r8 = 0; r8 < nrows; r8 += 8) {
    // beta * C: the eight rows' old values are requested together, ahead of their use (one dependent global load per row
    // made the accumulate-into-C epilogue 18 us longer than the plain one at 4096 x 768 x 768)
    float4 cc[8];
    if (use_beta) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        cc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int rr = r8 + u;
        if (rr < nrows && row0 + rr < Mlive) {
          const float* p = out + (size_t)rr * ld + lane * 4;
          if (full4) {
            cc[u] = *reinterpret_cast<const float4*>(p);
          } else {
            cc[u].x = p[0];
            if (n + 1 < ncols) cc[u].y = p[1];
            if (n + 2 < ncols) cc[u].z = p[2];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int rr = r8 + u;
      if (rr >= nrows) break;
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(stage + (uint32_t)((rr * EPI_LD + lane * 4) * 4)));
      float* p = out + (size_t)rr * ld + lane * 4;
      if (!partial) {
        const bool live = row0 + rr < Mlive;
        v.x = live ? fmaf(alpha, v.x, b[0]) : 0.f; v.y = live ? fmaf(alpha, v.y, b[1]) : 0.f;
        v.z = live ? fmaf(alpha, v.z, b[2]) : 0.f; v.w = live ? fmaf(alpha, v.w, b[3]) : 0.f;
        if (use_beta && live) {
          v.x = fmaf(beta, cc[u].x, v.x); v.y = fmaf(beta, cc[u].y, v.y); v.z = fmaf(beta, cc[u].z, v.z); v.w = fmaf(beta, cc[u].w, v.w);
        }
      }
      if (full4) {
        *reinterpret_cast<float4*>(p) = v;
      } else {
        p[0] = v.x;
        if (n + 1 < ncols) p[1] = v.y;
        if (n + 2 < ncols) p[2] = v.z;
      }
      if (out_lo != nullptr) {  // ld_lo is a multiple of 4 and >= roundup(ncols, 4): the whole float4 is in bounds
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = n + 1 < ncols ? v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u) : 0.f;
        l.z = n + 2 < ncols ? v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u) : 0.f;
        l.w = n + 3 < ncols ? v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u) : 0.f;
        *reinterpret_cast<float4*>(out_lo + (size_t)rr * ld_lo + lane * 4) = l;
      }
    }
  }
}

// One CTA = one BM x BN output tile (tile_m, tile_n).  A_MN / B_MN are compile-time constants in gemm_tc_kernel and
// run-time values in the grouped kernel (they only select descriptor bits and TMA box coordinates).
template <int BN>
__device__ __forceinline__ void tc_cta(const CUtensorMap* pAh, const CUtensorMap* pAl, const CUtensorMap* pBh,
                                       const CUtensorMap* pBl, const TcArgs& g, const bool A_MN, const bool B_MN,
                                       const int tile_m, const int tile_n, const int nsplit, const int zsplit, const int bz1,
                                       const int bz2) {
  using C_ = Cfg<BN>;
  constexpr int STAGES = C_::STAGES, STAGE_BYTES = C_::STAGE_BYTES, TILE_B = C_::TILE_B, NACC = C_::NACC;
  extern __shared__ uint8_t smem_raw[];
  int M = g.M, K = g.K;
  if (g.ragged_dim == 1) M = ragged_rows(M, g.ragged);
  if (g.ragged_dim == 2) K = ragged_rows(K, g.ragged);
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  if (g.ragged_dim == 1 && m0 >= M) return;  // whole tile beyond the ragged end (uniform per CTA)
  const int nkb = (K + BKT - 1) / BKT;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], tfull[2], tempty[2], tmem_ptr
  const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  const int kb_per = (nkb + nsplit - 1) / nsplit;
  const int z_kb0 = min(nkb, zsplit * kb_per), z_kb1 = min(nkb, z_kb0 + kb_per);
  const int nchunks = (z_kb1 - z_kb0 + KC - 1) / KC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {  // hide the descriptor fetch behind the barrier / TMEM set-up
    prefetch_tmap(pAh); prefetch_tmap(pAl); prefetch_tmap(pBh); prefetch_tmap(pBl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, C_::EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
  // 384-thread variant: the control warpgroup hands its registers to the two epilogue warpgroups, whose fp32
  // promotion tile (128 columns per thread) does not fit the 168-register launch allocation
  if (BN == 256) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = z_kb0; kb < z_kb1; ++kb) {
      const int it = kb - z_kb0;
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t sa_h = base + s * STAGE_BYTES, sa_l = sa_h + TILE_A, sb_h = sa_l + TILE_A, sb_l = sb_h + TILE_B;
      const uint32_t fb = bar_full + 8 * s;
      mbar_expect_tx(fb, STAGE_BYTES);
      const int k0 = kb * BKT;
      if (!A_MN) {
        tma_tile(sa_h, pAh, fb, k0, m0, g.batched, bz2, bz1);
        tma_tile(sa_l, pAl, fb, k0, m0, g.batched, bz2, bz1);
      } else {
#pragma unroll
        for (int j = 0; j < BM / 32; ++j) {
          tma_tile(sa_h + j * 4096, pAh, fb, m0 + 32 * j, k0, g.batched, bz2, bz1);
          tma_tile(sa_l + j * 4096, pAl, fb, m0 + 32 * j, k0, g.batched, bz2, bz1);
        }
      }
      if (!B_MN) {
        tma_tile(sb_h, pBh, fb, k0, n0, g.batched, bz2, bz1);
        tma_tile(sb_l, pBl, fb, k0, n0, g.batched, bz2, bz1);
      } else {
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          tma_tile(sb_h + j * 4096, pBh, fb, n0 + 32 * j, k0, g.batched, bz2, bz1);
          tma_tile(sb_l + j * 4096, pBl, fb, n0 + 32 * j, k0, g.batched, bz2, bz1);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bar_tempty + 8 * buf, ((c >> 1) & 1) ^ 1);  // epilogue has drained this TMEM buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc_big = tmem_base + (uint32_t)(buf * NACC * BN), tacc_small = tacc_big + (NACC == 2 ? BN : 0);
      const int kb0 = z_kb0 + c * KC;
      const int kb_end = min(z_kb1, kb0 + KC);
      for (int kb = kb0; kb < kb_end; ++kb) {
        const int it = kb - z_kb0;
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa_h = base + s * STAGE_BYTES, sa_l = sa_h + TILE_A, sb_h = sa_l + TILE_A, sb_l = sb_h + TILE_B;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t sa = prod == 0 ? sa_l : sa_h;  // lo*hi, hi*lo, hi*hi
          const uint32_t sb = prod == 1 ? sb_l : sb_h;
#pragma unroll
          for (int ks = 0; ks < BKT / 8; ++ks) {
            const uint64_t ad = make_desc(sa + (A_MN ? ks * 1024 : ks * 32), A_MN);
            const uint64_t bd = make_desc(sb + (B_MN ? ks * 1024 : ks * 32), B_MN);
            const bool first = (kb == kb0) && ks == 0 && (prod == 0 || (NACC == 2 && prod == 2));
            umma_tf32(prod == 2 ? tacc_big : tacc_small, ad, bd, idesc, first ? 0u : 1u);
          }
        }
        umma_commit(bar_empty + 8 * s);  // arrives when the MMAs that read this stage have completed
      }
      umma_commit(bar_tfull + 8 * buf);
    }
  }
  } else {
    // ===== epilogue =====
    if (BN == 256) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int ch = ((warp - 4) >> 2) * 128;  // this warp's 128-column half of the tile (BN = 256: two warpgroups)
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bar_tfull + 8 * buf, (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NACC * BN + ch + c0), v);
        if (NACC == 2) {
          uint32_t u[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NACC * BN + BN + ch + c0), u);
#pragma unroll
          for (int j = 0; j < 32; ++j)  // fp32 round-to-nearest promotion
            acc[c0 + j] += __uint_as_float(v[j]) + __uint_as_float(u[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
    // the operand ring is idle now (every load was consumed by an MMA that has completed): stage + coalesced stores
    const uint32_t stage = base + (uint32_t)((warp - 4) * EPI_WARP_BYTES);
    const int row0 = m0 + q * 32;
    if (nsplit > 1) {
      // split-K: this split's partial tile goes to the workspace [z][M][ldp]; splitk_reduce_kernel sums the
      // splits in a fixed order (deterministic, unlike atomics) and applies alpha / beta / bias
      epilogue_store(acc, stage, lane, g.partial + ((size_t)zsplit * g.M + row0) * g.ldp + n0 + ch, (size_t)g.ldp, row0, n0 + ch,
                     g.M, g.M, g.ldp, true, 1.f, 0.f, nullptr);
    } else {
      epilogue_store(acc, stage, lane, g.C + bz1 * g.c_s1 + bz2 * g.c_s2 + (size_t)row0 * g.ldc + n0 + ch, (size_t)g.ldc, row0,
                     n0 + ch, g.M, M, g.N, false, g.alpha, g.beta, g.bias,
                     g.C_lo != nullptr ? g.C_lo + (size_t)row0 * g.ldc_lo + n0 + ch : nullptr, (size_t)g.ldc_lo);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}


template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(Cfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcArgs g) {
  // split-K: blockIdx.z owns a contiguous range of k-blocks (partial tiles are summed by splitk_reduce_kernel);
  // batched: blockIdx.z is the batch index and every CTA runs the whole contraction
  const int nsplit = g.batched ? 1 : (int)gridDim.z, zsplit = g.batched ? 0 : (int)blockIdx.z;
  const int bz1 = g.batched ? (int)blockIdx.z / g.batch2 : 0, bz2 = g.batched ? (int)blockIdx.z % g.batch2 : 0;
  tc_cta<BN>(&mapAh, &mapAl, &mapBh, &mapBl, g, A_MN, B_MN, (int)blockIdx.y, (int)blockIdx.x, nsplit, zsplit, bz1, bz2);
}

// Grouped launch: up to GRP_MAX independent small products (the d x d x d weight-space folds and un-folds, which are
// too small to fill the GPU one at a time) as ONE grid; blockIdx.z = problem, 128 x 128 tiles, no split-K.
constexpr int GRP_MAX = 4;
struct GroupArgs {
  CUtensorMap maps[GRP_MAX][4];  // A, A_lo, B, B_lo
  TcArgs g[GRP_MAX];
  int a_mn[GRP_MAX], b_mn[GRP_MAX];
};
__global__ void __launch_bounds__(Cfg<128>::THREADS, 1) gemm_tc_group_kernel(const __grid_constant__ GroupArgs grp) {
  const int p = blockIdx.z;
  const TcArgs& g = grp.g[p];
  if ((int)blockIdx.y * BM >= g.M || (int)blockIdx.x * 128 >= g.N) return;  // this problem has fewer tiles (uniform per CTA)
  tc_cta<128>(&grp.maps[p][0], &grp.maps[p][1], &grp.maps[p][2], &grp.maps[p][3], g, grp.a_mn[p] != 0, grp.b_mn[p] != 0,
              (int)blockIdx.y, (int)blockIdx.x, 1, 0, 0, 0);
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (one TPC) compute one 256 x 256 output tile: each CTA stages ITS 128 rows of A and ITS 128
// columns of B (64 KiB per k-block instead of 96 KiB for the same 128 x 256 per-CTA output -> 3 stages fit), the
// leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 with M = 256, N = 256, and the tensor cores of both SMs
// read both B halves.  Each CTA's TMEM holds its own 128 rows x 256 columns, so the epilogue is the 128 x 256 one.
//   full[s]   lives in the leader; both producers' TMA loads complete_tx on it (expect_tx = 2 x 64 KiB)
//   empty[s]  one per CTA, arrived by the leader's tcgen05.commit multicast to both CTAs
//   tfull[b]  one per CTA (multicast commit), tempty[b] in the leader, 16 arrivals (8 epilogue warps x 2 CTAs)
struct Cfg2 {
  static constexpr int BN = 256, BNH = 128;
  static constexpr int STAGES = 3;
  static constexpr int TILE_B = BNH * BKT * 4;
  static constexpr int STAGE_BYTES = 2 * TILE_A + 2 * TILE_B;  // 64 KiB per CTA
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data into the executing CTA's smem, completion bytes onto `bar` (a shared::cluster address,
// here always the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg2::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl, const TcArgs g) {
  using C_ = Cfg2;
  constexpr int STAGES = C_::STAGES, STAGE_BYTES = C_::STAGE_BYTES, TILE_B = C_::TILE_B, BN = C_::BN, BNH = C_::BNH;
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) {
    TC_STAMP(0);
    if (g.trace != nullptr) {
      long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      g.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + 7] = gt;
    }
  }
  int M = g.M, K = g.K;
  if (g.ragged_dim == 1) M = ragged_rows(M, g.ragged);
  if (g.ragged_dim == 2) K = ragged_rows(K, g.ragged);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m_pair = (int)(blockIdx.x >> 1) * (2 * BM);
  const int m0 = m_pair + (int)rank * BM, n0 = blockIdx.y * BN;
  if (g.ragged_dim == 1 && m_pair >= M) return;  // whole pair tile beyond the ragged end (uniform per cluster)
  const int nkb = (K + BKT - 1) / BKT;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  const int nsplit = (int)gridDim.z, zsplit = (int)blockIdx.z;
  const int kb_per = (nkb + nsplit - 1) / nsplit;
  const int z_kb0 = min(nkb, zsplit * kb_per), z_kb1 = min(nkb, z_kb0 + kb_per);
  const int nchunks = (z_kb1 - z_kb0 + KC - 1) / KC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {  // hide the descriptor fetch behind the barrier / TMEM set-up
    prefetch_tmap(&mapAh); prefetch_tmap(&mapAl); prefetch_tmap(&mapBh); prefetch_tmap(&mapBl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 2 * C_::EPI_WARPS);  // epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // barrier inits and the TMEM allocation of BOTH CTAs are visible before any remote arrive / MMA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs; completion on the leader's full barrier) =====
    const uint32_t nb0 = n0 + (int)rank * BNH;
    for (int kb = z_kb0; kb < z_kb1; ++kb) {
      const int it = kb - z_kb0;
      const int s = it % STAGES, ph = (it / STAGES) & 1;
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      const uint32_t sa_h = base + s * STAGE_BYTES, sa_l = sa_h + TILE_A, sb_h = sa_l + TILE_A, sb_l = sb_h + TILE_B;
      if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);
      const uint32_t fb = mapa_shared(bar_full + 8 * s, 0);
      const int k0 = kb * BKT;
      if (!A_MN) {
        tma_load_2d_pair(sa_h, &mapAh, fb, k0, m0);
        tma_load_2d_pair(sa_l, &mapAl, fb, k0, m0);
      } else {
#pragma unroll
        for (int j = 0; j < BM / 32; ++j) {
          tma_load_2d_pair(sa_h + j * 4096, &mapAh, fb, m0 + 32 * j, k0);
          tma_load_2d_pair(sa_l + j * 4096, &mapAl, fb, m0 + 32 * j, k0);
        }
      }
      if (!B_MN) {
        tma_load_2d_pair(sb_h, &mapBh, fb, k0, nb0);
        tma_load_2d_pair(sb_l, &mapBl, fb, k0, nb0);
      } else {
#pragma unroll
        for (int j = 0; j < BNH / 32; ++j) {
          tma_load_2d_pair(sb_h + j * 4096, &mapBh, fb, nb0 + 32 * j, k0);
          tma_load_2d_pair(sb_l + j * 4096, &mapBl, fb, nb0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===== MMA issuer (leader CTA only) =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bar_tempty + 8 * buf, ((c >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this TMEM buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
      const int kb0 = z_kb0 + c * KC;
      const int kb_end = min(z_kb1, kb0 + KC);
      for (int kb = kb0; kb < kb_end; ++kb) {
        const int it = kb - z_kb0;
        const int s = it % STAGES, ph = (it / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (kb == z_kb0) TC_STAMP(2);
        const uint32_t sa_h = base + s * STAGE_BYTES, sa_l = sa_h + TILE_A, sb_h = sa_l + TILE_A, sb_l = sb_h + TILE_B;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t sa = prod == 0 ? sa_l : sa_h;  // lo*hi, hi*lo, hi*hi
          const uint32_t sb = prod == 1 ? sb_l : sb_h;
#pragma unroll
          for (int ks = 0; ks < BKT / 8; ++ks) {
            const uint64_t ad = make_desc(sa + (A_MN ? ks * 1024 : ks * 32), A_MN);
            const uint64_t bd = make_desc(sb + (B_MN ? ks * 1024 : ks * 32), B_MN);
            const bool first = (kb == kb0) && ks == 0 && prod == 0;
            umma_tf32_pair(tacc, ad, bd, idesc, first ? 0u : 1u);
          }
        }
        umma_commit_pair(bar_empty + 8 * s);  // frees this stage in BOTH CTAs
      }
      umma_commit_pair(bar_tfull + 8 * buf);
    }
    TC_STAMP(3);
  }
  } else {
    // ===== epilogue (both CTAs; each drains its own 128 rows) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;
    const int ch = ((warp - 4) >> 2) * 128;
    const uint32_t tempty_leader = mapa_shared(bar_tempty, 0);
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int buf = c & 1;
      mbar_wait(bar_tfull + 8 * buf, (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (c == nchunks - 1 && threadIdx.x == 128) TC_STAMP(4);
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + ch + c0), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * buf);
    }
    const uint32_t stage = base + (uint32_t)((warp - 4) * EPI_WARP_BYTES);
    const int row0 = m0 + q * 32;
    if (nsplit > 1) {
      epilogue_store(acc, stage, lane, g.partial + ((size_t)zsplit * g.M + row0) * g.ldp + n0 + ch, (size_t)g.ldp, row0, n0 + ch,
                     g.M, g.M, g.ldp, true, 1.f, 0.f, nullptr);
    } else {
      epilogue_store(acc, stage, lane, g.C + (size_t)row0 * g.ldc + n0 + ch, (size_t)g.ldc, row0, n0 + ch, g.M, M, g.N, false,
                     g.alpha, g.beta, g.bias, g.C_lo != nullptr ? g.C_lo + (size_t)row0 * g.ldc_lo + n0 + ch : nullptr,
                     (size_t)g.ldc_lo);
    }
  }
  if (threadIdx.x == 128) TC_STAMP(5);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // the peer may still be reading this CTA's smem / arriving on its barriers until here
  if (threadIdx.x == 0) TC_STAMP(6);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// C = alpha * sum_z partial[z] + beta*C + bias for rows < M_eff; zeros for ragged pad rows inside touched tiles
__global__ void splitk_reduce_kernel(float* __restrict__ C, int ldc, int M, int N, float alpha, float beta,
                                     const float* __restrict__ bias, const float* __restrict__ partial, int ldp, int splitk,
                                     const int32_t* __restrict__ ragged, int ragged_dim, float* __restrict__ C_lo, int ldc_lo) {
  int Meff = M;
  if (ragged_dim == 1) Meff = ragged_rows(M, ragged);
  int Mtouch = (Meff + BM - 1) / BM * BM;
  if (Mtouch > M) Mtouch = M;
  const int n4 = ldp >> 2;
  const size_t total = (size_t)Mtouch * n4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n4), c = (int)(i % n4) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < Meff)
      for (int z0 = 0; z0 < splitk; z0 += 8) {  // eight partial tiles requested together, added in split order
        float4 p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          p[u] = z0 + u < splitk ? __ldcs(reinterpret_cast<const float4*>(partial + ((size_t)(z0 + u) * M + r) * ldp + c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) { s.x += p[u].x; s.y += p[u].y; s.z += p[u].z; s.w += p[u].w; }
      }
    const float v[4] = {s.x, s.y, s.z, s.w};
    float* out = C + (size_t)r * ldc + c;
    if (c + 3 < N && beta == 0.f && (((uintptr_t)out) & 15) == 0 && (C_lo == nullptr || (((uintptr_t)(C_lo + (size_t)r * ldc_lo + c)) & 15) == 0)) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < Meff) {
        x = make_float4(alpha * v[0], alpha * v[1], alpha * v[2], alpha * v[3]);
        if (bias != nullptr) { x.x += bias[c]; x.y += bias[c + 1]; x.z += bias[c + 2]; x.w += bias[c + 3]; }
      }
      *reinterpret_cast<float4*>(out) = x;
      if (C_lo != nullptr)
        *reinterpret_cast<float4*>(C_lo + (size_t)r * ldc_lo + c) =
            make_float4(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u),
                        x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
      continue;
    }
    for (int e = 0; e < 4 && c + e < N; ++e) {
      float x = 0.f;
      if (r < Meff) {
        x = alpha * v[e];
        if (beta != 0.f) x = fmaf(beta, out[e], x);
        if (bias != nullptr) x += bias[c + e];
      }
      out[e] = x;
      if (C_lo != nullptr) C_lo[(size_t)r * ldc_lo + c + e] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    }
  }
}

// ---------------------------------------------------------------- operand split
// tcgen05.mma kind::tf32 reads fp32 containers and ignores the low 13 mantissa bits (measured: feeding the raw
// operand or its pre-truncated copy gives bit-identical results, round-1 experiment), so the "hi" operand IS the
// original tensor and only lo = x - trunc_tf32(x) has to be materialised.  rows x cols (ld_src) -> [rows][ld_dst].
__global__ void split_lo_kernel(const float* __restrict__ src, int ld_src, int rows, int cols, float* __restrict__ lo,
                                int ld_dst, const int32_t* __restrict__ ragged, int ragged_rows_flag) {
  int live_rows = rows;
  if (ragged_rows_flag) {
    const int m = ragged_rows(rows, ragged);
    live_rows = (m + 127) / 128 * 128;  // pad rows are zeros written by the producers
    if (live_rows > rows) live_rows = rows;
  }
  const int c4n = ld_dst >> 2;
  const size_t total = (size_t)live_rows * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4n), c = (int)(i % c4n) * 4;
    const float* p = src + (size_t)r * ld_src + c;
    float4 x;
    if (c + 3 < cols) x = __ldg(reinterpret_cast<const float4*>(p));
    else {
      x.x = c + 0 < cols ? p[0] : 0.f; x.y = c + 1 < cols ? p[1] : 0.f;
      x.z = c + 2 < cols ? p[2] : 0.f; x.w = c + 3 < cols ? p[3] : 0.f;
    }
    float4 l;
    l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
    l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
    l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
    l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
    *reinterpret_cast<float4*>(lo + (size_t)r * ld_dst + c) = l;
  }
}

// Several small tensors (the weight matrices of a module) in ONE launch: optional plain copy (hi_dst: the packed
// operand of a fused projection) and / or the lo split.  blockIdx.y = task.
constexpr int MS_MAX = 16;
struct MultiSplitArgs {
  const float* src[MS_MAX];
  float* hi[MS_MAX];
  float* lo[MS_MAX];
  int rows[MS_MAX], cols[MS_MAX], ld_src[MS_MAX], ld_hi[MS_MAX], ld_lo[MS_MAX];
};
__global__ void multi_split_kernel(const __grid_constant__ MultiSplitArgs a) {
  const int t = blockIdx.y;
  const float* __restrict__ src = a.src[t];
  float* __restrict__ hi = a.hi[t];
  float* __restrict__ lo = a.lo[t];
  const int cols = a.cols[t], c4n = (cols + 3) >> 2;
  const size_t total = (size_t)a.rows[t] * c4n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4n), c = (int)(i % c4n) * 4;
    const float* p = src + (size_t)r * a.ld_src[t] + c;
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = c + e < cols ? __ldg(p + e) : 0.f;
    if (hi != nullptr) {
      float* h = hi + (size_t)r * a.ld_hi[t] + c;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < cols) h[e] = x[e];
    }
    if (lo != nullptr) {
      float* l = lo + (size_t)r * a.ld_lo[t] + c;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c + e < cols) l[e] = x[e] - __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
    }
  }
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

// 2-D fp32 row-major [rows][cols] (ld), box = {32 cols, box_rows}, 128B swizzle
int make_map(CUtensorMap* m, const float* ptr, int rows, int cols, int ld, int box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// 4-D fp32 view [batch1][batch2][rows][cols] with element strides (s1, s2, ld, 1); box = {32 cols, box_rows, 1, 1}
int make_map4(CUtensorMap* m, const float* ptr, int rows, int cols, long ld, long s2, long s1, int batch2, int batch1, int box_rows,
              bool mn_major) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch2, (cuuint64_t)batch1};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)s2 * 4, (cuuint64_t)s1 * 4};
  cuuint32_t box[4] = {32u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Kernel variants: 0 = 128x128 tiles, 1 = 128x256 tiles, 2 = CTA pairs (256x256 per cluster of two).
constexpr int V128 = 0, V256 = 1, VPAIR = 2;
inline int variant_bn(int v) { return v == V128 ? 128 : 256; }
// CTAs that cover the output once
inline long tile_ctas(int M, int N, int v) {
  if (v == VPAIR) return (long)ceil_div(N, 256) * 2 * ceil_div(M, 2 * BM);
  return (long)ceil_div(N, variant_bn(v)) * ceil_div(M, BM);
}

// split-K when the output has too few tiles to occupy the 148 SMs (weight gradients: M, N = d; K = rows)
inline int choose_splitk(int M, int N, int K, int v) {
  const long tiles = tile_ctas(M, N, v);
  const int nkb = ceil_div(K, BKT);
  int splitk = 1;
  if (tiles * 2 <= 148 && nkb >= 16) {
    splitk = (int)(148 / tiles);
    if (splitk > nkb / 8) splitk = nkb / 8;
    if (splitk > 16) splitk = 16;
    if (splitk < 1) splitk = 1;
  }
  return splitk;
}

// Expected fraction of live rows of a ragged operand (the true count lives on the device and is never read back):
// Time-IMM-like batches hold U{1..N_max} notes per sample.  IMMTSF_RAGGED_FILL overrides it.
inline double ragged_fill() {
  static double f = -1.0;
  if (f < 0.0) {
    const char* e = getenv("IMMTSF_RAGGED_FILL");
    f = e ? atof(e) : 0.6;
    if (!(f > 0.0 && f <= 1.0)) f = 0.6;
  }
  return f;
}

// Variant: estimated clocks = waves x (k-blocks per CTA x clocks per k-block + fixed prologue/epilogue), where a
// k-block costs the larger of its tensor time (768 clk for a 128x128 tile, 1536 for 128x256) and the time the L2
// needs to feed every co-resident CTA (64 / 96 / 64 KiB per k-block; ~6300 B/clk chip-wide).  Measured with
// immtsf_gemm_trace: ~15 k clocks of prologue + epilogue per CTA.  IMMTSF_TC_BN=128|256|512 forces one variant
// (512 = pairs; tests, experiments).  ragged_dim: the bound is device-resident, so the estimate uses ragged_fill().
inline int choose_variant(int M, int N, int K, int ragged_dim = 0) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("IMMTSF_TC_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 128) return V128;
  if (forced == 256) return V256;
  if (forced == 512) return VPAIR;
  if (N <= 128) return V128;
  const int Me = ragged_dim == 1 ? max(BM, (int)(M * ragged_fill())) : M;
  const int Ke = ragged_dim == 2 ? max(BKT, (int)(K * ragged_fill())) : K;
  double best = 0.0;
  int best_v = V128;
  for (int v = 0; v < 3; ++v) {
    if (v == VPAIR && M <= BM) continue;
    const int sk = choose_splitk(M, N, K, v);  // (the launch uses the allocated extent)
    const long ctas = tile_ctas(Me, N, v) * sk;
    const double waves = (double)((ctas + 147) / 148);
    const double resident = (double)(ctas < 148 ? ctas : 148);
    const double tens = v == V128 ? 768.0 : 1536.0, kib = v == V256 ? 96.0 : 64.0;
    const double feed = resident * kib * 1024.0 / 6300.0;
    const double t_kb = tens > feed ? tens : feed;
    const double cost = waves * ((double)ceil_div(ceil_div(Ke, BKT), sk) * t_kb + (v == VPAIR ? 16000.0 : 15000.0)) +
                        (sk > 1 ? 9000.0 : 0.0);  // + the partial-tile round trip and reduce launch of split-K
    if (v == 0 || cost < best) { best = cost; best_v = v; }
  }
  return best_v;
}

}  // namespace

// ---------------------------------------------------------------- per-launch timing (bench.py's roofline)
namespace {
long long* g_trace = nullptr;
struct ProfRec { cudaEvent_t e0, e1; int M, N, K, ragged_dim; };
ProfRec* g_prof = nullptr;
int g_prof_cap = 0, g_prof_n = 0, g_prof_on = 0;
}  // namespace

// Diagnostics: when set, every CTA of the CTA-pair kernel writes 8 clock64 stamps (entry, prologue done, first operands
// landed, MMAs issued, accumulator complete, stores done, exit, -) to buf[cta * 8 + slot]; NULL switches it off.
extern "C" int immtsf_gemm_trace(long long* buf) {
  g_trace = buf;
  return IMMTSF_OK;
}

// Start recording a CUDA-event pair around every gemm_tc_kernel launch (on the launch stream).
extern "C" int immtsf_profile_begin(int max_records) {
  IMMTSF_REQUIRE(max_records > 0, "profile_begin: max_records must be > 0");
  if (g_prof_cap < max_records) {
    ProfRec* p = (ProfRec*)realloc(g_prof, sizeof(ProfRec) * (size_t)max_records);
    IMMTSF_REQUIRE(p != nullptr, "profile_begin: out of memory");
    g_prof = p;
    for (int i = g_prof_cap; i < max_records; ++i) {
      cudaEventCreate(&g_prof[i].e0);
      cudaEventCreate(&g_prof[i].e1);
    }
    g_prof_cap = max_records;
  }
  g_prof_n = 0;
  g_prof_on = 1;
  return IMMTSF_OK;
}
// Stop recording; waits for the recorded launches and returns their shapes and durations (ms). Returns the count.
extern "C" int immtsf_profile_end(int* M, int* N, int* K, int* ragged_dim, float* ms, int cap) {
  g_prof_on = 0;
  int n = g_prof_n < cap ? g_prof_n : cap;
  for (int i = 0; i < n; ++i) {
    cudaEventSynchronize(g_prof[i].e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1);
    M[i] = g_prof[i].M; N[i] = g_prof[i].N; K[i] = g_prof[i].K; ragged_dim[i] = g_prof[i].ragged_dim; ms[i] = t;
  }
  g_prof_n = 0;
  return n;
}

// workspace: A_lo | B_lo (each dense with ld rounded up to 4 floats, 256 B aligned) | split-K partial tiles
size_t immtsf_gemm_tc_workspace(int transA, int transB, int M, int N, int K) {
  const size_t ra = transA ? K : M, ca = transA ? M : K, rb = transB ? N : K, cb = transB ? K : N;
  const size_t a = align_up(ra * align_up(ca, 4) * 4, 256), b = align_up(rb * align_up(cb, 4) * 4, 256);
  int sk = 1;  // the variant depends on ragged_dim, which this query does not know: size for the largest split
  for (int v = 0; v < 3; ++v) sk = max(sk, choose_splitk(M, N, K, v));
  const size_t p = sk > 1 ? align_up((size_t)sk * M * align_up(N, 4) * 4, 256) : 0;
  return a + b + p + 256;
}

// forced != 0: only hard requirements (alignment, driver entry point); else also the size heuristic
int immtsf_gemm_tc_eligible(int forced, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                            const float* C, int ldc) {
  if (M < 1 || N < 1 || K < 1) return 0;
  if (((uintptr_t)A & 15) != 0 || (lda & 3) != 0 || ((uintptr_t)B & 15) != 0 || (ldb & 3) != 0) return 0;  // TMA: 16 B strides
  if (((uintptr_t)C & 15) != 0 || (ldc & 3) != 0) return 0;  // epilogue stores 128-bit
  if (!forced) {
    if (M < 64 || N < 32 || K < 32) return 0;  // tiny / skinny: CUDA cores
    if ((double)M * N * K < 4.0e6) return 0;
  }
  return get_encode() != nullptr;
}

static int launch_split_lo(const float* src, int ld, int rows, int cols, float* lo, int ld_lo, const int32_t* ragged,
                           int ragged_flag, cudaStream_t st) {
  const size_t tot = (size_t)rows * (ld_lo / 4);
  int grid = (int)((tot + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  split_lo_kernel<<<grid, 256, 0, st>>>(src, ld, rows, cols, lo, ld_lo, ragged, ragged_flag);
  IMMTSF_CHECK_LAUNCH("split_lo");
  return IMMTSF_OK;
}

// lo[rows][roundup(cols,4)] = x - trunc_tf32(x); rows bounded by roundup(*ragged, 128) when ragged != NULL
extern "C" int immtsf_split_lo(const float* src, int ld, int rows, int cols, float* lo, int ld_lo, const int32_t* ragged,
                               void* stream) {
  if (rows == 0 || cols == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(src && lo, "split_lo: null pointer");
  IMMTSF_REQUIRE(((uintptr_t)src & 15) == 0 && (ld & 3) == 0 && ((uintptr_t)lo & 15) == 0 && (ld_lo & 3) == 0 && ld_lo >= cols,
                 "split_lo: operands must be 16B aligned with ld %% 4 == 0 and ld_lo >= cols");
  return launch_split_lo(src, ld, rows, cols, lo, ld_lo, ragged, ragged != nullptr, (cudaStream_t)stream);
}

// dst[c][r] = src[r][c] (32 x 32 tiles through shared memory) and, when lo != NULL, lo[c][r] = dst - trunc_tf32(dst).
// The tcgen05 kernel reads K-major operands with 128B-swizzled boxes at full rate; an MN-major fp32 operand goes through the
// "32B atom" path, measured 1.7x slower on the data-gradient products (dx = dy W: 46 us against 28 us at 4096 x 768 x 768).
// Transposing the WEIGHT once (2.4 MB) is cheaper than paying that on every row of the batch.
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ src, int ld, int rows, int cols,
                                                               float* __restrict__ dst, int ldd, float* __restrict__ lo, int ldl) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? __ldg(src + (size_t)r * ld + c) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;  // output row = source column
    if (c < cols && r < rows) {
      const float v = tile[tx][j];
      dst[(size_t)c * ldd + r] = v;
      if (lo != nullptr) lo[(size_t)c * ldl + r] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    }
  }
}

extern "C" int immtsf_transpose_split(const float* src, int ld, int rows, int cols, float* dst, int ldd, float* lo, int ldl,
                                      void* stream) {
  if (rows == 0 || cols == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(src && dst && ld >= cols && ldd >= rows && (lo == nullptr || ldl >= rows), "transpose_split: bad arguments");
  transpose_split_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32)), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, cols, dst, ldd,
                                                                                                       lo, ldl);
  IMMTSF_CHECK_LAUNCH("transpose_split");
  return IMMTSF_OK;
}

// n <= 16 tasks: task i reads src[i] (rows[i] x cols[i], ld_src[i]) and writes a copy to hi[i] (nullable, ld_hi[i])
// and src - trunc_tf32(src) to lo[i] (nullable, ld_lo[i]).  One launch for all of them.
extern "C" int immtsf_multi_split(int n, const float* const* src, const int* ld_src, const int* rows, const int* cols,
                                  float* const* hi, const int* ld_hi, float* const* lo, const int* ld_lo, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(n > 0 && n <= MS_MAX, "multi_split: 1..16 tasks per call");
  IMMTSF_REQUIRE(src && ld_src && rows && cols && hi && ld_hi && lo && ld_lo, "multi_split: null array");
  MultiSplitArgs a;
  size_t most = 0;
  for (int i = 0; i < MS_MAX; ++i) {
    const int j = i < n ? i : 0;
    IMMTSF_REQUIRE(src[j] != nullptr && rows[j] >= 0 && cols[j] >= 0 && ld_src[j] >= cols[j], "multi_split: bad task %d", j);
    IMMTSF_REQUIRE((hi[j] == nullptr || ld_hi[j] >= cols[j]) && (lo[j] == nullptr || ld_lo[j] >= cols[j]),
                   "multi_split: destination leading dimension too small (task %d)", j);
    a.src[i] = src[j]; a.hi[i] = hi[j]; a.lo[i] = lo[j];
    a.rows[i] = rows[j]; a.cols[i] = cols[j]; a.ld_src[i] = ld_src[j]; a.ld_hi[i] = ld_hi[j]; a.ld_lo[i] = ld_lo[j];
    const size_t tot = (size_t)rows[j] * ((cols[j] + 3) / 4);
    if (i < n && tot > most) most = tot;
  }
  int gx = (int)((most + 255) / 256);
  if (gx > 148 * 2) gx = 148 * 2;
  if (gx < 1) gx = 1;
  multi_split_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(a);
  IMMTSF_CHECK_LAUNCH("multi_split");
  return IMMTSF_OK;
}

int immtsf_gemm_tc(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* A_lo,
                   int lda_lo, const float* B, int ldb, const float* B_lo, int ldb_lo, float beta, float* C, int ldc,
                   float* C_lo, int ldc_lo, const float* bias, const int32_t* ragged, int ragged_dim, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
  if (C_lo != nullptr && (((uintptr_t)C_lo & 15) != 0 || (ldc_lo & 3) != 0 || ldc_lo < (int)align_up(N, 4))) {
    immtsf_set_error("gemm_tc: C_lo must be 16B aligned with ldc_lo %% 4 == 0 and ldc_lo >= roundup(N, 4)");
    return IMMTSF_ERR_ARG;
  }
  const size_t need = immtsf_gemm_tc_workspace(transA, transB, M, N, K);
  if (workspace == nullptr || workspace_bytes < need) {  // (an upper bound over the variants)
    immtsf_set_error("gemm_tc: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IMMTSF_ERR_ARG;
  }
  const int ra = transA ? K : M, ca = transA ? M : K, rb = transB ? N : K, cb = transB ? K : N;
  int lda2 = (int)align_up(ca, 4), ldb2 = (int)align_up(cb, 4);
  const size_t abytes = align_up((size_t)ra * lda2 * 4, 256), bbytes = align_up((size_t)rb * ldb2 * 4, 256);
  uint8_t* w = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const float* Al = A_lo; const float* Bl = B_lo;
  // rows of A are ragged when (ragged_dim==1 && !transA) or (ragged_dim==2 && transA); rows of B when ragged_dim==2 && !transB
  const int a_ragged = (ragged_dim == 1 && !transA) || (ragged_dim == 2 && transA);
  const int b_ragged = (ragged_dim == 2 && !transB);
  if (Al == nullptr) {
    int rc = launch_split_lo(A, lda, ra, ca, (float*)w, lda2, ragged, a_ragged, st);
    if (rc) return rc;
    Al = (const float*)w;
  } else {
    lda2 = lda_lo;
  }
  if (Bl == nullptr) {
    int rc = launch_split_lo(B, ldb, rb, cb, (float*)(w + abytes), ldb2, ragged, b_ragged, st);
    if (rc) return rc;
    Bl = (const float*)(w + abytes);
  } else {
    ldb2 = ldb_lo;
  }
  CUtensorMap mAh, mAl, mBh, mBl;
  // K-major operand [rows=MN][cols=K]: box 32 x 128 (or 256) ; MN-major operand [rows=K][cols=MN]: box 32 x 32
  const int variant = choose_variant(M, N, K, ragged_dim);
  const int bn = variant_bn(variant);
  // pairs: every CTA fetches its own 128-column half of B
  const int boxA = transA ? 32 : BM, boxB = transB ? (variant == VPAIR ? Cfg2::BNH : bn) : 32;
  if (make_map(&mAh, A, ra, ca, lda, boxA, transA != 0) || make_map(&mAl, Al, ra, ca, lda2, boxA, transA != 0) ||
      make_map(&mBh, B, rb, cb, ldb, boxB, transB == 0) || make_map(&mBl, Bl, rb, cb, ldb2, boxB, transB == 0)) {
    immtsf_set_error("gemm_tc: cuTensorMapEncodeTiled failed");
    return IMMTSF_ERR_LAUNCH;
  }
  TcArgs g;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.bias = bias;
  g.ragged = ragged; g.ragged_dim = ragged_dim;
  g.batched = 0; g.batch2 = 1; g.c_s1 = 0; g.c_s2 = 0; g.trace = g_trace; g.C_lo = C_lo; g.ldc_lo = ldc_lo;
  dim3 grid(ceil_div(N, bn), ceil_div(M, BM));
  if (variant == VPAIR) grid = dim3(2 * ceil_div(M, 2 * BM), ceil_div(N, 256));  // x: CTA pairs along M (cluster 2x1x1)
  const int splitk = choose_splitk(M, N, K, variant);
  g.partial = nullptr; g.ldp = 0;
  if (splitk > 1) {
    grid.z = splitk;
    g.partial = (float*)(w + abytes + bbytes);
    g.ldp = (int)align_up(N, 4);
  }
  static bool attr_done = false;
  if (!attr_done) {
#define TC_ATTR(a, b)                                                                                                      \
  cudaFuncSetAttribute(gemm_tc_kernel<a, b, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES);     \
  cudaFuncSetAttribute(gemm_tc_kernel<a, b, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES)
    TC_ATTR(false, false); TC_ATTR(false, true); TC_ATTR(true, false); TC_ATTR(true, true);
#undef TC_ATTR
    cudaFuncSetAttribute(gemm_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
    attr_done = true;
  }
  ProfRec* rec = (g_prof_on && g_prof_n < g_prof_cap) ? &g_prof[g_prof_n++] : nullptr;
  if (rec) { rec->M = M; rec->N = N; rec->K = K; rec->ragged_dim = ragged_dim; cudaEventRecord(rec->e0, st); }
  // UMMA "B is K-major" means stored [N][K], i.e. transB=1
#define TC_LAUNCH(a, b)                                                                                                   \
  do {                                                                                                                    \
    if (variant == VPAIR) gemm_tc2_kernel<a, b><<<grid, Cfg2::THREADS, Cfg2::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, g);      \
    else if (bn == 128) gemm_tc_kernel<a, b, 128><<<grid, Cfg<128>::THREADS, Cfg<128>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, g); \
    else gemm_tc_kernel<a, b, 256><<<grid, Cfg<256>::THREADS, Cfg<256>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, g);          \
  } while (0)
  if (!transA && transB) TC_LAUNCH(false, false);
  else if (!transA && !transB) TC_LAUNCH(false, true);
  else if (transA && !transB) TC_LAUNCH(true, true);
  else TC_LAUNCH(true, false);
#undef TC_LAUNCH
  if (rec) cudaEventRecord(rec->e1, st);
  IMMTSF_CHECK_LAUNCH("gemm_tc");
  if (splitk > 1) {
    const size_t tot = (size_t)M * (g.ldp / 4);
    int rg = (int)((tot + 255) / 256); if (rg > 148 * 8) rg = 148 * 8;
    splitk_reduce_kernel<<<rg, 256, 0, st>>>(C, ldc, M, N, alpha, beta, bias, g.partial, g.ldp, splitk, ragged, ragged_dim, C_lo,
                                             ldc_lo);
    IMMTSF_CHECK_LAUNCH("splitk_reduce");
  }
  return IMMTSF_OK;
}


// ---------------------------------------------------------------- grouped products
// n <= 4 independent products C_i = alpha_i op(A_i) op(B_i) + beta_i C_i in one launch (see gemm_tc_group_kernel).
// Every operand comes with its lo part (immtsf_split_lo / a producer's C_lo); C_lo[i] (nullable) as in immtsf_gemm_ex.
extern "C" int immtsf_gemm_group(int n, const int* transA, const int* transB, const int* M, const int* N, const int* K,
                                 const float* alpha, const float* const* A, const float* const* A_lo, const int* lda,
                                 const int* lda_lo, const float* const* B, const float* const* B_lo, const int* ldb,
                                 const int* ldb_lo, const float* beta, float* const* C, const int* ldc, float* const* C_lo,
                                 const int* ldc_lo, void* stream) {
  if (n == 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(n > 0 && n <= GRP_MAX, "gemm_group: 1..4 problems per call");
  IMMTSF_REQUIRE(get_encode() != nullptr, "gemm_group: cuTensorMapEncodeTiled unavailable");
  GroupArgs grp;
  int gx = 1, gy = 1;
  for (int i = 0; i < n; ++i) {
    IMMTSF_REQUIRE(M[i] > 0 && N[i] > 0 && K[i] > 0 && A[i] && A_lo[i] && B[i] && B_lo[i] && C[i], "gemm_group: bad problem %d", i);
    IMMTSF_REQUIRE(immtsf_gemm_tc_eligible(1, M[i], N[i], K[i], A[i], lda[i], B[i], ldb[i], C[i], ldc[i]) &&
                       ((uintptr_t)A_lo[i] & 15) == 0 && (lda_lo[i] & 3) == 0 && ((uintptr_t)B_lo[i] & 15) == 0 && (ldb_lo[i] & 3) == 0,
                   "gemm_group: problem %d: operands must be 16B aligned with leading dimensions %% 4 == 0", i);
    IMMTSF_REQUIRE(C_lo[i] == nullptr || (((uintptr_t)C_lo[i] & 15) == 0 && (ldc_lo[i] & 3) == 0 && ldc_lo[i] >= (int)align_up(N[i], 4)),
                   "gemm_group: problem %d: bad C_lo", i);
    const int tA = transA[i], tB = transB[i];
    const int ra = tA ? K[i] : M[i], ca = tA ? M[i] : K[i], rb = tB ? N[i] : K[i], cb = tB ? K[i] : N[i];
    const int boxA = tA ? 32 : BM, boxB = tB ? 128 : 32;
    if (make_map(&grp.maps[i][0], A[i], ra, ca, lda[i], boxA, tA != 0) || make_map(&grp.maps[i][1], A_lo[i], ra, ca, lda_lo[i], boxA, tA != 0) ||
        make_map(&grp.maps[i][2], B[i], rb, cb, ldb[i], boxB, tB == 0) || make_map(&grp.maps[i][3], B_lo[i], rb, cb, ldb_lo[i], boxB, tB == 0)) {
      immtsf_set_error("gemm_group: cuTensorMapEncodeTiled failed (problem %d)", i);
      return IMMTSF_ERR_LAUNCH;
    }
    TcArgs& g = grp.g[i];
    g.C = C[i]; g.ldc = ldc[i]; g.M = M[i]; g.N = N[i]; g.K = K[i]; g.alpha = alpha[i]; g.beta = beta[i]; g.bias = nullptr;
    g.ragged = nullptr; g.ragged_dim = 0; g.partial = nullptr; g.ldp = 0;
    g.batched = 0; g.batch2 = 1; g.c_s1 = 0; g.c_s2 = 0; g.trace = nullptr; g.C_lo = C_lo[i]; g.ldc_lo = ldc_lo[i];
    grp.a_mn[i] = tA ? 1 : 0;   // UMMA "A is MN-major" = stored [K][M]
    grp.b_mn[i] = tB ? 0 : 1;   // UMMA "B is K-major" = stored [N][K] = transB
    gx = max(gx, ceil_div(N[i], 128));
    gy = max(gy, ceil_div(M[i], BM));
  }
  for (int i = n; i < GRP_MAX; ++i) {
    for (int j = 0; j < 4; ++j) grp.maps[i][j] = grp.maps[0][j];
    grp.g[i] = grp.g[0]; grp.a_mn[i] = grp.a_mn[0]; grp.b_mn[i] = grp.b_mn[0];
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gemm_tc_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES);
    attr_done = true;
  }
  // per-launch timing (bench.py): one record for the group, M scaled so that 2*M*N*K equals the group's FLOPs
  ProfRec* rec = (g_prof_on && g_prof_n < g_prof_cap) ? &g_prof[g_prof_n++] : nullptr;
  if (rec) {
    double f = 0.0;
    for (int i = 0; i < n; ++i) f += (double)M[i] * N[i] * K[i];
    rec->M = (int)(f / ((double)N[0] * K[0]) + 0.5); rec->N = N[0]; rec->K = K[0]; rec->ragged_dim = 0;
    cudaEventRecord(rec->e0, (cudaStream_t)stream);
  }
  gemm_tc_group_kernel<<<dim3(gx, gy, n), Cfg<128>::THREADS, Cfg<128>::SMEM_BYTES, (cudaStream_t)stream>>>(grp);
  if (rec) cudaEventRecord(rec->e1, (cudaStream_t)stream);
  IMMTSF_CHECK_LAUNCH("gemm_tc_group");
  return IMMTSF_OK;
}

// ---------------------------------------------------------------- batched products (attention contractions)
// C(b1,b2)[M,N] = alpha * op(A(b1,b2)) op(B(b1,b2)) + beta * C(b1,b2), X(b1,b2) = X + b1*x_s1 + b2*x_s2 (element strides).
// Operands are read through 4-D tensor maps, so tiles that overhang a batch's rows / columns are zero-filled by TMA
// and the contraction never bleeds into the neighbouring batch.  The lo parts are split over each operand's whole
// flat extent into the workspace.
static size_t flat_extent(int rows, int cols, long ld, long s1, long s2, int batch1, int batch2) {
  return (size_t)((long)(batch1 - 1) * s1 + (long)(batch2 - 1) * s2 + (long)(rows - 1) * ld + cols);
}

extern "C" size_t immtsf_gemm_batched_workspace_bytes(int transA, int transB, int M, int N, int K, int lda, long a_s1, long a_s2,
                                                      int ldb, long b_s1, long b_s2, int batch1, int batch2) {
  const int ra = transA ? K : M, ca = transA ? M : K, rb = transB ? N : K, cb = transB ? K : N;
  return align_up(flat_extent(ra, ca, lda, a_s1, a_s2, batch1, batch2) * 4 + 16, 256) +
         align_up(flat_extent(rb, cb, ldb, b_s1, b_s2, batch1, batch2) * 4 + 16, 256) + 256;
}

extern "C" int immtsf_gemm_batched(int transA, int transB, int M, int N, int K, float alpha, const float* A, const float* A_lo,
                                   int lda, long a_s1, long a_s2, const float* B, const float* B_lo, int ldb, long b_s1,
                                   long b_s2, float beta, float* C, int ldc, long c_s1, long c_s2, int batch1, int batch2,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0 || batch1 <= 0 || batch2 <= 0) return IMMTSF_OK;
  IMMTSF_REQUIRE(A && B && C && K >= 1, "gemm_batched: null operand or K < 1");
  IMMTSF_REQUIRE((long)batch1 * batch2 <= 65535, "gemm_batched: at most 65535 batches");
  auto ok16 = [](const void* p, long ld, long s1, long s2) { return ((uintptr_t)p & 15) == 0 && (ld & 3) == 0 && (s1 & 3) == 0 && (s2 & 3) == 0; };
  IMMTSF_REQUIRE(ok16(A, lda, a_s1, a_s2) && ok16(B, ldb, b_s1, b_s2) && ok16(C, ldc, c_s1, c_s2),
                 "gemm_batched: operands must be 16B aligned with every stride a multiple of 4 floats");
  IMMTSF_REQUIRE(get_encode() != nullptr, "gemm_batched: cuTensorMapEncodeTiled unavailable");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = immtsf_gemm_batched_workspace_bytes(transA, transB, M, N, K, lda, a_s1, a_s2, ldb, b_s1, b_s2, batch1, batch2);
  IMMTSF_REQUIRE(workspace != nullptr && workspace_bytes >= need, "gemm_batched: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
  const int ra = transA ? K : M, ca = transA ? M : K, rb = transB ? N : K, cb = transB ? K : N;
  const size_t ea = flat_extent(ra, ca, lda, a_s1, a_s2, batch1, batch2), eb = flat_extent(rb, cb, ldb, b_s1, b_s2, batch1, batch2);
  uint8_t* w = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  // lo over the flat extents (1 x extent "matrices"; the tail beyond a multiple of 4 is handled by the kernel),
  // unless the caller hands in lo buffers it made earlier with immtsf_split_lo(src, ., 1, extent, ...)
  const float* Al = A_lo;
  const float* Bl = B_lo;
  if (Al == nullptr) {
    int rc = launch_split_lo(A, (int)align_up(ea, 4), 1, (int)ea, (float*)w, (int)align_up(ea, 4), nullptr, 0, st);
    if (rc) return rc;
    Al = (const float*)w;
  }
  if (Bl == nullptr) {
    float* dst = (float*)(w + align_up(ea * 4 + 16, 256));
    int rc = launch_split_lo(B, (int)align_up(eb, 4), 1, (int)eb, dst, (int)align_up(eb, 4), nullptr, 0, st);
    if (rc) return rc;
    Bl = dst;
  }
  const int bn = (N > 128 && ceil_div(N, 256) * ceil_div(M, BM) * batch1 * batch2 >= 148) ? 256 : 128;
  const int boxA = transA ? 32 : BM, boxB = transB ? bn : 32;
  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map4(&mAh, A, ra, ca, lda, a_s2, a_s1, batch2, batch1, boxA, transA != 0) ||
      make_map4(&mAl, Al, ra, ca, lda, a_s2, a_s1, batch2, batch1, boxA, transA != 0) ||
      make_map4(&mBh, B, rb, cb, ldb, b_s2, b_s1, batch2, batch1, boxB, transB == 0) ||
      make_map4(&mBl, Bl, rb, cb, ldb, b_s2, b_s1, batch2, batch1, boxB, transB == 0)) {
    immtsf_set_error("gemm_batched: cuTensorMapEncodeTiled failed");
    return IMMTSF_ERR_LAUNCH;
  }
  TcArgs g;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta; g.bias = nullptr;
  g.ragged = nullptr; g.ragged_dim = 0; g.partial = nullptr; g.ldp = 0;
  g.batched = 1; g.batch2 = batch2; g.c_s1 = c_s1; g.c_s2 = c_s2; g.trace = nullptr; g.C_lo = nullptr; g.ldc_lo = 0;
  dim3 grid(ceil_div(N, bn), ceil_div(M, BM), batch1 * batch2);
  static bool attr_done = false;
  if (!attr_done) {
#define TC_ATTR(a, b)                                                                                                      \
  cudaFuncSetAttribute(gemm_tc_kernel<a, b, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM_BYTES);     \
  cudaFuncSetAttribute(gemm_tc_kernel<a, b, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM_BYTES)
    TC_ATTR(false, false); TC_ATTR(false, true); TC_ATTR(true, false); TC_ATTR(true, true);
#undef TC_ATTR
    attr_done = true;
  }
#define TC_LAUNCH(a, b)                                                                                                   \
  do {                                                                                                                    \
    if (bn == 128) gemm_tc_kernel<a, b, 128><<<grid, Cfg<128>::THREADS, Cfg<128>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, g); \
    else gemm_tc_kernel<a, b, 256><<<grid, Cfg<256>::THREADS, Cfg<256>::SMEM_BYTES, st>>>(mAh, mAl, mBh, mBl, g);          \
  } while (0)
  if (!transA && transB) TC_LAUNCH(false, false);
  else if (!transA && !transB) TC_LAUNCH(false, true);
  else if (transA && !transB) TC_LAUNCH(true, true);
  else TC_LAUNCH(true, false);
#undef TC_LAUNCH
  IMMTSF_CHECK_LAUNCH("gemm_tc_batched");
  return IMMTSF_OK;
}
