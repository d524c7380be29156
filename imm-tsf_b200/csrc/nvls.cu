// In-switch (NVLS) all-reduce of the data-parallel gradient statistics: one hand-written kernel per collective, over the
// symmetric-memory arena every rank maps (its own replica, every peer's replica over NVLink, and the MULTICAST address that
// the NVSwitch resolves to all replicas).  SURVEY.md 8e: the only collective of the path is the SUM all-reduce of gradients.
//
// NCCL's all-reduce of the three packed buffers of a cfg2 step costs 34 / 35 / 52 us at 8 GPUs (ring, LL protocol; measured,
// tools/bench_allreduce.py) and its kernels are serialised on the communicator; the statistics are a few MB, so the
// collective is latency-, not bandwidth-bound.  Here, per launch:
//   1. barrier over the ranks (flags in every rank's arena, written with system-scope atomics over NVLink): all inputs final;
//   2. rank r owns the r-th slice: multimem.ld_reduce (the switch adds the N replicas and returns the sum -- one load
//      instead of N) and multimem.st (the switch writes the sum to all N replicas -- one store instead of N);
//   3. barrier: every slice has landed everywhere.
// Every element is summed exactly once, by one rank, so all ranks end up with bit-identical values.  A rank that never
// arrives would leave the others spinning: the spin is bounded and a timeout raises a device-side error flag instead of
// hanging the GPU.
#include "common.cuh"
#include "../../include/immtsf.h"

namespace {

constexpr int NV_MAX_WORLD = 16;
constexpr int NV_MAX_BLOCKS = 32;
constexpr int NV_THREADS = 512;
constexpr unsigned long long NV_SPIN_LIMIT = 1ull << 25;  // x (poll + 40 ns back-off) ~ seconds; then give up and flag the error

__device__ __forceinline__ float4 mm_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Flags: flags[(block * world + src) ] on rank dst = "src has arrived at this barrier of this block".  put: 0 -> 1 on every
// peer (waits for the previous use to have been consumed); wait: 1 -> 0 on my own flags.  Reusable without epochs.
__device__ __forceinline__ bool rank_barrier(float* const* bases, size_t flag_off_floats, int rank, int world, int* err) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int peer = threadIdx.x;
    unsigned int* remote = reinterpret_cast<unsigned int*>(bases[peer] + flag_off_floats) + (size_t)blockIdx.x * world + rank;
    unsigned int* local = reinterpret_cast<unsigned int*>(bases[rank] + flag_off_floats) + (size_t)blockIdx.x * world + peer;
    unsigned long long spins = 0;
    __threadfence_system();
    // (the waiting CTAs sit on SMs that other lanes' kernels are using: back off between polls instead of hammering the
    // issue slots and the NVLink with atomics)
    while (atomicCAS_system(remote, 0u, 1u) != 0u) {
      if (++spins > NV_SPIN_LIMIT) { s_bad = 1; break; }
      __nanosleep(40);
    }
    spins = 0;
    while (atomicCAS_system(local, 1u, 0u) != 1u) {
      if (++spins > NV_SPIN_LIMIT) { s_bad = 1; break; }
      __nanosleep(40);
    }
    __threadfence_system();
  }
  __syncthreads();
  if (s_bad && threadIdx.x == 0 && err != nullptr) *err = 1;
  return s_bad == 0;
}

__global__ void __launch_bounds__(NV_THREADS) nvls_allreduce_kernel(float* __restrict__ mc, float* const* __restrict__ bases,
                                                                    size_t off_floats, size_t n_floats, size_t flag_off_floats,
                                                                    int rank, int world, int* err) {
  if (!rank_barrier(bases, flag_off_floats, rank, world, err)) return;
  // slice of this rank, in float4 units (n_floats is a multiple of 4; the last rank takes the remainder)
  const size_t n4 = n_floats >> 2, per = (n4 + world - 1) / world;
  const size_t lo = per * rank < n4 ? per * rank : n4, hi = lo + per < n4 ? lo + per : n4;
  float* p = mc + off_floats;
  // four independent in-switch reductions in flight per thread (each is a round trip through the NVSwitch)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < hi) v[u] = mm_ld_reduce_add(p + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < hi) mm_st(p + 4 * (i + u * stride), v[u]);
  }
  rank_barrier(bases, flag_off_floats, rank, world, err);
}

}  // namespace

extern "C" size_t immtsf_nvls_flag_bytes(int world) { return (size_t)NV_MAX_BLOCKS * (world > 0 ? world : 1) * sizeof(unsigned int); }

// mc_base: multicast address of the arena; peer_bases_dev: DEVICE array of `world` pointers, entry r = the address at which
// THIS process maps rank r's replica; [off_floats, off_floats + n_floats) is reduced in place on every rank (16-byte aligned,
// n_floats % 4 == 0); flag_off_floats: offset of the zero-initialised flag area (immtsf_nvls_flag_bytes) inside the arena.
extern "C" int immtsf_nvls_allreduce_f32(void* mc_base, void* const* peer_bases_dev, size_t off_floats, size_t n_floats,
                                         size_t flag_off_floats, int rank, int world, int* err_flag_dev, void* stream) {
  if (n_floats == 0 || world <= 1) return IMMTSF_OK;
  IMMTSF_REQUIRE(mc_base && peer_bases_dev, "nvls_allreduce: null pointer");
  IMMTSF_REQUIRE(world <= NV_MAX_WORLD && rank >= 0 && rank < world, "nvls_allreduce: world size %d / rank %d unsupported", world, rank);
  IMMTSF_REQUIRE((off_floats & 3) == 0 && (n_floats & 3) == 0 && ((uintptr_t)mc_base & 15) == 0,
                 "nvls_allreduce: offsets and sizes must be multiples of 4 floats");
  const size_t n4 = n_floats >> 2, per = (n4 + world - 1) / world;
  int blocks = (int)((per + NV_THREADS - 1) / NV_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > NV_MAX_BLOCKS) blocks = NV_MAX_BLOCKS;
  nvls_allreduce_kernel<<<blocks, NV_THREADS, 0, (cudaStream_t)stream>>>((float*)mc_base, (float* const*)peer_bases_dev, off_floats,
                                                                         n_floats, flag_off_floats, rank, world, err_flag_dev);
  IMMTSF_CHECK_LAUNCH("nvls_allreduce");
  return IMMTSF_OK;
}
