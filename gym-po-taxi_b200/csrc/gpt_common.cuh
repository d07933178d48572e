// gpt_common.cuh — device helpers shared by the fused env-step kernels (sm_100a).
//
//   * Philox4x32-7 counter RNG (no RNG state in HBM: key = seed, counter = (global env id, step, stream))
//   * streaming 128/64/32-bit global loads / stores (every per-env byte is touched once per step)
//   * TMA bulk copy (cp.async.bulk + mbarrier) that stages the packed static tables into shared memory
//   * exact division of small integers by run-time constants via multiply-high
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef GPT_RESPAWN_INLINE_SINGLE
#define GPT_RESPAWN_INLINE_SINGLE 1   // single-step ROOMS / MSRooms kernels inline the respawn too (measured: hansen8 259 -> 277 G, MSRooms 257 -> 281 G; 0 = out of line)
#endif

namespace gpt {

constexpr int kWarp = 32;
constexpr int kQuad = 4;                       // consecutive envs handled as one vector
constexpr int kQuadsPerThread = 4;             // -> 16 envs per thread
constexpr int kEnvsPerThread = kQuad * kQuadsPerThread;
constexpr int kTileEnvs = kWarp * kEnvsPerThread;  // 512 == GPT_ENV_ALIGN
constexpr int kQuadStride = kWarp * kQuad;     // 128 envs between a thread's consecutive quads

// ------------------------------------------------------------------------------------------
// Philox4x32 (Salmon et al., SC'11).  ctr = {env_lo, env_hi, step_lo, step_hi^stream}, key = seed.
// ------------------------------------------------------------------------------------------
// The key is the same for every thread of a launch, so the host expands the ten round keys once
// (key + r*W) and passes them in the kernel parameters: each round reads its key straight from the
// constant bank instead of spending two integer adds per round per call.
struct RngKey {
  uint32_t rk[20];             // rk[2r], rk[2r+1] = Philox key words of round r (the first kRounds rounds are used)
  uint32_t step_lo, step_hi;   // step counter (incremented by the host once per launch / per step)
};
__host__ __device__ inline void expand_round_keys(RngKey& k, uint64_t seed) {
  uint32_t x = (uint32_t)seed, y = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.rk[2 * r] = x;
    k.rk[2 * r + 1] = y;
    x += 0x9E3779B9u;
    y += 0xBB67AE85u;
  }
}
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint4 c, const RngKey& k) {
  static_assert(ROUNDS >= 7 && ROUNDS <= 10, "Philox4x32 is Crush-resistant from 7 rounds on (Salmon et al., SC'11, Table 2)");
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.rk[2 * r], lo1, hi0 ^ c.w ^ k.rk[2 * r + 1], lo0);
  }
  return c;
}

// Two independent blocks under the same key with their rounds interleaved: each round key is read once and feeds both
// blocks (written out so that the compiler keeps the keys in uniform registers instead of copying them around).
template <int ROUNDS>
__device__ __forceinline__ void philox4x32_x2(uint4& a, uint4& b, const RngKey& k) {
  static_assert(ROUNDS >= 7 && ROUNDS <= 10, "see philox4x32");
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t k0 = k.rk[2 * r], k1 = k.rk[2 * r + 1];
    const uint32_t ah0 = __umulhi(M0, a.x), al0 = M0 * a.x, ah1 = __umulhi(M1, a.z), al1 = M1 * a.z;
    const uint32_t bh0 = __umulhi(M0, b.x), bl0 = M0 * b.x, bh1 = __umulhi(M1, b.z), bl1 = M1 * b.z;
    a = make_uint4(ah1 ^ a.y ^ k0, al1, ah0 ^ a.w ^ k1, al0);
    b = make_uint4(bh1 ^ b.y ^ k0, bl1, bh0 ^ b.w ^ k1, bl0);
  }
}

// Rounds.  Every draw of every family uses Philox4x32-7: the same counter-based generator at the smallest round count
// that passes BigCrush (Salmon et al., SC'11, Table 2; Random123 ships it as philox4x32_7), i.e. without the three
// rounds of safety margin of the default Philox4x32-10.  The per-step slip / noise blocks were up to 30 % of a fused
// kernel's instructions at 10 rounds, and the rare respawn path — taken by ~30 % of the warp iterations under a random
// policy — is latency the whole warp waits for.  Documented deviation: the law of every draw is
// unchanged; tests/philox_host.py restates the generator and checks it against Random123's known-answer vectors.
constexpr int kRounds = 7;
constexpr int kResetRounds = kRounds, kStepRounds = kRounds;

// one Philox block for (global env id, step, stream)
template <int ROUNDS = kResetRounds>
__device__ __forceinline__ uint4 env_random(const RngKey& k, uint64_t env, uint32_t stream) {
  return philox4x32<ROUNDS>(make_uint4((uint32_t)env, (uint32_t)(env >> 32), k.step_lo, k.step_hi ^ (stream << 24)), k);
}

// Graph mode (DEVCTR kernel instantiations): the step counter comes from device memory instead of the launch parameters.
template <bool DEVCTR, int ROUNDS = kResetRounds>
__device__ __forceinline__ uint4 rnd_block(const RngKey& k, uint64_t ctr_dev, uint64_t id, uint32_t stream) {
  if constexpr (DEVCTR)
    return philox4x32<ROUNDS>(make_uint4((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)ctr_dev, ((uint32_t)(ctr_dev >> 32) & 0x00FFFFFFu) ^ (stream << 24)), k);
  else
    return env_random<ROUNDS>(k, id, stream);
}

// unbiased-enough bounded integer: floor(u * n / 2^32); bias <= n * 2^-32
__device__ __forceinline__ uint32_t bounded(uint32_t u, uint32_t n) { return __umulhi(u, n); }

// ------------------------------------------------------------------------------------------
// exact n / d for 0 <= n < 2^16, 1 <= d < 2^16:  q = (n * ceil(2^32/d)) >> 32
// ------------------------------------------------------------------------------------------
struct FastDiv {
  uint32_t magic, d;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.magic = d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return f.d <= 1 ? n : __umulhi(n, f.magic); }

// ------------------------------------------------------------------------------------------
// streaming global access (evict-first: nothing is re-read within a step)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_stream(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldcs(p); }
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(int4* p, int4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint2* p, uint2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(uint32_t* p, uint32_t v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// ------------------------------------------------------------------------------------------
// TMA bulk copy of the static-table blob: global -> shared, completion on an mbarrier.
// SASS: UBLKCP + SYNCS.  `bytes` must be a multiple of 16, both addresses 16-byte aligned.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global bulk copy (TMA store, bulk-group completion): the fused rollout kernels stage a CTA's step outputs
// in shared memory and write each array with ONE of these instead of per-thread stores (scripts/stream_pattern_probe4.cu:
// 6.2 TB/s against 5.6 TB/s on the rollout access pattern).  Generic-proxy writes to the source must be followed by
// fence_proxy_async() (every writer) and a CTA barrier before the issuing thread calls this.
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of the calling thread's committed bulk groups have not finished READING their shared source
template <int N> __device__ __forceinline__ void tma_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Called by every thread of the CTA.  Thread 0 arms the barrier and issues one bulk copy; the
// caller overlaps its first global loads with the copy and calls stage_wait() before the first
// table lookup.
__device__ __forceinline__ void stage_tables_begin(void* smem_dst, const void* gmem_blob, uint32_t bytes, uint64_t* bar) {
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(smem_dst, gmem_blob, bytes, bar);
  }
}
__device__ __forceinline__ void stage_tables_wait(uint64_t* bar) { mbar_wait(bar, 0); }

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): consecutive step kernels on one stream are data dependent
// (step N+1 reads the state step N wrote), but the launch latency, CTA scheduling and the TMA staging
// of the static tables of step N+1 need not wait.  Every kernel signals launch_dependents as soon as it
// starts and calls pdl_wait() before its first access to per-env memory; pdl_wait() returns once the
// preceding grid has completed and its writes are visible.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// Graph mode (DEVCTR kernel instantiations): the Philox step counter lives in device memory — a captured CUDA graph
// replays the same launch parameters, so the kernel itself has to find "which step is this" and advance it.
//   ctr[0] = step counter, ctr[1] = arrival count of the running grid.
// Every thread reads the counter; behind the CTA's barrier (all of the CTA's reads are done) thread 0 announces the CTA
// on the arrival count, and the LAST CTA to arrive — by then every CTA of the grid has read the counter — stores
// counter + n_steps and clears the arrival count.  The next launch reads it after stream order / griddepcontrol.wait
// (= this grid complete and its writes visible), so PDL and fused multi-step launches work under capture and no
// separate tick kernel is needed.  Call once, by all non-exited threads of the CTA, after pdl_wait().
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t devctr_fetch_and_advance(const uint64_t* ctr_ptr, uint32_t n_steps) {
  unsigned long long* c = (unsigned long long*)ctr_ptr;
  const unsigned long long v = __ldcg(c);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(reinterpret_cast<unsigned int*>(c + 1), 1u) == gridDim.x - 1u) {
      __threadfence();
      *reinterpret_cast<volatile unsigned int*>(c + 1) = 0u;
      *reinterpret_cast<volatile unsigned long long*>(c) = v + n_steps;
    }
  }
  return (uint64_t)v;
}

// ------------------------------------------------------------------------------------------
// episode statistics: per-thread partial sums -> warp shuffle reduce -> one atomicAdd per warp and field
// stats[0] episodes, [1] sum of returns, [2] sum of lengths, [3] sum of squared returns, [4] env-steps
// ------------------------------------------------------------------------------------------
struct EpisodeAcc {
  float episodes = 0.f, ret = 0.f, len = 0.f, ret2 = 0.f, steps = 0.f;
  __device__ __forceinline__ void finish(float episode_return, int episode_length) {
    episodes += 1.f;
    ret += episode_return;
    len += (float)episode_length;
    ret2 += episode_return * episode_return;
  }
  __device__ __forceinline__ void flush(double* stats) {
    float v[5] = {episodes, ret, len, ret2, steps};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double d = (double)v[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
      if ((threadIdx.x & 31u) == 0 && d != 0.0) atomicAdd(stats + i, d);
    }
  }
};

}  // namespace gpt
