// stream_pattern_probe3.cu — round-2 follow-up of stream_pattern_probe{,2}.cu: WHY does the fused rollout pattern
// (state in, T x {action in; obs / reward / terminated / truncated out at slot t}, state out) stop at ~5.3 TB/s when a
// plain copy reaches 6.45 TB/s?  Compute-free kernels, hypotheses tested one at a time:
//   pad      rollout-slot stride = capacity + pad envs (power-of-two strides between the 42 concurrent streams?)
//   skew     every array's base shifted by a different odd multiple of 4 KB
//   waves    env counts that are whole multiples of the resident CTA count (tail of the last wave?)
//   bulk     every warp stages its step outputs in shared memory and writes them with TMA bulk stores
//            (cp.async.bulk.global.shared::cta, 1 KB / 1 KB / 256 B / 256 B per warp and step)
//   consec   a thread owns 16 CONSECUTIVE envs: flag stores become 16-byte stores
//   nstream  pure write of N concurrent contiguous streams (how many streams does the write path sustain?)
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/stream_pattern_probe3 scripts/stream_pattern_probe3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct P {
  int32_t *s, *e; const int8_t* act; int32_t* obs; float* rew; uint8_t *term, *trunc;
  int64_t stride;  // rollout-slot stride in envs
  int T;
};

// ---- the product arrangement: lane l owns quads at tile + j*128 + 4l ------------------------------------------------
template <int QPT, int THREADS>
__global__ void __launch_bounds__(THREADS) k_base(P p) {
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t base = ((int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5)) * (128 * QPT) + lane * 4;
  int4 sv[QPT], ev[QPT];
#pragma unroll
  for (int j = 0; j < QPT; ++j) { sv[j] = __ldcs((const int4*)(p.s + base + j * 128)); ev[j] = __ldcs((const int4*)(p.e + base + j * 128)); }
#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t o = (int64_t)t * p.stride + base + j * 128;
      const uint32_t a = __ldcs((const uint32_t*)(p.act + o));
      sv[j].x += a; sv[j].y ^= a; ev[j].x += 1; ev[j].w += a & 1;
      __stcs((int4*)(p.obs + o), sv[j]);
      __stcs((int4*)(p.rew + o), make_int4(ev[j].x, 1, 2, 3));
      __stcs((uint32_t*)(p.term + o), a & 0x01010101u);
      __stcs((uint32_t*)(p.trunc + o), (a >> 1) & 0x01010101u);
    }
  }
#pragma unroll
  for (int j = 0; j < QPT; ++j) { __stcs((int4*)(p.s + base + j * 128), sv[j]); __stcs((int4*)(p.e + base + j * 128), ev[j]); }
}

// ---- a thread owns 16 consecutive envs (4 quads): 16-byte flag stores ------------------------------------------------
__global__ void __launch_bounds__(128) k_consec(P p) {
  const int64_t base = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 16;
  int4 sv[4], ev[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sv[j] = __ldcs((const int4*)(p.s + base + j * 4)); ev[j] = __ldcs((const int4*)(p.e + base + j * 4)); }
#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
    const int64_t o = (int64_t)t * p.stride + base;
    const int4 a = __ldcs((const int4*)(p.act + o));
    const uint32_t aw[4] = {(uint32_t)a.x, (uint32_t)a.y, (uint32_t)a.z, (uint32_t)a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sv[j].x += aw[j]; sv[j].y ^= aw[j]; ev[j].x += 1; ev[j].w += aw[j] & 1;
      __stcs((int4*)(p.obs + o + j * 4), sv[j]);
      __stcs((int4*)(p.rew + o + j * 4), make_int4(ev[j].x, 1, 2, 3));
    }
    __stcs((int4*)(p.term + o), make_int4(a.x & 0x01010101, a.y & 0x01010101, a.z & 0x01010101, a.w & 0x01010101));
    __stcs((int4*)(p.trunc + o), make_int4((a.x >> 1) & 0x01010101, (a.y >> 1) & 0x01010101, (a.z >> 1) & 0x01010101, (a.w >> 1) & 0x01010101));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { __stcs((int4*)(p.s + base + j * 4), sv[j]); __stcs((int4*)(p.e + base + j * 4), ev[j]); }
}

// ---- warp-level TMA bulk stores ---------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
// per warp and buffer: obs 1 KB | rew 1 KB | term 256 B | trunc 256 B = 2560 B; NBUF buffers
template <int NBUF>
__global__ void __launch_bounds__(128) k_bulk(P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* wb = smem + warp * (2560 * NBUF);
  const int64_t tile = ((int64_t)blockIdx.x * 4 + warp) * 256;
  const int64_t base = tile + lane * 4;
  int4 sv[2], ev[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) { sv[j] = __ldcs((const int4*)(p.s + base + j * 128)); ev[j] = __ldcs((const int4*)(p.e + base + j * 128)); }
#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
    uint8_t* b = wb + (t % NBUF) * 2560;
    if (t >= NBUF) {  // the bulk store that last read this buffer must have finished reading it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t o = (int64_t)t * p.stride + base + j * 128;
      const uint32_t a = __ldcs((const uint32_t*)(p.act + o));
      sv[j].x += a; sv[j].y ^= a; ev[j].x += 1; ev[j].w += a & 1;
      *(int4*)(b + (j * 128 + lane * 4) * 4) = sv[j];
      *(int4*)(b + 1024 + (j * 128 + lane * 4) * 4) = make_int4(ev[j].x, 1, 2, 3);
      *(uint32_t*)(b + 2048 + j * 128 + lane * 4) = a & 0x01010101u;
      *(uint32_t*)(b + 2304 + j * 128 + lane * 4) = (a >> 1) & 0x01010101u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const int64_t o = (int64_t)t * p.stride + tile;
      bulk_store(p.obs + o, b, 1024);
      bulk_store(p.rew + o, b + 1024, 1024);
      bulk_store(p.term + o, b + 2048, 256);
      bulk_store(p.trunc + o, b + 2304, 256);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) { __stcs((int4*)(p.s + base + j * 128), sv[j]); __stcs((int4*)(p.e + base + j * 128), ev[j]); }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- CTA-level TMA bulk stores (4 KB / 4 KB / 1 KB / 1 KB per CTA and step, one CTA barrier per step) --------------
template <int NBUF>
__global__ void __launch_bounds__(128) k_bulk_cta(P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * 1024;
  const uint32_t loc = warp * 256 + lane * 4;  // env index within the CTA tile (j adds 128)
  const int64_t base = tile + loc;
  int4 sv[2], ev[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) { sv[j] = __ldcs((const int4*)(p.s + base + j * 128)); ev[j] = __ldcs((const int4*)(p.e + base + j * 128)); }
#pragma unroll 1
  for (int t = 0; t < p.T; ++t) {
    uint8_t* b = smem + (t % NBUF) * 10240;
    if (t >= NBUF) {
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t o = (int64_t)t * p.stride + base + j * 128;
      const uint32_t a = __ldcs((const uint32_t*)(p.act + o));
      sv[j].x += a; sv[j].y ^= a; ev[j].x += 1; ev[j].w += a & 1;
      *(int4*)(b + (loc + j * 128) * 4) = sv[j];
      *(int4*)(b + 4096 + (loc + j * 128) * 4) = make_int4(ev[j].x, 1, 2, 3);
      *(uint32_t*)(b + 8192 + loc + j * 128) = a & 0x01010101u;
      *(uint32_t*)(b + 9216 + loc + j * 128) = (a >> 1) & 0x01010101u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const int64_t o = (int64_t)t * p.stride + tile;
      bulk_store(p.obs + o, b, 4096);
      bulk_store(p.rew + o, b + 4096, 4096);
      bulk_store(p.term + o, b + 8192, 1024);
      bulk_store(p.trunc + o, b + 9216, 1024);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) { __stcs((int4*)(p.s + base + j * 128), sv[j]); __stcs((int4*)(p.e + base + j * 128), ev[j]); }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- pure write of N concurrent streams (stream n = buf + n*stride_bytes), 16 B per thread and stream ---------------
template <int N>
__global__ void __launch_bounds__(128) k_nstream(uint8_t* buf, int64_t stride_bytes) {
  const int64_t off = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 16;
#pragma unroll
  for (int n = 0; n < N; ++n) __stcs((int4*)(buf + n * stride_bytes + off), make_int4(n, 1, 2, (int)off));
}
// the same bytes written by a loop over streams with T serial "steps" (a CTA visits its chunk of every stream in turn)
__global__ void __launch_bounds__(128) k_nstream_loop(uint8_t* buf, int64_t stride_bytes, int n_streams) {
  const int64_t off = ((int64_t)blockIdx.x * 128 + threadIdx.x) * 16;
#pragma unroll 1
  for (int n = 0; n < n_streams; ++n) __stcs((int4*)(buf + n * stride_bytes + off), make_int4(n, 1, 2, (int)off));
}

template <typename F> static float timeit(F launch, int reps = 100) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float sum = 0;
  for (int r = 0; r < reps + 10; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 10) sum += ms;
  }
  CK(cudaGetLastError());
  return sum / reps;
}
static void report(const char* name, float ms, double bytes) { printf("%-58s %8.1f us %6.0f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); fflush(stdout); }

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 10;
  const int64_t Bmax = (1ll << 22) + (1ll << 20), padmax = 1 << 16, skewmax = 64 * 4096;
  uint8_t* pool; const int64_t slot = Bmax + padmax;
  // one pool, arrays carved out of it so that skews are under control
  const int64_t sz_s = Bmax * 4, sz_act = slot * T, sz_obs = slot * T * 4, sz_fl = slot * T;
  const int64_t total = 2 * sz_s + sz_act + 2 * sz_obs + 2 * sz_fl + 8 * skewmax + (1 << 21);
  CK(cudaMalloc(&pool, total)); CK(cudaMemset(pool, 1, total));
  auto carve = [&](int64_t skew) {
    P p; uint8_t* c = pool; auto al = [&](int64_t n, int k) { uint8_t* r = c + k * skew; c += ((n + skewmax + (1 << 21) / 8 - 1) / 4096 + 1) * 4096; return r; };
    p.s = (int32_t*)al(sz_s, 0); p.e = (int32_t*)al(sz_s, 1); p.act = (const int8_t*)al(sz_act, 2); p.obs = (int32_t*)al(sz_obs, 3);
    p.rew = (float*)al(sz_obs, 5); p.term = (uint8_t*)al(sz_fl, 7); p.trunc = (uint8_t*)al(sz_fl, 11); p.T = T; return p;
  };
  char name[128];
  const int64_t B22 = 1ll << 22;
  const double per_env = 11.0 * T + 16.0;

  printf("== pad / skew (2^22 envs, T=%d, q2 t128) ==\n", T);
  for (int64_t pad : {0ll, 512ll, 1536ll, 2560ll, 4608ll, 33ll * 512, 127ll * 512}) {
    P p = carve(0); p.stride = B22 + pad;
    snprintf(name, sizeof name, "pad %lld envs", (long long)pad);
    report(name, timeit([&] { k_base<2, 128><<<(int)(B22 / 1024), 128>>>(p); }), B22 * per_env);
  }
  for (int64_t skew : {4096ll, 3 * 4096ll, 5 * 4096ll}) {
    P p = carve(skew); p.stride = B22;
    snprintf(name, sizeof name, "skew %lld B x {0,1,2,3,5,7,11}", (long long)skew);
    report(name, timeit([&] { k_base<2, 128><<<(int)(B22 / 1024), 128>>>(p); }), B22 * per_env);
    p.stride = B22 + 2560;
    snprintf(name, sizeof name, "skew %lld B + pad 2560", (long long)skew);
    report(name, timeit([&] { k_base<2, 128><<<(int)(B22 / 1024), 128>>>(p); }), B22 * per_env);
  }

  printf("== whole waves (q2 t128: 1024 envs per CTA; resident = 148 x R) ==\n");
  for (int64_t ctas : {4096ll, 148ll * 8 * 3, 148ll * 8 * 4, 148ll * 16 * 2, 148ll * 12 * 3, 148ll * 10 * 3, 148ll * 24}) {
    const int64_t B = ctas * 1024; P p = carve(0); p.stride = B;
    snprintf(name, sizeof name, "%lld CTAs = %.2f x 148x8 (%lld envs)", (long long)ctas, ctas / 1184.0, (long long)B);
    report(name, timeit([&] { k_base<2, 128><<<(int)ctas, 128>>>(p); }), B * per_env);
  }

  printf("== arrangements (2^22 envs) ==\n");
  {
    P p = carve(0); p.stride = B22;
    report("base q2 t128", timeit([&] { k_base<2, 128><<<(int)(B22 / 1024), 128>>>(p); }), B22 * per_env);
    report("base q1 t128", timeit([&] { k_base<1, 128><<<(int)(B22 / 512), 128>>>(p); }), B22 * per_env);
    report("base q4 t128", timeit([&] { k_base<4, 128><<<(int)(B22 / 2048), 128>>>(p); }), B22 * per_env);
    report("consec 16 envs per thread (16-B flag stores)", timeit([&] { k_consec<<<(int)(B22 / 2048), 128>>>(p); }), B22 * per_env);
    CK(cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2560 * 2));
    CK(cudaFuncSetAttribute(k_bulk<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2560 * 3));
    CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2560 * 4));
    report("warp TMA bulk stores, 2 buffers", timeit([&] { k_bulk<2><<<(int)(B22 / 1024), 128, 4 * 2560 * 2>>>(p); }), B22 * per_env);
    report("warp TMA bulk stores, 3 buffers", timeit([&] { k_bulk<3><<<(int)(B22 / 1024), 128, 4 * 2560 * 3>>>(p); }), B22 * per_env);
    report("warp TMA bulk stores, 4 buffers", timeit([&] { k_bulk<4><<<(int)(B22 / 1024), 128, 4 * 2560 * 4>>>(p); }), B22 * per_env);
    CK(cudaFuncSetAttribute(k_bulk_cta<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10240 * 2));
    CK(cudaFuncSetAttribute(k_bulk_cta<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 10240 * 3));
    report("CTA TMA bulk stores, 2 buffers", timeit([&] { k_bulk_cta<2><<<(int)(B22 / 1024), 128, 10240 * 2>>>(p); }), B22 * per_env);
    report("CTA TMA bulk stores, 3 buffers", timeit([&] { k_bulk_cta<3><<<(int)(B22 / 1024), 128, 10240 * 3>>>(p); }), B22 * per_env);
    p.stride = B22 + 2560;
    report("warp TMA bulk stores, 3 buffers, pad 2560", timeit([&] { k_bulk<3><<<(int)(B22 / 1024), 128, 4 * 2560 * 3>>>(p); }), B22 * per_env);
  }

  printf("== N concurrent pure write streams (each 2^22 x 16 B = 64 MB apart unless padded) ==\n");
  {
    const int64_t n16 = (1ll << 20);  // threads: 16 MB per stream
    for (int64_t sb : {16ll << 20, (16ll << 20) + 4608}) {
      snprintf(name, sizeof name, "1 stream (stride %lld)", (long long)sb); report(name, timeit([&] { k_nstream<1><<<(int)(n16 / 128), 128>>>(pool, sb); }), 16.0 * n16 * 1);
      snprintf(name, sizeof name, "2 streams"); report(name, timeit([&] { k_nstream<2><<<(int)(n16 / 128), 128>>>(pool, sb); }), 16.0 * n16 * 2);
      snprintf(name, sizeof name, "4 streams"); report(name, timeit([&] { k_nstream<4><<<(int)(n16 / 128), 128>>>(pool, sb); }), 16.0 * n16 * 4);
      snprintf(name, sizeof name, "8 streams"); report(name, timeit([&] { k_nstream<8><<<(int)(n16 / 128), 128>>>(pool, sb); }), 16.0 * n16 * 8);
      snprintf(name, sizeof name, "16 streams"); report(name, timeit([&] { k_nstream<16><<<(int)(n16 / 128), 128>>>(pool, sb); }), 16.0 * n16 * 16);
      snprintf(name, sizeof name, "20 streams, rolled loop"); report(name, timeit([&] { k_nstream_loop<<<(int)(n16 / 128), 128>>>(pool, sb, 20); }), 16.0 * n16 * 20);
    }
  }
  return 0;
}
