// stream_pattern_probe2.cu — variants of the fused-launch access pattern (see stream_pattern_probe.cu), to find out
// which part of the pattern costs the gap to the plain-stream bandwidth.  Prints mean us and GB/s per variant.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
struct P { int32_t *s, *e; const int8_t* act; int32_t* obs; float* rew; uint8_t *term, *trunc; int64_t cap; int T; int64_t n_tiles; };

template <int MODE> __device__ __forceinline__ void st16(void* p, int4 v) {
  if (MODE == 0) __stcs((int4*)p, v); else if (MODE == 1) *(int4*)p = v; else __stcg((int4*)p, v);
}
template <int MODE> __device__ __forceinline__ void st4(void* p, uint32_t v) {
  if (MODE == 0) __stcs((uint32_t*)p, v); else if (MODE == 1) *(uint32_t*)p = v; else __stcg((uint32_t*)p, v);
}
// QPT quads per thread; THREADS per CTA; flags: WORDS = obs/reward stores, BYTES = term/trunc stores, ACT = action loads,
// PERSIST = grid-stride loop over warp tiles
template <int QPT, int THREADS, int MODE, bool WORDS, bool BYTES, bool ACT, bool PERSIST>
__global__ void __launch_bounds__(THREADS) k(P p) {
  const uint32_t lane = threadIdx.x & 31u;
  constexpr int EPW = 128 * QPT;
  const int64_t warps = (int64_t)gridDim.x * (THREADS / 32);
  for (int64_t w = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); w < p.n_tiles; w += warps) {
    const int64_t base = w * EPW + lane * 4;
    int4 sv[QPT], ev[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + j * 128;
      sv[j] = __ldcs((const int4*)(p.s + q)); ev[j] = __ldcs((const int4*)(p.e + q));
    }
#pragma unroll 1
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
      for (int j = 0; j < QPT; ++j) {
        const int64_t q = base + j * 128, o = (int64_t)t * p.cap + q;
        uint32_t a = 0x01020304u;
        if (ACT) a = __ldcs((const uint32_t*)(p.act + o));
        sv[j].x += a; sv[j].y ^= a; ev[j].x += 1; ev[j].w += a & 1;
        if (WORDS) {
          st16<MODE>(p.obs + o, sv[j]);
          st16<MODE>(p.rew + o, make_int4(ev[j].x, 1, 2, 3));
        }
        if (BYTES) {
          st4<MODE>(p.term + o, a & 0x01010101u);
          st4<MODE>(p.trunc + o, (a >> 1) & 0x01010101u);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + j * 128;
      __stcs((int4*)(p.s + q), sv[j]); __stcs((int4*)(p.e + q), ev[j]);
    }
    if (!PERSIST) break;
  }
}

template <typename K> static void run(const char* name, K kern, P p, int grid, int threads, double bytes) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float sum = 0, best = 1e30f; const int reps = 100;
  for (int r = 0; r < reps + 10; ++r) {
    CK(cudaEventRecord(e0)); kern<<<grid, threads>>>(p); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 10) { sum += ms; if (ms < best) best = ms; }
  }
  CK(cudaGetLastError());
  printf("%-44s mean %7.1f us %5.0f GB/s | best %7.1f us %5.0f GB/s\n", name, sum / reps * 1e3, bytes / (sum / reps * 1e-3) / 1e9, best * 1e3, bytes / (best * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 22, T = argc > 2 ? atoi(argv[2]) : 10;
  const int64_t B = 1ll << lg;
  P p; p.cap = B; p.T = T;
  CK(cudaMalloc(&p.s, B * 4)); CK(cudaMalloc(&p.e, B * 4));
  CK(cudaMalloc((void**)&p.act, B * T)); CK(cudaMalloc(&p.obs, B * T * 4)); CK(cudaMalloc(&p.rew, B * T * 4));
  CK(cudaMalloc(&p.term, B * T)); CK(cudaMalloc(&p.trunc, B * T));
  CK(cudaMemset(p.s, 0, B * 4)); CK(cudaMemset(p.e, 0, B * 4)); CK(cudaMemset((void*)p.act, 1, B * T));
  const double full = (double)B * (11.0 * T + 16.0);
  p.n_tiles = B / 256;
  run("base q2 t128 .cs", k<2, 128, 0, true, true, true, false>, p, (int)(B / 1024), 128, full);
  run("plain stores", k<2, 128, 1, true, true, true, false>, p, (int)(B / 1024), 128, full);
  run(".cg stores", k<2, 128, 2, true, true, true, false>, p, (int)(B / 1024), 128, full);
  run("no action loads", k<2, 128, 0, true, true, false, false>, p, (int)(B / 1024), 128, (double)B * (10.0 * T + 16.0));
  run("words only (obs, reward)", k<2, 128, 0, true, false, true, false>, p, (int)(B / 1024), 128, (double)B * (9.0 * T + 16.0));
  run("bytes only (term, trunc)", k<2, 128, 0, false, true, true, false>, p, (int)(B / 1024), 128, (double)B * (3.0 * T + 16.0));
  run("q2 t256", k<2, 256, 0, true, true, true, false>, p, (int)(B / 2048), 256, full);
  run("q2 t64", k<2, 64, 0, true, true, true, false>, p, (int)(B / 512), 64, full);
  p.n_tiles = B / 512;
  run("q4 t128", k<4, 128, 0, true, true, true, false>, p, (int)(B / 2048), 128, full);
  run("q4 t128 plain", k<4, 128, 1, true, true, true, false>, p, (int)(B / 2048), 128, full);
  p.n_tiles = B / 128;
  run("q1 t128", k<1, 128, 0, true, true, true, false>, p, (int)(B / 512), 128, full);
  run("q1 t256", k<1, 256, 0, true, true, true, false>, p, (int)(B / 1024), 256, full);
  p.n_tiles = B / 256;
  run("persistent q2 t128 x8/SM", k<2, 128, 0, true, true, true, true>, p, 148 * 8, 128, full);
  run("persistent q2 t128 x16/SM", k<2, 128, 0, true, true, true, true>, p, 148 * 16, 128, full);
  run("persistent q2 t256 x8/SM", k<2, 256, 0, true, true, true, true>, p, 148 * 8, 256, full);
  return 0;
}
