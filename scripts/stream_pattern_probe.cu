// stream_pattern_probe.cu — what does HBM deliver for the ACCESS PATTERN of the fused Taxi launch, with no compute?
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/bin/stream_pattern_probe scripts/stream_pattern_probe.cu
//   scripts/bin/stream_pattern_probe [log2_envs=22] [T=10]
//
// The fused kernel (csrc/gpt_taxi.cu taxi_table_multi_kernel) moves, per warp tile of 256 envs: state in (s, elapsed:
// 2 x int4 per quad), then for each of T steps one action word in and obs int4 / reward float4 / terminated u32 /
// truncated u32 out at rollout slot t, then the state out.  This probe issues exactly those loads and stores (values
// derived from the loaded words so nothing is optimised away) in three arrangements:
//   fused    one launch, every thread loops over the T steps (the product kernel's arrangement)
//   fused_sync  the same with a __syncthreads() per step (the CTA's 4 KB per array go out together)
//   stepwise T launches, one step each, no state traffic (all CTAs sweep slot t before slot t+1)
// and prints the achieved GB/s on the same algorithmic bytes (11 B per env-step + 16 B per env per launch).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct P {
  int32_t *s, *e;
  const int8_t* act;
  int32_t* obs; float* rew; uint8_t *term, *trunc;
  int64_t cap; int T;
};

template <bool SYNC, bool STATE>
__global__ void __launch_bounds__(128) k_fused(P p, int t0, int nt) {
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t base = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 5)) * 256 + lane * 4;
  int4 sv[2], ev[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int64_t q = base + j * 128;
    if (STATE) { sv[j] = __ldcs((const int4*)(p.s + q)); ev[j] = __ldcs((const int4*)(p.e + q)); }
    else { sv[j] = make_int4(1, 2, 3, 4); ev[j] = make_int4(5, 6, 7, 8); }
  }
#pragma unroll 1
  for (int t = t0; t < t0 + nt; ++t) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t q = base + j * 128, o = (int64_t)t * p.cap + q;
      const uint32_t a = __ldcs((const uint32_t*)(p.act + o));
      sv[j].x += a; sv[j].y ^= a; ev[j].x += 1; ev[j].w += a & 1;
      __stcs((int4*)(p.obs + o), sv[j]);
      __stcs((float4*)(p.rew + o), make_float4(-0.05f, 1.f, -0.5f, (float)ev[j].x));
      __stcs((uint32_t*)(p.term + o), a & 0x01010101u);
      __stcs((uint32_t*)(p.trunc + o), (a >> 1) & 0x01010101u);
    }
    if (SYNC) __syncthreads();
  }
  if (STATE) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int64_t q = base + j * 128;
      __stcs((int4*)(p.s + q), sv[j]);
      __stcs((int4*)(p.e + q), ev[j]);
    }
  }
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 22, T = argc > 2 ? atoi(argv[2]) : 10;
  const int64_t B = 1ll << lg;
  P p; p.cap = B; p.T = T;
  CK(cudaMalloc(&p.s, B * 4)); CK(cudaMalloc(&p.e, B * 4));
  CK(cudaMalloc((void**)&p.act, B * T)); CK(cudaMalloc(&p.obs, B * T * 4)); CK(cudaMalloc(&p.rew, B * T * 4));
  CK(cudaMalloc(&p.term, B * T)); CK(cudaMalloc(&p.trunc, B * T));
  CK(cudaMemset(p.s, 0, B * 4)); CK(cudaMemset(p.e, 0, B * 4)); CK(cudaMemset((void*)p.act, 1, B * T));
  const int grid = (int)(B / 1024), reps = 200;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = (double)B * (11.0 * T + 16.0);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 20; ++r) {
      CK(cudaEventRecord(e0));
      if (mode == 0) k_fused<false, true><<<grid, 128>>>(p, 0, T);
      else if (mode == 1) k_fused<true, true><<<grid, 128>>>(p, 0, T);
      else for (int t = 0; t < T; ++t) k_fused<false, false><<<grid, 128>>>(p, t, 1);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r >= 20) { sum += ms; if (ms < best) best = ms; }
    }
    const double b = mode == 2 ? (double)B * 11.0 * T : bytes;
    printf("%-10s T=%d 2^%d envs: mean %.1f us (%.0f GB/s), best %.1f us (%.0f GB/s)\n",
           mode == 0 ? "fused" : (mode == 1 ? "fused_sync" : "stepwise"), T, lg, sum / reps * 1e3, b / (sum / reps * 1e-3) / 1e9,
           best * 1e3, b / (best * 1e-3) / 1e9);
  }
  return 0;
}
