// gpt_wrappers.cu — fused, family-independent episode-statistics and reward-normalisation wrappers for sm_100a
// (SURVEY.md §8f row 3).
//
// The reference's author stacks gymnasium's RecordEpisodeStatistics and NormalizeReward on the vector envs
// (gym_po/tester.py:36-41).  Those wrappers are numpy passes over [B] arrays; here both run on the device, on the
// env's own output arrays, so a training loop never brings reward / done to the host:
//
//   wrap_accumulate_kernel  (one pass, HBM-bound)
//     RecordEpisodeStatistics: ep_return += r, ep_length += 1; where done: report (return, length) in the info
//       arrays, add the episode to the global statistics vector {episodes, sum r, sum l, sum r^2, env-steps}
//       (per-thread partial sums -> warp shuffle -> one atomicAdd(double) per warp and field), clear.
//     NormalizeReward, first half: disc = disc * gamma * (1 - terminated) + r, and the batch moments
//       (sum disc, sum disc^2) of this step in float64.
//   wrap_normalize_kernel   (second pass over reward only)
//     merges the batch moments into the running mean / variance (Chan et al. parallel update, exactly
//     gymnasium's RunningMeanStd.update) and writes reward / sqrt(var + epsilon).  Every thread computes the
//     (identical) merge from the 5 doubles it reads; thread 0 of block 0 publishes the new running state into the
//     OTHER half of a double buffer, so there is no grid-wide barrier and no race.
//
// Algorithmic bytes per env-step: RECORD 4+1+1 read, 8+8 state r/w, 8 info written = 30 B; NORMALIZE adds
// 8 B (disc r/w) in pass 1 and 8 B in pass 2.
#include "gpt_internal.h"

struct gpt_wrap {
  int device = 0;
  int flags = 0;
  int64_t num_envs = 0, capacity = 0;
  double gamma = 0.99, epsilon = 1e-8;
  // device: [0..7] statistics vector (layout of gpt_stats_ptr), [8..13] two halves of {count, mean, var},
  // [14..17] two halves of the batch moments {sum, sum of squares}
  double* d_state = nullptr;
  uint64_t step = 0;
  int64_t launches = 0;
};

namespace gpt {

struct WrapParams {
  const float* reward;
  const uint8_t* terminated;
  const uint8_t* truncated;
  float* ep_return;
  int32_t* ep_length;
  float* last_return;
  int32_t* last_length;
  float* disc_return;
  float* norm_reward;
  double* state;
  int64_t num_envs, capacity;
  float gamma;
  double epsilon;
  int32_t cur;      // which half of the double buffers this step reads (rms) / accumulates into (moments)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

template <bool RECORD, bool NORMALIZE>
__global__ void __launch_bounds__(128) wrap_accumulate_kernel(const __grid_constant__ WrapParams P) {
  pdl_launch_dependents();
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kQuad;
  pdl_wait();
  EpisodeAcc acc;
  double s1 = 0.0, s2 = 0.0;
  if (q < P.capacity) {
    const float4 r4 = __ldcs(reinterpret_cast<const float4*>(P.reward + q));
    const uint32_t t4 = ld_stream(reinterpret_cast<const uint32_t*>(P.terminated + q));
    const uint32_t u4 = ld_stream(reinterpret_cast<const uint32_t*>(P.truncated + q));
    const float r[4] = {r4.x, r4.y, r4.z, r4.w};
    if constexpr (RECORD) {
      const float4 er4 = __ldcs(reinterpret_cast<const float4*>(P.ep_return + q));
      const int4 el4 = ld_stream(reinterpret_cast<const int4*>(P.ep_length + q));
      float er[4] = {er4.x, er4.y, er4.z, er4.w}, lr[4];
      int32_t el[4] = {el4.x, el4.y, el4.z, el4.w}, ll[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool done = (((t4 | u4) >> (8 * k)) & 0xFFu) != 0;
        er[k] += r[k];
        el[k] += 1;
        lr[k] = done ? er[k] : 0.f;
        ll[k] = done ? el[k] : 0;
        if (q + k < P.num_envs) {
          acc.steps += 1.f;
          if (done) acc.finish(er[k], el[k]);
        }
        er[k] = done ? 0.f : er[k];
        el[k] = done ? 0 : el[k];
      }
      st_stream(reinterpret_cast<float4*>(P.ep_return + q), make_float4(er[0], er[1], er[2], er[3]));
      st_stream(reinterpret_cast<int4*>(P.ep_length + q), make_int4(el[0], el[1], el[2], el[3]));
      st_stream(reinterpret_cast<float4*>(P.last_return + q), make_float4(lr[0], lr[1], lr[2], lr[3]));
      st_stream(reinterpret_cast<int4*>(P.last_length + q), make_int4(ll[0], ll[1], ll[2], ll[3]));
    }
    if constexpr (NORMALIZE) {
      const float4 d4 = __ldcs(reinterpret_cast<const float4*>(P.disc_return + q));
      float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool term = ((t4 >> (8 * k)) & 0xFFu) != 0;
        d[k] = (term ? 0.f : d[k] * P.gamma) + r[k];
        if (q + k < P.num_envs) {
          s1 += (double)d[k];
          s2 += (double)d[k] * (double)d[k];
        }
      }
      st_stream(reinterpret_cast<float4*>(P.disc_return + q), make_float4(d[0], d[1], d[2], d[3]));
    }
  }
  if constexpr (RECORD) acc.flush(P.state);
  if constexpr (NORMALIZE) {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31u) == 0) {
      atomicAdd(P.state + 14 + 2 * P.cur, s1);
      atomicAdd(P.state + 15 + 2 * P.cur, s2);
    }
  }
}

__global__ void __launch_bounds__(128) wrap_normalize_kernel(const __grid_constant__ WrapParams P) {
  pdl_launch_dependents();
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kQuad;
  pdl_wait();
  // RunningMeanStd.update with the batch moments of this step (gymnasium wrappers/normalize.py)
  const double* rms = P.state + 8 + 3 * P.cur;
  const double count = rms[0], mean = rms[1], var = rms[2];
  const double n = (double)P.num_envs, s1 = P.state[14 + 2 * P.cur], s2 = P.state[15 + 2 * P.cur];
  const double bmean = s1 / n;
  double bvar = s2 / n - bmean * bmean;
  bvar = bvar > 0.0 ? bvar : 0.0;
  const double delta = bmean - mean, tot = count + n;
  const double new_mean = mean + delta * n / tot;
  const double m2 = var * count + bvar * n + delta * delta * count * n / tot;
  const double new_var = m2 / tot;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double* nxt = P.state + 8 + 3 * (P.cur ^ 1);
    nxt[0] = tot;
    nxt[1] = new_mean;
    nxt[2] = new_var;
    P.state[14 + 2 * (P.cur ^ 1)] = 0.0;   // the next step's moment accumulators (last read one step ago)
    P.state[15 + 2 * (P.cur ^ 1)] = 0.0;
  }
  if (q >= P.capacity) return;
  const float inv = (float)(1.0 / sqrt(new_var + P.epsilon));
  const float4 r4 = __ldcs(reinterpret_cast<const float4*>(P.reward + q));
  st_stream(reinterpret_cast<float4*>(P.norm_reward + q), make_float4(r4.x * inv, r4.y * inv, r4.z * inv, r4.w * inv));
}

}  // namespace gpt

using namespace gpt;

extern "C" {

int gpt_wrap_create(int device, int64_t num_envs, int flags, double gamma, double epsilon, gpt_wrap** out) {
  if (!out) return fail(GPT_E_ARG, "gpt_wrap_create: NULL argument");
  *out = nullptr;
  if (num_envs < 1) return fail(GPT_E_ARG, "gpt_wrap_create: num_envs must be >= 1");
  if (!(flags & (GPT_WRAP_RECORD | GPT_WRAP_NORMALIZE)) || (flags & ~(GPT_WRAP_RECORD | GPT_WRAP_NORMALIZE)))
    return fail(GPT_E_ARG, "gpt_wrap_create: flags must combine GPT_WRAP_RECORD / GPT_WRAP_NORMALIZE");
  if (!(gamma >= 0.0 && gamma <= 1.0) || !(epsilon >= 0.0)) return fail(GPT_E_ARG, "gpt_wrap_create: bad gamma / epsilon");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(GPT_E_CUDA, "gpt_wrap_create: no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(GPT_E_ARG, "gpt_wrap_create: bad device ordinal");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  gpt_wrap* w = new (std::nothrow) gpt_wrap();
  if (!w) return fail(GPT_E_ARG, "gpt_wrap_create: out of host memory");
  w->device = device;
  w->flags = flags;
  w->num_envs = num_envs;
  w->capacity = (num_envs + GPT_ENV_ALIGN - 1) / GPT_ENV_ALIGN * GPT_ENV_ALIGN;
  w->gamma = gamma;
  w->epsilon = epsilon;
  e = cudaMalloc((void**)&w->d_state, 18 * sizeof(double));
  if (e != cudaSuccess) { delete w; return cuda_fail(e, "cudaMalloc(wrapper state)"); }
  double init[18] = {0};
  init[8] = 1e-4; init[9] = 0.0; init[10] = 1.0;   // RunningMeanStd(): count = 1e-4, mean 0, var 1
  e = cudaMemcpy(w->d_state, init, sizeof(init), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(w->d_state); delete w; return cuda_fail(e, "cudaMemcpy(wrapper state)"); }
  *out = w;
  return GPT_OK;
}

int gpt_wrap_destroy(gpt_wrap* w) {
  if (!w) return GPT_OK;
  cudaFree(w->d_state);
  delete w;
  return GPT_OK;
}

int gpt_wrap_step(gpt_wrap* w, const gpt_wrap_io* io, void* stream) {
  if (!w || !io) return fail(GPT_E_ARG, "gpt_wrap_step: NULL argument");
  const bool rec = w->flags & GPT_WRAP_RECORD, nrm = w->flags & GPT_WRAP_NORMALIZE;
  if (!io->reward || !io->terminated || !io->truncated) return fail(GPT_E_UNBOUND, "gpt_wrap_step: reward / terminated / truncated missing");
  if (rec && (!io->ep_return || !io->ep_length || !io->last_return || !io->last_length))
    return fail(GPT_E_UNBOUND, "gpt_wrap_step: RECORD needs ep_return, ep_length, last_return, last_length");
  if (nrm && (!io->disc_return || !io->norm_reward)) return fail(GPT_E_UNBOUND, "gpt_wrap_step: NORMALIZE needs disc_return, norm_reward");
  WrapParams P{};
  P.reward = io->reward; P.terminated = io->terminated; P.truncated = io->truncated;
  P.ep_return = io->ep_return; P.ep_length = io->ep_length; P.last_return = io->last_return; P.last_length = io->last_length;
  P.disc_return = io->disc_return; P.norm_reward = io->norm_reward;
  P.state = w->d_state;
  P.num_envs = w->num_envs;
  P.capacity = w->capacity;
  P.gamma = (float)w->gamma;
  P.epsilon = w->epsilon;
  P.cur = (int32_t)(w->step & 1u);
  const int threads = 128;
  const int nblocks = (int)((w->capacity / kQuad + threads - 1) / threads);
  using K = void (*)(const WrapParams);
  K k = rec ? (nrm ? (K)wrap_accumulate_kernel<true, true> : (K)wrap_accumulate_kernel<true, false>) : (K)wrap_accumulate_kernel<false, true>;
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl((void*)k, dim3(nblocks), dim3(threads), 0, (cudaStream_t)stream, args);
  w->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "wrap_accumulate_kernel launch");
  if (nrm) {
    e = launch_pdl((void*)wrap_normalize_kernel, dim3(nblocks), dim3(threads), 0, (cudaStream_t)stream, args);
    w->launches += 1;
    if (e != cudaSuccess) return cuda_fail(e, "wrap_normalize_kernel launch");
    w->step += 1;
  }
  return GPT_OK;
}

int gpt_wrap_state_ptr(gpt_wrap* w, void** device_ptr, int32_t* rms_offset) {
  if (!w || !device_ptr) return fail(GPT_E_ARG, "gpt_wrap_state_ptr: NULL argument");
  *device_ptr = w->d_state;
  if (rms_offset) *rms_offset = 8 + 3 * (int32_t)(w->step & 1u);
  return GPT_OK;
}

int64_t gpt_wrap_launch_count(const gpt_wrap* w) { return w ? w->launches : 0; }

}  // extern "C"
