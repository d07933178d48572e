// gpt_car.cu — fused car-flag ("heaven / hell with a priest") step for sm_100a.
//
// One kernel = CarVecEnv.step (reference gym_po/envs/car_flag.py:114-141) + _reset_mask (:97-112) +
// _obs (:143-144, the observation IS the float32 state row) and the discrete-action wrapper (:286-303).
// Arithmetic follows numpy's dtype promotion exactly: the state is float32; with float32 forces all
// math is float32, with float64 forces (or the float64 linspace table of the discrete wrapper) it is
// float64 and rounded to float32 when stored; the priest window is compared in float64.  Compiled
// with -fmad=false, so replayed trajectories are bit-identical to the reference.
//
// HBM layout: obs/state float32 [cap,3] (position, velocity, priest indicator) | flags uint8 (bit0: heaven
// is at +1, bit1: priest is at +0.5) | elapsed int32 | action (float32 | float64 | int8)  ->  the same
// arrays + reward float32, terminated uint8, truncated uint8.  44 B per env-step with float32 actions.
// One thread handles 4 consecutive envs (48 B of state = three 16-byte accesses).
#include "gpt_internal.h"

namespace gpt {

enum : int { kCarF32 = 0, kCarF64 = 1, kCarDiscrete = 2 };

struct CarParams {
  float* s;            // [cap,3]
  uint8_t* flags;
  int32_t* elapsed;
  const void* actions;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  const double* rp_reset_pos;
  const int8_t* rp_heaven;
  const int8_t* rp_priest;
  const double* table;  // discrete forces (device global, n_actions doubles)
  int64_t env_offset;
  int32_t first_tile, n_tiles, time_limit, n_actions;
  double priest_lo[2], priest_hi[2];   // [priest at -0.5, priest at +0.5]: priests -/+ PRIEST_THRESHOLD in float64
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};

template <typename A> __device__ __forceinline__ A car_clip(A v, A lo, A hi) { return fmin(fmax(v, lo), hi); }

template <typename A, int KIND, bool REPLAY, bool DEVCTR = false>
__global__ void __launch_bounds__(128) car_step_kernel(const __grid_constant__ CarParams P) {
  pdl_launch_dependents();
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t q = first + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kQuad;
  if (q >= last) return;
  pdl_wait();
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, 1u);
  float4 sv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) sv[i] = __ldcs(reinterpret_cast<const float4*>(P.s + q * 3) + i);
  float st[12] = {sv[0].x, sv[0].y, sv[0].z, sv[0].w, sv[1].x, sv[1].y, sv[1].z, sv[1].w, sv[2].x, sv[2].y, sv[2].z, sv[2].w};
  const uint32_t fl4 = ld_stream(reinterpret_cast<const uint32_t*>(P.flags + q));
  const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
  int32_t ev[4] = {e4.x, e4.y, e4.z, e4.w};
  float rv[4];
  uint32_t tw = 0, trw = 0, flw = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t env = q + k;
    uint32_t fl = (fl4 >> (8 * k)) & 0xFFu;
    const float heaven = (fl & 1u) ? 1.f : -1.f;
    const uint32_t pi = (fl >> 1) & 1u;
    ev[k] += 1;
    A force;
    if (KIND == kCarDiscrete) {
      uint32_t a = (uint32_t)(uint8_t) reinterpret_cast<const int8_t*>(P.actions)[env];
      a = a < (uint32_t)P.n_actions ? a : (uint32_t)P.n_actions - 1;
      force = (A)P.table[a];
    } else if (KIND == kCarF32) {
      force = (A) reinterpret_cast<const float*>(P.actions)[env];
    } else {
      force = (A) reinterpret_cast<const double*>(P.actions)[env];
    }
    force = car_clip<A>(force, (A)-1.0, (A)1.0);                                         // (:117)
    A vel = car_clip<A>((A)st[3 * k + 1] + force * (A)0.0015, (A)-0.07, (A)0.07);        // (:119-121)
    const A pos = car_clip<A>((A)st[3 * k] + vel, (A)-1.1, (A)1.1);                      // (:122)
    if (pos == (A)-1.1 && vel < (A)0) vel = (A)0;                                         // (:123)
    const bool done = fabs(pos) >= (A)1.0;                                                // (:124)
    const float side = pos > (A)0 ? 1.f : (pos < (A)0 ? -1.f : 0.f);
    rv[k] = done ? (side == heaven ? 1.f : (side == -heaven ? -1.f : 0.f)) : 0.f;         // (:126-128)
    const bool trunc = ev[k] >= P.time_limit;                                             // (:129) note >=
    const double pd = (double)pos;
    const float ind = (pd >= P.priest_lo[pi] && pd <= P.priest_hi[pi]) ? heaven : 0.f;    // (:130-135)
    if (!done) {                                                                          // (:137-139)
      st[3 * k] = (float)pos;
      st[3 * k + 1] = (float)vel;
      st[3 * k + 2] = ind;
    }
    tw |= (done ? 1u : 0u) << (8 * k);
    trw |= (trunc ? 1u : 0u) << (8 * k);
    if (done | trunc) {  // _reset_mask (:97-112): position uniform(-0.2, 0.2), heaven side, priest side
      double p0;
      uint32_t hv, pr;
      if (REPLAY) {
        p0 = P.rp_reset_pos[env];
        hv = P.rp_heaven[env] > 0;
        pr = P.rp_priest[env] > 0;
      } else {
        const uint4 r = rnd_block<DEVCTR>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 0u);
        const double u = (double)(((uint64_t)r.x << 21) ^ (uint64_t)(r.y >> 11)) * (1.0 / 9007199254740992.0);
        p0 = -0.2 + 0.4 * u;          // numpy uniform: low + (high - low) * random()
        hv = r.z >> 31;
        pr = r.w >> 31;
      }
      st[3 * k] = (float)p0;
      st[3 * k + 1] = 0.f;
      st[3 * k + 2] = 0.f;
      ev[k] = 0;
      fl = hv | (pr << 1);
    }
    flw |= fl << (8 * k);
  }
  float4* so = reinterpret_cast<float4*>(P.s + q * 3);
  st_stream(so, make_float4(st[0], st[1], st[2], st[3]));
  st_stream(so + 1, make_float4(st[4], st[5], st[6], st[7]));
  st_stream(so + 2, make_float4(st[8], st[9], st[10], st[11]));
  st_stream(reinterpret_cast<uint32_t*>(P.flags + q), flw);
  st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
  st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
  st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
  st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
}

int car_create(gpt_env* env, const gpt_config* c) {
  if (c->track_stats) return fail(GPT_E_ARG, "car: track_stats is implemented for the Taxi and ROOMS families only");
  if (c->car_num_actions < 0 || c->car_num_actions > 127) return fail(GPT_E_ARG, "car: num_actions must be in [0, 127]");
  if (c->car_num_actions > 0 && !c->car_action_table) return fail(GPT_E_ARG, "car: action table missing");
  if (c->time_limit >= 0x7F7F7F7E) return fail(GPT_E_ARG, "car: time_limit too large");
  std::vector<uint8_t> blob(16, 0);
  if (c->car_num_actions > 0) {
    std::vector<double> t(c->car_action_table, c->car_action_table + c->car_num_actions);
    blob.clear();
    blob_append(blob, t);
  }
  if (int rc = upload_blob(env, blob)) return rc;
  add_array(env, "obs", GPT_ROLE_STATE, GPT_DT_F32, 3);   // the observation is the live state row (car_flag.py:143-144)
  add_array(env, "flags", GPT_ROLE_STATE, GPT_DT_U8, 1);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  add_array(env, "reward", GPT_ROLE_OUTPUT, GPT_DT_F32, 1);
  add_array(env, "terminated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "truncated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "replay_reset_pos", GPT_ROLE_REPLAY, GPT_DT_F64, 1);
  add_array(env, "replay_heaven", GPT_ROLE_REPLAY, GPT_DT_I8, 1);
  add_array(env, "replay_priest", GPT_ROLE_REPLAY, GPT_DT_I8, 1);
  if (c->car_num_actions > 0) {
    env->action_dtype = GPT_DT_I8;
  } else {
    env->action_dtype = c->c_action_f64 ? GPT_DT_F64 : GPT_DT_F32;
  }
  env->action_cols = 1;
  add_array(env, "actions", GPT_ROLE_ACTION, env->action_dtype, 1);
  return GPT_OK;
}

int car_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  CarParams P{};
  P.s = (float*)env->ptr("obs");
  P.flags = (uint8_t*)env->ptr("flags");
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = a.actions;
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.s || !P.flags || !P.elapsed || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "car: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "car: actions is NULL");
  const bool reset = a.mode == kModeReset;
  const int kind = c.car_num_actions > 0 ? kCarDiscrete : (c.c_action_f64 ? kCarF64 : kCarF32);
  if (reset) {  // reset() = every env truncates (poisoned elapsed), forces come from a zeroed scratch: the reward array
    cudaError_t e = cudaMemsetAsync(P.elapsed, 0x7F, (size_t)env->capacity * sizeof(int32_t), a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.reward, 0, (size_t)env->capacity * sizeof(float), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(car reset)");
    // the forces are irrelevant in this launch (every env respawns); any readable memory of the right size will do
    P.actions = kind == kCarDiscrete ? (const void*)P.terminated : (kind == kCarF32 ? (const void*)P.reward : (const void*)P.s);
  }
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  if (replay) {
    P.rp_reset_pos = (const double*)env->ptr("replay_reset_pos");
    P.rp_heaven = (const int8_t*)env->ptr("replay_heaven");
    P.rp_priest = (const int8_t*)env->ptr("replay_priest");
    if (!P.rp_reset_pos || !P.rp_heaven || !P.rp_priest) return fail(GPT_E_UNBOUND, "car: replay arrays must be bound in replay mode");
  }
  P.table = (const double*)env->d_blob;
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.time_limit = c.time_limit;
  P.n_actions = c.car_num_actions;
  P.priest_lo[0] = -0.5 - 0.2;   // self.priests - PRIEST_THRESHOLD, float64 (car_flag.py:131-132)
  P.priest_hi[0] = -0.5 + 0.2;
  P.priest_lo[1] = 0.5 - 0.2;
  P.priest_hi[1] = 0.5 + 0.2;
  P.rng = make_rng_key(env);
  const int threads = 128;
  const int64_t quads = (int64_t)a.n_tiles * (kTileEnvs / kQuad);
  const int nblocks = (int)((quads + threads - 1) / threads);
  if (nblocks <= 0) return GPT_OK;
  using K = void (*)(const CarParams);
  K k;
  if (kind == kCarF32) k = replay ? (K)car_step_kernel<float, kCarF32, true> : (K)car_step_kernel<float, kCarF32, false>;
  else if (kind == kCarF64) k = replay ? (K)car_step_kernel<double, kCarF64, true> : (K)car_step_kernel<double, kCarF64, false>;
  else k = replay ? (K)car_step_kernel<double, kCarDiscrete, true> : (K)car_step_kernel<double, kCarDiscrete, false>;
  P.ctr_ptr = env->d_counter;
  if (env->graph_mode && !replay)   // graph mode: step counter in device memory
    k = kind == kCarF32 ? (K)car_step_kernel<float, kCarF32, false, true>
                        : (kind == kCarF64 ? (K)car_step_kernel<double, kCarF64, false, true> : (K)car_step_kernel<double, kCarDiscrete, false, true>);
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl((const void*)k, dim3(nblocks), dim3(threads), 0, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "car_step_kernel launch");
  if (reset) {
    e = cudaMemsetAsync(P.terminated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.truncated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.reward, 0, (size_t)env->capacity * sizeof(float), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reset outputs)");
  }
  return GPT_OK;
}

}  // namespace gpt
