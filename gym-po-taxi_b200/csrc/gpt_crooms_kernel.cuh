// gpt_crooms_kernel.cuh — fused continuous-position ROOMS step and point-mass Tag step for sm_100a.
//
// CROOMS = CRoomsEnv.step (reference gym_po/envs/rooms/crooms.py:276-298 with _apply_action :300-331,
// _out_of_bounds :333-338, the sample_action closures :175-198, _reset_some :268-274 and
// grid_to_coord / coord_to_grid, rooms/utils.py:7-20).  The reference computes in float64 with numpy;
// this kernel computes in float64 with the same operation order and NO fused multiply-add (the file is
// compiled with -fmad=false), so that on replayed random draws positions, rewards and flags are
// bit-identical.  TAG = the pursuit rules of AntTagEnv (ant_tag.py:105-123, :144-153) on a point mass
// that moves with the CROOMS motion model (see oracle/tag.py and DESIGN.md).
//
// Precision is a template parameter: float64 (default; bit-exact parity with numpy) or float32 (fast mode:
// half the state bytes, single-precision Box-Muller; checked against the oracle within 1e-5 relative).
//
// HBM layout: agent real2 [cap] | goal real2 [cap] (random-goal envs) | velocity real2 [cap]
// (use_velocity) | elapsed int32 | action (int8 | float32x2 | float64x2)  ->  the same state arrays,
// obs, reward float32, terminated uint8, truncated uint8.  One thread handles 4 consecutive envs.
#pragma once
#ifndef GPT_CONT_RESPAWN_ATTR
#define GPT_CONT_RESPAWN_ATTR __forceinline__   // respawn of the continuous envs, inlined (measured: CRooms 115.1 -> 116.7 G, Tag 93.9 -> 95.6 G vs __noinline__)
#endif
#include "gpt_rooms_kernel.cuh"

namespace gpt {

enum : int { kActI8 = 0, kActF32 = 1, kActF64 = 2 };

template <typename R> struct RealTraits;
// quad access: the (y, x) pairs of 4 consecutive envs are contiguous (32 B float / 64 B double), so a lane moves
// them with 128-bit accesses and a warp touches one fully used contiguous segment
template <> struct RealTraits<double> {
  using V2 = double2;
  static __device__ __forceinline__ V2 make(double a, double b) { return make_double2(a, b); }
  static __device__ __forceinline__ void load4(const void* base, int64_t q, V2 (&v)[4]) {
    const double2* p = reinterpret_cast<const double2*>(base) + q;
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __ldcs(p + k);
  }
  static __device__ __forceinline__ void store4(void* base, int64_t q, const V2 (&v)[4]) {
    double2* p = reinterpret_cast<double2*>(base) + q;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(p + k, v[k]);
  }
};
template <> struct RealTraits<float> {
  using V2 = float2;
  static __device__ __forceinline__ V2 make(float a, float b) { return make_float2(a, b); }
  static __device__ __forceinline__ void load4(const void* base, int64_t q, V2 (&v)[4]) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(base) + q);
    const float4 a = __ldcs(p), b = __ldcs(p + 1);
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
  }
  static __device__ __forceinline__ void store4(void* base, int64_t q, const V2 (&v)[4]) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float2*>(base) + q);
    __stcs(p, make_float4(v[0].x, v[0].y, v[1].x, v[1].y));
    __stcs(p + 1, make_float4(v[2].x, v[2].y, v[3].x, v[3].y));
  }
};

// yx actions of a quad (float32 or float64 [cap,2]) widened / narrowed to the state precision
template <typename R>
__device__ __forceinline__ void load_actions4(const void* actions, int act_kind, int64_t q, typename RealTraits<R>::V2 (&push)[4]) {
  if (act_kind == kActF32) {
    float2 a[4];
    RealTraits<float>::load4(actions, q, a);
#pragma unroll
    for (int k = 0; k < 4; ++k) push[k] = RealTraits<R>::make((R)a[k].x, (R)a[k].y);
  } else {
    double2 a[4];
    RealTraits<double>::load4(actions, q, a);
#pragma unroll
    for (int k = 0; k < 4; ++k) push[k] = RealTraits<R>::make((R)a[k].x, (R)a[k].y);
  }
}

struct CRoomsParams {
  void* agent;      // real2 [cap]
  void* goal;       // real2 [cap] (random-goal envs)
  void* velocity;   // real2 [cap] (use_velocity)
  int32_t* elapsed;
  const void* actions;
  void* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  const double* rp_u;
  const double2* rp_noise;
  const double2* rp_resample;
  const int32_t* rp_reset_agent;
  const int32_t* rp_reset_goal;
  const uint8_t* blob;
  // `reserved0` is unused.  It stays because the float32 kernel is sensitive to ptxas' register allocation: without
  // this word in the parameter block the same source gets 72 instead of 71 registers and runs 11 % slower
  // (measured on B200: 103.5 vs 116.9 G env-steps/s, same instruction count).
  uint32_t blob_bytes, nb8_off, room_off, sid_off, valid_off, reserved0, thr64_off, rows_off, stage_off, grid_off, alias_off;
  uint32_t log2n;
  int64_t env_offset;
  int32_t first_tile, n_tiles, mode;
  int32_t h, w, n_actions, n_valid, n_rooms, time_limit, hansen_n, grid_n;
  int32_t act_kind, rgoal, use_velocity, has_noise;
  FastDiv div_w;
  double cell_size, action_std, action_power, goal_threshold, max_y, max_x, goal_y, goal_x;
  float r_step, r_wall, r_goal;
  float f_cell, f_inv_cell, f_half, f_std, f_pow, f_max_y, f_max_x, f_thr2, f_goal_y, f_goal_x;   // float32 fast mode
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};

// two standard normals from 4 x 32 random bits (Box-Muller on 53-bit uniforms)
__device__ __forceinline__ double2 normal_pair(uint4 r) {
  const double u1 = ((double)(((uint64_t)r.x << 21) ^ (uint64_t)(r.y >> 11)) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
  const double u2 = (double)(((uint64_t)r.z << 21) ^ (uint64_t)(r.w >> 11)) * (1.0 / 9007199254740992.0);          // [0,1)
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  return make_double2(rad * c, rad * s);
}

// float32 mode: two standard normals from 2 x 24 random bits (tails reach 5.8 sigma), on the special-function
// unit: lg2, sqrt, sin, cos are one MUFU instruction each (absolute error of sin/cos on [-pi, pi): 2^-21.4)
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 normal_pair_f(uint32_t a, uint32_t b) {
  const float u1 = __fmaf_rn((float)(a >> 8), 0x1p-24f, 0x1p-24f);                               // (0,1]
  const float ang = __fmaf_rn((float)(b >> 8), 6.28318530718f * 0x1p-24f, -3.14159265359f);      // [-pi, pi)
  const float rad = fast_sqrt(-1.38629436112f * __log2f(u1));                                    // sqrt(-2 ln u1)
  return make_float2(rad * __cosf(ang), rad * __sinf(ang));
}

template <typename R> __device__ __forceinline__ R clipd(R v, R lo, R hi) { return fmin(fmax(v, lo), hi); }

// _reset_some (crooms.py:268-274): goal cell first (random-goal envs), then agent cell.  Rare, out of line.
template <bool REPLAY, bool DEVCTR = false>
__device__ GPT_CONT_RESPAWN_ATTR uint32_t crooms_respawn(const CRoomsParams& P, const uint16_t* valid, int64_t env, uint64_t ctr_dev = 0) {
  uint32_t ac, gc = 0;
  if (REPLAY) {
    if (P.rgoal) gc = (uint32_t)P.rp_reset_goal[env];
    ac = (uint32_t)P.rp_reset_agent[env];
  } else {
    const uint4 r = rnd_block<DEVCTR>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 1u);
    if (P.rgoal) gc = valid[bounded(r.y, (uint32_t)P.n_valid)];
    ac = valid[bounded(r.x, (uint32_t)P.n_valid)];
  }
  return ac | (gc << 16);
}

#ifndef GPT_CROOMS_MINB
#define GPT_CROOMS_MINB 7   // measured on B200 (2^22 envs, f32): 1 -> 83 G, 6 -> 106 G, 7 -> 113.5 G, 8 -> 110.6 G env-steps/s
#endif
// SPEC: the configuration the constructor defaults to (yx float32 actions, action noise, no velocity, fixed goal) as
// compile-time constants — the per-env transition becomes ONE basic block, so the four envs of a quad interleave.
template <typename R, int OBS, bool REPLAY, bool DEVCTR = false, bool SPEC = false>
__global__ void __launch_bounds__(128, GPT_CROOMS_MINB) crooms_step_kernel(const __grid_constant__ CRoomsParams P) {
  using V2 = typename RealTraits<R>::V2;
  constexpr bool kFast = sizeof(R) == 4;   // float32 fast mode: reciprocal multiply, squared distances, MUFU noise
  const bool rgoal = SPEC ? false : P.rgoal != 0, use_velocity = SPEC ? false : P.use_velocity != 0, has_noise = SPEC ? true : P.has_noise != 0;
  const int act_kind = SPEC ? (int)kActF32 : P.act_kind;
  const R cell_size = kFast ? (R)P.f_cell : (R)P.cell_size, inv_cell = kFast ? (R)P.f_inv_cell : (R)0, half = kFast ? (R)P.f_half : (R)(P.cell_size / 2);
  const R a_std = kFast ? (R)P.f_std : (R)P.action_std, a_pow = kFast ? (R)P.f_pow : (R)P.action_power;
  const R max_y = kFast ? (R)P.f_max_y : (R)P.max_y, max_x = kFast ? (R)P.f_max_x : (R)P.max_x;
  const R thr = kFast ? (R)0 : (R)P.goal_threshold, thr2 = kFast ? (R)P.f_thr2 : (R)0;
  const R goal_y = kFast ? (R)P.f_goal_y : (R)P.goal_y, goal_x = kFast ? (R)P.f_goal_x : (R)P.goal_x;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.blob_bytes, &bar);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  // a warp owns 128 consecutive envs (one quad per lane)
  const int64_t wq = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t q = first + wq * kQuadStride + lane * kQuad;
  if (q >= last) return;
  const bool reset_all = P.mode == kModeReset;
  const uint32_t n = (uint32_t)P.n_actions;
  auto cell_of = [&](R v) -> int { return kFast ? (int)floor(v * inv_cell) : (int)floor(v / cell_size); };

  pdl_wait();
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, 1u);
  V2 pos[4], gpos[4], vel[4], push[4];
  int32_t ev[4] = {0, 0, 0, 0};
  uint32_t abytes = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    pos[k] = RealTraits<R>::make(0, 0);
    vel[k] = RealTraits<R>::make(0, 0);
    push[k] = RealTraits<R>::make(0, 0);
    gpos[k] = RealTraits<R>::make(goal_y, goal_x);
  }
  if (!reset_all) {
    const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
    RealTraits<R>::load4(P.agent, q, pos);
    if (rgoal) RealTraits<R>::load4(P.goal, q, gpos);
    if (use_velocity) RealTraits<R>::load4(P.velocity, q, vel);
    if (act_kind == kActI8) abytes = ld_stream(reinterpret_cast<const uint32_t*>(reinterpret_cast<const int8_t*>(P.actions) + q));
    else load_actions4<R>(P.actions, act_kind, q, push);
  }

  stage_tables_wait(&bar);
  RoomsTables T;
  T.nb8 = smem + P.nb8_off;
  T.room = smem + P.room_off;
  T.sid = reinterpret_cast<const uint16_t*>(smem + P.sid_off);
  T.valid = reinterpret_cast<const uint16_t*>(smem + P.valid_off);
  T.thr64 = reinterpret_cast<const double*>(smem + P.thr64_off);
  T.rows = reinterpret_cast<const uint64_t*>(smem + P.rows_off);
  const int8_t* grid = reinterpret_cast<const int8_t*>(smem + P.grid_off);
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);
  const int gn = P.grid_n;
  ObsCtx OC;
  OC.w = P.w; OC.n_rooms = P.n_rooms; OC.n_valid = P.n_valid; OC.hansen_n = P.hansen_n; OC.gn = gn; OC.div_w = P.div_w;
  OC.fixed_goal = 0; OC.gy = 0; OC.gx = 0;   // goal cell always derived from the goal position
  uint8_t* stage = smem + P.stage_off + warp * (uint32_t)(kQuadStride * gn * gn);

  float rv[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t tw = 0, trw = 0, again = reset_all ? 0xFu : 0u;
  uint32_t o32[4] = {0, 0, 0, 0}, o32b[4] = {0, 0, 0, 0};

  if (!reset_all) {
    uint4 slipq = make_uint4(0, 0, 0, 0);
    if (!REPLAY && act_kind == kActI8) {  // one Philox block feeds the slip draws of the quad
      const uint64_t gq = (uint64_t)(P.env_offset + q) >> 2;
      slipq = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, gq, 3u);
    }
    const uint32_t slipv[4] = {slipq.x, slipq.y, slipq.z, slipq.w};
    // float32 fast mode: 3 Philox blocks per quad = 3 words per env: 24 + 24 bits for the action-noise pair, 16 + 16
    // bits for the in-cell jitter pair (which is clipped at +-1 sigma, so its tail resolution is irrelevant)
    uint32_t qw[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (kFast && !REPLAY) {
      const uint64_t gq = (uint64_t)(P.env_offset + q) >> 2;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint4 b = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, gq, (uint32_t)(4 + j));   // per-step noise: Philox4x32-7
        qw[4 * j] = b.x; qw[4 * j + 1] = b.y; qw[4 * j + 2] = b.z; qw[4 * j + 3] = b.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t env = q + k;
      ev[k] += 1;
      // ---- noisy action (crooms.py:175-178 / :188-196) ----
      uint4 r0 = make_uint4(0, 0, 0, 0);
      if constexpr (kFast && !REPLAY) r0 = make_uint4(qw[3 * k], qw[3 * k + 1], qw[3 * k + 2] & 0xFFFF0000u, qw[3 * k + 2] << 16);
      else if (!REPLAY) r0 = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 0u);
      if (act_kind == kActI8) {
        uint32_t a = (abytes >> (8 * k)) & 0xFFu;
        a = a < n ? a : n - 1;
        uint32_t d8;
        if (REPLAY) {
          const double u = P.rp_u[env];
          const double* row = T.thr64 + a * n;
          uint32_t a2 = 0;
          for (uint32_t i = 0; i < n; ++i) a2 += row[i] < u ? 1u : 0u;
          a2 = a2 < n ? a2 : n - 1;
          d8 = n == 4 ? a2 * 2 : a2;
        } else {  // Walker alias column (see gpt_rooms_kernel.cuh), entries in ordinal-direction units
          const uint32_t u = slipv[k];
          const uint2 e = alias[a * 8 + (u >> (32u - P.log2n))];
          d8 = ((u << P.log2n) < e.x) ? (e.y & 0xFFu) : (e.y >> 8);
        }
        push[k] = RealTraits<R>::make((R)dir_dy(d8), (R)dir_dx(d8));
      }
      if (has_noise) {
        V2 z;
        if constexpr (REPLAY) {
          const double2 zz = P.rp_noise[env];  // already scaled by action_std (numpy normal(scale=std))
          z = RealTraits<R>::make((R)zz.x, (R)zz.y);
        } else if constexpr (kFast) {
          const float2 zz = normal_pair_f(r0.x, r0.y);
          z = RealTraits<R>::make(zz.x * a_std, zz.y * a_std);
        } else {
          const double2 zz = normal_pair(r0);
          z = RealTraits<R>::make((R)zz.x * a_std, (R)zz.y * a_std);
        }
        push[k].x = push[k].x + z.x;
        push[k].y = push[k].y + z.y;
      }
      push[k].x = push[k].x * a_pow;
      push[k].y = push[k].y * a_pow;
      // ---- _apply_action (crooms.py:300-331) ----
      V2 target;
      if (use_velocity) {
        vel[k].x = clipd<R>(vel[k].x + push[k].x, (R)-5.0, (R)5.0);
        vel[k].y = clipd<R>(vel[k].y + push[k].y, (R)-5.0, (R)5.0);
        target = RealTraits<R>::make(pos[k].x + vel[k].x, pos[k].y + vel[k].y);
      } else {
        target = RealTraits<R>::make(pos[k].x + push[k].x, pos[k].y + push[k].y);
      }
      target.x = clipd<R>(target.x, (R)0.0, max_y);
      target.y = clipd<R>(target.y, (R)0.0, max_x);
      const bool blocked = grid[cell_of(target.x) * P.w + cell_of(target.y)] < 0;
      // blocked: stay in the current cell at a jittered position, velocity zeroed (:317-330)
      if constexpr (kFast && !REPLAY) {  // branch-free: the second normal pair comes from the same Philox block
        const float2 zz = normal_pair_f(r0.z, r0.w);
        const R cy = floor(pos[k].x * inv_cell) * cell_size + half;
        const R cx = floor(pos[k].y * inv_cell) * cell_size + half;
        // upper bound: the reference's `centre + half - 1e-8` is below float32 resolution; (1 - 2^-23) * (centre + half)
        // is 1-2 ulp below the cell's far edge, so a clipped jitter never lands in the next (possibly wall) cell
        const R jy = clipd<R>(cy + zz.x * (R)0.5, cy - half, (cy + half) * (R)0.99999988f);
        const R jx = clipd<R>(cx + zz.y * (R)0.5, cx - half, (cx + half) * (R)0.99999988f);
        pos[k].x = blocked ? jy : target.x;
        pos[k].y = blocked ? jx : target.y;
        vel[k].x = blocked ? (R)0 : vel[k].x;
        vel[k].y = blocked ? (R)0 : vel[k].y;
      } else {
        if (!blocked) {
          pos[k] = target;
        } else {
          const R cy = (R)cell_of(pos[k].x) * cell_size + half;
          const R cx = (R)cell_of(pos[k].y) * cell_size + half;
          V2 z;
          if constexpr (REPLAY) {
            const double2 zz = P.rp_resample[env];   // normal(scale=0.5)
            z = RealTraits<R>::make((R)zz.x, (R)zz.y);
          } else {
            const double2 zz = normal_pair(rnd_block<DEVCTR>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 2u));
            z = RealTraits<R>::make((R)zz.x * (R)0.5, (R)zz.y * (R)0.5);
          }
          const R hy = kFast ? (cy + half) * (R)0.99999988f : cy + half - (R)1e-8;   // float32: see above
          const R hx = kFast ? (cx + half) * (R)0.99999988f : cx + half - (R)1e-8;
          pos[k].x = clipd<R>(cy + z.x, cy - half, hy);
          pos[k].y = clipd<R>(cx + z.y, cx - half, hx);
          vel[k] = RealTraits<R>::make(0, 0);
        }
      }
      // ---- reward / done (:290-296) ----
      const R dy = pos[k].x - gpos[k].x, dx = pos[k].y - gpos[k].y;
      const bool at_goal = kFast ? (dy * dy + dx * dx <= thr2) : (sqrt(dy * dy + dx * dx) <= thr);
      rv[k] = at_goal ? P.r_goal : (blocked ? P.r_wall : P.r_step);
      const bool trunc = ev[k] > P.time_limit;
      tw |= (at_goal ? 1u : 0u) << (8 * k);
      trw |= (trunc ? 1u : 0u) << (8 * k);
      again |= ((at_goal | trunc) ? 1u : 0u) << k;
    }
  }
  if (again) {  // rare: goal and agent respawn at unit-cell centres, velocity zero
#pragma unroll 1
    for (uint32_t m = again; m; m &= m - 1) {
      const int k = __ffs(m) - 1;
      const uint32_t fresh = crooms_respawn<REPLAY, DEVCTR>(P, T.valid, q + k, ctr_dev);
      const uint32_t ac = fresh & 0xFFFFu, gc = fresh >> 16;
      const uint32_t ay = fdiv(ac, P.div_w), gy = fdiv(gc, P.div_w);
      const V2 np = RealTraits<R>::make((R)ay + (R)0.5, (R)(ac - ay * P.w) + (R)0.5);
      const V2 ng = RealTraits<R>::make((R)gy + (R)0.5, (R)(gc - gy * P.w) + (R)0.5);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i == k) {
          ev[i] = 0;
          pos[i] = np;
          if (rgoal) gpos[i] = ng;
          vel[i] = RealTraits<R>::make(0, 0);
        }
      }
    }
  }
  RealTraits<R>::store4(P.agent, q, pos);
  if (rgoal) RealTraits<R>::store4(P.goal, q, gpos);
  if (use_velocity) RealTraits<R>::store4(P.velocity, q, vel);
  st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
  if (!reset_all) {
    st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
  }
  // ---- observation ----
  if constexpr (OBS == GPT_OBS_VEC_MDP) {
    RealTraits<R>::store4(P.obs, q, pos);
  } else if constexpr (OBS == GPT_OBS_VEC_MDP_GOAL) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const V2 two[4] = {pos[k], gpos[k], pos[k], gpos[k]};
      V2* o = reinterpret_cast<V2*>(P.obs) + 2 * (q + k);
      __stcs(o, two[0]);
      __stcs(o + 1, two[1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t cell = (uint32_t)(cell_of(pos[k].x) * P.w + cell_of(pos[k].y));
      const uint32_t gcell = (uint32_t)(cell_of(gpos[k].x) * P.w + cell_of(gpos[k].y));
      cell_obs<OBS, 0>(T, OC, cell, gcell, stage + (uint32_t)(lane * kQuad + k) * (uint32_t)(gn * gn), o32[k], o32b[k]);
    }
    store_obs<OBS>(P.obs, q, first + wq * kQuadStride, P.hansen_n, gn, lane, stage, o32, o32b);
  }
}

// ------------------------------------------------------------------------------------------
// Tag
// ------------------------------------------------------------------------------------------
struct TagParams {
  void* agent;    // real2 [cap]
  void* target;   // real2 [cap]
  int32_t* elapsed;
  const void* actions;
  void* obs;      // real2 [cap]
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  const double2* rp_noise;
  const int8_t* rp_choice;
  const double2* rp_spawn_agent;
  const double2* rp_spawn_target;
  int64_t env_offset;
  int32_t first_tile, n_tiles, mode, time_limit, act_kind;
  double action_std, action_power;
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};


// reset_model (ant_tag.py:88-103): agent uniform in the cage, target redrawn while within the minimum distance.
// Rare, out of line; returns (agent, target) through registers.
template <typename R, bool REPLAY, bool DEVCTR = false>
__device__ GPT_CONT_RESPAWN_ATTR void tag_respawn(const TagParams& P, int64_t env, typename RealTraits<R>::V2* pos_out, typename RealTraits<R>::V2* tgt_out, uint64_t ctr_dev = 0) {
  using V2 = typename RealTraits<R>::V2;
  constexpr R kMinSpawn = (R)5.0;
  V2 pos, tgt;
  if (REPLAY) {
    const double2 sa = P.rp_spawn_agent[env], st = P.rp_spawn_target[env];
    pos = RealTraits<R>::make((R)sa.x, (R)sa.y);
    tgt = RealTraits<R>::make((R)st.x, (R)st.y);
  } else {
    const uint4 r = rnd_block<DEVCTR>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 1u);
    const double s = 2.0 * 4.5 / 4294967296.0;
    pos = RealTraits<R>::make((R)((double)r.x * s - 4.5), (R)((double)r.y * s - 4.5));
    tgt = pos;
    uint32_t attempt = 0;
    do {
      const uint4 t = rnd_block<DEVCTR>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 16u + (attempt >> 1));
      tgt = (attempt & 1u) ? RealTraits<R>::make((R)((double)t.z * s - 4.5), (R)((double)t.w * s - 4.5))
                           : RealTraits<R>::make((R)((double)t.x * s - 4.5), (R)((double)t.y * s - 4.5));
      ++attempt;
      const R dx = pos.x - tgt.x, dy = pos.y - tgt.y;
      if (sqrt(dx * dx + dy * dy) > kMinSpawn) break;
    } while (attempt < 400u);
  }
  *pos_out = pos;
  *tgt_out = tgt;
}

#ifndef GPT_TAG_MINB
#define GPT_TAG_MINB 1
#endif
template <typename R, bool REPLAY, bool DEVCTR = false>
__global__ void __launch_bounds__(128, GPT_TAG_MINB) tag_step_kernel(const __grid_constant__ TagParams P) {
  using V2 = typename RealTraits<R>::V2;
  constexpr bool kFast = sizeof(R) == 4;
  constexpr R kCage = (R)4.5, kVisible = (R)3.0, kTagRadius = (R)1.5, kTargetStep = (R)0.5, kArena = (R)5.0;
  const R a_std = (R)P.action_std, a_pow = (R)P.action_power;
  pdl_launch_dependents();
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t q = first + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kQuad;
  if (q >= last) return;
  const bool reset_all = P.mode == kModeReset;
  pdl_wait();
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, 1u);
  float rv[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t tw = 0, trw = 0, again = reset_all ? 0xFu : 0u;
  int32_t ev[4] = {0, 0, 0, 0};
  V2 pos[4], tgt[4], push[4], obs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) pos[k] = tgt[k] = push[k] = RealTraits<R>::make(0, 0);
  if (!reset_all) {
    const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
    RealTraits<R>::load4(P.agent, q, pos);
    RealTraits<R>::load4(P.target, q, tgt);
    load_actions4<R>(P.actions, P.act_kind, q, push);
    // float32 fast mode: 2 Philox blocks per quad = 2 words per env: 2 x 24 bits for the noise pair, the 2 low bits
    // of the first word (not used by the noise) pick the target's move
    uint32_t qw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if constexpr (kFast && !REPLAY) {
      const uint64_t gq = (uint64_t)(P.env_offset + q) >> 2;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint4 b = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, gq, (uint32_t)(4 + j));   // per-step noise: Philox4x32-7
        qw[4 * j] = b.x; qw[4 * j + 1] = b.y; qw[4 * j + 2] = b.z; qw[4 * j + 3] = b.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t env = q + k;
      ev[k] += 1;
      uint4 r0 = make_uint4(0, 0, 0, 0);
      if constexpr (kFast && !REPLAY) r0 = make_uint4(qw[2 * k], qw[2 * k + 1], qw[2 * k] << 30, 0u);
      else if (!REPLAY) r0 = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 0u);
      V2 z;
      uint32_t choice;
      if constexpr (REPLAY) {
        const double2 zz = P.rp_noise[env];
        z = RealTraits<R>::make((R)zz.x, (R)zz.y);
        choice = (uint32_t)P.rp_choice[env];
      } else if constexpr (kFast) {  // one Philox block per env: 2 x 24 bits for the noise, 2 bits for the target's move
        const float2 zz = normal_pair_f(r0.x, r0.y);
        z = RealTraits<R>::make(zz.x * a_std, zz.y * a_std);
        choice = r0.z >> 30;
      } else {
        const double2 zz = normal_pair(r0);
        z = RealTraits<R>::make((R)zz.x * a_std, (R)zz.y * a_std);
        choice = rnd_block<DEVCTR, kStepRounds>(P.rng, ctr_dev, (uint64_t)(P.env_offset + env), 3u).x >> 30;
      }
      push[k].x = (push[k].x + z.x) * a_pow;
      push[k].y = (push[k].y + z.y) * a_pow;
      pos[k].x = clipd<R>(pos[k].x + push[k].x, -kArena, kArena);
      pos[k].y = clipd<R>(pos[k].y + push[k].y, -kArena, kArena);
      // target moves relative to the agent's NEW position (ant_tag.py:105-123, :139-141)
      R vx = pos[k].x - tgt[k].x, vy = pos[k].y - tgt[k].y;
      if constexpr (kFast) {
        const R inv = fast_rsqrt(vx * vx + vy * vy);
        vx = vx * inv;
        vy = vy * inv;
      } else {
        const R nrm = sqrt(vx * vx + vy * vy);
        vx = vx / nrm;
        vy = vy / nrm;
      }
      R mx = 0, my = 0;
      if (choice == 0) { mx = -vx; my = -vy; }
      else if (choice == 1) { mx = vy; my = -vx; }
      else if (choice == 2) { mx = -vy; my = vx; }
      const R nx = mx * kTargetStep + tgt[k].x, ny = my * kTargetStep + tgt[k].y;
      if (!(fabs(nx) > kCage || fabs(ny) > kCage)) tgt[k] = RealTraits<R>::make(nx, ny);
      const R dx = pos[k].x - tgt[k].x, dy = pos[k].y - tgt[k].y;
      const bool tagged = kFast ? (dx * dx + dy * dy <= kTagRadius * kTagRadius)
                                : (sqrt(dx * dx + dy * dy) <= kTagRadius);     // (:147-150)
      rv[k] = tagged ? 1.f : 0.f;
      const bool trunc = ev[k] >= P.time_limit;                       // gymnasium TimeLimit (envs/__init__.py:15-19)
      tw |= (tagged ? 1u : 0u) << (8 * k);
      trw |= (trunc ? 1u : 0u) << (8 * k);
      again |= ((tagged | trunc) ? 1u : 0u) << k;
    }
  }
  if (again) {
#pragma unroll 1
    for (uint32_t m = again; m; m &= m - 1) {
      const int k = __ffs(m) - 1;
      V2 np, nt;
      tag_respawn<R, REPLAY, DEVCTR>(P, q + k, &np, &nt, ctr_dev);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i == k) {
          ev[i] = 0;
          pos[i] = np;
          tgt[i] = nt;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const R dx = pos[k].x - tgt[k].x, dy = pos[k].y - tgt[k].y;
    const bool visible = kFast ? (dx * dx + dy * dy < kVisible * kVisible) : (sqrt(dx * dx + dy * dy) < kVisible);   // (:153, :83-85)
    obs[k] = visible ? tgt[k] : RealTraits<R>::make(0, 0);
  }
  RealTraits<R>::store4(P.agent, q, pos);
  RealTraits<R>::store4(P.target, q, tgt);
  RealTraits<R>::store4(P.obs, q, obs);
  st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
  if (!reset_all) {
    st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
  }
}

// Instantiations: the float64 kernels (bit-exact parity with numpy) live in gpt_crooms.cu, compiled with -fmad=false;
// the float32 fast-mode kernels live in gpt_crooms_f32.cu, compiled with FMA contraction enabled.
// `spec` = the launch has the default configuration (see SPEC above); only the float32 Philox kernels carry that
// instantiation, everything else ignores the hint.
template <typename R, int OBS>
static void* crooms_pick_rr(bool replay, bool devctr, bool spec) {
  using K = void (*)(const CRoomsParams);
  if constexpr (sizeof(R) == 4) {
    if (spec && !replay) return devctr ? (void*)(K)crooms_step_kernel<R, OBS, false, true, true> : (void*)(K)crooms_step_kernel<R, OBS, false, false, true>;
  }
  if (devctr) return replay ? nullptr : (void*)(K)crooms_step_kernel<R, OBS, false, true>;   // graph mode (Philox)
  return replay ? (void*)(K)crooms_step_kernel<R, OBS, true> : (void*)(K)crooms_step_kernel<R, OBS, false>;
}
template <typename R>
static void* crooms_pick_obs(int obs, bool replay, bool devctr = false, bool spec = false) {
  switch (obs) {
    case GPT_OBS_ROOM: return crooms_pick_rr<R, GPT_OBS_ROOM>(replay, devctr, spec);
    case GPT_OBS_ROOM_GOAL: return crooms_pick_rr<R, GPT_OBS_ROOM_GOAL>(replay, devctr, spec);
    case GPT_OBS_MDP: return crooms_pick_rr<R, GPT_OBS_MDP>(replay, devctr, spec);
    case GPT_OBS_MDP_GOAL: return crooms_pick_rr<R, GPT_OBS_MDP_GOAL>(replay, devctr, spec);
    case GPT_OBS_VEC_MDP: return crooms_pick_rr<R, GPT_OBS_VEC_MDP>(replay, devctr, spec);
    case GPT_OBS_VEC_MDP_GOAL: return crooms_pick_rr<R, GPT_OBS_VEC_MDP_GOAL>(replay, devctr, spec);
    case GPT_OBS_HANSEN: return crooms_pick_rr<R, GPT_OBS_HANSEN>(replay, devctr, spec);
    case GPT_OBS_VEC_HANSEN: return crooms_pick_rr<R, GPT_OBS_VEC_HANSEN>(replay, devctr, spec);
    case GPT_OBS_VEC_HANSEN_GOAL: return crooms_pick_rr<R, GPT_OBS_VEC_HANSEN_GOAL>(replay, devctr, spec);
    case GPT_OBS_GRID: return crooms_pick_rr<R, GPT_OBS_GRID>(replay, devctr, spec);
  }
  return nullptr;
}
template <typename R>
static void* tag_pick_rr(bool replay, bool devctr = false) {
  using K = void (*)(const TagParams);
  if (devctr) return replay ? nullptr : (void*)(K)tag_step_kernel<R, false, true>;   // graph mode (Philox)
  return replay ? (void*)(K)tag_step_kernel<R, true> : (void*)(K)tag_step_kernel<R, false>;
}
void* crooms_pick_f32(int obs, bool replay, bool devctr, bool spec);   // gpt_crooms_f32.cu
void* tag_pick_f32(bool replay, bool devctr);

}  // namespace gpt
