// gpt_crooms.cu — fused continuous-position ROOMS step and point-mass Tag step for sm_100a.
//
// CROOMS = CRoomsEnv.step (reference gym_po/envs/rooms/crooms.py:276-298 with _apply_action :300-331,
// _out_of_bounds :333-338, the sample_action closures :175-198, _reset_some :268-274 and
// grid_to_coord / coord_to_grid, rooms/utils.py:7-20).  The reference computes in float64 with numpy;
// this kernel computes in float64 with the same operation order and NO fused multiply-add (the file is
// compiled with -fmad=false), so that on replayed random draws positions, rewards and flags are
// bit-identical.  TAG = the pursuit rules of AntTagEnv (ant_tag.py:105-123, :144-153) on a point mass
// that moves with the CROOMS motion model (see oracle/tag.py and DESIGN.md).
//
// HBM layout: agent double2 [cap] | goal double2 [cap] (random-goal envs) | velocity double2 [cap]
// (use_velocity) | elapsed int32 | action (int8 | float32x2 | float64x2)  ->  the same state arrays,
// obs, reward float32, terminated uint8, truncated uint8.  One thread handles 4 consecutive envs.
#include "gpt_rooms_kernel.cuh"

namespace gpt {

enum : int { kActI8 = 0, kActF32 = 1, kActF64 = 2 };

struct CRoomsParams {
  double2* agent;
  double2* goal;
  double2* velocity;
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
  uint32_t blob_bytes, nb8_off, room_off, sid_off, valid_off, thr32_off, thr64_off, rows_off, stage_off, grid_off;
  int64_t env_offset;
  int32_t first_tile, n_tiles, mode;
  int32_t h, w, n_actions, n_valid, n_rooms, time_limit, hansen_n, grid_n;
  int32_t act_kind, rgoal, use_velocity, has_noise;
  FastDiv div_w;
  double cell_size, action_std, action_power, goal_threshold, max_y, max_x, goal_y, goal_x;
  float r_step, r_wall, r_goal;
  RngKey rng;
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

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

template <int OBS, bool REPLAY>
__global__ void __launch_bounds__(128) crooms_step_kernel(const __grid_constant__ CRoomsParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
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

  stage_tables_wait(&bar);
  RoomsTables T;
  T.nb8 = smem + P.nb8_off;
  T.room = smem + P.room_off;
  T.sid = reinterpret_cast<const uint16_t*>(smem + P.sid_off);
  T.valid = reinterpret_cast<const uint16_t*>(smem + P.valid_off);
  T.thr32 = reinterpret_cast<const uint32_t*>(smem + P.thr32_off);
  T.thr64 = reinterpret_cast<const double*>(smem + P.thr64_off);
  T.rows = reinterpret_cast<const uint64_t*>(smem + P.rows_off);
  const int8_t* grid = reinterpret_cast<const int8_t*>(smem + P.grid_off);
  const int gn = P.grid_n;
  ObsCtx OC;
  OC.w = P.w; OC.n_rooms = P.n_rooms; OC.n_valid = P.n_valid; OC.hansen_n = P.hansen_n; OC.gn = gn; OC.div_w = P.div_w;
  OC.fixed_goal = 0; OC.gy = 0; OC.gx = 0;   // goal cell always derived from the goal position
  uint8_t* stage = smem + P.stage_off + warp * (uint32_t)(kQuadStride * gn * gn);

  float rv[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t tw = 0, trw = 0;
  uint32_t o32[4] = {0, 0, 0, 0}, o32b[4] = {0, 0, 0, 0};
  int32_t ev[4] = {0, 0, 0, 0};
  if (!reset_all) {
    const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
  }

#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t env = q + k;
    double2 pos = make_double2(0, 0), vel = make_double2(0, 0);
    double2 gpos = make_double2(P.goal_y, P.goal_x);
    bool again = reset_all;
    if (!reset_all) {
      pos = P.agent[env];
      if (P.rgoal) gpos = P.goal[env];
      if (P.use_velocity) vel = P.velocity[env];
      ev[k] += 1;
      // ---- noisy action (crooms.py:175-178 / :188-196) ----
      double2 push;
      uint4 r0 = make_uint4(0, 0, 0, 0);
      if (!REPLAY) r0 = env_random(P.rng, (uint64_t)(P.env_offset + env), 0u);
      if (P.act_kind == kActI8) {
        uint32_t a = (uint32_t)(uint8_t) reinterpret_cast<const int8_t*>(P.actions)[env];
        a = a < n ? a : n - 1;
        uint32_t a2 = 0;
        if (REPLAY) {
          const double u = P.rp_u[env];
          const double* row = T.thr64 + a * n;
          for (uint32_t i = 0; i < n; ++i) a2 += row[i] < u ? 1u : 0u;
          a2 = a2 < n ? a2 : n - 1;
        } else {
          const uint4 rs = env_random(P.rng, (uint64_t)(P.env_offset + env), 3u);
          const uint32_t* row = T.thr32 + a * 8;   // 8-wide rows in ordinal-direction units (gpt_rooms.cu)
          a2 = row[3] < rs.x ? 4u : 0u;
          a2 += row[a2 + 1] < rs.x ? 2u : 0u;
          a2 += row[a2] < rs.x ? 1u : 0u;
        }
        const uint32_t d8 = (REPLAY && n == 4) ? a2 * 2 : a2;
        push = make_double2((double)dir_dy(d8), (double)dir_dx(d8));
      } else if (P.act_kind == kActF32) {
        const float2 a = reinterpret_cast<const float2*>(P.actions)[env];
        push = make_double2((double)a.x, (double)a.y);
      } else {
        push = reinterpret_cast<const double2*>(P.actions)[env];
      }
      if (P.has_noise) {
        double2 z;
        if (REPLAY) {
          z = P.rp_noise[env];                 // already scaled by action_std (numpy normal(scale=std))
        } else {
          z = normal_pair(r0);
          z.x *= P.action_std;
          z.y *= P.action_std;
        }
        push.x = push.x + z.x;
        push.y = push.y + z.y;
      }
      push.x = push.x * P.action_power;
      push.y = push.y * P.action_power;
      // ---- _apply_action (crooms.py:300-331) ----
      double2 target;
      if (P.use_velocity) {
        vel.x = clipd(vel.x + push.x, -5.0, 5.0);
        vel.y = clipd(vel.y + push.y, -5.0, 5.0);
        target = make_double2(pos.x + vel.x, pos.y + vel.y);
      } else {
        target = make_double2(pos.x + push.x, pos.y + push.y);
      }
      target.x = clipd(target.x, 0.0, P.max_y);
      target.y = clipd(target.y, 0.0, P.max_x);
      const int ty = (int)floor(target.x / P.cell_size), tx = (int)floor(target.y / P.cell_size);
      const bool blocked = grid[ty * P.w + tx] < 0;
      if (!blocked) {
        pos = target;
      } else {  // stay in the current cell at a jittered position, velocity zeroed (:317-330)
        const double half = P.cell_size / 2;
        const double cy = floor(pos.x / P.cell_size) * P.cell_size + half;
        const double cx = floor(pos.y / P.cell_size) * P.cell_size + half;
        double2 z;
        if (REPLAY) {
          z = P.rp_resample[env];              // normal(scale=0.5)
        } else {
          z = normal_pair(env_random(P.rng, (uint64_t)(P.env_offset + env), 2u));
          z.x *= 0.5;
          z.y *= 0.5;
        }
        pos.x = clipd(cy + z.x, cy - half, cy + half - 1e-8);
        pos.y = clipd(cx + z.y, cx - half, cx + half - 1e-8);
        vel = make_double2(0, 0);
      }
      // ---- reward / done (:290-296) ----
      const double dy = pos.x - gpos.x, dx = pos.y - gpos.y;
      const bool at_goal = sqrt(dy * dy + dx * dx) <= P.goal_threshold;
      rv[k] = at_goal ? P.r_goal : (blocked ? P.r_wall : P.r_step);
      const bool trunc = ev[k] > P.time_limit;
      tw |= (at_goal ? 1u : 0u) << (8 * k);
      trw |= (trunc ? 1u : 0u) << (8 * k);
      again = at_goal | trunc;
    }
    if (again) {  // _reset_some (:268-274): goal, agent at unit-cell centres, velocity zero
      ev[k] = 0;
      uint32_t ac, gc = 0;
      if (REPLAY) {
        if (P.rgoal) gc = (uint32_t)P.rp_reset_goal[env];
        ac = (uint32_t)P.rp_reset_agent[env];
      } else {
        const uint4 r = env_random(P.rng, (uint64_t)(P.env_offset + env), 1u);
        if (P.rgoal) gc = T.valid[bounded(r.y, (uint32_t)P.n_valid)];
        ac = T.valid[bounded(r.x, (uint32_t)P.n_valid)];
      }
      if (P.rgoal) {
        const uint32_t y = fdiv(gc, P.div_w);
        gpos = make_double2((double)y + 0.5, (double)(gc - y * P.w) + 0.5);
      }
      const uint32_t y = fdiv(ac, P.div_w);
      pos = make_double2((double)y + 0.5, (double)(ac - y * P.w) + 0.5);
      vel = make_double2(0, 0);
    }
    P.agent[env] = pos;
    if (P.rgoal) P.goal[env] = gpos;
    if (P.use_velocity) P.velocity[env] = vel;

    // ---- observation ----
    if constexpr (OBS == GPT_OBS_VEC_MDP) {
      reinterpret_cast<double2*>(P.obs)[env] = pos;
    } else if constexpr (OBS == GPT_OBS_VEC_MDP_GOAL) {
      reinterpret_cast<double2*>(P.obs)[2 * env] = pos;
      reinterpret_cast<double2*>(P.obs)[2 * env + 1] = gpos;
    } else {
      const uint32_t cell = (uint32_t)((int)floor(pos.x / P.cell_size) * P.w + (int)floor(pos.y / P.cell_size));
      const uint32_t gcell = (uint32_t)((int)floor(gpos.x / P.cell_size) * P.w + (int)floor(gpos.y / P.cell_size));
      cell_obs<OBS, 0>(T, OC, cell, gcell, stage + (uint32_t)(lane * kQuad + k) * (uint32_t)(gn * gn), o32[k], o32b[k]);
    }
  }
  st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
  if (!reset_all) {
    st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
  }
  if constexpr (OBS != GPT_OBS_VEC_MDP && OBS != GPT_OBS_VEC_MDP_GOAL)
    store_obs<OBS>(P.obs, q, first + wq * kQuadStride, P.hansen_n, gn, lane, stage, o32, o32b);
}

// ------------------------------------------------------------------------------------------
// Tag
// ------------------------------------------------------------------------------------------
struct TagParams {
  double2* agent;
  double2* target;
  int32_t* elapsed;
  const void* actions;
  double2* obs;
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
};

constexpr double kCage = 4.5, kVisible = 3.0, kTagRadius = 1.5, kMinSpawn = 5.0, kTargetStep = 0.5, kArena = 5.0;

template <bool REPLAY>
__global__ void __launch_bounds__(128) tag_step_kernel(const __grid_constant__ TagParams P) {
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t q = first + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kQuad;
  if (q >= last) return;
  const bool reset_all = P.mode == kModeReset;
  float rv[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t tw = 0, trw = 0;
  int32_t ev[4] = {0, 0, 0, 0};
  if (!reset_all) {
    const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t env = q + k;
    double2 pos = make_double2(0, 0), tgt = make_double2(0, 0);
    bool again = reset_all;
    if (!reset_all) {
      pos = P.agent[env];
      tgt = P.target[env];
      ev[k] += 1;
      uint4 r0 = make_uint4(0, 0, 0, 0);
      if (!REPLAY) r0 = env_random(P.rng, (uint64_t)(P.env_offset + env), 0u);
      double2 push;
      if (P.act_kind == kActF32) {
        const float2 a = reinterpret_cast<const float2*>(P.actions)[env];
        push = make_double2((double)a.x, (double)a.y);
      } else {
        push = reinterpret_cast<const double2*>(P.actions)[env];
      }
      double2 z;
      if (REPLAY) {
        z = P.rp_noise[env];
      } else {
        z = normal_pair(r0);
        z.x *= P.action_std;
        z.y *= P.action_std;
      }
      push.x = (push.x + z.x) * P.action_power;
      push.y = (push.y + z.y) * P.action_power;
      pos.x = clipd(pos.x + push.x, -kArena, kArena);
      pos.y = clipd(pos.y + push.y, -kArena, kArena);
      // target moves relative to the agent's NEW position (ant_tag.py:105-123, :139-141)
      uint32_t choice;
      if (REPLAY) choice = (uint32_t)P.rp_choice[env];
      else choice = env_random(P.rng, (uint64_t)(P.env_offset + env), 3u).x >> 30;
      double vx = pos.x - tgt.x, vy = pos.y - tgt.y;
      const double nrm = sqrt(vx * vx + vy * vy);
      vx = vx / nrm;
      vy = vy / nrm;
      double mx = 0.0, my = 0.0;
      if (choice == 0) { mx = -vx; my = -vy; }
      else if (choice == 1) { mx = vy; my = -vx; }
      else if (choice == 2) { mx = -vy; my = vx; }
      const double nx = mx * kTargetStep + tgt.x, ny = my * kTargetStep + tgt.y;
      if (!(fabs(nx) > kCage || fabs(ny) > kCage)) tgt = make_double2(nx, ny);
      const double dx = pos.x - tgt.x, dy = pos.y - tgt.y;
      const bool tagged = sqrt(dx * dx + dy * dy) <= kTagRadius;     // (:147-150)
      rv[k] = tagged ? 1.f : 0.f;
      const bool trunc = ev[k] >= P.time_limit;                       // gymnasium TimeLimit (envs/__init__.py:15-19)
      tw |= (tagged ? 1u : 0u) << (8 * k);
      trw |= (trunc ? 1u : 0u) << (8 * k);
      again = tagged | trunc;
    }
    if (again) {  // reset_model (ant_tag.py:88-103): target redrawn while within min distance
      ev[k] = 0;
      if (REPLAY) {
        pos = P.rp_spawn_agent[env];
        tgt = P.rp_spawn_target[env];
      } else {
        const uint4 r = env_random(P.rng, (uint64_t)(P.env_offset + env), 1u);
        const double s = 2.0 * kCage / 4294967296.0;
        pos = make_double2((double)r.x * s - kCage, (double)r.y * s - kCage);
        uint32_t attempt = 0;
        do {
          const uint4 t = env_random(P.rng, (uint64_t)(P.env_offset + env), 16u + (attempt >> 1));
          tgt = (attempt & 1u) ? make_double2((double)t.z * s - kCage, (double)t.w * s - kCage)
                               : make_double2((double)t.x * s - kCage, (double)t.y * s - kCage);
          ++attempt;
          const double dx = pos.x - tgt.x, dy = pos.y - tgt.y;
          if (sqrt(dx * dx + dy * dy) > kMinSpawn) break;
        } while (attempt < 400u);
      }
    }
    P.agent[env] = pos;
    P.target[env] = tgt;
    const double dx = pos.x - tgt.x, dy = pos.y - tgt.y;
    P.obs[env] = sqrt(dx * dx + dy * dy) < kVisible ? tgt : make_double2(0.0, 0.0);   // (:153, :83-85)
  }
  st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
  if (!reset_all) {
    st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int rooms_build_tables(gpt_env* env, const gpt_config* c, bool discrete_actions);

static int action_kind(const gpt_config* c) {
  if (c->rooms_n_actions > 0) return kActI8;
  return c->c_action_f64 == 1 ? kActF64 : kActF32;
}

int crooms_create(gpt_env* env, const gpt_config* c) {
  if (c->track_stats) return fail(GPT_E_ARG, "crooms: track_stats is implemented for the Taxi and ROOMS families only");
  if (!(c->c_cell_size > 0)) return fail(GPT_E_ARG, "crooms: cell_size must be > 0");
  const int kind = c->rooms_obs_kind;
  if ((kind == GPT_OBS_HANSEN || kind == GPT_OBS_VEC_HANSEN || kind == GPT_OBS_VEC_HANSEN_GOAL) && c->rooms_obs_n != 4 && c->rooms_obs_n != 8)
    return fail(GPT_E_ARG, "crooms: hansen obs_n must be 4 or 8");
  if (int rc = rooms_build_tables(env, c, c->rooms_n_actions > 0)) return rc;
  const bool rgoal = c->rooms_goal_y < 0;
  if (!rgoal && (c->rooms_goal_y >= c->rooms_h || c->rooms_goal_x >= c->rooms_w || c->rooms_goal_x < 0) &&
      (kind == GPT_OBS_ROOM_GOAL || kind == GPT_OBS_MDP_GOAL))
    return fail(GPT_E_ARG, "crooms: goal-indexed observations need the fixed goal inside the grid");
  const int ak = action_kind(c);
  add_array(env, "agent", GPT_ROLE_STATE, GPT_DT_F64, 2);
  if (rgoal) add_array(env, "goal", GPT_ROLE_STATE, GPT_DT_F64, 2);
  if (c->c_use_velocity) add_array(env, "velocity", GPT_ROLE_STATE, GPT_DT_F64, 2);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  switch (kind) {
    case GPT_OBS_VEC_MDP: add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_F64, 2); break;
    case GPT_OBS_VEC_MDP_GOAL: add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_F64, 4); break;
    case GPT_OBS_ROOM: case GPT_OBS_ROOM_GOAL: case GPT_OBS_MDP: case GPT_OBS_MDP_GOAL: case GPT_OBS_HANSEN:
      add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_I32, 1); break;
    case GPT_OBS_VEC_HANSEN: case GPT_OBS_VEC_HANSEN_GOAL: add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_U8, c->rooms_obs_n); break;
    case GPT_OBS_GRID: add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_U8, c->rooms_obs_n * c->rooms_obs_n); break;
    default: return fail(GPT_E_ARG, "crooms: unknown obs kind");
  }
  add_array(env, "reward", GPT_ROLE_OUTPUT, GPT_DT_F32, 1);
  add_array(env, "terminated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "truncated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "replay_u", GPT_ROLE_REPLAY, GPT_DT_F64, 1);
  add_array(env, "replay_noise", GPT_ROLE_REPLAY, GPT_DT_F64, 2);
  add_array(env, "replay_resample", GPT_ROLE_REPLAY, GPT_DT_F64, 2);
  add_array(env, "replay_reset_agent", GPT_ROLE_REPLAY, GPT_DT_I32, 1);
  add_array(env, "replay_reset_goal", GPT_ROLE_REPLAY, GPT_DT_I32, 1);
  env->action_dtype = ak == kActI8 ? GPT_DT_I8 : (ak == kActF32 ? GPT_DT_F32 : GPT_DT_F64);
  env->action_cols = ak == kActI8 ? 1 : 2;
  add_array(env, "actions", GPT_ROLE_ACTION, env->action_dtype, env->action_cols);
  return GPT_OK;
}

template <int OBS>
static void* pick_c(bool replay) {
  using K = void (*)(const CRoomsParams);
  return replay ? (void*)(K)crooms_step_kernel<OBS, true> : (void*)(K)crooms_step_kernel<OBS, false>;
}

int crooms_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool rgoal = c.rooms_goal_y < 0;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  CRoomsParams P{};
  P.agent = (double2*)env->ptr("agent");
  P.goal = rgoal ? (double2*)env->ptr("goal") : nullptr;
  P.velocity = c.c_use_velocity ? (double2*)env->ptr("velocity") : nullptr;
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = a.actions;
  P.obs = env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.agent || (rgoal && !P.goal) || (c.c_use_velocity && !P.velocity) || !P.elapsed || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "crooms: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "crooms: actions is NULL");
  const int oi = env->find("obs");
  const size_t obs_row = (size_t)env->arrays[oi].desc.cols * env->arrays[oi].desc.elem_size;
  P.obs = (uint8_t*)P.obs + a.out_row * obs_row;
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  if (replay) {
    P.rp_u = (const double*)env->ptr("replay_u");
    P.rp_noise = (const double2*)env->ptr("replay_noise");
    P.rp_resample = (const double2*)env->ptr("replay_resample");
    P.rp_reset_agent = (const int32_t*)env->ptr("replay_reset_agent");
    P.rp_reset_goal = (const int32_t*)env->ptr("replay_reset_goal");
    if (!P.rp_u || !P.rp_noise || !P.rp_resample || !P.rp_reset_agent || !P.rp_reset_goal)
      return fail(GPT_E_UNBOUND, "crooms: replay arrays must be bound in replay mode");
  }
  P.blob = env->d_blob;
  P.blob_bytes = env->blob_bytes;
  P.nb8_off = env->rl.nb8_off;
  P.room_off = env->rl.room_off;
  P.sid_off = env->rl.sid_off;
  P.valid_off = env->rl.valid_off;
  P.thr32_off = env->rl.thr32_off;
  P.thr64_off = env->rl.thr64_off;
  P.rows_off = env->rl.rows_off;
  P.grid_off = env->rl.grid_off;
  P.stage_off = (env->blob_bytes + 127u) & ~127u;
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.mode = a.mode;
  P.h = c.rooms_h;
  P.w = c.rooms_w;
  P.n_actions = c.rooms_n_actions;
  P.n_valid = env->rl.n_valid;
  P.n_rooms = env->rl.n_rooms;
  P.time_limit = c.time_limit;
  const bool grid = c.rooms_obs_kind == GPT_OBS_GRID;
  P.hansen_n = grid ? 0 : c.rooms_obs_n;
  P.grid_n = grid ? c.rooms_obs_n : 0;
  P.act_kind = action_kind(&c);
  P.rgoal = rgoal;
  P.use_velocity = c.c_use_velocity;
  P.has_noise = P.act_kind != kActI8 || c.c_action_std != 0.0;
  P.div_w = make_fastdiv((uint32_t)c.rooms_w);
  P.cell_size = c.c_cell_size;
  P.action_std = c.c_action_std;
  P.action_power = c.c_action_power;
  P.goal_threshold = c.c_goal_threshold;
  P.max_y = (double)(c.rooms_h - 1) - 1e-6;   // gridshape - 1 - 1e-6 (crooms.py:312-314)
  P.max_x = (double)(c.rooms_w - 1) - 1e-6;
  P.goal_y = rgoal ? 0.0 : (double)c.rooms_goal_y + 0.5;   // fixed goal at the unit-cell centre (:222-226)
  P.goal_x = rgoal ? 0.0 : (double)c.rooms_goal_x + 0.5;
  P.r_step = c.rooms_step_reward;
  P.r_wall = c.rooms_wall_reward;
  P.r_goal = c.rooms_goal_reward;
  P.rng = make_rng_key(env);

  const int threads = 128, warps = threads / 32;
  const int64_t wquads = (int64_t)a.n_tiles * (kTileEnvs / kQuadStride);
  const int nblocks = (int)((wquads + warps - 1) / warps);
  if (nblocks <= 0) return GPT_OK;
  size_t smem = env->blob_bytes;
  if (grid) smem = P.stage_off + (size_t)warps * kQuadStride * P.grid_n * P.grid_n;
  void* k = nullptr;
  switch (c.rooms_obs_kind) {
    case GPT_OBS_ROOM: k = pick_c<GPT_OBS_ROOM>(replay); break;
    case GPT_OBS_ROOM_GOAL: k = pick_c<GPT_OBS_ROOM_GOAL>(replay); break;
    case GPT_OBS_MDP: k = pick_c<GPT_OBS_MDP>(replay); break;
    case GPT_OBS_MDP_GOAL: k = pick_c<GPT_OBS_MDP_GOAL>(replay); break;
    case GPT_OBS_VEC_MDP: k = pick_c<GPT_OBS_VEC_MDP>(replay); break;
    case GPT_OBS_VEC_MDP_GOAL: k = pick_c<GPT_OBS_VEC_MDP_GOAL>(replay); break;
    case GPT_OBS_HANSEN: k = pick_c<GPT_OBS_HANSEN>(replay); break;
    case GPT_OBS_VEC_HANSEN: k = pick_c<GPT_OBS_VEC_HANSEN>(replay); break;
    case GPT_OBS_VEC_HANSEN_GOAL: k = pick_c<GPT_OBS_VEC_HANSEN_GOAL>(replay); break;
    case GPT_OBS_GRID: k = pick_c<GPT_OBS_GRID>(replay); break;
  }
  if (!k) return fail(GPT_E_ARG, "crooms: no kernel for this obs kind");
  if (smem > 40 * 1024) {  // static + dynamic shared memory above 48 KB needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(crooms)");
  }
  void* args[] = {(void*)&P};
  cudaError_t e = cudaLaunchKernel(k, dim3(nblocks), dim3(threads), args, smem, a.stream);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "crooms_step_kernel launch");
  return GPT_OK;
}

int tag_create(gpt_env* env, const gpt_config* c) {
  if (c->track_stats) return fail(GPT_E_ARG, "tag: track_stats is implemented for the Taxi and ROOMS families only");
  add_array(env, "agent", GPT_ROLE_STATE, GPT_DT_F64, 2);
  add_array(env, "target", GPT_ROLE_STATE, GPT_DT_F64, 2);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_F64, 2);
  add_array(env, "reward", GPT_ROLE_OUTPUT, GPT_DT_F32, 1);
  add_array(env, "terminated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "truncated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "replay_noise", GPT_ROLE_REPLAY, GPT_DT_F64, 2);
  add_array(env, "replay_choice", GPT_ROLE_REPLAY, GPT_DT_I8, 1);
  add_array(env, "replay_spawn_agent", GPT_ROLE_REPLAY, GPT_DT_F64, 2);
  add_array(env, "replay_spawn_target", GPT_ROLE_REPLAY, GPT_DT_F64, 2);
  env->action_dtype = c->c_action_f64 == 1 ? GPT_DT_F64 : GPT_DT_F32;
  env->action_cols = 2;
  add_array(env, "actions", GPT_ROLE_ACTION, env->action_dtype, 2);
  std::vector<uint8_t> blob(16, 0);
  return upload_blob(env, blob);
}

int tag_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  TagParams P{};
  P.agent = (double2*)env->ptr("agent");
  P.target = (double2*)env->ptr("target");
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = a.actions;
  P.obs = (double2*)env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.agent || !P.target || !P.elapsed || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "tag: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "tag: actions is NULL");
  P.obs += a.out_row;
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  if (replay) {
    P.rp_noise = (const double2*)env->ptr("replay_noise");
    P.rp_choice = (const int8_t*)env->ptr("replay_choice");
    P.rp_spawn_agent = (const double2*)env->ptr("replay_spawn_agent");
    P.rp_spawn_target = (const double2*)env->ptr("replay_spawn_target");
    if (!P.rp_noise || !P.rp_choice || !P.rp_spawn_agent || !P.rp_spawn_target)
      return fail(GPT_E_UNBOUND, "tag: replay arrays must be bound in replay mode");
  }
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.mode = a.mode;
  P.time_limit = c.time_limit;
  P.act_kind = c.c_action_f64 == 1 ? kActF64 : kActF32;
  P.action_std = c.c_action_std;
  P.action_power = c.c_action_power;
  P.rng = make_rng_key(env);
  const int threads = 128;
  const int64_t quads = (int64_t)a.n_tiles * (kTileEnvs / kQuad);
  const int nblocks = (int)((quads + threads - 1) / threads);
  if (nblocks <= 0) return GPT_OK;
  if (replay) tag_step_kernel<true><<<nblocks, threads, 0, a.stream>>>(P);
  else tag_step_kernel<false><<<nblocks, threads, 0, a.stream>>>(P);
  env->launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tag_step_kernel launch");
  return GPT_OK;
}

}  // namespace gpt
