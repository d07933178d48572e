// gpt_crooms.cu — host side and float64 instantiations of the fused continuous-position ROOMS step and point-mass Tag step for sm_100a.
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
#include "gpt_crooms_kernel.cuh"

namespace gpt {

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
  const int rdt = c->c_state_f32 ? GPT_DT_F32 : GPT_DT_F64;
  add_array(env, "agent", GPT_ROLE_STATE, rdt, 2);
  if (rgoal) add_array(env, "goal", GPT_ROLE_STATE, rdt, 2);
  if (c->c_use_velocity) add_array(env, "velocity", GPT_ROLE_STATE, rdt, 2);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  switch (kind) {
    case GPT_OBS_VEC_MDP: add_array(env, "obs", GPT_ROLE_OUTPUT, rdt, 2); break;
    case GPT_OBS_VEC_MDP_GOAL: add_array(env, "obs", GPT_ROLE_OUTPUT, rdt, 4); break;
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

int crooms_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool rgoal = c.rooms_goal_y < 0;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  CRoomsParams P{};
  P.agent = env->ptr("agent");
  P.goal = rgoal ? env->ptr("goal") : nullptr;
  P.velocity = c.c_use_velocity ? env->ptr("velocity") : nullptr;
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
  P.thr64_off = env->rl.thr64_off;
  P.rows_off = env->rl.rows_off;
  P.grid_off = env->rl.grid_off;
  P.alias_off = env->rl.alias_off;
  P.log2n = c.rooms_n_actions == 4 ? 2u : 3u;
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
  P.f_cell = (float)P.cell_size; P.f_inv_cell = (float)(1.0 / P.cell_size); P.f_half = (float)(P.cell_size / 2);
  P.f_std = (float)P.action_std; P.f_pow = (float)P.action_power; P.f_max_y = (float)P.max_y; P.f_max_x = (float)P.max_x;
  P.f_thr2 = (float)(P.goal_threshold * P.goal_threshold); P.f_goal_y = (float)P.goal_y; P.f_goal_x = (float)P.goal_x;
  P.rng = make_rng_key(env);

  const int threads = 128, warps = threads / 32;
  const int64_t wquads = (int64_t)a.n_tiles * (kTileEnvs / kQuadStride);
  const int nblocks = (int)((wquads + warps - 1) / warps);
  if (nblocks <= 0) return GPT_OK;
  size_t smem = env->blob_bytes;
  if (grid) smem = P.stage_off + (size_t)warps * kQuadStride * P.grid_n * P.grid_n;
  const bool devctr = env->graph_mode && !replay;   // graph mode: step counter in device memory
  P.ctr_ptr = env->d_counter;
  // the constructor's default configuration has its own instantiation (flags folded at compile time)
  const bool spec = P.act_kind == kActF32 && P.has_noise && !P.use_velocity && !P.rgoal;
  void* k = c.c_state_f32 ? crooms_pick_f32(c.rooms_obs_kind, replay, devctr, spec) : crooms_pick_obs<double>(c.rooms_obs_kind, replay, devctr);
  if (!k) return fail(GPT_E_ARG, "crooms: no kernel for this obs kind");
  if (smem > 40 * 1024) {  // static + dynamic shared memory above 48 KB needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(crooms)");
  }
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl(k, dim3(nblocks), dim3(threads), smem, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "crooms_step_kernel launch");
  return GPT_OK;
}

int tag_create(gpt_env* env, const gpt_config* c) {
  if (c->track_stats) return fail(GPT_E_ARG, "tag: track_stats is implemented for the Taxi and ROOMS families only");
  const int rdt = c->c_state_f32 ? GPT_DT_F32 : GPT_DT_F64;
  add_array(env, "agent", GPT_ROLE_STATE, rdt, 2);
  add_array(env, "target", GPT_ROLE_STATE, rdt, 2);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  add_array(env, "obs", GPT_ROLE_OUTPUT, rdt, 2);
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
  P.agent = env->ptr("agent");
  P.target = env->ptr("target");
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = a.actions;
  P.obs = env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.agent || !P.target || !P.elapsed || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "tag: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "tag: actions is NULL");
  P.obs = (uint8_t*)P.obs + a.out_row * (c.c_state_f32 ? sizeof(float2) : sizeof(double2));
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
  const bool devctr = env->graph_mode && !replay;
  P.ctr_ptr = env->d_counter;
  void* k = c.c_state_f32 ? tag_pick_f32(replay, devctr) : tag_pick_rr<double>(replay, devctr);
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl(k, dim3(nblocks), dim3(threads), 0, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "tag_step_kernel launch");
  return GPT_OK;
}

}  // namespace gpt
