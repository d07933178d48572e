// gpt_rooms.cu — host side of the fused ROOMS step: table building, array schema, launch.
// The kernel is in gpt_rooms_kernel.cuh; its instantiations in gpt_rooms_k*.cu.
#include "gpt_rooms_kernel.cuh"

namespace gpt {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static const int kDY[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int kDX[8] = {0, 1, 1, 1, 0, -1, -1, -1};

// Walker alias tables for the slip (Philox mode): per intended action, n columns of {threshold, dir | alias<<8}
std::vector<uint32_t> build_slip_alias(int n, const double* thr64) {
  const int shift = n == 4 ? 1 : 0;
  std::vector<uint32_t> alias((size_t)n * 8 * 2, 0u);
  for (int a = 0; a < n; ++a) {
    std::vector<double> q(n);
    for (int j = 0; j < n; ++j) {
      const double hi = j == n - 1 ? 1.0 : thr64[a * n + j];   // the last threshold never counts (= clamp to n-1)
      const double lo = j == 0 ? 0.0 : thr64[a * n + j - 1];
      q[j] = (hi > lo ? hi - lo : 0.0) * n;
    }
    std::vector<int> small, large, al(n);
    std::vector<double> pr(n, 1.0);
    for (int j = 0; j < n; ++j) { al[j] = j; (q[j] < 1.0 ? small : large).push_back(j); }
    while (!small.empty() && !large.empty()) {
      const int sidx = small.back(), lidx = large.back();
      small.pop_back();
      pr[sidx] = q[sidx];
      al[sidx] = lidx;
      q[lidx] = (q[lidx] + q[sidx]) - 1.0;
      if (q[lidx] < 1.0) { large.pop_back(); small.push_back(lidx); }
    }
    for (int j = 0; j < n; ++j) {
      const double t = pr[j] * 4294967296.0;
      alias[((size_t)a * 8 + j) * 2] = t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0 ? 0u : (uint32_t)t);
      alias[((size_t)a * 8 + j) * 2 + 1] = (uint32_t)(j << shift) | ((uint32_t)(al[j] << shift) << 8);
    }
  }
  return alias;
}

int rooms_build_tables(gpt_env* env, const gpt_config* c, bool discrete_actions) {
  const int h = c->rooms_h, w = c->rooms_w;
  if (h < 3 || w < 3 || h > 255 || w > 255 || (int64_t)h * w >= 32768) return fail(GPT_E_ARG, "rooms: grid shape out of range");
  if (!c->rooms_grid) return fail(GPT_E_ARG, "rooms: grid missing");
  const int8_t* g = c->rooms_grid;
  const int nc = h * w;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if ((y == 0 || x == 0 || y == h - 1 || x == w - 1) && g[y * w + x] >= 0)
        return fail(GPT_E_ARG, "rooms: the map must have a solid wall border");
  std::vector<uint8_t> nb8(nc, 0), room(nc, 0xFF);
  std::vector<uint16_t> sid(nc, 0), valid;
  int run = 0, max_room = -1;
  for (int i = 0; i < nc; ++i) {
    if (g[i] >= 0) {
      ++run;
      valid.push_back((uint16_t)i);
      room[i] = (uint8_t)g[i];
      if (g[i] > max_room) max_room = g[i];
      const int y = i / w, x = i % w;
      for (int d = 0; d < 8; ++d)
        if (g[(y + kDY[d]) * w + (x + kDX[d])] >= 0) nb8[i] |= (uint8_t)(1u << d);
    }
    sid[i] = (uint16_t)(run > 0 ? run - 1 : 0);   // (grid>=0).cumsum()-1  (observations.py:27)
  }
  if (valid.empty()) return fail(GPT_E_ARG, "rooms: no walkable cell");
  // number of rooms = number of distinct non-wall ids (observations.py:40)
  std::vector<bool> seen(256, false);
  int n_rooms = 0;
  for (int i = 0; i < nc; ++i)
    if (g[i] >= 0 && !seen[(uint8_t)g[i]]) { seen[(uint8_t)g[i]] = true; ++n_rooms; }
  env->rl.n_valid = (int32_t)valid.size();
  env->rl.n_rooms = n_rooms;
  env->rl.n_cells = nc;

  std::vector<double> thr64;
  if (discrete_actions) {
    const int n = c->rooms_n_actions;
    if (n != 4 && n != 8) return fail(GPT_E_ARG, "rooms: n_actions must be 4 (cardinal) or 8 (ordinal)");
    if (!c->rooms_slip_cumsum) return fail(GPT_E_ARG, "rooms: slip_cumsum missing");
    thr64.assign(c->rooms_slip_cumsum, c->rooms_slip_cumsum + n * n);   // replay mode compares the recorded u with these
  }
  // walkable-bit rows padded by the window radius (grid obs)
  std::vector<uint64_t> rows;
  const int gn = c->rooms_obs_kind == GPT_OBS_GRID ? c->rooms_obs_n : 0;
  if (gn) {
    if (gn < 1 || gn > 15) return fail(GPT_E_ARG, "rooms: grid obs_n must be in [1, 15]");
    const int off = gn / 2;
    if (w + 2 * off > 64) return fail(GPT_E_ARG, "rooms: map too wide for the bit-row window (w + 2*(n//2) <= 64)");
    rows.assign(h + 2 * off + 1, 0ull);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        if (g[y * w + x] >= 0) rows[y + off] |= 1ull << (x + off);
  }
  // move table: [cell*8 + ordinal dir] -> next cell, bit 15 set when the target is a wall (agent stays)
  std::vector<uint16_t> move((size_t)nc * 8);
  for (int i = 0; i < nc; ++i)
    for (int d = 0; d < 8; ++d) {
      const int y = i / w + kDY[d], x = i % w + kDX[d];
      const bool free_target = g[i] >= 0 && y >= 0 && y < h && x >= 0 && x < w && g[y * w + x] >= 0;
      move[(size_t)i * 8 + d] = (uint16_t)(free_target ? (y * w + x) : (i | 0x8000));
    }
  // fixed goal: every non-window observation is a function of the agent cell alone -> tabulate it
  std::vector<uint32_t> obstab;
  const int kind = c->rooms_obs_kind;
  if (c->family == GPT_FAMILY_ROOMS && c->rooms_goal_y >= 0 && kind != GPT_OBS_GRID) {
    const int gy = c->rooms_goal_y, gx = c->rooms_goal_x;
    const bool inside = gy < h && gx >= 0 && gx < w;
    const int gcell = inside ? gy * w + gx : 0;
    const bool two = (kind == GPT_OBS_VEC_HANSEN || kind == GPT_OBS_VEC_HANSEN_GOAL) && c->rooms_obs_n == 8;
    obstab.assign((size_t)nc * (two ? 2 : 1), 0u);
    auto spread4 = [](uint32_t b) { return (b * 0x00204081u) & 0x01010101u; };
    for (int i = 0; i < nc; ++i) {
      const int y = i / w, x = i % w;
      const uint32_t nb = nb8[i];
      const uint32_t b4 = (nb & 1u) | ((nb >> 1) & 2u) | ((nb >> 2) & 4u) | ((nb >> 3) & 8u);
      int gd = -1;  // ordinal index of the neighbour holding the goal
      for (int d = 0; d < 8; ++d)
        if (y + kDY[d] == gy && x + kDX[d] == gx) gd = d;
      uint32_t lo = 0, hi = 0;
      switch (kind) {
        case GPT_OBS_ROOM: lo = room[i]; break;
        case GPT_OBS_ROOM_GOAL: lo = room[i] + (uint32_t)n_rooms * room[gcell]; break;
        case GPT_OBS_MDP: lo = sid[i]; break;
        case GPT_OBS_MDP_GOAL: lo = sid[i] + (uint32_t)valid.size() * sid[gcell]; break;
        case GPT_OBS_VEC_MDP: lo = (uint32_t)y | ((uint32_t)x << 8); break;
        case GPT_OBS_VEC_MDP_GOAL: lo = (uint32_t)y | ((uint32_t)x << 8) | ((uint32_t)gy << 16) | ((uint32_t)gx << 24); break;
        case GPT_OBS_HANSEN:
          if (c->rooms_obs_n == 8) lo = nb * (gd < 0 ? 1u : (uint32_t)gd + 1u);
          else lo = b4 * ((gd >= 0 && !(gd & 1)) ? (uint32_t)(gd >> 1) + 1u : 1u);
          break;
        case GPT_OBS_VEC_HANSEN:
        case GPT_OBS_VEC_HANSEN_GOAL: {
          const int g8 = kind == GPT_OBS_VEC_HANSEN_GOAL ? gd : -1;
          if (c->rooms_obs_n == 8) {
            lo = spread4(nb & 15u);
            hi = spread4(nb >> 4);
            if (g8 >= 0 && g8 < 4) lo = (lo & ~(0xFFu << (8 * g8))) | (2u << (8 * g8));
            else if (g8 >= 4) hi = (hi & ~(0xFFu << (8 * (g8 - 4)))) | (2u << (8 * (g8 - 4)));
          } else {
            lo = spread4(b4);
            if (g8 >= 0 && !(g8 & 1)) lo = (lo & ~(0xFFu << (4 * g8))) | (2u << (4 * g8));
          }
          break;
        }
      }
      if (two) { obstab[(size_t)i * 2] = lo; obstab[(size_t)i * 2 + 1] = hi; }
      else obstab[i] = lo;
    }
  }
  // merged table for 16-bit observations: next | on-goal<<14 | blocked<<15 | obs(next)<<16
  std::vector<uint32_t> moveobs;
  const bool merged = !obstab.empty() && (kind == GPT_OBS_ROOM || kind == GPT_OBS_ROOM_GOAL || kind == GPT_OBS_MDP ||
                                          kind == GPT_OBS_HANSEN || kind == GPT_OBS_VEC_MDP);
  if (merged) {
    if (nc > 0x3FFF) return fail(GPT_E_ARG, "rooms: grid too large for the merged move table");
    const int gy = c->rooms_goal_y, gx = c->rooms_goal_x;
    const int gcell = (gy < h && gx >= 0 && gx < w) ? gy * w + gx : -1;
    moveobs.resize((size_t)nc * 8);
    for (size_t i = 0; i < moveobs.size(); ++i) {
      const uint32_t nxt = move[i] & 0x7FFFu;
      if (obstab[nxt] > 0xFFFFu) return fail(GPT_E_ARG, "rooms: internal: observation does not fit 16 bits");
      moveobs[i] = nxt | ((int)nxt == gcell ? 0x4000u : 0u) | (move[i] & 0x8000u) | (obstab[nxt] << 16);
    }
    move.clear();   // the merged kernels never read the plain move table
  }
  std::vector<uint32_t> alias;
  if (discrete_actions) alias = build_slip_alias(c->rooms_n_actions, thr64.data());
  std::vector<uint8_t> blob;
  env->rl.move_off = blob_append(blob, move);
  env->rl.moveobs_off = blob_append(blob, moveobs);
  env->rl.alias_off = blob_append(blob, alias);
  env->rl.obstab_off = blob_append(blob, obstab);
  env->rl.nb8_off = blob_append(blob, nb8);
  env->rl.room_off = blob_append(blob, room);
  env->rl.sid_off = blob_append(blob, sid);
  env->rl.valid_off = blob_append(blob, valid);
  env->rl.thr64_off = blob_append(blob, thr64);
  env->rl.rows_off = blob_append(blob, rows);
  if (!discrete_actions || c->family == GPT_FAMILY_CROOMS) {  // continuous env: wall test on floor(pos / cell)
    std::vector<int8_t> grid(g, g + nc);
    env->rl.grid_off = blob_append(blob, grid);
  }
  return upload_blob(env, blob);
}

static int obs_desc(const gpt_config* c, int* dtype, int* cols) {
  switch (c->rooms_obs_kind) {
    case GPT_OBS_ROOM: case GPT_OBS_ROOM_GOAL: case GPT_OBS_MDP: case GPT_OBS_MDP_GOAL: case GPT_OBS_HANSEN:
      *dtype = GPT_DT_I32; *cols = 1; return GPT_OK;
    case GPT_OBS_VEC_MDP: *dtype = GPT_DT_U8; *cols = 2; return GPT_OK;
    case GPT_OBS_VEC_MDP_GOAL: *dtype = GPT_DT_U8; *cols = 4; return GPT_OK;
    case GPT_OBS_VEC_HANSEN: case GPT_OBS_VEC_HANSEN_GOAL: *dtype = GPT_DT_U8; *cols = c->rooms_obs_n; return GPT_OK;
    case GPT_OBS_GRID: *dtype = GPT_DT_U8; *cols = c->rooms_obs_n * c->rooms_obs_n; return GPT_OK;
  }
  return fail(GPT_E_ARG, "rooms: unknown obs kind");
}

int rooms_create(gpt_env* env, const gpt_config* c) {
  if ((c->rooms_obs_kind == GPT_OBS_HANSEN || c->rooms_obs_kind == GPT_OBS_VEC_HANSEN || c->rooms_obs_kind == GPT_OBS_VEC_HANSEN_GOAL) &&
      c->rooms_obs_n != 4 && c->rooms_obs_n != 8)
    return fail(GPT_E_ARG, "rooms: hansen obs_n must be 4 or 8");
  if (c->env_offset % GPT_ENV_ALIGN != 0) return fail(GPT_E_ARG, "rooms: env_offset must be a multiple of GPT_ENV_ALIGN");
  if (int rc = rooms_build_tables(env, c, true)) return rc;
  const bool rgoal = c->rooms_goal_y < 0;
  if (!rgoal) {
    // A fixed goal outside the grid is legal and simply unreachable: the reference's default goal for
    // layouts '32'/'32b' is ENDS = (x 47, y 32) on a 25 x 49 grid (layouts.py:197-205, rooms.py:153-158).
    const bool inside = c->rooms_goal_y < c->rooms_h && c->rooms_goal_x >= 0 && c->rooms_goal_x < c->rooms_w;
    if (!inside && (c->rooms_obs_kind == GPT_OBS_ROOM_GOAL || c->rooms_obs_kind == GPT_OBS_MDP_GOAL))
      return fail(GPT_E_ARG, "rooms: goal-indexed observations need the fixed goal inside the grid");
    if (c->rooms_goal_y > 255 || c->rooms_goal_x > 255 || c->rooms_goal_x < 0) return fail(GPT_E_ARG, "rooms: fixed goal out of range");
  }
  int odt, ocols;
  if (int rc = obs_desc(c, &odt, &ocols)) return rc;
  add_array(env, "pos", GPT_ROLE_STATE, GPT_DT_U16, 1);
  if (rgoal) add_array(env, "goal", GPT_ROLE_STATE, GPT_DT_U16, 1);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  if (c->track_stats) add_array(env, "ep_return", GPT_ROLE_STATE, GPT_DT_F32, 1);
  add_array(env, "obs", GPT_ROLE_OUTPUT, odt, ocols);
  add_array(env, "reward", GPT_ROLE_OUTPUT, GPT_DT_F32, 1);
  add_array(env, "terminated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "truncated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "replay_u", GPT_ROLE_REPLAY, GPT_DT_F64, 1);
  add_array(env, "replay_reset_agent", GPT_ROLE_REPLAY, GPT_DT_I32, 1);
  add_array(env, "replay_reset_goal", GPT_ROLE_REPLAY, GPT_DT_I32, 1);
  add_array(env, "actions", GPT_ROLE_ACTION, GPT_DT_I8, 1);
  env->action_dtype = GPT_DT_I8;
  env->action_cols = 1;
  return GPT_OK;
}


static void* pick_kernel(int obs, int grid_n, bool rgoal, bool replay, int stats) {
  switch (obs) {
    case GPT_OBS_ROOM: case GPT_OBS_ROOM_GOAL: case GPT_OBS_MDP: case GPT_OBS_MDP_GOAL: return rooms_pick_table(obs, rgoal, replay, stats);
    case GPT_OBS_VEC_MDP: case GPT_OBS_VEC_MDP_GOAL: case GPT_OBS_HANSEN: return rooms_pick_vec(obs, rgoal, replay, stats);
    case GPT_OBS_VEC_HANSEN: case GPT_OBS_VEC_HANSEN_GOAL: return rooms_pick_vhansen(obs, rgoal, replay, stats);
    case GPT_OBS_GRID:
      if (grid_n == 3 || grid_n == 5) return rooms_pick_grid_small(grid_n, rgoal, replay, stats);
      if (grid_n == 7 || grid_n == 9) return rooms_pick_grid_large(grid_n, rgoal, replay, stats);
      return rooms_pick_grid_any(rgoal, replay, stats);
  }
  return nullptr;
}

// (the fused kernels keep the elapsed counters as biased 16-bit pairs: time limits beyond 32766 step one launch at a time)
bool rooms_can_fuse(const gpt_env* env) {
  return env->cfg.rng_mode == GPT_RNG_PHILOX && !env->cfg.track_stats && env->cfg.time_limit >= 0 && env->cfg.time_limit <= 32766;
}

int rooms_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool rgoal = c.rooms_goal_y < 0;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  RoomsParams P{};
  P.pos = (uint16_t*)env->ptr("pos");
  P.goal = rgoal ? (uint16_t*)env->ptr("goal") : nullptr;
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = (const int8_t*)a.actions;
  P.obs = env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.pos || (rgoal && !P.goal) || !P.elapsed || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "rooms: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "rooms: actions is NULL");
  const bool reset = a.mode == kModeReset;
  if (reset) {
    // reset() = every env truncates: poison `elapsed` (0x7F7F7F7F > any time limit), step once with any
    // valid action bytes (the `terminated` array holds 0/1), then clear the flags the step wrote.
    if (c.time_limit >= 0x7F7F7F7E) return fail(GPT_E_ARG, "rooms: time_limit too large");
    cudaError_t e = cudaMemsetAsync(P.elapsed, 0x7F, (size_t)env->capacity * sizeof(int32_t), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(elapsed)");
    P.actions = (const int8_t*)P.terminated;
  }
  const int oi = env->find("obs");
  const size_t obs_row = (size_t)env->arrays[oi].desc.cols * env->arrays[oi].desc.elem_size;
  P.obs = (uint8_t*)P.obs + a.out_row * obs_row;
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  if (replay) {
    P.rp_u = (const double*)env->ptr("replay_u");
    P.rp_reset_agent = (const int32_t*)env->ptr("replay_reset_agent");
    P.rp_reset_goal = (const int32_t*)env->ptr("replay_reset_goal");
    if (!P.rp_u || !P.rp_reset_agent || !P.rp_reset_goal) return fail(GPT_E_UNBOUND, "rooms: replay arrays must be bound in replay mode");
  }
  if (c.track_stats) {
    P.ep_return = (float*)env->ptr("ep_return");
    P.stats = env->d_stats;
    if (!P.ep_return) return fail(GPT_E_UNBOUND, "rooms: ep_return must be bound when track_stats=1");
  }
  P.num_envs = reset ? 0 : c.num_envs;   // a reset() does not count as finished episodes
  P.blob = env->d_blob;
  P.blob_bytes = env->blob_bytes;
  P.nb8_off = env->rl.nb8_off;
  P.room_off = env->rl.room_off;
  P.sid_off = env->rl.sid_off;
  P.valid_off = env->rl.valid_off;
  P.thr64_off = env->rl.thr64_off;
  P.rows_off = env->rl.rows_off;
  P.move_off = env->rl.move_off;
  P.obstab_off = env->rl.obstab_off;
  P.moveobs_off = env->rl.moveobs_off;
  P.alias_off = env->rl.alias_off;
  P.log2n = c.rooms_n_actions == 4 ? 2u : 3u;
  P.stage_off = (env->blob_bytes + 127u) & ~127u;
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.mode = a.mode;
  P.w = c.rooms_w;
  P.n_actions = c.rooms_n_actions;
  P.n_valid = env->rl.n_valid;
  P.n_rooms = env->rl.n_rooms;
  P.time_limit = c.time_limit;
  const bool grid = c.rooms_obs_kind == GPT_OBS_GRID;
  P.hansen_n = grid ? 0 : c.rooms_obs_n;
  P.grid_n = grid ? c.rooms_obs_n : 0;
  P.goal_y = rgoal ? 0 : c.rooms_goal_y;
  P.goal_x = rgoal ? 0 : c.rooms_goal_x;
  const bool goal_inside = !rgoal && c.rooms_goal_y < c.rooms_h && c.rooms_goal_x < c.rooms_w;
  P.goal_cell = goal_inside ? c.rooms_goal_y * c.rooms_w + c.rooms_goal_x : 0xFFFF;  // 0xFFFF never equals a cell
  P.div_w = make_fastdiv((uint32_t)c.rooms_w);
  P.r_step = c.rooms_step_reward;
  P.r_wall = c.rooms_wall_reward;
  P.r_goal = c.rooms_goal_reward;
  P.rng = make_rng_key(env);

  const int threads = 128, warps = threads / 32;
  int qpt = RoomsShape<GPT_OBS_MDP, 0>::kQpt;
  if (grid) {
    switch (P.grid_n) {
      case 3: qpt = RoomsShape<GPT_OBS_GRID, 3>::kQpt; break;
      case 5: qpt = RoomsShape<GPT_OBS_GRID, 5>::kQpt; break;
      case 7: qpt = RoomsShape<GPT_OBS_GRID, 7>::kQpt; break;
      case 9: qpt = RoomsShape<GPT_OBS_GRID, 9>::kQpt; break;
      default: qpt = RoomsShape<GPT_OBS_GRID, 0>::kQpt; break;
    }
  }
  const int64_t envs_per_cta = (int64_t)threads * kQuad * qpt;
  const int nblocks = (int)(((int64_t)a.n_tiles * kTileEnvs + envs_per_cta - 1) / envs_per_cta);
  if (nblocks <= 0) return GPT_OK;
  size_t smem = env->blob_bytes;
  if (grid) smem = P.stage_off + (size_t)warps * kQuadStride * P.grid_n * P.grid_n;
  const bool multi = a.n_steps > 1;
  if (multi && (replay || c.track_stats)) return fail(GPT_E_ARG, "rooms: fused multi-step launches need Philox mode without track_stats");
  P.n_steps = a.n_steps;
  P.act_stride = env->capacity;
  P.out_stride = a.out_stride_rows;
  // window obs: one TMA bulk store per warp tile (128 envs x n^2 bytes) when every tile starts 16-byte aligned; fused
  // in-place outputs keep the per-lane stores (bulk stores of different steps to the same rows would be unordered).
  // GPT_ROOMS_OBS_TMA=0 switches it off (A/B runs).
  static const bool obs_tma_off = getenv("GPT_ROOMS_OBS_TMA") && atoi(getenv("GPT_ROOMS_OBS_TMA")) == 0;
  P.obs_tma = grid && !obs_tma_off && ((uintptr_t)P.obs & 15u) == 0 && (a.out_stride_rows & 15) == 0 && !(multi && a.out_stride_rows == 0);
  const bool devctr = env->graph_mode && !replay && !c.track_stats;   // graph mode: step counter in device memory
  P.ctr_ptr = env->d_counter;
  void* k = pick_kernel(c.rooms_obs_kind, P.grid_n, rgoal, replay, devctr ? (multi ? 4 : 3) : (multi ? 2 : (c.track_stats != 0 ? 1 : 0)));
  if (!k) return fail(GPT_E_ARG, "rooms: no kernel for this obs kind");
  if (smem > 40 * 1024) {  // static + dynamic shared memory above 48 KB needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(rooms)");
  }
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl(k, dim3(nblocks), dim3(threads), smem, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "rooms_step_kernel launch");
  if (reset) {
    e = cudaMemsetAsync(P.terminated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.truncated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.reward, 0, (size_t)env->capacity * sizeof(float), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reset outputs)");
  }
  return GPT_OK;
}

}  // namespace gpt
