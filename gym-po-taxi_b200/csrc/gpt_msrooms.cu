// gpt_msrooms.cu — fused multistory FourRooms step for sm_100a.
//
// One kernel = MultistoryFourRoomsEnv.step (reference gym_po/envs/rooms/msrooms.py:392-413): ++elapsed, action
// slip (rooms/action_utils.py:73-90), move unless the target is a wall (:400-402, :415-417), stairs teleport for
// agents that moved (:419-428), reward step -> wall -> goal (:404-408), terminated = on goal, truncated =
// elapsed > limit (:409), same-step respawn goal-then-agent (:385-390; agent on the bottom floor, random goals on
// the top floor, :306-313) and the observation of the post-reset state (:131-254).
//
// The whole map logic is folded into tables on the host (msrooms_create), staged into shared memory with one TMA
// bulk copy per CTA:
//   move    uint16[cells*8]  [cell*8 + ordinal dir] -> cell after the move INCLUDING the stair teleport,
//                            bit 15 = target is a wall (agent stays)
//   obstab  uint32[cells*W]  fixed goal: the observation is a function of the agent cell alone (W = 1 or 2 words)
//   moveobs uint32[cells*8]  fixed goal, observation fits 16 bits: next | on-goal<<14 | blocked<<15 | obs(next)<<16
//   info / sid / h3 / vh     random goal: per-cell (z,y,x,grid value), dense id, base-3 Hansen sum, Hansen bytes
//   alias / thr64            slip (Philox: Walker alias per intended action; replay: float64 cumsum rows)
//   avalid / gvalid          spawn cells (bottom floor) / random-goal cells (top floor)
//
// HBM layout (SoA) and thread mapping as in gpt_rooms_kernel.cuh: pos uint16 | [goal uint16] | elapsed int32 |
// action int8 -> pos, [goal], elapsed, obs, reward f32, terminated u8, truncated u8.  B_alg = 23 B per env-step
// (fixed goal, scalar obs), like ROOMS.
#include "gpt_internal.h"

namespace gpt {

struct MsTables {
  const uint32_t* info;   // z | y<<8 | x<<16 | grid value<<24
  const uint16_t* sid;    // dense id of the cell among the walkable cells (cumsum(grid-1 >= 0) - 1, msrooms.py:229)
  const uint16_t* h3;     // sum_i digit_i 3^i over the 4 / 8 neighbours, digit = 0 wall, 2 walkable (:176-189)
  const uint32_t* vh;     // [cell*2] the same digits as bytes (:146-156)
};
struct MsObsCfg {
  int kind, hansen_n, n_free, room_n;
};

// ordinal index of (dy, dx) in N,NE,E,SE,S,SW,W,NW; 0xF for (0,0)
__host__ __device__ inline uint32_t ms_dir_of(int dy, int dx) {
  return (uint32_t)((0x3452F6107ull >> (4 * ((dy + 1) * 3 + (dx + 1)))) & 0xFull);
}

// Observation of one env from its agent cell and goal cell (msrooms.py:192-254).  Used on the host to tabulate
// the fixed-goal observation per cell and on the device for random goals.  Scalar kinds return the value in `lo`;
// byte-vector kinds return up to 8 little-endian packed bytes in lo/hi.
__host__ __device__ inline void ms_obs(const MsTables& T, const MsObsCfg& C, uint32_t cell, uint32_t gcell, uint32_t& lo, uint32_t& hi) {
  lo = 0;
  hi = 0;
  const uint32_t ia = T.info[cell], ig = T.info[gcell];
  switch (C.kind) {
    case GPT_OBS_ROOM: lo = ia >> 24; break;                                            // raw grid value (:213)
    case GPT_OBS_ROOM_GOAL:                                                             // (:208-211), may be negative
      lo = (uint32_t)(((int)(ia >> 24) - 4) + C.room_n * ((int)(ig >> 24) - 4));
      break;
    case GPT_OBS_MDP: lo = T.sid[cell]; break;
    case GPT_OBS_MDP_GOAL: lo = (uint32_t)T.sid[cell] + (uint32_t)C.n_free * (uint32_t)T.sid[gcell]; break;
    case GPT_OBS_VEC_MDP: lo = ia & 0xFFFFFFu; break;
    case GPT_OBS_VEC_MDP_GOAL:
      lo = (ia & 0xFFFFFFu) | (ig << 24);
      hi = (ig >> 8) & 0xFFFFu;
      break;
    default: {  // Hansen family: which neighbour (index into the 4 / 8 directions) holds the goal, if any
      const int dz = (int)(ig & 0xFFu) - (int)(ia & 0xFFu);
      const int dy = (int)((ig >> 8) & 0xFFu) - (int)((ia >> 8) & 0xFFu);
      const int dx = (int)((ig >> 16) & 0xFFu) - (int)((ia >> 16) & 0xFFu);
      uint32_t gd = 0xFu;
      if (dz == 0 && (unsigned)(dy + 1) <= 2u && (unsigned)(dx + 1) <= 2u) gd = ms_dir_of(dy, dx);
      if (C.hansen_n == 4) gd = (gd != 0xFu && !(gd & 1u)) ? gd >> 1 : 0xFu;
      if (C.kind == GPT_OBS_HANSEN) {
        lo = (uint32_t)T.h3[cell] * (gd == 0xFu ? 1u : gd + 1u);
      } else {
        lo = T.vh[cell * 2];
        hi = T.vh[cell * 2 + 1];
        if (C.kind == GPT_OBS_VEC_HANSEN_GOAL && gd != 0xFu) {
          if (gd < 4u) lo = (lo & ~(0xFFu << (8 * gd))) | (3u << (8 * gd));
          else hi = (hi & ~(0xFFu << (8 * (gd - 4)))) | (3u << (8 * (gd - 4)));
        }
      }
    }
  }
}

struct MsParams {
  uint16_t* pos;
  uint16_t* goal;
  int32_t* elapsed;
  const int8_t* actions;
  uint8_t* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  const double* rp_u;
  const int32_t* rp_reset_agent;
  const int32_t* rp_reset_goal;
  const uint8_t* blob;
  uint32_t blob_bytes, move_off, moveobs_off, obstab_off, info_off, sid_off, h3_off, vh_off, alias_off, thr64_off, avalid_off, gvalid_off;
  uint32_t log2n;
  int64_t env_offset;
  int32_t first_tile, n_tiles;
  int32_t n_actions, n_agent, n_goal, time_limit, goal_cell;
  MsObsCfg oc;
  float r_step, r_wall, r_goal;
  int32_t n_steps;      // fused multi-step launch (MULTI kernels): steps per launch
  int64_t act_stride;   // bytes between consecutive steps' action rows (= capacity)
  int64_t out_stride;   // rows between consecutive steps' outputs (0 = overwrite in place)
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};

constexpr int kMsThreads = 128, kMsQpt = 2;

// _reset_some (msrooms.py:385-390): goal first (random-goal envs), then agent.  Rare, out of line.
template <bool RGOAL, bool REPLAY, bool DEVCTR = false>
__device__ __forceinline__ uint32_t ms_respawn_inline(const MsParams& P, const uint16_t* avalid, const uint16_t* gvalid, int64_t env, uint32_t gcell, uint32_t t, uint64_t ctr_dev = 0) {
  uint32_t cell;
  if (REPLAY) {
    if (RGOAL) gcell = (uint32_t)P.rp_reset_goal[env];
    cell = (uint32_t)P.rp_reset_agent[env];
  } else {
    const uint64_t ctr = (DEVCTR ? ctr_dev : (((uint64_t)P.rng.step_hi << 32) | P.rng.step_lo)) + t;   // step index inside a fused launch
    const uint64_t ge = (uint64_t)(P.env_offset + env);
    const uint4 r = philox4x32<kRounds>(make_uint4((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)ctr,
                                                                ((uint32_t)(ctr >> 32) & 0x00FFFFFFu) ^ (1u << 24)), P.rng);
    if (RGOAL) gcell = gvalid[bounded(r.y, (uint32_t)P.n_goal)];
    cell = avalid[bounded(r.x, (uint32_t)P.n_agent)];
  }
  return cell | (gcell << 16);
}
// out-of-line copy for the single-step kernels
template <bool RGOAL, bool REPLAY>
__device__ __noinline__ uint32_t ms_respawn(const MsParams& P, const uint16_t* avalid, const uint16_t* gvalid, int64_t env, uint32_t gcell, uint32_t t) {
  return ms_respawn_inline<RGOAL, REPLAY>(P, avalid, gvalid, env, gcell, t);
}

// word W of the 4*OB contiguous observation bytes of a quad (env k contributes bytes e[k] >> 8*o, o < OB)
template <int OB>
__device__ __forceinline__ uint32_t quad_obs_word(const uint64_t (&e)[4], int w) {
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = 4 * w + i, k = b / OB, o = b % OB;
    r |= (uint32_t)((e[k] >> (8 * o)) & 0xFFull) << (8 * i);
  }
  return r;
}

// OB = observation bytes per env (4: int32 scalar or 4 Hansen bytes, 3: z,y,x, 6: z,y,x,gz,gy,gx, 8: 8 Hansen bytes)
// RGOAL = random goal (observation computed from the per-cell tables), otherwise tabulated per cell;
// MERGED = fixed goal and a 16-bit observation riding in the move-table entry.
// MULTI: gpt_step_many as ONE launch (state in registers for P.n_steps steps), bit-identical to single-step launches.
#ifndef GPT_MS_MINB_MULTI
#define GPT_MS_MINB_MULTI 7    // measured on B200: 6 -> 348 G, 7 -> 361 G env-steps/s (fused), single step 8 -> 251 G, 7 -> 256 G
#endif
#ifndef GPT_MS_MINB_SINGLE
#define GPT_MS_MINB_SINGLE 7
#endif
template <int OB, bool RGOAL, bool MERGED, bool REPLAY, bool MULTI = false, bool DEVCTR = false>
__global__ void __launch_bounds__(kMsThreads, MULTI ? GPT_MS_MINB_MULTI : GPT_MS_MINB_SINGLE) msrooms_step_kernel(const __grid_constant__ MsParams P) {
  static_assert(!MULTI || !REPLAY, "fused launches need Philox mode");
  constexpr int kEnvsPerWarp = kWarp * kQuad * kMsQpt;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.blob_bytes, &bar);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t wbase = first + ((int64_t)blockIdx.x * (kMsThreads / 32) + warp) * kEnvsPerWarp;
  if (wbase >= last) return;
  const int64_t base = wbase + lane * kQuad;
  const uint32_t n = (uint32_t)P.n_actions;
  const uint32_t dir_shift = n == 4 ? 1u : 0u;   // cardinal action i = ordinal direction 2i

  pdl_wait();   // reset() = poisoned `elapsed` + this same kernel (msrooms_launch)
  uint2 pos4[kMsQpt], goal4[kMsQpt];
  int4 e4[kMsQpt];
  uint32_t a4[kMsQpt];
#pragma unroll
  for (int j = 0; j < kMsQpt; ++j) {
    const int64_t q = base + j * kQuadStride;
    goal4[j] = make_uint2(0, 0);
    pos4[j] = ld_stream(reinterpret_cast<const uint2*>(P.pos + q));
    if (RGOAL) goal4[j] = ld_stream(reinterpret_cast<const uint2*>(P.goal + q));
    e4[j] = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    a4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + q));
  }

  stage_tables_wait(&bar);
  const uint16_t* move = reinterpret_cast<const uint16_t*>(smem + P.move_off);
  const uint32_t* moveobs = reinterpret_cast<const uint32_t*>(smem + P.moveobs_off);
  const uint32_t* obstab = reinterpret_cast<const uint32_t*>(smem + P.obstab_off);
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);
  const double* thr64 = reinterpret_cast<const double*>(smem + P.thr64_off);
  const uint16_t* avalid = reinterpret_cast<const uint16_t*>(smem + P.avalid_off);
  const uint16_t* gvalid = reinterpret_cast<const uint16_t*>(smem + P.gvalid_off);
  MsTables T;
  T.info = reinterpret_cast<const uint32_t*>(smem + P.info_off);
  T.sid = reinterpret_cast<const uint16_t*>(smem + P.sid_off);
  T.h3 = reinterpret_cast<const uint16_t*>(smem + P.h3_off);
  T.vh = reinterpret_cast<const uint32_t*>(smem + P.vh_off);
  const uint32_t col_shift = 32u - P.log2n;

  uint32_t cellq[kMsQpt][4], goalq[kMsQpt][4];   // state in registers (across the steps of a fused launch)
  int32_t evq[kMsQpt][4];
#pragma unroll
  for (int j = 0; j < kMsQpt; ++j) {
    cellq[j][0] = pos4[j].x & 0xFFFFu; cellq[j][1] = pos4[j].x >> 16; cellq[j][2] = pos4[j].y & 0xFFFFu; cellq[j][3] = pos4[j].y >> 16;
    goalq[j][0] = goal4[j].x & 0xFFFFu; goalq[j][1] = goal4[j].x >> 16; goalq[j][2] = goal4[j].y & 0xFFFFu; goalq[j][3] = goal4[j].y >> 16;
    evq[j][0] = e4[j].x; evq[j][1] = e4[j].y; evq[j][2] = e4[j].z; evq[j][3] = e4[j].w;
  }
  // MULTI: elapsed counters of a quad as biased 16-bit pairs (see rooms_step_kernel): one packed add per pair and step,
  // bit 15 of a half = "elapsed > time_limit"; the host fuses only with time_limit <= 32766
  const uint32_t cbias = 0x7FFFu - (uint32_t)P.time_limit;
  uint32_t cntq[MULTI ? kMsQpt : 1][2];
  if constexpr (MULTI) {
#pragma unroll
    for (int j = 0; j < kMsQpt; ++j) {
      uint32_t c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = (uint32_t)min(max(evq[j][k], 0), P.time_limit) + cbias;
      cntq[j][0] = c[0] | (c[2] << 16);
      cntq[j][1] = c[1] | (c[3] << 16);
    }
  }
  const int32_t n_steps = MULTI ? P.n_steps : 1;
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, (uint32_t)n_steps);
#pragma unroll 1
  for (int32_t t = 0; t < n_steps; ++t) {
  uint32_t a_next[kMsQpt];
  const bool more = MULTI && t + 1 < n_steps;
#pragma unroll
  for (int j = 0; j < kMsQpt; ++j) {   // prefetch the next step's action bytes (the only per-step read)
    a_next[j] = 0u;
    if (more) a_next[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + (int64_t)(t + 1) * P.act_stride + base + j * kQuadStride));
  }
  const int64_t orow = MULTI ? (int64_t)t * P.out_stride : 0;
  const uint64_t ctr = (DEVCTR ? ctr_dev : (((uint64_t)P.rng.step_hi << 32) | P.rng.step_lo)) + (uint32_t)t;
  const uint32_t ctr_lo = (uint32_t)ctr, ctr_hi = (uint32_t)(ctr >> 32) & 0x00FFFFFFu;
  // slip draws: one Philox4x32-7 block per quad and step (32 bits per env), the blocks of a quad pair computed together
  static_assert(kMsQpt % 2 == 0, "quads are processed in pairs");
  uint4 slipq[kMsQpt];
  if (!REPLAY) {
#pragma unroll
    for (int j = 0; j < kMsQpt; j += 2) {
      const uint64_t g0 = (uint64_t)(P.env_offset + base + j * kQuadStride) >> 2, g1 = (uint64_t)(P.env_offset + base + (j + 1) * kQuadStride) >> 2;
      slipq[j] = make_uint4((uint32_t)g0, (uint32_t)(g0 >> 32), ctr_lo, ctr_hi);
      slipq[j + 1] = make_uint4((uint32_t)g1, (uint32_t)(g1 >> 32), ctr_lo, ctr_hi);
      philox4x32_x2<kStepRounds>(slipq[j], slipq[j + 1], P.rng);
    }
  }
#pragma unroll
  for (int j = 0; j < kMsQpt; ++j) {
    const int64_t q = base + j * kQuadStride;
    uint32_t (&cellv)[4] = cellq[j];
    uint32_t (&goalv)[4] = goalq[j];
    int32_t (&ev)[4] = evq[j];
    float rv[4];
    uint32_t tw = 0, trw = 0, again = 0;
    uint32_t mraw[4] = {0, 0, 0, 0};   // merged move-table entries (fused launches: flag bytes by byte permutes)
    uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};

    const uint32_t slipv[4] = {REPLAY ? 0u : slipq[j].x, REPLAY ? 0u : slipq[j].y, REPLAY ? 0u : slipq[j].z, REPLAY ? 0u : slipq[j].w};

#pragma unroll
    for (int k = 0; k < 4; ++k) {  // straight-line transition of the 4 envs
      const uint32_t gcell = RGOAL ? goalv[k] : (uint32_t)P.goal_cell;
      if constexpr (!MULTI) ev[k] += 1;
      const uint32_t a = ((a4[j] >> (8 * k)) & 0xFFu) & (n - 1u);
      uint32_t d8;
      if (REPLAY) {  // a' = min(#{j : cumsum(P[a])_j < u}, n-1)   (action_utils.py:84-90)
        const double u = P.rp_u[q + k];
        const double* row = thr64 + a * n;
        uint32_t a2 = 0;
        for (uint32_t i = 0; i < n; ++i) a2 += row[i] < u ? 1u : 0u;
        d8 = (a2 < n ? a2 : n - 1) << dir_shift;
      } else {
        const uint32_t u = slipv[k];
        const uint2 e = alias[a * 8 + (u >> col_shift)];
        d8 = ((u << P.log2n) < e.x) ? (e.y & 0xFFu) : (e.y >> 8);
      }
      bool blocked, at_goal;
      if constexpr (MERGED) {
        const uint32_t m = moveobs[cellv[k] * 8 + d8];
        mraw[k] = m;
        cellv[k] = m & 0x3FFFu;
        blocked = (m & 0x8000u) != 0;
        at_goal = (m & 0x4000u) != 0;
        lo[k] = m >> 16;
      } else {
        const uint32_t mv = move[cellv[k] * 8 + d8];
        blocked = (mv & 0x8000u) != 0;
        cellv[k] = mv & 0x7FFFu;
        at_goal = cellv[k] == gcell;
      }
      rv[k] = at_goal ? P.r_goal : (blocked ? P.r_wall : P.r_step);
      if constexpr (!MULTI) {
        const bool trunc = ev[k] > P.time_limit;
        tw |= (at_goal ? 1u : 0u) << (8 * k);
        trw |= (trunc ? 1u : 0u) << (8 * k);
        again |= ((at_goal | trunc) ? 1u : 0u) << k;
      } else if constexpr (!MERGED) {
        tw |= (at_goal ? 1u : 0u) << (8 * k);
      }
      goalv[k] = gcell;
    }
    if constexpr (MULTI) {
      if constexpr (MERGED) {   // terminated bytes: the on-goal flag (bit 14) of the four merged entries
        const uint32_t hi01 = __byte_perm(mraw[0], mraw[1], 0x0051), hi23 = __byte_perm(mraw[2], mraw[3], 0x0051);
        tw = (__byte_perm(hi01, hi23, 0x5410) >> 6) & 0x01010101u;
      }
      cntq[j][0] += 0x00010001u;   // elapsed += 1 for the four envs; truncated bytes = bit 15 of the biased halves
      cntq[j][1] += 0x00010001u;
      trw = ((cntq[j][0] >> 15) & 0x00010001u) | ((cntq[j][1] >> 7) & 0x01000100u);
      again = tw | trw;   // one byte per env
    }
    if (again) {  // rare: respawn finished envs, one divergence point per quad
#pragma unroll 1
      for (uint32_t m = again; m; m &= m - 1) {
        const int k = MULTI ? (__ffs(m) - 1) >> 3 : __ffs(m) - 1;
        uint32_t g = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) g = i == k ? goalv[i] : g;
        // fused launches inline it: a CALL would wait for the in-flight action prefetch
        const uint32_t fresh = (MULTI || DEVCTR || GPT_RESPAWN_INLINE_SINGLE) ? ms_respawn_inline<RGOAL, REPLAY, DEVCTR>(P, avalid, gvalid, q + k, g, (uint32_t)t, ctr_dev) : ms_respawn<RGOAL, REPLAY>(P, avalid, gvalid, q + k, g, (uint32_t)t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i == k) {
            cellv[i] = fresh & 0xFFFFu;
            goalv[i] = fresh >> 16;
            ev[i] = 0;
            if constexpr (MERGED) lo[i] = obstab[fresh & 0xFFFFu];
          }
        }
        if constexpr (MULTI) {   // elapsed = 0: the env's counter half goes back to the bias
          const uint32_t half = (k & 2) ? 0xFFFF0000u : 0x0000FFFFu, fresh_c = (cbias | (cbias << 16)) & half;
          if (k & 1) cntq[j][1] = (cntq[j][1] & ~half) | fresh_c;
          else cntq[j][0] = (cntq[j][0] & ~half) | fresh_c;
        }
      }
    }
    // observation of the post-reset state
    uint64_t ob[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr (MERGED) {
      } else if constexpr (!RGOAL) {
        if constexpr (OB > 4) {
          const uint2 o = reinterpret_cast<const uint2*>(obstab)[cellv[k]];
          lo[k] = o.x;
          hi[k] = o.y;
        } else {
          lo[k] = obstab[cellv[k]];
        }
      } else {
        ms_obs(T, P.oc, cellv[k], goalv[k], lo[k], hi[k]);
      }
      ob[k] = (uint64_t)lo[k] | ((uint64_t)hi[k] << 32);
    }

    st_stream(reinterpret_cast<float4*>(P.reward + orow + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + orow + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + orow + q), trw);
    uint8_t* o = P.obs + (orow + q) * OB;   // the quad's 4*OB observation bytes are contiguous
    if constexpr (OB == 4) {
      st_stream(reinterpret_cast<int4*>(o), make_int4((int)lo[0], (int)lo[1], (int)lo[2], (int)lo[3]));
    } else if constexpr (OB == 8) {
      st_stream(reinterpret_cast<int4*>(o), make_int4((int)lo[0], (int)hi[0], (int)lo[1], (int)hi[1]));
      st_stream(reinterpret_cast<int4*>(o) + 1, make_int4((int)lo[2], (int)hi[2], (int)lo[3], (int)hi[3]));
    } else if constexpr (OB == 6) {
#pragma unroll
      for (int w = 0; w < 3; ++w)
        st_stream(reinterpret_cast<uint2*>(o) + w, make_uint2(quad_obs_word<6>(ob, 2 * w), quad_obs_word<6>(ob, 2 * w + 1)));
    } else {
#pragma unroll
      for (int w = 0; w < 3; ++w) st_stream(reinterpret_cast<uint32_t*>(o) + w, quad_obs_word<3>(ob, w));
    }
    a4[j] = a_next[j];
  }
  }  // steps of a fused launch
#pragma unroll
  for (int j = 0; j < kMsQpt; ++j) {
    const int64_t q = base + j * kQuadStride;
    const uint32_t (&cellv)[4] = cellq[j];
    const uint32_t (&goalv)[4] = goalq[j];
    st_stream(reinterpret_cast<uint2*>(P.pos + q), make_uint2(cellv[0] | (cellv[1] << 16), cellv[2] | (cellv[3] << 16)));
    if (RGOAL) st_stream(reinterpret_cast<uint2*>(P.goal + q), make_uint2(goalv[0] | (goalv[1] << 16), goalv[2] | (goalv[3] << 16)));
    if constexpr (MULTI)
      st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4((int)((cntq[j][0] & 0xFFFFu) - cbias), (int)((cntq[j][1] & 0xFFFFu) - cbias),
                                                                  (int)((cntq[j][0] >> 16) - cbias), (int)((cntq[j][1] >> 16) - cbias)));
    else
      st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(evq[j][0], evq[j][1], evq[j][2], evq[j][3]));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static const int kDY[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int kDX[8] = {0, 1, 1, 1, 0, -1, -1, -1};

static int ms_obs_bytes(const gpt_config* c) {
  switch (c->rooms_obs_kind) {
    case GPT_OBS_VEC_MDP: return 3;
    case GPT_OBS_VEC_MDP_GOAL: return 6;
    case GPT_OBS_VEC_HANSEN: case GPT_OBS_VEC_HANSEN_GOAL: return c->rooms_obs_n;
    default: return 4;
  }
}

int msrooms_create(gpt_env* env, const gpt_config* c) {
  const int h = c->rooms_h, w = c->rooms_w, S = c->ms_floors, n = c->rooms_n_actions, kind = c->rooms_obs_kind;
  if (h < 3 || w < 3 || h > 255 || w > 255 || S < 1 || S > 255) return fail(GPT_E_ARG, "msrooms: grid shape out of range");
  const int64_t nc64 = (int64_t)S * h * w;
  if (nc64 >= 32768) return fail(GPT_E_ARG, "msrooms: too many cells (floors*h*w must be < 32768)");
  const int nc = (int)nc64, fl = h * w;
  if (!c->rooms_grid || !c->rooms_slip_cumsum) return fail(GPT_E_ARG, "msrooms: floor map / slip_cumsum missing");
  if (n != 4 && n != 8) return fail(GPT_E_ARG, "msrooms: n_actions must be 4 (cardinal) or 8 (ordinal)");
  if (kind < GPT_OBS_ROOM || kind > GPT_OBS_VEC_HANSEN_GOAL) return fail(GPT_E_ARG, "msrooms: unsupported obs kind");
  const bool hansen = kind == GPT_OBS_HANSEN || kind == GPT_OBS_VEC_HANSEN || kind == GPT_OBS_VEC_HANSEN_GOAL;
  if (hansen && c->rooms_obs_n != 4 && c->rooms_obs_n != 8) return fail(GPT_E_ARG, "msrooms: hansen obs_n must be 4 or 8");
  if (c->env_offset % GPT_ENV_ALIGN != 0) return fail(GPT_E_ARG, "msrooms: env_offset must be a multiple of GPT_ENV_ALIGN");
  if (c->track_stats) return fail(GPT_E_ARG, "msrooms: track_stats is not supported for this family");
  if (c->time_limit >= 0x7F7F7F7E) return fail(GPT_E_ARG, "msrooms: time_limit too large");
  auto interior = [&](int y, int x) { return y > 0 && y < h - 1 && x > 0 && x < w - 1; };
  if (S > 1 && (!interior(c->ms_up_y, c->ms_up_x) || !interior(c->ms_down_y, c->ms_down_x)))
    return fail(GPT_E_ARG, "msrooms: stair cells must lie strictly inside the floor map");
  // 3-D walk grid (msrooms.py:69-90): 0 wall, 1 walkable, 2 stair down (floors 1..), 3 stair up (floors ..S-2)
  std::vector<uint8_t> g(nc, 0);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const bool free_cell = c->rooms_grid[y * w + x] > 0;
      if (free_cell && !interior(y, x)) return fail(GPT_E_ARG, "msrooms: the floor map must have a solid wall border");
      for (int z = 0; z < S; ++z) g[z * fl + y * w + x] = free_cell ? 1 : 0;
    }
  if (S > 1) {
    for (int z = 1; z < S; ++z) g[z * fl + c->ms_down_y * w + c->ms_down_x] = 2;
    for (int z = 0; z + 1 < S; ++z) g[z * fl + c->ms_up_y * w + c->ms_up_x] = 3;
  }
  int gmax = 0;
  for (int i = 0; i < nc; ++i) gmax = g[i] > gmax ? g[i] : gmax;
  // per-cell tables
  std::vector<uint32_t> info(nc), vh((size_t)nc * 2, 0u);
  std::vector<uint16_t> sid(nc, 0), h3(nc, 0), avalid, gvalid;
  const int hn = hansen ? c->rooms_obs_n : 4;
  int run = 0;
  for (int i = 0; i < nc; ++i) {
    const int z = i / fl, y = (i % fl) / w, x = i % w;
    info[i] = (uint32_t)z | ((uint32_t)y << 8) | ((uint32_t)x << 16) | ((uint32_t)g[i] << 24);
    if (g[i] > 0) {
      ++run;
      if (z == 0) avalid.push_back((uint16_t)i);
      if (z == S - 1) gvalid.push_back((uint16_t)i);
    }
    sid[i] = (uint16_t)(run > 0 ? run - 1 : 0);
    uint32_t pow3 = 1, sum = 0;
    for (int j = 0; j < hn; ++j) {
      const int d = hn == 4 ? 2 * j : j, yy = y + kDY[d], xx = x + kDX[d];
      const bool walk = yy >= 0 && yy < h && xx >= 0 && xx < w && g[z * fl + yy * w + xx] > 0;
      const uint32_t digit = walk ? 2u : 0u;   // every walkable cell (values 1..3) aliases to "stairs" = 2 (:154-155)
      sum += digit * pow3;
      pow3 *= 3;
      vh[(size_t)i * 2 + (j >> 2)] |= digit << (8 * (j & 3));
    }
    h3[i] = (uint16_t)sum;
  }
  if (avalid.empty() || gvalid.empty()) return fail(GPT_E_ARG, "msrooms: no walkable cell");
  const bool rgoal = c->ms_goal_cell < 0;
  if (!rgoal && (c->ms_goal_cell >= nc || g[c->ms_goal_cell] == 0)) return fail(GPT_E_ARG, "msrooms: the fixed goal must be a walkable cell");
  env->ms.n_agent = (int32_t)avalid.size();
  env->ms.n_goal = (int32_t)gvalid.size();
  env->ms.n_free = run;
  env->ms.room_n = gmax - 4;   // "number of rooms" as the reference computes it: grid.max() - len(GR_CNST) (:203-206)
  env->ms.n_cells = nc;
  env->ms.obs_bytes = ms_obs_bytes(c);
  if (kind == GPT_OBS_MDP_GOAL && (int64_t)run * run > 0x7FFFFFFF) return fail(GPT_E_ARG, "msrooms: mdp_goal observation overflows int32");
  // move table with the stair teleport folded in (:400-403, :419-428)
  std::vector<uint16_t> move((size_t)nc * 8);
  for (int i = 0; i < nc; ++i) {
    const int z = i / fl, y = (i % fl) / w, x = i % w;
    for (int d = 0; d < 8; ++d) {
      const int yy = y + kDY[d], xx = x + kDX[d];
      const bool ok = g[i] > 0 && yy >= 0 && yy < h && xx >= 0 && xx < w && g[z * fl + yy * w + xx] > 0;
      int nxt = i;
      if (ok) {
        nxt = z * fl + yy * w + xx;
        if (g[nxt] == 3) nxt = (z + 1) * fl + c->ms_down_y * w + c->ms_down_x;       // up the stairs
        else if (g[nxt] == 2) nxt = (z - 1) * fl + c->ms_up_y * w + c->ms_up_x;      // down the stairs
      }
      move[(size_t)i * 8 + d] = (uint16_t)(ok ? nxt : (i | 0x8000));
    }
  }
  // fixed goal: tabulate the observation per agent cell with the same function the device uses for random goals
  MsTables T{info.data(), sid.data(), h3.data(), vh.data()};
  MsObsCfg oc{kind, hn, run, env->ms.room_n};
  std::vector<uint32_t> obstab, moveobs;
  if (!rgoal) {
    const int words = env->ms.obs_bytes > 4 ? 2 : 1;
    obstab.assign((size_t)nc * words, 0u);
    bool fits16 = env->ms.obs_bytes == 4 && nc <= 0x3FFF && kind != GPT_OBS_VEC_HANSEN && kind != GPT_OBS_VEC_HANSEN_GOAL;
    for (int i = 0; i < nc; ++i) {
      uint32_t lo, hi;
      ms_obs(T, oc, (uint32_t)i, (uint32_t)c->ms_goal_cell, lo, hi);
      obstab[(size_t)i * words] = lo;
      if (words == 2) obstab[(size_t)i * 2 + 1] = hi;
      if (g[i] > 0 && lo > 0xFFFFu) fits16 = false;
    }
    if (fits16) {
      moveobs.resize((size_t)nc * 8);
      for (size_t i = 0; i < moveobs.size(); ++i) {
        const uint32_t nxt = move[i] & 0x7FFFu;
        moveobs[i] = nxt | ((int)nxt == c->ms_goal_cell ? 0x4000u : 0u) | (move[i] & 0x8000u) | ((obstab[nxt] & 0xFFFFu) << 16);
      }
      move.clear();
      env->ms.merged = true;
    }
  }
  std::vector<double> thr64(c->rooms_slip_cumsum, c->rooms_slip_cumsum + n * n);
  std::vector<uint32_t> alias = build_slip_alias(n, thr64.data());
  if (!rgoal) { info.clear(); sid.clear(); h3.clear(); vh.clear(); gvalid.clear(); }
  std::vector<uint8_t> blob;
  env->ms.move_off = blob_append(blob, move);
  env->ms.moveobs_off = blob_append(blob, moveobs);
  env->ms.alias_off = blob_append(blob, alias);
  env->ms.thr64_off = blob_append(blob, thr64);
  env->ms.obstab_off = blob_append(blob, obstab);
  env->ms.info_off = blob_append(blob, info);
  env->ms.sid_off = blob_append(blob, sid);
  env->ms.h3_off = blob_append(blob, h3);
  env->ms.vh_off = blob_append(blob, vh);
  env->ms.avalid_off = blob_append(blob, avalid);
  env->ms.gvalid_off = blob_append(blob, gvalid);
  if (int rc = upload_blob(env, blob)) return rc;

  add_array(env, "pos", GPT_ROLE_STATE, GPT_DT_U16, 1);
  if (rgoal) add_array(env, "goal", GPT_ROLE_STATE, GPT_DT_U16, 1);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  if (env->ms.obs_bytes == 4 && !(kind == GPT_OBS_VEC_HANSEN || kind == GPT_OBS_VEC_HANSEN_GOAL))
    add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_I32, 1);
  else
    add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_U8, env->ms.obs_bytes);
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

template <int OB>
static void* ms_pick(bool rgoal, bool merged, bool replay, bool multi, bool devctr) {
  using K = void (*)(const MsParams);
  K k;
  if (devctr && multi) {  // graph mode, fused multi-step launch
    if (rgoal) k = (K)msrooms_step_kernel<OB, true, false, false, true, true>;
    else if (merged && OB == 4) k = (K)msrooms_step_kernel<4, false, true, false, true, true>;
    else k = (K)msrooms_step_kernel<OB, false, false, false, true, true>;
    return (void*)k;
  }
  if (devctr) {  // graph mode: single step, Philox, step counter in device memory
    if (rgoal) k = (K)msrooms_step_kernel<OB, true, false, false, false, true>;
    else if (merged && OB == 4) k = (K)msrooms_step_kernel<4, false, true, false, false, true>;
    else k = (K)msrooms_step_kernel<OB, false, false, false, false, true>;
    return (void*)k;
  }
  if (multi) {  // Philox mode only
    if (rgoal) k = (K)msrooms_step_kernel<OB, true, false, false, true>;
    else if (merged && OB == 4) k = (K)msrooms_step_kernel<4, false, true, false, true>;
    else k = (K)msrooms_step_kernel<OB, false, false, false, true>;
    return (void*)k;
  }
  if (rgoal) k = replay ? (K)msrooms_step_kernel<OB, true, false, true> : (K)msrooms_step_kernel<OB, true, false, false>;
  else if (merged && OB == 4) k = replay ? (K)msrooms_step_kernel<4, false, true, true> : (K)msrooms_step_kernel<4, false, true, false>;
  else k = replay ? (K)msrooms_step_kernel<OB, false, false, true> : (K)msrooms_step_kernel<OB, false, false, false>;
  return (void*)k;
}

// (the fused kernels keep the elapsed counters as biased 16-bit pairs: time limits beyond 32766 step one launch at a time)
bool msrooms_can_fuse(const gpt_env* env) { return env->cfg.rng_mode == GPT_RNG_PHILOX && env->cfg.time_limit >= 0 && env->cfg.time_limit <= 32766; }

int msrooms_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  const bool rgoal = c.ms_goal_cell < 0;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  MsParams P{};
  P.pos = (uint16_t*)env->ptr("pos");
  P.goal = rgoal ? (uint16_t*)env->ptr("goal") : nullptr;
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.actions = (const int8_t*)a.actions;
  P.obs = (uint8_t*)env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.pos || (rgoal && !P.goal) || !P.elapsed || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "msrooms: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "msrooms: actions is NULL");
  const bool reset = a.mode == kModeReset;
  if (reset) {  // reset() = every env truncates: poison `elapsed`, step with any valid action bytes, clear the flags
    cudaError_t e = cudaMemsetAsync(P.elapsed, 0x7F, (size_t)env->capacity * sizeof(int32_t), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(elapsed)");
    P.actions = (const int8_t*)P.terminated;
  }
  P.obs += a.out_row * (int64_t)env->ms.obs_bytes;
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  if (replay) {
    P.rp_u = (const double*)env->ptr("replay_u");
    P.rp_reset_agent = (const int32_t*)env->ptr("replay_reset_agent");
    P.rp_reset_goal = (const int32_t*)env->ptr("replay_reset_goal");
    if (!P.rp_u || !P.rp_reset_agent || !P.rp_reset_goal) return fail(GPT_E_UNBOUND, "msrooms: replay arrays must be bound in replay mode");
  }
  P.blob = env->d_blob;
  P.blob_bytes = env->blob_bytes;
  const auto& L = env->ms;
  P.move_off = L.move_off; P.moveobs_off = L.moveobs_off; P.obstab_off = L.obstab_off; P.info_off = L.info_off;
  P.sid_off = L.sid_off; P.h3_off = L.h3_off; P.vh_off = L.vh_off; P.alias_off = L.alias_off; P.thr64_off = L.thr64_off;
  P.avalid_off = L.avalid_off; P.gvalid_off = L.gvalid_off;
  P.log2n = c.rooms_n_actions == 4 ? 2u : 3u;
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.n_actions = c.rooms_n_actions;
  P.n_agent = L.n_agent;
  P.n_goal = L.n_goal;
  P.time_limit = c.time_limit;
  P.goal_cell = rgoal ? 0xFFFF : c.ms_goal_cell;
  const bool hansen = c.rooms_obs_kind == GPT_OBS_HANSEN || c.rooms_obs_kind == GPT_OBS_VEC_HANSEN || c.rooms_obs_kind == GPT_OBS_VEC_HANSEN_GOAL;
  P.oc = MsObsCfg{c.rooms_obs_kind, hansen ? c.rooms_obs_n : 4, L.n_free, L.room_n};
  P.r_step = c.rooms_step_reward;
  P.r_wall = c.rooms_wall_reward;
  P.r_goal = c.rooms_goal_reward;
  P.rng = make_rng_key(env);

  const int64_t envs_per_cta = (int64_t)kMsThreads * kQuad * kMsQpt;
  const int nblocks = (int)(((int64_t)a.n_tiles * kTileEnvs + envs_per_cta - 1) / envs_per_cta);
  if (nblocks <= 0) return GPT_OK;
  const bool multi = a.n_steps > 1;
  if (multi && replay) return fail(GPT_E_ARG, "msrooms: fused multi-step launches need Philox mode");
  P.n_steps = a.n_steps;
  P.act_stride = env->capacity;
  P.out_stride = a.out_stride_rows;
  const bool devctr = env->graph_mode && !replay;
  P.ctr_ptr = env->d_counter;
  void* k = nullptr;
  switch (L.obs_bytes) {
    case 3: k = ms_pick<3>(rgoal, false, replay, multi, devctr); break;
    case 4: k = ms_pick<4>(rgoal, L.merged, replay, multi, devctr); break;
    case 6: k = ms_pick<6>(rgoal, false, replay, multi, devctr); break;
    case 8: k = ms_pick<8>(rgoal, false, replay, multi, devctr); break;
  }
  if (!k) return fail(GPT_E_ARG, "msrooms: no kernel for this observation layout");
  const size_t smem = env->blob_bytes;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(msrooms)");
  }
  void* args[] = {(void*)&P};
  cudaError_t e = launch_pdl(k, dim3(nblocks), dim3(kMsThreads), smem, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "msrooms_step_kernel launch");
  if (reset) {
    e = cudaMemsetAsync(P.terminated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.truncated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.reward, 0, (size_t)env->capacity * sizeof(float), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reset outputs)");
  }
  return GPT_OK;
}

}  // namespace gpt
