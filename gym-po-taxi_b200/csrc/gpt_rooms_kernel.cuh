// gpt_rooms_kernel.cuh — fused discrete ROOMS / FourRooms step for sm_100a.
//
// One kernel = RoomsEnv.step (reference gym_po/envs/rooms/rooms.py:198-222): ++elapsed, action slip
// (rooms/action_utils.py:73-90), move unless the target cell is a wall (:211-213, :224-226), reward with
// precedence step -> wall -> goal (:215-219), terminated = on goal, truncated = elapsed > limit (:220),
// same-step autoreset goal-then-agent (:191-196) and the observation of the post-reset state
// (rooms/observations.py:44-131 and the table/room obs closures of rooms.py:22-48).
//
// HBM layout (SoA): pos uint16 (flat cell y*W+x) | [goal uint16, random-goal envs only] | elapsed int32 |
// action int8  ->  pos, [goal], elapsed, obs, reward float32, terminated uint8, truncated uint8.
// Thread mapping as in gpt_taxi.cu (warp tile of 512 envs, four quads of 4 consecutive envs per lane).
//
// Static tables, one TMA bulk copy per CTA into shared memory:
//   nb8   uint8 [cells]  bit i = neighbour i (N,NE,E,SE,S,SW,W,NW) is walkable.  Drives BOTH the wall
//                        collision (blocked = bit of the slipped direction clear) and the Hansen obs.
//   room  uint8 [cells], sid uint16[cells] (dense cell id), valid uint16[n_valid] (spawn cells)
//   alias uint2[n*8]     Walker alias table per intended action (Philox-mode slip): {threshold, dir | alias dir << 8};
//                        the draws are Philox4x32-7 blocks, one per quad and step (gpt_common.cuh kStepRounds)
//   thr64 double[n*n]    cumsum(P[a]) exactly as numpy computes it (replay mode compares the recorded u)
//   rows  uint64[H+2*off] walkable-bit rows, padded by the window radius (grid obs only)
#pragma once
#include <utility>
#include "gpt_internal.h"

namespace gpt {

struct RoomsParams {
  uint16_t* pos;
  uint16_t* goal;
  int32_t* elapsed;
  const int8_t* actions;
  void* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  float* ep_return;   // running return per env (track_stats only)
  double* stats;      // device float64[8] (track_stats only)
  int64_t num_envs;   // rows >= num_envs are padding: excluded from the statistics
  const double* rp_u;
  const int32_t* rp_reset_agent;
  const int32_t* rp_reset_goal;
  const uint8_t* blob;
  uint32_t blob_bytes, nb8_off, room_off, sid_off, valid_off, thr64_off, rows_off, stage_off, move_off, obstab_off, alias_off, moveobs_off;
  uint32_t log2n;   // log2(n_actions)
  int64_t env_offset;
  int32_t first_tile, n_tiles, mode;
  int32_t w, n_actions, n_valid, n_rooms, time_limit;
  int32_t hansen_n, grid_n, goal_cell, goal_y, goal_x;
  int32_t obs_tma;   // window obs: the staged tile is written by one TMA bulk store per warp and quad row (16-byte aligned rows)
  FastDiv div_w;
  float r_step, r_wall, r_goal;
  // fused multi-step launch (gpt_step_many, MULTI kernels): n_steps consecutive steps, the state stays in registers
  int32_t n_steps;
  int64_t act_stride;   // bytes between consecutive steps' action rows (= capacity)
  int64_t out_stride;   // rows between consecutive steps' outputs (0 = overwrite in place)
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};

struct RoomsTables {
  const uint8_t* nb8;
  const uint8_t* room;
  const uint16_t* sid;
  const uint16_t* valid;
  const double* thr64;
  const uint64_t* rows;
};

// ordinal direction -> (dy, dx), 2-bit fields holding value+1
__device__ __forceinline__ int dir_dy(uint32_t d) { return (int)((0x1A90u >> (2 * d)) & 3u) - 1; }
__device__ __forceinline__ int dir_dx(uint32_t d) { return (int)((0x01A9u >> (2 * d)) & 3u) - 1; }
// (dy, dx) in {-1,0,1}^2 -> ordinal direction index, 0xF for (0,0)
__device__ __forceinline__ uint32_t dir_of(int dy, int dx) {
  return (uint32_t)((0x3452F6107ull >> (4 * ((dy + 1) * 3 + (dx + 1)))) & 0xFull);
}
// index (ordinal) of the neighbour holding the goal, 0xF if the goal is not adjacent
__device__ __forceinline__ uint32_t goal_dir(int y, int x, int gy, int gx) {
  const int dy = gy - y, dx = gx - x;
  return ((unsigned)(dy + 1) <= 2u && (unsigned)(dx + 1) <= 2u) ? dir_of(dy, dx) : 0xFu;
}
__device__ __forceinline__ uint32_t spread4(uint32_t bits4) { return (bits4 * 0x00204081u) & 0x01010101u; }


// what the observation functions need besides the tables
struct ObsCtx {
  int w, n_rooms, n_valid, hansen_n, gn;
  int fixed_goal, gy, gx;   // fixed goal: its (y, x) — may lie outside the grid (unreachable goal)
  FastDiv div_w;
};

// Observation of one env from its agent cell and goal cell (rooms.py:22-67, observations.py:44-131).
// Scalar kinds return the value in `lo`; byte-vector kinds return up to 8 packed bytes in lo/hi; the
// n x n window is written to `grid_dst` (n*n bytes in shared memory).
template <int OBS, int GRID_N>
__device__ __forceinline__ void cell_obs(const RoomsTables& T, const ObsCtx& C, uint32_t cell, uint32_t gcell,
                                         uint8_t* grid_dst, uint32_t& lo, uint32_t& hi) {
  if constexpr (OBS == GPT_OBS_ROOM) {
    lo = T.room[cell];
  } else if constexpr (OBS == GPT_OBS_ROOM_GOAL) {
    lo = T.room[cell] + (uint32_t)C.n_rooms * T.room[gcell];
  } else if constexpr (OBS == GPT_OBS_MDP) {
    lo = T.sid[cell];
  } else if constexpr (OBS == GPT_OBS_MDP_GOAL) {
    lo = T.sid[cell] + (uint32_t)C.n_valid * T.sid[gcell];
  } else {
    const int y = (int)fdiv(cell, C.div_w), x = (int)cell - y * C.w;
    int gy = C.gy, gx = C.gx;
    if (!C.fixed_goal) {
      gy = (int)fdiv(gcell, C.div_w);
      gx = (int)gcell - gy * C.w;
    }
    if constexpr (OBS == GPT_OBS_VEC_MDP) {
      lo = (uint32_t)y | ((uint32_t)x << 8);
    } else if constexpr (OBS == GPT_OBS_VEC_MDP_GOAL) {
      lo = (uint32_t)y | ((uint32_t)x << 8) | ((uint32_t)gy << 16) | ((uint32_t)gx << 24);
    } else if constexpr (OBS == GPT_OBS_HANSEN) {  // observations.py:44-71
      const uint32_t nb = T.nb8[cell];
      const uint32_t gd = goal_dir(y, x, gy, gx);
      if (C.hansen_n == 8) {
        lo = nb * (gd == 0xFu ? 1u : gd + 1u);
      } else {
        const uint32_t b4 = (nb & 1u) | ((nb >> 1) & 2u) | ((nb >> 2) & 4u) | ((nb >> 3) & 8u);
        lo = b4 * ((gd != 0xFu && !(gd & 1u)) ? (gd >> 1) + 1u : 1u);
      }
    } else if constexpr (OBS == GPT_OBS_VEC_HANSEN || OBS == GPT_OBS_VEC_HANSEN_GOAL) {  // :106-131
      const uint32_t nb = T.nb8[cell];
      uint32_t gd = 0xFu;
      if constexpr (OBS == GPT_OBS_VEC_HANSEN_GOAL) gd = goal_dir(y, x, gy, gx);
      if (C.hansen_n == 8) {
        lo = spread4(nb & 15u);
        hi = spread4(nb >> 4);
        if (gd < 4u) lo = (lo & ~(0xFFu << (8 * gd))) | (2u << (8 * gd));
        else if (gd < 8u) hi = (hi & ~(0xFFu << (8 * (gd - 4)))) | (2u << (8 * (gd - 4)));
      } else {
        const uint32_t b4 = (nb & 1u) | ((nb >> 1) & 2u) | ((nb >> 2) & 4u) | ((nb >> 3) & 8u);
        lo = spread4(b4);
        if (gd != 0xFu && !(gd & 1u)) lo = (lo & ~(0xFFu << (4 * gd))) | (2u << (4 * gd));
      }
    } else if constexpr (OBS == GPT_OBS_GRID) {  // observations.py:74-103
      const int gn = GRID_N > 0 ? GRID_N : C.gn;
      const int off = gn >> 1;
      const uint64_t mask = (1ull << gn) - 1ull;
      if constexpr (GRID_N > 0) {
#pragma unroll
        for (int r = 0; r < GRID_N; ++r) {
          const uint32_t bits = (uint32_t)((T.rows[y + r] >> x) & mask);
#pragma unroll
          for (int c = 0; c < GRID_N; ++c) grid_dst[r * GRID_N + c] = (uint8_t)((bits >> c) & 1u);
        }
      } else {  // run-time window size: rolled loops
#pragma unroll 1
        for (int r = 0; r < gn; ++r) {
          const uint32_t bits = (uint32_t)((T.rows[y + r] >> x) & mask);
#pragma unroll 1
          for (int c = 0; c < gn; ++c) grid_dst[r * gn + c] = (uint8_t)((bits >> c) & 1u);
        }
      }
      const int gr = gy - y + off, gc = gx - x + off;
      if ((unsigned)gr < (unsigned)gn && (unsigned)gc < (unsigned)gn) grid_dst[gr * gn + gc] = 2;
    }
  }
}

// ---- n x n window for a QUAD of envs, assembled as 32-bit words ---------------------------------
// A lane's 4 envs occupy 4*n*n bytes = n*n whole words of the [B,n,n] tensor, so the lane can build
// its slice of the warp's shared-memory tile with n*n conflict-free STS.32 (lane stride n*n words, odd
// for odd n) instead of 4*n*n byte stores.  Which (env, row, col) feeds which byte of which word is
// known at compile time; runs of bytes from one window row are expanded with one multiply (spread4).
template <int N, int W, int B>
__device__ __forceinline__ uint32_t window_word_part(const uint32_t (&bits)[4][N]) {
  if constexpr (B >= 4) {
    return 0u;
  } else {
    constexpr int idx = 4 * W + B, k = idx / (N * N), p = idx % (N * N), r = p / N, c = p % N;
    constexpr int len = (4 - B) < (N - c) ? (4 - B) : (N - c);
    return (spread4((bits[k][r] >> c) & ((1u << len) - 1u)) << (8 * B)) | window_word_part<N, W, B + len>(bits);
  }
}
template <int N, int... W>
__device__ __forceinline__ void window_words(const uint32_t (&bits)[4][N], uint32_t* dst, std::integer_sequence<int, W...>) {
  ((dst[W] = window_word_part<N, W, 0>(bits)), ...);
}
template <int N>
__device__ __forceinline__ void window_quad(const RoomsTables& T, const ObsCtx& C, const uint32_t (&cell)[4], const uint32_t (&gcell)[4],
                                            uint8_t* lane_dst) {
  constexpr int off = N / 2;
  uint32_t bits[4][N];
  int y[4], x[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    y[k] = (int)fdiv(cell[k], C.div_w);
    x[k] = (int)cell[k] - y[k] * C.w;
#pragma unroll
    for (int r = 0; r < N; ++r) bits[k][r] = (uint32_t)(T.rows[y[k] + r] >> x[k]) & ((1u << N) - 1u);
  }
  window_words<N>(bits, reinterpret_cast<uint32_t*>(lane_dst), std::make_integer_sequence<int, N * N>{});
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // the goal cell, if inside the window, reads 2 (same thread: ordered after the word stores)
    int gy = C.gy, gx = C.gx;
    if (!C.fixed_goal) {
      gy = (int)fdiv(gcell[k], C.div_w);
      gx = (int)gcell[k] - gy * C.w;
    }
    const int gr = gy - y[k] + off, gc = gx - x[k] + off;
    if ((unsigned)gr < (unsigned)N && (unsigned)gc < (unsigned)N) lane_dst[k * N * N + gr * N + gc] = 2;
  }
}

// Stores the observations of one quad (4 consecutive envs starting at env index q).  For the window
// obs the warp's 128 envs x n^2 bytes (contiguous in HBM) are streamed out of shared memory.
// `tile_tma` (window obs only): the warp's staged tile leaves with ONE TMA bulk store (lane 0) instead of a loop of
// per-lane 16-byte stores; the caller must call obs_tile_acquire() before the tile is written again and at kernel end.
template <int OBS>
__device__ __forceinline__ void store_obs(void* obs, int64_t q, int64_t warp_quad_base, int hansen_n, int gn, uint32_t lane,
                                          const uint8_t* stage, const uint32_t (&lo)[4], const uint32_t (&hi)[4], bool tile_tma = false) {
  if constexpr (OBS == GPT_OBS_VEC_MDP) {
    st_stream(reinterpret_cast<uint2*>((uint8_t*)obs + q * 2), make_uint2(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16)));
  } else if constexpr (OBS == GPT_OBS_VEC_HANSEN || OBS == GPT_OBS_VEC_HANSEN_GOAL) {
    if (hansen_n == 8) {
      int4* o = reinterpret_cast<int4*>((uint8_t*)obs + q * 8);
      st_stream(o, make_int4((int)lo[0], (int)hi[0], (int)lo[1], (int)hi[1]));
      st_stream(o + 1, make_int4((int)lo[2], (int)hi[2], (int)lo[3], (int)hi[3]));
    } else {
      st_stream(reinterpret_cast<int4*>((uint8_t*)obs + q * 4), make_int4((int)lo[0], (int)lo[1], (int)lo[2], (int)lo[3]));
    }
  } else if constexpr (OBS == GPT_OBS_GRID) {
    if (tile_tma) {
      fence_proxy_async();   // the lanes' tile slices become visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        tma_bulk_s2g((uint8_t*)obs + warp_quad_base * (gn * gn), stage, (uint32_t)(kQuadStride * gn * gn));
        tma_bulk_commit();
      }
    } else {
      __syncwarp();
      const uint32_t vecs = (uint32_t)(kQuadStride * gn * gn) >> 4;
      const int4* src = reinterpret_cast<const int4*>(stage);
      int4* dst = reinterpret_cast<int4*>((uint8_t*)obs + warp_quad_base * (gn * gn));
      for (uint32_t i = lane; i < vecs; i += 32) st_stream(dst + i, src[i]);
      __syncwarp();
    }
  } else {  // scalar int32 obs, or 4 packed bytes per env (VEC_MDP_GOAL)
    st_stream(reinterpret_cast<int4*>((uint8_t*)obs + q * 4), make_int4((int)lo[0], (int)lo[1], (int)lo[2], (int)lo[3]));
  }
}

// Window obs with tile_tma: wait until the warp's previous bulk store has finished READING the staged tile.
__device__ __forceinline__ void obs_tile_acquire(uint32_t lane) {
  if (lane == 0) tma_bulk_wait_read<0>();
  __syncwarp();
}

// Number of quads (4 envs) a thread handles and the CTA size, per observation kind.  Measured on B200
// (Taxi, same access pattern): 2 quads x 128 threads beats 4 x 256 — more, smaller CTAs backfill better.
#ifndef GPT_ROOMS_QPT_SCALAR
#define GPT_ROOMS_QPT_SCALAR 2
#endif
#ifndef GPT_ROOMS_QPT_GRID
#define GPT_ROOMS_QPT_GRID 4
#endif
#ifndef GPT_ROOMS_MINB_SCALAR
#define GPT_ROOMS_MINB_SCALAR 8
#endif
#ifndef GPT_ROOMS_MINB_GRID
#define GPT_ROOMS_MINB_GRID 1
#endif
// measured on B200 (2^22 envs, layout '4'): scalar obs 2 quads/thread; window obs n <= 5: 2 quads and a
// 64-register cap (82 % of HBM peak vs 73 %), n >= 7: 4 quads, no cap (85 % vs 74-81 %)
template <int OBS, int GRID_N> struct RoomsShape { static constexpr int kQpt = GPT_ROOMS_QPT_SCALAR, kThreads = 128, kMinBlocks = GPT_ROOMS_MINB_SCALAR; };
template <int GRID_N> struct RoomsShape<GPT_OBS_GRID, GRID_N> {
  static constexpr bool kSmall = GRID_N > 0 && GRID_N <= 5;
  static constexpr int kQpt = kSmall ? 2 : GPT_ROOMS_QPT_GRID, kThreads = 128, kMinBlocks = kSmall ? 8 : GPT_ROOMS_MINB_GRID;
};

// Rare path, deliberately out of line (one copy per kernel instead of one per unrolled env):
// _reset_some (rooms.py:191-196) — new goal first (random-goal envs), then new agent cell.
template <bool RGOAL, bool REPLAY, bool DEVCTR = false>
__device__ __forceinline__ uint32_t rooms_respawn_inline(const RoomsParams& P, const uint16_t* valid, int64_t env, uint32_t gcell, uint32_t t, uint64_t ctr_dev = 0) {
  uint32_t cell;
  if (REPLAY) {
    if (RGOAL) gcell = (uint32_t)P.rp_reset_goal[env];
    cell = (uint32_t)P.rp_reset_agent[env];
  } else {
    const uint64_t ctr = (DEVCTR ? ctr_dev : (((uint64_t)P.rng.step_hi << 32) | P.rng.step_lo)) + t;   // step index inside a fused launch
    const uint64_t ge = (uint64_t)(P.env_offset + env);
    const uint4 r = philox4x32<kRounds>(make_uint4((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)ctr,
                                                                ((uint32_t)(ctr >> 32) & 0x00FFFFFFu) ^ (1u << 24)), P.rng);
    if (RGOAL) gcell = valid[bounded(r.y, (uint32_t)P.n_valid)];
    cell = valid[bounded(r.x, (uint32_t)P.n_valid)];
  }
  return cell | (gcell << 16);   // values, not references: no local-memory round trip at the call site
}
// out-of-line copy for the single-step kernels
template <bool RGOAL, bool REPLAY>
__device__ __noinline__ uint32_t rooms_respawn(const RoomsParams& P, const uint16_t* valid, int64_t env, uint32_t gcell, uint32_t t) {
  return rooms_respawn_inline<RGOAL, REPLAY>(P, valid, env, gcell, t);
}

// resident CTAs per SM of the fused kernels; measured on B200 (2^22 envs, 8 steps per launch): hansen8 6 -> 331 G,
// 7 -> 346 G, 8 -> 340 G env-steps/s; window 5x5 6 -> 178 G, 7 -> 174 G, 8 -> 169 G
#ifndef GPT_ROOMS_MINB_MULTI
#define GPT_ROOMS_MINB_MULTI (OBS == GPT_OBS_GRID ? 6 : 7)
#endif
// MULTI: gpt_step_many as ONE launch — pos / goal / elapsed are read once, live in registers for P.n_steps steps and
// are written once; per step only the action byte is read and the outputs are written.  Bit-identical to n_steps
// single-step launches (Philox counters = (global env / quad id, first step + t)).
template <int OBS, bool RGOAL, bool REPLAY, int GRID_N, bool STATS, bool MULTI = false, bool DEVCTR = false>
__global__ void __launch_bounds__(RoomsShape<OBS, GRID_N>::kThreads,
                                  STATS ? 1 : (MULTI ? (RoomsShape<OBS, GRID_N>::kMinBlocks > 1 ? GPT_ROOMS_MINB_MULTI : 1) : RoomsShape<OBS, GRID_N>::kMinBlocks))
rooms_step_kernel(const __grid_constant__ RoomsParams P) {
  static_assert(!MULTI || (!REPLAY && !STATS), "fused launches: Philox mode, no in-kernel statistics");
  constexpr int QPT = RoomsShape<OBS, GRID_N>::kQpt;
  constexpr int kEnvsPerWarp = kWarp * kQuad * QPT;
  // fixed goal + non-window obs: the observation is a pure function of the agent cell -> one table lookup
  constexpr bool kObsTable = !RGOAL && OBS != GPT_OBS_GRID;
  // ... and when it also fits 16 bits it rides in the move-table entry: ONE lookup yields next cell, blocked,
  // on-goal and the observation of the next cell
  constexpr bool kMerged = kObsTable && (OBS == GPT_OBS_ROOM || OBS == GPT_OBS_ROOM_GOAL || OBS == GPT_OBS_MDP ||
                                         OBS == GPT_OBS_HANSEN || OBS == GPT_OBS_VEC_MDP);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.blob_bytes, &bar);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t wbase = first + ((int64_t)blockIdx.x * (RoomsShape<OBS, GRID_N>::kThreads / 32) + warp) * kEnvsPerWarp;
  if (wbase >= last) return;
  const int64_t base = wbase + lane * kQuad;
  const uint32_t n = (uint32_t)P.n_actions;
  const uint32_t dir_shift = n == 4 ? 1u : 0u;   // cardinal action i = ordinal direction 2i

  // reset() is not a separate code path: the host poisons `elapsed` so that every env truncates and
  // launches this same kernel (gpt_rooms.cu), which keeps the hot loop free of mode branches.
  pdl_wait();   // the previous step's writes are complete and visible from here on
  uint2 pos4[QPT], goal4[QPT];
  int4 e4[QPT];
  uint32_t a4[QPT];
  float4 ret4[STATS ? QPT : 1];
  EpisodeAcc acc;
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    goal4[j] = make_uint2(0, 0);
    pos4[j] = ld_stream(reinterpret_cast<const uint2*>(P.pos + q));
    if (RGOAL) goal4[j] = ld_stream(reinterpret_cast<const uint2*>(P.goal + q));
    e4[j] = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    a4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + q));
    if constexpr (STATS) ret4[j] = __ldcs(reinterpret_cast<const float4*>(P.ep_return + q));
  }

  stage_tables_wait(&bar);
  RoomsTables T;
  T.nb8 = smem + P.nb8_off;
  T.room = smem + P.room_off;
  T.sid = reinterpret_cast<const uint16_t*>(smem + P.sid_off);
  T.valid = reinterpret_cast<const uint16_t*>(smem + P.valid_off);
  T.thr64 = reinterpret_cast<const double*>(smem + P.thr64_off);
  T.rows = reinterpret_cast<const uint64_t*>(smem + P.rows_off);
  const uint16_t* move = reinterpret_cast<const uint16_t*>(smem + P.move_off);     // [cell*8 + dir] next cell | blocked << 15
  const uint32_t* obstab = reinterpret_cast<const uint32_t*>(smem + P.obstab_off); // [cell] or [cell*2] packed observation
  const uint32_t* moveobs = reinterpret_cast<const uint32_t*>(smem + P.moveobs_off); // [cell*8 + dir] next | goal<<14 | blocked<<15 | obs<<16
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);         // [a*8 + column] {threshold, dir | alias_dir<<8}
  const uint32_t col_shift = 32u - P.log2n;
  const int gn = GRID_N > 0 ? GRID_N : P.grid_n;
  ObsCtx OC;
  OC.w = P.w; OC.n_rooms = P.n_rooms; OC.n_valid = P.n_valid; OC.hansen_n = P.hansen_n; OC.gn = gn; OC.div_w = P.div_w;
  OC.fixed_goal = !RGOAL; OC.gy = P.goal_y; OC.gx = P.goal_x;
  uint8_t* stage = smem + P.stage_off + warp * (uint32_t)(kQuadStride * gn * gn);  // window obs only
  constexpr bool kObs8 = OBS == GPT_OBS_VEC_HANSEN || OBS == GPT_OBS_VEC_HANSEN_GOAL;
  const bool obs_two_words = kObs8 && P.hansen_n == 8;

  // state of the thread's envs, kept in registers (across the steps of a fused launch)
  uint32_t cellq[QPT][4], goalq[QPT][4];
  int32_t evq[QPT][4];
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    cellq[j][0] = pos4[j].x & 0xFFFFu; cellq[j][1] = pos4[j].x >> 16; cellq[j][2] = pos4[j].y & 0xFFFFu; cellq[j][3] = pos4[j].y >> 16;
    goalq[j][0] = goal4[j].x & 0xFFFFu; goalq[j][1] = goal4[j].x >> 16; goalq[j][2] = goal4[j].y & 0xFFFFu; goalq[j][3] = goal4[j].y >> 16;
    evq[j][0] = e4[j].x; evq[j][1] = e4[j].y; evq[j][2] = e4[j].z; evq[j][3] = e4[j].w;
  }
  // MULTI: the four elapsed counters of a quad live in two registers as biased 16-bit pairs, c = elapsed + 0x7FFF -
  // time_limit ([0] = envs 0 | 2, [1] = envs 1 | 3): one packed add per pair and step, bit 15 of a half = "elapsed >
  // time_limit" (rooms.py:220), so the truncated bytes are two shifts and masks and nothing is done per env.  The host
  // fuses only with time_limit <= 32766; an injected counter outside [0, time_limit] behaves like the limit (the env
  // truncates on its next step either way).
  const uint32_t cbias = 0x7FFFu - (uint32_t)P.time_limit;
  uint32_t cntq[MULTI ? QPT : 1][2];
  if constexpr (MULTI) {
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      uint32_t c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) c[k] = (uint32_t)min(max(evq[j][k], 0), P.time_limit) + cbias;
      cntq[j][0] = c[0] | (c[2] << 16);
      cntq[j][1] = c[1] | (c[3] << 16);
    }
  }
  const int32_t n_steps = MULTI ? P.n_steps : 1;
  // DEVCTR (graph mode): the step counter comes from device memory, so that a captured CUDA graph can be replayed
  uint64_t ctr_dev = 0;
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, (uint32_t)n_steps);
  const size_t obs_row = OBS == GPT_OBS_GRID ? (size_t)(gn * gn)
                         : (OBS == GPT_OBS_VEC_MDP ? 2 : ((OBS == GPT_OBS_VEC_HANSEN || OBS == GPT_OBS_VEC_HANSEN_GOAL) ? (size_t)P.hansen_n : 4));
#pragma unroll 1
  for (int32_t t = 0; t < n_steps; ++t) {
  uint32_t a_next[QPT];
  const bool more = MULTI && t + 1 < n_steps;
#pragma unroll
  for (int j = 0; j < QPT; ++j) {   // prefetch the next step's action bytes (the only per-step read)
    a_next[j] = 0u;
    if (more) a_next[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + (int64_t)(t + 1) * P.act_stride + base + j * kQuadStride));
  }
  const int64_t orow = MULTI ? (int64_t)t * P.out_stride : 0;
  const uint64_t ctr = (DEVCTR ? ctr_dev : (((uint64_t)P.rng.step_hi << 32) | P.rng.step_lo)) + (uint32_t)t;   // Philox step counter of this step
  const uint32_t ctr_lo = (uint32_t)ctr, ctr_hi = (uint32_t)(ctr >> 32) & 0x00FFFFFFu;
  // slip draws: one Philox4x32-7 block per quad and step (32 bits per env), the blocks of a quad pair computed together
  static_assert(QPT % 2 == 0, "quads are processed in pairs");
  uint4 slipq[QPT];
  if (!REPLAY) {
#pragma unroll
    for (int j = 0; j < QPT; j += 2) {
      const uint64_t g0 = (uint64_t)(P.env_offset + base + j * kQuadStride) >> 2, g1 = (uint64_t)(P.env_offset + base + (j + 1) * kQuadStride) >> 2;
      slipq[j] = make_uint4((uint32_t)g0, (uint32_t)(g0 >> 32), ctr_lo, ctr_hi);
      slipq[j + 1] = make_uint4((uint32_t)g1, (uint32_t)(g1 >> 32), ctr_lo, ctr_hi);
      philox4x32_x2<kStepRounds>(slipq[j], slipq[j + 1], P.rng);
    }
  }
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    uint32_t (&cellv)[4] = cellq[j];
    uint32_t (&goalv)[4] = goalq[j];
    int32_t (&ev)[4] = evq[j];
    float rv[4];
    uint32_t tw = 0, trw = 0;
    uint32_t mraw[4] = {0, 0, 0, 0}; // merged move-table entries of the quad (fused launches: flag bytes by byte permutes)
    uint32_t o32[4] = {0, 0, 0, 0};  // scalar obs, or packed bytes of the vector obs (lo)
    uint32_t o32b[4] = {0, 0, 0, 0}; // second word for 8-byte vector obs

    const uint32_t slipv[4] = {REPLAY ? 0u : slipq[j].x, REPLAY ? 0u : slipq[j].y, REPLAY ? 0u : slipq[j].z, REPLAY ? 0u : slipq[j].w};

    // ---- transition of the 4 envs: straight-line code, no branches, so the compiler can interleave the
    //      four dependent lookup chains (thresholds -> move table)
    uint32_t again = 0u;   // bit k: env k finished its episode (goal reached or time limit)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t env = q + k;
      const uint32_t gcell = RGOAL ? goalv[k] : (uint32_t)P.goal_cell;
      if constexpr (!MULTI) ev[k] += 1;
      const uint32_t a = ((a4[j] >> (8 * k)) & 0xFFu) & (n - 1u);   // n is 4 or 8; out-of-range bytes wrap
      // slipped action a' = #{j : cumsum(P[a])_j < u}, clamped to n-1   (action_utils.py:84-90)
      uint32_t a2 = 0;
      if (REPLAY) {
        const double u = P.rp_u[env];
        const double* row = T.thr64 + a * n;
        for (uint32_t i = 0; i < n; ++i) a2 += row[i] < u ? 1u : 0u;
        a2 = a2 < n ? a2 : n - 1;
      } else {
        // Walker alias table per intended action: the top log2(n) bits of the draw pick a column, the
        // remaining bits decide between the column's own direction and its alias — one 8-byte lookup and one
        // compare instead of a 3-step threshold search.  Entries are already in ordinal-direction units.
        const uint32_t u = slipv[k];
        const uint2 e = alias[a * 8 + (u >> col_shift)];
        a2 = ((u << P.log2n) < e.x) ? (e.y & 0xFFu) : (e.y >> 8);
      }
      const uint32_t d8 = REPLAY ? (a2 << dir_shift) : a2;
      bool blocked, at_goal;
      if constexpr (kMerged) {
        const uint32_t m = moveobs[cellv[k] * 8 + d8];
        mraw[k] = m;
        cellv[k] = m & 0x3FFFu;
        blocked = (m & 0x8000u) != 0;
        at_goal = (m & 0x4000u) != 0;
        o32[k] = m >> 16;
      } else {
        const uint32_t mv = move[cellv[k] * 8 + d8];   // grid[proposed] == -1 -> stay (rooms.py:212-213, :224-226)
        blocked = (mv & 0x8000u) != 0;
        cellv[k] = mv & 0x7FFFu;
        at_goal = cellv[k] == gcell;                     // (:216)
      }
      rv[k] = at_goal ? P.r_goal : (blocked ? P.r_wall : P.r_step);
      bool trunc = false;
      if constexpr (!MULTI) {
        trunc = ev[k] > P.time_limit;            // (:220)
        tw |= (at_goal ? 1u : 0u) << (8 * k);
        trw |= (trunc ? 1u : 0u) << (8 * k);
        again |= ((at_goal | trunc) ? 1u : 0u) << k;
      } else if constexpr (!kMerged) {
        tw |= (at_goal ? 1u : 0u) << (8 * k);
      }
      goalv[k] = gcell;
      if constexpr (STATS) {
        float& ret = k == 0 ? ret4[j].x : (k == 1 ? ret4[j].y : (k == 2 ? ret4[j].z : ret4[j].w));
        ret += rv[k];
        if (env < P.num_envs) {
          acc.steps += 1.f;
          if (at_goal | trunc) acc.finish(ret, ev[k]);
        }
        ret = (at_goal | trunc) ? 0.f : ret;
      }
    }
    if constexpr (MULTI) {
      // terminated bytes: the on-goal flag (bit 14) of the four merged entries, gathered with byte permutes
      if constexpr (kMerged) {
        const uint32_t hi01 = __byte_perm(mraw[0], mraw[1], 0x0051), hi23 = __byte_perm(mraw[2], mraw[3], 0x0051);
        tw = (__byte_perm(hi01, hi23, 0x5410) >> 6) & 0x01010101u;
      }
      // elapsed += 1 for the four envs; truncated bytes = bit 15 of the biased halves
      cntq[j][0] += 0x00010001u;
      cntq[j][1] += 0x00010001u;
      trw = ((cntq[j][0] >> 15) & 0x00010001u) | ((cntq[j][1] >> 7) & 0x01000100u);
      again = tw | trw;   // one byte per env
    }
    // ---- rare: respawn finished envs (one divergence point per quad, not per env) ----
    if (again) {
#pragma unroll 1
      for (uint32_t m = again; m; m &= m - 1) {
        const int k = MULTI ? (__ffs(m) - 1) >> 3 : __ffs(m) - 1;
        uint32_t g = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) g = i == k ? goalv[i] : g;
        // fused launches inline it: a CALL would wait for the in-flight action prefetch
        const uint32_t fresh = (MULTI || DEVCTR || GPT_RESPAWN_INLINE_SINGLE) ? rooms_respawn_inline<RGOAL, REPLAY, DEVCTR>(P, T.valid, q + k, g, (uint32_t)t, ctr_dev) : rooms_respawn<RGOAL, REPLAY>(P, T.valid, q + k, g, (uint32_t)t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i == k) {
            cellv[i] = fresh & 0xFFFFu;
            goalv[i] = fresh >> 16;
            ev[i] = 0;
            if constexpr (kMerged) o32[i] = obstab[fresh & 0xFFFFu];
          }
        }
        if constexpr (MULTI) {   // elapsed = 0: the env's counter half goes back to the bias
          const uint32_t half = (k & 2) ? 0xFFFF0000u : 0x0000FFFFu, fresh_c = (cbias | (cbias << 16)) & half;
          if (k & 1) cntq[j][1] = (cntq[j][1] & ~half) | fresh_c;
          else cntq[j][0] = (cntq[j][0] & ~half) | fresh_c;
        }
      }
    }
    // ---- observation of the (post-reset) state ------------------------------------------------
    if constexpr (OBS == GPT_OBS_GRID) {
      if (P.obs_tma) obs_tile_acquire(lane);   // the previous quad row's bulk store has read the tile
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if constexpr (kMerged) {
        // observation already taken from the move-table entry (or refreshed by the respawn above)
      } else if constexpr (kObsTable) {
        if (obs_two_words) {
          const uint2 o = reinterpret_cast<const uint2*>(obstab)[cellv[k]];
          o32[k] = o.x;
          o32b[k] = o.y;
        } else {
          o32[k] = obstab[cellv[k]];
        }
      } else if constexpr (!(OBS == GPT_OBS_GRID && GRID_N > 0)) {
        cell_obs<OBS, GRID_N>(T, OC, cellv[k], goalv[k], stage + (uint32_t)(lane * kQuad + k) * (uint32_t)(gn * gn), o32[k], o32b[k]);
      }
    }
    if constexpr (OBS == GPT_OBS_GRID && GRID_N > 0)
      window_quad<GRID_N>(T, OC, cellv, goalv, stage + (uint32_t)(lane * kQuad) * (uint32_t)(GRID_N * GRID_N));

    // ---- stores ------------------------------------------------------------------------------
    st_stream(reinterpret_cast<float4*>(P.reward + orow + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + orow + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + orow + q), trw);
    store_obs<OBS>((uint8_t*)P.obs + (size_t)orow * obs_row, q, wbase + j * kQuadStride, P.hansen_n, gn, lane, stage, o32, o32b, P.obs_tma != 0);
    if constexpr (STATS) st_stream(reinterpret_cast<float4*>(P.ep_return + q), ret4[j]);
    a4[j] = a_next[j];
  }
  }  // steps of a fused launch
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
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
  if constexpr (STATS) acc.flush(P.stats);
  if constexpr (OBS == GPT_OBS_GRID) {
    if (P.obs_tma) obs_tile_acquire(lane);   // shared memory must outlive the last bulk store's read
  }
}

template <int OBS, int GRID_N, bool STATS>
static void* pick_rr2(bool rgoal, bool replay) {
  using K = void (*)(const RoomsParams);
  K k = rgoal ? (replay ? (K)rooms_step_kernel<OBS, true, true, GRID_N, STATS> : (K)rooms_step_kernel<OBS, true, false, GRID_N, STATS>)
              : (replay ? (K)rooms_step_kernel<OBS, false, true, GRID_N, STATS> : (K)rooms_step_kernel<OBS, false, false, GRID_N, STATS>);
  return (void*)k;
}
// variant: 0 = plain single step, 1 = single step with in-kernel statistics, 2 = fused multi-step (Philox mode),
// 3 = single step with the device-resident step counter (graph mode, Philox), 4 = fused multi-step in graph mode
template <int OBS, int GRID_N>
static void* pick_rr(bool rgoal, bool replay, int variant) {
  using K = void (*)(const RoomsParams);
  if (variant == 4)
    return replay ? nullptr
                  : (void*)(rgoal ? (K)rooms_step_kernel<OBS, true, false, GRID_N, false, true, true> : (K)rooms_step_kernel<OBS, false, false, GRID_N, false, true, true>);
  if (variant == 3)
    return replay ? nullptr
                  : (void*)(rgoal ? (K)rooms_step_kernel<OBS, true, false, GRID_N, false, false, true> : (K)rooms_step_kernel<OBS, false, false, GRID_N, false, false, true>);
  if (variant == 2)
    return replay ? nullptr
                  : (void*)(rgoal ? (K)rooms_step_kernel<OBS, true, false, GRID_N, false, true> : (K)rooms_step_kernel<OBS, false, false, GRID_N, false, true>);
  return variant == 1 ? pick_rr2<OBS, GRID_N, true>(rgoal, replay) : pick_rr2<OBS, GRID_N, false>(rgoal, replay);
}

// kernel instantiations live in gpt_rooms_k*.cu so that they compile in parallel
void* rooms_pick_table(int obs, bool rgoal, bool replay, int variant);   // ROOM, ROOM_GOAL, MDP, MDP_GOAL
void* rooms_pick_vec(int obs, bool rgoal, bool replay, int variant);     // VEC_MDP, VEC_MDP_GOAL, HANSEN
void* rooms_pick_vhansen(int obs, bool rgoal, bool replay, int variant); // VEC_HANSEN, VEC_HANSEN_GOAL
void* rooms_pick_grid_small(int n, bool rgoal, bool replay, int variant);  // 3, 5
void* rooms_pick_grid_large(int n, bool rgoal, bool replay, int variant);  // 7, 9
void* rooms_pick_grid_any(bool rgoal, bool replay, int variant);           // run-time n <= 15

}  // namespace gpt
