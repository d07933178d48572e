// gpt_taxi.cu — fused Taxi POMDP step for sm_100a.
//
// One kernel = the whole of TaxiVecEnv.step (reference gym_po/envs/extended_taxi.py:244-287):
// ++elapsed, decode, move / pickup / dropoff, reward, terminated, truncated, passenger respawn
// (:354-364), full autoreset (:344-352) and the observation of the post-reset state (:366-372).
//
// Data layout in HBM (SoA, `capacity` rows, one stream per field):
//     s int32 | elapsed int32 | ndrop uint8 | action int8  ->  s, elapsed, ndrop, obs int32,
//     reward float32, terminated uint8, truncated uint8                       (29 B per env-step)
// Thread mapping: a warp owns a tile of 512 consecutive envs; lane l handles four "quads" of four
// consecutive envs at tile + j*128 + 4*l (j = 0..3), so every warp-wide access is one fully used
// 512 B (int32/float32 streams) or 128 B (byte streams) contiguous segment.
//
// Kernels implementing the same step:
//   * taxi_multi_tma_kernel / taxi_table_multi_kernel: T consecutive steps in ONE launch (gpt_step_many), state in
//     registers; the first moves its per-step I/O with the TMA engine (bulk-copied action rows, outputs staged in shared
//     memory and written by four bulk stores per CTA and step) and is the default, the second loads and stores per thread
//     (track_stats, unaligned streams, A/B runs).  See the comments at the kernels.
//   * taxi_table_kernel (used when ns <= 2048, i.e. both reference maps): the complete (state, action) ->
//     (next state, delivered?, illegal?) relation is tabulated on the host (ns x 6 uint16, 6 KB for the
//     5x5 map) and lives in shared memory, so the per-env main path is ONE data-dependent LDS plus ~20
//     ALU instructions.  The rare branches (autoreset ~0.5 % of env-steps, passenger respawn) are NOT in
//     the unrolled main path: a thread records them in a 32-bit mask and patches the affected envs with
//     scalar stores in a compact loop after its vector stores.
//   * taxi_arith_kernel: decode / wall-bit move / pickup-dropoff arithmetic, any map with ns < 65536.
//
// Static tables are packed on the host into one blob and staged into shared memory by
// one TMA bulk copy per CTA:
//     celltab uint16[cells x REP]: bits 0-3 = wall bits N,S,W,E around the cell (1 = blocked; this is
//         hansen_encodings, extended_taxi.py:102-114, which also decides motion — SURVEY.md A.1),
//         bits 8-15 = index of the named location on that cell (0xFF none).  Replicated per lane
//         (REP = 32) when small so that the data-dependent lookups are bank-conflict free.
//     alias uint2[n_valid]: Philox-mode autoreset sampler — Walker alias table of the law of
//         argmax(multinomial(ns, uniform over the valid states)) (extended_taxi.py:348-350), built on the host from
//         the thresholds in gpt_config.taxi_reset_cdf: {threshold, state | alias state << 16} per column.
#include <cstdlib>

#include "gpt_internal.h"

namespace gpt {

struct TaxiParams {
  int32_t* s;
  int32_t* elapsed;
  uint8_t* ndrop;
  const int8_t* actions;
  int32_t* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  float* ep_return;   // running return per env (track_stats only, else NULL)
  double* stats;      // device float64[8] (track_stats only, else NULL)
  int64_t num_envs;   // rows >= num_envs are padding: excluded from the statistics
  const int32_t* rp_reset_state;
  const int8_t* rp_new_p;
  const int8_t* rp_new_d;
  const uint8_t* blob;
  uint32_t blob_bytes, rep_shift;
  uint32_t trans_off, hobs_off;   // fused kernel: 32-bit transition table / per-state observation table offsets in the blob
  uint32_t trans16_off, alias_off, single_bytes;   // single-step kernel: compact table, reset alias table, bytes to stage
  FastDiv div_pd;                 // divide by (nlocs+1)*nlocs
  int64_t env_offset;
  int32_t first_tile, n_tiles;
  int32_t cols, nlocs, n_dropoffs, time_limit, n_valid, mode;
  FastDiv div_nlocs, div_nlocs1;
  float r_goal, r_bad, r_any;
  RngKey rng;
  const uint64_t* ctr_ptr;   // graph mode (DEVCTR kernels): device-resident Philox step counter, else unused
};

struct TaxiTables {
  const uint16_t* cell;
  const uint2* alias;
  uint32_t rep_shift, rep_lane;
  __device__ __forceinline__ uint32_t lookup(uint32_t c) const { return cell[(c << rep_shift) | rep_lane]; }
};

template <bool HANSEN, bool REPLAY>
__device__ __forceinline__ void taxi_env(const TaxiParams& P, const TaxiTables& T, int64_t env, bool reset_all,
                                         int32_t& s, int32_t& e, uint32_t& nd, uint32_t a, int32_t& obs, float& rew,
                                         uint32_t& term, uint32_t& trunc) {
  const uint32_t nlocs = (uint32_t)P.nlocs;
  uint32_t cell, p, d;
  bool full_reset = reset_all;
  bool respawn = false;
  rew = 0.f;
  term = trunc = 0u;
  uint32_t ent = 0;
  if (!reset_all) {
    e += 1;
    // decode s = ((cell*(nlocs+1)) + p)*nlocs + d           (extended_taxi.py:84-94)
    const uint32_t t = fdiv((uint32_t)s, P.div_nlocs);
    d = (uint32_t)s - t * nlocs;
    cell = fdiv(t, P.div_nlocs1);
    p = t - cell * (nlocs + 1);
    // move N/S/W/E unless the wall bit of that side is set    (:248-260)
    ent = T.lookup(cell);
    if (a < 4u && !((ent >> a) & 1u)) {
      const int32_t step = (a & 2u) ? 1 : P.cols;
      cell = (uint32_t)((int32_t)cell + ((a & 1u) ? step : -step));
      ent = T.lookup(cell);
    }
    // pickup / dropoff                                       (:262-275)
    const uint32_t here = ent >> 8;  // named location on this cell, 0xFF if none
    const bool act = a == 4u;
    const bool goal = act && p == nlocs && here == d;
    nd += goal ? 1u : 0u;
    const bool pickup = act && p < nlocs && here == p;
    p = pickup ? nlocs : p;
    const bool bad = act && !goal && !pickup;
    rew = goal ? P.r_goal : (bad ? P.r_bad : P.r_any);
    term = nd == (uint32_t)P.n_dropoffs;   // (:276-279)
    trunc = e > P.time_limit;
    full_reset = term | trunc;
    respawn = goal && !full_reset;         // (:283-285)
  }
  if (respawn | full_reset) {  // rare, divergent
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (!REPLAY) rnd = env_random(P.rng, (uint64_t)(P.env_offset + env), 0u);
    if (full_reset) {  // _reset_mask (:344-352)
      uint32_t fresh;
      if (REPLAY) {
        fresh = (uint32_t)P.rp_reset_state[env];
      } else {  // the law of argmax(multinomial(ns, uniform over valid states)) through its alias table (as taxi_fix_inline)
        const uint64_t w = (uint64_t)rnd.x * (uint32_t)P.n_valid;
        const uint2 e = T.alias[(uint32_t)(w >> 32)];
        fresh = ((uint32_t)w < e.x) ? (e.y & 0xFFFFu) : (e.y >> 16);
      }
      const uint32_t t = fdiv(fresh, P.div_nlocs);
      d = fresh - t * nlocs;
      cell = fdiv(t, P.div_nlocs1);
      p = t - cell * (nlocs + 1);
      e = 0;
      nd = 0;
      if (HANSEN) ent = T.lookup(cell);
    } else {  // _reset_passenger_and_destination (:354-364): p uniform, d uniform over the others
      if (REPLAY) {
        p = (uint32_t)P.rp_new_p[env];
        d = (uint32_t)P.rp_new_d[env];
      } else {
        p = bounded(rnd.y, nlocs);
        d = bounded(rnd.z, nlocs - 1);
        d += d >= p ? 1u : 0u;
      }
    }
  }
  s = (int32_t)((cell * (nlocs + 1) + p) * nlocs + d);   // encode (:97-99)
  obs = HANSEN ? (int32_t)(((ent & 15u) * (nlocs + 1) + p) * nlocs + d) : s;   // (:366-372)
}

template <bool HANSEN, bool REPLAY>
__global__ void __launch_bounds__(256) taxi_arith_kernel(const __grid_constant__ TaxiParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.blob_bytes, &bar);
  pdl_wait();

  const uint32_t lane = threadIdx.x & 31u;
  const int32_t tile = P.first_tile + (int32_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
  if (tile >= P.first_tile + P.n_tiles) return;
  const int64_t base = (int64_t)tile * kTileEnvs + lane * kQuad;
  const bool reset_all = P.mode == kModeReset;

  int4 s4[kQuadsPerThread], e4[kQuadsPerThread];
  uint32_t nd4[kQuadsPerThread], a4[kQuadsPerThread];
#pragma unroll
  for (int j = 0; j < kQuadsPerThread; ++j) {
    const int64_t q = base + j * kQuadStride;
    if (!reset_all) {
      s4[j] = ld_stream(reinterpret_cast<const int4*>(P.s + q));
      e4[j] = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
      nd4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.ndrop + q));
      a4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + q));
    } else {
      s4[j] = e4[j] = make_int4(0, 0, 0, 0);
      nd4[j] = a4[j] = 0u;
    }
  }

  stage_tables_wait(&bar);
  TaxiTables T;
  T.cell = reinterpret_cast<const uint16_t*>(smem);
  T.alias = reinterpret_cast<const uint2*>(smem + P.alias_off);
  T.rep_shift = P.rep_shift;
  T.rep_lane = P.rep_shift ? lane : 0u;

#pragma unroll
  for (int j = 0; j < kQuadsPerThread; ++j) {
    const int64_t q = base + j * kQuadStride;
    int32_t sv[4] = {s4[j].x, s4[j].y, s4[j].z, s4[j].w};
    int32_t ev[4] = {e4[j].x, e4[j].y, e4[j].z, e4[j].w};
    int32_t ov[4];
    float rv[4];
    uint32_t ndw = 0, tw = 0, trw = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t nd = (nd4[j] >> (8 * k)) & 0xFFu;
      const uint32_t a = (a4[j] >> (8 * k)) & 0xFFu;
      uint32_t term, trunc;
      taxi_env<HANSEN, REPLAY>(P, T, q + k, reset_all, sv[k], ev[k], nd, a, ov[k], rv[k], term, trunc);
      ndw |= (nd & 0xFFu) << (8 * k);
      tw |= term << (8 * k);
      trw |= trunc << (8 * k);
    }
    st_stream(reinterpret_cast<int4*>(P.s + q), make_int4(sv[0], sv[1], sv[2], sv[3]));
    st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.ndrop + q), ndw);
    st_stream(reinterpret_cast<int4*>(P.obs + q), make_int4(ov[0], ov[1], ov[2], ov[3]));
    if (!reset_all) {
      st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
      st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
      st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
    }
  }
}


// ------------------------------------------------------------------------------------------
// table-driven kernel (both reference maps): single steps AND fused multi-step launches
// ------------------------------------------------------------------------------------------
// trans: [state][8 actions] uint32, one 32-byte row per state.  An entry holds everything the step needs:
//   bits 31-16  observation of the next state (state id, or the Hansen re-encoding, extended_taxi.py:366-372)
//   bits 15-5   byte offset of the next state's row (the kernel keeps the state as this offset, so the next lookup
//               address is one 3-input add: table base + row offset + 4*action)
//   bit 1 / 0   illegal pickup-dropoff / delivery
// Action bytes 5..7 are no-op columns; larger bytes wrap modulo 8 (the reference raises IndexError).
constexpr int kRowShift = 5;
constexpr uint32_t kTransGoal = 1u, kTransBad = 2u, kTransRow = 0xFFE0u;
constexpr int kTransCols = 8;
constexpr int64_t kTableMaxStates = 65536 >> kRowShift;   // row offsets must fit 16 bits

// T consecutive steps from an action stream [T, capacity] in ONE launch (gpt_step_many): the state (s, elapsed,
// ndrop) is read once, lives in registers for the T steps and is written once; per step only the action byte is
// read and the outputs are written — 11 B per env-step + 18 B per env per launch instead of 29 B.  Results are
// bit-identical to T single-step launches: Philox counters are (global env id, first step + t).
struct TaxiMultiParams {
  TaxiParams p;
  int32_t n_steps;
  int64_t act_stride;   // bytes between consecutive steps' action rows (= capacity)
  int64_t out_stride;   // rows between consecutive steps' outputs (0 = overwrite in place)
};

// Rare branch: full reset (extended_taxi.py:344-352) or passenger respawn
// (:354-364) -> new state id.  `t` = step index inside a fused launch.
template <bool REPLAY, bool DEVCTR = false>
__device__ __forceinline__ uint32_t taxi_fix_inline(const TaxiParams& P, const uint2* alias, int64_t env, uint32_t t, uint32_t cur, bool full, uint64_t ctr_dev = 0) {
  if (REPLAY) {
    if (full) return (uint32_t)P.rp_reset_state[env];
    const uint32_t cell = fdiv(cur, P.div_pd);
    return (cell * (uint32_t)(P.nlocs + 1) + (uint32_t)P.rp_new_p[env]) * (uint32_t)P.nlocs + (uint32_t)P.rp_new_d[env];
  }
  const uint64_t ctr = (DEVCTR ? ctr_dev : (((uint64_t)P.rng.step_hi << 32) | P.rng.step_lo)) + t;   // Philox step counter of this step
  const uint64_t ge = (uint64_t)(P.env_offset + env);
  const uint4 rnd = philox4x32<kRounds>(make_uint4((uint32_t)ge, (uint32_t)(ge >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32) & 0x00FFFFFFu), P.rng);
  if (full) {  // the law of argmax(multinomial(ns, uniform over valid states)), sampled through its alias table
    const uint64_t w = (uint64_t)rnd.x * (uint32_t)P.n_valid;
    const uint2 e = alias[(uint32_t)(w >> 32)];
    return ((uint32_t)w < e.x) ? (e.y & 0xFFFFu) : (e.y >> 16);
  }
  const uint32_t cell = fdiv(cur, P.div_pd);   // taxi cell kept, p uniform, d uniform over the other locations
  const uint32_t p = bounded(rnd.y, (uint32_t)P.nlocs);
  uint32_t d = bounded(rnd.z, (uint32_t)P.nlocs - 1);
  d += d >= p ? 1u : 0u;
  return (cell * (uint32_t)(P.nlocs + 1) + p) * (uint32_t)P.nlocs + d;
}
// (inlined in the single-step kernel too: measured 253.9 -> 256.4 G against an out-of-line call)
#ifndef GPT_TAXI_MINB_MULTI
#define GPT_TAXI_MINB_MULTI 6   // 72 registers; measured on B200 (2^22 envs, 8 steps per launch): 4 -> 412 G, 6 -> 417 G, 7 -> 418 G
#endif
// ONE: num_passengers == 1 (the default).  A delivery then always terminates the episode, so `terminated` is the
// delivered bit of the table entry, there is no passenger respawn and the dropoff counter is 0 at every step
// boundary: the kernel neither loads nor tracks it and stores 0 (a counter injected with set_state is ignored here;
// the single-step kernel keeps the reference's arithmetic).
template <bool STATS, int QPT, int THREADS, bool ONE, bool DEVCTR = false>
__global__ void __launch_bounds__(THREADS, STATS ? 1 : GPT_TAXI_MINB_MULTI * 128 / THREADS) taxi_table_multi_kernel(const __grid_constant__ TaxiMultiParams M) {
  constexpr bool REPLAY = false;   // replayed draws are per step: replay mode uses the single-step kernel
  const TaxiParams& P = M.p;
  const int32_t n_steps = M.n_steps;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.blob_bytes, &bar);

  constexpr int kEnvsPerWarp = kWarp * kQuad * QPT;
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  // reset() is not a separate code path: the host poisons `elapsed` so that every env truncates and
  // launches this same kernel (taxi_launch), which keeps the hot loop free of mode branches.
  const uint8_t* trans = smem + P.trans_off;
  const uint16_t* hobs = reinterpret_cast<const uint16_t*>(smem + P.hobs_off);   // observation per state id (fix path)
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);
  EpisodeAcc acc;
  const int64_t wtile = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const int64_t base = first + wtile * kEnvsPerWarp + lane * kQuad;
  if (base >= last) return;
  pdl_wait();   // the previous launch's writes are complete and visible from here on
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory, advanced by n_steps for the next launch
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, (uint32_t)n_steps);
  uint32_t off[QPT][4], ndv[QPT][4], a4[QPT];   // off = state id << kRowShift
  int32_t ev[QPT][4];
  float ret[STATS ? QPT : 1][4];
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    const int4 s4 = ld_stream(reinterpret_cast<const int4*>(P.s + q));
    const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    const uint32_t nd4 = ONE ? 0u : ld_stream(reinterpret_cast<const uint32_t*>(P.ndrop + q));
    a4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + q));
    off[j][0] = (uint32_t)s4.x << kRowShift; off[j][1] = (uint32_t)s4.y << kRowShift;
    off[j][2] = (uint32_t)s4.z << kRowShift; off[j][3] = (uint32_t)s4.w << kRowShift;
    ev[j][0] = e4.x; ev[j][1] = e4.y; ev[j][2] = e4.z; ev[j][3] = e4.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) ndv[j][k] = (nd4 >> (8 * k)) & 0xFFu;
    if constexpr (STATS) {
      const float4 r4 = __ldcs(reinterpret_cast<const float4*>(P.ep_return + q));
      ret[j][0] = r4.x; ret[j][1] = r4.y; ret[j][2] = r4.z; ret[j][3] = r4.w;
    }
  }
  uint32_t a_next[QPT];
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    a_next[j] = 0u;
    if (n_steps > 1) a_next[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + M.act_stride + base + j * kQuadStride));
  }
  stage_tables_wait(&bar);

#pragma unroll 1
  for (int32_t t = 0; t < n_steps; ++t) {
    // prefetch the action bytes (the only per-step read) TWO steps ahead: under the write-dominated traffic of this
    // kernel a load takes longer than one loop iteration to come back
    uint32_t a_next2[QPT];
    const bool more = t + 2 < n_steps;
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      a_next2[j] = 0u;
      if (more) a_next2[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + (int64_t)(t + 2) * M.act_stride + base + j * kQuadStride));
    }
    const int64_t orow = (int64_t)t * M.out_stride;
    uint32_t fix_done = 0, fix_goal = 0;   // bit 8k + j: env k of quad j finished its episode / delivered a passenger
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + j * kQuadStride;
      float rv[4];
      int32_t ov[4];
      uint32_t tw = 0, trw = 0, gw = 0;   // terminated / truncated / delivered, one byte per env
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t aoff = (k == 0 ? (a4[j] << 2) : (a4[j] >> (8 * k - 2))) & 0x1Cu;   // 4 * (action & 7)
        const uint32_t ent = *reinterpret_cast<const uint32_t*>(trans + off[j][k] + aoff);
        const uint32_t goal = ent & kTransGoal;
        if constexpr (!ONE) ndv[j][k] += goal;
        off[j][k] = ent & kTransRow;
        ov[k] = (int32_t)(ent >> 16);
        ev[j][k] += 1;
        rv[k] = goal ? P.r_goal : ((ent & kTransBad) ? P.r_bad : P.r_any);
        const uint32_t term = ONE ? goal : (uint32_t)(ndv[j][k] == (uint32_t)P.n_dropoffs);      // (:276-279)
        const uint32_t trunc = ev[j][k] > P.time_limit;
        tw |= term << (8 * k);
        trw |= trunc << (8 * k);
        if constexpr (!ONE) gw |= goal << (8 * k);
        if constexpr (STATS) {
          ret[j][k] += rv[k];
          if (q + k < P.num_envs) {
            acc.steps += 1.f;
            if (term | trunc) acc.finish(ret[j][k], ev[j][k]);
          }
          ret[j][k] = (term | trunc) ? 0.f : ret[j][k];
        }
      }
      // the vector stores go out now; the rare envs are fixed below, ONCE per step for all of the thread's envs (their
      // observation is patched with a scalar store: same thread, same address, program order)
      fix_done |= ((tw | trw) & 0x01010101u) << j;
      fix_goal |= (gw & 0x01010101u) << j;
      st_stream(reinterpret_cast<int4*>(P.obs + orow + q), make_int4(ov[0], ov[1], ov[2], ov[3]));
      st_stream(reinterpret_cast<float4*>(P.reward + orow + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
      st_stream(reinterpret_cast<uint32_t*>(P.terminated + orow + q), tw);
      st_stream(reinterpret_cast<uint32_t*>(P.truncated + orow + q), trw);
      a4[j] = a_next[j];
      a_next[j] = a_next2[j];
    }
    if (fix_done | fix_goal) {  // rare: full reset of finished envs, passenger respawn after a delivery (:283-286)
#pragma unroll 1
      for (uint32_t m = fix_done | fix_goal; m; m &= m - 1) {
        const int bit = __ffs(m) - 1, k = bit >> 3, j = bit & 7;   // bit 8k + j <-> quad j, env k
        const bool full = (fix_done >> bit) & 1u;
        const int idx = j * 4 + k;
        uint32_t cur = 0;
#pragma unroll
        for (int i = 0; i < 4 * QPT; ++i) cur = i == idx ? off[i >> 2][i & 3] : cur;
        const int64_t env = base + j * kQuadStride + k;
        // inlined: a CALL here would wait for the in-flight action prefetch (ncu: 20 % of all stall samples)
        const uint32_t fresh = taxi_fix_inline<REPLAY, DEVCTR>(P, alias, env, (uint32_t)t, cur >> kRowShift, full, ctr_dev);
        P.obs[orow + env] = (int32_t)hobs[fresh];
#pragma unroll
        for (int i = 0; i < 4 * QPT; ++i) {
          if (i == idx) {
            off[i >> 2][i & 3] = fresh << kRowShift;
            ev[i >> 2][i & 3] = full ? 0 : ev[i >> 2][i & 3];
            if constexpr (!ONE) ndv[i >> 2][i & 3] = full ? 0u : ndv[i >> 2][i & 3];
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    st_stream(reinterpret_cast<int4*>(P.s + q), make_int4((int)(off[j][0] >> kRowShift), (int)(off[j][1] >> kRowShift),
                                                          (int)(off[j][2] >> kRowShift), (int)(off[j][3] >> kRowShift)));
    st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[j][0], ev[j][1], ev[j][2], ev[j][3]));
    st_stream(reinterpret_cast<uint32_t*>(P.ndrop + q),
              ONE ? 0u : ((ndv[j][0] & 0xFFu) | ((ndv[j][1] & 0xFFu) << 8) | ((ndv[j][2] & 0xFFu) << 16) | ((ndv[j][3] & 0xFFu) << 24)));
    if constexpr (STATS) st_stream(reinterpret_cast<float4*>(P.ep_return + q), make_float4(ret[j][0], ret[j][1], ret[j][2], ret[j][3]));
  }
  if constexpr (STATS) acc.flush(P.stats);
}


// ---- fused multi-step launch with TMA I/O (the default fused path) ----------------------------------------------------
// Same step as taxi_table_multi_kernel, bit-identical results, but the per-step I/O is moved by the TMA engine instead of
// per-thread loads and stores: the CTA's action rows [t][1024 envs] are fetched by bulk copies into a ring of
// kTmaActRows shared-memory rows (one mbarrier per row, refilled as rows are consumed), and every step's outputs are
// staged in shared memory (obs 4 KB | reward 4 KB | terminated 1 KB | truncated 1 KB) and written with four bulk stores
// per CTA and step (double-buffered; cp.async.bulk.wait_group.read guards the reuse).  On the rollout access pattern —
// 86 % writes into 4 output streams x T slots — this moves 6.2 TB/s where per-thread stores reach 5.6 TB/s
// (scripts/stream_pattern_probe4.cu).  One CTA barrier per step.
constexpr int kT16RowShift = 4;                // 16-bit table of the TMA kernel: 8 actions x 2 B per state
#ifndef GPT_TAXI_TMA_THREADS
#define GPT_TAXI_TMA_THREADS 128
#endif
constexpr int kTmaThreads = GPT_TAXI_TMA_THREADS;
constexpr int kTmaEnvs = kTmaThreads * 8;      // envs per CTA: threads x 2 quads x 4 envs
constexpr int kTmaStage = kTmaEnvs * 10;       // staging bytes per buffer
#ifndef GPT_TAXI_TMA_BUFS
#define GPT_TAXI_TMA_BUFS 2
#endif
#ifndef GPT_TAXI_TMA_ACT_ROWS
#define GPT_TAXI_TMA_ACT_ROWS 4   // with the 32-bit table (43.5 KB per CTA on the 5x5 map): 8 rows = 4 CTAs/SM 91.3 us, 4 rows = 5 CTAs/SM 87.1 us per 10-step launch
#endif
constexpr int kTmaBufs = GPT_TAXI_TMA_BUFS;          // power of two
constexpr int kTmaActRows = GPT_TAXI_TMA_ACT_ROWS;   // power of two
#ifndef GPT_TAXI_MINB_TMA
#define GPT_TAXI_MINB_TMA 6
#endif
struct TaxiTmaParams {
  TaxiMultiParams m;
  uint32_t tab_bytes;     // bytes of the blob's part this kernel reads (hobs | alias | 16-bit table), staged at shared offset 0
  uint32_t stage_off;     // shared-memory offset of the staging buffers (128-byte aligned), the action ring follows
};

// 16-bit table: entries [state][8 actions] (16-byte rows; the state is kept as the row's byte offset): next state id << 2 |
// illegal << 1 | delivered.  The observation is the next state id itself, or — Hansen observations — one more 16-bit
// lookup.  Half the size of the per-thread kernel's 32-bit table.
// TAB = 0: the 32-bit table of the per-thread kernel (observation in the entry); 1: the 16-bit table; 2: the 16-bit
// table + observation lookup (Hansen observations).  The host picks 0 while that leaves five CTAs per SM (the 5x5 map:
// 87.1 us per 10-step launch against 87.3 with the 16-bit table at six CTAs per SM) and the 16-bit table for larger maps
// (8x8 map: 108.1 -> 91.4 us).
template <bool ONE, int TAB, bool DEVCTR = false>
__global__ void __launch_bounds__(kTmaThreads, GPT_TAXI_MINB_TMA) taxi_multi_tma_kernel(const __grid_constant__ TaxiTmaParams TP) {
  constexpr int kShift = TAB == 0 ? kRowShift : kT16RowShift;   // log2 of a table row's bytes
  constexpr bool REPLAY = false;
  constexpr int QPT = 2;
  const TaxiMultiParams& M = TP.m;
  const TaxiParams& P = M.p;
  const int32_t n_steps = M.n_steps;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, abar[kTmaActRows];
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
#pragma unroll
    for (int i = 0; i < kTmaActRows; ++i) mbar_init(&abar[i], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {   // static tables: independent of the previous launch, staged before the dependency wait
    mbar_expect_tx(&bar, TP.tab_bytes);
    tma_bulk_g2s(smem, P.blob, TP.tab_bytes, &bar);
  }
  const uint8_t* trans = smem + P.trans_off;
  const uint16_t* hobs = reinterpret_cast<const uint16_t*>(smem + P.hobs_off);
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);
  uint8_t* stage = smem + TP.stage_off;
  uint8_t* acts = stage + kTmaBufs * kTmaStage;

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t tile = first + (int64_t)blockIdx.x * kTmaEnvs;          // the CTA's envs: [tile, tile + valid)
  const uint32_t valid = (uint32_t)(last - tile < kTmaEnvs ? last - tile : kTmaEnvs);   // a multiple of 512
  const uint32_t loc = warp * (kQuadStride * QPT) + lane * kQuad;       // thread's first quad inside the CTA tile (+ j*128)
  const bool active = loc < valid;                                      // whole warps: a warp owns 256 consecutive envs
  const int64_t base = tile + loc;
  pdl_wait();   // the previous launch's writes are complete and visible from here on
  if (threadIdx.x == 0) {   // the first action rows
    const int rows = n_steps < kTmaActRows ? n_steps : kTmaActRows;
    for (int r = 0; r < rows; ++r) {
      mbar_expect_tx(&abar[r], valid);
      tma_bulk_g2s(acts + r * kTmaEnvs, P.actions + (int64_t)r * M.act_stride + tile, valid, &abar[r]);
    }
  }
  uint64_t ctr_dev = 0;   // graph mode: step counter from device memory, advanced by n_steps for the next launch
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, (uint32_t)n_steps);
  uint32_t off[QPT][4], ndv[QPT][4];   // off = state id << kShift
  int32_t ev[QPT][4];
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { off[j][k] = 0u; ndv[j][k] = 0u; ev[j][k] = 0; }
    if (active) {
      const int64_t q = base + j * kQuadStride;
      const int4 s4 = ld_stream(reinterpret_cast<const int4*>(P.s + q));
      const int4 e4 = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
      const uint32_t nd4 = ONE ? 0u : ld_stream(reinterpret_cast<const uint32_t*>(P.ndrop + q));
      off[j][0] = (uint32_t)s4.x << kShift; off[j][1] = (uint32_t)s4.y << kShift;
      off[j][2] = (uint32_t)s4.z << kShift; off[j][3] = (uint32_t)s4.w << kShift;
      ev[j][0] = e4.x; ev[j][1] = e4.y; ev[j][2] = e4.z; ev[j][3] = e4.w;
#pragma unroll
      for (int k = 0; k < 4; ++k) ndv[j][k] = (nd4 >> (8 * k)) & 0xFFu;
    }
  }
  mbar_wait(&bar, 0);

#pragma unroll 1
  for (int32_t t = 0; t < n_steps; ++t) {
    uint8_t* buf = stage + (t & (kTmaBufs - 1)) * kTmaStage;
    const int slot = t & (kTmaActRows - 1);
    mbar_wait(&abar[slot], (uint32_t)(t / kTmaActRows) & 1u);
    if (active) {
      uint32_t fix_done = 0, fix_goal = 0;   // bit 8k + j: env k of quad j finished its episode / delivered a passenger
#pragma unroll
      for (int j = 0; j < QPT; ++j) {
        const uint32_t lq = loc + j * kQuadStride;
        const uint32_t a4 = *reinterpret_cast<const uint32_t*>(acts + slot * kTmaEnvs + lq);
        float rv[4];
        int32_t ov[4];
        uint32_t tw = 0, trw = 0, gw = 0;   // terminated / truncated / delivered, one byte per env
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t ent;
          if constexpr (TAB == 0) {
            const uint32_t aoff = (k == 0 ? (a4 << 2) : (a4 >> (8 * k - 2))) & 0x1Cu;   // 4 * (action & 7)
            ent = *reinterpret_cast<const uint32_t*>(trans + off[j][k] + aoff);
            off[j][k] = ent & kTransRow;
            ov[k] = (int32_t)(ent >> 16);
          } else {
            const uint32_t aoff = (k == 0 ? (a4 << 1) : (a4 >> (8 * k - 1))) & 0xEu;   // 2 * (action & 7)
            ent = *reinterpret_cast<const uint16_t*>(trans + off[j][k] + aoff);
            const uint32_t nxt = ent >> 2;
            off[j][k] = nxt << kShift;
            ov[k] = TAB == 2 ? (int32_t)hobs[nxt] : (int32_t)nxt;
          }
          const uint32_t goal = ent & kTransGoal;
          if constexpr (!ONE) ndv[j][k] += goal;
          ev[j][k] += 1;
          rv[k] = goal ? P.r_goal : ((ent & kTransBad) ? P.r_bad : P.r_any);
          const uint32_t term = ONE ? goal : (uint32_t)(ndv[j][k] == (uint32_t)P.n_dropoffs);      // (:276-279)
          const uint32_t trunc = ev[j][k] > P.time_limit;
          tw |= term << (8 * k);
          trw |= trunc << (8 * k);
          if constexpr (!ONE) gw |= goal << (8 * k);
        }
        fix_done |= ((tw | trw) & 0x01010101u) << j;
        fix_goal |= (gw & 0x01010101u) << j;
        *reinterpret_cast<int4*>(buf + lq * 4) = make_int4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(buf + 4 * kTmaEnvs + lq * 4) = make_float4(rv[0], rv[1], rv[2], rv[3]);
        *reinterpret_cast<uint32_t*>(buf + 8 * kTmaEnvs + lq) = tw;
        *reinterpret_cast<uint32_t*>(buf + 9 * kTmaEnvs + lq) = trw;
      }
      if (fix_done | fix_goal) {  // rare: full reset of finished envs, passenger respawn after a delivery (:283-286)
#pragma unroll 1
        for (uint32_t m = fix_done | fix_goal; m; m &= m - 1) {
          const int bit = __ffs(m) - 1, k = bit >> 3, j = bit & 7;   // bit 8k + j <-> quad j, env k
          const bool full = (fix_done >> bit) & 1u;
          const int idx = j * 4 + k;
          uint32_t cur = 0;
#pragma unroll
          for (int i = 0; i < 4 * QPT; ++i) cur = i == idx ? off[i >> 2][i & 3] : cur;
          const uint32_t le = loc + j * kQuadStride + k;
          const uint32_t fresh = taxi_fix_inline<REPLAY, DEVCTR>(P, alias, tile + le, (uint32_t)t, cur >> kShift, full, ctr_dev);
          reinterpret_cast<int32_t*>(buf)[le] = (int32_t)hobs[fresh];   // observation of the post-reset state (same thread wrote the quad)
#pragma unroll
          for (int i = 0; i < 4 * QPT; ++i) {
            if (i == idx) {
              off[i >> 2][i & 3] = fresh << kShift;
              ev[i >> 2][i & 3] = full ? 0 : ev[i >> 2][i & 3];
              if constexpr (!ONE) ndv[i >> 2][i & 3] = full ? 0u : ndv[i >> 2][i & 3];
            }
          }
        }
      }
    }
    fence_proxy_async();   // the staged outputs become visible to the TMA engine
    // the buffer step t+1 writes was last read by the bulk stores of step t-1
    if (threadIdx.x == 0) tma_bulk_wait_read<kTmaBufs - 2>();
    __syncthreads();
    if (threadIdx.x == 0) {
      // in-place outputs (out_stride 0: every step overwrites the same rows) are written by the LAST step only — bulk
      // stores of different groups are not ordered against each other
      if (M.out_stride != 0 || t == n_steps - 1) {
        const int64_t o = (int64_t)t * M.out_stride + tile;
        tma_bulk_s2g(P.obs + o, buf, 4 * valid);
        tma_bulk_s2g(P.reward + o, buf + 4 * kTmaEnvs, 4 * valid);
        tma_bulk_s2g(P.terminated + o, buf + 8 * kTmaEnvs, valid);
        tma_bulk_s2g(P.truncated + o, buf + 9 * kTmaEnvs, valid);
        tma_bulk_commit();
      }
      if (t + kTmaActRows < n_steps) {  // every thread has consumed this ring row (barrier above): refill it
        mbar_expect_tx(&abar[slot], valid);
        tma_bulk_g2s(acts + slot * kTmaEnvs, P.actions + (int64_t)(t + kTmaActRows) * M.act_stride + tile, valid, &abar[slot]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t q = base + j * kQuadStride;
      st_stream(reinterpret_cast<int4*>(P.s + q), make_int4((int)(off[j][0] >> kShift), (int)(off[j][1] >> kShift),
                                                            (int)(off[j][2] >> kShift), (int)(off[j][3] >> kShift)));
      st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[j][0], ev[j][1], ev[j][2], ev[j][3]));
      st_stream(reinterpret_cast<uint32_t*>(P.ndrop + q),
                ONE ? 0u : ((ndv[j][0] & 0xFFu) | ((ndv[j][1] & 0xFFu) << 8) | ((ndv[j][2] & 0xFFu) << 16) | ((ndv[j][3] & 0xFFu) << 24)));
    }
  }
  if (threadIdx.x == 0) tma_bulk_wait_read<0>();   // shared memory must outlive the last bulk stores' reads
}


// ---- single-step kernel (gpt_step, replay mode, reset): compact 16-bit table, rare envs patched in memory ----
constexpr uint32_t kT16State = 0x1FFFu, kT16Goal = 1u << 13, kT16Bad = 1u << 14;
constexpr int kT16Cols = 6;  // actions 0..4 + "no-op" column for out-of-range action bytes

// DEVCTR (graph mode): the Philox step counter is read from device memory instead of the launch parameters, so that
// a captured CUDA graph can be replayed (the kernel advances the counter itself: devctr_fetch_and_advance).
template <bool HANSEN, bool REPLAY, bool STATS, int QPT, int THREADS, bool DEVCTR = false>
__global__ void __launch_bounds__(THREADS, (STATS || QPT > 2) ? 1 : 1280 / THREADS) taxi_table_kernel(const __grid_constant__ TaxiParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  pdl_launch_dependents();
  stage_tables_begin(smem, P.blob, P.single_bytes, &bar);   // the 32-bit table of the fused kernel lies behind

  constexpr int kEnvsPerWarp = kWarp * kQuad * QPT;
  const uint32_t lane = threadIdx.x & 31u;
  const int64_t wtile = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const int64_t first = (int64_t)P.first_tile * kTileEnvs, last = first + (int64_t)P.n_tiles * kTileEnvs;
  const int64_t base = first + wtile * kEnvsPerWarp + lane * kQuad;
  if (base >= last) return;
  // reset() is not a separate code path: the host poisons `elapsed` so that every env truncates and
  // launches this same kernel (taxi_launch), which keeps the hot loop free of mode branches.

  pdl_wait();   // the previous step's writes are complete and visible from here on
  uint64_t ctr_dev = 0;
  if constexpr (DEVCTR) ctr_dev = devctr_fetch_and_advance(P.ctr_ptr, 1u);
  int4 s4[QPT], e4[QPT];
  uint32_t nd4[QPT], a4[QPT];
  float4 ret4[QPT];
  constexpr bool stats = STATS;
  EpisodeAcc acc;
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    ret4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    s4[j] = ld_stream(reinterpret_cast<const int4*>(P.s + q));
    e4[j] = ld_stream(reinterpret_cast<const int4*>(P.elapsed + q));
    nd4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.ndrop + q));
    a4[j] = ld_stream(reinterpret_cast<const uint32_t*>(P.actions + q));
    if (stats) ret4[j] = __ldcs(reinterpret_cast<const float4*>(P.ep_return + q));
  }

  stage_tables_wait(&bar);
  const uint16_t* trans = reinterpret_cast<const uint16_t*>(smem + P.trans16_off);
  const uint16_t* hobs = reinterpret_cast<const uint16_t*>(smem + P.hobs_off);
  const uint2* alias = reinterpret_cast<const uint2*>(smem + P.alias_off);

  uint32_t reset_mask = 0u;                                               // bit 4j+k: env needs a full reset
  uint32_t respawn_mask = 0u;                                             // bit 4j+k: new passenger + destination
  int32_t keep_s[4 * QPT];                                                // post-move states (respawn needs the taxi cell)

#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int64_t q = base + j * kQuadStride;
    int32_t sv[4] = {s4[j].x, s4[j].y, s4[j].z, s4[j].w};
    int32_t ev[4] = {e4[j].x, e4[j].y, e4[j].z, e4[j].w};
    int32_t ov[4];
    float rv[4];
    uint32_t ndw = 0, tw = 0, trw = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      {
        const uint32_t a = min((a4[j] >> (8 * k)) & 7u, (uint32_t)(kT16Cols - 1));   // bytes wrap modulo 8; 5..7 are no-ops
        const uint32_t ent = trans[(uint32_t)sv[k] * kT16Cols + a];
        const uint32_t goal = (ent >> 13) & 1u;
        const uint32_t nd = ((nd4[j] >> (8 * k)) & 0xFFu) + goal;
        sv[k] = (int32_t)(ent & kT16State);
        ev[k] += 1;
        rv[k] = goal ? P.r_goal : ((ent & kT16Bad) ? P.r_bad : P.r_any);
        const uint32_t term = nd == (uint32_t)P.n_dropoffs;
        const uint32_t trunc = ev[k] > P.time_limit;
        const uint32_t done = term | trunc;
        reset_mask |= done << (4 * j + k);
        respawn_mask |= (goal & ~done & 1u) << (4 * j + k);
        if (stats) {
          float& ret = k == 0 ? ret4[j].x : (k == 1 ? ret4[j].y : (k == 2 ? ret4[j].z : ret4[j].w));
          ret += rv[k];
          if (q + k < P.num_envs) {
            acc.steps += 1.f;
            if (done) acc.finish(ret, ev[k]);
          }
          ret = done ? 0.f : ret;
        }
        ndw |= (nd & 0xFFu) << (8 * k);
        tw |= term << (8 * k);
        trw |= trunc << (8 * k);
      }
      ov[k] = HANSEN ? (int32_t)hobs[sv[k]] : sv[k];
      keep_s[4 * j + k] = sv[k];
    }
    st_stream(reinterpret_cast<int4*>(P.s + q), make_int4(sv[0], sv[1], sv[2], sv[3]));
    st_stream(reinterpret_cast<int4*>(P.elapsed + q), make_int4(ev[0], ev[1], ev[2], ev[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.ndrop + q), ndw);
    st_stream(reinterpret_cast<int4*>(P.obs + q), make_int4(ov[0], ov[1], ov[2], ov[3]));
    st_stream(reinterpret_cast<float4*>(P.reward + q), make_float4(rv[0], rv[1], rv[2], rv[3]));
    st_stream(reinterpret_cast<uint32_t*>(P.terminated + q), tw);
    st_stream(reinterpret_cast<uint32_t*>(P.truncated + q), trw);
    if (stats) st_stream(reinterpret_cast<float4*>(P.ep_return + q), ret4[j]);
  }
  if (stats) acc.flush(P.stats);

  // ---- rare branches: patch the affected envs (same thread, later stores to the same addresses win)
  uint32_t todo = reset_mask | respawn_mask;
#pragma unroll 1
  while (todo) {
    const int b = __ffs(todo) - 1;
    todo &= todo - 1;
    const int64_t env = base + (b >> 2) * kQuadStride + (b & 3);
    const bool full = (reset_mask >> b) & 1u;
    int32_t cur = 0;
#pragma unroll
    for (int i = 0; i < 4 * QPT; ++i) cur = i == b ? keep_s[i] : cur;
    const uint32_t fresh = taxi_fix_inline<REPLAY, DEVCTR>(P, alias, env, 0u, (uint32_t)cur, full, ctr_dev);   // same sampler as the fused kernel
    if (full) {
      P.elapsed[env] = 0;
      P.ndrop[env] = 0;
    }
    P.s[env] = (int32_t)fresh;
    P.obs[env] = HANSEN ? (int32_t)hobs[fresh] : (int32_t)fresh;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int taxi_create(gpt_env* env, const gpt_config* c) {
  const int64_t cells = (int64_t)c->taxi_rows * c->taxi_cols;
  const int64_t ns = cells * (c->taxi_nlocs + 1) * c->taxi_nlocs;
  if (c->taxi_rows < 1 || c->taxi_cols < 1 || c->taxi_nlocs < 2 || c->taxi_nlocs > 254)
    return fail(GPT_E_ARG, "taxi: need rows, cols >= 1 and 2 <= nlocs <= 254");
  if (ns >= 65536) return fail(GPT_E_ARG, "taxi: rows*cols*(nlocs+1)*nlocs must be < 65536");
  if (c->taxi_n_dropoffs < 0 || c->taxi_n_dropoffs > 255) return fail(GPT_E_ARG, "taxi: num_passengers must be in [0, 255]");
  if (!c->taxi_wall_bits || !c->taxi_loc_cell || !c->taxi_valid_states || c->taxi_n_valid < 1)
    return fail(GPT_E_ARG, "taxi: wall_bits / loc_cell / valid_states missing");
  if (c->rng_mode == GPT_RNG_PHILOX && !c->taxi_reset_cdf) return fail(GPT_E_ARG, "taxi: reset_cdf required in Philox mode");

  if (const char* shape = getenv("GPT_TAXI_SHAPE")) env->taxi_shape = atoi(shape);
  const uint32_t rep = cells <= 256 ? 32u : 1u;
  env->taxi_rep_shift = rep == 32u ? 5u : 0u;
  std::vector<uint16_t> celltab((size_t)cells * rep);
  for (int64_t i = 0; i < cells; ++i) {
    uint16_t ent = (uint16_t)(c->taxi_wall_bits[i] & 15u) | 0xFF00u;
    for (int l = 0; l < c->taxi_nlocs; ++l) {
      if (c->taxi_loc_cell[l] < 0 || c->taxi_loc_cell[l] >= cells) return fail(GPT_E_ARG, "taxi: loc_cell out of range");
      if (c->taxi_loc_cell[l] == i) ent = (uint16_t)((ent & 0x00FFu) | (l << 8));
    }
    for (uint32_t r = 0; r < rep; ++r) celltab[i * rep + r] = ent;
  }
  std::vector<uint32_t> cdf(c->taxi_n_valid, 0xFFFFFFFFu);
  std::vector<uint16_t> valid(c->taxi_n_valid);
  for (int i = 0; i < c->taxi_n_valid; ++i) {
    if (c->taxi_valid_states[i] < 0 || c->taxi_valid_states[i] >= ns) return fail(GPT_E_ARG, "taxi: valid state out of range");
    valid[i] = (uint16_t)c->taxi_valid_states[i];
    if (c->taxi_reset_cdf) cdf[i] = c->taxi_reset_cdf[i];
  }
  cdf.back() = 0xFFFFFFFFu;
  // Walker alias table of the reset law (Philox mode): column j = {threshold, valid[j] | valid[alias_j] << 16}
  std::vector<uint32_t> alias((size_t)c->taxi_n_valid * 2, 0u);
  {
    const int n = c->taxi_n_valid;
    std::vector<double> q(n);
    for (int j = 0; j < n; ++j) {
      const double hi = j == n - 1 ? 4294967296.0 : (double)cdf[j] + 1.0, lo = j == 0 ? 0.0 : (double)cdf[j - 1] + 1.0;
      q[j] = (hi > lo ? hi - lo : 0.0) / 4294967296.0 * n;
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
      alias[(size_t)j * 2] = t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0 ? 0u : (uint32_t)t);
      alias[(size_t)j * 2 + 1] = (uint32_t)valid[j] | ((uint32_t)valid[al[j]] << 16);
    }
  }
  std::vector<uint8_t> blob;
  env->taxi_use_table = ns <= kTableMaxStates;
  if (env->taxi_use_table) {
    // tabulate the whole step relation with the same rule the arithmetic kernel applies per env
    const int nl = c->taxi_nlocs, cols = c->taxi_cols;
    std::vector<uint32_t> trans((size_t)ns * kTransCols);
    std::vector<uint16_t> hobs((size_t)ns), trans16((size_t)ns * kT16Cols), tfused16((size_t)ns * kTransCols);
    for (int64_t st = 0; st < ns; ++st) {   // observation per state: the id itself, or the Hansen re-encoding
      const int d0 = (int)(st % nl), p0 = (int)((st / nl) % (nl + 1)), cell0 = (int)(st / nl / (nl + 1));
      hobs[st] = c->taxi_hansen_obs ? (uint16_t)((((celltab[(size_t)cell0 * rep] & 15) * (nl + 1)) + p0) * nl + d0) : (uint16_t)st;
    }
    for (int64_t st = 0; st < ns; ++st) {
      const int d0 = (int)(st % nl), p0 = (int)((st / nl) % (nl + 1)), cell0 = (int)(st / nl / (nl + 1));
      for (int a = 0; a < kTransCols; ++a) {   // columns 5..7: no-op
        int cell = cell0, p = p0;
        uint16_t ent = celltab[(size_t)cell * rep];
        if (a < 4 && !((ent >> a) & 1)) {
          const int step = (a & 2) ? 1 : cols;
          const int to = cell + ((a & 1) ? step : -step);
          if (to >= 0 && to < cells) cell = to;   // only unreachable wall-cell states can point outside
          ent = celltab[(size_t)cell * rep];
        }
        const int here = ent >> 8;
        const bool act = a == 4;
        const bool goal = act && p == nl && here == d0;
        const bool pickup = act && p < nl && here == p;
        if (pickup) p = nl;
        const bool bad = act && !goal && !pickup;
        const int64_t s2 = ((int64_t)cell * (nl + 1) + p) * nl + d0;
        trans[(size_t)st * kTransCols + a] = ((uint32_t)hobs[s2] << 16) | ((uint32_t)s2 << kRowShift) | (goal ? kTransGoal : 0u) | (bad ? kTransBad : 0u);
        tfused16[(size_t)st * kTransCols + a] = (uint16_t)((s2 << 2) | (goal ? kTransGoal : 0u) | (bad ? kTransBad : 0u));   // ns <= 2048: 11 + 2 bits
        if (a < kT16Cols) trans16[(size_t)st * kT16Cols + a] = (uint16_t)(s2 | (goal ? kT16Goal : 0u) | (bad ? kT16Bad : 0u));
      }
    }
    // blob order: what the single-step kernel stages (compact table, observations, alias) first, the fused
    // kernel's 32-bit table last
    env->taxi_trans16_off = blob_append(blob, trans16);
    env->taxi_hobs_off = blob_append(blob, hobs);
    env->taxi_alias_off = blob_append(blob, alias);
    env->taxi_single_bytes = align16((uint32_t)blob.size());
    env->taxi_trans_off = blob_append(blob, trans);
    env->taxi_t32_end = align16((uint32_t)blob.size());      // fused kernels with the 32-bit table stage hobs .. here
    // 16-bit section of the TMA fused kernel (larger maps): its own copies of hobs and alias, then the 16-bit table
    env->taxi_s16_off = blob_append(blob, hobs);
    env->taxi_s16_alias_off = blob_append(blob, alias);
    env->taxi_tfused16_off = blob_append(blob, tfused16);
  } else {
    blob_append(blob, celltab);
    env->taxi_alias_off = blob_append(blob, alias);
  }
  if (int rc = upload_blob(env, blob)) return rc;

  add_array(env, "s", GPT_ROLE_STATE, GPT_DT_I32, 1);
  add_array(env, "elapsed", GPT_ROLE_STATE, GPT_DT_I32, 1);
  add_array(env, "ndrop", GPT_ROLE_STATE, GPT_DT_U8, 1);
  if (c->track_stats) {
    if (!env->taxi_use_table) return fail(GPT_E_ARG, "taxi: track_stats needs the table kernel (ns <= 2048)");
    add_array(env, "ep_return", GPT_ROLE_STATE, GPT_DT_F32, 1);
  }
  add_array(env, "obs", GPT_ROLE_OUTPUT, GPT_DT_I32, 1);
  add_array(env, "reward", GPT_ROLE_OUTPUT, GPT_DT_F32, 1);
  add_array(env, "terminated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "truncated", GPT_ROLE_OUTPUT, GPT_DT_U8, 1);
  add_array(env, "replay_reset_state", GPT_ROLE_REPLAY, GPT_DT_I32, 1);
  add_array(env, "replay_new_p", GPT_ROLE_REPLAY, GPT_DT_I8, 1);
  add_array(env, "replay_new_d", GPT_ROLE_REPLAY, GPT_DT_I8, 1);
  add_array(env, "actions", GPT_ROLE_ACTION, GPT_DT_I8, 1);
  env->action_dtype = GPT_DT_I8;
  env->action_cols = 1;
  return GPT_OK;
}

bool taxi_can_fuse(const gpt_env* env) {
  return env->taxi_use_table && env->cfg.rng_mode == GPT_RNG_PHILOX;   // replay draws are per step
}

int taxi_launch(gpt_env* env, const LaunchArgs& a) {
  const gpt_config& c = env->cfg;
  TaxiParams P{};
  P.s = (int32_t*)env->ptr("s");
  P.elapsed = (int32_t*)env->ptr("elapsed");
  P.ndrop = (uint8_t*)env->ptr("ndrop");
  P.actions = (const int8_t*)a.actions;
  P.obs = (int32_t*)env->ptr("obs");
  P.reward = (float*)env->ptr("reward");
  P.terminated = (uint8_t*)env->ptr("terminated");
  P.truncated = (uint8_t*)env->ptr("truncated");
  if (!P.s || !P.elapsed || !P.ndrop || !P.obs || !P.reward || !P.terminated || !P.truncated)
    return fail(GPT_E_UNBOUND, "taxi: state/output arrays must be bound before reset/step");
  if (a.mode == kModeStep && !P.actions) return fail(GPT_E_ARG, "taxi: actions is NULL");
  const bool table_reset = a.mode == kModeReset && env->taxi_use_table;
  if (table_reset) {
    // reset() = every env truncates: poison `elapsed` (0x7F7F7F7F > any time limit), step once with any
    // action bytes (the `terminated` array holds 0/1), then clear the outputs the step wrote.
    if (c.time_limit >= 0x7F7F7F7E) return fail(GPT_E_ARG, "taxi: time_limit too large");
    cudaError_t e = cudaMemsetAsync(P.elapsed, 0x7F, (size_t)env->capacity * sizeof(int32_t), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(elapsed)");
    P.actions = (const int8_t*)P.terminated;
  }
  P.obs += a.out_row;
  P.reward += a.out_row;
  P.terminated += a.out_row;
  P.truncated += a.out_row;
  const bool replay = c.rng_mode == GPT_RNG_REPLAY;
  if (replay) {
    P.rp_reset_state = (const int32_t*)env->ptr("replay_reset_state");
    P.rp_new_p = (const int8_t*)env->ptr("replay_new_p");
    P.rp_new_d = (const int8_t*)env->ptr("replay_new_d");
    if (!P.rp_reset_state || !P.rp_new_p || !P.rp_new_d) return fail(GPT_E_UNBOUND, "taxi: replay arrays must be bound in replay mode");
  }
  P.blob = env->d_blob;
  P.blob_bytes = env->taxi_use_table ? env->taxi_t32_end : env->blob_bytes;   // (the 16-bit section behind is the TMA kernel's)
  P.rep_shift = env->taxi_rep_shift;
  P.env_offset = c.env_offset;
  P.first_tile = a.first_tile;
  P.n_tiles = a.n_tiles;
  P.cols = c.taxi_cols;
  P.nlocs = c.taxi_nlocs;
  P.n_dropoffs = c.taxi_n_dropoffs;
  P.time_limit = c.time_limit;
  P.n_valid = c.taxi_n_valid;
  P.mode = a.mode;
  P.div_nlocs = make_fastdiv((uint32_t)c.taxi_nlocs);
  P.div_nlocs1 = make_fastdiv((uint32_t)c.taxi_nlocs + 1);
  P.r_goal = c.taxi_reward_goal;
  P.r_bad = c.taxi_reward_bad;
  P.r_any = c.taxi_reward_any;
  P.rng = make_rng_key(env);

  if (c.track_stats) {
    P.ep_return = (float*)env->ptr("ep_return");
    P.stats = env->d_stats;
    if (!P.ep_return) return fail(GPT_E_UNBOUND, "taxi: ep_return must be bound when track_stats=1");
  }
  P.num_envs = table_reset ? 0 : c.num_envs;   // a reset() does not count as finished episodes
  P.trans_off = env->taxi_trans_off;
  P.hobs_off = env->taxi_hobs_off;
  P.div_pd = make_fastdiv((uint32_t)(c.taxi_nlocs + 1) * (uint32_t)c.taxi_nlocs);
  if (a.n_tiles <= 0) return GPT_OK;
  void* args[] = {(void*)&P};
  const void* kernel;
  int threads, grid;
  size_t smem = P.blob_bytes;
  TaxiMultiParams M;
  TaxiTmaParams TP;
  const bool hansen = c.taxi_hansen_obs != 0;
  P.trans16_off = env->taxi_trans16_off;
  P.alias_off = env->taxi_alias_off;
  P.single_bytes = env->taxi_single_bytes;
  if (env->taxi_use_table && a.n_steps > 1) {  // fused multi-step launch (gpt_step_many)
    if (replay) return fail(GPT_E_ARG, "taxi: fused multi-step launches need Philox mode");
    M.p = P;
    M.n_steps = a.n_steps;
    M.act_stride = env->capacity;
    M.out_stride = a.out_stride_rows;
    using KM = void (*)(const TaxiMultiParams);
    const bool one = c.taxi_n_dropoffs == 1;
    KM km = c.track_stats ? (one ? (KM)taxi_table_multi_kernel<true, 2, 128, true> : (KM)taxi_table_multi_kernel<true, 2, 128, false>)
                          : (one ? (KM)taxi_table_multi_kernel<false, 2, 128, true> : (KM)taxi_table_multi_kernel<false, 2, 128, false>);
    const int qpt = 2;   // launch shapes 1x128, 1x256, 2x256 measured within 2 % of 2x128 (profiles/README.md) and removed
    threads = 128;
    const bool devctr_multi = env->graph_mode && !c.track_stats;   // graph mode: step counter in device memory
    if (devctr_multi) {
      M.p.ctr_ptr = env->d_counter;
      km = one ? (KM)taxi_table_multi_kernel<false, 2, 128, true, true> : (KM)taxi_table_multi_kernel<false, 2, 128, false, true>;
    }
    // default: the TMA-I/O kernel (bulk-copied action rows and outputs); it needs 16-byte aligned rows in every
    // stream, i.e. aligned base pointers and a rollout-slot stride that is a multiple of 16 rows.
    // gpt_set_fused_steps(env, GPT_FUSED_THREADS) or GPT_TAXI_FUSED_LEGACY=1 select the per-thread load/store kernel (A/B runs).
    static const bool legacy_env = getenv("GPT_TAXI_FUSED_LEGACY") != nullptr;
    const bool legacy = env->fused_io == 2 || (legacy_env && env->fused_io == 0);
    const uintptr_t align_or = (uintptr_t)P.actions | (uintptr_t)P.obs | (uintptr_t)P.reward | (uintptr_t)P.terminated | (uintptr_t)P.truncated |
                               (uintptr_t)env->d_blob | (uintptr_t)(a.out_stride_rows & 15) | (uintptr_t)((env->taxi_hobs_off | env->taxi_s16_off) & 15u);
    if (!legacy && !c.track_stats && (align_or & 15u) == 0) {
      // table format: the 32-bit table (hobs | alias | trans) while it leaves five CTAs per SM, else the 16-bit section
      const size_t io_bytes = (size_t)kTmaBufs * kTmaStage + (size_t)kTmaActRows * kTmaEnvs;
      const uint32_t t32_bytes = env->taxi_t32_end - env->taxi_hobs_off;
      const bool t32 = (((size_t)t32_bytes + 127u) & ~(size_t)127u) + io_bytes + 1024 <= 227 * 1024 / 5;
      const uint32_t sec_off = t32 ? env->taxi_hobs_off : env->taxi_s16_off;
      const uint32_t tab_bytes = t32 ? t32_bytes : env->blob_bytes - env->taxi_s16_off;
      const uint32_t stage_off = (tab_bytes + 127u) & ~127u;
      const size_t need = (size_t)stage_off + io_bytes;
      if (need <= 200 * 1024) {
        TP.m = M;
        TP.m.p.blob = env->d_blob + sec_off;
        TP.m.p.hobs_off = 0;
        TP.m.p.alias_off = (t32 ? env->taxi_alias_off : env->taxi_s16_alias_off) - sec_off;
        TP.m.p.trans_off = (t32 ? env->taxi_trans_off : env->taxi_tfused16_off) - sec_off;
        TP.tab_bytes = tab_bytes;
        TP.stage_off = stage_off;
        using KT = void (*)(const TaxiTmaParams);
#define GPT_TAXI_TMA_PICK2(O, D) (t32 ? (KT)taxi_multi_tma_kernel<O, 0, D> : (hansen ? (KT)taxi_multi_tma_kernel<O, 2, D> : (KT)taxi_multi_tma_kernel<O, 1, D>))
#define GPT_TAXI_TMA_PICK(D) (one ? GPT_TAXI_TMA_PICK2(true, D) : GPT_TAXI_TMA_PICK2(false, D))
        KT kt = devctr_multi ? GPT_TAXI_TMA_PICK(true) : GPT_TAXI_TMA_PICK(false);
#undef GPT_TAXI_TMA_PICK
#undef GPT_TAXI_TMA_PICK2
        km = nullptr;
        kernel = (const void*)kt;
        args[0] = (void*)&TP;
        threads = kTmaThreads;
        grid = (int)(((int64_t)a.n_tiles * kTileEnvs + kTmaEnvs - 1) / kTmaEnvs);
        smem = need;
        // five or six CTAs per SM need the largest shared-memory carveout
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
      }
    }
    if (km) {   // per-thread I/O kernel
      const int64_t envs_per_cta = (int64_t)threads * kQuad * qpt;
      grid = (int)(((int64_t)a.n_tiles * kTileEnvs + envs_per_cta - 1) / envs_per_cta);
      kernel = (const void*)km;
      args[0] = (void*)&M;
    }
  } else if (env->taxi_use_table) {
    using K = void (*)(const TaxiParams);
    // launch shape: GPT_TAXI_SHAPE = "<quads per thread>x<threads>" (tuning knob; default 2x128)
#define GPT_TAXI_PICK2(S, Q, T)                                                                                  \
  (hansen ? (replay ? (K)taxi_table_kernel<true, true, S, Q, T> : (K)taxi_table_kernel<true, false, S, Q, T>)      \
          : (replay ? (K)taxi_table_kernel<false, true, S, Q, T> : (K)taxi_table_kernel<false, false, S, Q, T>))
#define GPT_TAXI_PICK(Q, T) (c.track_stats ? GPT_TAXI_PICK2(true, Q, T) : GPT_TAXI_PICK2(false, Q, T))
    K k;
    int qpt;
    P.ctr_ptr = env->d_counter;
    if (env->graph_mode && !replay && !c.track_stats) {  // graph mode: step counter in device memory
      k = hansen ? (K)taxi_table_kernel<true, false, false, 2, 128, true> : (K)taxi_table_kernel<false, false, false, 2, 128, true>;
      qpt = 2; threads = 128;
    } else
    switch (env->taxi_shape) {  // measured on B200 at 2^22 envs: 2x128 18.9 us, 4x128 19.3 us, 1x256 20.3 us per step
      case 4128: k = GPT_TAXI_PICK(4, 128); qpt = 4; threads = 128; break;
      default: k = GPT_TAXI_PICK(2, 128); qpt = 2; threads = 128; break;
    }
#undef GPT_TAXI_PICK
#undef GPT_TAXI_PICK2
    const int64_t envs_per_cta = (int64_t)threads * kQuad * qpt;
    grid = (int)(((int64_t)a.n_tiles * kTileEnvs + envs_per_cta - 1) / envs_per_cta);
    kernel = (const void*)k;
    smem = env->taxi_single_bytes;
  } else {
    using K = void (*)(const TaxiParams);
    threads = 256;
    grid = (a.n_tiles + 7) / 8;
    kernel = (const void*)(hansen ? (replay ? (K)taxi_arith_kernel<true, true> : (K)taxi_arith_kernel<true, false>)
                                  : (replay ? (K)taxi_arith_kernel<false, true> : (K)taxi_arith_kernel<false, false>));
  }
  if (smem > 40 * 1024) {  // static + dynamic shared memory above 48 KB needs the opt-in
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(taxi)");
  }
  cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(threads), smem, a.stream, args);
  env->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "taxi step kernel launch");
  if (table_reset) {
    e = cudaMemsetAsync(P.terminated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.truncated, 0, (size_t)env->capacity, a.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.reward, 0, (size_t)env->capacity * sizeof(float), a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(reset outputs)");
  }
  return GPT_OK;
}

}  // namespace gpt
