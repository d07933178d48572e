/*
 * gpt_b200.h — C ABI of libgpt_b200.so: the B200-native (sm_100a) fused environment step of
 * gym-po-taxi (transition + reward + terminated/truncated + same-step autoreset + observation).
 *
 * This is the drop-in boundary.  The reference has no FFI: its "interface" for
 * this path is the Python vector-env surface
 *     TaxiVecEnv.reset / .step          gym_po/envs/extended_taxi.py:232-242, :244-287
 *     RoomsEnv.reset / .step            gym_po/envs/rooms/rooms.py:177-189, :198-222
 *     CRoomsEnv.reset / .step           gym_po/envs/rooms/crooms.py:251-266, :276-298
 *     AntTagEnv pursuit rules           gym_po/envs/ant_tag.py:105-123, :144-153
 *     CarVecEnv.reset / .step           gym_po/envs/car_flag.py:87-95, :114-141
 *     MultistoryFourRoomsEnv.reset/.step gym_po/envs/rooms/msrooms.py:371-383, :392-428
 * Each entry point below says which of those it replaces.  The Python host classes in
 * gym-po-taxi_b200/gym_po/ bind these symbols with ctypes.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every call returns 0 on success or a negative GPT_E_* code;
 *     gpt_last_error() gives the message of the last failure on the calling thread.
 *   - the CALLER owns every per-env array (state, outputs, replay draws): it allocates them on the
 *     env's device with the dtype/shape gpt_array_info() reports and binds them with gpt_bind().
 *     The library never allocates or frees per-env memory and never allocates inside gpt_step().
 *     The handle owns only the packed static tables (map, wall bits, thresholds; <= 64 KB).
 *   - every per-env array has `capacity` rows, capacity = num_envs rounded up to GPT_ENV_ALIGN;
 *     rows >= num_envs are padding the kernels may read and write.
 *   - STATE arrays index the kernels' shared-memory tables: they must hold in-range values.  gpt_reset() clears
 *     every bound STATE array itself before it runs, so freshly allocated memory is fine as long as gpt_reset() is
 *     the first call; a caller that calls gpt_step() without a reset (legal in the reference) or writes the state
 *     arrays directly (set_state) must zero-initialise them / keep the values in range — the hot kernels do not
 *     range-check.
 *   - calls on one handle are not thread-safe (neither is the reference); handles are independent.
 *   - gpt_reset / gpt_step / gpt_step_many are ASYNCHRONOUS on the given CUDA stream; no hidden
 *     device synchronisation.  `stream` is a cudaStream_t passed as void* (0 = legacy default).
 *   - there is no CPU fallback: without a CUDA device gpt_create() fails with GPT_E_CUDA.
 */
#ifndef GPT_B200_H
#define GPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GPT_API __attribute__((visibility("default")))
#else
#define GPT_API
#endif

#define GPT_ABI_VERSION 2
#define GPT_ENV_ALIGN 512 /* envs per warp tile: 32 lanes x 16 envs */

/* error codes */
#define GPT_OK 0
#define GPT_E_ARG (-1)     /* bad argument / unsupported configuration */
#define GPT_E_CUDA (-2)    /* CUDA runtime error (message has cudaGetErrorString) */
#define GPT_E_UNBOUND (-3) /* a required array has not been bound */
#define GPT_E_DLPACK (-4)  /* DLPack tensor on wrong device / dtype / shape / not contiguous */

/* env families */
#define GPT_FAMILY_TAXI 0   /* TaxiVecEnv            extended_taxi.py:149-372 */
#define GPT_FAMILY_ROOMS 1  /* RoomsEnv              rooms/rooms.py:71-226 */
#define GPT_FAMILY_CROOMS 2 /* CRoomsEnv             rooms/crooms.py:91-338 */
#define GPT_FAMILY_TAG 3    /* point-mass Tag built from AntTagEnv's pursuit rules, ant_tag.py:105-153 */
#define GPT_FAMILY_CAR 4    /* CarVecEnv / DiscreteActionCarVecEnv   car_flag.py:23-144, :286-303 */
#define GPT_FAMILY_MSROOMS 5 /* MultistoryFourRoomsEnv              rooms/msrooms.py:257-428 */

/* random-number modes */
#define GPT_RNG_PHILOX 0 /* Philox4x32-7, key = seed, counter = (global env id | quad id, step, stream) */
#define GPT_RNG_REPLAY 1 /* consume pre-drawn values from the bound REPLAY arrays (parity tests) */

/* ROOMS / CROOMS observation kinds (substring dispatch of rooms/rooms.py:19-67 resolved by the host) */
#define GPT_OBS_ROOM 0          /* grid[a]                                  int32 [B]      */
#define GPT_OBS_ROOM_GOAL 1     /* grid[a] + n_rooms*grid[g]                int32 [B]      */
#define GPT_OBS_MDP 2           /* dense cell id                            int32 [B]      */
#define GPT_OBS_MDP_GOAL 3      /* id(a) + n_cells*id(g)                    int32 [B]      */
#define GPT_OBS_VEC_MDP 4       /* (y,x)          uint8 [B,2]  (CROOMS: real [B,2])      */
#define GPT_OBS_VEC_MDP_GOAL 5  /* (y,x,gy,gx)    uint8 [B,4]  (CROOMS: real [B,4])      */
#define GPT_OBS_HANSEN 6        /* sum(empty_i 2^i) * goal multiplier       int32 [B]      observations.py:44-71 */
#define GPT_OBS_VEC_HANSEN 7    /* 0 wall / 1 empty per neighbour           uint8 [B,n]    observations.py:106-131 */
#define GPT_OBS_VEC_HANSEN_GOAL 8 /* ... and 2 = goal                        uint8 [B,n]    */
#define GPT_OBS_GRID 9          /* n x n egocentric window 0/1/2            uint8 [B,n,n]  observations.py:74-103 */

/* array roles / dtypes reported by gpt_array_info */
#define GPT_ROLE_STATE 0  /* persistent env state, read+written by every step */
#define GPT_ROLE_OUTPUT 1 /* obs / reward / terminated / truncated, written by every step */
#define GPT_ROLE_REPLAY 2 /* pre-drawn random values, read in GPT_RNG_REPLAY mode only */
#define GPT_ROLE_ACTION 3 /* describes the action array gpt_step() expects (never bound) */

#define GPT_DT_U8 0
#define GPT_DT_I8 1
#define GPT_DT_U16 2
#define GPT_DT_I32 3
#define GPT_DT_F32 4
#define GPT_DT_F64 5

typedef struct gpt_env gpt_env; /* opaque */

/* Construction parameters.  Mirrors the reference constructors' kwargs after the host has parsed
 * the map (extended_taxi.py:158-169, rooms/rooms.py:84-100, rooms/crooms.py:104-125).
 * Unused fields must be 0 / NULL.  Pointers are HOST pointers, copied during gpt_create(). */
typedef struct gpt_config {
  int32_t abi_version; /* = GPT_ABI_VERSION */
  int32_t family;      /* GPT_FAMILY_* */
  int32_t rng_mode;    /* GPT_RNG_* */
  int32_t device;      /* CUDA device ordinal */
  int64_t num_envs;    /* B on this device */
  int64_t env_offset;  /* global id of local env 0 (multi-GPU sharding: Philox streams are keyed by
                          global env id, so results do not depend on the GPU count) */
  uint64_t seed;       /* Philox key */
  int32_t time_limit;  /* truncated = elapsed > time_limit (Tag: >=, gymnasium TimeLimit) */
  int32_t track_stats; /* 1: accumulate episode statistics on device (gpt_stats_read) */

  /* ---- TAXI ---- */
  int32_t taxi_rows, taxi_cols;    /* navigable grid (5x5 / 8x8) */
  int32_t taxi_nlocs;              /* named pickup/dropoff locations */
  int32_t taxi_n_dropoffs;         /* num_passengers */
  int32_t taxi_hansen_obs;         /* 0: obs = state id, 1: Hansen-4 wall-bit obs */
  float taxi_reward_goal, taxi_reward_bad, taxi_reward_any;
  const uint8_t* taxi_wall_bits;   /* [rows*cols] bit0 N, bit1 S, bit2 W, bit3 E; 1 = move blocked
                                      (== hansen_encodings, extended_taxi.py:102-114) */
  const int32_t* taxi_loc_cell;    /* [nlocs] cell index r*cols+c of each location */
  int32_t taxi_n_valid;            /* number of valid reset states */
  const int32_t* taxi_valid_states;/* [n_valid] ascending state ids (extended_taxi.py:205-218) */
  const uint32_t* taxi_reset_cdf;  /* [n_valid] Philox mode: inverse-CDF thresholds (x 2^32) of the law
                                      of argmax(multinomial(ns, uniform-over-valid)) over the valid
                                      states in ascending order; may be NULL in replay mode */

  /* ---- ROOMS / CROOMS ---- */
  int32_t rooms_h, rooms_w;        /* grid shape */
  const int8_t* rooms_grid;        /* [h*w] -1 wall, k >= 0 room id (rooms/layouts.py:222-232) */
  int32_t rooms_n_actions;         /* 8 ordinal, 4 cardinal, 0 = continuous yx (CROOMS) */
  const double* rooms_slip_cumsum; /* [n*n] float64 row-wise cumsum of the slip matrix
                                      (action_utils.py:38-48, :85-87), as numpy computes it */
  int32_t rooms_obs_kind;          /* GPT_OBS_* */
  int32_t rooms_obs_n;             /* hansen: 4 or 8; grid: window size n */
  int32_t rooms_goal_y, rooms_goal_x; /* fixed goal cell, or -1,-1 = random goal every episode */
  float rooms_step_reward, rooms_wall_reward, rooms_goal_reward;

  /* ---- CROOMS / TAG (continuous) ---- */
  double c_cell_size, c_action_std, c_action_power, c_goal_threshold;
  int32_t c_use_velocity;
  int32_t c_action_f64; /* continuous (yx) actions are float32 [B,2] (0) or float64 [B,2] (1) */
  int32_t c_state_f32;  /* 0: float64 positions (bit-exact vs the numpy reference); 1: float32 fast mode */
  int32_t car_num_actions;          /* CAR: 0 = continuous force [B,1] (float32, or float64 if c_action_f64); n > 0 =
                                       DiscreteActionCarVecEnv with n evenly spaced forces */
  const double* car_action_table;   /* CAR: [car_num_actions] = np.linspace(-1, 1, n) (car_flag.py:291) */

  /* ---- MSROOMS (multistory FourRooms) ----
   * uses rooms_h / rooms_w / rooms_grid = ONE floor's map [h*w], 0 wall, > 0 walkable (FR_MAP convention,
   * msrooms.py:50-66), rooms_n_actions, rooms_slip_cumsum, rooms_obs_kind (not GPT_OBS_GRID), rooms_obs_n and the
   * three rewards.  Cells are flat ids z*h*w + y*w + x.  Observation dtypes: scalar kinds int32 [B] (ROOM_GOAL may
   * be negative, like the reference), VEC_MDP uint8 [B,3] (z,y,x), VEC_MDP_GOAL uint8 [B,6], VEC_HANSEN* uint8 [B,n]
   * with 0 wall / 2 walkable / 3 goal (msrooms.py:131-160). */
  int32_t ms_floors;                /* grid_z */
  int32_t ms_goal_cell;             /* fixed goal (flat cell id), or -1 = random goal on the top floor every episode */
  int32_t ms_up_y, ms_up_x;         /* stair-up cell of every floor but the top one (msrooms.py:21-24: (1,11)) */
  int32_t ms_down_y, ms_down_x;     /* stair-down cell of every floor but the bottom one ((11,1)); taking the stairs up
                                       lands on the stair-down cell of the floor above and vice versa (:419-428) */
} gpt_config;

typedef struct gpt_array_desc {
  char name[32];     /* e.g. "s", "elapsed", "obs", "reward", "terminated", "replay_u" */
  int32_t role;      /* GPT_ROLE_* */
  int32_t dtype;     /* GPT_DT_* */
  int32_t cols;      /* elements per env (row-major [capacity, cols]) */
  int32_t elem_size; /* bytes per element */
} gpt_array_desc;

/* Host-buffer step (end-to-end path): pinned host arrays with num_envs rows. */
typedef struct gpt_host_io {
  const void* actions; /* in:  [B] int8 (discrete) or [B,2] float32/float64 per the ACTION desc */
  void* obs;           /* out: dtype/cols of the "obs" array */
  float* reward;       /* out: [B] */
  uint8_t* terminated; /* out: [B] */
  uint8_t* truncated;  /* out: [B] */
  void* stream;        /* the caller's cudaStream_t (0 = legacy default): the call's internal copy/compute streams
                          start after all work queued on it (gpt_reset, gpt_step, state uploads) and it is made to
                          wait for them, so step_host may be mixed freely with the stream-asynchronous calls */
} gpt_host_io;

/* --- lifecycle ----------------------------------------------------------------------------
 * replaces the reference constructors (table building on the host side stays in Python). */
GPT_API int gpt_create(const gpt_config* cfg, gpt_env** out);
GPT_API int gpt_destroy(gpt_env* env);
GPT_API int64_t gpt_capacity(const gpt_env* env); /* num_envs rounded up to GPT_ENV_ALIGN */

/* --- array schema + binding -------------------------------------------------------------- */
GPT_API int gpt_array_count(const gpt_env* env);
GPT_API int gpt_array_info(const gpt_env* env, int index, gpt_array_desc* out);
GPT_API int gpt_find_array(const gpt_env* env, const char* name); /* index or GPT_E_ARG */
GPT_API int gpt_bind(gpt_env* env, int index, void* device_ptr, int64_t capacity_rows);
/* DLPack ingestion (zero-copy): `managed` is a `DLManagedTensor*` (capsule "dltensor"); the
 * library validates device, dtype, shape [capacity(,cols)] and contiguity, binds the data pointer
 * and does NOT take ownership (never calls the deleter). */
GPT_API int gpt_bind_dlpack(gpt_env* env, int index, void* managed);

/* --- the hot path -------------------------------------------------------------------------
 * gpt_reset  replaces <Env>.reset(seed=...)   : full reset of every env + first observation
 * gpt_step   replaces <Env>.step(actions)     : one fused kernel launch
 * `actions`: device pointer, int8 [capacity] (discrete) or float32/float64 [capacity,2] (yx). */
GPT_API int gpt_reset(gpt_env* env, int has_seed, uint64_t seed, void* stream);
GPT_API int gpt_step(gpt_env* env, const void* actions, void* stream);
GPT_API int gpt_step_dlpack(gpt_env* env, void* managed_actions, void* stream);
/* T consecutive steps from an action stream [T, capacity]; outputs of step t go to the bound
 * output arrays offset by t*out_stride_rows rows (0 = in place: the arrays hold the LAST step's results afterwards;
 * otherwise >= capacity — rollout slots may be padded). */
GPT_API int gpt_step_many(gpt_env* env, const void* actions, int32_t n_steps, int64_t out_stride_rows, void* stream);
/* Taxi (table kernel), ROOMS and MSRooms in Philox mode run gpt_step_many as ONE fused launch: the state stays in
 * registers for the n_steps steps, only actions are read and outputs written per step; results are bit-identical to
 * n_steps single-step launches.  gpt_set_fused_steps(env, mode): 0 = one launch per step (A/B measurements),
 * 1 = fused, the family's default I/O path, 2 = fused with TMA I/O where the family has it (Taxi, where it is the
 * default: action rows by bulk loads, outputs staged in shared memory and written by bulk stores), 3 = fused with
 * per-thread loads/stores.  ROOMS / MSRooms have the per-thread path only: their fused step is bound by instruction
 * issue, not by its I/O; they fuse with time_limit <= 32766 (16-bit elapsed counters in registers) and
 * step one launch at a time beyond that.  The window observations leave through one TMA bulk store per warp tile. */
#define GPT_FUSED_OFF 0
#define GPT_FUSED_DEFAULT 1
#define GPT_FUSED_TMA 2
#define GPT_FUSED_THREADS 3
GPT_API int gpt_set_fused_steps(gpt_env* env, int mode);
/* Graph mode (every family; Philox mode, no track_stats; Taxi with the table kernel): the Philox step counter moves from the
 * launch parameters into device memory and the step kernels read AND advance it themselves (the last CTA of a grid to
 * fetch it stores counter + steps), so that gpt_step() / gpt_step_many() calls captured into a CUDA graph draw fresh
 * random numbers on every replay — programmatic dependent launch and the fused multi-step launches keep working under
 * capture.  gpt_step_host is unavailable in graph mode.  Call outside of stream capture (it synchronises `stream`). */
GPT_API int gpt_set_graph_mode(gpt_env* env, int enable, void* stream);
/* end-to-end: H2D(actions) -> fused step -> D2H(obs, reward, terminated, truncated), chunked and
 * pipelined on internal streams; returns after the results are in the host buffers. */
GPT_API int gpt_step_host(gpt_env* env, const gpt_host_io* io);

/* --- RNG / bookkeeping -------------------------------------------------------------------- */
GPT_API int gpt_get_counter(const gpt_env* env, uint64_t* counter); /* Philox step counter */
GPT_API int gpt_set_counter(gpt_env* env, uint64_t counter);
GPT_API int gpt_set_env_offset(gpt_env* env, int64_t env_offset);
/* episode statistics (track_stats=1): {n_episodes, sum_return, sum_length, sum_return^2, n_steps, 0,0,0}
 * as float64[8] in DEVICE memory owned by the handle; gpt_stats_ptr exposes it so the host can
 * all-reduce it across ranks (NCCL via torch.distributed) without a copy. */
GPT_API int gpt_stats_ptr(gpt_env* env, void** device_ptr);
GPT_API int gpt_stats_reset(gpt_env* env, void* stream);

/* --- wrapper layer -----------------------------------------------------
 * Device-side equivalents of the two gymnasium wrappers the reference's author stacks on the vector envs
 * (gym_po/tester.py:36-41): RecordEpisodeStatistics and NormalizeReward, fused and family independent.  They
 * read an env's OUTPUT arrays (reward, terminated, truncated: `capacity` rows) right after gpt_step on the same
 * stream.  All per-env arrays are caller-owned, capacity = num_envs rounded up to GPT_ENV_ALIGN, zero-initialised
 * by the caller before the first step. */
#define GPT_WRAP_RECORD 1    /* RecordEpisodeStatistics */
#define GPT_WRAP_NORMALIZE 2 /* NormalizeReward(gamma, epsilon) */
typedef struct gpt_wrap gpt_wrap; /* opaque */
typedef struct gpt_wrap_io {
  const float* reward;       /* in  [cap] */
  const uint8_t* terminated; /* in  [cap] */
  const uint8_t* truncated;  /* in  [cap] */
  float* ep_return;          /* RECORD state [cap]: running episode return */
  int32_t* ep_length;        /* RECORD state [cap]: running episode length */
  float* last_return;        /* RECORD out   [cap]: info["episode"]["r"] — the finished episode's return where
                                terminated|truncated, else 0 */
  int32_t* last_length;      /* RECORD out   [cap]: info["episode"]["l"] */
  float* disc_return;        /* NORMALIZE state [cap]: returns = returns*gamma*(1-terminated) + reward */
  float* norm_reward;        /* NORMALIZE out   [cap]: reward / sqrt(var(returns) + epsilon); may alias reward */
} gpt_wrap_io;
GPT_API int gpt_wrap_create(int device, int64_t num_envs, int flags, double gamma, double epsilon, gpt_wrap** out);
GPT_API int gpt_wrap_destroy(gpt_wrap* w);
/* one accumulate launch (+ one normalise launch with GPT_WRAP_NORMALIZE), asynchronous on `stream` */
GPT_API int gpt_wrap_step(gpt_wrap* w, const gpt_wrap_io* io, void* stream);
/* float64 device vector owned by the handle: [0..4] {episodes, sum_return, sum_length, sum_return^2, env_steps}
 * (same layout as gpt_stats_ptr: all-reduce it across ranks), and at *rms_offset the CURRENT running
 * {count, mean, var} of the discounted returns. */
GPT_API int gpt_wrap_state_ptr(gpt_wrap* w, void** device_ptr, int32_t* rms_offset);
GPT_API int64_t gpt_wrap_launch_count(const gpt_wrap* w);

/* --- diagnostics -------------------------------------------------------------------------- */
/* Debug range check of a discrete action array [capacity] int8 (the reference raises IndexError for an action
 * outside [0, n), extended_taxi.py:248 / rooms.py:210; the hot kernels mask the byte instead): *n_bad = number of
 * the first num_envs bytes outside [0, n).  Synchronises `stream`.  Continuous-action envs report 0. */
GPT_API int gpt_check_actions(gpt_env* env, const void* actions, void* stream, int64_t* n_bad);
/* Copies one of the handle's static device tables to host memory, so that a test can rebuild on the host the
 * random draws a Philox-mode kernel consumes from the very tables the kernel reads.  Names: "reset_alias" (Taxi:
 * uint32 pairs {threshold, state | alias_state << 16} per valid state), "slip_alias" (ROOMS / MSROOMS: per intended
 * action 8 columns of uint32 pairs {threshold, dir | alias_dir << 8}), "spawn_cells" / "goal_cells" (uint16 flat
 * cell ids).  host_out = NULL only queries *n_bytes. */
GPT_API int gpt_table_read(const gpt_env* env, const char* name, void* host_out, int64_t capacity_bytes, int64_t* n_bytes);
GPT_API const char* gpt_last_error(void);
GPT_API int gpt_abi_version(void);
/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
GPT_API int64_t gpt_launch_count(const gpt_env* env);

#ifdef __cplusplus
}
#endif
#endif /* GPT_B200_H */
