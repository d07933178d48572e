// gpt_internal.h — host-side handle shared by the C ABI (gpt_api.cu) and the per-family launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gpt_b200.h"
#include "gpt_common.cuh"

namespace gpt {

struct ArraySlot {
  gpt_array_desc desc;
  void* ptr = nullptr;
  int64_t rows = 0;
};

enum LaunchMode { kModeStep = 0, kModeReset = 1 };

struct HostPath {  // gpt_step_host(): chunked H2D -> step -> D2H pipeline
  static constexpr int kStreams = 4;
  cudaStream_t streams[kStreams] = {};
  cudaEvent_t done[kStreams] = {};
  cudaEvent_t fork = nullptr;   // recorded on the caller's stream: the internal streams wait for it
  void* d_actions = nullptr;  // device staging for the host actions (capacity rows)
  int n_chunks = 2;
  bool ready = false;
};

}  // namespace gpt

struct gpt_env {
  gpt_config cfg{};
  int64_t capacity = 0;
  int32_t n_tiles = 0;
  std::vector<gpt::ArraySlot> arrays;
  // packed static tables (device) — staged into shared memory by every CTA with one TMA bulk copy
  uint8_t* d_blob = nullptr;
  uint32_t blob_bytes = 0;
  double* d_stats = nullptr;
  uint64_t counter = 0;  // Philox step counter
  int64_t launches = 0;
  bool no_fused_steps = false;   // gpt_set_fused_steps(env, 0)
  int fused_io = 0;              // gpt_set_fused_steps(env, 2 / 3): 0 = family default, 1 = TMA I/O, 2 = per-thread loads/stores
  // graph mode (gpt_set_graph_mode): the Philox step counter lives in device memory; the DEVCTR kernels read and advance
  // it themselves (devctr_fetch_and_advance).  d_counter[0] = step counter, d_counter[1] = arrival count of the grid.
  bool graph_mode = false;
  uint64_t* d_counter = nullptr;
  unsigned long long* d_bad = nullptr;   // gpt_check_actions scratch
  gpt::HostPath host;

  // ---- taxi ----
  uint32_t taxi_rep_shift = 0, taxi_trans_off = 0, taxi_hobs_off = 0, taxi_alias_off = 0, taxi_trans16_off = 0, taxi_single_bytes = 0,
           taxi_tfused16_off = 0, taxi_t32_end = 0, taxi_s16_off = 0, taxi_s16_alias_off = 0;
  bool taxi_use_table = false;
  int taxi_shape = 0;  // launch-shape tuning knob (GPT_TAXI_SHAPE), 0 = default
  // ---- rooms / crooms ----
  struct RoomsLayout {
    uint32_t nb8_off = 0, yx_off = 0, room_off = 0, sid_off = 0, valid_off = 0, thr64_off = 0,
             rows_off = 0, grid_off = 0, move_off = 0, obstab_off = 0, alias_off = 0, moveobs_off = 0;
    int32_t n_valid = 0, n_rooms = 0, n_cells = 0;
  } rl;
  // ---- multistory rooms ----
  struct MsLayout {
    uint32_t move_off = 0, moveobs_off = 0, obstab_off = 0, info_off = 0, sid_off = 0, h3_off = 0, vh_off = 0, alias_off = 0,
             thr64_off = 0, avalid_off = 0, gvalid_off = 0;
    int32_t n_agent = 0, n_goal = 0, n_free = 0, room_n = 0, n_cells = 0, obs_bytes = 4;
    bool merged = false;
  } ms;
  int32_t action_dtype = GPT_DT_I8, action_cols = 1;

  int find(const char* name) const;
  void* ptr(const char* name) const;  // nullptr if unbound / missing
};

namespace gpt {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

void add_array(gpt_env* env, const char* name, int role, int dtype, int cols);
int upload_blob(gpt_env* env, const std::vector<uint8_t>& blob);

// per-family: validate the config, build the blob, declare the array schema
int taxi_create(gpt_env* env, const gpt_config* cfg);
int rooms_create(gpt_env* env, const gpt_config* cfg);
int crooms_create(gpt_env* env, const gpt_config* cfg);
int tag_create(gpt_env* env, const gpt_config* cfg);
int car_create(gpt_env* env, const gpt_config* cfg);
int msrooms_create(gpt_env* env, const gpt_config* cfg);

// per-family: launch one fused kernel.  `out_row` offsets the OUTPUT arrays (gpt_step_many);
// `first_tile`/`n_tiles` restrict the launch to a tile range (gpt_step_host chunks); when
// `actions` is non-null it points at row 0 of the action array.
struct LaunchArgs {
  int mode = kModeStep;
  const void* actions = nullptr;
  int64_t out_row = 0;
  int32_t first_tile = 0;
  int32_t n_tiles = 0;
  cudaStream_t stream = nullptr;
  // fused multi-step launch (gpt_step_many on families whose *_can_fuse() says so): n_steps consecutive steps from
  // an action stream [n_steps, capacity], outputs of step t at out_row + t*out_stride_rows
  int32_t n_steps = 1;
  int64_t out_stride_rows = 0;
};
bool taxi_can_fuse(const gpt_env* env);
bool rooms_can_fuse(const gpt_env* env);
bool msrooms_can_fuse(const gpt_env* env);
int taxi_launch(gpt_env* env, const LaunchArgs& a);
int rooms_launch(gpt_env* env, const LaunchArgs& a);
int crooms_launch(gpt_env* env, const LaunchArgs& a);
int tag_launch(gpt_env* env, const LaunchArgs& a);
int car_launch(gpt_env* env, const LaunchArgs& a);
int msrooms_launch(gpt_env* env, const LaunchArgs& a);

// Walker alias tables for the action slip (Philox mode), shared by ROOMS and MSROOMS: per intended action, n
// columns of {threshold (x 2^32), dir | alias_dir << 8} in ordinal-direction units, rows padded to 8 columns.
std::vector<uint32_t> build_slip_alias(int n_actions, const double* cumsum_rows);

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in gpt_common.cuh).
// GPT_NO_PDL=1 falls back to a plain launch (A/B measurements).
inline cudaError_t launch_pdl(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, void** args) {
  static const bool no_pdl = getenv("GPT_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelExC(&cfg, func, args);
}

inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

template <typename T>
inline uint32_t blob_append(std::vector<uint8_t>& blob, const std::vector<T>& v) {
  uint32_t off = align16((uint32_t)blob.size());
  blob.resize(off + align16((uint32_t)(v.size() * sizeof(T))), 0);
  if (!v.empty()) memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
  return off;
}

inline RngKey make_rng_key(const gpt_env* env) {
  RngKey k;
  expand_round_keys(k, env->cfg.seed);
  k.step_lo = (uint32_t)env->counter;
  k.step_hi = (uint32_t)(env->counter >> 32) & 0x00FFFFFFu;
  return k;
}

}  // namespace gpt
