// gpt_api.cu — the extern "C" surface declared in include/gpt_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "gpt_internal.h"

// Minimal DLPack (v0.x ABI) definitions — only what gpt_bind_dlpack / gpt_step_dlpack read.
extern "C" {
typedef struct {
  int32_t device_type;
  int32_t device_id;
} GptDLDevice;
typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} GptDLDataType;
typedef struct {
  void* data;
  GptDLDevice device;
  int32_t ndim;
  GptDLDataType dtype;
  int64_t* shape;
  int64_t* strides;
  uint64_t byte_offset;
} GptDLTensor;
typedef struct GptDLManagedTensor {
  GptDLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct GptDLManagedTensor*);
} GptDLManagedTensor;
}
enum { kDLCUDA = 2, kDLCUDAManaged = 13, kDLInt = 0, kDLUInt = 1, kDLFloat = 2, kDLBool = 6 };

namespace gpt {

// Optional NVTX ranges around the entry points (GPT_NVTX=1): shows the host-side cost of a step next to the kernels
// in an Nsight Systems timeline.  Header-only NVTX3: a no-op unless a profiler is attached.
struct NvtxRange {
  static bool enabled() {
    static const bool on = getenv("GPT_NVTX") != nullptr;
    return on;
  }
  explicit NvtxRange(const char* name) : active(enabled()) {
    if (active) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (active) nvtxRangePop();
  }
  bool active;
};

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
  return GPT_E_CUDA;
}

static int elem_size(int dtype) {
  switch (dtype) {
    case GPT_DT_U8:
    case GPT_DT_I8: return 1;
    case GPT_DT_U16: return 2;
    case GPT_DT_I32:
    case GPT_DT_F32: return 4;
    case GPT_DT_F64: return 8;
  }
  return 0;
}

void add_array(gpt_env* env, const char* name, int role, int dtype, int cols) {
  ArraySlot s;
  memset(&s.desc, 0, sizeof(s.desc));
  snprintf(s.desc.name, sizeof(s.desc.name), "%s", name);
  s.desc.role = role;
  s.desc.dtype = dtype;
  s.desc.cols = cols;
  s.desc.elem_size = elem_size(dtype);
  env->arrays.push_back(s);
}

int upload_blob(gpt_env* env, const std::vector<uint8_t>& blob) {
  env->blob_bytes = align16((uint32_t)blob.size());
  if (env->blob_bytes > 200 * 1024) return fail(GPT_E_ARG, "static tables exceed 200 KB of shared memory");
  cudaError_t e = cudaMalloc((void**)&env->d_blob, env->blob_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(blob)");
  std::vector<uint8_t> padded(env->blob_bytes, 0);
  memcpy(padded.data(), blob.data(), blob.size());
  e = cudaMemcpy(env->d_blob, padded.data(), env->blob_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(blob)");
  return GPT_OK;
}

// debug range check of a discrete action array (gpt_check_actions): counts bytes outside [0, n)
__global__ void count_bad_actions_kernel(const int8_t* a, int64_t n_rows, int n_actions, unsigned long long* out) {
  unsigned long long bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x)
    bad += (unsigned)(int)a[i] >= (unsigned)n_actions ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
  if ((threadIdx.x & 31u) == 0 && bad) atomicAdd(out, bad);
}

static bool graph_mode_family(const gpt_env* env) {
  const gpt_config& c = env->cfg;
  if (c.rng_mode != GPT_RNG_PHILOX || c.track_stats) return false;
  return c.family != GPT_FAMILY_TAXI || env->taxi_use_table;   // every family; Taxi only with the table kernel
}

static int launch(gpt_env* env, const LaunchArgs& a) {
  switch (env->cfg.family) {
    case GPT_FAMILY_TAXI: return taxi_launch(env, a);
    case GPT_FAMILY_ROOMS: return rooms_launch(env, a);
    case GPT_FAMILY_CROOMS: return crooms_launch(env, a);
    case GPT_FAMILY_TAG: return tag_launch(env, a);
    case GPT_FAMILY_CAR: return car_launch(env, a);
    case GPT_FAMILY_MSROOMS: return msrooms_launch(env, a);
  }
  return fail(GPT_E_ARG, "unknown family");
}

static int check_dlpack(const gpt_env* env, const GptDLManagedTensor* m, int dtype, int cols, int64_t min_rows, void** out,
                        int64_t* rows_out) {
  if (!m) return fail(GPT_E_DLPACK, "dlpack: NULL tensor");
  const GptDLTensor& t = m->dl_tensor;
  if (t.device.device_type != kDLCUDA && t.device.device_type != kDLCUDAManaged)
    return fail(GPT_E_DLPACK, "dlpack: tensor is not on a CUDA device");
  if (t.device.device_id != env->cfg.device) return fail(GPT_E_DLPACK, "dlpack: tensor is on a different CUDA device than the env");
  int code = -1, bits = elem_size(dtype) * 8;
  switch (dtype) {
    case GPT_DT_U8: code = kDLUInt; break;
    case GPT_DT_I8: code = kDLInt; break;
    case GPT_DT_U16: code = kDLUInt; break;
    case GPT_DT_I32: code = kDLInt; break;
    case GPT_DT_F32:
    case GPT_DT_F64: code = kDLFloat; break;
  }
  const bool bool_as_u8 = dtype == GPT_DT_U8 && t.dtype.code == kDLBool && t.dtype.bits == 8;
  // int16 storage is accepted for uint16 arrays (torch has no general uint16 support)
  const bool i16_as_u16 = dtype == GPT_DT_U16 && t.dtype.code == kDLInt && t.dtype.bits == 16;
  if (!bool_as_u8 && !i16_as_u16 && (t.dtype.code != code || t.dtype.bits != bits || t.dtype.lanes != 1))
    return fail(GPT_E_DLPACK, "dlpack: dtype mismatch");
  if (t.ndim < 1 || t.ndim > 3) return fail(GPT_E_DLPACK, "dlpack: expected 1 to 3 dimensions");
  int64_t inner = 1;
  for (int i = 1; i < t.ndim; ++i) inner *= t.shape[i];
  if (inner != cols) return fail(GPT_E_DLPACK, "dlpack: trailing dimensions do not match the array's columns");
  if (t.shape[0] < min_rows) return fail(GPT_E_DLPACK, "dlpack: fewer rows than the env capacity");
  if (t.strides) {  // must be row-major contiguous
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; --i) {
      if (t.shape[i] != 1 && t.strides[i] != expect) return fail(GPT_E_DLPACK, "dlpack: tensor is not contiguous");
      expect *= t.shape[i];
    }
  }
  char* p = (char*)t.data + t.byte_offset;
  if (((uintptr_t)p & 15u) != 0) return fail(GPT_E_DLPACK, "dlpack: data pointer is not 16-byte aligned");
  *out = p;
  if (rows_out) *rows_out = t.shape[0];
  return GPT_OK;
}

static int host_path_init(gpt_env* env) {
  HostPath& h = env->host;
  if (h.ready) return GPT_OK;
  for (int i = 0; i < HostPath::kStreams; ++i) {
    cudaError_t e = cudaStreamCreateWithFlags(&h.streams[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate(host path)");
    e = cudaEventCreateWithFlags(&h.done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate(host path)");
  }
  cudaError_t ef = cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming);
  if (ef != cudaSuccess) return cuda_fail(ef, "cudaEventCreate(host path)");
  const size_t abytes = (size_t)env->capacity * env->action_cols * elem_size(env->action_dtype);
  cudaError_t e = cudaMalloc(&h.d_actions, abytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(host path actions)");
  e = cudaMemset(h.d_actions, 0, abytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemset(host path actions)");
  h.n_chunks = 2;  // measured: 1 -> 4.71, 2 -> 4.83, 4 -> 4.54 G env-steps/s (Taxi 2^22, PCIe Gen5)
  if (const char* e2 = getenv("GPT_HOST_CHUNKS")) h.n_chunks = atoi(e2);
  h.ready = true;
  return GPT_OK;
}

}  // namespace gpt

using namespace gpt;

int gpt_env::find(const char* name) const {
  for (size_t i = 0; i < arrays.size(); ++i)
    if (strncmp(arrays[i].desc.name, name, sizeof(arrays[i].desc.name)) == 0) return (int)i;
  return -1;
}
void* gpt_env::ptr(const char* name) const {
  int i = find(name);
  return i < 0 ? nullptr : arrays[i].ptr;
}

extern "C" {

int gpt_abi_version(void) { return GPT_ABI_VERSION; }
const char* gpt_last_error(void) { return g_error.c_str(); }

int gpt_create(const gpt_config* cfg, gpt_env** out) {
  if (!cfg || !out) return fail(GPT_E_ARG, "gpt_create: NULL argument");
  *out = nullptr;
  if (cfg->abi_version != GPT_ABI_VERSION) return fail(GPT_E_ARG, "gpt_create: abi_version mismatch");
  if (cfg->num_envs < 1) return fail(GPT_E_ARG, "gpt_create: num_envs must be >= 1");
  if (cfg->rng_mode != GPT_RNG_PHILOX && cfg->rng_mode != GPT_RNG_REPLAY) return fail(GPT_E_ARG, "gpt_create: bad rng_mode");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GPT_E_CUDA, std::string("gpt_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(GPT_E_ARG, "gpt_create: bad device ordinal");
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");

  gpt_env* env = new (std::nothrow) gpt_env();
  if (!env) return fail(GPT_E_ARG, "gpt_create: out of host memory");
  env->cfg = *cfg;
  env->capacity = (cfg->num_envs + GPT_ENV_ALIGN - 1) / GPT_ENV_ALIGN * GPT_ENV_ALIGN;
  if (env->capacity / GPT_ENV_ALIGN > 0x7FFFFFF0LL) {
    delete env;
    return fail(GPT_E_ARG, "gpt_create: num_envs too large");
  }
  env->n_tiles = (int32_t)(env->capacity / GPT_ENV_ALIGN);
  int rc;
  switch (cfg->family) {
    case GPT_FAMILY_TAXI: rc = taxi_create(env, cfg); break;
    case GPT_FAMILY_ROOMS: rc = rooms_create(env, cfg); break;
    case GPT_FAMILY_CROOMS: rc = crooms_create(env, cfg); break;
    case GPT_FAMILY_TAG: rc = tag_create(env, cfg); break;
    case GPT_FAMILY_CAR: rc = car_create(env, cfg); break;
    case GPT_FAMILY_MSROOMS: rc = msrooms_create(env, cfg); break;
    default: rc = fail(GPT_E_ARG, "gpt_create: unknown family");
  }
  if (rc == GPT_OK) {
    e = cudaMalloc((void**)&env->d_stats, 8 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(env->d_stats, 0, 8 * sizeof(double));
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(stats)");
  }
  if (rc != GPT_OK) {
    std::string keep = g_error;
    gpt_destroy(env);
    g_error = keep;
    return rc;
  }
  // host pointers in the copied config must not be dereferenced after create
  env->cfg.taxi_wall_bits = nullptr;
  env->cfg.taxi_loc_cell = nullptr;
  env->cfg.taxi_valid_states = nullptr;
  env->cfg.taxi_reset_cdf = nullptr;
  env->cfg.rooms_grid = nullptr;
  env->cfg.rooms_slip_cumsum = nullptr;
  env->cfg.car_action_table = nullptr;
  *out = env;
  return GPT_OK;
}

int gpt_destroy(gpt_env* env) {
  if (!env) return GPT_OK;
  cudaSetDevice(env->cfg.device);
  if (env->d_blob) cudaFree(env->d_blob);
  if (env->d_stats) cudaFree(env->d_stats);
  if (env->d_counter) cudaFree(env->d_counter);
  if (env->d_bad) cudaFree(env->d_bad);
  if (env->host.ready) {
    for (int i = 0; i < HostPath::kStreams; ++i) {
      if (env->host.done[i]) cudaEventDestroy(env->host.done[i]);
      if (i == 0 && env->host.fork) cudaEventDestroy(env->host.fork);
      if (env->host.streams[i]) cudaStreamDestroy(env->host.streams[i]);
    }
    if (env->host.d_actions) cudaFree(env->host.d_actions);
  }
  delete env;
  return GPT_OK;
}

int64_t gpt_capacity(const gpt_env* env) { return env ? env->capacity : 0; }
int gpt_array_count(const gpt_env* env) { return env ? (int)env->arrays.size() : 0; }

int gpt_array_info(const gpt_env* env, int index, gpt_array_desc* out) {
  if (!env || !out || index < 0 || index >= (int)env->arrays.size()) return fail(GPT_E_ARG, "gpt_array_info: bad index");
  *out = env->arrays[index].desc;
  return GPT_OK;
}

int gpt_find_array(const gpt_env* env, const char* name) {
  if (!env || !name) return fail(GPT_E_ARG, "gpt_find_array: NULL argument");
  int i = env->find(name);
  return i < 0 ? fail(GPT_E_ARG, std::string("gpt_find_array: no array named ") + name) : i;
}

int gpt_bind(gpt_env* env, int index, void* device_ptr, int64_t capacity_rows) {
  if (!env || index < 0 || index >= (int)env->arrays.size()) return fail(GPT_E_ARG, "gpt_bind: bad index");
  ArraySlot& s = env->arrays[index];
  if (s.desc.role == GPT_ROLE_ACTION) return fail(GPT_E_ARG, "gpt_bind: the action array is passed to gpt_step, not bound");
  if (!device_ptr) return fail(GPT_E_ARG, "gpt_bind: NULL pointer");
  if (capacity_rows < env->capacity) return fail(GPT_E_ARG, "gpt_bind: array has fewer rows than gpt_capacity()");
  if (((uintptr_t)device_ptr & 15u) != 0) return fail(GPT_E_ARG, "gpt_bind: pointer must be 16-byte aligned");
  s.ptr = device_ptr;
  s.rows = capacity_rows;
  return GPT_OK;
}

int gpt_bind_dlpack(gpt_env* env, int index, void* managed) {
  if (!env || index < 0 || index >= (int)env->arrays.size()) return fail(GPT_E_ARG, "gpt_bind_dlpack: bad index");
  const ArraySlot& s = env->arrays[index];
  void* p = nullptr;
  int64_t rows = 0;
  if (int rc = check_dlpack(env, (const GptDLManagedTensor*)managed, s.desc.dtype, s.desc.cols, env->capacity, &p, &rows)) return rc;
  return gpt_bind(env, index, p, rows);
}

int gpt_reset(gpt_env* env, int has_seed, uint64_t seed, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_reset: NULL env");
  NvtxRange range("gpt_reset");
  if (has_seed) {
    env->cfg.seed = seed;
    env->counter = 0;
  }
  LaunchArgs a;
  a.mode = kModeReset;
  a.n_tiles = env->n_tiles;
  a.stream = (cudaStream_t)stream;
  // reset() runs the ordinary step kernel on a poisoned `elapsed` (every env truncates), and that kernel indexes the
  // shared-memory tables with whatever the state arrays hold: clear them first, so that a caller who bound freshly
  // cudaMalloc'ed (uninitialised) memory cannot make the first reset read out of bounds.
  for (const ArraySlot& sl : env->arrays) {
    if (sl.desc.role != GPT_ROLE_STATE || !sl.ptr) continue;
    cudaError_t e = cudaMemsetAsync(sl.ptr, 0, (size_t)env->capacity * sl.desc.cols * sl.desc.elem_size, a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(state arrays)");
  }
  if (env->graph_mode && has_seed) {  // the seed restarts the counter: mirror it to the device (not capturable, like reset itself)
    cudaError_t e = cudaMemcpyAsync(env->d_counter, &env->counter, sizeof(uint64_t), cudaMemcpyHostToDevice, a.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(a.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(step counter)");
  }
  int rc = launch(env, a);
  env->counter += 1;   // graph mode: the kernel advanced the device counter itself (this host mirror is not used then)
  return rc;
}

int gpt_step(gpt_env* env, const void* actions, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_step: NULL env");
  NvtxRange range("gpt_step");
  LaunchArgs a;
  a.mode = kModeStep;
  a.actions = actions;
  a.n_tiles = env->n_tiles;
  a.stream = (cudaStream_t)stream;
  int rc = launch(env, a);
  env->counter += 1;
  return rc;
}

int gpt_step_dlpack(gpt_env* env, void* managed_actions, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_step_dlpack: NULL env");
  void* p = nullptr;
  if (int rc = check_dlpack(env, (const GptDLManagedTensor*)managed_actions, env->action_dtype, env->action_cols, env->capacity,
                            &p, nullptr))
    return rc;
  return gpt_step(env, p, stream);
}

int gpt_step_many(gpt_env* env, const void* actions, int32_t n_steps, int64_t out_stride_rows, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_step_many: NULL env");
  NvtxRange range("gpt_step_many");
  if (n_steps < 0 || out_stride_rows < 0) return fail(GPT_E_ARG, "gpt_step_many: negative argument");
  if (out_stride_rows != 0 && out_stride_rows < env->capacity) return fail(GPT_E_ARG, "gpt_step_many: out_stride_rows < capacity");
  const size_t arow = (size_t)env->action_cols * elem_size(env->action_dtype);
  if (out_stride_rows) {
    for (const ArraySlot& s : env->arrays)
      if (s.desc.role == GPT_ROLE_OUTPUT && s.rows < (int64_t)(n_steps - 1) * out_stride_rows + env->capacity)
        return fail(GPT_E_ARG, "gpt_step_many: bound output arrays are too small for n_steps*out_stride_rows");
  }
  static const bool no_fuse = getenv("GPT_NO_FUSED_STEPS") != nullptr;
  const bool fuse = !no_fuse && !env->no_fused_steps && n_steps > 1 &&
                    ((env->cfg.family == GPT_FAMILY_TAXI && taxi_can_fuse(env)) || (env->cfg.family == GPT_FAMILY_ROOMS && rooms_can_fuse(env)) ||
                     (env->cfg.family == GPT_FAMILY_MSROOMS && msrooms_can_fuse(env)));
  if (fuse) {  // one launch for all n_steps: state stays in registers, only actions are read and outputs written per step
    LaunchArgs a;
    a.mode = kModeStep;
    a.actions = actions;
    a.n_tiles = env->n_tiles;
    a.stream = (cudaStream_t)stream;
    a.n_steps = n_steps;
    a.out_stride_rows = out_stride_rows;
    int rc = launch(env, a);
    env->counter += (uint64_t)n_steps;
    return rc;
  }
  for (int32_t t = 0; t < n_steps; ++t) {
    LaunchArgs a;
    a.mode = kModeStep;
    a.actions = (const char*)actions + (size_t)t * env->capacity * arow;
    a.out_row = t * out_stride_rows;
    a.n_tiles = env->n_tiles;
    a.stream = (cudaStream_t)stream;
    // outputs at row offsets need (t+1)*stride rows in every bound OUTPUT array
    if (out_stride_rows) {
      for (const ArraySlot& s : env->arrays)
        if (s.desc.role == GPT_ROLE_OUTPUT && s.rows < a.out_row + env->capacity)
          return fail(GPT_E_ARG, "gpt_step_many: bound output arrays are too small for n_steps*out_stride_rows");
    }
    int rc = launch(env, a);
    env->counter += 1;
    if (rc) return rc;
  }
  return GPT_OK;
}

int gpt_set_fused_steps(gpt_env* env, int mode) {
  if (!env) return fail(GPT_E_ARG, "gpt_set_fused_steps: NULL env");
  if (mode < GPT_FUSED_OFF || mode > GPT_FUSED_THREADS) return fail(GPT_E_ARG, "gpt_set_fused_steps: mode must be 0..3");
  env->no_fused_steps = mode == GPT_FUSED_OFF;
  env->fused_io = mode == GPT_FUSED_TMA ? 1 : (mode == GPT_FUSED_THREADS ? 2 : 0);
  return GPT_OK;
}

int gpt_set_graph_mode(gpt_env* env, int enable, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_set_graph_mode: NULL env");
  cudaError_t e = cudaSetDevice(env->cfg.device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (enable && env->graph_mode) return GPT_OK;   // already on: the device counter is the truth, leave it alone
  if (enable) {
    if (!graph_mode_family(env))
      return fail(GPT_E_ARG, "gpt_set_graph_mode: needs Philox mode without track_stats (Taxi: a map small enough for the table kernel)");
    if (!env->d_counter) {
      e = cudaMalloc((void**)&env->d_counter, 2 * sizeof(uint64_t));   // {step counter, arrival count}
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(step counter)");
    }
    e = cudaMemsetAsync(env->d_counter, 0, 2 * sizeof(uint64_t), (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(env->d_counter, &env->counter, sizeof(uint64_t), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(step counter)");
    env->graph_mode = true;
  } else if (env->graph_mode) {  // back to launch-parameter counters: fetch the device value (graph replays advanced it)
    e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaMemcpy(&env->counter, env->d_counter, sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(step counter)");
    env->graph_mode = false;
  }
  return GPT_OK;
}

int gpt_step_host(gpt_env* env, const gpt_host_io* io) {
  if (!env || !io || !io->actions || !io->obs || !io->reward || !io->terminated || !io->truncated)
    return fail(GPT_E_ARG, "gpt_step_host: NULL argument");
  if (env->graph_mode) return fail(GPT_E_ARG, "gpt_step_host: not available in graph mode");
  NvtxRange range("gpt_step_host");
  cudaError_t e = cudaSetDevice(env->cfg.device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (int rc = host_path_init(env)) return rc;
  HostPath& h = env->host;
  const int obs_i = env->find("obs");
  char* d_obs = (char*)env->ptr("obs");
  float* d_rew = (float*)env->ptr("reward");
  uint8_t* d_term = (uint8_t*)env->ptr("terminated");
  uint8_t* d_trunc = (uint8_t*)env->ptr("truncated");
  if (obs_i < 0 || !d_obs || !d_rew || !d_term || !d_trunc) return fail(GPT_E_UNBOUND, "gpt_step_host: output arrays must be bound");
  const size_t obs_row = (size_t)env->arrays[obs_i].desc.cols * env->arrays[obs_i].desc.elem_size;
  const size_t act_row = (size_t)env->action_cols * elem_size(env->action_dtype);
  const int64_t B = env->cfg.num_envs;

  // chunk the tile range over the internal streams so H2D, compute and D2H overlap
  // chunks: enough to overlap the H2D of chunk i+1 with the D2H of chunk i, few enough that every copy
  // stays large (PCIe efficiency); GPT_HOST_CHUNKS overrides (tuning knob)
  int n_chunks = h.n_chunks;
  if (n_chunks > env->n_tiles) n_chunks = env->n_tiles;
  if (n_chunks < 1) n_chunks = 1;
  const int32_t per = (env->n_tiles + n_chunks - 1) / n_chunks;
  int rc = GPT_OK;
  // The internal streams are non-blocking: order them behind whatever the caller queued on ITS stream (gpt_reset,
  // gpt_step, set_state / set_replay copies) before the first H2D copy, and the caller's stream behind them at the end.
  cudaStream_t caller = (cudaStream_t)io->stream;
  e = cudaEventRecord(h.fork, caller);
  if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord(host path fork)");
  const int used = n_chunks < HostPath::kStreams ? n_chunks : HostPath::kStreams;
  for (int i = 0; i < used; ++i) {
    e = cudaStreamWaitEvent(h.streams[i], h.fork, 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent(host path fork)");
  }
  for (int c = 0; c < n_chunks && rc == GPT_OK; ++c) {
    const int32_t t0 = c * per;
    const int32_t nt = (t0 + per <= env->n_tiles) ? per : env->n_tiles - t0;
    if (nt <= 0) break;
    cudaStream_t st = h.streams[c % HostPath::kStreams];
    const int64_t r0 = (int64_t)t0 * GPT_ENV_ALIGN;
    int64_t r1 = r0 + (int64_t)nt * GPT_ENV_ALIGN;
    if (r1 > B) r1 = B;
    if (r1 <= r0) break;
    const size_t rows = (size_t)(r1 - r0);
    e = cudaMemcpyAsync((char*)h.d_actions + r0 * act_row, (const char*)io->actions + r0 * act_row, rows * act_row,
                        cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "H2D actions"); break; }
    LaunchArgs a;
    a.mode = kModeStep;
    a.actions = h.d_actions;
    a.first_tile = t0;
    a.n_tiles = nt;
    a.stream = st;
    rc = launch(env, a);
    if (rc) break;
    e = cudaMemcpyAsync((char*)io->obs + r0 * obs_row, d_obs + r0 * obs_row, rows * obs_row, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(io->reward + r0, d_rew + r0, rows * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(io->terminated + r0, d_term + r0, rows, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(io->truncated + r0, d_trunc + r0, rows, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "D2H outputs"); break; }
  }
  env->counter += 1;
  for (int i = 0; i < used; ++i) {  // join: later work on the caller's stream sees this step's state
    e = cudaEventRecord(h.done[i], h.streams[i]);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(caller, h.done[i], 0);
    if (e != cudaSuccess && rc == GPT_OK) rc = cuda_fail(e, "cudaStreamWaitEvent(host path join)");
  }
  for (int i = 0; i < used; ++i) {
    e = cudaStreamSynchronize(h.streams[i]);
    if (e != cudaSuccess && rc == GPT_OK) rc = cuda_fail(e, "cudaStreamSynchronize(host path)");
  }
  return rc;
}

int gpt_get_counter(const gpt_env* env, uint64_t* counter) {
  if (!env || !counter) return fail(GPT_E_ARG, "gpt_get_counter: NULL argument");
  *counter = env->counter;
  if (env->graph_mode) {  // the device value is the truth (graph replays advance it); synchronises the device
    cudaError_t e = cudaSetDevice(env->cfg.device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(counter, env->d_counter, sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(step counter)");
  }
  return GPT_OK;
}
int gpt_set_counter(gpt_env* env, uint64_t counter) {
  if (!env) return fail(GPT_E_ARG, "gpt_set_counter: NULL env");
  env->counter = counter;
  if (env->graph_mode) {
    cudaError_t e = cudaSetDevice(env->cfg.device);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(env->d_counter, &env->counter, sizeof(uint64_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(step counter)");
  }
  return GPT_OK;
}
int gpt_set_env_offset(gpt_env* env, int64_t env_offset) {
  if (!env || env_offset < 0) return fail(GPT_E_ARG, "gpt_set_env_offset: bad argument");
  env->cfg.env_offset = env_offset;
  return GPT_OK;
}
int gpt_stats_ptr(gpt_env* env, void** device_ptr) {
  if (!env || !device_ptr) return fail(GPT_E_ARG, "gpt_stats_ptr: NULL argument");
  *device_ptr = env->d_stats;
  return GPT_OK;
}
int gpt_stats_reset(gpt_env* env, void* stream) {
  if (!env) return fail(GPT_E_ARG, "gpt_stats_reset: NULL env");
  cudaError_t e = cudaMemsetAsync(env->d_stats, 0, 8 * sizeof(double), (cudaStream_t)stream);
  return e == cudaSuccess ? GPT_OK : cuda_fail(e, "cudaMemsetAsync(stats)");
}
int64_t gpt_launch_count(const gpt_env* env) { return env ? env->launches : 0; }

int gpt_table_read(const gpt_env* env, const char* name, void* host_out, int64_t capacity_bytes, int64_t* n_bytes) {
  if (!env || !name || !n_bytes) return fail(GPT_E_ARG, "gpt_table_read: NULL argument");
  const gpt_config& c = env->cfg;
  const std::string nm(name);
  int64_t off = -1, bytes = 0;
  if (c.family == GPT_FAMILY_TAXI) {
    if (nm == "reset_alias") { off = env->taxi_alias_off; bytes = (int64_t)c.taxi_n_valid * 8; }
  } else if (c.family == GPT_FAMILY_ROOMS) {
    if (nm == "slip_alias") { off = env->rl.alias_off; bytes = (int64_t)c.rooms_n_actions * 64; }
    if (nm == "spawn_cells" || nm == "goal_cells") { off = env->rl.valid_off; bytes = (int64_t)env->rl.n_valid * 2; }
  } else if (c.family == GPT_FAMILY_MSROOMS) {
    if (nm == "slip_alias") { off = env->ms.alias_off; bytes = (int64_t)c.rooms_n_actions * 64; }
    if (nm == "spawn_cells") { off = env->ms.avalid_off; bytes = (int64_t)env->ms.n_agent * 2; }
    if (nm == "goal_cells") { off = env->ms.gvalid_off; bytes = c.ms_goal_cell < 0 ? (int64_t)env->ms.n_goal * 2 : 0; }   // fixed goal: no table
  }
  if (off < 0) return fail(GPT_E_ARG, "gpt_table_read: this env has no table named " + nm);
  *n_bytes = bytes;
  if (!host_out || bytes == 0) return GPT_OK;   // size query / empty table (e.g. goal_cells of a fixed-goal env)
  if (capacity_bytes < bytes) return fail(GPT_E_ARG, "gpt_table_read: host buffer too small");
  cudaError_t e = cudaSetDevice(c.device);
  if (e == cudaSuccess) e = cudaMemcpy(host_out, env->d_blob + off, (size_t)bytes, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? GPT_OK : cuda_fail(e, "cudaMemcpy(table)");
}

int gpt_check_actions(gpt_env* env, const void* actions, void* stream, int64_t* n_bad) {
  if (!env || !actions || !n_bad) return fail(GPT_E_ARG, "gpt_check_actions: NULL argument");
  const gpt_config& c = env->cfg;
  int n = 0;
  switch (c.family) {
    case GPT_FAMILY_TAXI: n = 5; break;
    case GPT_FAMILY_ROOMS: case GPT_FAMILY_MSROOMS: n = c.rooms_n_actions; break;
    case GPT_FAMILY_CROOMS: n = env->action_dtype == GPT_DT_I8 ? c.rooms_n_actions : 0; break;
    case GPT_FAMILY_CAR: n = c.car_num_actions; break;
    default: n = 0;
  }
  *n_bad = 0;
  if (n <= 0 || env->action_dtype != GPT_DT_I8) return GPT_OK;   // continuous actions: nothing to range-check
  cudaError_t e = cudaSetDevice(c.device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  if (!env->d_bad) {
    e = cudaMalloc((void**)&env->d_bad, sizeof(unsigned long long));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(action check)");
  }
  cudaStream_t st = (cudaStream_t)stream;
  e = cudaMemsetAsync(env->d_bad, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(action check)");
  const int64_t B = c.num_envs;
  count_bad_actions_kernel<<<(unsigned)((B + 255) / 256 < 1184 ? (B + 255) / 256 : 1184), 256, 0, st>>>((const int8_t*)actions, B, n, env->d_bad);
  env->launches += 1;
  unsigned long long bad = 0;
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, env->d_bad, sizeof(bad), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "gpt_check_actions");
  *n_bad = (int64_t)bad;
  return GPT_OK;
}

}  // extern "C"
