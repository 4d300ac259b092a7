// Context, scratch memory, NCCL plumbing (dlopen), operator dispatch and the micro-benchmarks of libcggp_b200.
#include <dlfcn.h>

#include <cstring>

#include "pipe_common.cuh"

int cggp_matvec_simple(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX, int64_t n,
                       const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, const void* V, int64_t ldv,
                       int B, void* W, int64_t ldw, const int* active);
int cggp_matvec_pipe(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                     const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                     int B, double* W, int64_t ldw, const int* active);
bool cggp_matvec_pipe_supported(cggp_ctx* ctx, int dtype, int64_t m, int D, int B);
int cggp_kuf_times_pipe(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                        const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* Y, int64_t ldy,
                        int P, double* W, int64_t ldw);

static std::string g_err;  // errors raised without a ctx

extern "C" const char* cggp_version(void) { return "cggp_b200 0.1 (sm_100a)"; }

extern "C" int cggp_ctx_create(int device, cggp_ctx** out) {
  if (!out) return CGGP_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
    return CGGP_ERR_CUDA;  // no CPU fallback by design
  }
  if (device < 0 || device >= count) {
    g_err = "device index out of range";
    return CGGP_ERR_INVALID;
  }
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_err = cudaGetErrorString(e);
    return CGGP_ERR_CUDA;
  }
  struct Restore {  // the caller's current device is left as it was
    int d;
    ~Restore() { if (d >= 0) cudaSetDevice(d); }
  } restore{prev_device};
  cggp_ctx* ctx = new cggp_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  if (cudaMalloc(&ctx->cg_state, 16 * sizeof(int)) != cudaSuccess ||
      cudaMallocHost(&ctx->cg_state_host, 16 * sizeof(int)) != cudaSuccess) {
    g_err = "ctx allocation failed";
    delete ctx;
    return CGGP_ERR_CUDA;
  }
  cudaMemset(ctx->cg_state, 0, 16 * sizeof(int));
  *out = ctx;
  return CGGP_OK;
}

struct Ws2 {
  void* p = nullptr;
  size_t bytes = 0;
};
static Ws2& ws2_of(cggp_ctx* ctx) {
  static Ws2 slots[64];
  return slots[ctx->device % 64];
}

extern "C" int cggp_ctx_destroy(cggp_ctx* ctx) {
  if (!ctx) return CGGP_OK;
  cudaSetDevice(ctx->device);
  cggp_ctx_comm_destroy(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  Ws2& w = ws2_of(ctx);
  if (w.p) { cudaFree(w.p); w.p = nullptr; w.bytes = 0; }
  if (ctx->exp_tab) cudaFree(ctx->exp_tab);
  if (ctx->xa2) cudaFree(ctx->xa2);
  if (ctx->cg_state) cudaFree(ctx->cg_state);
  if (ctx->cg_state_host) cudaFreeHost(ctx->cg_state_host);
  for (int s = 0; s < CGGP_PROF_SECTIONS; ++s)
    for (cudaEvent_t e : ctx->prof_ev[s]) cudaEventDestroy(e);
  delete ctx;
  return CGGP_OK;
}

extern "C" int cggp_ctx_set_stream(cggp_ctx* ctx, void* stream) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  ctx->stream = (cudaStream_t)stream;
  return CGGP_OK;
}

extern "C" const char* cggp_last_error(cggp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
extern "C" int64_t cggp_launch_count(cggp_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------------
// Section timing
// ---------------------------------------------------------------------------------------------------------
static const size_t PROF_CAP = 1 << 16;  // event pairs per section; later launches are not recorded

ProfScope::ProfScope(cggp_ctx* c, int s) : ctx(c), sec(s), live(false) {
  if (!c || !c->prof_on || c->prof_used[s] >= PROF_CAP) return;
  std::vector<cudaEvent_t>& ev = c->prof_ev[s];
  const size_t i = c->prof_used[s];
  while (ev.size() < 2 * (i + 1)) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    ev.push_back(e);
  }
  live = cudaEventRecord(ev[2 * i], c->stream) == cudaSuccess;
}
ProfScope::~ProfScope() {
  if (!live) return;
  const size_t i = ctx->prof_used[sec];
  cudaEventRecord(ctx->prof_ev[sec][2 * i + 1], ctx->stream);
  ctx->prof_used[sec] = i + 1;
}

extern "C" int cggp_profile_enable(cggp_ctx* ctx, int on) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  ctx->prof_on = on != 0;
  for (int s = 0; s < CGGP_PROF_SECTIONS; ++s) ctx->prof_used[s] = 0;
  return CGGP_OK;
}

extern "C" int cggp_profile_read(cggp_ctx* ctx, int section, double* host_ms_total, int64_t* host_count) {
  if (!ctx || section < 0 || section >= CGGP_PROF_SECTIONS || !host_ms_total || !host_count) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double total = 0.0;
  for (size_t i = 0; i < ctx->prof_used[section]; ++i) {
    float ms = 0.f;
    CGGP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[section][2 * i], ctx->prof_ev[section][2 * i + 1]));
    total += ms;
  }
  *host_ms_total = total;
  *host_count = (int64_t)ctx->prof_used[section];
  return CGGP_OK;
}

int cggp_ws_reserve(cggp_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return CGGP_OK;
  CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->ws) cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  CGGP_CUDA(ctx, cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return CGGP_OK;
}
int cggp_ws2_reserve(cggp_ctx* ctx, size_t bytes) {
  Ws2& w = ws2_of(ctx);
  if (bytes <= w.bytes) return CGGP_OK;
  CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (w.p) cudaFree(w.p);
  w.p = nullptr;
  w.bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  CGGP_CUDA(ctx, cudaMalloc(&w.p, want));
  w.bytes = want;
  return CGGP_OK;
}
void* cggp_ws2_ptr(cggp_ctx* ctx) { return ws2_of(ctx).p; }

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen: the process already carries the torch-bundled libnccl.so.2; linking a second copy would
// clash, so the five entry points are resolved at run time.
// ---------------------------------------------------------------------------------------------------------
struct Id128 { char b[128]; };  // layout of ncclUniqueId
namespace {
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;

static int nccl_load(cggp_ctx* ctx) {
  if (g_nccl.h) return CGGP_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (!h) h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) CGGP_FAIL(ctx, CGGP_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    CGGP_FAIL(ctx, CGGP_ERR_COMM, "libnccl is missing expected symbols");
  g_nccl.h = h;
  return CGGP_OK;
}

extern "C" int cggp_comm_unique_id(void* host_id128) {
  if (!host_id128) return CGGP_ERR_INVALID;
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  int e = g_nccl.GetUniqueId(host_id128);
  if (e != 0) {
    g_err = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
    return CGGP_ERR_COMM;
  }
  return CGGP_OK;
}

extern "C" int cggp_ctx_comm_init(cggp_ctx* ctx, const void* host_id128, int rank, int world) {
  if (!ctx || !host_id128 || world < 1 || rank < 0 || rank >= world) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  cggp_ctx_comm_destroy(ctx);
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return CGGP_OK;
  int rc = nccl_load(ctx);
  if (rc) return rc;
  CGGP_CUDA(ctx, cudaSetDevice(ctx->device));
  Id128 id;
  memcpy(id.b, host_id128, 128);
  int e = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
  if (e != 0) CGGP_FAIL(ctx, CGGP_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
  return CGGP_OK;
}

extern "C" int cggp_peer_close(cggp_ctx* ctx);
extern "C" int cggp_ctx_comm_destroy(cggp_ctx* ctx) {
  cggp_peer_close(ctx);
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  ctx->comm = nullptr;
  ctx->world = 1;
  ctx->rank = 0;
  return CGGP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory.  The per-iteration collective of the path is tiny (M x 8 bytes = 32 KiB
// at c3): NCCL spends ~60 us on it, most of it launch and protocol latency.  Here every rank copies its vector into a
// slot of its own IPC-shared buffer, publishes a sequence number, waits for the other ranks' numbers and sums all
// slots in RANK ORDER straight out of peer memory (NVLink loads) - one kernel, bit-identical results on all ranks.
// Two slots alternate: a rank reaches call k + 2 only after every peer has published k + 1, i.e. finished reading k.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <typename T>
__global__ void peer_allreduce_kernel(T* __restrict__ buf, int64_t count, char* const* __restrict__ peers, int rank,
                                      int world, unsigned seq, int64_t slot_bytes, int* __restrict__ counter) {
  const int slot = (int)(seq & 1u);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  T* mine = reinterpret_cast<T*>(peers[rank] + slot * slot_bytes);
  if (i < count) mine[i] = buf[i];
  __syncthreads();  // the block's stores are ordered before thread 0's fence (cumulativity through the barrier)
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(counter, 1) == (int)gridDim.x - 1) {  // the last block of this rank: the whole vector is in the slot
      *counter = 0;
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned*>(peers[rank] + 2 * slot_bytes) + slot, seq);
    }
    for (int r = 0; r < world; ++r) {
      const unsigned* flag = reinterpret_cast<const unsigned*>(peers[r] + 2 * slot_bytes) + slot;
      while ((int)(ld_acquire_sys(flag) - seq) < 0) {
      }
    }
  }
  __syncthreads();
  if (i < count) {
    T v = T(0);
    for (int r = 0; r < world; ++r) v += __ldcv(reinterpret_cast<const T*>(peers[r] + slot * slot_bytes) + i);
    buf[i] = v;
  }
}

extern "C" int cggp_peer_close(cggp_ctx* ctx) {
  if (!ctx) return CGGP_OK;
  for (int r = 0; r < ctx->peer_world; ++r)
    if (ctx->peer_ptrs_host[r] && ctx->peer_ptrs_host[r] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_ptrs_host[r]);
  if (ctx->peer_ptrs_dev) cudaFree(ctx->peer_ptrs_dev);
  if (ctx->peer_counter) cudaFree(ctx->peer_counter);
  if (ctx->peer_local) cudaFree(ctx->peer_local);
  for (int r = 0; r < 16; ++r) ctx->peer_ptrs_host[r] = nullptr;
  ctx->peer_ptrs_dev = nullptr;
  ctx->peer_counter = nullptr;
  ctx->peer_local = nullptr;
  ctx->peer_world = 0;
  ctx->peer_seq = 0;
  return CGGP_OK;
}

// step 1 (every rank): allocate this rank's buffer (2 slots of slot_bytes + flags) and export its 64-byte IPC handle
extern "C" int cggp_peer_alloc(cggp_ctx* ctx, int64_t slot_bytes, void* host_handle64) {
  if (!ctx || !host_handle64 || slot_bytes <= 0) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cggp_peer_close(ctx);
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  CGGP_CUDA(ctx, cudaMalloc(&ctx->peer_local, (size_t)(2 * slot_bytes + 256)));
  CGGP_CUDA(ctx, cudaMemset(ctx->peer_local, 0, (size_t)(2 * slot_bytes + 256)));
  CGGP_CUDA(ctx, cudaMalloc(&ctx->peer_counter, sizeof(int)));
  CGGP_CUDA(ctx, cudaMemset(ctx->peer_counter, 0, sizeof(int)));
  cudaIpcMemHandle_t h;
  CGGP_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->peer_local));
  memcpy(host_handle64, &h, 64);
  ctx->peer_slot_bytes = slot_bytes;
  CGGP_CUDA(ctx, cudaDeviceSynchronize());
  return CGGP_OK;
}

// step 2 (every rank, after the handles went round): map all ranks' buffers; `host_handles` = world x 64 bytes in rank
// order.  Call only after cggp_ctx_comm_init (rank / world).
extern "C" int cggp_peer_open(cggp_ctx* ctx, const void* host_handles, int world) {
  if (!ctx || !host_handles) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!ctx->peer_local) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "cggp_peer_alloc first");
  if (world != ctx->world || world < 2 || world > 16) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "peer table: bad world size %d", world);
  for (int r = 0; r < world; ++r) {
    if (r == ctx->rank) {
      ctx->peer_ptrs_host[r] = ctx->peer_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)host_handles + 64 * r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      ctx->peer_world = r;  // close what was opened
      cggp_peer_close(ctx);
      CGGP_FAIL(ctx, CGGP_ERR_COMM, "cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(e));
    }
    ctx->peer_ptrs_host[r] = p;
  }
  ctx->peer_world = world;
  CGGP_CUDA(ctx, cudaMalloc((void**)&ctx->peer_ptrs_dev, sizeof(void*) * 16));
  CGGP_CUDA(ctx, cudaMemcpy(ctx->peer_ptrs_dev, ctx->peer_ptrs_host, sizeof(void*) * 16, cudaMemcpyHostToDevice));
  ctx->peer_seq = 0;
  return CGGP_OK;
}

extern "C" int cggp_peer_enabled(cggp_ctx* ctx) { return ctx && ctx->peer_ptrs_dev && ctx->peer_world == ctx->world; }

extern "C" int cggp_allreduce_sum(cggp_ctx* ctx, int dtype, void* buf, int64_t count) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (ctx->world == 1 || count == 0) return CGGP_OK;
  const int64_t bytes = count * (dtype == CGGP_F64 ? 8 : 4);
  if (cggp_peer_enabled(ctx) && bytes <= ctx->peer_slot_bytes) {
    ProfScope prof(ctx, 3);
    const unsigned seq = ++ctx->peer_seq;
    const unsigned blocks = (unsigned)((count + 1023) / 1024);  // <= 128 blocks of 1024 threads: all co-resident
    if (dtype == CGGP_F64)
      peer_allreduce_kernel<double><<<blocks, 1024, 0, ctx->stream>>>((double*)buf, count, (char* const*)ctx->peer_ptrs_dev,
                                                                     ctx->rank, ctx->world, seq, ctx->peer_slot_bytes,
                                                                     ctx->peer_counter);
    else
      peer_allreduce_kernel<float><<<blocks, 1024, 0, ctx->stream>>>((float*)buf, count, (char* const*)ctx->peer_ptrs_dev,
                                                                    ctx->rank, ctx->world, seq, ctx->peer_slot_bytes,
                                                                    ctx->peer_counter);
    CGGP_LAUNCH_CHECK(ctx);
    return CGGP_OK;
  }
  if (!ctx->comm) CGGP_FAIL(ctx, CGGP_ERR_COMM, "communicator not initialised");
  const int nccl_dtype = dtype == CGGP_F64 ? 8 : 7;  // ncclFloat64 / ncclFloat32
  ProfScope prof(ctx, 3);
  int e = g_nccl.AllReduce(buf, buf, (size_t)count, nccl_dtype, 0 /* ncclSum */, ctx->comm, ctx->stream);
  if (e != 0) CGGP_FAIL(ctx, CGGP_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
  ctx->launches += 1;
  return CGGP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Matrix-free product dispatch
// ---------------------------------------------------------------------------------------------------------
int cggp_matvec_dispatch(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX,
                         int64_t n, const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, const void* V,
                         int64_t ldv, int B, void* W, int64_t ldw, int variant, const int* active) {
  if (kind < CGGP_SE || kind > CGGP_MATERN52) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind);
  const bool can_pipe = cggp_matvec_pipe_supported(ctx, dtype, m, D, B);
  ProfScope prof(ctx, 0);
  if (variant == 3 && !can_pipe)
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec does not support dtype=%d m=%lld D=%d B=%d", dtype,
              (long long)m, D, B);
  if ((variant == 0 && can_pipe) || variant == 3)
    return cggp_matvec_pipe(ctx, kind, variance, (const double*)PX, (const double*)nX, n, (const double*)PZ,
                            (const double*)nZ, m, D, ldp, (const double*)V, ldv, B, (double*)W, ldw, active);
  return cggp_matvec_simple(ctx, dtype, kind, variance, PX, nX, n, PZ, nZ, m, D, ldp, V, ldv, B, W, ldw, active);
}

extern "C" int cggp_kuf_kfu_matvec(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX,
                                   const void* nX, int64_t n, const void* PZ, const void* nZ, int64_t m, int D,
                                   int64_t ldp, const void* V, int64_t ldv, int B, void* W, int64_t ldw, int variant) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (B <= 0 || m <= 0) return CGGP_OK;
  if (variant != 0 && variant != 1 && variant != 3) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "variant must be 0, 1 or 3");
  return cggp_matvec_dispatch(ctx, dtype, kind, variance, PX, nX, n, PZ, nZ, m, D, ldp, V, ldv, B, W, ldw, variant,
                              nullptr);
}

int cggp_matvec_tf32(cggp_ctx* ctx, int kind, double variance, const float* Xb, const float* Xs, const float* xn,
                     int64_t n, const float* Zb, const float* Zs, const float* zn, int64_t m, int D, const float* V,
                     int64_t ldv, int B, float* W, int64_t ldw, int nsplit, const int* active);

extern "C" int cggp_kuf_kfu_matvec_tf32(cggp_ctx* ctx, int kind, double variance, const void* Xb, const void* Xs,
                                        const void* xn, int64_t n, const void* Zb, const void* Zs, const void* zn,
                                        int64_t m, int D, const void* V, int64_t ldv, int B, void* W, int64_t ldw,
                                        int nsplit) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (B <= 0 || m <= 0) return CGGP_OK;
  if (kind < CGGP_SE || kind > CGGP_MATERN52) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind);
  if (nsplit != 1 && nsplit != 3 && nsplit != 16)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "nsplit must be 1, 3 (TF32) or 16 (3xFP16)");
  ProfScope prof(ctx, 0);
  return cggp_matvec_tf32(ctx, kind, variance, (const float*)Xb, (const float*)Xs, (const float*)xn, n,
                          (const float*)Zb, (const float*)Zs, (const float*)zn, m, D, (const float*)V, ldv, B,
                          (float*)W, ldw, nsplit, nullptr);
}

extern "C" int cggp_kuf_times(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX,
                              int64_t n, const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, const void* Y,
                              int64_t ldy, int P, void* W, int64_t ldw) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (P <= 0 || m <= 0) return CGGP_OK;
  if (kind < CGGP_SE || kind > CGGP_MATERN52) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind);
  if (!cggp_matvec_pipe_supported(ctx, dtype, m, D, P))
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "fused Kuf @ Y needs float64 and D <= 31 (dtype=%d, D=%d)", dtype, D);
  ProfScope prof(ctx, 0);
  return cggp_kuf_times_pipe(ctx, kind, variance, (const double*)PX, (const double*)nX, n, (const double*)PZ,
                             (const double*)nZ, m, D, ldp, (const double*)Y, ldy, P, (double*)W, ldw);
}

// ---------------------------------------------------------------------------------------------------------
// Micro-benchmarks (roofline denominators)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mb_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void mb_kernel(int which, double* out, int iters, double s, const int2* etab_g) {
  double acc = 0;
  __shared__ int2 etab_s[1024];
  if (which == 5) {
    for (int j = threadIdx.x; j < 1024; j += blockDim.x) etab_s[j] = etab_g[j];
    __syncthreads();
  }
  if (which == 0) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], s, 1e-9);
    for (int i = 0; i < 8; ++i) acc += a[i];
  } else if (which == 1 || which == 4) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = 0;
    const double a = threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) mb_dmma884(c[2 * i], c[2 * i + 1], a, s);
    for (int i = 0; i < 16; ++i) acc += c[i];
  } else if (which == 2) {
    const FastExpTable tab = fast_exp_table();
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = -(threadIdx.x * 1e-2 + i);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc += fast_exp(a[i], tab); a[i] += s; }
  } else if (which == 3) {
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-2 + i + 1;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc += fast_sqrt_pos(a[i]); a[i] += s; }
  } else if (which == 5) {  // the exp of the pipelined kernels: 1024-entry shared-memory table, degree-3 polynomial
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-2 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc += fast_exp_neg_core_smem<10>(a[i], etab_s); a[i] += s; }
  } else {  // 6: the sqrt of the pipelined kernels: MUFU.RSQ64H seed + one third-order step
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 1e-2 + i + 1;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc += fast_sqrt_pos_cubic(a[i]); a[i] += s; }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

extern "C" int cggp_microbench(cggp_ctx* ctx, int which, int iters, double* host_gops) {
  if (!ctx || !host_gops || which < 0 || which > 6 || iters < 1) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  const int2* etab = nullptr;
  if (which == 5) {
    int rce = kpipe::exp_table_device(ctx, &etab);
    if (rce) return rce;
  }
  const int nb = ctx->sm_count * 8, nt = 256;
  int rc = cggp_ws_reserve(ctx, sizeof(double) * nb * nt);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  CGGP_CUDA(ctx, cudaEventCreate(&e0));
  CGGP_CUDA(ctx, cudaEventCreate(&e1));
  const double s = (which == 0 || which == 1 || which == 4) ? 0.999 : 1e-6;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    mb_kernel<<<nb, nt, 0, ctx->stream>>>(which, (double*)ctx->ws, iters, s, etab);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    ctx->launches += 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CGGP_CUDA(ctx, cudaGetLastError());
  const double threads = (double)nb * nt;
  double ops;
  if (which == 0) ops = threads * iters * 8 * 2.0;
  else if (which == 1 || which == 4) ops = (threads / 32) * iters * 8 * (8 * 8 * 4 * 2.0);
  else ops = threads * iters * 4.0;
  *host_gops = ops / (best * 1e-3) / 1e9;
  return CGGP_OK;
}
