// pcf_api.cu -- the C ABI of include/pcf.h: context set, NCCL plumbing, per-method drivers.
//
// Job shapes:
//   * pcf_init(G): one process, G GPUs, one host thread per GPU for the duration of a call
//     (the front ends' trailing [gpus] argument);
//   * pcf_init_rank(rank, world, device, id): one process per GPU (torchrun), NCCL communicator
//     bootstrapped from a caller-distributed unique id.
// In both, units (paths / antithetic pairs / term pairs) are split into contiguous global index
// ranges and the partial moments meet in one small all-reduce: P2P stores into NVLink-mapped mailboxes issued by the
// reducing kernel itself (csrc/xchg.cuh), or one ncclAllReduce on the compute stream where peer access is unavailable
// (replaces MPI_Reduce, reference src/mc_eur_mpi.cpp:36 etc.).
// NCCL is bound lazily through dlopen("libnccl.so.2") so that single-GPU use has no NCCL
// dependency and a host process that already carries NCCL (PyTorch) shares its copy.
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>
#include <chrono>
#include <cmath>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"
#include "binom_math.cuh"

namespace pcf {

// ---- method drivers implemented in the kernel files --------------------------------------------
int run_mc_eur(Ctx& c, const pcf_params& p, Shard pairs, const double* d_replay, const PeerLink& link);
int run_mc_asia(Ctx& c, const pcf_params& p, Shard paths, const double* d_replay, const PeerLink& link);
int run_mc_basket(Ctx& c, const pcf_params& p, const double* L_host, Shard paths, const double* d_replay,
                  const PeerLink& link, const BasketHost* spec);
int run_mc_amer(Ctx& c, const pcf_params& p, Shard pairs, const double* d_replay, size_t ws_offset);
size_t amer_workspace_bytes(long long local_pairs, int M);
int run_binom_tree(Ctx& c, const pcf_params& p, bool american);
int run_binom(Ctx& c, const pcf_params& p, Shard pairs, bool add_mid, const PeerLink& link);
int run_philox_kat(Ctx& c, const unsigned int ctr[4], const unsigned int key[2], uint32_t* d_out);
int run_normal_stream(Ctx& c, uint64_t seed, uint32_t stream, uint64_t index0, long long count, int T,
                      double scale, double* d_out);
int run_fp64_peak(Ctx& c, double seconds_target, double* dfma_per_sec);
int run_hbm_peak(Ctx& c, long long bytes, double* bytes_per_sec);
void build_math_tables(MathTables& t);
int upload_binom_tables(Ctx& c);

// ---- error text --------------------------------------------------------------------------------
static std::mutex g_err_mu;
static std::string g_last_error;
void set_last_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = s;
}

// ---- NCCL, bound at run time -------------------------------------------------------------------
typedef struct { char internal[128]; } NcclId;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return PCF_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
    if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) {
    set_last_error(std::string("dlopen(libnccl.so.2): ") + dlerror());
    return PCF_ENCCL;
  }
#define PCF_SYM(field, name)                                   \
  *(void**)(&g_nccl.field) = dlsym(h, name);                   \
  if (!g_nccl.field) {                                         \
    set_last_error(std::string("dlsym ") + name + " failed");  \
    return PCF_ENCCL;                                          \
  }
  PCF_SYM(GetUniqueId, "ncclGetUniqueId");
  PCF_SYM(CommInitRank, "ncclCommInitRank");
  PCF_SYM(CommInitAll, "ncclCommInitAll");
  PCF_SYM(CommDestroy, "ncclCommDestroy");
  PCF_SYM(AllReduce, "ncclAllReduce");
  PCF_SYM(GetErrorString, "ncclGetErrorString");
#undef PCF_SYM
  g_nccl.lib = h;
  return PCF_OK;
}

#define PCF_NCCL(expr)                                                                   \
  do {                                                                                   \
    int _r = (expr);                                                                     \
    if (_r != 0) {                                                                       \
      set_last_error(std::string(#expr) + ": " + g_nccl.GetErrorString(_r));             \
      return PCF_ENCCL;                                                                  \
    }                                                                                    \
  } while (0)

int allreduce_sum(Ctx& c, double* d_buf, int count) {
  if (c.world <= 1) return PCF_OK;
  if (!c.comm) {
    set_last_error("multi-GPU job without peer mapping or NCCL communicator");
    return PCF_ENOINIT;
  }
  PCF_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, c.comm, c.stream));
  return PCF_OK;
}

// ---- context set -------------------------------------------------------------------------------
static std::vector<Ctx> g_ctx;

static int ctx_open(Ctx& c, int device, int rank, int world) {
  c.device = device; c.rank = rank; c.world = world;
  PCF_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PCF_CUDA(cudaGetDeviceProperties(&prop, device));
  c.sm_count = prop.multiProcessorCount;
  PCF_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  PCF_CUDA(cudaEventCreate(&c.ev0));
  PCF_CUDA(cudaEventCreate(&c.ev1));
  PCF_CUDA(cudaMalloc(&c.d_partials, sizeof(double) * kMaxBlocks * kMaxMoments * 2));
  PCF_CUDA(cudaMalloc(&c.d_ticket, sizeof(unsigned int)));
  PCF_CUDA(cudaMalloc(&c.d_out, sizeof(double) * 64));
  PCF_CUDA(cudaMemset(c.d_ticket, 0, sizeof(unsigned int)));
  PCF_CUDA(cudaMemset(c.d_out, 0, sizeof(double) * 64));
  {
    void* h = nullptr;
    PCF_CUDA(cudaHostAlloc(&h, sizeof(HostOut), cudaHostAllocMapped | cudaHostAllocPortable));
    std::memset(h, 0, sizeof(HostOut));
    c.h_res = (HostOut*)h;
    void* d = nullptr;
    PCF_CUDA(cudaHostGetDevicePointer(&d, h, 0));
    HostOut* dv = (HostOut*)d;
    c.res_dev = dv->vals;
    c.flag_dev = &dv->flag;
    c.perr_dev = &dv->peer_error;
  }
  PCF_CUDA(cudaMalloc(&c.mailbox, sizeof(Mailbox)));
  PCF_CUDA(cudaMemset(c.mailbox, 0, sizeof(Mailbox)));
  c.link = PeerLink{};
  c.link.rank = rank;
  c.link.world = world;
  c.link.peer[rank] = c.mailbox;
  c.peer_ok = false;
  c.xchg_seq = 0;
  static MathTables host_tables;
  static std::once_flag once;
  std::call_once(once, [] { build_math_tables(host_tables); });
  MathTables* dt = nullptr;
  PCF_CUDA(cudaMalloc(&dt, sizeof(MathTables)));
  PCF_CUDA(cudaMemcpy(dt, &host_tables, sizeof(MathTables), cudaMemcpyHostToDevice));
  c.d_tables = dt;
  PCF_TRY(upload_binom_tables(c));  // Stirling-error table of the binomial kernel: constant, uploaded once per context
  return PCF_OK;
}

static std::vector<void*> g_ipc_opened;

static void ctx_close(Ctx& c) {
  cudaSetDevice(c.device);
  for (void* ptr : g_ipc_opened) cudaIpcCloseMemHandle(ptr);
  g_ipc_opened.clear();
  if (c.mailbox) cudaFree(c.mailbox);
  if (c.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c.comm);
  if (c.workspace) cudaFree(c.workspace);
  if (c.h_res) cudaFreeHost(c.h_res);
  if (c.d_tables) cudaFree((void*)c.d_tables);
  if (c.d_out) cudaFree(c.d_out);
  if (c.d_ticket) cudaFree(c.d_ticket);
  if (c.d_partials) cudaFree(c.d_partials);
  if (c.ev0) cudaEventDestroy(c.ev0);
  if (c.ev1) cudaEventDestroy(c.ev1);
  if (c.stream) cudaStreamDestroy(c.stream);
  c = Ctx();
}

int ctx_reserve(Ctx& c, size_t bytes) {
  if (bytes <= c.workspace_bytes) return PCF_OK;
  if (c.workspace) {
    PCF_CUDA(cudaFree(c.workspace));
    c.workspace = nullptr;
    c.workspace_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&c.workspace, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("cudaMalloc(workspace, " + std::to_string(bytes) + " B): " + cudaGetErrorString(e));
    return PCF_ENOMEM;
  }
  c.workspace_bytes = bytes;
  return PCF_OK;
}

// Runs fn(ctx) for every local context: inline when there is one, one host thread per GPU otherwise.
template <class F>
static int for_each_ctx(F fn) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  if (g_ctx.size() == 1) {
    cudaSetDevice(g_ctx[0].device);
    return fn(g_ctx[0]);
  }
  std::vector<int> st(g_ctx.size(), PCF_OK);
  std::vector<std::thread> th;
  for (size_t i = 0; i < g_ctx.size(); ++i)
    th.emplace_back([&, i] {
      cudaSetDevice(g_ctx[i].device);
      st[i] = fn(g_ctx[i]);
    });
  for (auto& t : th) t.join();
  for (int s : st)
    if (s != PCF_OK) return s;
  return PCF_OK;
}

// Uploads this GPU's slice [off, off+len) of a host replay stream into the context workspace.
static int upload_replay(Ctx& c, const double* host, long long off, long long len, size_t ws_offset,
                         const double** d_ptr) {
  PCF_TRY(ctx_reserve(c, ws_offset + (size_t)(len > 0 ? len : 1) * sizeof(double)));
  double* d = (double*)((char*)c.workspace + ws_offset);
  if (len > 0)
    PCF_CUDA(cudaMemcpyAsync(d, host + off, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  *d_ptr = d;
  return PCF_OK;
}

// Brackets the device work of one call on one context. Every rank consumes the same `n_exch` exchange sequence numbers
// whether or not its own call succeeds, and a rank that fails tells its peers (xchg.cuh poison range), so one rank's
// ENOMEM neither hangs the others nor desynchronises the next call.
template <class F>
static int guarded_call(Ctx& c, int n_exch, F body) {
  const unsigned long long seq0 = c.xchg_seq;
  c.call_first = seq0 + 1;
  c.call_last = seq0 + (unsigned long long)n_exch;
  c.h_res->flag = 0;
  c.h_res->peer_error = 0;
  c.launches = 0;
  int s = body();
  if (use_peer(c)) {
    c.xchg_seq = c.call_last;
    if (s != PCF_OK) {
      cudaGetLastError();
      if (launch_xchg_poison(c) == PCF_OK) cudaStreamSynchronize(c.stream);
    }
  }
  return s;
}

// Common tail of every method: [NCCL all-reduce of k doubles + D2H on the fallback path] -> event -> sync. On the
// default path the kernel's last block has already written the job-wide sums into host-mapped memory (reduce.cuh).
static int finish_call(Ctx& c, int k, double* host_vals, double* seconds_kernel, int* flag) {
  const bool nccl_path = c.world > 1 && !use_peer(c);
  if (nccl_path) PCF_TRY(allreduce_sum(c, c.d_out, k));
  PCF_CUDA(cudaEventRecord(c.ev1, c.stream));
  if (nccl_path) PCF_CUDA(cudaMemcpyAsync(c.h_res->vals, c.d_out, sizeof(double) * k, cudaMemcpyDeviceToHost, c.stream));
  PCF_CUDA(cudaStreamSynchronize(c.stream));
  if (c.h_res->peer_error) {
    set_last_error("peer-memory exchange: a GPU of the job failed or did not answer within the timeout");
    return PCF_ENCCL;
  }
  for (int i = 0; i < k; ++i) host_vals[i] = c.h_res->vals[i];
  float ms = 0.f;
  PCF_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
  *seconds_kernel = ms * 1e-3;
  *flag = c.h_res->flag;
  return PCF_OK;
}

struct CallOut {
  double vals[8] = {0};
  double seconds_kernel = 0;
  int launches = 0;
  int flag = 0;
};

static int check_common(const pcf_params* p, pcf_result* out) {
  if (!p || !out) return PCF_EINVAL;
  std::memset(out, 0, sizeof(*out));
  if (p->cp != 1 && p->cp != -1) return PCF_EINVAL_PAYOFF;
  if (p->N <= 0) return PCF_EINVAL;
  if (g_ctx.empty()) return PCF_ENOINIT;
  return PCF_OK;
}

// The kernels' exponentials scale by 2^k through the exponent field and assume |x| stays inside exp()'s finite range
// (fastmath.cuh). A parameter set whose largest possible argument |drift| + |vol| * zmax leaves it -- where the reference
// would print inf, 0 or NaN -- is refused here instead of being priced wrongly. zmax: the largest |z| the Philox stream
// can produce (8.5), or the largest |w| of the caller's replay stream.
static int check_exp_range(double drift, double vol, double zmax) {
  const double x = std::fabs(drift) + std::fabs(vol) * zmax;
  if (!(x <= 700.0)) {  // also catches NaN / inf parameters
    set_last_error("parameters put exp() arguments up to " + std::to_string(x) + " in play (limit 700)");
    return PCF_EINVAL;
  }
  return PCF_OK;
}
static double replay_absmax(const double* w, long long n) {
  double m = 0.0;
  for (long long i = 0; i < n; ++i) {
    const double a = std::fabs(w[i]);
    if (!(a <= m)) m = a;  // NaN propagates
  }
  return m;
}
constexpr double kHostZMax = 8.5;  // rng.cuh kZMax

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void fill_mc_result(pcf_result* out, const std::vector<CallOut>& co, double disc, long long n,
                           long long units, double t0) {
  out->sum = co[0].vals[0];
  out->sumsq = co[0].vals[1];
  out->n = n;
  out->units = units;
  out->price = (disc * out->sum) / (double)n;
  double mean = out->sum / (double)n, var = out->sumsq / (double)n - mean * mean;
  out->std_error = (n > 1 && var > 0) ? disc * std::sqrt(var / (double)(n - 1)) : 0.0;
  for (auto& c : co) {
    if (c.seconds_kernel > out->seconds_kernel) out->seconds_kernel = c.seconds_kernel;
    out->launches += c.launches;
  }
  out->gpus = g_ctx[0].world;
  out->seconds_total = now_s() - t0;
}

// Builds the NCCL communicator of a single-process job on first need (pcf_init maps the mailboxes instead when it can).
static int ensure_comm_all() {
  if (g_ctx.size() <= 1 || g_ctx[0].comm) return PCF_OK;
  PCF_TRY(nccl_load());
  const int gpus = (int)g_ctx.size();
  std::vector<void*> comms(gpus);
  std::vector<int> devs(gpus);
  for (int i = 0; i < gpus; ++i) devs[i] = g_ctx[i].device;
  PCF_NCCL(g_nccl.CommInitAll(comms.data(), gpus, devs.data()));
  for (int i = 0; i < gpus; ++i) g_ctx[i].comm = comms[i];
  return PCF_OK;
}

}  // namespace pcf

using namespace pcf;

// ================================================================================================
extern "C" {

int pcf_init(int gpus) {
  if (!g_ctx.empty()) return PCF_OK;
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);  // kernel images load with the context (T_overall), not inside the first pricing call
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_last_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    return PCF_ECUDA;
  }
  if (gpus > ndev) {  // reference src/mc_eur_mpi.cpp:58-62: fail loudly at init
    set_last_error("pcf_init: " + std::to_string(gpus) + " GPUs requested, " + std::to_string(ndev) + " visible");
    return PCF_EINVAL;
  }
  if (gpus <= 0) gpus = ndev;
  g_ctx.resize(gpus);
  for (int i = 0; i < gpus; ++i) {
    int s = ctx_open(g_ctx[i], i, i, gpus);
    if (s != PCF_OK) { pcf_shutdown(); return s; }
  }
  if (gpus > 1) {
    // NVLink peer mapping of every GPU's mailbox (xchg.cuh); NCCL only when peer access is unavailable
    bool peer = (gpus <= kMaxWorld) && !getenv("PCF_NO_PEER");
    for (int i = 0; i < gpus && peer; ++i)
      for (int j = 0; j < gpus && peer; ++j) {
        if (i == j) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, i, j) != cudaSuccess || !can) { peer = false; break; }
        cudaSetDevice(i);
        cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
        cudaGetLastError();
      }
    if (peer) {
      for (int i = 0; i < gpus; ++i) {
        for (int r = 0; r < gpus; ++r) g_ctx[i].link.peer[r] = g_ctx[r].mailbox;
        g_ctx[i].peer_ok = true;
      }
    } else {
      int s = ensure_comm_all();
      if (s != PCF_OK) { pcf_shutdown(); return s; }
    }
  }
  return PCF_OK;
}

int pcf_nccl_unique_id(unsigned char id[128]) {
  PCF_TRY(nccl_load());
  NcclId nid;
  PCF_NCCL(g_nccl.GetUniqueId(&nid));
  std::memcpy(id, nid.internal, 128);
  return PCF_OK;
}

int pcf_init_rank(int rank, int world, int device, const unsigned char* nccl_id) {
  if (!g_ctx.empty()) return PCF_OK;
  if (world < 1 || rank < 0 || rank >= world) return PCF_EINVAL;
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_last_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    return PCF_ECUDA;
  }
  g_ctx.resize(1);
  int s = ctx_open(g_ctx[0], device, rank, world);
  if (s != PCF_OK) { pcf_shutdown(); return s; }
  if (world > 1 && nccl_id) {  // NCCL communicator: the fallback when the mailboxes cannot be peer-mapped
    s = nccl_load();
    if (s != PCF_OK) { pcf_shutdown(); return s; }
    NcclId nid;
    std::memcpy(nid.internal, nccl_id, 128);
    int r = g_nccl.CommInitRank(&g_ctx[0].comm, world, nid, rank);
    if (r != 0) {
      set_last_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
      pcf_shutdown();
      return PCF_ENCCL;
    }
  }
  return PCF_OK;
}

int pcf_ipc_export(unsigned char handle[64]) {
  if (g_ctx.size() != 1) return PCF_ENOINIT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  PCF_CUDA(cudaSetDevice(g_ctx[0].device));
  cudaIpcMemHandle_t h;
  PCF_CUDA(cudaIpcGetMemHandle(&h, g_ctx[0].mailbox));
  std::memcpy(handle, &h, 64);
  return PCF_OK;
}

int pcf_ipc_import(const unsigned char* handles, int world) {
  if (g_ctx.size() != 1) return PCF_ENOINIT;
  Ctx& c = g_ctx[0];
  if (!handles || world != c.world || world > kMaxWorld) return PCF_EINVAL;
  PCF_CUDA(cudaSetDevice(c.device));
  for (int r = 0; r < world; ++r) {
    if (r == c.rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * r, 64);
    void* ptr = nullptr;
    PCF_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc_opened.push_back(ptr);
    c.link.peer[r] = (Mailbox*)ptr;
  }
  c.peer_ok = true;
  return PCF_OK;
}

int pcf_peer_enable(int on) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  if (!on && g_ctx[0].world > 1) {
    // switching to ncclAllReduce needs a communicator: built here for a single-process job; a one-process-per-GPU job
    // must have passed an NCCL id to pcf_init_rank
    if (g_ctx.size() > 1) PCF_TRY(ensure_comm_all());
    if (!g_ctx[0].comm) {
      set_last_error("pcf_peer_enable(0): no NCCL communicator (pcf_init_rank was called without an NCCL id)");
      return PCF_ENOINIT;
    }
  }
  for (auto& c : g_ctx) {
    bool mapped = true;
    for (int r = 0; r < c.world; ++r) mapped = mapped && c.link.peer[r] != nullptr;
    c.peer_ok = on && mapped;
  }
  return (on && !g_ctx[0].peer_ok && g_ctx[0].world > 1) ? PCF_EINVAL : PCF_OK;
}

int pcf_peer_active(void) { return (!g_ctx.empty() && use_peer(g_ctx[0])) ? 1 : 0; }

int pcf_shutdown(void) {
  for (auto& c : g_ctx) ctx_close(c);
  g_ctx.clear();
  return PCF_OK;
}

int pcf_world_size(void) { return g_ctx.empty() ? 0 : g_ctx[0].world; }

// ------------------------------------------------------------------------------------------------
int pcf_mc_eur(const pcf_params* p, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && p->replay && p->replay_len < p->N) s = PCF_EINVAL;
  if (s == PCF_OK)
    s = p->replay ? check_exp_range((p->r - p->sigma * p->sigma / 2) * p->T, p->sigma, replay_absmax(p->replay, p->N))
                  : check_exp_range((p->r - p->sigma * p->sigma / 2) * p->T, p->sigma * std::sqrt(p->T), kHostZMax);
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const double t0 = now_s();
  const long long pairs = (p->N + 1) / 2;
  std::vector<CallOut> co(g_ctx.size());
  s = for_each_ctx([&](Ctx& c) -> int {
    CallOut& o = co[&c - &g_ctx[0]];
    return guarded_call(c, 1, [&]() -> int {
      Shard sh = shard_of(pairs, c.rank, c.world);
      const double* d_rep = nullptr;
      if (p->replay) {
        long long off = 2 * sh.begin, end = std::min(p->N, 2 * sh.end);
        PCF_TRY(upload_replay(c, p->replay, off, end - off, 0, &d_rep));
      }
      PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
      const PeerLink l = next_link(c);
      PCF_TRY(run_mc_eur(c, *p, sh, d_rep, l));
      PCF_TRY(finish_call(c, 2, o.vals, &o.seconds_kernel, &o.flag));
      o.launches = c.launches;
      return PCF_OK;
    });
  });
  if (s == PCF_OK) fill_mc_result(out, co, std::exp(-p->r * p->T), p->N, p->N, t0);
  out->status = s;
  return s;
}

int pcf_mc_asia(const pcf_params* p, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && p->M <= 0) s = PCF_EINVAL;
  if (s == PCF_OK && p->replay && p->replay_len < p->N * (long long)p->M) s = PCF_EINVAL;
  if (s == PCF_OK) {
    const double dt = p->T / p->M, adt = (p->r - p->sigma * p->sigma / 2) * dt;
    s = p->replay ? check_exp_range(adt, p->sigma, replay_absmax(p->replay, p->N * (long long)p->M))
                  : check_exp_range(adt, p->sigma * std::sqrt(dt), kHostZMax);
  }
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const double t0 = now_s();
  std::vector<CallOut> co(g_ctx.size());
  s = for_each_ctx([&](Ctx& c) -> int {
    CallOut& o = co[&c - &g_ctx[0]];
    return guarded_call(c, 1, [&]() -> int {
      Shard sh = shard_of(p->N, c.rank, c.world);
      const double* d_rep = nullptr;
      if (p->replay) PCF_TRY(upload_replay(c, p->replay, sh.begin * p->M, sh.size() * p->M, 0, &d_rep));
      PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
      const PeerLink l = next_link(c);
      PCF_TRY(run_mc_asia(c, *p, sh, d_rep, l));
      PCF_TRY(finish_call(c, 2, o.vals, &o.seconds_kernel, &o.flag));
      o.launches = c.launches;
      return PCF_OK;
    });
  });
  if (s == PCF_OK) fill_mc_result(out, co, std::exp(-p->r * p->T), p->N, p->N * (long long)p->M, t0);
  out->status = s;
  return s;
}

int pcf_chol_equicorr(int d, double rho, double* L) {
  if (d < 1 || d > PCF_MAX_ASSETS || !L) return PCF_EINVAL;
  // covar(i,j) = i==j ? 1 : rho  (reference include/mvn.h:55-60), row-by-row Cholesky (mvn.h:63-70)
  std::memset(L, 0, sizeof(double) * d * d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = (i == j) ? 1.0 : rho;
      for (int k = 0; k < j; ++k) s -= L[i * d + k] * L[j * d + k];
      if (i == j) {
        if (!(s > 0.0)) return PCF_ENOTPD;
        L[i * d + i] = std::sqrt(s);
      } else {
        L[i * d + j] = s / L[j * d + j];
      }
    }
  return PCF_OK;
}

// Shared driver of the two basket entry points. `spec` == nullptr: the reference's basket.
static int basket_call(const pcf_params* p, const double* L, const BasketHost* spec, pcf_result* out) {
  {
    // largest exponent in play: |drift_a| + sigma_a * sum_k |A[a][k]| * zmax
    const int d = p->assets;
    const double zmax = p->replay ? replay_absmax(p->replay, p->N * (long long)d) : kHostZMax;
    for (int a = 0; a < d; ++a) {
      double rowsum = 0;
      for (int k = 0; k < d; ++k) rowsum += std::fabs(L[a * d + k]);
      const double sg = spec ? spec->sigma[a] : p->sigma;
      int s = check_exp_range((p->r - sg * sg / 2) * p->T, sg * rowsum, zmax);
      if (s != PCF_OK) { out->status = s; return s; }
    }
  }
  const double t0 = now_s();
  std::vector<CallOut> co(g_ctx.size());
  int s = for_each_ctx([&](Ctx& c) -> int {
    CallOut& o = co[&c - &g_ctx[0]];
    return guarded_call(c, 1, [&]() -> int {
      Shard sh = shard_of(p->N, c.rank, c.world);
      const double* d_rep = nullptr;
      if (p->replay)
        PCF_TRY(upload_replay(c, p->replay, sh.begin * p->assets, sh.size() * p->assets, 0, &d_rep));
      PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
      const PeerLink l = next_link(c);
      PCF_TRY(run_mc_basket(c, *p, L, sh, d_rep, l, spec));
      PCF_TRY(finish_call(c, 2, o.vals, &o.seconds_kernel, &o.flag));
      o.launches = c.launches;
      return PCF_OK;
    });
  });
  if (s == PCF_OK) fill_mc_result(out, co, std::exp(-p->r * p->T), p->N, p->N, t0);
  out->status = s;
  return s;
}

int pcf_mc_eur_multi(const pcf_params* p, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && (p->assets < 1 || p->assets > PCF_MAX_ASSETS)) s = PCF_EINVAL;
  if (s == PCF_OK && p->replay && p->replay_len < p->N * (long long)p->assets) s = PCF_EINVAL;
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const int d = p->assets;
  double L[PCF_MAX_ASSETS * PCF_MAX_ASSETS];
  s = pcf_chol_equicorr(d, p->rho, L);
  if (s == PCF_OK) return basket_call(p, L, nullptr, out);
  // mvn.h:68-76: LLT reported a non-positive pivot -> eigenvectors * sqrt(eigenvalues) of the same matrix. A positive
  // SEMI-definite matrix (rho = 1, rho = -1/(d-1)) prices through this branch as it does in the reference; a matrix with
  // a genuinely negative eigenvalue (rho < -1/(d-1), where the reference's cwiseSqrt yields NaN samples) is PCF_ENOTPD.
  double cov[PCF_MAX_ASSETS * PCF_MAX_ASSETS], S0[PCF_MAX_ASSETS], sg[PCF_MAX_ASSETS], w[PCF_MAX_ASSETS];
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) cov[i * d + j] = (i != j) ? p->rho : 1.0;  // mvn.h:55-60
  s = pcf_normal_transform(d, cov, L, nullptr);
  if (s != PCF_OK) { out->status = s; return s; }
  for (int i = 0; i < d; ++i) {
    S0[i] = p->S0;
    sg[i] = p->sigma;
    w[i] = 1.0 / (double)d;  // mc_eur_multi.cpp:23
  }
  BasketHost spec{S0, sg, w, true};
  return basket_call(p, L, &spec, out);
}

// Cyclic Jacobi eigen-decomposition of a symmetric d x d matrix (row-major): V's columns are the eigenvectors,
// lam the eigenvalues. d <= 32, so a few sweeps of d(d-1)/2 rotations on the host.
static void jacobi_eigen(int d, const double* Sin, double* V, double* lam) {
  std::vector<double> A(Sin, Sin + (size_t)d * d);
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) V[i * d + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) (i == j ? diag : off) += A[i * d + j] * A[i * d + j];
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int pi = 0; pi < d - 1; ++pi)
      for (int q = pi + 1; q < d; ++q) {
        const double apq = A[pi * d + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * d + q] - A[pi * d + pi]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < d; ++k) {  // columns pi, q of A
          const double akp = A[k * d + pi], akq = A[k * d + q];
          A[k * d + pi] = cs * akp - sn * akq;
          A[k * d + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < d; ++k) {  // rows pi, q of A
          const double apk = A[pi * d + k], aqk = A[q * d + k];
          A[pi * d + k] = cs * apk - sn * aqk;
          A[q * d + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < d; ++k) {
          const double vkp = V[k * d + pi], vkq = V[k * d + q];
          V[k * d + pi] = cs * vkp - sn * vkq;
          V[k * d + q] = sn * vkp + cs * vkq;
        }
      }
  }
  // Eigen's documented convention (SelfAdjointEigenSolver): eigenvalues in increasing order, eigenvectors permuted with
  // them. The order matters: column k of the transform multiplies the k-th normal of a path (mvn.h:74-79).
  std::vector<int> order(d);
  for (int i = 0; i < d; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return A[x * d + x] < A[y * d + y]; });
  std::vector<double> Vs((size_t)d * d);
  for (int j = 0; j < d; ++j) {
    lam[j] = A[order[j] * d + order[j]];
    for (int i = 0; i < d; ++i) Vs[i * d + j] = V[i * d + order[j]];
  }
  std::copy(Vs.begin(), Vs.end(), V);
}

int pcf_normal_transform(int d, const double* cov, double* A, int* used_eigen) {
  if (d < 1 || d > PCF_MAX_ASSETS || !cov || !A) return PCF_EINVAL;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < i; ++j) {
      const double a = cov[i * d + j], b = cov[j * d + i];
      if (!(std::fabs(a - b) <= 1e-12 * (std::fabs(a) + std::fabs(b)) || a == b)) return PCF_EINVAL;  // also rejects NaN
    }
  if (used_eigen) *used_eigen = 0;
  // row-by-row Cholesky of the lower triangle (mvn.h:63-70)
  std::memset(A, 0, sizeof(double) * d * d);
  bool pd = true;
  for (int i = 0; i < d && pd; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = cov[i * d + j];
      for (int k = 0; k < j; ++k) s -= A[i * d + k] * A[j * d + k];
      if (i == j) {
        if (!(s > 0.0)) { pd = false; break; }
        A[i * d + i] = std::sqrt(s);
      } else {
        A[i * d + j] = s / A[j * d + j];
      }
    }
  if (pd) return PCF_OK;
  // mvn.h:72-76: eigenvectors * sqrt(eigenvalues). Eigenvalues in [-tol, 0) are rounding noise of a semi-definite matrix
  // and count as 0 (Eigen's cwiseSqrt would give NaN there); anything more negative is not a covariance matrix.
  std::vector<double> V((size_t)d * d), lam(d);
  jacobi_eigen(d, cov, V.data(), lam.data());
  double lmax = 0;
  for (int i = 0; i < d; ++i) lmax = std::max(lmax, std::fabs(lam[i]));
  for (int i = 0; i < d; ++i)
    if (!(lam[i] >= -1e-10 * lmax)) return PCF_ENOTPD;
  for (int i = 0; i < d; ++i)
    for (int k = 0; k < d; ++k) A[i * d + k] = V[i * d + k] * std::sqrt(std::max(lam[k], 0.0));
  if (used_eigen) *used_eigen = 1;
  return PCF_OK;
}

int pcf_mc_basket(const pcf_params* p, const pcf_basket* b, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && (!b || p->assets < 1 || p->assets > PCF_MAX_ASSETS)) s = PCF_EINVAL;
  if (s == PCF_OK && p->replay && p->replay_len < p->N * (long long)p->assets) s = PCF_EINVAL;
  const int d = (s == PCF_OK) ? p->assets : 0;
  double A[PCF_MAX_ASSETS * PCF_MAX_ASSETS];
  double S0[PCF_MAX_ASSETS], sg[PCF_MAX_ASSETS], w[PCF_MAX_ASSETS];
  int eig = 0;
  if (s == PCF_OK) {
    if (b->transform) std::memcpy(A, b->transform, sizeof(double) * d * d);
    else if (b->cov) s = pcf_normal_transform(d, b->cov, A, &eig);
    else s = pcf_chol_equicorr(d, p->rho, A);
  }
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  bool full = eig != 0;
  if (b->transform)
    for (int i = 0; i < d; ++i)
      for (int k = i + 1; k < d; ++k) full = full || A[i * d + k] != 0.0;
  for (int i = 0; i < d; ++i) {
    S0[i] = b->S0 ? b->S0[i] : p->S0;
    sg[i] = b->sigma ? b->sigma[i] : p->sigma;
    w[i] = b->weight ? b->weight[i] : 1.0 / (double)d;
  }
  BasketHost spec{S0, sg, w, full};
  return basket_call(p, A, &spec, out);
}

int pcf_mc_amer(const pcf_params* p, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && p->M <= 0) s = PCF_EINVAL;
  if (s == PCF_OK && (p->N % 2) != 0) s = PCF_EODD_N;  // reference include/common.h:180
  if (s == PCF_OK && p->replay && p->replay_len < (p->N / 2) * (long long)p->M) s = PCF_EINVAL;
  if (s == PCF_OK) {
    const double dt = p->T / p->M, adt = (p->r - p->sigma * p->sigma / 2) * dt;
    s = p->replay ? check_exp_range(adt, p->sigma, replay_absmax(p->replay, (p->N / 2) * (long long)p->M))
                  : check_exp_range(adt, p->sigma * std::sqrt(dt), kHostZMax);
  }
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const double t0 = now_s();
  const long long pairs = p->N / 2;
  std::vector<CallOut> co(g_ctx.size());
  s = for_each_ctx([&](Ctx& c) -> int {
    CallOut& o = co[&c - &g_ctx[0]];
    return guarded_call(c, p->M, [&]() -> int {
      Shard sh = shard_of(pairs, c.rank, c.world);
      const size_t rep_bytes = p->replay ? (((size_t)sh.size() * p->M * 8 + 255) / 256) * 256 + 256 : 0;
      PCF_TRY(ctx_reserve(c, rep_bytes + amer_workspace_bytes(sh.size(), p->M)));
      const double* d_rep = nullptr;
      if (p->replay) PCF_TRY(upload_replay(c, p->replay, sh.begin * p->M, sh.size() * p->M, 0, &d_rep));
      PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
      PCF_TRY(run_mc_amer(c, *p, sh, d_rep, rep_bytes));
      PCF_TRY(finish_call(c, 2, o.vals, &o.seconds_kernel, &o.flag));
      o.launches = c.launches;
      return PCF_OK;
    });
  });
  if (s == PCF_OK) {
    for (auto& c : co)
      if (c.flag) s = c.flag;  // PCF_ESINGULAR raised on the device
  }
  if (s == PCF_OK) {
    fill_mc_result(out, co, 1.0, p->N, p->N * (long long)p->M, t0);
    // mc_amer.cpp:113: floored at the immediate-exercise value
    out->price = std::max(payoff(p->S0, p->E, p->cp), out->sum / (double)p->N);
  }
  out->status = s;
  return s;
}

int pcf_binom_embar(const pcf_params* p, pcf_result* out) {
  int s = check_common(p, out);
  if (s == PCF_OK && p->N > 2147483647LL) s = PCF_EINVAL;  // reference N is an int
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const double t0 = now_s();
  const long long until = (p->N % 2 != 0) ? (p->N + 1) / 2 : p->N / 2;  // binom_embar.cpp:31-33
  long long lo = 0;
  if (p->flags & PCF_FLAG_BINOM_WINDOW) {
    // Hoeffding: P(|X - Np| >= t) <= 2 exp(-2 t^2 / N); a term below 2^-1075 is exactly 0 in FP64.
    // Pairs (i, N-i) with i < min(Np, Nq) - W carry only such terms.
    double u, d, pp, q;
    binom_lattice(p->r, p->sigma, p->T, p->N, u, d, pp, q);
    double W = 19.4 * std::sqrt((double)p->N) + 2.0;
    double m = std::min(pp, q) * (double)p->N - W;
    if (m > 0) lo = std::min((long long)m, until);
  }
  std::vector<CallOut> co(g_ctx.size());
  s = for_each_ctx([&](Ctx& c) -> int {
    CallOut& o = co[&c - &g_ctx[0]];
    return guarded_call(c, 1, [&]() -> int {
      Shard sh = shard_of(until - lo, c.rank, c.world);
      sh.begin += lo; sh.end += lo;
      PCF_CUDA(cudaEventRecord(c.ev0, c.stream));
      const PeerLink l = next_link(c);
      PCF_TRY(run_binom(c, *p, sh, (p->N % 2 == 0) && c.rank == 0, l));
      PCF_TRY(finish_call(c, 1, o.vals, &o.seconds_kernel, &o.flag));
      o.launches = c.launches;
      return PCF_OK;
    });
  });
  if (s == PCF_OK) {
    out->sum = co[0].vals[0];
    out->price = std::exp(-p->r * p->T) * out->sum;  // binom_embar.cpp:49
    out->n = p->N + 1;
    out->units = p->N + 1;
    for (auto& c : co) {
      if (c.seconds_kernel > out->seconds_kernel) out->seconds_kernel = c.seconds_kernel;
      out->launches += c.launches;
    }
    out->gpus = g_ctx[0].world;
    out->seconds_total = now_s() - t0;
  }
  out->status = s;
  return s;
}

// Backward-induction trees. The layers are a serial chain of small, L2-resident rows: the path does not
// shard, so a multi-GPU job runs it on its first GPU only (every rank of a one-process-per-GPU job computes its own
// replica: there is nothing to exchange).
static int binom_tree_call(const pcf_params* p, pcf_result* out, bool american) {
  int s = check_common(p, out);
  if (s == PCF_OK && p->N > 10000000LL) s = PCF_EINVAL;  // O(N^2): 5e13 node updates at the cap
  if (s != PCF_OK) { if (out) out->status = s; return s; }
  const double t0 = now_s();
  Ctx& c = g_ctx[0];
  CallOut o;
  s = [&]() -> int {
    PCF_CUDA(cudaSetDevice(c.device));
    c.launches = 0;
    PCF_TRY(run_binom_tree(c, *p, american));  // records c.ev0 after the pow tables are resident; root -> c.h_res (pinned)
    PCF_CUDA(cudaEventRecord(c.ev1, c.stream));
    PCF_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    PCF_CUDA(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    o.vals[0] = c.h_res->vals[0];
    o.seconds_kernel = ms * 1e-3;
    o.launches = c.launches;
    return PCF_OK;
  }();
  if (s == PCF_OK) {
    out->price = o.vals[0];  // binom_vanilla_*.cpp: `return v_ij[0]` (the discounting is inside the recurrence)
    out->sum = o.vals[0];
    out->n = p->N + 1;
    out->units = p->N * (p->N + 1) / 2;  // node updates
    out->seconds_kernel = o.seconds_kernel;
    out->launches = o.launches;
    out->gpus = 1;
    out->seconds_total = now_s() - t0;
  }
  out->status = s;
  return s;
}

int pcf_binom_vanilla_eur(const pcf_params* p, pcf_result* out) { return binom_tree_call(p, out, false); }
int pcf_binom_vanilla_amer(const pcf_params* p, pcf_result* out) { return binom_tree_call(p, out, true); }

// ------------------------------------------------------------------------------------------------
int pcf_normal_stream(unsigned long long seed, unsigned int stream, unsigned long long index0,
                      long long count, int T, double scale, double* out_host) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  if (count < 0 || T <= 0 || !out_host) return PCF_EINVAL;
  Ctx& c = g_ctx[0];
  PCF_CUDA(cudaSetDevice(c.device));
  size_t bytes = (size_t)count * T * sizeof(double);
  double* d = nullptr;
  PCF_CUDA(cudaMalloc(&d, bytes ? bytes : 8));
  int s = run_normal_stream(c, seed, stream, index0, count, T, scale, d);
  if (s == PCF_OK) {
    cudaError_t e = cudaMemcpyAsync(out_host, d, bytes, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); s = PCF_ECUDA; }
  }
  cudaFree(d);
  return s;
}

int pcf_philox4x32_10(const unsigned int ctr[4], const unsigned int key[2], unsigned int out[4]) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  Ctx& c = g_ctx[0];
  PCF_CUDA(cudaSetDevice(c.device));
  uint32_t* d = (uint32_t*)(c.d_out + 32);
  PCF_TRY(run_philox_kat(c, ctr, key, d));
  PCF_CUDA(cudaMemcpyAsync(out, d, 16, cudaMemcpyDeviceToHost, c.stream));
  PCF_CUDA(cudaStreamSynchronize(c.stream));
  return PCF_OK;
}


int pcf_fp64_peak(double seconds_target, double* dfma_per_sec) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  PCF_CUDA(cudaSetDevice(g_ctx[0].device));
  return run_fp64_peak(g_ctx[0], seconds_target, dfma_per_sec);
}

int pcf_hbm_peak(long long bytes, double* bytes_per_sec) {
  if (g_ctx.empty()) return PCF_ENOINIT;
  PCF_CUDA(cudaSetDevice(g_ctx[0].device));
  return run_hbm_peak(g_ctx[0], bytes, bytes_per_sec);
}

int pcf_device_info(char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                    long long* mem_bytes) {
  int dev = g_ctx.empty() ? 0 : g_ctx[0].device;
  cudaDeviceProp prop;
  PCF_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (name && name_len > 0) {
    std::strncpy(name, prop.name, (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (mem_bytes) *mem_bytes = (long long)prop.totalGlobalMem;
  return PCF_OK;
}

const char* pcf_strerror(int status) {
  switch (status) {
    case PCF_OK: return "ok";
    case PCF_EINVAL_PAYOFF: return "Unknown payoff function";
    case PCF_EODD_N: return "N needs to be divisible by 2 for finding paths";
    case PCF_ESINGULAR: return "Detereminant is not > 0";
    case PCF_EINVAL: return "invalid argument";
    case PCF_ENOTPD: return "correlation matrix is not positive definite";
    case PCF_ECUDA: return "CUDA failure (no device, or runtime error)";
    case PCF_ENCCL: return "NCCL failure";
    case PCF_ENOINIT: return "pcf_init() has not been called";
    case PCF_ENOMEM: return "device memory allocation failed";
    default: return "unknown status";
  }
}

const char* pcf_last_error(void) {
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_last_error;
  return copy.c_str();
}

}  // extern "C"
