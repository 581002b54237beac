// common.cuh -- per-GPU context, error plumbing and shard arithmetic shared by the method files.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../../include/pcf.h"
#include "fastmath.cuh"
#include "xchg.cuh"

namespace pcf {

void set_last_error(const std::string& s);

// Launch-shape A/B knobs exist only in a -DPCF_TUNING build (`make lib TUNING=1`, used by tools/tune_*.py). The shipped
// library has one instantiation per kernel family and reads no environment variable on a pricing call.
#ifdef PCF_TUNING
inline const char* tuning_env(const char* name) { return getenv(name); }
#else
inline const char* tuning_env(const char*) { return nullptr; }
#endif

#define PCF_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::pcf::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
      return (_e == cudaErrorMemoryAllocation) ? PCF_ENOMEM : PCF_ECUDA;                      \
    }                                                                                         \
  } while (0)

#define PCF_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != PCF_OK) return _s; \
  } while (0)

constexpr int kMaxBlocks = 148 * 16;  // upper bound on any reduction grid
constexpr int kMaxMoments = 8;

struct HostOut {
  double vals[8];
  int flag;        // device-raised status (PCF_ESINGULAR), 0 = none
  int peer_error;  // the peer-memory exchange timed out or was poisoned
};

// One context = one GPU of the job (one per process under torchrun, `gpus` of them in the
// single-process front ends).
struct Ctx {
  int device = 0;
  int rank = 0, world = 1;   // position in the job
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  void* comm = nullptr;      // ncclComm_t, null when world == 1
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double* d_partials = nullptr;    // kMaxBlocks * kMaxMoments * 2 doubles
  unsigned int* d_ticket = nullptr;
  double* d_out = nullptr;         // small device scratch vector (64 doubles): per-date moments of mc_amer, KAT output
  const MathTables* d_tables = nullptr;  // fastmath.cuh lookup tables (global memory; staged to smem per block)
  // Results and status words live in host-MAPPED pinned memory: the last block of the reducing kernel writes them
  // straight into host memory, so a call ends with one stream synchronise and no device-to-host copy.
  HostOut* h_res = nullptr;        // host view
  double* res_dev = nullptr;       // device alias of h_res->vals
  int* flag_dev = nullptr;         // device alias of h_res->flag  (PCF_ESINGULAR raised by the sweep kernel)
  int* perr_dev = nullptr;         // device alias of h_res->peer_error (xchg.cuh: timeout / poisoned mailbox)
  void* workspace = nullptr;       // grow-only scratch (mc_amer path store, replay streams)
  size_t workspace_bytes = 0;
  int launches = 0;                // kernels launched in the current call
  // NVLink peer-memory exchange (xchg.cuh). peer_ok: every rank's mailbox is mapped here; otherwise the job
  // falls back to ncclAllReduce on the compute stream.
  Mailbox* mailbox = nullptr;
  PeerLink link{};
  bool peer_ok = false;
  unsigned long long xchg_seq = 0;
  unsigned long long call_first = 0, call_last = 0;  // sequence numbers owned by the call in flight
};

// Link for the next exchange of this call sequence (every rank issues the same sequence of exchanges).
// Without peer mapping the returned link has world == 1, which turns publishing and gathering off.
// `gather`: the publishing block also collects every rank's sums (end-of-run moments).
inline PeerLink next_link(Ctx& c, bool gather = true) {
  PeerLink l = c.link;
  l.host_err = c.perr_dev;
  l.gather = gather ? 1 : 0;
  l.call_first = c.call_first;
  l.call_last = c.call_last;
  if (c.world > 1 && c.peer_ok) {
    l.seq = ++c.xchg_seq;
  } else {
    l.world = 1;
    l.seq = 0;
  }
  return l;
}
inline bool use_peer(const Ctx& c) { return c.world > 1 && c.peer_ok; }
// Where a method's last block writes its end-of-run sums: host-mapped memory, except on the NCCL fallback path, whose
// all-reduce runs in place on a device buffer.
inline double* final_out(Ctx& c) { return (c.world > 1 && !c.peer_ok) ? c.d_out : c.res_dev; }
int launch_xchg_poison(Ctx& c);  // raises `error` in every peer's mailbox

int ctx_reserve(Ctx& c, size_t bytes);  // ensures c.workspace >= bytes
int allreduce_sum(Ctx& c, double* d_buf, int count);  // in place, on c.stream; no-op if world == 1

// Host-side description of a general basket, all arrays of length d
struct BasketHost {
  const double* S0;
  const double* sigma;
  const double* weight;
  bool full;  // the transform is a full matrix (eigen fallback), not a lower triangle
};

// Contiguous block partition of `units` over the job: rank g takes
// [g*ceil(U/G), min(U,(g+1)*ceil(U/G))).
struct Shard {
  long long begin, end;
  __host__ __device__ long long size() const { return end > begin ? end - begin : 0; }
};
inline Shard shard_of(long long units, int rank, int world) {
  long long per = (units + world - 1) / world;
  long long b = per * rank, e = b + per;
  if (b > units) b = units;
  if (e > units) e = units;
  return Shard{b, e};
}

// Reduction-grid sizing: persistent-style grids in multiples of the SM count.
inline int grid_for(const Ctx& c, long long work_items, int block, int blocks_per_sm) {
  long long need = (work_items + block - 1) / block;
  long long cap = (long long)c.sm_count * blocks_per_sm;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// reference include/common.h:56-60
__host__ __device__ __forceinline__ double payoff(double St, double E, int cp) {
  double v = (double)cp * (St - E);
  return v > 0.0 ? v : 0.0;
}

}  // namespace pcf
