// xchg.cuh -- moment exchange over NVLink peer memory, fused into the producing and consuming kernels.
//
// Replaces the reference's MPI_Reduce / MPI_Allreduce / MPI_Bcast of 1-3 doubles (src/mc_eur_mpi.cpp:36,
// src/mc_amer_mpi.cpp:82-92,131). Payloads are <= 64 B, so the cost is pure latency: instead of a separate
// collective launch, the LAST BLOCK of the reducing kernel stores its K sums straight into a mailbox in every
// peer GPU's memory (P2P stores through NVSwitch), fences at system scope and raises a per-source sequence flag;
// the consumer (the American sweep kernel at its next exercise date, or the same last block for the end-of-run
// moments) spins on its LOCAL flags and adds the K x world values in rank order -- bit-identical on every GPU, no
// host round trip, no extra kernel.
// Mailboxes are double-buffered by sequence parity: a GPU cannot run two exchanges ahead because each consumer
// needs every peer's flag for the current one.
//
// Failure handling: every wait is bounded by the GPU's nanosecond timer (kXchgTimeoutNs). A mailbox carries a POISON
// RANGE of sequence numbers [poison_lo, poison_hi]: the exchanges of one failed call. A peer whose host failed before
// it could launch its kernels writes the range of that call into every other rank's mailbox (xchg_poison_kernel); a
// waiter that times out writes the range of its own call into its own mailbox. A wait whose sequence number lies in the
// range falls through at once, sets the host-mapped status word and yields NaN sums -- never a sum over a stale slot --
// so the call ends promptly with PCF_ENCCL on every rank, and because every rank consumes the same sequence numbers
// whether or not its call succeeded, the next call is unaffected (nothing has to be cleared).
#pragma once
#include <cstdint>

namespace pcf {

constexpr int kMaxWorld = 8;
constexpr int kXchgVals = 8;
constexpr unsigned long long kXchgTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s

struct Mailbox {
  double vals[2][kMaxWorld][kXchgVals];       // [sequence parity][source rank][k]
  unsigned long long flags[2][kMaxWorld];     // sequence number published by the source after its vals
  unsigned long long poison_lo, poison_hi;    // exchanges lo..hi belong to a failed call (0, 0: none)
};

struct PeerLink {
  Mailbox* peer[kMaxWorld];  // peer[r]: rank r's mailbox as mapped into THIS device's address space (peer[rank] local)
  int rank, world;
  unsigned long long seq;    // sequence number of this exchange (host-incremented, identical on every rank)
  unsigned long long call_first, call_last;  // sequence numbers the current call owns
  int* host_err;             // host-mapped status word of this context (written on timeout / poison)
  int gather;                // 1: the publishing block also waits for every rank and writes the job-wide sums
};

__device__ __forceinline__ unsigned long long xchg_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The exchange functions take the sequence number separately: the American sweep kernel of date m publishes exchange
// `seq` and gathers exchange `seq - 1` with one PeerLink parameter.
__device__ __forceinline__ bool xchg_poisoned(const PeerLink& L, unsigned long long seq) {
  Mailbox* me = L.peer[L.rank];
  const unsigned long long lo = *(volatile unsigned long long*)&me->poison_lo;
  const unsigned long long hi = *(volatile unsigned long long*)&me->poison_hi;
  return lo != 0ull && seq >= lo && seq <= hi;
}

// Spins until *f == seq. Returns false on timeout or poison (and records it).
__device__ __forceinline__ bool xchg_wait_flag(const PeerLink& L, unsigned long long seq, volatile unsigned long long* f,
                                               unsigned int backoff_ns = 0) {
  if (*f == seq) return true;
  Mailbox* me = L.peer[L.rank];
  const unsigned long long t0 = xchg_now_ns();
  unsigned int spins = 0;
  while (*f != seq) {
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((++spins & 255u) == 1u) {
      const bool late = xchg_now_ns() - t0 > kXchgTimeoutNs;
      if (late || xchg_poisoned(L, seq)) {
        if (late) {  // the rest of this call's waits fall through at once
          *(volatile unsigned long long*)&me->poison_hi = L.call_last;
          __threadfence();
          *(volatile unsigned long long*)&me->poison_lo = L.call_first;
        }
        if (L.host_err) *(volatile int*)L.host_err = 1;
        return false;
      }
    }
  }
  return true;
}

// Called by all threads that take part in `sync` (a whole block with __syncthreads, or the consumer warps of a
// warp-specialised kernel with their named barrier; at least world*K threads, numbered from threadIdx.x == 0);
// `vals` (K doubles) must be readable by every one of them (shared memory).
template <int K, class Sync>
__device__ __forceinline__ void peer_publish(const PeerLink& L, unsigned long long seq, const double* vals, Sync sync) {
  const int slot = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < L.world * K) {
    const int r = t / K, k = t - r * K;
    volatile double* dst = &L.peer[r]->vals[slot][L.rank][k];
    *dst = vals[k];
  }
  sync();
  if (t < L.world) {
    // release: the barrier orders every thread's value stores before this fence, and fences are cumulative, so one
    // system-scope fence in the flag writer covers them all
    __threadfence_system();
    volatile unsigned long long* f = &L.peer[t]->flags[slot][L.rank];
    *f = seq;
  }
}
template <int K>
__device__ __forceinline__ void peer_publish(const PeerLink& L, const double* vals) {
  peer_publish<K>(L, L.seq, vals, [] { __syncthreads(); });
}

// One warp (lanes 0..31 of the caller) waits for every rank's publication and adds the K values in rank order into
// out[0..K) (shared or global memory). Failure -> NaN. Ends with __syncwarp().
template <int K>
__device__ __forceinline__ void peer_gather_warp(const PeerLink& L, unsigned long long seq, double* out,
                                                 unsigned int backoff_ns = 0) {
  static_assert(K <= 32, "one lane per value");
  const int slot = (int)(seq & 1ull);
  Mailbox* me = L.peer[L.rank];
  const int t = threadIdx.x & 31;
  bool ok = true;
  if (t < L.world) {
    ok = xchg_wait_flag(L, seq, &me->flags[slot][t], backoff_ns);
    __threadfence_system();
  }
  ok = __all_sync(0xffffffffu, ok);
  if (t < K) {
    double s = 0.0;
    for (int r = 0; r < L.world; ++r) s = __dadd_rn(s, *(volatile double*)&me->vals[slot][r][t]);
    out[t] = ok ? s : __longlong_as_double(0x7ff8000000000000LL);
  }
  __syncwarp();
}

// Same, called by all threads of one block (>= 32 threads); ends with __syncthreads().
template <int K>
__device__ __forceinline__ void peer_gather(const PeerLink& L, double* out) {
  if (threadIdx.x < 32) peer_gather_warp<K>(L, L.seq, out);
  __syncthreads();
}

}  // namespace pcf
