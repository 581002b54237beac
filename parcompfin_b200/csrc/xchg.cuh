// xchg.cuh -- moment exchange over NVLink peer memory, fused into the producing and consuming kernels.
//
// Replaces the reference's MPI_Reduce / MPI_Allreduce / MPI_Bcast of 1-3 doubles (src/mc_eur_mpi.cpp:36,
// src/mc_amer_mpi.cpp:82-92,131). Payloads are <= 64 B, so the cost is pure latency: instead of a separate
// collective launch, the LAST BLOCK of the reducing kernel stores its K sums straight into a mailbox in every
// peer GPU's memory (P2P stores through NVSwitch), fences at system scope and raises a per-source sequence flag;
// the consuming kernel (the American decision kernel, or a one-warp finisher for the end-of-run moments) spins on
// its LOCAL flags and adds the K x world values in rank order -- bit-identical on every GPU, no host round trip,
// no extra kernel between the moments and decision passes.
// Mailboxes are double-buffered by sequence parity: a GPU cannot run two exchanges ahead because each consumer
// needs every peer's flag for the current one.
#pragma once
#include <cstdint>

namespace pcf {

constexpr int kMaxWorld = 8;
constexpr int kXchgVals = 8;

struct Mailbox {
  double vals[2][kMaxWorld][kXchgVals];       // [sequence parity][source rank][k]
  unsigned long long flags[2][kMaxWorld];     // sequence number published by the source after its vals
  int error;                                  // set when a spin wait times out
};

struct PeerLink {
  Mailbox* peer[kMaxWorld];  // peer[r]: rank r's mailbox as mapped into THIS device's address space (peer[rank] local)
  int rank, world;
  unsigned long long seq;    // sequence number of this exchange (host-incremented, identical on every rank)
};

// Called by all threads of one block; `vals` (K doubles) must be readable by every thread (shared memory).
template <int K>
__device__ __forceinline__ void peer_publish(const PeerLink& L, const double* vals) {
  const int slot = (int)(L.seq & 1ull);
  const int t = threadIdx.x;
  if (t < L.world * K) {
    const int r = t / K, k = t - r * K;
    volatile double* dst = &L.peer[r]->vals[slot][L.rank][k];
    *dst = vals[k];
    __threadfence_system();
  }
  __syncthreads();
  if (t < L.world) {
    __threadfence_system();
    volatile unsigned long long* f = &L.peer[t]->flags[slot][L.rank];
    *f = L.seq;
  }
}

// Called by all threads of one block (>= 32 threads); out[K] in shared memory. Ends with __syncthreads().
template <int K>
__device__ __forceinline__ void peer_gather(const PeerLink& L, double* out) {
  const int slot = (int)(L.seq & 1ull);
  Mailbox* me = L.peer[L.rank];
  const int t = threadIdx.x;
  if (t < L.world) {
    volatile unsigned long long* f = &me->flags[slot][t];
    unsigned long long spins = 0;
    while (*f != L.seq) {
      if (++spins > (1ull << 31)) {  // a peer died: ~10 s at 1.9 GHz; report instead of hanging the GPU
        me->error = 1;
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (t < K) {
    double s = 0.0;
    for (int r = 0; r < L.world; ++r) s = __dadd_rn(s, *(volatile double*)&me->vals[slot][r][t]);
    out[t] = s;
  }
  __syncthreads();
}

// Same, called by ONE full warp (warp-specialised kernels whose other warps are busy); ends with __syncwarp().
template <int K>
__device__ __forceinline__ void peer_gather_warp(const PeerLink& L, double* out) {
  static_assert(K <= 32, "one lane per value");
  const int slot = (int)(L.seq & 1ull);
  Mailbox* me = L.peer[L.rank];
  const int t = threadIdx.x & 31;
  if (t < L.world) {
    volatile unsigned long long* f = &me->flags[slot][t];
    unsigned long long spins = 0;
    while (*f != L.seq) {
      if (++spins > (1ull << 31)) {
        me->error = 1;
        break;
      }
    }
    __threadfence_system();
  }
  __syncwarp();
  if (t < K) {
    double s = 0.0;
    for (int r = 0; r < L.world; ++r) s = __dadd_rn(s, *(volatile double*)&me->vals[slot][r][t]);
    out[t] = s;
  }
  __syncwarp();
}

}  // namespace pcf
