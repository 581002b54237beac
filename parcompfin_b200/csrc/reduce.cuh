// reduce.cuh -- compensated FP64 moment reduction: thread (Neumaier) -> warp shuffle -> block ->
// grid (last-block-done, fixed order => bit-reproducible for a fixed launch shape).
// Replaces the reference's serial `result += payoff(...)` (src/mc_eur.cpp:23-25) and its
// OpenMP `reduction(+:result)` / MPI_Reduce (src/mc_eur_omp.cpp, src/mc_eur_mpi.cpp:36).
#pragma once
#include <cstdint>
#include "xchg.cuh"

namespace pcf {

// value = hi + lo, |lo| << |hi|
struct Comp {
  double hi, lo;
  __device__ __forceinline__ Comp() : hi(0.0), lo(0.0) {}
  __device__ __forceinline__ Comp(double h, double l) : hi(h), lo(l) {}
  // Neumaier / TwoSum accumulation of one term (branch-free Knuth TwoSum: 6 adds)
  __device__ __forceinline__ void add(double x) {
    double s = __dadd_rn(hi, x);
    double bp = __dadd_rn(s, -hi);
    double e = __dadd_rn(__dadd_rn(hi, -__dadd_rn(s, -bp)), __dadd_rn(x, -bp));
    hi = s;
    lo = __dadd_rn(lo, e);
  }
  __device__ __forceinline__ void merge(const Comp& o) {
    add(o.hi);
    lo = __dadd_rn(lo, o.lo);
  }
  __device__ __forceinline__ double value() const { return __dadd_rn(hi, lo); }
};

// Cheap blocked accumulator for hot loops: plain adds into `run`, folded into the compensated
// total every kFold terms (1 + 7/kFold adds per term instead of 7).
template <int kFold>
struct BlockedComp {
  Comp total;
  double run;
  int cnt;
  __device__ __forceinline__ BlockedComp() : run(0.0), cnt(0) {}
  __device__ __forceinline__ void add(double x) {
    run += x;
    if (++cnt == kFold) {
      total.add(run);
      run = 0.0;
      cnt = 0;
    }
  }
  __device__ __forceinline__ Comp finish() {
    total.add(run);
    run = 0.0;
    cnt = 0;
    return total;
  }
};

__device__ __forceinline__ double shfl_down_f64(double v, int delta) {
  return __shfl_down_sync(0xffffffffu, v, delta);
}

__device__ __forceinline__ Comp warp_reduce(Comp v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    Comp o(shfl_down_f64(v.hi, d), shfl_down_f64(v.lo, d));
    v.merge(o);
  }
  return v;
}

// Reduces K compensated values per thread over the block; result valid in thread 0.
// smem: K * 2 * 32 doubles.
template <int K>
__device__ __forceinline__ void block_reduce(Comp (&v)[K], double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    v[k] = warp_reduce(v[k]);
    if (lane == 0) {
      smem[(k * 32 + warp) * 2 + 0] = v[k].hi;
      smem[(k * 32 + warp) * 2 + 1] = v[k].lo;
    }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      Comp w = (lane < nwarp) ? Comp(smem[(k * 32 + lane) * 2], smem[(k * 32 + lane) * 2 + 1]) : Comp();
      v[k] = warp_reduce(w);
    }
  }
}

// Grid-level finish. Every block writes its K (hi,lo) partials to `partials[block][k]`; the last
// block to arrive (ticket counter) folds all partials in block order and writes out[k] = hi+lo.
// `ticket` must be zero on entry and is reset to zero on exit, so the buffer is reusable across
// launches on the same stream. smem as for block_reduce.
// When `link` names a multi-GPU job (world > 1) the last block also publishes the K sums to every peer's mailbox
// (xchg.cuh): the reduction and the collective are one kernel. With link->gather set it then waits for every rank's
// publication and overwrites out[k] with the job-wide sums (added in rank order: bit-identical on every GPU) -- the
// end-of-run all-reduce of the reference (src/mc_eur_mpi.cpp:36) without a second launch. `out` may be host-mapped
// pinned memory: the host reads the result right after the stream synchronises, no device-to-host copy.
template <int K>
__device__ __forceinline__ void grid_reduce(Comp (&v)[K], double* smem, double* partials,
                                            unsigned int* ticket, double* out, const PeerLink* link = nullptr) {
  block_reduce<K>(v, smem);
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      partials[((size_t)blockIdx.x * K + k) * 2 + 0] = v[k].hi;
      partials[((size_t)blockIdx.x * K + k) * 2 + 1] = v[k].lo;
    }
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Comp acc[K];
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const volatile double* p = partials + ((size_t)b * K + k) * 2;
      acc[k].merge(Comp(p[0], p[1]));
    }
  }
  __syncthreads();  // smem reuse
  block_reduce<K>(acc, smem);
  __syncthreads();  // smem reuse below
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double r = acc[k].value();
      out[k] = r;
      smem[k] = r;
    }
    *ticket = 0u;
  }
  if (link != nullptr && link->world > 1) {
    __syncthreads();
    peer_publish<K>(*link, smem);
    if (link->gather) {
      __syncthreads();
      peer_gather<K>(*link, smem + 16);
      if ((int)threadIdx.x < K) out[threadIdx.x] = smem[16 + threadIdx.x];
    }
  }
}

}  // namespace pcf
