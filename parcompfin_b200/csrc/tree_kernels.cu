// tree_kernels.cu -- the backward-induction binomial trees, sm_100a FP64.
//   reference src/binom_vanilla_eur.cpp:15-41 (European) and src/binom_vanilla_amer.cpp:15-42 (American):
//     v_N[i] = payoff(S0 u^i d^(N-i));   v_n[i] = (p v_{n+1}[i+1] + q v_{n+1}[i]) / R   [, max with payoff(S0 u^i d^(n-i))]
//   The run-scripts call them right before every Monte Carlo sweep to produce `comparison`
//   (runscript_mc_eur.sh:23, runscript_mc_amer.sh:24); O(N^2), 45 s on the CPU at N = 1e5.
//
// Parity contract: every node value is formed with the reference's operations in the reference's order --
// p*v[i+1] and q*v[i] rounded separately, their sum rounded, the quotient by R rounded (x86-64 has no FMA
// contraction in the reference build) -- so the root equals the reference's to the last bit:
//   * u^i and d^j are TABLES computed on the host with the same glibc pow() the reference calls per node
//     (2(N+1) calls instead of N^2), and S is formed as (S0*pow(u,i))*pow(d,n-i), the reference's association;
//   * the division by the constant R is q0 = RN(x*z), r = fma(-q0, R, x), q = fma(r, z, q0) with z = RN(1/R):
//     the value rounded last is (x/R)(1 - delta*eps), |delta*eps| <= 2^-105, i.e. the correctly rounded
//     quotient unless x/R lies within 2^-105 (relative) of a rounding boundary -- probability ~2^-51 per node,
//     ~1e-6 per N = 1e5 tree -- or x is subnormal. tests/test_gpu_parity.py states 1e-13 relative and observes
//     bit equality.
//
// Parallelisation: time-skewed (trapezoid) tiling with no intra-step synchronisation. A warp owns 32*kR
// consecutive nodes of a layer in registers (kR per lane), advances them kSteps layers -- per layer one shuffle
// brings the right neighbour's first value -- and writes the 32*kR - kSteps left-most nodes, which are the ones
// whose whole dependency cone was inside the warp. One launch = kSteps layers of the whole tree; layers ping-pong
// between two HBM (L2-resident) buffers. The American tree also needs d^(n-i): per warp the window of the d-power
// table that its cone touches is staged in shared memory (padded against bank conflicts) and every lane keeps a
// kR-entry sliding window of it in registers, refilled with ONE shared load per layer.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <type_traits>
#include <vector>
#include "common.cuh"
#include "binom_math.cuh"

namespace pcf {

struct TreeArgs {
  const double* vin;   // layer n0: entries 0..n0
  double* vout;        // layer n0 - steps
  const double* pu;    // pow(u, i), i = 0..N   (American)
  const double* pd;    // pow(d, j), j = 0..N   (American)
  long long n0;
  int steps;           // <= kSteps
  int trig;            // programmatic dependent launch: 1 = release the next launch at kernel start, 2 = after the layers
  int stride;          // tree_cta_kernel: nodes a CTA finishes per launch
  double p, q, R, z;   // z = RN(1/R)
  double S0, sgn, nE;  // payoff(S) = max(fma(sgn, S, nE), 0) = max(cp*(S - E), 0)
};

// reference: (p*v[i+1] + q*v[i])/R, binom_vanilla_eur.cpp:35 / binom_vanilla_amer.cpp:34
__device__ __forceinline__ double tree_node(double lo, double hi, const TreeArgs& a) {
  const double x = __dadd_rn(__dmul_rn(a.p, hi), __dmul_rn(a.q, lo));
  const double q0 = __dmul_rn(x, a.z);
  const double r = fma(-q0, a.R, x);
  return fma(r, a.z, q0);
}

constexpr int kTreeWarps = 4;  // warps per CTA (independent of each other)

template <int kR, int kSteps, bool kAmer>
__global__ void __launch_bounds__(kTreeWarps * 32) tree_steps_kernel(TreeArgs a) {
  constexpr int kL = 32 * kR;            // nodes per warp
  constexpr int kStride = kL - kSteps;   // nodes a warp finishes
  constexpr int kWin = kL + kSteps;      // d-power window per warp (unpadded entries)
  constexpr int kPad = kWin + kWin / 8 + 8;
  __shared__ double s_pd[kAmer ? kTreeWarps * kPad : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kTreeWarps + warp;
  const long long base = gw * kStride;
  const long long n_out = a.n0 - a.steps;
  if (base > n_out) return;  // warps are independent: no block-level barrier below
  const long long i0 = base + (long long)lane * kR;
  // Programmatic dependent launch: let the next launch of the chain be scheduled now (its CTAs park in
  // griddepcontrol.wait), stage everything that does not depend on the previous launch, then wait for the previous
  // launch's layer to be complete and visible. Without the launch attribute both instructions are no-ops.
  if (a.trig == 1) asm volatile("griddepcontrol.launch_dependents;");

  double v[kR], A[kR], W[kR];
  // d-power window: entry k holds pd[lo + k], lo = n0 - kSteps - base - kL + 1 (clamped reads below 0 are never used
  // by a node inside the tree); padded index k + (k >> 3) makes the lane stride 9 doubles (conflict-free LDS.64)
  double* my_pd = s_pd + (kAmer ? warp * kPad : 0);
  const long long lo = a.n0 - kSteps - base - kL + 1;
  if (kAmer) {
#pragma unroll
    for (int j = 0; j < kR; ++j) {
      const long long i = i0 + j;
      A[j] = (i <= a.n0) ? __dmul_rn(a.S0, a.pu[i]) : 0.0;  // S0*pow(u,i), binom_vanilla_amer.cpp:33
    }
    for (int k = lane; k < kWin; k += 32) {
      const long long idx = lo + k;
      my_pd[k + (k >> 3)] = (idx >= 0 && idx <= a.n0) ? a.pd[idx] : 0.0;
    }
    __syncwarp();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int j = 0; j < kR; ++j) {
    const long long i = i0 + j;
    v[j] = (i <= a.n0) ? a.vin[i] : 0.0;
  }
  // node (i0 + j) of the layer reached after s+1 steps (n = n0 - 1 - s) needs pd[n - i0 - j] = window entry
  // k(s, j) = kSteps + kL - 2 - lane*kR - (s + j): it depends on s + j only, so the lane keeps entries t = s..s+kR-1 in
  // W[t % kR] and replaces the one that falls out of the window with ONE shared load per step.
  const int kbase = kSteps + kL - 2 - lane * kR;
  if (kAmer) {
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      const int k = kbase - t;
      W[t] = my_pd[k + (k >> 3)];
    }
  }
  const int groups = (a.steps + kR - 1) / kR;
  for (int g = 0; g < groups; ++g) {
#pragma unroll
    for (int kk = 0; kk < kR; ++kk) {
      const int s = g * kR + kk;
      if (s < a.steps) {  // warp-uniform
        const double halo = __shfl_down_sync(0xffffffffu, v[0], 1);  // lane 31: outside the cone, never written back
#pragma unroll
        for (int j = 0; j < kR; ++j) {
          const double hi = (j + 1 < kR) ? v[j + 1] : halo;
          double nv = tree_node(v[j], hi, a);
          if (kAmer) {
            // payoff(S0*pow(u,i)*pow(d,n-i)) and std::max(jatk, sij), binom_vanilla_amer.cpp:33-35
            const double c = fma(a.sgn, __dmul_rn(A[j], W[(kk + j) % kR]), a.nE);
            const double sij = c > 0.0 ? c : 0.0;
            nv = (nv < sij) ? sij : nv;
          }
          v[j] = nv;
        }
        if (kAmer) {
          const int k = kbase - (s + kR);  // entry t = s + kR replaces t = s
          W[kk] = (k >= 0) ? my_pd[k + (k >> 3)] : 0.0;
        }
      }
    }
  }
  if (a.trig == 2) asm volatile("griddepcontrol.launch_dependents;");
#pragma unroll
  for (int j = 0; j < kR; ++j) {
    const long long i = i0 + j;
    if (lane * kR + j < kStride && i <= n_out) a.vout[i] = v[j];
  }
}

// f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{}): a loop whose index is a constant
// expression inside the body
template <int N, int I = 0, typename F>
__device__ __forceinline__ void unrolled(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    unrolled<N, I + 1>(f);
  }
}

// CTA-cooperative trapezoid (see the header). Geometry, all in CTA-local node positions c = warp*kOwn + lane*kR + j:
//   kL = 32 kR nodes per warp, kOwn = kL - kH of them owned, kC = kW kOwn + kH nodes per CTA; after s layers the
//   positions [0, kC - s) are exact, so a launch of a.steps <= kK layers finishes a.stride = kC - kK of them.
template <int kR, int kW, int kKsel>
struct CtaShape {
  static constexpr int kH = kR == 3 ? 9 : kR == 6 ? 12 : 8;  // halo layers: a whole number of lanes
  static constexpr int kL = 32 * kR;
  static constexpr int kOwn = kL - kH;
  static constexpr int kC = kW * kOwn + kH;
  static constexpr int kK = kKsel ? kKsel : (kC >= 440 ? 128 : 64);  // layers per launch (0 = default)
  static constexpr int kStride = kC - kK;
};

template <int kR, int kW, int kKsel, bool kAmer>
__global__ void __launch_bounds__(kW * 32) tree_cta_kernel(TreeArgs a) {
  using Sh = CtaShape<kR, kW, kKsel>;
  constexpr int kH = Sh::kH, kOwn = Sh::kOwn, kC = Sh::kC, kK = Sh::kK;
  static_assert(kH % kR == 0 && kH >= kR && kH < 32 * kR, "halo must be a whole number of lanes");
  static_assert(kC > kK && kK >= kH, "a launch must finish at least one node per CTA");
  constexpr int kHaloLanes = kH / kR;
  constexpr int kWin = kK + kC - 1;  // d-power window of the CTA; entry k at s_pd[k + 1], s_pd[0] = 0 stands for k = -1
  __shared__ double s_pd[kAmer ? kWin + 1 : 1];
  __shared__ double s_x[2][kW][kH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = (long long)blockIdx.x * a.stride;
  const long long n_out = a.n0 - a.steps;
  if (base > n_out) return;  // whole CTA: no barrier is skipped by part of a block
  const int c0 = warp * kOwn + lane * kR;
  const long long i0 = base + c0;
  if (a.trig == 1) asm volatile("griddepcontrol.launch_dependents;");

  double v[kR], A[kR], W[kR];
  // window entry k holds pd[lo + k]; node c of the layer reached after s+1 layers (n = n0 - 1 - s) needs
  // pd[n - base - c] = entry (steps - 1 - s) + (kC - 1 - c)
  const long long lo = a.n0 - a.steps - base - kC + 1;
  if (kAmer) {
#pragma unroll
    for (int j = 0; j < kR; ++j) {
      const long long i = i0 + j;
      A[j] = (i <= a.n0) ? __dmul_rn(a.S0, a.pu[i]) : 0.0;  // S0*pow(u,i), binom_vanilla_amer.cpp:33
    }
    const int win = a.steps + kC - 1;
    for (int k = threadIdx.x; k < win; k += kW * 32) {
      const long long idx = lo + k;
      s_pd[k + 1] = (idx >= 0 && idx <= a.n0) ? a.pd[idx] : 0.0;
    }
    if (threadIdx.x == 0) s_pd[0] = 0.0;
    __syncthreads();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
  for (int j = 0; j < kR; ++j) {
    const long long i = i0 + j;
    v[j] = (i <= a.n0) ? a.vin[i] : 0.0;
  }
  // the lane keeps entries t = s .. s+kR-1 (t = s + j) in W[t % kR]; entry t is window index kbase - t >= -1
  // (kC - 1 - c0 >= kR - 1 for every lane, t <= steps - 1 + kR), i.e. wp[-t]: one LDS with an immediate offset per layer
  const int kbase = a.steps - 1 + kC - 1 - c0;
  const double* wp = s_pd + 1 + kbase;
  double sgn = a.sgn, nE = a.nE;
  if (kAmer) {
    asm volatile("" : "+d"(sgn), "+d"(nE));  // keep them in registers: ptxas otherwise reloads them (LDC) per layer
#pragma unroll
    for (int t = 0; t < kR; ++t) W[t] = wp[-t];
  }
  // One layer (s = r kH + ss; kH % kR == 0, so s % kR == ss % kR and the register indices are static).
  // The kR nodes of a lane are independent within a layer. Written node by node, ptxas keeps each node's
  // six-instruction dependent chain contiguous (one scratch register pair, kR x 48 cycles per layer); written
  // stage by stage it issues the kR chains interleaved, which is what keeps the FP64 pipe busy.
  auto layer = [&](int s, auto ss_tag) {
    constexpr int ss = decltype(ss_tag)::value;
    const double halo = __shfl_down_sync(0xffffffffu, v[0], 1);  // lane 31: the warp's halo decays by one node
    double x[kR], q0[kR], ex[kR];
#pragma unroll
    for (int j = 0; j < kR; ++j) x[j] = __dmul_rn(a.q, v[j]);
#pragma unroll
    for (int j = 0; j < kR; ++j) x[j] = __dadd_rn(__dmul_rn(a.p, (j + 1 < kR) ? v[j + 1] : halo), x[j]);
#pragma unroll
    for (int j = 0; j < kR; ++j) q0[j] = __dmul_rn(x[j], a.z);
    if (kAmer) {
      // cp*(S - E) with S = (S0 u^i) d^(n-i), binom_vanilla_amer.cpp:33. The reference takes
      // max(continuation, max(cp*(S - E), 0)); the continuation value is >= 0 whenever p, q >= 0 (the host sends
      // lattices with a negative probability to tree_steps_kernel), so the inner max with 0 changes nothing.
#pragma unroll
      for (int j = 0; j < kR; ++j) ex[j] = fma(sgn, __dmul_rn(A[j], W[(ss + j) % kR]), nE);
    }
#pragma unroll
    for (int j = 0; j < kR; ++j) x[j] = fma(-q0[j], a.R, x[j]);
#pragma unroll
    for (int j = 0; j < kR; ++j) v[j] = fma(x[j], a.z, q0[j]);
    if (kAmer) {
#pragma unroll
      for (int j = 0; j < kR; ++j) v[j] = (v[j] < ex[j]) ? ex[j] : v[j];  // std::max(jatk, sij), :34-35
      W[ss % kR] = wp[-(s + kR)];  // entry t = s + kR replaces t = s
    }
  };
  // kH layers back to back. Full rounds carry no per-layer test, so a round is ONE basic block and the scheduler can
  // start the shuffle-independent nodes of layer s+1 under the tail of layer s; only the last, partial round tests.
  auto round_of = [&](int r, auto guarded) {
    unrolled<kH>([&](auto ss_tag) {
      const int s = r * kH + decltype(ss_tag)::value;
      if (!decltype(guarded)::value || s < a.steps) layer(s, ss_tag);  // block-uniform
    });
  };
  const int rounds = (a.steps + kH - 1) / kH, full = a.steps / kH;
  for (int r = 0; r < rounds; ++r) {
    if (r < full) round_of(r, std::false_type{});
    else round_of(r, std::true_type{});
    if (r + 1 < rounds) {  // refill every warp's halo from its right neighbour's first kH nodes
      if (lane < kHaloLanes) {
#pragma unroll
        for (int j = 0; j < kR; ++j) s_x[r & 1][warp][lane * kR + j] = v[j];
      }
      __syncthreads();  // one barrier per round: the buffers alternate
      if (warp + 1 < kW && lane >= 32 - kHaloLanes) {
#pragma unroll
        for (int j = 0; j < kR; ++j) v[j] = s_x[r & 1][warp + 1][(lane - (32 - kHaloLanes)) * kR + j];
      }
    }
  }
  if (a.trig == 2) asm volatile("griddepcontrol.launch_dependents;");
#pragma unroll
  for (int j = 0; j < kR; ++j) {
    const long long i = i0 + j;
    if (lane * kR + j < kOwn && c0 + j < a.stride && i <= n_out) a.vout[i] = v[j];
  }
}

// terminal layer: eur max((S-E)*cp, 0) (binom_vanilla_eur.cpp:30), amer payoff(S,E,cp) (binom_vanilla_amer.cpp:29) --
// the same value (multiplication by +-1 is exact)
__global__ void tree_terminal_kernel(double* __restrict__ v, const double* __restrict__ pu, const double* __restrict__ pd,
                                     long long N, double S0, double sgn, double nE) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= N; i += (long long)gridDim.x * blockDim.x) {
    const double S = __dmul_rn(__dmul_rn(S0, pu[i]), pd[N - i]);
    const double c = fma(sgn, S, nE);
    v[i] = c > 0.0 ? c : 0.0;
  }
}

template <int kR, int kSteps>
static int tree_launch_all(Ctx& c, TreeArgs a, long long N, bool amer, double* buf0, double* buf1) {
  constexpr int kStride = 32 * kR - kSteps;
  // PCF_TREE_PDL=0 turns the programmatic dependent launch off (A/B knob); the first launch of the chain follows a
  // plain kernel, for which the attribute is harmless
  // Programmatic dependent launch (profiles/r1_notes.md, r1_tune_tree_pdl.log): releasing the next launch AFTER the
  // layers (mode 2) hides the launch gap -- 11.1 -> 8.0 ms (European) / 20.0 -> 16.8 ms (American) at N = 1e5, 85 -> 62 ms
  // at N = 4e5. Releasing it at kernel start (mode 1) parks the CTAs of the next launches on the SMs while this one
  // computes and skews their placement: 2x SLOWER for N >= 1e5. PCF_TREE_PDL = 0 | 1 | 2 overrides.
  const char* pe = tuning_env("PCF_TREE_PDL");
  const int trig = pe ? atoi(pe) : 2;
  const bool pdl = trig != 0;
  a.trig = trig;
  long long n = N;
  double* in = buf0;
  double* out = buf1;
  while (n > 0) {
    const int steps = (int)std::min<long long>(kSteps, n);
    a.vin = in; a.vout = out; a.n0 = n; a.steps = steps;
    const long long warps = (n - steps + 1 + kStride - 1) / kStride;
    const int grid = (int)((warps + kTreeWarps - 1) / kTreeWarps);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kTreeWarps * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (amer) PCF_CUDA(cudaLaunchKernelEx(&cfg, tree_steps_kernel<kR, kSteps, true>, a));
    else PCF_CUDA(cudaLaunchKernelEx(&cfg, tree_steps_kernel<kR, kSteps, false>, a));
    c.launches++;
    n -= steps;
    std::swap(in, out);
  }
  // root -> c.h_res->vals[0] (pinned)
  PCF_CUDA(cudaMemcpyAsync(c.h_res->vals, in, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// ---- per-launch shape selection for tree_cta_kernel -------------------------------------------------------------
struct CtaCandidate {
  int kR, kW, kKsel, kK, stride;
  void (*eur)(TreeArgs);
  void (*amer)(TreeArgs);
};
#define PCF_CTA_ROW(R, W, K, EUR, AMER) \
  { R, W, K, CtaShape<R, W, K>::kK, CtaShape<R, W, K>::kStride, EUR, AMER }
#define PCF_CTA_BOTH(R, W, K) PCF_CTA_ROW(R, W, K, (tree_cta_kernel<R, W, K, false>), (tree_cta_kernel<R, W, K, true>))
#ifdef PCF_TUNING
#define PCF_CTA_EUR(R, W, K) PCF_CTA_BOTH(R, W, K)
#define PCF_CTA_AMER(R, W, K) PCF_CTA_BOTH(R, W, K)
#else
#define PCF_CTA_EUR(R, W, K) PCF_CTA_ROW(R, W, K, (tree_cta_kernel<R, W, K, false>), nullptr)
#define PCF_CTA_AMER(R, W, K) PCF_CTA_ROW(R, W, K, nullptr, (tree_cta_kernel<R, W, K, true>))
#endif
#define PCF_CTA_SHAPE(R, W) PCF_CTA_BOTH(R, W, 0)
static const CtaCandidate kCtaShapes[] = {
    // the shapes the measured per-width tables below select (kRulesEur / kRulesAmer), in the flavour that selects them
    PCF_CTA_EUR(4, 4, 256), PCF_CTA_EUR(4, 4, 0), PCF_CTA_BOTH(3, 8, 256), PCF_CTA_BOTH(3, 8, 0), PCF_CTA_BOTH(4, 8, 0),
    PCF_CTA_AMER(2, 8, 256), PCF_CTA_AMER(2, 8, 0), PCF_CTA_AMER(2, 12, 0), PCF_CTA_AMER(6, 4, 0), PCF_CTA_AMER(2, 16, 0),
    PCF_CTA_AMER(6, 8, 0),
#ifdef PCF_TUNING
    // every shape tools/tune_tree4.py measured (profiles/r1s_tune_tree_shapes.log)
    PCF_CTA_SHAPE(2, 4),  PCF_CTA_SHAPE(3, 4),  PCF_CTA_SHAPE(8, 4),  PCF_CTA_SHAPE(1, 8),  PCF_CTA_SHAPE(8, 8),
    PCF_CTA_SHAPE(3, 12), PCF_CTA_SHAPE(4, 12), PCF_CTA_SHAPE(6, 12), PCF_CTA_SHAPE(8, 12), PCF_CTA_SHAPE(1, 16),
    PCF_CTA_SHAPE(3, 16), PCF_CTA_SHAPE(4, 16), PCF_CTA_SHAPE(6, 16), PCF_CTA_SHAPE(8, 16), PCF_CTA_SHAPE(1, 20),
    PCF_CTA_SHAPE(8, 20), PCF_CTA_BOTH(2, 12, 256),
#endif
};
#undef PCF_CTA_SHAPE
#undef PCF_CTA_BOTH
#undef PCF_CTA_EUR
#undef PCF_CTA_AMER
#undef PCF_CTA_ROW

// Shape of the launch that starts at an n0-node layer. The table is MEASURED (tools/tune_tree4.py: T(N) of every pinned
// shape on a grid of N; the slope between two grid points is the cost of one layer at that width; the cheapest shape
// per interval is listed, profiles/r1s_tune_tree_shapes.log) on a 148-SM B200 and scaled by the SM count. What it
// encodes: (1) ptxas interleaves the independent node chains of a lane for kR <= 4 but serialises them for kR >= 6
// (one scratch register pair per node), so more than four nodes per lane only pay when the layer is many waves wide;
// (2) the FP64 pipe interleaves the chains of ONE warp better than those of several warps (tools/ubench: 8 warps x 1
// chain 3.0 cycles per instruction, 1 warp x 4 chains 2.2), so the European tree prefers one kR = 4 warp per
// sub-partition; (3) the American node carries three more values per node (S0 u^i, the d-power window, the exercise
// value) and prefers kR = 2 with 8-16 warps until the layer is wider than one wave; (4) 256 layers per launch pay
// only while the layer is narrow (the CTA's decaying edge is a larger share of a small CTA).
struct ShapeRule { long long n_max; int kR, kW, kKsel; };
static const ShapeRule kRulesEur[] = {{30000, 4, 4, 256}, {50000, 4, 4, 0}, {60000, 3, 8, 256}, {80000, 3, 8, 0},
                                      {100000, 4, 4, 0}, {125000, 4, 8, 0}, {200000, 4, 4, 0}, {250000, 4, 8, 0},
                                      {300000, 4, 4, 0}, {-1, 4, 8, 0}};
static const ShapeRule kRulesAmer[] = {{30000, 2, 8, 256}, {50000, 2, 8, 0}, {60000, 3, 8, 256}, {80000, 2, 12, 0},
                                       {90000, 6, 4, 0}, {100000, 2, 16, 0}, {125000, 4, 8, 0}, {150000, 3, 8, 0},
                                       {200000, 6, 4, 0}, {250000, 4, 8, 0}, {700000, 6, 4, 0}, {-1, 6, 8, 0}};

static const CtaCandidate* tree_find_shape(int kR, int kW, int kKsel, bool amer) {
  for (const CtaCandidate& s : kCtaShapes)
    if (s.kR == kR && s.kW == kW && s.kKsel == kKsel && (amer ? s.amer : s.eur) != nullptr) return &s;
  return nullptr;
}

static const CtaCandidate* tree_pick_shape(long long n0, int sms, bool amer, int fixed_r, int fixed_w, int fixed_k) {
  if (fixed_r) return tree_find_shape(fixed_r, fixed_w, fixed_k, amer);
  const double scaled = (double)n0 * 148.0 / (double)std::max(sms, 1);
  for (const ShapeRule* r = amer ? kRulesAmer : kRulesEur;; ++r)
    if (r->n_max < 0 || scaled <= (double)r->n_max) return tree_find_shape(r->kR, r->kW, r->kKsel, amer);
}

// fixed_r/fixed_w != 0 pin one shape for the whole tree (PCF_TREE=1<R><WW>[<K/64>], tests and tuning)
static int tree_launch_cta(Ctx& c, TreeArgs a, long long N, bool amer, double* buf0, double* buf1, int fixed_r, int fixed_w,
                           int fixed_k) {
  const char* pe = tuning_env("PCF_TREE_PDL");
  const int trig = pe ? atoi(pe) : 2;
  a.trig = trig;
  long long n = N;
  double* in = buf0;
  double* out = buf1;
  while (n > 0) {
    const CtaCandidate* ps = tree_pick_shape(n, c.sm_count, amer, fixed_r, fixed_w, fixed_k);
    if (!ps) {
      set_last_error("unknown PCF_TREE");
      return PCF_EINVAL;
    }
    const CtaCandidate& s = *ps;
    const int steps = (int)std::min<long long>(s.kK, n);
    a.vin = in; a.vout = out; a.n0 = n; a.steps = steps; a.stride = s.stride;
    const long long grid = (n - steps + 1 + s.stride - 1) / s.stride;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(s.kW * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = trig != 0 ? 1 : 0;
    PCF_CUDA(cudaLaunchKernelEx(&cfg, amer ? s.amer : s.eur, a));
    c.launches++;
    n -= steps;
    std::swap(in, out);
  }
  PCF_CUDA(cudaMemcpyAsync(c.h_res->vals, in, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  PCF_CUDA(cudaGetLastError());
  return PCF_OK;
}

// pow tables with the reference's own libm call, spread over the host's cores (2(N+1) calls; a serial loop would
// cost as much as the whole device computation at N = 1e5). One set of threads fills both tables.
static void pow_tables(double u, double d, long long N, double* out_u, double* out_d) {
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (N < 4096) nt = 1;
  auto work = [=](unsigned t) {
    for (long long i = t; i <= N; i += nt) {
      out_u[i] = pow(u, (double)(int)i);  // pow(u,i): i is an int in the reference
      out_d[i] = pow(d, (double)(int)i);
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

size_t tree_workspace_bytes(long long N) { return 4 * ((size_t)(N + 1) * sizeof(double) + 256); }

// Enqueues the whole tree on c.stream; root value -> c.h_res->vals[0]. The pow tables are uploaded BEFORE c.ev0 is recorded
// (inputs resident when the device clock starts); the host clock of the call covers them.
int run_binom_tree(Ctx& c, const pcf_params& p, bool american) {
  const long long N = p.N;
  double u, d, pp, q;
  binom_lattice(p.r, p.sigma, p.T, N, u, d, pp, q);
  const double dt = (double)p.T / (double)N;
  const double R = exp(p.r * dt);
  PCF_TRY(ctx_reserve(c, tree_workspace_bytes(N)));
  const size_t slot = (size_t)(N + 1) * sizeof(double) + 256;
  char* ws = (char*)c.workspace;
  double* d_pu = (double*)ws;
  double* d_pd = (double*)(ws + slot);
  double* buf0 = (double*)(ws + 2 * slot);
  double* buf1 = (double*)(ws + 3 * slot);
  std::vector<double> h(2 * (size_t)(N + 1));
  pow_tables(u, d, N, h.data(), h.data() + (N + 1));
  PCF_CUDA(cudaMemcpyAsync(d_pu, h.data(), (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, c.stream));
  PCF_CUDA(cudaMemcpyAsync(d_pd, h.data() + (N + 1), (size_t)(N + 1) * 8, cudaMemcpyHostToDevice, c.stream));
  PCF_CUDA(cudaStreamSynchronize(c.stream));  // `h` dies with this frame
  PCF_CUDA(cudaEventRecord(c.ev0, c.stream));

  TreeArgs a{};
  a.pu = d_pu; a.pd = d_pd;
  a.p = pp; a.q = q; a.R = R; a.z = 1.0 / R;
  a.S0 = p.S0; a.sgn = (double)p.cp; a.nE = -a.sgn * p.E;
  tree_terminal_kernel<<<grid_for(c, N + 1, 256, 8), 256, 0, c.stream>>>(buf0, d_pu, d_pd, N, p.S0, a.sgn, a.nE);
  c.launches++;
  // launch shape (PCF_TUNING builds): unset = CTA-cooperative kernel, shape chosen per launch; PCF_TREE=1<nodes per lane>
  // <warps per CTA, two digits>[<layers per launch / 64>] pins one CTA shape (e.g. 1216, 14044); PCF_TREE=<nodes per lane><layers per launch / 8> selects the warp-trapezoid
  // kernel of the first build
  const char* e = tuning_env("PCF_TREE");
  // tree_cta_kernel's American node relies on continuation values >= 0, i.e. on p, q >= 0; a lattice whose rounded
  // probabilities leave [0, 1] (p is within an ulp of 0 or 1 when sigma -> 0) runs the warp kernel, which keeps both maxima
  const bool cta_ok = !american || (pp >= 0.0 && q >= 0.0);
  if (!e && cta_ok && !(p.flags & PCF_FLAG_TREE_WARP)) return tree_launch_cta(c, a, N, american, buf0, buf1, 0, 0, 0);
  if (!e) return N > 250000 ? tree_launch_all<4, 32>(c, a, N, american, buf0, buf1)
                            : tree_launch_all<4, 64>(c, a, N, american, buf0, buf1);
#ifdef PCF_TUNING
  const int shape = atoi(e);
  if (shape >= 1000 && shape < 20000 && !cta_ok) {
    set_last_error("PCF_TREE: the CTA kernel needs p, q >= 0 for the American tree");
    return PCF_EINVAL;
  }
  if (shape >= 1000 && shape < 2000) return tree_launch_cta(c, a, N, american, buf0, buf1, (shape / 100) % 10, shape % 100, 0);
  if (shape >= 10000 && shape < 20000)
    return tree_launch_cta(c, a, N, american, buf0, buf1, (shape / 1000) % 10, (shape / 10) % 100, 64 * (shape % 10));
  switch (shape) {
    case 22: return tree_launch_all<2, 16>(c, a, N, american, buf0, buf1);
    case 44: return tree_launch_all<4, 32>(c, a, N, american, buf0, buf1);
    case 48: return tree_launch_all<4, 64>(c, a, N, american, buf0, buf1);
    case 68: return tree_launch_all<6, 64>(c, a, N, american, buf0, buf1);
    case 88: return tree_launch_all<8, 64>(c, a, N, american, buf0, buf1);
    case 84: return tree_launch_all<8, 32>(c, a, N, american, buf0, buf1);
    default:
      set_last_error("unknown PCF_TREE");
      return PCF_EINVAL;
  }
#else
  return PCF_EINVAL;  // unreachable: tuning_env() is null in this build
#endif
}

}  // namespace pcf
