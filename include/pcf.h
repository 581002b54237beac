/* include/pcf.h -- C ABI of libpcf.so, the B200 (sm_100a) implementation of the parcompfin
 * Monte Carlo / binomial pricing hot path.
 *
 * The reference (moledoc/parcompfin) has no FFI of its own: every method is a free function in
 * the same translation unit as its main(). Each entry point below replaces exactly one of those
 * functions; the five C++ front ends in parcompfin_b200/host/ keep the reference's argv layout
 * and CSV row and call nothing else. Plain C types only, caller owns every struct, the library
 * owns all device memory between pcf_init*() and pcf_shutdown(). Nothing throws across this
 * boundary: every function returns a pcf_status (0 = OK). There is no CPU fallback: without a
 * CUDA device every compute call returns PCF_ECUDA.
 *
 * Thread-compatible, not thread-safe: one context set per process, calls are serialised by the
 * caller (the reference programs make one call per process).
 */
#ifndef PCF_H_
#define PCF_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PCF_ABI_VERSION 1
#if defined(__GNUC__)
#define PCF_API __attribute__((visibility("default")))
#else
#define PCF_API
#endif

typedef enum pcf_status {
  PCF_OK = 0,
  PCF_EINVAL_PAYOFF = 1, /* cp not +1/-1: reference throws "Unknown payoff function", src/mc_eur.cpp:42 */
  PCF_EODD_N = 2,        /* mc_amer with odd N: reference include/common.h:180 throws                     */
  PCF_ESINGULAR = 3,     /* regression determinant <= 0: reference include/common.h:115-117 throws        */
  PCF_EINVAL = 4,        /* other invalid argument (N <= 0, M <= 0, assets out of range, short replay)    */
  PCF_ENOTPD = 5,        /* covariance matrix has a genuinely negative eigenvalue. A positive SEMI-definite
                            matrix is NOT an error: like the reference (include/mvn.h:68-76) every basket
                            entry point falls back from Cholesky to eigenvectors * sqrt(eigenvalues); where
                            the reference's cwiseSqrt then meets a negative eigenvalue it prices NaN -- that
                            case is reported with this code instead                                       */
  PCF_ECUDA = 10,        /* CUDA runtime failure or no device; pcf_last_error() has the text              */
  PCF_ENCCL = 11,        /* NCCL failure or libnccl.so.2 not loadable                                     */
  PCF_ENOINIT = 12,      /* pcf_init()/pcf_init_rank() not called                                         */
  PCF_ENOMEM = 13        /* device allocation failed (mc_amer path store does not fit)                    */
} pcf_status;

/* flags for pcf_params.flags */
#define PCF_FLAG_BINOM_WINDOW 0x1u /* binom_embar: skip term pairs whose weight underflows (reported
                                      separately from the full term sum; default off)               */

#define PCF_FLAG_AMER_LSM 0x2u     /* mc_amer: textbook Longstaff-Schwartz decision instead of the reference's
                                      (SURVEY F1 / 8f.3): exercise when the TRUE payoff exceeds the fitted
                                      continuation value and book that payoff. Off by default: parity with the
                                      reference means reproducing its rule.                                     */

#define PCF_FLAG_BINOM_NOSCREEN 0x4u /* binom_embar: evaluate every term pair with the full-accuracy saddle-point
                                      routine. By default a pair is first screened with a two-logarithm evaluation of
                                      its log-weights (abs. error < 1e-4) and settled as 0 when both are below the
                                      underflow rule's threshold with margin; the sum is bit-identical either way
                                      (tests assert it), the flag exists so both rates can be reported.            */

#define PCF_FLAG_BASKET_GENERAL 0x8u /* mc_eur_multi: price the reference's equicorrelation basket with the general
                                      triangular-product kernel instead of the constant-column fast path (same chain of
                                      FMAs, bit-identical sums; tests assert it)                                       */

#define PCF_FLAG_TREE_WARP 0x10u   /* binom_vanilla_*: warp-trapezoid tiling (every warp independent) instead of the
                                      CTA-cooperative one picked per launch; same per-node operations, bit-identical
                                      root (tests assert it). Lattices with p or q outside [0,1] always take it.       */

/* Normal-stream ids (word 3 of the Philox counter), one per method. */
#define PCF_STREAM_EUR 0u
#define PCF_STREAM_ASIA 1u
#define PCF_STREAM_BASKET 2u
#define PCF_STREAM_AMER 3u

/* The reference's parameter set (positional argv of every program, e.g. src/mc_asia.cpp:44-51). */
typedef struct pcf_params {
  double S0, E, r, sigma, T;
  int cp;           /* +1 call, -1 put                                                              */
  long long N;      /* paths (MC) or lattice steps (binom_embar)                                    */
  int M;            /* monitoring / exercise dates (mc_asia, mc_amer); ignored elsewhere            */
  int assets;       /* basket size d (mc_eur_multi), 1..PCF_MAX_ASSETS                              */
  double rho;       /* equicorrelation (mc_eur_multi)                                               */
  unsigned long long seed; /* Philox key (native mode)                                              */
  /* Replay mode: host pointer to the post-scaling normal variates in the reference's draw order
   * (mc_eur w[n]; mc_asia dB[n*M+m]; mc_amer w[p*M+(m-1)], p < N/2; mc_eur_multi Z[n*d+a]).
   * NULL = native mode (Philox4x32-10 keyed by global index). */
  const double* replay;
  long long replay_len;
  unsigned int flags;
} pcf_params;

#define PCF_MAX_ASSETS 32

typedef struct pcf_result {
  double price;          /* what the reference function returns                                      */
  double sum, sumsq;     /* global sum and sum of squares of the undiscounted payoffs (MC) / the
                            undiscounted term sum and 0 (binom)                                      */
  double std_error;      /* standard error of `price` (0 for binom)                                  */
  long long n;           /* samples (paths) or terms accumulated, all GPUs                           */
  long long units;       /* path-steps (N*max(M,1)) or terms, all GPUs: the throughput unit          */
  double seconds_kernel; /* CUDA-event time on the launching stream, first kernel to last collective,
                            this process's slowest GPU                                               */
  double seconds_total;  /* host wall time of the call (parameter upload, launches, result read-back) */
  int launches;          /* kernels this process launched for the call                               */
  int gpus;              /* GPUs that shared the work                                                */
  int status;            /* same value the function returned                                         */
} pcf_result;

/* --- lifetime --------------------------------------------------------------------------------- */
/* Single process driving the first `gpus` visible devices (0 = all): the front ends' trailing
 * [gpus] argument, mirroring the [threads] argument of the reference's _omp programs
 * (src/mc_eur_omp.cpp:41). gpus > visible devices is PCF_EINVAL. With gpus > 1 the GPUs exchange their partial
 * moments through NVLink peer-memory mailboxes (csrc/xchg.cuh); an NCCL communicator is built only when peer access
 * is unavailable or PCF_NO_PEER is set, or later by pcf_peer_enable(0). */
PCF_API int pcf_init(int gpus);
/* One process per GPU (torchrun / mpirun style). `nccl_id` is the 128-byte NCCL unique id created
 * by rank 0 with pcf_nccl_unique_id() and distributed by the caller (any transport); may be NULL
 * when world == 1. Replaces MPI_Init + MPI_Comm_rank/size of the reference's _mpi programs
 * (src/mc_eur_mpi.cpp:58-66). */
PCF_API int pcf_init_rank(int rank, int world, int device, const unsigned char* nccl_id);
PCF_API int pcf_nccl_unique_id(unsigned char id[128]);
/* NVLink peer-memory exchange between processes (one per GPU, same node): every rank exports the CUDA IPC handle
 * of its 1 KB mailbox after pcf_init_rank(), the caller all-gathers the world x 64 bytes by any transport and
 * hands them to pcf_ipc_import(). From then on partial moments travel as P2P stores issued by the reducing
 * kernel itself (csrc/xchg.cuh) instead of ncclAllReduce. pcf_init(G) sets this up by itself.
 * pcf_peer_enable(0) switches back to NCCL (requires a communicator); pcf_peer_active() reports the mode. */
PCF_API int pcf_ipc_export(unsigned char handle[64]);
PCF_API int pcf_ipc_import(const unsigned char* handles, int world);
PCF_API int pcf_peer_enable(int on);
PCF_API int pcf_peer_active(void);
PCF_API int pcf_shutdown(void);
PCF_API int pcf_world_size(void);

/* --- the hot path ----------------------------------------------------------------------------- */
/* replaces mc_eur(),   reference src/mc_eur.cpp:5-27                                             */
PCF_API int pcf_mc_eur(const pcf_params* p, pcf_result* out);
/* replaces mc_eur() + mvnorm(), reference src/mc_eur_multi.cpp:6-35, include/mvn.h:42-82 -- both branches of
 * mvn.h:68-76: Cholesky factor, or (rho = 1, rho = -1/(d-1): LLT meets a zero pivot) the eigen-decomposition.
 * Eigenvalues within 1e-10 * lambda_max below zero are rounding noise of a semi-definite matrix and count as 0
 * (the reference's sqrt of such a value is NaN by accident of rounding). */
PCF_API int pcf_mc_eur_multi(const pcf_params* p, pcf_result* out);
/* General basket: the same pricing loop (src/mc_eur_multi.cpp:23-34) with what the reference hard-wires
 * made explicit. Every pointer is a HOST array and may be NULL, which selects the reference's value:
 *   S0[d], sigma[d]  per-asset spot / volatility           (NULL: p->S0 / p->sigma for every asset, :30)
 *   weight[d]        basket weights                        (NULL: 1/d, :30)
 *   cov[d*d]         row-major symmetric covariance of the driving normals (NULL: equicorrelation p->rho,
 *                    include/mvn.h:55-60); factored as mvn.h:63-76 does: Cholesky, else eigenvectors * sqrt(eigenvalues)
 *                    when only positive SEMI-definite; PCF_ENOTPD when it has a negative eigenvalue
 *   transform[d*d]   row-major A with Bt = A Z, overrides cov (what mvn.h calls normTransform); replay parity tests
 *                    hand the oracle and the GPU the same A
 * d = p->assets. Replay layout as pcf_mc_eur_multi: Z[n*d + a]. */
typedef struct pcf_basket {
  const double* S0;
  const double* sigma;
  const double* weight;
  const double* cov;
  const double* transform;
} pcf_basket;
PCF_API int pcf_mc_basket(const pcf_params* p, const pcf_basket* b, pcf_result* out);
/* The factor pcf_mc_basket builds from `cov` (host computation): A A^T = cov. *used_eigen = 1 when the Cholesky
 * factorisation failed and the eigen-decomposition of mvn.h:72-76 was used (A is then a full matrix). */
PCF_API int pcf_normal_transform(int d, const double* cov, double* A, int* used_eigen);
/* replaces mc_asia(),  reference src/mc_asia.cpp:5-40                                            */
PCF_API int pcf_mc_asia(const pcf_params* p, pcf_result* out);
/* replaces mc_amer() + pathsfinder() + inverse()/mat_vec_mul(), reference src/mc_amer.cpp:5-114,
 * include/common.h:98-141,168-208                                                                */
PCF_API int pcf_mc_amer(const pcf_params* p, pcf_result* out);
/* replaces binom() + comb(), reference src/binom_embar.cpp:5-50, include/common.h:63-72          */
PCF_API int pcf_binom_embar(const pcf_params* p, pcf_result* out);

/* Backward-induction trees: the programs the reference's run-scripts call to produce `comparison`
 * (runscript_mc_eur.sh:23, runscript_mc_amer.sh:24). N <= 1e7 (the work is N^2/2 node updates). One GPU: the
 * layers are a serial chain, the path does not shard. `units` = node updates.
 * replaces binom(), reference src/binom_vanilla_eur.cpp:5-41                                      */
PCF_API int pcf_binom_vanilla_eur(const pcf_params* p, pcf_result* out);
/* replaces binom(), reference src/binom_vanilla_amer.cpp:5-42                                     */
PCF_API int pcf_binom_vanilla_amer(const pcf_params* p, pcf_result* out);

/* --- diagnostics / test support ----------------------------------------------------------------- */
/* The normal variates the native-mode kernels consume ("normal stream v1"):
 *   Philox4x32-10, key = seed, counter = (index lo, index hi, t/2, stream); X1 = x1:x0, X2 = x3:x2;
 *   u1 = 1 - (X1>>12)*2^-52, u2 = ((X2>>6)+1/2)*2^-58; z(index, t) = sqrt(-2 ln u1) *
 *   (t even ? cos : sin)(2 pi u2).  Writes out[i*T + t] = scale*z(index0+i, t) to HOST memory, so a
 *   native run can be replayed through the CPU oracle. Runs on the first context's GPU. */
PCF_API int pcf_normal_stream(unsigned long long seed, unsigned int stream, unsigned long long index0,
                      long long count, int T, double scale, double* out_host);
/* Raw Philox4x32-10 block (known-answer tests). Runs on the GPU. */
PCF_API int pcf_philox4x32_10(const unsigned int ctr[4], const unsigned int key[2], unsigned int out[4]);
/* Lower Cholesky factor of the d x d equicorrelation matrix, row-major, as uploaded to the basket
 * kernel (host computation; reference include/mvn.h:53-70). */
PCF_API int pcf_chol_equicorr(int d, double rho, double* L);
/* FP64 roofline denominator: sustained DFMA warp-instruction issue measured by a register-resident
 * FMA-chain microbenchmark on the first context's GPU. Returns thread-level DFMA per second
 * (multiply by 2 for flop/s). */
PCF_API int pcf_fp64_peak(double seconds_target, double* dfma_per_sec);
/* Device copy bandwidth (read+write bytes per second) of a `bytes`-sized buffer, same GPU. */
PCF_API int pcf_hbm_peak(long long bytes, double* bytes_per_sec);
PCF_API int pcf_device_info(char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                    long long* mem_bytes);

PCF_API const char* pcf_strerror(int status);
PCF_API const char* pcf_last_error(void); /* text of the last CUDA/NCCL failure in this process */

#ifdef __cplusplus
}
#endif
#endif /* PCF_H_ */
