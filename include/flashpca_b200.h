/*
 * flashpca_b200.h -- C ABI of the B200-native FlashPCA2 hot path.
 *
 * Drop-in boundary: the matrix-free operator family that upstream flashpca
 * hands to Spectra (svdwide.h:27-30,77-106) over Data::read_snp_block
 * (data.cpp:215-335), plus a whole-solve entry that keeps the Lanczos basis on
 * the device.  Plain pointers and sizes only; no C++/torch types.  Every entry
 * point cites the upstream interface it replaces (paths relative to the
 * flashpca source tree).
 *
 * Conventions
 *   - all matrices are column-major double, as Eigen::MatrixXd upstream;
 *   - functions return 0 on success, non-zero on error; the message is
 *     available from fpb_last_error() (upstream throws std::runtime_error,
 *     data.cpp:160,188,287; the C++ wrappers in flashpca_b200/host re-throw);
 *   - a handle is not thread-safe and not re-entrant, like upstream's
 *     SVDWideOnline (it mutates nops/trace/dat.X, svdwide.cpp:21-68);
 *   - "_dev" variants take device pointers valid on the handle's device and
 *     enqueue on the handle's stream (fpb_stream()); the plain variants take
 *     host pointers and include the host<->device copies.
 *   - there is no CPU fallback: every call fails if no CUDA device is usable.
 */
#ifndef FLASHPCA_B200_H
#define FLASHPCA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPB_STANDARDISE_BINOM 2  /* util.h:36 STANDARDISE_BINOM  */
#define FPB_STANDARDISE_BINOM2 3 /* util.h:37 STANDARDISE_BINOM2 */

#define FPB_DIVISOR_NONE 0 /* randompca.h:50 */
#define FPB_DIVISOR_N1 1   /* randompca.h:51 */
#define FPB_DIVISOR_P 2    /* randompca.h:52 */

typedef struct fpb_handle fpb_handle;

/* Library/ABI version and the message of the last failed call on this thread
 * (h may be NULL for failures of fpb_create*). */
int fpb_abi_version(void);
const char *fpb_last_error(const fpb_handle *h);

/* ---- construction: replaces Data::get_size + Data::prepare (data.cpp:150-206)
 * and the SVDWideOnline constructor (svdwide.h:51-73). ------------------------
 *
 * fpb_create: stage `nsnps` SNP columns of packed PLINK genotypes (SNP-major,
 * np = ceil(N/4) bytes per SNP, the 3-byte bed header already skipped) from
 * host memory into HBM, once.  `bed_payload` is only read during the call.
 * stand_method: FPB_STANDARDISE_BINOM | FPB_STANDARDISE_BINOM2
 * (data.cpp:279-288; anything else fails with upstream's message).
 * preloaded_meansd: NULL, or nsnps x 2 column-major (mean, sd) used instead of
 * the computed statistics (Data::use_preloaded_maf, data.cpp:293-297).
 * device: CUDA ordinal. */
int fpb_create(fpb_handle **out, const unsigned char *bed_payload, uint64_t n_individuals,
               uint64_t nsnps, int stand_method, const double *preloaded_meansd, int device);

/* fpb_create_from_file: same, reading SNP columns [snp_begin, snp_begin +
 * snp_count) of a .bed file through pinned staging buffers (the contiguous
 * byte range 3 + np*snp_begin ..., data.cpp:218).  snp_count == 0 means "to
 * the end of the file"; the file's SNP count is derived from its size exactly
 * as data.cpp:166-170 does (no magic-byte validation). */
int fpb_create_from_file(fpb_handle **out, const char *bed_path, uint64_t n_individuals,
                         uint64_t snp_begin, uint64_t snp_count, int stand_method,
                         const double *preloaded_meansd, int device);

/* fpb_create_streaming: out-of-HBM mode for beds larger than one GPU's memory
 * (replaces the disk loop of Data::read_snp_block, data.cpp:215-335).
 * The recoded 2-bit matrix is kept in pinned host memory in slabs of
 * snps_per_slab SNP columns; every operator call streams the slabs through two
 * device buffers (the copy of slab b+1 overlaps the kernels of slab b) and
 * accumulates y = sum_b X_b X_b' x in slab order, the reference's block loop
 * (svdwide.cpp:48-59).  Statistics and missing-genotype lists of every slab
 * stay on the device.  All operator entry points, fpb_pca and the multi-GPU
 * calls accept the handle; fpb_get_bed does not. */
int fpb_create_streaming(fpb_handle **out, const char *bed_path, uint64_t n_individuals,
                         uint64_t snp_begin, uint64_t snp_count, uint64_t snps_per_slab,
                         int stand_method, const double *preloaded_meansd, int device);

/* fpb_create_synthetic: generate the packed genotypes directly in HBM
 * (bench-only input path; counter-based integer hash, reproduced bit for bit
 * on the host by flashpca_b200/synth.py).  pop_of_individual: N bytes;
 * thresholds: npop x nsnps uint32, row-major by population (allele-frequency
 * thresholds, p * 2^32); SNP ids snp_offset.. are hashed so shards of one
 * matrix agree with the single-GPU matrix. */
int fpb_create_synthetic(fpb_handle **out, uint64_t n_individuals, uint64_t nsnps,
                         uint64_t snp_offset, const unsigned char *pop_of_individual,
                         const uint32_t *thresholds, uint32_t npop, uint32_t missing_threshold,
                         uint64_t seed, int stand_method, int device);

/* fpb_create_dense: the in-memory matrix path -- RandomPCA::pca_fast(MatrixXd&, ...)
 * (randompca.cpp:121-166) with SVDWide (svdwide.h:9-30, svdwide.cpp:4-12).  x is an
 * N x P column-major matrix of dosages, NaN = missing; it is copied to HBM and
 * standardised there exactly as standardise() does (util.cpp:24-192):
 * stand_method 0 none, 1 sd, 2 binom, 3 binom2, 4 center (util.h:34-38).
 * The handle then serves the same operator family; fpb_get_meansd returns the
 * (mean, sd) table standardise() returns, fpb_get_trace the sum of squares
 * (randompca.cpp:154), fpb_get_dense the standardised matrix. */
int fpb_create_dense(fpb_handle **out, const double *x, uint64_t n_individuals, uint64_t nsnps,
                     int stand_method, int device);
int fpb_get_dense(fpb_handle *h, double *out_x);

void fpb_destroy(fpb_handle *h);

/* ---- shape, statistics -------------------------------------------------- */
uint64_t fpb_rows(const fpb_handle *h);  /* SVDWideOnline::rows(), svdwide.h:77: N   */
uint64_t fpb_cols(const fpb_handle *h);  /* SVDWideOnline::cols(), svdwide.h:78: N   */
uint64_t fpb_nsnps(const fpb_handle *h); /* Data::nsnps of this handle (shard-local) */
void *fpb_stream(const fpb_handle *h);   /* cudaStream_t the handle launches on      */

/* Data::X_meansd (data.cpp:290-291): nsnps x 2 column-major (mean, sd). */
int fpb_get_meansd(fpb_handle *h, double *out_meansd);
/* SVDWideOnline::trace (svdwide.h:37, svdwide.cpp:44-45,60-61): sum X_ij^2 of
 * the standardised matrix, un-divided; shard-local when sharded. */
int fpb_get_trace(fpb_handle *h, double *out_trace);
/* Copy the staged packed genotypes back (nsnps * ceil(N/4) bytes, upstream
 * layout, pad bits as staged) -- used by tests to feed the oracle. */
int fpb_get_bed(fpb_handle *h, unsigned char *out_payload);

/* ---- the operator family ------------------------------------------------- */
/* SVDWideOnline::perform_op (svdwide.cpp:21-68): y = X X' x, un-normalised.
 * x: N doubles (read only), y: N doubles (fully overwritten), no aliasing. */
int fpb_perform_op(fpb_handle *h, const double *x_in, double *y_out);
/* perform_op_mat / perform_op_multi (svdwide.cpp:71-118, 229-275):
 * Y (N x k) = X X' M (N x k). */
int fpb_perform_op_multi(fpb_handle *h, const double *m_in, uint32_t k, double *y_out);
/* crossprod / crossprod2 (svdwide.cpp:122-153, 157-188): Y (nsnps x k) = X' M (N x k). */
int fpb_crossprod(fpb_handle *h, const double *x_in, double *y_out);
int fpb_crossprod_multi(fpb_handle *h, const double *m_in, uint32_t k, double *y_out);
/* prod / prod3 (svdwide.cpp:193-226, 312-343): Y (N x k) = X V (nsnps x k). */
int fpb_prod(fpb_handle *h, const double *v_in, double *y_out);
int fpb_prod_multi(fpb_handle *h, const double *v_in, uint32_t k, double *y_out);

/* Device-pointer variants (same maths, no host copies, asynchronous on
 * fpb_stream(h); when a communicator is attached perform_op/prod all-reduce
 * their N x k result across ranks on the same stream). */
int fpb_perform_op_dev(fpb_handle *h, const double *d_x, double *d_y);
int fpb_perform_op_multi_dev(fpb_handle *h, const double *d_m, uint32_t k, double *d_y);
int fpb_crossprod_multi_dev(fpb_handle *h, const double *d_m, uint32_t k, double *d_y);
int fpb_prod_multi_dev(fpb_handle *h, const double *d_v, uint32_t k, double *d_y);
int fpb_sync(fpb_handle *h);

/* ---- SNP-sharded multi-GPU: one handle (one process) per GPU ---------------
 * Each rank stages a contiguous SNP range; X X' x = sum_g X_g X_g' x
 * (the block sum of svdwide.cpp:48-59, distributed).  The communicator is
 * NCCL; the 128-byte unique id is created by rank 0 (fpb_comm_unique_id) and
 * distributed by the host program (bench.py uses torch.distributed). */
int fpb_comm_unique_id(unsigned char id_out[128]);
int fpb_comm_init(fpb_handle *h, const unsigned char id[128], int nranks, int rank);
/* The shard sum itself: fpb_comm_init also maps one exchange region per rank into every other
 * rank (CUDA IPC, handles exchanged over the new communicator); when that works the sum runs as
 * one kernel over NVLink peer memory, fused with the last step of the op (fixed summation order,
 * bit-identical on every rank), else as ncclAllReduce.  fpb_comm_kind: 0 = single shard,
 * 1 = ncclAllReduce, 2 = peer-memory kernel.  FPB_PEER=0 keeps NCCL.
 * The host-pointer fpb_perform_op / fpb_perform_op_multi expect the same input on every rank (as the
 * replicated Lanczos drivers provide it): with the peer-memory path each rank uploads only its
 * slice of it and the slices are exchanged over NVLink (FPB_SLICE_UPLOAD=0: every rank uploads all).
 * fpb_comm_link_local links n handles of ONE process (shards on one GPU, or on GPUs with peer
 * access) as ranks 0..n-1 without NCCL; calls on the n handles must be issued concurrently
 * (one host thread per handle): each rank's kernel waits for the others. */
int fpb_comm_kind(const fpb_handle *h);
int fpb_comm_link_local(fpb_handle **handles, int n);

/* ---- whole solve: RandomPCA::pca_fast(Data&, ...) (randompca.cpp:168-218) ----
 * Implicitly restarted Lanczos with the schedule of Spectra 0.8.1
 * SymEigsSolver<double, LARGEST_ALGE, Op>(op, nev, ncv) -- init() + compute
 * (maxiter, tol) -- with the Krylov basis resident in HBM.
 * Outputs (any may be NULL): evals[nev] = Ritz values of X X' (descending,
 * un-divided: upstream divides afterwards, randompca.cpp:190), evecs N x nev
 * column-major unit vectors.  *nconv_out = converged count (success iff
 * == nev, as randompca.cpp:186,212-217), *nops_out = perform_op calls,
 * *niter_out = restarts + 1.  Returns 0 when the iteration ran (check
 * *nconv_out for convergence), non-zero on a CUDA / argument error. */
int fpb_pca(fpb_handle *h, uint32_t nev, uint32_t ncv, uint32_t maxiter, double tol,
            double *evals_out, double *evecs_out, uint32_t *nconv_out, uint32_t *nops_out,
            uint32_t *niter_out);

/* Block Krylov variant of the whole solve -- an EXTENSION, not upstream's algorithm: block Lanczos
 * (block = columns per pass over the packed matrix, <= 8; 0 = 8) with full re-orthogonalisation and
 * Rayleigh-Ritz on the accumulated space, no restart, at most max_passes passes (0 = 40).  On B200
 * one 8-column pass of the operator costs about two single-vector ops (tcgen05 block kernels), so
 * this converges k = 20 in a fraction of the time of Spectra's single-vector schedule; the
 * convergence test and tolerance are Spectra's, the trajectory is not.  Same outputs as fpb_pca
 * (evals descending, un-divided); *npasses_out = operator passes of `block` columns. */
int fpb_pca_block(fpb_handle *h, uint32_t nev, uint32_t block, uint32_t max_passes, double tol,
                  double *evals_out, double *evecs_out, uint32_t *nconv_out, uint32_t *npasses_out);

/* Self-check of the last fpb_pca / fpb_pca_block result without leaving the device: RandomPCA::check
 * (randompca.cpp:663-703) applied to the solver's own eigenpairs,
 *   err_out[j] = || X X' u_j / div - u_j d_j ||^2,  d_j = lambda_j / div,  j < nev
 * (mse = sum_j err_j / (N nev), "< 1e-8" per README.md:207).  One block call
 * fpb_perform_op_multi_dev on the N x nev eigenvectors; collective when a communicator is
 * attached (every rank must call it). */
int fpb_pca_residual(fpb_handle *h, double div, double *err_out, uint32_t nev);

/* Per-op device timings of the last fpb_pca call, milliseconds (CUDA events on
 * the handle's stream); returns the number written (<= cap). */
uint32_t fpb_pca_op_times(const fpb_handle *h, float *ms_out, uint32_t cap);
/* Host wall-clock of the last fpb_pca call, seconds: [0] Lanczos iteration
 * (setup included), [1] eigenvector assembly V*ritz, [2] download of the
 * eigenvectors to the caller's buffer, [3] total. */
void fpb_pca_phase_times(const fpb_handle *h, double out_seconds[4]);

/* ---- measurement helpers (bench.py) --------------------------------------
 * Enqueue `reps` back-to-back device perform_op calls and return the mean
 * milliseconds per call (CUDA events on fpb_stream(h)).  ms_kernels_out (may be
 * NULL, else 4 floats) receives, from one further op: [0] the X'x half, [1] the
 * X t half (each with its small kernels), [2], [3] the contraction kernel of
 * each half alone (events recorded immediately around its launch; 0 on the
 * generic FP64 path).  When the fused single-pass kernel is in use
 * (fpb_path_info & FPB_PATH_FUSED): [0] the whole op, [2] the fused kernel
 * alone, [1] = [3] = 0. */
int fpb_time_perform_op(fpb_handle *h, const double *d_x, double *d_y, uint32_t reps,
                        float *ms_per_op_out, float *ms_kernels_out);

/* Same op sequence with a CUDA event after every op: ms_each_out[r] (reps floats) = device time
 * of op r.  bench.py reports mean (= total / reps) and median from these. */
int fpb_time_perform_op_steps(fpb_handle *h, const double *d_x, double *d_y, uint32_t reps,
                              float *ms_each_out);

/* Which compute path the handle selected at staging (bit mask). */
#define FPB_PATH_DENSE 1u       /* in-memory matrix (fpb_create_dense) */
#define FPB_PATH_TENSOR 2u      /* exact int8-sliced tensor-core contraction */
#define FPB_PATH_TMA 4u         /* TMA + mbarrier pipelines */
#define FPB_PATH_SINGLE_COPY 8u /* both halves read the one SNP-major copy */
#define FPB_PATH_FUSED 16u      /* perform_op reads HBM once (fused two-phase kernel) */
#define FPB_PATH_STREAMING 32u  /* genotypes in pinned host memory, streamed per op */
unsigned fpb_path_info(const fpb_handle *h);

/* Free and total memory of a device, bytes (hosts use it to choose between
 * fpb_create_from_file and fpb_create_streaming). */
int fpb_device_memory(int device, uint64_t *free_bytes, uint64_t *total_bytes);

/* Debug: with FPB_FUSED_DEBUG=1 in the environment at staging, the fused kernel
 * records globaltimer stamps of its cross-CTA protocol for the first 256 slabs
 * of CTA 0 and of the last CTA: out[2][256][8] (count >= 4096). */
int fpb_fused_debug(fpb_handle *h, unsigned long long *out, uint64_t count);
/* Number of kernels this library has launched on the handle so far. */
uint64_t fpb_launch_count(const fpb_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* FLASHPCA_B200_H */
