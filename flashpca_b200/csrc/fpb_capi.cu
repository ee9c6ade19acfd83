// fpb_capi.cu -- C ABI (include/flashpca_b200.h) over the sm_100a kernels.
//
// Host-side state of one staged genotype matrix (or SNP shard of one):
//   d_gs     nsnps x pitch_s bytes  SNP-major 2-bit dosage codes (fpb_kernels.cuh)
//   d_gi     N x pitch_i bytes      individual-major copy (two-copy kernel variants only)
//   d_scale  nsnps x (mean, 1/sd)   d_lut  nsnps x double4 (generic path table)
//   d_meansd nsnps x 2              Data::X_meansd (data.cpp:290-291)
//   CSR lists of the missing genotypes by SNP and by individual (tensor path)
//   per-op scratch: digit slices, split partials, a/b/corr vectors
// Two compute paths share the staged data: the tensor path (fpb_imma.cuh, used
// when the missing rate is <= 3 %) and the generic FP64 path (fpb_kernels.cuh).
// FPB_PATH=generic|imma in the environment forces one (tests exercise both).
// No CPU fallback exists: if CUDA is unusable every entry point fails.
// Host-side pieces with a life of their own are textual includes of this translation unit:
//   fpb_host_fused_setup.inl / fpb_host_fused.inl   fused single-pass kernel (setup, launch, errors)
//   fpb_host_pairs.inl                              two vectors per pass (block variants)
//   fpb_host_streaming.inl / _create.inl            out-of-HBM streaming mode
#include "../../include/flashpca_b200.h"

#include <cuda.h>  // CUtensorMap types only; cuTensorMapEncodeTiled is resolved at run time
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "fpb_dense.cuh"
#include "fpb_fused.cuh"
#include "fpb_imma.cuh"
#include "fpb_irlm.cuh"
#include "fpb_block.cuh"
#include "fpb_kernels.cuh"
#include "fpb_umma.cuh"
#include "fpb_peer.cuh"

namespace {

thread_local std::string g_err;
// fpb_create_streaming stages every slab straight into one of its two slab buffers: when set,
// alloc_common uses this allocation for d_gs instead of a fresh one (the caller detaches it again)
thread_local uint8_t* g_borrow_gs = nullptr;

struct Tiling {
  int W = 1;
  uint32_t block = 256, chunks = 1, splits = 1, snps_per_split = 1;
};

// minimal NCCL surface, resolved with dlopen so the library has no link-time
// dependency on NCCL (single-GPU users never load it).
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;  // optional
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot load NCCL: ") + dlerror();
      return false;
    }
    GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(
        lib, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(
        lib, "ncclAllGather");
    CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
      err = "NCCL library lacks required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

}  // namespace

struct fpb_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;            // CSR gathers, overlapped with the contraction
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint64_t n = 0, nsnps = 0, np = 0;
  uint64_t pitch_s = 0, pitch_i = 0;
  int stand_method = 0;
  int sm_count = 148;
  uint8_t* d_gs = nullptr;
  uint8_t* d_gi = nullptr;
  double4* d_lut = nullptr;
  double2* d_scale = nullptr;
  double* d_meansd = nullptr;
  // generic path scratch
  double* d_t = nullptr;
  double4* d_coef = nullptr;
  double* d_c0 = nullptr;
  double* d_gpart = nullptr;  // generic path: fixed-order partials (allocated on first use)
  Tiling tl;
  // in-memory matrix path (fpb_create_dense): N x P doubles, already standardised
  bool dense = false;
  double* d_X = nullptr;
  uint32_t dense_splits = 1, dense_cps = 1;
  // tensor path
  bool use_imma = false;
  uint64_t nmissing = 0;
  // column-blocked missing-genotype lists (fpb_imma.cuh): by SNP and by individual
  uint64_t* d_rowptr_s = nullptr;  // staging only
  uint64_t *d_seg_s = nullptr, *d_seg_i = nullptr;
  uint16_t *d_col16_s = nullptr, *d_col16_i = nullptr;
  uint32_t gtiles_s = 0, gtiles_i = 0;
  uint32_t gather_sms = 0;  // > 0: k_sell_gather_p on that many dedicated SMs
  // first SNP row of the window the first half leaves in L2 for the second half (evict_last)
  uint32_t l2_keep_row = 0xFFFFFFFFu;
  uint4* d_slices = nullptr;
  double* d_part = nullptr;
  double *d_a = nullptr, *d_corr = nullptr;
  double *d_pmax = nullptr, *d_psum = nullptr;  // (max|v|, sum) block partials of the next vector
  uint32_t nparts = 0;
  double *d_mx = nullptr, *d_mc = nullptr;  // per-SNP / per-individual missing-genotype sums
  fpb::VecScale* d_sc = nullptr;  // [0] = x, [1] = (step of a, sum of b)
  uint32_t nchunks_s = 0, nchunks_i = 0, splits_s = 1, splits_i = 1, cps_s = 1, cps_i = 1;
  // TMA kernel variant: tensor maps over gs / gi, 128-byte stages
  bool use_tma = false;
  fpb::TmaDesc tm_s, tm_i;
  uint32_t nstages_s = 0, nstages_i = 0, tsplits_s = 1, tsplits_i = 1, sps_s = 1, sps_i = 1;
  // single-copy mode: the second half also reads gs (k_imma_gemv_tma_t); gi is not kept
  bool single_copy = false;
  uint32_t ttiles = 0, ttsplits = 1, ttps = 1;
  uint64_t part_stride = 0;
  // persistent forms of the two TMA kernels: used per half when the one-item-per-CTA grid would have
  // fewer than ~10 waves (pipeline fill and the partial last wave then dominate); FPB_PERSIST=0|1 forces
  bool persist_s = false, persist_t = false;
  // wide-stripe second half (k_imma_gemv_tma_tw): 256-byte stripes, 128-row stages
  bool wide_t = false;
  uint32_t wtiles = 0, wsplits = 1, wtps = 1, wstripes = 0;
  uint32_t psplits_s = 1, psps_s = 1, psplits_t = 1, ptps_t = 1;
  // fused single-pass perform_op (fpb_fused.cuh)
  bool use_fused = false;
  fpb::TmaDesc tm_f;                       // box 128 B x kFRows rows over gs
  uint32_t f_grid = 0, f_nslabs = 0, f_nstripes = 0, f_gpad = 0, f_window = 4, f_pol1 = 0, f_pol2 = 1, f_prefetch = 1;
  double *d_fpart = nullptr, *d_ybuf = nullptr, *d_arep = nullptr;
  uint64_t f_rep_stride = 0;
  uint32_t* d_fsync = nullptr;             // watchdog error word
  unsigned long long* d_fdbg = nullptr;    // FPB_FUSED_DEBUG: time stamps of two CTAs
  bool fused_used = false;                 // an op ran through the fused kernel since the last check
  // second set of per-vector scratch for the two-vector kernels (allocated on first use)
  struct Scratch {
    uint4* slices = nullptr;
    double *part = nullptr, *a = nullptr, *corr = nullptr, *pmax = nullptr, *psum = nullptr;
    double *mx = nullptr, *mc = nullptr;
    fpb::VecScale* sc = nullptr;
    uint32_t nparts = 0;
  } L1;
  size_t slice_bytes = 0, part_elems = 0, max_parts = 0;
  // tcgen05 block path (fpb_umma.cuh, fpb_host_umma.inl): 8 lanes of per-vector scratch, allocated
  // on the first block call with k >= 3
  struct Umma {
    bool ready = false, used = false;
    uint8_t *s_i = nullptr, *s_j = nullptr;      // digit slices of the two halves, all lanes interleaved
    double *part = nullptr, *a = nullptr, *corr = nullptr, *pmax = nullptr, *psum = nullptr;
    double *mx = nullptr, *mc = nullptr;
    fpb::VecScale* sc = nullptr;                  // [lane] first half, [8 + lane] second half
    uint32_t* err = nullptr;                      // watchdog word
    uint32_t nst = 0, splits_t = 1, sps_t = 1, nbox = 0, splits_v = 1, bps_v = 1;
    uint64_t sstride_t = 0, sstride_v = 0, vstride = 0;
    size_t si_bytes = 0, sj_bytes = 0;
  } U;
  // out-of-HBM streaming (fpb_create_streaming): the parent owns SNP slabs ("kids": ordinary
  // handles whose genotypes live in pinned host memory), two device slab buffers and a copy stream
  std::vector<fpb_handle*> kids;
  std::vector<uint8_t*> kid_host;          // pinned recoded genotypes of each kid (pitch_s x nsnps_kid)
  std::vector<uint64_t> kid_off;           // first SNP of each kid
  uint8_t* sbuf[2] = {nullptr, nullptr};
  long long sbuf_holds[2] = {-1, -1};       // slab currently resident in each buffer
  fpb::TmaDesc tm_s_alt[2], tm_f_alt[2];   // kid: tensor maps over the parent's two slab buffers
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  double* d_ytmp = nullptr;
  bool borrowed = false;                   // kid: streams and d_gs belong to the parent
  bool shared_scratch = false;             // kid: per-op scratch (slices, partials, a/corr, gather sums)
                                           // belongs to the parent, one set sized for the largest slab
  std::vector<char> kid_host_pinned;       // parent: kid_host[b] came from cudaMallocHost (else malloc)
  // pinned bounce buffers for large device -> pageable-host downloads (fpb_pca eigenvectors)
  unsigned char* h_bounce[2] = {nullptr, nullptr};
  cudaEvent_t ev_bounce[2] = {nullptr, nullptr};
  // host-pointer API staging (grown on demand)
  double* d_in = nullptr;
  double* d_out = nullptr;
  size_t in_cap = 0, out_cap = 0;
  double trace = 0.0;
  NcclComm comm = nullptr;
  int nranks = 1, rank = 0;
  // peer-memory shard sum (fpb_peer.cuh, fpb_host_peer.inl): replaces ncclAllReduce when the
  // exchange regions of all ranks could be mapped
  struct Peer {
    bool ok = false;
    void* region = nullptr;                     // this rank's exchange region
    uint32_t* h_err = nullptr;                  // watchdog word (mapped pinned host memory)
    void* mapped[fpb::kPeerMax] = {};           // IPC mappings of the other ranks' regions
    fpb::PeerView view = {};
    uint64_t cap = 0;                           // doubles per exchange
    uint32_t grid = 0;                          // CTAs of the exchange kernel (same on every rank)
    bool fuse = false;                          // the op in flight may sum inside its finalize kernel
    bool summed = false;                        // ... and did
    bool used = false;
  } P;
  uint64_t launches = 0;
  std::vector<float> op_ms;
  std::vector<double> last_evals;  // Ritz values of the last fpb_pca call (fpb_pca_residual)
  // CUDA graphs of the single-vector perform_op, one per (x, y) pointer pair: the ~14 launches of
  // an op (two streams, one ncclAllReduce when sharded) replay as one graph launch
  struct OpGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;
    uint32_t seen = 0;
  };
  std::unordered_map<uint64_t, OpGraph> op_graphs;
  bool graphs_ok = true;           // cleared when a capture fails: plain launches from then on
  fpb::Irlm* solver = nullptr;  // Lanczos workspace, kept between fpb_pca calls
  fpb::BlockKrylov* bsolver = nullptr;  // block Krylov workspace (fpb_pca_block)
  bool last_solve_block = false;        // which solver holds the eigenvectors of the last solve
  // fpb_time_perform_op: events bracketing the two contraction-kernel launches
  bool time_gemv = false;
  cudaEvent_t kev[4] = {nullptr, nullptr, nullptr, nullptr};
  double pca_phase_s[4] = {0, 0, 0, 0};  // setup+iterate, eigenvectors, download, total
  std::string err;
};

#define FPB_FAIL(h, msg)                           \
  do {                                             \
    g_err = (msg);                                 \
    if (h) (h)->err = g_err;                       \
    return 1;                                      \
  } while (0)

#define FPB_CUDA(h, call)                                                                 \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      g_err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;        \
      if (h) (h)->err = g_err;                                                            \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

namespace {

Tiling pick_tiling(uint64_t words_per_row, uint64_t nsnps, int sm_count) {
  Tiling t;
  t.W = words_per_row >= 4096 ? 2 : 1;
  const uint32_t cands[] = {256, 224, 192, 160, 128};
  uint64_t best_waste = ~0ull;
  for (uint32_t b : cands) {
    uint64_t per = (uint64_t)b * t.W;
    uint64_t chunks = (words_per_row + per - 1) / per;
    uint64_t waste = chunks * per - words_per_row;
    if (waste < best_waste) {
      best_waste = waste;
      t.block = b;
      t.chunks = (uint32_t)chunks;
    }
  }
  uint64_t target = (uint64_t)sm_count * 8;
  uint64_t splits = std::max<uint64_t>(1, target / t.chunks);
  splits = std::min<uint64_t>(splits, std::max<uint64_t>(1, nsnps / 64));
  splits = std::min<uint64_t>(splits, 65535);
  t.snps_per_split = (uint32_t)((nsnps + splits - 1) / splits);
  t.splits = (uint32_t)((nsnps + t.snps_per_split - 1) / t.snps_per_split);
  return t;
}

constexpr int kWarps = 8;           // warps per CTA of k_imma_gemv (128 rows)
constexpr uint32_t kVecBlocks = 256;  // blocks of the vector max/sum reduction

int alloc_common(fpb_handle* h, uint64_t n, uint64_t nsnps, int stand_method, int device) {
  if (n == 0 || nsnps == 0) FPB_FAIL(h, "empty genotype matrix (N == 0 or nsnps == 0)");
  if (nsnps > 0xFFFFFFF0ull || n > 0xFFFFFFF0ull) FPB_FAIL(h, "matrix too large for one handle");
  if (stand_method != FPB_STANDARDISE_BINOM && stand_method != FPB_STANDARDISE_BINOM2)
    FPB_FAIL(h, std::string("unknown standardisation method: ") + std::to_string(stand_method));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    FPB_FAIL(h, std::string("no usable CUDA device (flashpca_b200 has no CPU fallback): ") +
                    cudaGetErrorString(e));
  if (device < 0 || device >= ndev) FPB_FAIL(h, "invalid CUDA device ordinal");
  h->device = device;
  FPB_CUDA(h, cudaSetDevice(device));
  cudaDeviceProp prop;
  FPB_CUDA(h, cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  // main stream: everything ordered; side stream: the sparse missing-genotype
  // gathers (they only depend on the input vector, so they run ahead of the
  // contraction kernel and join before the finalize step)
  if (getenv("FPB_DEBUG_PRIO")) {  // debug: interleave the two streams' blocks on the SMs
    int lo = 0, hi = 0;
    FPB_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    FPB_CUDA(h, cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));
    FPB_CUDA(h, cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, lo));
  } else {
    FPB_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    FPB_CUDA(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  }
  FPB_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  FPB_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  h->n = n;
  h->nsnps = nsnps;
  h->np = (n + 3) / 4;
  h->pitch_s = (h->np + 63) / 64 * 64;
  h->pitch_i = ((nsnps + 3) / 4 + 63) / 64 * 64;
  h->stand_method = stand_method;
  if (g_borrow_gs) h->d_gs = g_borrow_gs;
  else FPB_CUDA(h, cudaMalloc(&h->d_gs, h->pitch_s * nsnps));
  FPB_CUDA(h, cudaMalloc(&h->d_lut, sizeof(double4) * nsnps));
  FPB_CUDA(h, cudaMalloc(&h->d_scale, sizeof(double2) * nsnps));
  FPB_CUDA(h, cudaMalloc(&h->d_meansd, sizeof(double) * 2 * nsnps));
  FPB_CUDA(h, cudaMalloc(&h->d_t, sizeof(double) * nsnps));
  h->tl = pick_tiling(h->pitch_s / 4, nsnps, h->sm_count);
  return 0;
}

// exclusive scan of per-row counts on the host -> device rowptr (rows + 1)
int build_rowptr(fpb_handle* h, const uint32_t* d_counts, uint64_t rows, uint64_t** d_rowptr,
                 uint64_t* total) {
  std::vector<uint32_t> cnt(rows);
  FPB_CUDA(h, cudaMemcpyAsync(cnt.data(), d_counts, sizeof(uint32_t) * rows,
                              cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  std::vector<uint64_t> ptr(rows + 1);
  uint64_t acc = 0;
  for (uint64_t r = 0; r < rows; r++) {
    ptr[r] = acc;
    acc += cnt[r];
  }
  ptr[rows] = acc;
  *total = acc;
  FPB_CUDA(h, cudaMalloc(d_rowptr, sizeof(uint64_t) * (rows + 1)));
  FPB_CUDA(h, cudaMemcpyAsync(*d_rowptr, ptr.data(), sizeof(uint64_t) * (rows + 1),
                              cudaMemcpyHostToDevice, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// Column splits of the contraction grid: enough CTAs for >= ~20 waves (2 CTAs
// per SM resident) so the last partial wave costs little, but >= 16 chunks per
// split so the pipeline prologue stays amortised.
void pick_splits(uint32_t rows, uint32_t nchunks, int sm_count, uint32_t* splits, uint32_t* cps) {
  uint32_t tiles = (rows + 16 * kWarps - 1) / (16 * kWarps);
  uint32_t want = (40u * sm_count + tiles - 1) / tiles;
  uint32_t s = std::min<uint32_t>(want, std::max<uint32_t>(1, nchunks / 16));
  s = std::max<uint32_t>(1, std::min<uint32_t>(s, 64));
  *cps = (nchunks + s - 1) / s;
  *splits = (nchunks + *cps - 1) / *cps;
}

// Column-blocked, row-sliced (SELL-32 per tile) lists (fpb_imma.cuh) from a CSR
// (d_rowptr: rows + 1, d_col: ascending columns per row).  Returns 2 when the
// padded layout would be wasteful (very uneven rows): the caller then uses the
// generic path.
int build_blocked_lists(fpb_handle* h, uint64_t rows, uint64_t veclen, const uint64_t* d_rowptr,
                        const uint32_t* d_col, uint64_t** d_blkoff, uint16_t** d_col16,
                        uint32_t* ntiles_out) {
  const uint32_t ntiles = (uint32_t)((veclen + fpb::kGatherTile - 1) / fpb::kGatherTile);
  const uint32_t nblk = (uint32_t)((rows + 31) / 32);
  const uint32_t gw = (uint32_t)((rows * 32 + 255) / 256);
  uint32_t *d_counts = nullptr, *d_sizes = nullptr;
  FPB_CUDA(h, cudaMalloc(&d_counts, sizeof(uint32_t) * (size_t)ntiles * rows));
  FPB_CUDA(h, cudaMalloc(&d_sizes, sizeof(uint32_t) * (size_t)ntiles * nblk));
  FPB_CUDA(h, cudaMemsetAsync(d_counts, 0, sizeof(uint32_t) * (size_t)ntiles * rows, h->stream));
  fpb::k_bcsr_count<<<gw, 256, 0, h->stream>>>(d_rowptr, d_col, rows, d_counts);
  uint64_t nwarps = (uint64_t)ntiles * nblk;
  fpb::k_sell_sizes<<<(uint32_t)((nwarps * 32 + 255) / 256), 256, 0, h->stream>>>(
      d_counts, rows, nblk, ntiles, d_sizes);
  uint64_t padded = 0;
  if (build_rowptr(h, d_sizes, nwarps, d_blkoff, &padded)) return 1;
  int rc = 0;
  if (padded > 4 * h->nmissing + (64ull << 20)) {
    rc = 2;  // padding would dominate
  } else {
    FPB_CUDA(h, cudaMalloc(d_col16, sizeof(uint16_t) * std::max<uint64_t>(padded, 1)));
    FPB_CUDA(h, cudaMemsetAsync(*d_col16, 0x30, sizeof(uint16_t) * padded, h->stream));
    fpb::k_sell_fill<<<gw, 256, 0, h->stream>>>(d_rowptr, d_col, rows, nblk, *d_blkoff, d_counts,
                                                *d_col16);
    h->launches += 3;
    FPB_CUDA(h, cudaStreamSynchronize(h->stream));
    FPB_CUDA(h, cudaGetLastError());
  }
  cudaFree(d_counts);
  cudaFree(d_sizes);
  *ntiles_out = ntiles;
  return rc;
}

// Transpose a CSR (rows_in x rows_out pattern) with a stable radix sort of
// (column << 32 | row) keys: the result lists ascend, so every later sum has a
// fixed order.  Staging only; CUB is library code outside the hot path.
int transpose_csr(fpb_handle* h, uint64_t rows_in, uint64_t rows_out, const uint64_t* d_rowptr,
                  const uint32_t* d_col, uint64_t** d_rowptr_t, uint32_t** d_col_t) {
  const uint64_t nnz = h->nmissing;
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  uint32_t* d_cnt = nullptr;
  void* d_tmp = nullptr;
  FPB_CUDA(h, cudaMalloc(&d_keys, sizeof(uint64_t) * nnz));
  FPB_CUDA(h, cudaMalloc(&d_keys2, sizeof(uint64_t) * nnz));
  FPB_CUDA(h, cudaMalloc(&d_cnt, sizeof(uint32_t) * rows_out));
  FPB_CUDA(h, cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t) * rows_out, h->stream));
  fpb::k_csr_to_keys<<<(uint32_t)((rows_in * 32 + 255) / 256), 256, 0, h->stream>>>(
      d_rowptr, d_col, rows_in, d_keys);
  size_t tmp_bytes = 0;
  FPB_CUDA(h, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys2, nnz, 0, 64,
                                             h->stream));
  FPB_CUDA(h, cudaMalloc(&d_tmp, tmp_bytes));
  FPB_CUDA(h, cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, nnz, 0, 64,
                                             h->stream));
  FPB_CUDA(h, cudaMalloc(d_col_t, sizeof(uint32_t) * nnz));
  fpb::k_keys_to_csr<<<(uint32_t)((nnz + 255) / 256), 256, 0, h->stream>>>(d_keys2, nnz, *d_col_t,
                                                                          d_cnt);
  h->launches += 2;
  uint64_t total = 0;
  int rc = build_rowptr(h, d_cnt, rows_out, d_rowptr_t, &total);
  cudaFree(d_keys);
  cudaFree(d_keys2);
  cudaFree(d_cnt);
  cudaFree(d_tmp);
  if (rc) return 1;
  if (total != nnz) FPB_FAIL(h, "internal error: transposed list size mismatch");
  return 0;
}

// Column/row splits of a TMA contraction grid (1 CTA per SM resident).  With
// `units` independent tiles and `nstages` pipeline stages to divide per tile, a
// split count s gives units*s CTAs of ceil(nstages/s) stages each; the kernel
// takes about ceil(units*s / SMs) rounds of (stages + pipeline fill).  Pick the s
// that minimises that estimate (keeps the last round full and the fill amortised).
void pick_splits_waves(uint32_t units, uint32_t nstages, int sm_count, uint32_t* splits,
                       uint32_t* per_split) {
  const double fill = 4.0;  // pipeline fill + epilogue of a CTA, in stage times (~3 us / 0.7 us)
  uint32_t best = 1;
  double best_cost = 1e300;
  const uint32_t smax = std::max<uint32_t>(1, std::min<uint32_t>(64, nstages / 8));
  for (uint32_t sp = 1; sp <= smax; sp++) {
    const uint32_t per = (nstages + sp - 1) / sp;
    const uint32_t real = (nstages + per - 1) / per;  // splits actually launched
    const double rounds = std::ceil((double)units * real / sm_count);
    const double cost = rounds * (per + fill);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = sp;
    }
  }
  *per_split = (nstages + best - 1) / best;
  *splits = (nstages + *per_split - 1) / *per_split;
}

// Splits of the reduction axis for the persistent kernels: `units` output tiles x s splits
// are dealt round-robin to one CTA per SM; pick the s that minimises
// (most items any CTA gets) x (stages per item + epilogue).
void pick_splits_persist(uint32_t units, uint32_t nstages, int sm_count, uint32_t* splits,
                         uint32_t* per_split) {
  const uint32_t G = (uint32_t)sm_count;
  const uint32_t smax = std::max<uint32_t>(1, std::min<uint32_t>(64, nstages / 12));
  uint32_t best = 1;
  double best_cost = 1e300;
  for (uint32_t sp = 1; sp <= smax; sp++) {
    const uint32_t per = (nstages + sp - 1) / sp;
    const uint32_t real = (nstages + per - 1) / per;
    const uint64_t items = (uint64_t)units * real;
    const double rounds = std::ceil((double)items / G);
    const double cost = rounds * (per + 1.5);
    if (cost < best_cost * 0.9999) {
      best_cost = cost;
      best = sp;
    }
  }
  *per_split = (nstages + best - 1) / best;
  *splits = (nstages + *per_split - 1) / *per_split;
}

// debug: FPB_DEBUG_SPLITS1 / FPB_DEBUG_SPLITS2 override the split count of the first /
// second contraction kernel (tuning sweeps)
void override_splits(const char* env, uint32_t nstages, uint32_t* splits, uint32_t* per_split) {
  const char* v = getenv(env);
  if (!v) return;
  uint32_t sp = (uint32_t)std::max(1, atoi(v));
  sp = std::min<uint32_t>(sp, nstages);
  *per_split = (nstages + sp - 1) / sp;
  *splits = (nstages + *per_split - 1) / *per_split;
}

void pick_splits_tma(uint32_t rows, uint32_t nstages, int sm_count, uint32_t* splits,
                     uint32_t* sps) {
  pick_splits_waves((rows + fpb::kTmaRows - 1) / fpb::kTmaRows, nstages, sm_count, splits, sps);
}

// 2-D uint8 tensor map over a packed matrix: dim0 = bytes of a row, dim1 = rows;
// box = 128 B x box_rows rows, 128-byte swizzle, out-of-bounds filled with zeros.
int make_tensor_map(fpb_handle* h, const uint8_t* base, uint64_t pitch, uint64_t rows,
                    fpb::TmaDesc* out, uint32_t box_rows = fpb::kTmaRows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FPB_CUDA(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      FPB_FAIL(h, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeFn)fn;
  }
  static_assert(sizeof(CUtensorMap) == sizeof(fpb::TmaDesc), "tensor map size");
  cuuint64_t dims[2] = {pitch, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {(cuuint32_t)fpb::kTmaStageCols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                       const_cast<uint8_t*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
    FPB_FAIL(h, std::string("cuTensorMapEncodeTiled failed, code ") + std::to_string((int)rc));
  return 0;
}

#include "fpb_host_fused_setup.inl"

// raw bed bytes are in d_gs (pitch_s): recode, statistics (data.cpp:257-322),
// trace, and -- when the missing rate allows -- the tensor path's second copy
// and CSR lists.
int finish_create(fpb_handle* h, const double* preloaded_meansd) {
  uint64_t nwords = h->nsnps * (h->pitch_s / 4);
  fpb::k_recode_rows<<<(uint32_t)((nwords + 255) / 256), 256, 0, h->stream>>>(h->d_gs, h->nsnps,
                                                                              h->n, h->pitch_s);
  h->launches++;
  if (preloaded_meansd)
    FPB_CUDA(h, cudaMemcpyAsync(h->d_meansd, preloaded_meansd, sizeof(double) * 2 * h->nsnps,
                                cudaMemcpyHostToDevice, h->stream));
  double* d_tracej = nullptr;
  uint32_t* d_cnt = nullptr;
  FPB_CUDA(h, cudaMalloc(&d_tracej, sizeof(double) * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&d_cnt, sizeof(uint32_t) * std::max(h->nsnps, h->n)));
  uint32_t gs = (uint32_t)((h->nsnps * 32 + 255) / 256);
  fpb::k_snp_stats<<<gs, 256, 0, h->stream>>>(h->d_gs, h->nsnps, h->n, h->pitch_s,
                                              h->stand_method, preloaded_meansd != nullptr,
                                              h->d_meansd, h->d_lut, h->d_scale, d_tracej, d_cnt);
  h->launches++;
  std::vector<double> tr(h->nsnps);
  FPB_CUDA(h, cudaMemcpyAsync(tr.data(), d_tracej, sizeof(double) * h->nsnps,
                              cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  FPB_CUDA(h, cudaGetLastError());
  double s = 0.0;
  for (double v : tr) s += v;  // SNP order, like the block loop of svdwide.cpp:44-61
  h->trace = s;
  cudaFree(d_tracej);

  if (build_rowptr(h, d_cnt, h->nsnps, &h->d_rowptr_s, &h->nmissing)) return 1;
  const char* force = getenv("FPB_PATH");
  h->use_imma = (double)h->nmissing <= 0.03 * (double)h->n * (double)h->nsnps;
  if (force && !strcmp(force, "generic")) h->use_imma = false;
  if (force && !strcmp(force, "imma")) h->use_imma = true;

  if (!h->use_imma) {
    cudaFree(h->d_rowptr_s);
    h->d_rowptr_s = nullptr;
    cudaFree(d_cnt);
    FPB_CUDA(h, cudaMalloc(&h->d_coef, sizeof(double4) * h->nsnps));
    FPB_CUDA(h, cudaMalloc(&h->d_c0, sizeof(double)));
    return 0;
  }

  // ---- tensor path staging: kernel variant, missing-genotype lists, scratch
  {
    const char* gv = getenv("FPB_GEMV");
    h->use_tma = !(gv && !strcmp(gv, "ldg"));
    h->single_copy = h->use_tma && !(gv && !strcmp(gv, "tma2"));  // tma2 = two-copy TMA variant
  }
  if (!h->single_copy) {
    // two-copy variants keep an individual-major copy for the second half
    FPB_CUDA(h, cudaMalloc(&h->d_gi, h->pitch_i * h->n));
    FPB_CUDA(h, cudaMemsetAsync(h->d_gi, 0, h->pitch_i * h->n, h->stream));
    dim3 grid((uint32_t)((h->n + 127) / 128), (uint32_t)((h->nsnps + 127) / 128));
    fpb::k_transpose_2bit<<<grid, 256, 0, h->stream>>>(h->d_gs, h->nsnps, h->pitch_s, h->d_gi,
                                                       h->n, h->pitch_i);
    h->launches++;
  }
  if (h->nmissing) {
    // CSR by SNP (columns = individuals, ascending), then its transpose by individual
    uint32_t *d_col_s = nullptr, *d_col_i = nullptr;
    uint64_t* d_rowptr_i = nullptr;
    FPB_CUDA(h, cudaMalloc(&d_col_s, sizeof(uint32_t) * h->nmissing));
    fpb::k_fill_missing_csr<<<gs, 256, 0, h->stream>>>(h->d_gs, h->nsnps, h->pitch_s,
                                                       h->d_rowptr_s, d_col_s);
    h->launches++;
    int rc = build_blocked_lists(h, h->nsnps, h->n, h->d_rowptr_s, d_col_s, &h->d_seg_s,
                                 &h->d_col16_s, &h->gtiles_s);
    if (rc == 0) {
      rc = transpose_csr(h, h->nsnps, h->n, h->d_rowptr_s, d_col_s, &d_rowptr_i, &d_col_i);
      if (rc == 0)
        rc = build_blocked_lists(h, h->n, h->nsnps, d_rowptr_i, d_col_i, &h->d_seg_i,
                                 &h->d_col16_i, &h->gtiles_i);
    }
    cudaFree(d_col_s);
    cudaFree(d_col_i);
    cudaFree(d_rowptr_i);
    if (rc == 1) return 1;
    if (rc == 2) {
      // very uneven missingness: padded lists would dominate -> generic FP64 path
      cudaFree(h->d_gi);
      h->d_gi = nullptr;
      cudaFree(h->d_rowptr_s);
      h->d_rowptr_s = nullptr;
      cudaFree(d_cnt);
      h->use_imma = false;
      FPB_CUDA(h, cudaMalloc(&h->d_coef, sizeof(double4) * h->nsnps));
      FPB_CUDA(h, cudaMalloc(&h->d_c0, sizeof(double)));
      return 0;
    }
  }
  cudaFree(h->d_rowptr_s);
  h->d_rowptr_s = nullptr;
  cudaFree(d_cnt);
  h->nchunks_s = (uint32_t)((h->pitch_s + fpb::kChunkBytes - 1) / fpb::kChunkBytes);
  h->nchunks_i = (uint32_t)((h->pitch_i + fpb::kChunkBytes - 1) / fpb::kChunkBytes);
  pick_splits((uint32_t)h->nsnps, h->nchunks_s, h->sm_count, &h->splits_s, &h->cps_s);
  pick_splits((uint32_t)h->n, h->nchunks_i, h->sm_count, &h->splits_i, &h->cps_i);
  h->part_stride = std::max(h->n, h->nsnps);
  {
    h->nstages_s = (uint32_t)((h->pitch_s + fpb::kTmaStageCols - 1) / fpb::kTmaStageCols);
    h->nstages_i = (uint32_t)((h->pitch_i + fpb::kTmaStageCols - 1) / fpb::kTmaStageCols);
    pick_splits_tma((uint32_t)h->nsnps, h->nstages_s, h->sm_count, &h->tsplits_s, &h->sps_s);
    override_splits("FPB_DEBUG_SPLITS1", h->nstages_s, &h->tsplits_s, &h->sps_s);
    pick_splits_tma((uint32_t)h->n, h->nstages_i, h->sm_count, &h->tsplits_i, &h->sps_i);
    if (h->single_copy) {
      // second half over gs: CTAs = 128-byte column stripes x splits of the 256-row tiles
      h->ttiles = (uint32_t)((h->nsnps + fpb::kTmaRows - 1) / fpb::kTmaRows);
      // measured (tools/sweep_splits.sh, 500k x 100k): the column-stripe kernel is fastest with
      // ~48 row tiles per CTA (8 splits: 1.87 ms vs 1.92 at 3 and 2.01 at 24); few stripes
      // (small N) need more splits to occupy every SM
      uint32_t sp = std::max<uint32_t>((h->ttiles + 24) / 48,
                                       (h->sm_count + h->nstages_s - 1) / h->nstages_s);
      sp = std::max<uint32_t>(1, std::min<uint32_t>(sp, std::min<uint32_t>(64, std::max<uint32_t>(1, h->ttiles / 8))));
      h->ttps = (h->ttiles + sp - 1) / sp;
      h->ttsplits = (h->ttiles + h->ttps - 1) / h->ttps;
      override_splits("FPB_DEBUG_SPLITS2", h->ttiles, &h->ttsplits, &h->ttps);
    }
    {
      // Measured (profiles/r01_persist_sweep.txt): at 10k x 100k the persistent kernels are 9 % faster
      // (0.141 -> 0.129 ms per op); at 500k x 100k the first half is unchanged (1.75 ms) and the
      // second half is slower in lockstep (2.6 vs 1.88 ms), so large grids stay one-shot.
      const char* pv = getenv("FPB_PERSIST");
      const uint32_t ntile_s = (uint32_t)((h->nsnps + fpb::kTmaRows - 1) / fpb::kTmaRows);
      const uint32_t few = 10u * (uint32_t)h->sm_count;
      h->persist_s = h->single_copy && (pv ? atoi(pv) != 0 : ntile_s * h->tsplits_s < few);
      // second half: only with few column stripes (small N); with many stripes the persistent
      // CTAs run in lockstep over the SNP rows and lose DRAM efficiency (8-GPU shard 500k x 12.5k:
      // 0.345 vs 0.275 ms, profiles/r01_scaling_persist.txt)
      h->persist_t = h->single_copy && (pv ? atoi(pv) != 0 : h->nstages_s < (uint32_t)h->sm_count);
      if (h->persist_s) {
        pick_splits_persist(ntile_s, h->nstages_s, h->sm_count, &h->psplits_s, &h->psps_s);
        override_splits("FPB_DEBUG_SPLITS1", h->nstages_s, &h->psplits_s, &h->psps_s);
      }
      if (h->persist_t) {
        pick_splits_persist(h->nstages_s, h->ttiles, h->sm_count, &h->psplits_t, &h->ptps_t);
        override_splits("FPB_DEBUG_SPLITS2", h->ttiles, &h->psplits_t, &h->ptps_t);
      }
    }
    if (h->single_copy) {
      // L2 window: the last SNP rows read by the first half stay in L2 (evict_last) and the second
      // half finds them there.  Measured (profiles/r02_l2_keep_window_sweep.txt): -3 % per op at
      // 10k x 100k (250 MB) with 24-64 MB, nothing at 1.5 GB and above -- on for matrices below
      // 1 GB.  FPB_L2_KEEP_MB overrides the size (0 = off).
      const char* kv = getenv("FPB_L2_KEEP_MB");
      const double keep_mb = kv ? atof(kv) : ((double)h->pitch_s * (double)h->nsnps < 1e9 ? 32.0 : 0.0);
      const uint32_t ntile_s = (uint32_t)((h->nsnps + fpb::kTmaRows - 1) / fpb::kTmaRows);
      const uint64_t tile_bytes = (uint64_t)h->pitch_s * fpb::kTmaRows;
      const uint32_t keep_tiles = (uint32_t)std::min<uint64_t>(ntile_s, (uint64_t)(keep_mb * 1e6) / tile_bytes);
      h->l2_keep_row = keep_tiles ? (ntile_s - keep_tiles) * fpb::kTmaRows : 0xFFFFFFFFu;
    }
    if (h->use_tma) {
      if (make_tensor_map(h, h->d_gs, h->pitch_s, h->nsnps, &h->tm_s)) return 1;
      if (!h->single_copy && make_tensor_map(h, h->d_gi, h->pitch_i, h->n, &h->tm_i)) return 1;
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_t,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_p,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_t_p,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
    }
    if (h->single_copy) {
      if (make_tensor_map(h, h->d_gs, h->pitch_s, h->nsnps, &h->tm_f, fpb::kFRows)) return 1;
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_2v,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_t_2v,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_imma_gemv_tma_tw,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fpb::kTmaSmemBytes));
      // wide-stripe second half: same SNP ranges per split as the narrow kernel
      const char* wv = getenv("FPB_P2WIDE");
      h->wstripes = (uint32_t)((h->pitch_s + 255) / 256);
      h->wtiles = (uint32_t)((h->nsnps + fpb::kTmaWRows - 1) / fpb::kTmaWRows);
      h->wtps = 2 * h->ttps;
      h->wsplits = (h->wtiles + h->wtps - 1) / h->wtps;
      // measured at 500k x 100k: 1.94 ms vs 1.93 ms for the 128-byte stripes (profiles/
      // r01_wide_sweep.txt) -- no gain, so it is a variant (FPB_P2WIDE=1), not the default
      h->wide_t = wv && atoi(wv) == 1;
      if (setup_fused(h)) return 1;
    }
  }
  uint64_t max_chunks = std::max(h->nchunks_s, h->nchunks_i);
  // the K-major slices of the single-copy second half need 8 bytes per SNP, padded to whole tiles
  uint64_t slice_bytes = std::max<uint64_t>(sizeof(uint4) * max_chunks * fpb::kChunkWords * 8,
                                            (uint64_t)(h->ttiles + 1) * fpb::kTmaTSliceBytes);
  h->slice_bytes = slice_bytes;
  FPB_CUDA(h, cudaMalloc(&h->d_slices, slice_bytes));
  h->part_elems = h->part_stride *
                  std::max(std::max(std::max(std::max(h->splits_s, h->splits_i),
                                             std::max(h->tsplits_s, h->tsplits_i)),
                                    h->ttsplits),
                           std::max(std::max(h->psplits_s, h->psplits_t), h->wsplits));
  FPB_CUDA(h, cudaMalloc(&h->d_part, sizeof(double) * h->part_elems));
  FPB_CUDA(h, cudaMalloc(&h->d_a, sizeof(double) * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&h->d_corr, sizeof(double) * h->nsnps));
  const size_t max_parts = std::max<size_t>(kVecBlocks, (h->nsnps + 255) / 256);
  h->max_parts = max_parts;
  FPB_CUDA(h, cudaMalloc(&h->d_pmax, sizeof(double) * max_parts));
  FPB_CUDA(h, cudaMalloc(&h->d_psum, sizeof(double) * max_parts));
  FPB_CUDA(h, cudaMalloc(&h->d_sc, sizeof(fpb::VecScale) * 2));
  if (h->nmissing) {
    FPB_CUDA(h, cudaMalloc(&h->d_mx, sizeof(double) * h->nsnps * h->gtiles_s));
    FPB_CUDA(h, cudaMalloc(&h->d_mc, sizeof(double) * h->n * h->gtiles_i));
    FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_sell_gather,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     fpb::kGatherSmem));
    FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_sell_gather_p,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     fpb::kTmaSmemBytes));
    // FPB_GATHER_SMS = n: the single-vector op's gathers run on n dedicated SMs beside the
    // contraction kernel (0: on all SMs in front of it)
    if (const char* g = getenv("FPB_GATHER_SMS")) h->gather_sms = std::max(0, atoi(g));
  }
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  FPB_CUDA(h, cudaGetLastError());
  return 0;
}

int ensure_staging(fpb_handle* h, size_t in_elems, size_t out_elems) {
  if (in_elems > h->in_cap) {
    if (h->d_in) cudaFree(h->d_in);
    h->d_in = nullptr;
    FPB_CUDA(h, cudaMalloc(&h->d_in, sizeof(double) * in_elems));
    h->in_cap = in_elems;
  }
  if (out_elems > h->out_cap) {
    if (h->d_out) cudaFree(h->d_out);
    h->d_out = nullptr;
    FPB_CUDA(h, cudaMalloc(&h->d_out, sizeof(double) * out_elems));
    h->out_cap = out_elems;
  }
  return 0;
}

// ------------------------------ generic FP64 path --------------------------

// per-(chunk, warp) partials of t and per-split partials of y: summed in a fixed order, so the
// generic path is as bit-reproducible as the tensor path
int ensure_generic_scratch(fpb_handle* h) {
  if (h->d_gpart) return 0;
  const Tiling& t = h->tl;
  const size_t rows_t = (size_t)t.chunks * (t.block / 32);
  const size_t need = std::max(rows_t * h->nsnps, (size_t)t.splits * h->n);
  FPB_CUDA(h, cudaMalloc(&h->d_gpart, sizeof(double) * need));
  return 0;
}

void generic_crossprod(fpb_handle* h, const double* d_x, double* d_t) {
  const Tiling& t = h->tl;
  if (ensure_generic_scratch(h)) return;
  dim3 grid(t.chunks, t.splits);
  if (t.W == 1)
    fpb::k_crossprod<1><<<grid, t.block, 0, h->stream>>>(h->d_gs, h->pitch_s, h->n,
                                                          (uint32_t)h->nsnps, t.snps_per_split,
                                                          d_x, h->d_lut, h->d_gpart);
  else
    fpb::k_crossprod<2><<<grid, t.block, 0, h->stream>>>(h->d_gs, h->pitch_s, h->n,
                                                          (uint32_t)h->nsnps, t.snps_per_split,
                                                          d_x, h->d_lut, h->d_gpart);
  fpb::k_sum_rows<<<(uint32_t)((h->nsnps + 255) / 256), 256, 0, h->stream>>>(
      h->d_gpart, t.chunks * (t.block / 32), h->nsnps, d_t);
  h->launches += 2;
}

void generic_prod(fpb_handle* h, const double* d_v, double* d_y) {
  const Tiling& t = h->tl;
  if (ensure_generic_scratch(h)) return;
  uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  fpb::k_prod_coef<<<gb, 256, 0, h->stream>>>(h->d_lut, d_v, (uint32_t)h->nsnps, h->d_coef);
  fpb::k_sum_a0<<<1, 1024, 0, h->stream>>>(h->d_coef, (uint32_t)h->nsnps, h->d_c0);
  double* dst = t.splits > 1 ? h->d_gpart : d_y;
  dim3 grid(t.chunks, t.splits);
  if (t.W == 1)
    fpb::k_prod<1><<<grid, t.block, 0, h->stream>>>(h->d_gs, h->pitch_s, h->n,
                                                     (uint32_t)h->nsnps, t.snps_per_split,
                                                     h->d_coef, h->d_c0, dst);
  else
    fpb::k_prod<2><<<grid, t.block, 0, h->stream>>>(h->d_gs, h->pitch_s, h->n,
                                                     (uint32_t)h->nsnps, t.snps_per_split,
                                                     h->d_coef, h->d_c0, dst);
  h->launches += 3;
  if (t.splits > 1) {
    fpb::k_sum_rows<<<(uint32_t)((h->n + 255) / 256), 256, 0, h->stream>>>(h->d_gpart, t.splits,
                                                                          h->n, d_y);
    h->launches++;
  }
}

// ------------------------------ tensor path --------------------------------

// (max|v|, sum v) block partials of a vector produced outside the library's kernels
void vec_partials(fpb_handle* h, const double* d_v, uint64_t len) {
  fpb::k_vec_partial<<<kVecBlocks, 256, 0, h->stream>>>(d_v, len, h->d_pmax, h->d_psum);
  h->nparts = kVecBlocks;
  h->launches++;
}

// part[split][row] = sum_s 128^s sum_col G[row][col] * digit_s(v[col]);
// snp_major selects gs (rows = SNPs) or gi (rows = individuals).  Returns the
// number of splits written.
uint32_t imma_contract(fpb_handle* h, bool snp_major, const double* d_v, uint64_t vlen, int slot) {
  if (!snp_major && h->single_copy) {
    // F = E a from the SNP-major copy: K-major slices of a, byte-transposing loads
    const uint32_t ngroups4 = h->ttiles * (fpb::kTmaRows / 4);
    fpb::k_slice_vec_k<<<(ngroups4 + 127) / 128, 128, 0, h->stream>>>(
        d_v, vlen, ngroups4, h->d_pmax, h->d_psum, h->nparts, h->d_sc + slot,
        reinterpret_cast<uint32_t*>(h->d_slices));
    if (h->time_gemv) cudaEventRecord(h->kev[2], h->stream);
    uint32_t used = h->ttsplits;
    if (h->wide_t) {
      dim3 grid(h->wstripes, h->wsplits);
      fpb::k_imma_gemv_tma_tw<<<grid, (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes,
                                h->stream>>>(h->tm_f, (uint32_t)h->n,
                                             reinterpret_cast<const uint32_t*>(h->d_slices),
                                             h->wtiles, h->wtps, h->d_part, h->part_stride);
      used = h->wsplits;
    } else if (h->persist_t) {
      const uint32_t nitems = h->nstages_s * h->psplits_t;
      fpb::k_imma_gemv_tma_t_p<<<std::min<uint32_t>((uint32_t)h->sm_count, nitems),
                                 (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes, h->stream>>>(
          h->tm_s, (uint32_t)h->n, reinterpret_cast<const uint32_t*>(h->d_slices), h->ttiles,
          h->ptps_t, h->psplits_t, nitems, h->d_part, h->part_stride);
      used = h->psplits_t;
    } else {
      dim3 grid(h->nstages_s, h->ttsplits);
      fpb::k_imma_gemv_tma_t<<<grid, (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes,
                               h->stream>>>(h->tm_s, (uint32_t)h->n,
                                            reinterpret_cast<const uint32_t*>(h->d_slices),
                                            h->ttiles, h->ttps, h->d_part, h->part_stride);
    }
    if (h->time_gemv) cudaEventRecord(h->kev[3], h->stream);
    h->launches += 2;
    return used;
  }
  const uint8_t* G = snp_major ? h->d_gs : h->d_gi;
  const uint64_t pitch = snp_major ? h->pitch_s : h->pitch_i;
  const uint32_t rows = (uint32_t)(snp_major ? h->nsnps : h->n);
  const uint32_t nchunks = snp_major ? h->nchunks_s : h->nchunks_i;
  uint32_t nwq = nchunks * fpb::kChunkWords;  // covers the TMA stages as well
  fpb::k_slice_vec<<<(nwq + 127) / 128, 128, 0, h->stream>>>(
      d_v, vlen, nwq, h->d_pmax, h->d_psum, h->nparts, h->d_sc + slot, h->d_slices);
  h->launches += 2;
  if (h->time_gemv) cudaEventRecord(h->kev[snp_major ? 0 : 2], h->stream);
  struct StopTimer {  // records the closing event when the launch has been enqueued
    fpb_handle* h;
    bool snp;
    ~StopTimer() {
      if (h->time_gemv) cudaEventRecord(h->kev[snp ? 1 : 3], h->stream);
    }
  } stop_timer{h, snp_major};
  if (h->use_tma && snp_major && h->persist_s) {
    const uint32_t ntile = (rows + fpb::kTmaRows - 1) / fpb::kTmaRows;
    const uint32_t nitems = ntile * h->psplits_s;
    fpb::k_imma_gemv_tma_p<<<std::min<uint32_t>((uint32_t)h->sm_count, nitems),
                             (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes, h->stream>>>(
        h->tm_s, rows, h->d_slices, h->nstages_s, h->psps_s, h->psplits_s, nitems, h->d_part,
        h->part_stride, h->l2_keep_row);
    return h->psplits_s;
  }
  if (h->use_tma) {
    const uint32_t nstages = snp_major ? h->nstages_s : h->nstages_i;
    const uint32_t splits = snp_major ? h->tsplits_s : h->tsplits_i;
    const uint32_t sps = snp_major ? h->sps_s : h->sps_i;
    dim3 grid((rows + fpb::kTmaRows - 1) / fpb::kTmaRows, splits);
    // debug: FPB_DEBUG_SMEM_MIN requests only the bytes the ring needs, which lets
    // blocks of other kernels share the SM (used to reproduce the co-residency race)
    static const int smem_bytes = getenv("FPB_DEBUG_SMEM_MIN") ? fpb::kTmaSmemUsed : fpb::kTmaSmemBytes;
    fpb::k_imma_gemv_tma<<<grid, (fpb::kTmaConsumerWarps + 1) * 32, smem_bytes,
                           h->stream>>>(snp_major ? h->tm_s : h->tm_i, rows, h->d_slices, nstages,
                                        sps, h->d_part, h->part_stride,
                                        snp_major ? h->l2_keep_row : 0xFFFFFFFFu);
    return splits;
  }
  const uint32_t splits = snp_major ? h->splits_s : h->splits_i;
  const uint32_t cps = snp_major ? h->cps_s : h->cps_i;
  dim3 grid((rows + 16 * kWarps - 1) / (16 * kWarps), splits);
  fpb::k_imma_gemv<kWarps><<<grid, kWarps * 32, 0, h->stream>>>(G, pitch, rows, h->d_slices,
                                                                nchunks, cps, h->d_part,
                                                                h->part_stride);
  return splits;
}

// Sparse missing-genotype sums on the side stream.  fork_mark() pins the point
// of the main stream the gather depends on (its input vector is complete);
// gather_launch() enqueues it (it runs ahead of / next to the slicing kernels);
// join_gather() makes the main stream wait for the result before the finalize
// kernel.
void fork_mark(fpb_handle* h) { cudaEventRecord(h->ev_fork, h->stream); }
void gather_launch(fpb_handle* h, bool by_snp, const double* vec) {
  const uint64_t nrows = by_snp ? h->nsnps : h->n, veclen = by_snp ? h->n : h->nsnps;
  const uint32_t ntiles = by_snp ? h->gtiles_s : h->gtiles_i;
  cudaStreamWaitEvent(h->side, h->ev_fork, 0);
  // about two full waves of 2 CTAs per SM (rounded down so the last wave is not a sliver)
  const uint32_t nblk = (uint32_t)((nrows + 31) / 32);
  uint32_t chunks = std::max<uint32_t>(1, (4u * h->sm_count) / ntiles);
  uint32_t blocks_per_cta = (nblk + chunks - 1) / chunks;
  chunks = (nblk + blocks_per_cta - 1) / blocks_per_cta;
  if (h->gather_sms > 0) {
    // a few dedicated SMs next to the contraction kernel instead of all SMs in front of it
    fpb::k_sell_gather_p<<<h->gather_sms, fpb::kGatherThreadsP, fpb::kTmaSmemBytes, h->side>>>(
        by_snp ? h->d_seg_s : h->d_seg_i, by_snp ? h->d_col16_s : h->d_col16_i, vec, veclen, nrows,
        nblk, blocks_per_cta, chunks, ntiles * chunks, by_snp ? h->d_mx : h->d_mc);
  } else {
    dim3 grid(ntiles, chunks);
    fpb::k_sell_gather<<<grid, fpb::kGatherThreads, fpb::kGatherSmem, h->side>>>(
        by_snp ? h->d_seg_s : h->d_seg_i, by_snp ? h->d_col16_s : h->d_col16_i, vec, veclen, nrows,
        nblk, blocks_per_cta, by_snp ? h->d_mx : h->d_mc);
  }
  cudaEventRecord(h->ev_join, h->side);
  h->launches++;
}
void join_gather(fpb_handle* h) { cudaStreamWaitEvent(h->stream, h->ev_join, 0); }

// first half: t = X'x (d_t) and/or the a, corr inputs of the second half
void imma_crossprod(fpb_handle* h, const double* d_x, double* d_t, bool second_half) {
  static const bool gather_after = getenv("FPB_DEBUG_GATHER_AFTER") != nullptr;
  if (h->nmissing) {
    fork_mark(h);
    if (!gather_after) gather_launch(h, true, d_x);
  }
  vec_partials(h, d_x, h->n);
  const uint32_t nsplits = imma_contract(h, true, d_x, h->n, 0);
  if (h->nmissing && gather_after) gather_launch(h, true, d_x);
  if (h->nmissing) join_gather(h);
  uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  fpb::k_finalize_crossprod<<<gb, 256, 0, h->stream>>>(
      h->d_part, nsplits, h->part_stride, (uint32_t)h->nsnps, h->d_sc + 0, h->d_scale,
      h->nmissing ? h->d_mx : nullptr, h->gtiles_s, d_t, second_half ? h->d_a : nullptr,
      h->d_corr, h->d_pmax, h->d_psum);
  if (second_half) h->nparts = gb;
  h->launches++;
}

void peer_finalize_prod(fpb_handle* h, uint32_t nsplits, double* d_y);

// second half from a, corr and the (max|a|, sum b) partials already in the handle
void imma_prod_tail(fpb_handle* h, double* d_y) {
  static const bool gather_after = getenv("FPB_DEBUG_GATHER_AFTER") != nullptr;
  if (h->nmissing) {
    fork_mark(h);
    if (!gather_after) gather_launch(h, false, h->d_corr);
  }
  const uint32_t nsplits = imma_contract(h, false, h->d_a, h->nsnps, 1);
  if (h->nmissing && gather_after) gather_launch(h, false, h->d_corr);
  if (h->nmissing) join_gather(h);
  if (h->P.ok && h->P.fuse) return peer_finalize_prod(h, nsplits, d_y);
  uint32_t gb = (uint32_t)((h->n + 255) / 256);
  fpb::k_finalize_prod<<<gb, 256, 0, h->stream>>>(h->d_part, nsplits, h->part_stride, h->n,
                                                  h->d_sc + 1, h->nmissing ? h->d_mc : nullptr,
                                                  h->gtiles_i, d_y);
  h->launches++;
}

#include "fpb_host_fused.inl"

#include "fpb_host_pairs.inl"

#include "fpb_host_umma.inl"

#include "fpb_host_peer.inl"

// device-reported protocol failures of the persistent / tcgen05 kernels, at the API's sync points
int check_async(fpb_handle* h) {
  if (check_fused(h)) return 1;
  if (check_peer(h)) return 1;
  return check_umma(h);
}

// in-memory matrix path (svdwide.cpp:4-12)
void dense_crossprod(fpb_handle* h, const double* d_x, double* d_t) {
  fpb::k_dense_gemv_t<<<(uint32_t)h->nsnps, 256, 0, h->stream>>>(h->d_X, h->n, d_x, d_t);
  h->launches++;
}
void dense_prod(fpb_handle* h, const double* d_v, double* d_y) {
  dim3 grid((uint32_t)((h->n + 255) / 256), h->dense_splits);
  fpb::k_dense_gemv_n<<<grid, 256, 0, h->stream>>>(h->d_X, h->n, (uint32_t)h->nsnps,
                                                    h->dense_cps, d_v, h->d_part);
  fpb::k_sum_splits<<<(uint32_t)((h->n + 255) / 256), 256, 0, h->stream>>>(
      h->d_part, h->dense_splits, h->n, d_y);
  h->launches += 2;
}

// t (nsnps) = X' x
void streaming_crossprod(fpb_handle* h, const double* d_x, double* d_t);
void streaming_prod(fpb_handle* h, const double* d_v, double* d_y);
void streaming_perform_op(fpb_handle* h, const double* d_x, double* d_y);

void launch_crossprod(fpb_handle* h, const double* d_x, double* d_t) {
  if (!h->kids.empty()) return streaming_crossprod(h, d_x, d_t);
  if (h->dense) dense_crossprod(h, d_x, d_t);
  else if (h->use_imma) imma_crossprod(h, d_x, d_t, false);
  else generic_crossprod(h, d_x, d_t);
}

// y (N) = X v
void launch_prod(fpb_handle* h, const double* d_v, double* d_y) {
  if (!h->kids.empty()) return streaming_prod(h, d_v, d_y);
  if (h->dense) {
    dense_prod(h, d_v, d_y);
  } else if (h->use_imma) {
    uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
    fpb::k_prod_inputs<<<gb, 256, 0, h->stream>>>(d_v, h->d_scale, (uint32_t)h->nsnps, h->d_a,
                                                  h->d_corr, h->d_pmax, h->d_psum);
    h->nparts = gb;
    h->launches++;
    imma_prod_tail(h, d_y);
  } else {
    generic_prod(h, d_v, d_y);
  }
}

// y (N) = X X' x
void launch_perform_op(fpb_handle* h, const double* d_x, double* d_y) {
  if (!h->kids.empty()) return streaming_perform_op(h, d_x, d_y);
  if (h->dense) {  // y = mat * (mat' x), svdwide.cpp:10
    dense_crossprod(h, d_x, h->d_t);
    dense_prod(h, h->d_t, d_y);
  } else if (h->use_imma && h->use_fused) {
    fused_perform_op(h, d_x, d_y);
  } else if (h->use_imma) {
    imma_crossprod(h, d_x, nullptr, true);
    imma_prod_tail(h, d_y);
  } else {
    generic_crossprod(h, d_x, h->d_t);
    generic_prod(h, h->d_t, d_y);
  }
}

int peer_allreduce(fpb_handle* h, double* d_buf, size_t count);

int allreduce(fpb_handle* h, double* d_buf, size_t count) {
  if (h->P.ok) return peer_allreduce(h, d_buf, count);
  if (!h->comm) return 0;
  int rc = g_nccl.AllReduce(d_buf, d_buf, count, kNcclFloat64, kNcclSum, h->comm, h->stream);
  if (rc != 0)
    FPB_FAIL(h, std::string("ncclAllReduce failed: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  return 0;
}

// Device -> pageable host copy through two pinned 16 MiB bounce buffers: the DMA of chunk c+1
// overlaps the host copy of chunk c, and the host copy itself runs on four threads (the destination
// is usually freshly allocated: the first touch of its pages, not the copy, is what costs -- a plain
// cudaMemcpy to pageable memory runs at ~4 GB/s, 20 ms for the 80 MB of eigenvectors at N = 500k).
int download_pageable(fpb_handle* h, void* dst, const void* d_src, size_t bytes) {
  constexpr size_t kChunk = 16u << 20;
  if (bytes <= (1u << 20)) {
    FPB_CUDA(h, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, h->stream));
    FPB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
  }
  for (int i = 0; i < 2; i++)
    if (!h->h_bounce[i]) {
      FPB_CUDA(h, cudaMallocHost(&h->h_bounce[i], kChunk));
      FPB_CUDA(h, cudaEventCreateWithFlags(&h->ev_bounce[i], cudaEventDisableTiming));
    }
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  auto issue = [&](size_t c) {
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    cudaMemcpyAsync(h->h_bounce[c & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost,
                    h->stream);
    cudaEventRecord(h->ev_bounce[c & 1], h->stream);
  };
  auto host_copy = [](char* to, const unsigned char* from, size_t len) {
    constexpr int kThreads = 4;
    const size_t per = ((len + kThreads - 1) / kThreads + 4095) & ~(size_t)4095;
    std::thread th[kThreads - 1];
    int started = 0;
    for (int t = 1; t < kThreads; t++) {
      const size_t o = (size_t)t * per;
      if (o >= len) break;
      th[started++] = std::thread([=] { memcpy(to + o, from + o, std::min(per, len - o)); });
    }
    memcpy(to, from, std::min(per, len));
    for (int t = 0; t < started; t++) th[t].join();
  };
  issue(0);
  for (size_t c = 0; c < nchunks; c++) {
    FPB_CUDA(h, cudaEventSynchronize(h->ev_bounce[c & 1]));
    if (c + 1 < nchunks) issue(c + 1);
    const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
    host_copy((char*)dst + off, h->h_bounce[c & 1], len);
  }
  FPB_CUDA(h, cudaGetLastError());
  return 0;
}

int check_launch(fpb_handle* h) {
  FPB_CUDA(h, cudaGetLastError());
  return 0;
}


#include "fpb_host_streaming.inl"

}  // namespace

extern "C" {

int fpb_abi_version(void) { return 1; }

const char* fpb_last_error(const fpb_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int fpb_create(fpb_handle** out, const unsigned char* bed_payload, uint64_t n, uint64_t nsnps,
               int stand_method, const double* preloaded_meansd, int device) {
  if (!out || !bed_payload) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  *out = nullptr;
  fpb_handle* h = new fpb_handle();
  if (alloc_common(h, n, nsnps, stand_method, device) ||
      [&]() -> int {
        FPB_CUDA(h, cudaMemcpy2DAsync(h->d_gs, h->pitch_s, bed_payload, h->np, h->np, nsnps,
                                      cudaMemcpyHostToDevice, h->stream));
        return 0;
      }() ||
      finish_create(h, preloaded_meansd)) {
    g_err = h->err;
    fpb_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

int fpb_create_from_file(fpb_handle** out, const char* bed_path, uint64_t n, uint64_t snp_begin,
                         uint64_t snp_count, int stand_method, const double* preloaded_meansd,
                         int device) {
  if (!out || !bed_path) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  *out = nullptr;
  if (n == 0) FPB_FAIL((fpb_handle*)nullptr, "empty genotype matrix (N == 0 or nsnps == 0)");
  FILE* f = fopen(bed_path, "rb");
  if (!f)
    FPB_FAIL((fpb_handle*)nullptr, std::string("[Data::read_bed] Error reading file ") + bed_path +
                                       ", error " + strerror(errno));
  // data.cpp:163-170: len = filesize - 3, np = ceil(N/4), nsnps = len / np
  fseeko(f, 0, SEEK_END);
  uint64_t fsz = (uint64_t)ftello(f);
  uint64_t len = fsz >= 3 ? fsz - 3 : 0;
  uint64_t np = (n + 3) / 4;
  uint64_t file_snps = len / np;
  if (snp_begin > file_snps) snp_begin = file_snps;
  if (snp_count == 0 || snp_begin + snp_count > file_snps) snp_count = file_snps - snp_begin;
  fpb_handle* h = new fpb_handle();
  int rc = alloc_common(h, n, snp_count, stand_method, device);
  if (!rc) {
    rc = [&]() -> int {
      // double-buffered pinned staging, 64 MiB slabs of whole SNP rows; a slab is read by several
      // threads at once (pread on disjoint row ranges: one thread copies ~3 GB/s out of the page
      // cache, the H2D copy it feeds runs at 55 GB/s), and the copy of slab b overlaps the read of
      // slab b + 1.  FPB_READ_THREADS overrides the thread count (default min(8, cores)).
      uint64_t rows_per_slab = std::max<uint64_t>(1, (64ull << 20) / np);
      rows_per_slab = std::min<uint64_t>(rows_per_slab, std::max<uint64_t>(1, snp_count));
      const char* rt = getenv("FPB_READ_THREADS");
      const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
      const unsigned nthr = rt && atoi(rt) > 0 ? (unsigned)atoi(rt) : std::min(8u, hw);
      const int fd = fileno(f);
      unsigned char* pin[2] = {nullptr, nullptr};
      cudaEvent_t ev[2];
      for (int b = 0; b < 2; b++) {
        FPB_CUDA(h, cudaMallocHost(&pin[b], rows_per_slab * np));
        FPB_CUDA(h, cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
      }
      const uint64_t file_off = 3 + np * snp_begin;
      auto read_range = [&](unsigned char* dst, uint64_t off, uint64_t len) -> bool {
        while (len) {
          const ssize_t got = pread(fd, dst, len, (off_t)off);
          if (got <= 0) {
            if (got < 0 && errno == EINTR) continue;
            return false;
          }
          dst += got;
          off += (uint64_t)got;
          len -= (uint64_t)got;
        }
        return true;
      };
      int b = 0;
      int failed = 0;
      for (uint64_t r = 0; r < snp_count && !failed; r += rows_per_slab, b ^= 1) {
        const uint64_t rows = std::min(rows_per_slab, snp_count - r);
        const uint64_t bytes = rows * np, off0 = file_off + r * np;
        cudaEventSynchronize(ev[b]);
        const unsigned nt = (unsigned)std::min<uint64_t>(nthr, std::max<uint64_t>(1, bytes >> 20));
        std::vector<char> okv(nt, 1);
        auto part = [&](unsigned i) {
          const uint64_t lo = bytes * i / nt, hi = bytes * (i + 1) / nt;
          okv[i] = read_range(pin[b] + lo, off0 + lo, hi - lo) ? 1 : 0;
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < nt; i++) pool.emplace_back(part, i);
        part(0);
        for (auto& th : pool) th.join();
        for (unsigned i = 0; i < nt; i++) failed |= !okv[i];
        if (failed) {
          h->err = std::string("[Data::read_bed] Error reading file ") + bed_path;
          break;
        }
        if (cudaMemcpy2DAsync(h->d_gs + r * h->pitch_s, h->pitch_s, pin[b], np, np, rows,
                              cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
          h->err = "CUDA error staging bed";
          failed = 1;
        }
        cudaEventRecord(ev[b], h->stream);
      }
      cudaStreamSynchronize(h->stream);
      for (int q = 0; q < 2; q++) {
        cudaFreeHost(pin[q]);
        cudaEventDestroy(ev[q]);
      }
      return failed;
    }();
  }
  fclose(f);
  if (!rc) rc = finish_create(h, preloaded_meansd);
  if (rc) {
    g_err = h->err;
    if (g_borrow_gs) h->d_gs = nullptr;  // the slab buffer belongs to the streaming parent
    fpb_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

#include "fpb_host_streaming_create.inl"

int fpb_create_synthetic(fpb_handle** out, uint64_t n, uint64_t nsnps, uint64_t snp_offset,
                         const unsigned char* pop_of_individual, const uint32_t* thresholds,
                         uint32_t npop, uint32_t missing_threshold, uint64_t seed,
                         int stand_method, int device) {
  if (!out || !pop_of_individual || !thresholds)
    FPB_FAIL((fpb_handle*)nullptr, "null argument");
  *out = nullptr;
  fpb_handle* h = new fpb_handle();
  int rc = alloc_common(h, n, nsnps, stand_method, device);
  if (!rc) {
    rc = [&]() -> int {
      uint8_t* d_pop = nullptr;
      uint32_t* d_thr = nullptr;
      FPB_CUDA(h, cudaMalloc(&d_pop, n));
      FPB_CUDA(h, cudaMalloc(&d_thr, sizeof(uint32_t) * (size_t)npop * nsnps));
      FPB_CUDA(h, cudaMemcpyAsync(d_pop, pop_of_individual, n, cudaMemcpyHostToDevice, h->stream));
      FPB_CUDA(h, cudaMemcpyAsync(d_thr, thresholds, sizeof(uint32_t) * (size_t)npop * nsnps,
                                  cudaMemcpyHostToDevice, h->stream));
      uint64_t total = nsnps * h->np;
      uint64_t blocks = (total + 255) / 256;
      if (blocks > 0x7FFFFFFFull) FPB_FAIL(h, "synthetic matrix too large for one launch");
      fpb::k_synth_bed<<<(uint32_t)blocks, 256, 0, h->stream>>>(
          h->d_gs, nsnps, n, h->pitch_s, snp_offset, d_pop, d_thr, missing_threshold, seed);
      h->launches++;
      FPB_CUDA(h, cudaStreamSynchronize(h->stream));
      cudaFree(d_pop);
      cudaFree(d_thr);
      return 0;
    }();
  }
  if (!rc) rc = finish_create(h, nullptr);
  if (rc) {
    g_err = h->err;
    fpb_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

int fpb_create_dense(fpb_handle** out, const double* x_host, uint64_t n, uint64_t p,
                     int stand_method, int device) {
  if (!out || !x_host) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  *out = nullptr;
  if (stand_method < 0 || stand_method > 4)
    FPB_FAIL((fpb_handle*)nullptr, "unknown standardization method");  // util.cpp:179
  if (n == 0 || p == 0 || p > 0xFFFFFFF0ull)
    FPB_FAIL((fpb_handle*)nullptr, "empty genotype matrix (N == 0 or nsnps == 0)");
  fpb_handle* h = new fpb_handle();
  int rc = [&]() -> int {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      FPB_FAIL(h, std::string("no usable CUDA device (flashpca_b200 has no CPU fallback): ") +
                      cudaGetErrorString(e));
    if (device < 0 || device >= ndev) FPB_FAIL(h, "invalid CUDA device ordinal");
    h->device = device;
    FPB_CUDA(h, cudaSetDevice(device));
    FPB_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->dense = true;
    h->n = n;
    h->nsnps = p;
    h->stand_method = stand_method;
    FPB_CUDA(h, cudaMalloc(&h->d_X, sizeof(double) * n * p));
    FPB_CUDA(h, cudaMalloc(&h->d_meansd, sizeof(double) * 2 * p));
    FPB_CUDA(h, cudaMalloc(&h->d_t, sizeof(double) * p));
    double* d_colsq = nullptr;
    FPB_CUDA(h, cudaMalloc(&d_colsq, sizeof(double) * p));
    FPB_CUDA(h, cudaMemcpyAsync(h->d_X, x_host, sizeof(double) * n * p, cudaMemcpyHostToDevice,
                                h->stream));
    fpb::k_dense_standardise<<<(uint32_t)p, 256, 0, h->stream>>>(h->d_X, n, (uint32_t)p,
                                                                 stand_method, h->d_meansd,
                                                                 d_colsq);
    h->launches++;
    std::vector<double> sq(p);
    FPB_CUDA(h, cudaMemcpyAsync(sq.data(), d_colsq, sizeof(double) * p, cudaMemcpyDeviceToHost,
                                h->stream));
    FPB_CUDA(h, cudaStreamSynchronize(h->stream));
    FPB_CUDA(h, cudaGetLastError());
    cudaFree(d_colsq);
    double tr = 0.0;
    for (double v : sq) tr += v;
    h->trace = tr;
    // column splits of the X t product so that small-N matrices still fill the GPU
    uint64_t rowblocks = (n + 255) / 256;
    uint64_t splits = std::max<uint64_t>(1, std::min<uint64_t>(p / 64 + 1, (592 + rowblocks - 1) / rowblocks));
    h->dense_cps = (uint32_t)((p + splits - 1) / splits);
    h->dense_splits = (uint32_t)((p + h->dense_cps - 1) / h->dense_cps);
    FPB_CUDA(h, cudaMalloc(&h->d_part, sizeof(double) * n * h->dense_splits));
    return 0;
  }();
  if (rc) {
    g_err = h->err;
    fpb_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

void fpb_destroy(fpb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->side) cudaStreamSynchronize(h->side);
  if (h->copy) cudaStreamSynchronize(h->copy);
  for (fpb_handle* kid : h->kids) {
    if (kid->borrowed) {  // streams and the slab buffer belong to this handle
      kid->stream = nullptr;
      kid->side = nullptr;
      kid->d_gs = nullptr;
    }
    if (kid->shared_scratch) {  // freed once, below, through this handle's own fields
      kid->d_slices = nullptr;
      kid->d_part = kid->d_a = kid->d_corr = kid->d_pmax = kid->d_psum = kid->d_mx = kid->d_mc = nullptr;
      kid->d_sc = nullptr;
    }
    fpb_destroy(kid);
  }
  for (size_t b = 0; b < h->kid_host.size(); b++)
    if (h->kid_host[b]) {
      if (b < h->kid_host_pinned.size() && !h->kid_host_pinned[b]) free(h->kid_host[b]);
      else cudaFreeHost(h->kid_host[b]);
    }
  for (int i = 0; i < 2; i++) {
    cudaFree(h->sbuf[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
  }
  cudaFree(h->d_ytmp);
  if (h->copy) cudaStreamDestroy(h->copy);
  delete h->solver;
  delete h->bsolver;
  for (int i = 0; i < 4; i++)
    if (h->kev[i]) cudaEventDestroy(h->kev[i]);
  peer_release(h);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  cudaFree(h->d_X);
  cudaFree(h->d_gs);
  cudaFree(h->d_gi);
  cudaFree(h->d_scale);
  cudaFree(h->d_rowptr_s);
  cudaFree(h->d_seg_s);
  cudaFree(h->d_seg_i);
  cudaFree(h->d_col16_s);
  cudaFree(h->d_col16_i);
  cudaFree(h->d_slices);
  cudaFree(h->d_part);
  cudaFree(h->d_a);
  cudaFree(h->d_corr);
  cudaFree(h->d_pmax);
  cudaFree(h->d_psum);
  cudaFree(h->d_sc);
  cudaFree(h->d_mx);
  cudaFree(h->d_mc);
  for (auto& kv : h->op_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  free_umma(h);
  cudaFree(h->L1.slices);
  cudaFree(h->L1.part);
  cudaFree(h->L1.a);
  cudaFree(h->L1.corr);
  cudaFree(h->L1.pmax);
  cudaFree(h->L1.psum);
  cudaFree(h->L1.sc);
  cudaFree(h->L1.mx);
  cudaFree(h->L1.mc);
  for (int i = 0; i < 2; i++) {
    if (h->h_bounce[i]) cudaFreeHost(h->h_bounce[i]);
    if (h->ev_bounce[i]) cudaEventDestroy(h->ev_bounce[i]);
  }
  cudaFree(h->d_fpart);
  cudaFree(h->d_ybuf);
  cudaFree(h->d_arep);
  cudaFree(h->d_fsync);
  cudaFree(h->d_fdbg);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  cudaFree(h->d_lut);
  cudaFree(h->d_meansd);
  cudaFree(h->d_t);
  cudaFree(h->d_coef);
  cudaFree(h->d_c0);
  cudaFree(h->d_gpart);
  cudaFree(h->d_in);
  cudaFree(h->d_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

uint64_t fpb_rows(const fpb_handle* h) { return h ? h->n : 0; }
uint64_t fpb_cols(const fpb_handle* h) { return h ? h->n : 0; }
uint64_t fpb_nsnps(const fpb_handle* h) { return h ? h->nsnps : 0; }
void* fpb_stream(const fpb_handle* h) { return h ? (void*)h->stream : nullptr; }
uint64_t fpb_launch_count(const fpb_handle* h) { return h ? h->launches : 0; }

int fpb_get_meansd(fpb_handle* h, double* out_meansd) {
  if (!h || !out_meansd) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  if (!h->kids.empty()) {  // concatenate the slabs' (mean, sd) columns
    std::vector<double> tmp;
    for (size_t b = 0; b < h->kids.size(); b++) {
      const uint64_t cnt = h->kids[b]->nsnps;
      tmp.resize(2 * cnt);
      if (fpb_get_meansd(h->kids[b], tmp.data())) FPB_FAIL(h, h->kids[b]->err);
      std::copy(tmp.begin(), tmp.begin() + cnt, out_meansd + h->kid_off[b]);
      std::copy(tmp.begin() + cnt, tmp.end(), out_meansd + h->nsnps + h->kid_off[b]);
    }
    return 0;
  }
  FPB_CUDA(h, cudaMemcpyAsync(out_meansd, h->d_meansd, sizeof(double) * 2 * h->nsnps,
                              cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int fpb_get_trace(fpb_handle* h, double* out_trace) {
  if (!h || !out_trace) FPB_FAIL(h, "null argument");
  *out_trace = h->trace;
  return 0;
}

int fpb_get_bed(fpb_handle* h, unsigned char* out_payload) {
  if (!h || !out_payload) FPB_FAIL(h, "null argument");
  if (h->dense) FPB_FAIL(h, "fpb_get_bed: the handle holds a dense matrix, not a bed");
  if (!h->kids.empty()) FPB_FAIL(h, "fpb_get_bed: not available in streaming mode");
  FPB_CUDA(h, cudaSetDevice(h->device));
  uint8_t* d_tmp = nullptr;
  uint64_t total = h->nsnps * h->np;
  FPB_CUDA(h, cudaMalloc(&d_tmp, total));
  fpb::k_decode_rows<<<(uint32_t)((total + 255) / 256), 256, 0, h->stream>>>(
      h->d_gs, d_tmp, h->nsnps, h->np, h->pitch_s);
  h->launches++;
  FPB_CUDA(h, cudaMemcpyAsync(out_payload, d_tmp, total, cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_tmp);
  return 0;
}

int fpb_sync(fpb_handle* h) {
  if (!h) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  FPB_CUDA(h, cudaGetLastError());
  return check_async(h);
}

// ------------------------------ device-pointer ops -------------------------

int fpb_crossprod_multi_dev(fpb_handle* h, const double* d_m, uint32_t k, double* d_y) {
  if (!h || !d_m || !d_y || k == 0) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  uint32_t c = 0;
  if (k >= 3 && umma_capable(h)) {
    if (ensure_umma(h)) return 1;
    for (; k - c >= 3; c += std::min(kUmmaLanes, k - c))
      umma_crossprod_block(h, d_m + (uint64_t)c * h->n, std::min(kUmmaLanes, k - c),
                           d_y + (uint64_t)c * h->nsnps, false);
  }
  if (k - c >= 2 && pair_capable(h)) {
    if (ensure_lane1(h)) return 1;
    for (; c + 1 < k; c += 2)
      imma_crossprod_pair(h, d_m + (uint64_t)c * h->n, d_m + (uint64_t)(c + 1) * h->n,
                          d_y + (uint64_t)c * h->nsnps, d_y + (uint64_t)(c + 1) * h->nsnps, false);
  }
  for (; c < k; c++)
    launch_crossprod(h, d_m + (uint64_t)c * h->n, d_y + (uint64_t)c * h->nsnps);
  return check_launch(h);
}

int fpb_prod_multi_dev(fpb_handle* h, const double* d_v, uint32_t k, double* d_y) {
  if (!h || !d_v || !d_y || k == 0) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  uint32_t c = 0;
  if (k >= 3 && umma_capable(h)) {
    if (ensure_umma(h)) return 1;
    for (; k - c >= 3; c += std::min(kUmmaLanes, k - c)) {
      umma_prod_inputs(h, d_v + (uint64_t)c * h->nsnps, std::min(kUmmaLanes, k - c));
      umma_prod_block(h, std::min(kUmmaLanes, k - c), d_y + (uint64_t)c * h->n);
    }
  }
  if (k - c >= 2 && pair_capable(h)) {
    if (ensure_lane1(h)) return 1;
    for (; c + 1 < k; c += 2) {
      prod_inputs_pair(h, d_v + (uint64_t)c * h->nsnps, d_v + (uint64_t)(c + 1) * h->nsnps);
      imma_prod_tail_pair(h, d_y + (uint64_t)c * h->n, d_y + (uint64_t)(c + 1) * h->n);
    }
  }
  // a single column may sum the shards inside its finalize kernel (fpb_peer.cuh)
  h->P.fuse = h->P.ok && k == 1;
  h->P.summed = false;
  for (; c < k; c++) launch_prod(h, d_v + (uint64_t)c * h->nsnps, d_y + (uint64_t)c * h->n);
  h->P.fuse = false;
  if (check_launch(h)) return 1;
  if (h->P.summed) return 0;
  return allreduce(h, d_y, (size_t)h->n * k);
}

int fpb_perform_op_multi_dev(fpb_handle* h, const double* d_m, uint32_t k, double* d_y) {
  if (!h || !d_m || !d_y || k == 0) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  uint32_t c = 0;
  if (k >= 3 && umma_capable(h)) {
    if (ensure_umma(h)) return 1;
    for (; k - c >= 3; c += std::min(kUmmaLanes, k - c)) {
      umma_crossprod_block(h, d_m + (uint64_t)c * h->n, std::min(kUmmaLanes, k - c), nullptr, true);
      umma_prod_block(h, std::min(kUmmaLanes, k - c), d_y + (uint64_t)c * h->n);
    }
  }
  if (k - c >= 2 && pair_capable(h) && !h->use_fused) {
    if (ensure_lane1(h)) return 1;
    for (; c + 1 < k; c += 2) {
      imma_crossprod_pair(h, d_m + (uint64_t)c * h->n, d_m + (uint64_t)(c + 1) * h->n, nullptr,
                          nullptr, true);
      imma_prod_tail_pair(h, d_y + (uint64_t)c * h->n, d_y + (uint64_t)(c + 1) * h->n);
    }
  }
  h->P.fuse = h->P.ok && k == 1;
  h->P.summed = false;
  for (; c < k; c++) launch_perform_op(h, d_m + (uint64_t)c * h->n, d_y + (uint64_t)c * h->n);
  h->P.fuse = false;
  if (check_launch(h)) return 1;
  if (h->P.summed) return 0;
  return allreduce(h, d_y, (size_t)h->n * k);
}

namespace {
bool graph_capable(const fpb_handle* h) {
  static const bool off = getenv("FPB_GRAPH") && atoi(getenv("FPB_GRAPH")) == 0;
  // (not with NCCL's all-reduce inside: a captured ncclAllReduce next to the eager collectives of
  // the host program's own communicator hung the 2-GPU run; the peer-memory shard sum is a plain
  // kernel and replays with the rest of the op)
  return !off && h->graphs_ok && (!h->comm || h->P.ok) && h->kids.empty() && !h->dense && h->use_imma &&
         !h->use_fused && !h->time_gemv;
}
}  // namespace

// y = X X' x, device pointers.  A (x, y) pointer pair that comes back a third time is captured once
// and replayed as a CUDA graph from then on.
int fpb_perform_op_dev(fpb_handle* h, const double* d_x, double* d_y) {
  if (!h || !d_x || !d_y) FPB_FAIL(h, "null argument");
  if (!graph_capable(h)) return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
  FPB_CUDA(h, cudaSetDevice(h->device));
  const uint64_t key = (uint64_t)(uintptr_t)d_x * 0x9E3779B97F4A7C15ull ^ (uint64_t)(uintptr_t)d_y;
  auto it = h->op_graphs.find(key);
  if (it == h->op_graphs.end()) {
    // pointer pairs that recur (the Lanczos loop's fixed buffers, the host API's staging buffers,
    // a benchmark loop) get a graph on their third call; one-off pairs never pay for a capture
    if (h->op_graphs.size() >= 64) return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
    h->op_graphs[key] = fpb_handle::OpGraph();
    return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
  }
  if (!it->second.exec && ++it->second.seen < 2) return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
  if (!it->second.exec) {
    const uint64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      h->graphs_ok = false;
      return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
    }
    const int rc = fpb_perform_op_multi_dev(h, d_x, 1, d_y);
    const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    cudaGraphExec_t exec = nullptr;
    if (rc || ce != cudaSuccess || !graph ||
        cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      h->graphs_ok = false;
      h->launches = l0;
      h->op_graphs.erase(key);
      return fpb_perform_op_multi_dev(h, d_x, 1, d_y);
    }
    cudaGraphDestroy(graph);
    it->second.exec = exec;
    it->second.launches = h->launches - l0;
    h->launches = l0;
  }
  FPB_CUDA(h, cudaGraphLaunch(it->second.exec, h->stream));
  h->launches += it->second.launches;
  if (h->P.ok) h->P.used = true;  // the replayed op ends in the peer kernel: look at its watchdog word
  return 0;
}

// ------------------------------ host-pointer ops ---------------------------

static int host_op(fpb_handle* h, const double* in, uint32_t k, double* out, uint64_t in_rows,
                   uint64_t out_rows, int (*dev_fn)(fpb_handle*, const double*, uint32_t, double*)) {
  if (!h || !in || !out || k == 0) FPB_FAIL(h, "null argument");
  if (in == out) FPB_FAIL(h, "input and output must not alias");
  FPB_CUDA(h, cudaSetDevice(h->device));
  if (ensure_staging(h, (size_t)in_rows * k, (size_t)out_rows * k)) return 1;
  static const bool slice_upload = !(getenv("FPB_SLICE_UPLOAD") && atoi(getenv("FPB_SLICE_UPLOAD")) == 0);
  if (h->P.ok && slice_upload && dev_fn == fpb_perform_op_multi_dev) {
    // perform_op's input is the same on every rank: upload a slice, gather the rest over NVLink
    if (peer_upload_replicated(h, in, h->d_in, (size_t)in_rows * k)) return 1;
  } else {
    FPB_CUDA(h, cudaMemcpyAsync(h->d_in, in, sizeof(double) * in_rows * k, cudaMemcpyHostToDevice,
                                h->stream));
  }
  if (k == 1 && dev_fn == fpb_perform_op_multi_dev) {  // graph-replayed single-vector op
    if (fpb_perform_op_dev(h, h->d_in, h->d_out)) return 1;
  } else if (dev_fn(h, h->d_in, k, h->d_out)) {
    return 1;
  }
  FPB_CUDA(h, cudaMemcpyAsync(out, h->d_out, sizeof(double) * out_rows * k,
                              cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return check_async(h);
}

int fpb_perform_op_multi(fpb_handle* h, const double* m_in, uint32_t k, double* y_out) {
  return host_op(h, m_in, k, y_out, h ? h->n : 0, h ? h->n : 0, fpb_perform_op_multi_dev);
}
int fpb_perform_op(fpb_handle* h, const double* x_in, double* y_out) {
  return fpb_perform_op_multi(h, x_in, 1, y_out);
}
int fpb_crossprod_multi(fpb_handle* h, const double* m_in, uint32_t k, double* y_out) {
  return host_op(h, m_in, k, y_out, h ? h->n : 0, h ? h->nsnps : 0, fpb_crossprod_multi_dev);
}
int fpb_crossprod(fpb_handle* h, const double* x_in, double* y_out) {
  return fpb_crossprod_multi(h, x_in, 1, y_out);
}
int fpb_prod_multi(fpb_handle* h, const double* v_in, uint32_t k, double* y_out) {
  return host_op(h, v_in, k, y_out, h ? h->nsnps : 0, h ? h->n : 0, fpb_prod_multi_dev);
}
int fpb_prod(fpb_handle* h, const double* v_in, double* y_out) {
  return fpb_prod_multi(h, v_in, 1, y_out);
}

// ------------------------------ multi-GPU ----------------------------------

int fpb_comm_unique_id(unsigned char id_out[128]) {
  if (!id_out) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  if (!g_nccl.load(g_err)) return 1;
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) FPB_FAIL((fpb_handle*)nullptr, "ncclGetUniqueId failed");
  memcpy(id_out, id.internal, 128);
  return 0;
}

int fpb_get_dense(fpb_handle* h, double* out_x) {
  if (!h || !out_x) FPB_FAIL(h, "null argument");
  if (!h->dense) FPB_FAIL(h, "fpb_get_dense: the handle holds a bed, not a dense matrix");
  FPB_CUDA(h, cudaSetDevice(h->device));
  FPB_CUDA(h, cudaMemcpyAsync(out_x, h->d_X, sizeof(double) * h->n * h->nsnps,
                              cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

static void drop_op_graphs(fpb_handle* h);

int fpb_comm_init(fpb_handle* h, const unsigned char id_in[128], int nranks, int rank) {
  if (!h || !id_in) FPB_FAIL(h, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) FPB_FAIL(h, "invalid rank / nranks");
  if (!g_nccl.load(g_err)) {
    h->err = g_err;
    return 1;
  }
  FPB_CUDA(h, cudaSetDevice(h->device));
  NcclUniqueId id;
  memcpy(id.internal, id_in, 128);
  int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (rc != 0)
    FPB_FAIL(h, std::string("ncclCommInitRank failed: ") +
                    (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  h->nranks = nranks;
  h->rank = rank;
  drop_op_graphs(h);
  return peer_setup_ipc(h);
}

// graphs captured before the communicator existed do not contain the shard sum
static void drop_op_graphs(fpb_handle* h) {
  for (auto& kv : h->op_graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->op_graphs.clear();
}

// 0 = single shard, 1 = ncclAllReduce, 2 = peer-memory kernel (fpb_peer.cuh)
int fpb_comm_kind(const fpb_handle* h) {
  if (!h) return 0;
  return h->P.ok ? 2 : h->comm ? 1 : 0;
}

// Shards held by handles of this process (one GPU or several with peer access): handle i becomes
// rank i of n; the shard sum runs over plain device pointers.  Calls on the n handles must be
// issued concurrently (one host thread per handle) or at least all enqueued before any is waited
// for -- each rank's kernel waits for the others.
int fpb_comm_link_local(fpb_handle** hs, int n) {
  if (!hs || n < 1 || n > fpb::kPeerMax) FPB_FAIL((fpb_handle*)nullptr, "invalid handle list");
  for (int i = 0; i < n; i++) {
    if (!hs[i]) FPB_FAIL((fpb_handle*)nullptr, "null handle");
    if (hs[i]->comm || hs[i]->P.ok) FPB_FAIL(hs[i], "handle already has a communicator");
    if (hs[i]->n != hs[0]->n) FPB_FAIL(hs[i], "shards must have the same number of individuals");
  }
  bool shared_device = false;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++) shared_device |= hs[i]->device == hs[j]->device;
  if (shared_device && n > 4) FPB_FAIL(hs[0], "at most 4 linked shards may share a GPU");
  for (int i = 0; i < n; i++) {
    FPB_CUDA(hs[i], cudaSetDevice(hs[i]->device));
    for (int j = 0; j < n; j++)
      if (hs[j]->device != hs[i]->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, hs[i]->device, hs[j]->device);
        if (!can) FPB_FAIL(hs[i], "no peer access between the devices of the shards");
        cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          FPB_FAIL(hs[i], "cudaDeviceEnablePeerAccess failed");
        cudaGetLastError();
      }
    if (peer_alloc(hs[i])) return 1;
  }
  for (int i = 0; i < n; i++) {
    fpb_handle::Peer& P = hs[i]->P;
    for (int g = 0; g < n; g++)
      peer_slot(P.view, g, static_cast<unsigned char*>(hs[g]->P.region), P.cap);
    P.view.rank = i;
    P.view.world = n;
    P.grid = shared_device ? fpb::kPeerCtasShared : fpb::kPeerCtas;
    P.ok = true;
    hs[i]->nranks = n;
    hs[i]->rank = i;
    drop_op_graphs(hs[i]);
  }
  return 0;
}

// ------------------------------ whole solve --------------------------------

int fpb_pca(fpb_handle* h, uint32_t nev, uint32_t ncv, uint32_t maxiter, double tol,
            double* evals_out, double* evecs_out, uint32_t* nconv_out, uint32_t* nops_out,
            uint32_t* niter_out) {
  if (!h) FPB_FAIL(h, "null argument");
  if (nev < 1 || ncv <= nev || ncv > h->n)
    FPB_FAIL(h, "invalid nev/ncv (need 1 <= nev < ncv <= N)");
  FPB_CUDA(h, cudaSetDevice(h->device));
  h->op_ms.clear();
  std::vector<cudaEvent_t> evs;
  int op_rc = 0;
  auto op = [&](const double* d_in, double* d_out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, h->stream);
    op_rc |= fpb_perform_op_dev(h, d_in, d_out);
    cudaEventRecord(b, h->stream);
    evs.push_back(a);
    evs.push_back(b);
  };
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t0 = now();
  if (h->solver && !h->solver->matches(h->n, nev, ncv)) {
    delete h->solver;
    h->solver = nullptr;
  }
  if (!h->solver) h->solver = new fpb::Irlm(h->n, nev, ncv, h->stream, op);
  else h->solver->set_op(op);
  fpb::Irlm& solver = *h->solver;
  fpb::IrlmResult res;
  solver.run(maxiter, tol, res);
  cudaStreamSynchronize(h->stream);
  const auto t1 = now();
  for (size_t i = 0; i + 1 < evs.size(); i += 2) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
    h->op_ms.push_back(ms);
    cudaEventDestroy(evs[i]);
    cudaEventDestroy(evs[i + 1]);
  }
  if (op_rc || check_async(h)) {
    solver.set_op(nullptr);
    return 1;
  }
  if (!solver.error.empty()) {
    solver.set_op(nullptr);
    FPB_FAIL(h, solver.error);
  }
  FPB_CUDA(h, cudaGetLastError());
  if (evals_out) memcpy(evals_out, res.evals.data(), sizeof(double) * nev);
  h->last_evals = res.evals;
  h->last_solve_block = false;
  auto t2 = t1, t3 = t1;
  if (evecs_out) {
    if (ensure_staging(h, 0, (size_t)h->n * nev)) return 1;
    solver.eigenvectors(h->d_out);
    t2 = now();
    FPB_CUDA(h, cudaStreamSynchronize(h->stream));
    t2 = now();
    if (download_pageable(h, evecs_out, h->d_out, sizeof(double) * h->n * nev)) return 1;
    t3 = now();
  }
  h->pca_phase_s[0] = secs(t0, t1);
  h->pca_phase_s[1] = secs(t1, t2);
  h->pca_phase_s[2] = secs(t2, t3);
  h->pca_phase_s[3] = secs(t0, t3);
  solver.set_op(nullptr);
  if (nconv_out) *nconv_out = res.nconv;
  if (nops_out) *nops_out = res.nops;
  if (niter_out) *niter_out = res.niter;
  FPB_CUDA(h, cudaGetLastError());
  return 0;
}

int fpb_pca_block(fpb_handle* h, uint32_t nev, uint32_t block, uint32_t max_passes, double tol,
                  double* evals_out, double* evecs_out, uint32_t* nconv_out, uint32_t* npasses_out) {
  if (!h) FPB_FAIL(h, "null argument");
  if (block == 0) block = fpb::kBlkMaxB;
  if (max_passes == 0) max_passes = 40;
  if (nev < 1 || block > (uint32_t)fpb::kBlkMaxB || (uint64_t)nev + block > h->n)
    FPB_FAIL(h, "invalid nev/block (need 1 <= nev, block <= 8, nev + block <= N)");
  max_passes = (uint32_t)std::min<uint64_t>(max_passes, h->n / block);
  if ((uint64_t)block * max_passes < nev) FPB_FAIL(h, "max_passes too small for nev");
  FPB_CUDA(h, cudaSetDevice(h->device));
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  const auto t0 = now();
  auto op = [h](const double* d_in, uint32_t k, double* d_out) -> int {
    return fpb_perform_op_multi_dev(h, d_in, k, d_out);
  };
  if (h->bsolver && !h->bsolver->matches(h->n, nev, block, max_passes)) {
    delete h->bsolver;
    h->bsolver = nullptr;
  }
  if (!h->bsolver) h->bsolver = new fpb::BlockKrylov(h->n, nev, block, max_passes, h->stream, op);
  else h->bsolver->set_op(op);
  fpb::BlockResult res;
  const int rc = h->bsolver->run(tol, res);
  const auto t1 = now();
  if (getenv("FPB_IRLM_TRACE"))
    fprintf(stderr, "[block] %u passes: operator + V'AV %.1f ms, orthogonalisation %.1f ms, Rayleigh-Ritz %.1f ms\n",
            res.npasses, h->bsolver->t_op * 1e3, h->bsolver->t_orth * 1e3, h->bsolver->t_ritz * 1e3);
  if (rc || check_async(h)) {
    if (rc) FPB_FAIL(h, h->bsolver->error.empty() ? std::string("block Krylov solve failed") : h->bsolver->error);
    return 1;
  }
  FPB_CUDA(h, cudaGetLastError());
  if (evals_out) memcpy(evals_out, res.evals.data(), sizeof(double) * nev);
  h->last_evals = res.evals;
  h->last_solve_block = true;
  auto t2 = t1;
  if (evecs_out) {
    if (download_pageable(h, evecs_out, h->bsolver->eigenvectors(), sizeof(double) * h->n * nev)) return 1;
    t2 = now();
  }
  h->pca_phase_s[0] = secs(t0, t1);
  h->pca_phase_s[1] = 0.0;
  h->pca_phase_s[2] = secs(t1, t2);
  h->pca_phase_s[3] = secs(t0, t2);
  if (nconv_out) *nconv_out = res.nconv;
  if (npasses_out) *npasses_out = res.npasses;
  return 0;
}

int fpb_pca_residual(fpb_handle* h, double div, double* err_out, uint32_t nev) {
  if (!h || !err_out || !(div > 0.0)) FPB_FAIL(h, "null argument");
  if ((h->last_solve_block ? !h->bsolver : !h->solver) || h->last_evals.size() != nev || nev == 0)
    FPB_FAIL(h, "fpb_pca_residual: no fpb_pca result with this many eigenvectors on the handle");
  FPB_CUDA(h, cudaSetDevice(h->device));
  if (ensure_staging(h, (size_t)h->n * nev + nev, (size_t)h->n * nev + nev)) return 1;
  if (h->last_solve_block)
    FPB_CUDA(h, cudaMemcpyAsync(h->d_in, h->bsolver->eigenvectors(), sizeof(double) * h->n * nev,
                                cudaMemcpyDeviceToDevice, h->stream));
  else
    h->solver->eigenvectors(h->d_in);  // U, N x nev
  if (fpb_perform_op_multi_dev(h, h->d_in, nev, h->d_out)) return 1;  // all-reduced when sharded
  double* d_lam = h->d_in + (size_t)h->n * nev;
  double* d_err = h->d_out + (size_t)h->n * nev;
  FPB_CUDA(h, cudaMemcpyAsync(d_lam, h->last_evals.data(), sizeof(double) * nev,
                              cudaMemcpyHostToDevice, h->stream));
  fpb::k_check_resid<<<nev, 1024, 0, h->stream>>>(h->d_out, h->d_in, d_lam, h->n, div, d_err);
  h->launches++;
  FPB_CUDA(h, cudaMemcpyAsync(err_out, d_err, sizeof(double) * nev, cudaMemcpyDeviceToHost,
                              h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  FPB_CUDA(h, cudaGetLastError());
  return check_async(h);
}

void fpb_pca_phase_times(const fpb_handle* h, double out_seconds[4]) {
  if (!h || !out_seconds) return;
  for (int i = 0; i < 4; i++) out_seconds[i] = h->pca_phase_s[i];
}

uint32_t fpb_pca_op_times(const fpb_handle* h, float* ms_out, uint32_t cap) {
  if (!h || !ms_out) return 0;
  uint32_t m = (uint32_t)std::min<size_t>(cap, h->op_ms.size());
  for (uint32_t i = 0; i < m; i++) ms_out[i] = h->op_ms[i];
  return m;
}

int fpb_time_perform_op(fpb_handle* h, const double* d_x, double* d_y, uint32_t reps,
                        float* ms_per_op_out, float* ms_kernels_out) {
  if (!h || !d_x || !d_y || reps == 0) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  cudaEvent_t e0, e1, e2, e3;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
  FPB_CUDA(h, cudaEventRecord(e0, h->stream));
  for (uint32_t r = 0; r < reps; r++)
    if (fpb_perform_op_dev(h, d_x, d_y)) return 1;
  FPB_CUDA(h, cudaEventRecord(e1, h->stream));
  FPB_CUDA(h, cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_per_op_out) *ms_per_op_out = ms / reps;
  if (ms_kernels_out) {
    // one more op with events around each half, and around each contraction kernel
    for (int i = 0; i < 4; i++)
      if (!h->kev[i]) cudaEventCreate(&h->kev[i]);
    h->time_gemv = h->use_imma && h->kids.empty();
    const bool fused = h->use_imma && h->use_fused && !h->dense;
    cudaEventRecord(e0, h->stream);
    if (fused) {
      // one launch does both halves: [0] = the whole op, [2] = the fused kernel alone
      launch_perform_op(h, d_x, d_y);
      cudaEventRecord(e1, h->stream);
      cudaEventRecord(e2, h->stream);
    } else {
      launch_crossprod(h, d_x, h->d_t);
      cudaEventRecord(e1, h->stream);
      cudaEventRecord(e2, h->stream);
      launch_prod(h, h->d_t, d_y);
    }
    cudaEventRecord(e3, h->stream);
    h->time_gemv = false;
    FPB_CUDA(h, cudaEventSynchronize(e3));
    cudaEventElapsedTime(&ms_kernels_out[0], e0, e1);
    cudaEventElapsedTime(&ms_kernels_out[1], e2, e3);
    ms_kernels_out[2] = ms_kernels_out[3] = 0.f;
    if (fused) {
      ms_kernels_out[1] = 0.f;
      cudaEventElapsedTime(&ms_kernels_out[2], h->kev[0], h->kev[1]);
    } else if (h->use_imma) {
      cudaEventElapsedTime(&ms_kernels_out[2], h->kev[0], h->kev[1]);
      cudaEventElapsedTime(&ms_kernels_out[3], h->kev[2], h->kev[3]);
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
  FPB_CUDA(h, cudaGetLastError());
  return check_async(h);
}

int fpb_time_perform_op_steps(fpb_handle* h, const double* d_x, double* d_y, uint32_t reps,
                              float* ms_each_out) {
  if (!h || !d_x || !d_y || !ms_each_out || reps == 0) FPB_FAIL(h, "null argument");
  FPB_CUDA(h, cudaSetDevice(h->device));
  std::vector<cudaEvent_t> ev(reps + 1);
  for (auto& e : ev) FPB_CUDA(h, cudaEventCreate(&e));
  FPB_CUDA(h, cudaEventRecord(ev[0], h->stream));
  for (uint32_t r = 0; r < reps; r++) {
    if (fpb_perform_op_dev(h, d_x, d_y)) return 1;
    FPB_CUDA(h, cudaEventRecord(ev[r + 1], h->stream));
  }
  FPB_CUDA(h, cudaEventSynchronize(ev[reps]));
  for (uint32_t r = 0; r < reps; r++) cudaEventElapsedTime(&ms_each_out[r], ev[r], ev[r + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  FPB_CUDA(h, cudaGetLastError());
  return check_async(h);
}

int fpb_fused_debug(fpb_handle* h, unsigned long long* out, uint64_t count) {
  if (!h || !out) FPB_FAIL(h, "null argument");
  const uint64_t have = 2ull * fpb::kFDbgSlabs * fpb::kFDbgEvents;
  if (!h->d_fdbg || count < have) FPB_FAIL(h, "no fused debug stamps (set FPB_FUSED_DEBUG=1)");
  FPB_CUDA(h, cudaMemcpy(out, h->d_fdbg, sizeof(unsigned long long) * have,
                         cudaMemcpyDeviceToHost));
  return 0;
}

int fpb_device_memory(int device, uint64_t* free_bytes, uint64_t* total_bytes) {
  if (!free_bytes || !total_bytes) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    FPB_FAIL((fpb_handle*)nullptr,
             std::string("no usable CUDA device (flashpca_b200 has no CPU fallback): ") +
                 cudaGetErrorString(e));
  if (device < 0 || device >= ndev) FPB_FAIL((fpb_handle*)nullptr, "invalid CUDA device ordinal");
  FPB_CUDA((fpb_handle*)nullptr, cudaSetDevice(device));
  size_t fr = 0, tot = 0;
  FPB_CUDA((fpb_handle*)nullptr, cudaMemGetInfo(&fr, &tot));
  *free_bytes = fr;
  *total_bytes = tot;
  return 0;
}

unsigned fpb_path_info(const fpb_handle* h) {
  if (!h) return 0;
  if (!h->kids.empty()) return FPB_PATH_STREAMING | fpb_path_info(h->kids[0]);
  return (h->dense ? FPB_PATH_DENSE : 0u) | (h->use_imma ? FPB_PATH_TENSOR : 0u) |
         (h->use_imma && h->use_tma ? FPB_PATH_TMA : 0u) |
         (h->use_imma && h->single_copy ? FPB_PATH_SINGLE_COPY : 0u) |
         (h->use_imma && h->use_fused ? FPB_PATH_FUSED : 0u);
}

}  // extern "C"
