// Host side of the fused single-pass kernel: staging-time setup.  Textually included by fpb_capi.cu
// (inside its anonymous namespace).
// Fused single-pass perform_op (fpb_fused.cuh): one persistent CTA per SM owns
// <= kFSpc column stripes.  Opt-in (FPB_FUSED=1): it reads HBM once per op (ncu:
// 13.2 GB vs 25.1 GB) but on B200 the two halves do not overlap on an SM -- the
// kernel is bound by instruction issue (mma.sync + LOP3 decode), 4.2 ms against
// 3.7 ms for the two HBM-bound kernels (profiles/r01_fused_*).
int setup_fused(fpb_handle* h) {
  h->f_nstripes = (uint32_t)((h->pitch_s + 127) / 128);
  h->f_grid = std::min<uint32_t>((uint32_t)h->sm_count, h->f_nstripes);
  const uint32_t spc = (h->f_nstripes + h->f_grid - 1) / h->f_grid;
  const bool feasible = spc <= (uint32_t)fpb::kFSpc;
  bool want = false;
  if (const char* fv = getenv("FPB_FUSED")) want = feasible && atoi(fv) != 0;
  if (want) {
    int coop = 0;
    FPB_CUDA(h, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    if (!coop) want = false;
  }
  h->use_fused = want;
  if (!want) return 0;
  h->f_nslabs = (uint32_t)((h->nsnps + fpb::kFRows - 1) / fpb::kFRows);
  h->f_gpad = (fpb::kFP1Groups * h->f_grid + 31) / 32 * 32;
  if (const char* wv = getenv("FPB_FUSED_WINDOW"))
    h->f_window = (uint32_t)std::min(std::max(atoi(wv), 2), fpb::kFASlots);  // >= 2: lagged slot release
  if (const char* pv = getenv("FPB_FUSED_POL1")) h->f_pol1 = (uint32_t)atoi(pv);
  if (const char* pv = getenv("FPB_FUSED_POL2")) h->f_pol2 = (uint32_t)atoi(pv);
  if (const char* pv = getenv("FPB_FUSED_PREFETCH")) h->f_prefetch = (uint32_t)std::max(0, atoi(pv));
  FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_fused_op, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpb::kFSmemBytes));
  const size_t part_bytes = sizeof(double) * (size_t)fpb::kFASlots * fpb::kFRows * h->f_gpad;
  FPB_CUDA(h, cudaMalloc(&h->d_fpart, part_bytes));
  FPB_CUDA(h, cudaMemsetAsync(h->d_fpart, 0xFF, part_bytes, h->stream));  // all-ones = free
  FPB_CUDA(h, cudaMalloc(&h->d_ybuf, sizeof(double) * h->n));
  h->f_rep_stride = (uint64_t)h->f_nslabs * fpb::kFRows;
  FPB_CUDA(h, cudaMalloc(&h->d_arep, sizeof(double) * fpb::kFReplicas * h->f_rep_stride));
  FPB_CUDA(h, cudaMalloc(&h->d_fsync, sizeof(uint32_t)));
  FPB_CUDA(h, cudaMemsetAsync(h->d_fsync, 0, sizeof(uint32_t), h->stream));
  if (getenv("FPB_FUSED_DEBUG")) {
    const size_t db = sizeof(unsigned long long) * 2 * fpb::kFDbgSlabs * fpb::kFDbgEvents;
    FPB_CUDA(h, cudaMalloc(&h->d_fdbg, db));
    FPB_CUDA(h, cudaMemsetAsync(h->d_fdbg, 0, db, h->stream));
  }
  return 0;
}
