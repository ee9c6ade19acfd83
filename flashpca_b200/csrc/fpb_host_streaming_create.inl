// Out-of-HBM streaming: fpb_create_streaming.  Textually included by fpb_capi.cu (inside extern "C").
int fpb_create_streaming(fpb_handle** out, const char* bed_path, uint64_t n, uint64_t snp_begin,
                         uint64_t snp_count, uint64_t snps_per_slab, int stand_method,
                         const double* preloaded_meansd, int device) {
  if (!out || !bed_path) FPB_FAIL((fpb_handle*)nullptr, "null argument");
  *out = nullptr;
  if (n == 0) FPB_FAIL((fpb_handle*)nullptr, "empty genotype matrix (N == 0 or nsnps == 0)");
  if (snps_per_slab == 0) FPB_FAIL((fpb_handle*)nullptr, "snps_per_slab must be positive");
  {
    const char* gv = getenv("FPB_GEMV");
    if (gv && (!strcmp(gv, "ldg") || !strcmp(gv, "tma2")))
      FPB_FAIL((fpb_handle*)nullptr, "streaming mode needs the single-copy kernels (unset FPB_GEMV)");
  }
  FILE* f = fopen(bed_path, "rb");
  if (!f)
    FPB_FAIL((fpb_handle*)nullptr, std::string("[Data::read_bed] Error reading file ") + bed_path +
                                       ", error " + strerror(errno));
  fseeko(f, 0, SEEK_END);
  const uint64_t fsz = (uint64_t)ftello(f);
  fclose(f);
  const uint64_t np = (n + 3) / 4, file_snps = (fsz >= 3 ? fsz - 3 : 0) / np;  // data.cpp:163-170
  if (snp_begin > file_snps) snp_begin = file_snps;
  if (snp_count == 0 || snp_begin + snp_count > file_snps) snp_count = file_snps - snp_begin;
  if (snp_count == 0) FPB_FAIL((fpb_handle*)nullptr, "empty genotype matrix (N == 0 or nsnps == 0)");

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
    FPB_CUDA(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    FPB_CUDA(h, cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      FPB_CUDA(h, cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      FPB_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    h->n = n;
    h->nsnps = snp_count;
    h->np = np;
    h->pitch_s = (np + 63) / 64 * 64;
    h->stand_method = stand_method;
    const uint64_t slab = std::min(snps_per_slab, snp_count);
    for (int i = 0; i < 2; i++) FPB_CUDA(h, cudaMalloc(&h->sbuf[i], h->pitch_s * slab));
    FPB_CUDA(h, cudaMalloc(&h->d_ytmp, sizeof(double) * n));
    FPB_CUDA(h, cudaMalloc(&h->d_t, sizeof(double) * snp_count));  // fpb_time_perform_op's X'x scratch
    std::vector<double> msd;
    for (uint64_t off = 0; off < snp_count; off += slab) {
      const uint64_t cnt = std::min(slab, snp_count - off);
      const double* kid_msd = nullptr;
      if (preloaded_meansd) {  // nsnps x 2 column-major -> the slab's cnt x 2
        msd.resize(2 * cnt);
        std::copy(preloaded_meansd + off, preloaded_meansd + off + cnt, msd.begin());
        std::copy(preloaded_meansd + snp_count + off, preloaded_meansd + snp_count + off + cnt,
                  msd.begin() + cnt);
        kid_msd = msd.data();
      }
      fpb_handle* kid = nullptr;
      if (fpb_create_from_file(&kid, bed_path, n, snp_begin + off, cnt, stand_method, kid_msd,
                               device))
        FPB_FAIL(h, g_err);
      h->kids.push_back(kid);
      h->kid_off.push_back(off);
      h->kid_host.push_back(nullptr);
      // the recoded genotypes leave HBM: pinned host memory is their home from now on
      FPB_CUDA(h, cudaMallocHost(&h->kid_host.back(), kid->pitch_s * cnt));
      FPB_CUDA(h, cudaMemcpy(h->kid_host.back(), kid->d_gs, kid->pitch_s * cnt,
                             cudaMemcpyDeviceToHost));
      cudaFree(kid->d_gs);
      kid->d_gs = nullptr;
      if (kid->use_imma && kid->use_tma)
        for (int i = 0; i < 2; i++) {
          if (make_tensor_map(kid, h->sbuf[i], kid->pitch_s, cnt, &kid->tm_s_alt[i]) ||
              make_tensor_map(kid, h->sbuf[i], kid->pitch_s, cnt, &kid->tm_f_alt[i], fpb::kFRows))
            FPB_FAIL(h, kid->err);
        }
      cudaStreamDestroy(kid->stream);
      cudaStreamDestroy(kid->side);
      kid->stream = h->stream;
      kid->side = h->side;
      kid->borrowed = true;
      h->trace += kid->trace;  // slab order, like the block loop of svdwide.cpp:44-61
      h->launches += kid->launches;
      kid->launches = 0;
    }
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
