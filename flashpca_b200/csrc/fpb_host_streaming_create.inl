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
      // the slab is staged (read, recoded, statistics, missing-genotype lists) inside slab buffer 0:
      // no third slab-sized allocation next to the two buffers
      g_borrow_gs = h->sbuf[0];
      const int crc = fpb_create_from_file(&kid, bed_path, n, snp_begin + off, cnt, stand_method,
                                           kid_msd, device);
      g_borrow_gs = nullptr;
      if (crc) FPB_FAIL(h, g_err);
      h->kids.push_back(kid);
      h->kid_off.push_back(off);
      h->kid_host.push_back(nullptr);
      h->kid_host_pinned.push_back(1);
      kid->borrowed = true;  // from here on fpb_destroy(kid) must not free the slab buffer
      // the recoded genotypes leave HBM: host memory is their home from now on.  Pinned when the
      // host allows it (the copy engine then overlaps the kernels); a bed that cannot be pinned
      // (larger than the lockable memory) falls back to pageable memory, slab by slab.
      if (cudaMallocHost(&h->kid_host.back(), kid->pitch_s * cnt) != cudaSuccess) {
        cudaGetLastError();
        h->kid_host.back() = (uint8_t*)malloc(kid->pitch_s * cnt);
        h->kid_host_pinned.back() = 0;
        if (!h->kid_host.back()) FPB_FAIL(h, "out of host memory for the streamed genotypes");
      }
      FPB_CUDA(h, cudaMemcpy(h->kid_host.back(), kid->d_gs, kid->pitch_s * cnt,
                             cudaMemcpyDeviceToHost));
      kid->d_gs = nullptr;
      // per-op scratch: one set for all slabs (they run one after the other on one stream), sized
      // for the largest; freed here so that staging the next slab does not stack allocations
      if (kid->use_imma) {
        cudaFree(kid->d_slices); cudaFree(kid->d_part); cudaFree(kid->d_a); cudaFree(kid->d_corr);
        cudaFree(kid->d_pmax); cudaFree(kid->d_psum); cudaFree(kid->d_sc); cudaFree(kid->d_mx);
        cudaFree(kid->d_mc);
        kid->d_slices = nullptr;
        kid->d_part = kid->d_a = kid->d_corr = kid->d_pmax = kid->d_psum = kid->d_mx = kid->d_mc = nullptr;
        kid->d_sc = nullptr;
        kid->shared_scratch = true;
      }
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
      h->trace += kid->trace;  // slab order, like the block loop of svdwide.cpp:44-61
      h->launches += kid->launches;
      kid->launches = 0;
    }
    // the shared per-op scratch, held in this (parent) handle's own fields
    size_t slice_b = 0, part_e = 0, snps_e = 0, parts_e = 0, mx_e = 0, mc_e = 0;
    for (fpb_handle* kid : h->kids)
      if (kid->shared_scratch) {
        slice_b = std::max(slice_b, kid->slice_bytes);
        part_e = std::max(part_e, kid->part_elems);
        snps_e = std::max<size_t>(snps_e, kid->nsnps);
        parts_e = std::max(parts_e, kid->max_parts);
        if (kid->nmissing) {
          mx_e = std::max<size_t>(mx_e, (size_t)kid->nsnps * kid->gtiles_s);
          mc_e = std::max<size_t>(mc_e, (size_t)kid->n * kid->gtiles_i);
        }
      }
    if (slice_b) {
      FPB_CUDA(h, cudaMalloc(&h->d_slices, slice_b));
      FPB_CUDA(h, cudaMalloc(&h->d_part, sizeof(double) * part_e));
      FPB_CUDA(h, cudaMalloc(&h->d_a, sizeof(double) * snps_e));
      FPB_CUDA(h, cudaMalloc(&h->d_corr, sizeof(double) * snps_e));
      FPB_CUDA(h, cudaMalloc(&h->d_pmax, sizeof(double) * parts_e));
      FPB_CUDA(h, cudaMalloc(&h->d_psum, sizeof(double) * parts_e));
      FPB_CUDA(h, cudaMalloc(&h->d_sc, sizeof(fpb::VecScale) * 2));
      if (mx_e) FPB_CUDA(h, cudaMalloc(&h->d_mx, sizeof(double) * mx_e));
      if (mc_e) FPB_CUDA(h, cudaMalloc(&h->d_mc, sizeof(double) * mc_e));
      for (fpb_handle* kid : h->kids)
        if (kid->shared_scratch) {
          kid->d_slices = h->d_slices;
          kid->d_part = h->d_part;
          kid->d_a = h->d_a;
          kid->d_corr = h->d_corr;
          kid->d_pmax = h->d_pmax;
          kid->d_psum = h->d_psum;
          kid->d_sc = h->d_sc;
          kid->d_mx = h->d_mx;
          kid->d_mc = h->d_mc;
        }
    }
    h->sbuf_holds[0] = (long long)h->kids.size() - 1;  // the last slab staged is still in buffer 0 ...
    if (((h->kids.size() - 1) & 1) != 0) h->sbuf_holds[0] = -1;  // ... usable only if that is its buffer
    return 0;
  }();
  g_borrow_gs = nullptr;
  if (rc) {
    g_err = h->err;
    fpb_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}
