// Block variants on the tcgen05 path (fpb_umma.cuh): up to 8 vectors per pass over the packed
// matrix.  Textually included by fpb_capi.cu (inside its anonymous namespace).
// perform_op_mat / perform_op_multi, crossprod2, prod3 (svdwide.cpp:71-118, 157-188, 229-275,
// 312-343) with k >= 3 columns take their columns 8 (or 4) at a time: one k_umma_xt / k_umma_xv
// launch contracts all of them (N = 8 digit slices x vectors), the per-vector small kernels
// (missing-genotype gathers, scale, slicing, finalize) run once per lane on lane-indexed scratch.

constexpr uint32_t kUmmaLanes = 8;

bool umma_capable(const fpb_handle* h) {
  static const bool off = getenv("FPB_UMMA") && atoi(getenv("FPB_UMMA")) == 0;
  return !off && h->kids.empty() && !h->dense && h->use_imma && h->use_tma && h->single_copy;
}

int ensure_umma(fpb_handle* h) {
  fpb_handle::Umma& U = h->U;
  if (U.ready) return 0;
  const uint32_t sm = (uint32_t)h->sm_count;
  // first half: CTA = 512 SNP rows x a split of the 128-byte column stages.  At most 128 stages
  // (65536 individuals) per split: 65536 x 255 x 127 < 2^31 bounds the int32 accumulators.
  U.nst = (uint32_t)((h->pitch_s + 127) / 128);
  const uint32_t tiles_t = (uint32_t)((h->nsnps + 511) / 512);
  uint32_t sp = std::max<uint32_t>((U.nst + 127) / 128, (4 * sm + tiles_t - 1) / tiles_t);
  sp = std::max<uint32_t>(std::min<uint32_t>(sp, std::max<uint32_t>((U.nst + 127) / 128, U.nst / 8)), 1);
  U.sps_t = (U.nst + sp - 1) / sp;
  U.splits_t = (U.nst + U.sps_t - 1) / U.sps_t;
  // second half: CTA = one 128-byte stripe x a split of the 128-row boxes; at most 1024 boxes
  // (131072 SNPs) per split: 131072 x 255 x 64 < 2^31
  U.nbox = (uint32_t)((h->nsnps + 127) / 128);
  // (measured at 500k x 100k: 6 row splits 3.06 ms, 3 splits 3.27 ms per 8-column pass: ~32 waves)
  uint32_t sv = std::max<uint32_t>((U.nbox + 1023) / 1024, (32 * sm + U.nst - 1) / U.nst);
  sv = std::max<uint32_t>(std::min<uint32_t>(sv, std::max<uint32_t>((U.nbox + 1023) / 1024, U.nbox / 8)), 1);
  U.bps_v = (U.nbox + sv - 1) / sv;
  U.splits_v = (U.nbox + U.bps_v - 1) / U.bps_v;
  U.sstride_t = (h->nsnps + 63) / 64 * 64;
  U.sstride_v = (h->n + 63) / 64 * 64;
  U.vstride = std::max<uint64_t>(U.sstride_t * U.splits_t, U.sstride_v * U.splits_v);
  U.si_bytes = (size_t)U.nst * 16 * kUmmaLanes * 256;
  U.sj_bytes = (size_t)U.nbox * 4 * kUmmaLanes * 256;
  FPB_CUDA(h, cudaMalloc(&U.s_i, U.si_bytes));
  FPB_CUDA(h, cudaMalloc(&U.s_j, U.sj_bytes));
  FPB_CUDA(h, cudaMalloc(&U.part, sizeof(double) * kUmmaLanes * U.vstride));
  FPB_CUDA(h, cudaMalloc(&U.a, sizeof(double) * kUmmaLanes * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&U.corr, sizeof(double) * kUmmaLanes * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&U.pmax, sizeof(double) * kUmmaLanes * h->max_parts));
  FPB_CUDA(h, cudaMalloc(&U.psum, sizeof(double) * kUmmaLanes * h->max_parts));
  FPB_CUDA(h, cudaMalloc(&U.sc, sizeof(fpb::VecScale) * 2 * kUmmaLanes));
  if (h->nmissing) {
    FPB_CUDA(h, cudaMalloc(&U.mx, sizeof(double) * kUmmaLanes * h->nsnps * h->gtiles_s));
    FPB_CUDA(h, cudaMalloc(&U.mc, sizeof(double) * kUmmaLanes * h->n * h->gtiles_i));
  }
  FPB_CUDA(h, cudaMalloc(&U.err, sizeof(uint32_t)));
  FPB_CUDA(h, cudaMemsetAsync(U.err, 0, sizeof(uint32_t), h->stream));
  FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_umma_xt<4, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpb::kUSmemBytes));
  FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_umma_xt<8, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpb::kUSmemBytes));
  FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_umma_xv<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpb::kUSmemBytes));
  FPB_CUDA(h, cudaFuncSetAttribute(fpb::k_umma_xv<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpb::kUSmemBytes));
  U.ready = true;
  return 0;
}

void free_umma(fpb_handle* h) {
  fpb_handle::Umma& U = h->U;
  cudaFree(U.s_i); cudaFree(U.s_j); cudaFree(U.part); cudaFree(U.a); cudaFree(U.corr);
  cudaFree(U.pmax); cudaFree(U.psum); cudaFree(U.sc); cudaFree(U.mx); cudaFree(U.mc); cudaFree(U.err);
  U = fpb_handle::Umma();
}

// missing-genotype sums of one lane on the side stream (fork_mark() first, join_gather() after)
void gather_launch_to(fpb_handle* h, bool by_snp, const double* vec, double* out) {
  const uint64_t nrows = by_snp ? h->nsnps : h->n, veclen = by_snp ? h->n : h->nsnps;
  const uint32_t ntiles = by_snp ? h->gtiles_s : h->gtiles_i;
  cudaStreamWaitEvent(h->side, h->ev_fork, 0);
  const uint32_t nblk = (uint32_t)((nrows + 31) / 32);
  uint32_t chunks = std::max<uint32_t>(1, (4u * h->sm_count) / ntiles);
  uint32_t blocks_per_cta = (nblk + chunks - 1) / chunks;
  chunks = (nblk + blocks_per_cta - 1) / blocks_per_cta;
  dim3 grid(ntiles, chunks);
  fpb::k_sell_gather<<<grid, fpb::kGatherThreads, fpb::kGatherSmem, h->side>>>(
      by_snp ? h->d_seg_s : h->d_seg_i, by_snp ? h->d_col16_s : h->d_col16_i, vec, veclen, nrows,
      nblk, blocks_per_cta, out);
  cudaEventRecord(h->ev_join, h->side);
  h->launches++;
}

// First halves of nv (3..8) vectors, columns of d_x (leading dimension N): t = X'x into the columns
// of d_t (leading dimension nsnps; may be null) and/or the a, corr inputs of the second half.
void umma_crossprod_block(fpb_handle* h, const double* d_x, uint32_t nv, double* d_t, bool second_half) {
  fpb_handle::Umma& U = h->U;
  const uint32_t NV = nv <= 4 ? 4 : 8;
  const uint32_t nkb = U.nst * 16;
  if (nv < NV) cudaMemsetAsync(U.s_i, 0, (size_t)nkb * NV * 256, h->stream);  // unused lanes: zero digits
  if (h->nmissing) fork_mark(h);
  for (uint32_t v = 0; v < nv; v++) {
    const double* xv = d_x + (uint64_t)v * h->n;
    if (h->nmissing) gather_launch_to(h, true, xv, U.mx + (uint64_t)v * h->nsnps * h->gtiles_s);
    double* pm = U.pmax + (uint64_t)v * h->max_parts;
    double* ps = U.psum + (uint64_t)v * h->max_parts;
    fpb::k_vec_partial<<<kVecBlocks, 256, 0, h->stream>>>(xv, h->n, pm, ps);
    fpb::k_slice_umma_i<<<(nkb + 127) / 128, 128, 0, h->stream>>>(
        xv, h->n, nkb, NV, v, pm, ps, kVecBlocks, U.sc + v, reinterpret_cast<uint4*>(U.s_i));
    h->launches += 2;
  }
  dim3 grid((uint32_t)((h->nsnps + 511) / 512), U.splits_t);
  if (NV == 4)
    fpb::k_umma_xt<4, 4, 1><<<grid, fpb::UmmaXtCfg<4, 4, 1>::Threads, fpb::kUSmemBytes, h->stream>>>(
        h->tm_f, (uint32_t)h->nsnps, U.s_i, U.nst, U.sps_t, U.part, U.vstride, U.sstride_t, U.err);
  else
    fpb::k_umma_xt<8, 4, 1><<<grid, fpb::UmmaXtCfg<8, 4, 1>::Threads, fpb::kUSmemBytes, h->stream>>>(
        h->tm_f, (uint32_t)h->nsnps, U.s_i, U.nst, U.sps_t, U.part, U.vstride, U.sstride_t, U.err);
  h->launches++;
  U.used = true;
  if (h->nmissing) join_gather(h);
  const uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  for (uint32_t v = 0; v < nv; v++) {
    fpb::k_finalize_crossprod<<<gb, 256, 0, h->stream>>>(
        U.part + (uint64_t)v * U.vstride, U.splits_t, U.sstride_t, (uint32_t)h->nsnps, U.sc + v,
        h->d_scale, h->nmissing ? U.mx + (uint64_t)v * h->nsnps * h->gtiles_s : nullptr, h->gtiles_s,
        d_t ? d_t + (uint64_t)v * h->nsnps : nullptr,
        second_half ? U.a + (uint64_t)v * h->nsnps : nullptr, U.corr + (uint64_t)v * h->nsnps,
        U.pmax + (uint64_t)v * h->max_parts, U.psum + (uint64_t)v * h->max_parts);
    h->launches++;
  }
}

// a, corr and the (max|a|, sum b) partials of the lanes from user vectors (columns of d_v, ld nsnps)
void umma_prod_inputs(fpb_handle* h, const double* d_v, uint32_t nv) {
  fpb_handle::Umma& U = h->U;
  const uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  for (uint32_t v = 0; v < nv; v++) {
    fpb::k_prod_inputs<<<gb, 256, 0, h->stream>>>(
        d_v + (uint64_t)v * h->nsnps, h->d_scale, (uint32_t)h->nsnps, U.a + (uint64_t)v * h->nsnps,
        U.corr + (uint64_t)v * h->nsnps, U.pmax + (uint64_t)v * h->max_parts,
        U.psum + (uint64_t)v * h->max_parts);
    h->launches++;
  }
}

// Second halves of nv vectors from the lanes' a, corr and partials: y = X v into the columns of d_y
void umma_prod_block(fpb_handle* h, uint32_t nv, double* d_y) {
  fpb_handle::Umma& U = h->U;
  const uint32_t NV = nv <= 4 ? 4 : 8;
  const uint32_t nkb = U.nbox * 4;
  const uint32_t gbs = (uint32_t)((h->nsnps + 255) / 256);  // blocks that wrote the lanes' partials
  if (nv < NV) cudaMemsetAsync(U.s_j, 0, (size_t)nkb * NV * 256, h->stream);
  if (h->nmissing) fork_mark(h);
  for (uint32_t v = 0; v < nv; v++) {
    if (h->nmissing)
      gather_launch_to(h, false, U.corr + (uint64_t)v * h->nsnps, U.mc + (uint64_t)v * h->n * h->gtiles_i);
    fpb::k_slice_umma_j<<<(nkb + 127) / 128, 128, 0, h->stream>>>(
        U.a + (uint64_t)v * h->nsnps, h->nsnps, nkb, NV, v, U.pmax + (uint64_t)v * h->max_parts,
        U.psum + (uint64_t)v * h->max_parts, gbs, U.sc + kUmmaLanes + v, reinterpret_cast<uint4*>(U.s_j));
    h->launches++;
  }
  dim3 grid(U.nst, U.splits_v);
  if (NV == 4)
    fpb::k_umma_xv<4, 4><<<grid, (8 + 4 + 1) * 32, fpb::kUSmemBytes, h->stream>>>(
        h->tm_f, (uint32_t)h->n, U.s_j, U.nbox, U.bps_v, U.part, U.vstride, U.sstride_v, U.err);
  else
    fpb::k_umma_xv<8, 4><<<grid, (8 + 4 + 1) * 32, fpb::kUSmemBytes, h->stream>>>(
        h->tm_f, (uint32_t)h->n, U.s_j, U.nbox, U.bps_v, U.part, U.vstride, U.sstride_v, U.err);
  h->launches++;
  U.used = true;
  if (h->nmissing) join_gather(h);
  const uint32_t gb = (uint32_t)((h->n + 255) / 256);
  for (uint32_t v = 0; v < nv; v++) {
    fpb::k_finalize_prod<<<gb, 256, 0, h->stream>>>(
        U.part + (uint64_t)v * U.vstride, U.splits_v, U.sstride_v, h->n, U.sc + kUmmaLanes + v,
        h->nmissing ? U.mc + (uint64_t)v * h->n * h->gtiles_i : nullptr, h->gtiles_i,
        d_y + (uint64_t)v * h->n);
    h->launches++;
  }
}

// A timed-out wait inside a tcgen05 kernel (a protocol failure) is reported through a device
// word instead of a hang; surfaced at the API's synchronisation points.
int check_umma(fpb_handle* h) {
  fpb_handle::Umma& U = h->U;
  if (!U.used) return 0;
  U.used = false;
  uint32_t code = 0;
  FPB_CUDA(h, cudaMemcpyAsync(&code, U.err, sizeof(code), cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (code) {
    cudaMemsetAsync(U.err, 0, sizeof(code), h->stream);
    char buf[64];
    snprintf(buf, sizeof buf, "tcgen05 block kernel: wait timed out (code %x)", code);
    FPB_FAIL(h, buf);
  }
  return 0;
}
