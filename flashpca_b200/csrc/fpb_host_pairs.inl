// Two vectors per pass: host orchestration of the block variants.  Textually included by
// fpb_capi.cu (inside its anonymous namespace).
// ------------------------------ two vectors per pass ------------------------
// The block variants process their columns in pairs: per-vector small kernels run
// once per lane (lane 1 = a second set of scratch buffers, swapped in around the
// calls), the two contraction kernels read the packed matrix once for both.

bool pair_capable(const fpb_handle* h) {
  static const bool off = getenv("FPB_PAIR") && atoi(getenv("FPB_PAIR")) == 0;
  return !off && h->kids.empty() && !h->dense && h->use_imma && h->use_tma && h->single_copy;
}

int ensure_lane1(fpb_handle* h) {
  if (h->L1.slices) return 0;
  FPB_CUDA(h, cudaMalloc(&h->L1.slices, h->slice_bytes));
  FPB_CUDA(h, cudaMalloc(&h->L1.part, sizeof(double) * h->part_elems));
  FPB_CUDA(h, cudaMalloc(&h->L1.a, sizeof(double) * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&h->L1.corr, sizeof(double) * h->nsnps));
  FPB_CUDA(h, cudaMalloc(&h->L1.pmax, sizeof(double) * h->max_parts));
  FPB_CUDA(h, cudaMalloc(&h->L1.psum, sizeof(double) * h->max_parts));
  FPB_CUDA(h, cudaMalloc(&h->L1.sc, sizeof(fpb::VecScale) * 2));
  if (h->nmissing) {
    FPB_CUDA(h, cudaMalloc(&h->L1.mx, sizeof(double) * h->nsnps * h->gtiles_s));
    FPB_CUDA(h, cudaMalloc(&h->L1.mc, sizeof(double) * h->n * h->gtiles_i));
  }
  return 0;
}
void swap_lane(fpb_handle* h) {
  std::swap(h->d_slices, h->L1.slices);
  std::swap(h->d_part, h->L1.part);
  std::swap(h->d_a, h->L1.a);
  std::swap(h->d_corr, h->L1.corr);
  std::swap(h->d_pmax, h->L1.pmax);
  std::swap(h->d_psum, h->L1.psum);
  std::swap(h->d_sc, h->L1.sc);
  std::swap(h->d_mx, h->L1.mx);
  std::swap(h->d_mc, h->L1.mc);
  std::swap(h->nparts, h->L1.nparts);
}

// first halves of two vectors: t = X'x (d_t*, optional) and/or the inputs of the second half
void imma_crossprod_pair(fpb_handle* h, const double* d_x0, const double* d_x1, double* d_t0,
                         double* d_t1, bool second_half) {
  const uint32_t nwq = h->nchunks_s * fpb::kChunkWords;
  if (h->nmissing) fork_mark(h);
  for (int l = 0; l < 2; l++) {
    const double* d_x = l ? d_x1 : d_x0;
    if (l) swap_lane(h);
    if (h->nmissing) gather_launch(h, true, d_x);
    vec_partials(h, d_x, h->n);
    fpb::k_slice_vec<<<(nwq + 127) / 128, 128, 0, h->stream>>>(
        d_x, h->n, nwq, h->d_pmax, h->d_psum, h->nparts, h->d_sc + 0, h->d_slices);
    h->launches++;
    if (l) swap_lane(h);
  }
  const uint32_t rows = (uint32_t)h->nsnps;
  dim3 grid((rows + fpb::kTmaRows - 1) / fpb::kTmaRows, h->tsplits_s);
  fpb::k_imma_gemv_tma_2v<<<grid, (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes,
                            h->stream>>>(h->tm_s, rows, h->d_slices, h->L1.slices, h->nstages_s,
                                         h->sps_s, h->d_part, h->L1.part, h->part_stride);
  h->launches++;
  if (h->nmissing) join_gather(h);
  const uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  for (int l = 0; l < 2; l++) {
    if (l) swap_lane(h);
    fpb::k_finalize_crossprod<<<gb, 256, 0, h->stream>>>(
        h->d_part, h->tsplits_s, h->part_stride, (uint32_t)h->nsnps, h->d_sc + 0, h->d_scale,
        h->nmissing ? h->d_mx : nullptr, h->gtiles_s, l ? d_t1 : d_t0,
        second_half ? h->d_a : nullptr, h->d_corr, h->d_pmax, h->d_psum);
    if (second_half) h->nparts = gb;
    h->launches++;
    if (l) swap_lane(h);
  }
}

// second halves of two vectors from the a, corr and partials in the two lanes
void imma_prod_tail_pair(fpb_handle* h, double* d_y0, double* d_y1) {
  const uint32_t ngroups4 = h->ttiles * (fpb::kTmaRows / 4);
  if (h->nmissing) fork_mark(h);
  for (int l = 0; l < 2; l++) {
    if (l) swap_lane(h);
    if (h->nmissing) gather_launch(h, false, h->d_corr);
    fpb::k_slice_vec_k<<<(ngroups4 + 127) / 128, 128, 0, h->stream>>>(
        h->d_a, h->nsnps, ngroups4, h->d_pmax, h->d_psum, h->nparts, h->d_sc + 1,
        reinterpret_cast<uint32_t*>(h->d_slices));
    h->launches++;
    if (l) swap_lane(h);
  }
  dim3 grid(h->nstages_s, h->ttsplits);
  fpb::k_imma_gemv_tma_t_2v<<<grid, (fpb::kTmaConsumerWarps + 1) * 32, fpb::kTmaSmemBytes,
                              h->stream>>>(h->tm_s, (uint32_t)h->n,
                                           reinterpret_cast<const uint32_t*>(h->d_slices),
                                           reinterpret_cast<const uint32_t*>(h->L1.slices), h->ttiles,
                                           h->ttps, h->d_part, h->L1.part, h->part_stride);
  h->launches++;
  if (h->nmissing) join_gather(h);
  const uint32_t gb = (uint32_t)((h->n + 255) / 256);
  for (int l = 0; l < 2; l++) {
    if (l) swap_lane(h);
    fpb::k_finalize_prod<<<gb, 256, 0, h->stream>>>(h->d_part, h->ttsplits, h->part_stride, h->n,
                                                    h->d_sc + 1, h->nmissing ? h->d_mc : nullptr,
                                                    h->gtiles_i, l ? d_y1 : d_y0);
    h->launches++;
    if (l) swap_lane(h);
  }
}

void prod_inputs_pair(fpb_handle* h, const double* d_v0, const double* d_v1) {
  const uint32_t gb = (uint32_t)((h->nsnps + 255) / 256);
  for (int l = 0; l < 2; l++) {
    if (l) swap_lane(h);
    fpb::k_prod_inputs<<<gb, 256, 0, h->stream>>>(l ? d_v1 : d_v0, h->d_scale, (uint32_t)h->nsnps,
                                                  h->d_a, h->d_corr, h->d_pmax, h->d_psum);
    h->nparts = gb;
    h->launches++;
    if (l) swap_lane(h);
  }
}
