// Host side of the fused single-pass kernel: launch sequence and error surfacing.  Textually
// included by fpb_capi.cu (inside its anonymous namespace).
// y = X X' x in one pass over the packed matrix (fpb_fused.cuh).  The missing-genotype
// sums of the first half only depend on x and run before the fused kernel; the
// second gather needs every corr_j and runs after it.
void fused_perform_op(fpb_handle* h, const double* d_x, double* d_y) {
  if (h->nmissing) {
    fork_mark(h);
    gather_launch(h, true, d_x);
  }
  vec_partials(h, d_x, h->n);
  const uint32_t nwq = h->nchunks_s * fpb::kChunkWords;
  fpb::k_slice_vec<<<(nwq + 127) / 128, 128, 0, h->stream>>>(
      d_x, h->n, nwq, h->d_pmax, h->d_psum, h->nparts, h->d_sc + 0, h->d_slices);
  cudaMemsetAsync(h->d_arep, 0xFF, sizeof(double) * fpb::kFReplicas * h->f_rep_stride,
                  h->stream);  // all-ones = not written
  if (h->nmissing) join_gather(h);
  fpb::FusedArgs a;
  a.n = (uint32_t)h->n;
  a.nsnps = (uint32_t)h->nsnps;
  a.nslabs = h->f_nslabs;
  a.nstripes = h->f_nstripes;
  a.window = h->f_window;
  a.gpad = h->f_gpad;
  a.mx_tiles = h->nmissing ? h->gtiles_s : 0;
  a.pol1 = h->f_pol1;
  a.pol2 = h->f_pol2;
  a.prefetch = h->f_prefetch;
  {
    static const char* dm = getenv("FPB_FUSED_DBGMODE");
    a.dbg_mode = dm ? (uint32_t)atoi(dm) : 0u;
  }
  a.xslices = h->d_slices;
  a.sc_x = h->d_sc + 0;
  a.scale = h->d_scale;
  a.mxv = h->nmissing ? h->d_mx : nullptr;
  a.part = h->d_fpart;
  a.a_out = h->d_a;
  a.a_rep = h->d_arep;
  a.rep_stride = h->f_rep_stride;
  a.corr_out = h->d_corr;
  a.ybuf = h->d_ybuf;
  a.f_out = h->d_part;
  a.err = h->d_fsync;
  a.dbg = h->d_fdbg;
  void* params[] = {(void*)&h->tm_f, (void*)&a};
  if (h->time_gemv) cudaEventRecord(h->kev[0], h->stream);
  cudaLaunchCooperativeKernel((const void*)fpb::k_fused_op, dim3(h->f_grid), dim3(fpb::kFThreads),
                              params, (size_t)fpb::kFSmemBytes, h->stream);
  if (h->time_gemv) cudaEventRecord(h->kev[1], h->stream);
  h->fused_used = true;
  if (h->nmissing) {
    fork_mark(h);
    gather_launch(h, false, h->d_corr);
  }
  fpb::k_fused_sum_b<<<1, 1024, 0, h->stream>>>(h->d_a, h->d_scale, (uint32_t)h->nsnps,
                                                 h->d_sc + 1);
  if (h->nmissing) join_gather(h);
  const uint32_t gb = (uint32_t)((h->n + 255) / 256);
  fpb::k_finalize_prod<<<gb, 256, 0, h->stream>>>(h->d_part, 1, h->part_stride, h->n, h->d_sc + 1,
                                                  h->nmissing ? h->d_mc : nullptr, h->gtiles_i,
                                                  d_y);
  h->launches += 4;  // slicing, fused op, Sb, finalize (gathers and partials count themselves)
}

// The fused kernel reports a timed-out wait (a protocol failure) through a device
// word instead of hanging; surfaced at the API's synchronisation points.
int check_fused(fpb_handle* h) {
  for (fpb_handle* kid : h->kids)
    if (check_fused(kid)) {
      h->err = kid->err;
      return 1;
    }
  if (!h->fused_used) return 0;
  h->fused_used = false;
  uint32_t code = 0;
  uint32_t* d_err = h->d_fsync;
  FPB_CUDA(h, cudaMemcpyAsync(&code, d_err, sizeof(code), cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (code) {
    cudaMemsetAsync(d_err, 0, sizeof(code), h->stream);
    cudaMemsetAsync(h->d_fpart, 0xFF,
                    sizeof(double) * (size_t)fpb::kFASlots * fpb::kFRows * h->f_gpad, h->stream);
    FPB_FAIL(h, "fused perform_op kernel: wait timed out (code " + std::to_string(code) + ")");
  }
  return 0;
}
