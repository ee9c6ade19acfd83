// Out-of-HBM streaming: per-op slab loop.  Textually included by fpb_capi.cu (inside its anonymous
// namespace).
// ------------------------------ out-of-HBM streaming -----------------------
// y = sum_b X_b X_b' x over SNP slabs, the reference's own block loop (svdwide.cpp:48-59,
// Data::read_snp_block per block) with the disk re-read replaced by a pinned-host -> HBM
// copy that overlaps the previous slab's kernels.

__global__ void k_axpy1(double* __restrict__ y, const double* __restrict__ t, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) y[i] += t[i];
}

// make slab b resident in buffer b & 1 (copy stream), order the compute stream behind it
void stream_in(fpb_handle* h, size_t b) {
  const int i = (int)(b & 1);
  fpb_handle* kid = h->kids[b];
  if (h->sbuf_holds[i] != (long long)b) {  // (with one or two slabs everything stays resident)
    cudaStreamWaitEvent(h->copy, h->ev_done[i], 0);  // kernels of slab b - 2 are done with the buffer
    cudaMemcpyAsync(h->sbuf[i], h->kid_host[b], kid->pitch_s * kid->nsnps, cudaMemcpyHostToDevice,
                    h->copy);
    cudaEventRecord(h->ev_copied[i], h->copy);
    h->sbuf_holds[i] = (long long)b;
  }
  cudaStreamWaitEvent(h->stream, h->ev_copied[i], 0);
  kid->d_gs = h->sbuf[i];
  kid->tm_s = kid->tm_s_alt[i];
  kid->tm_f = kid->tm_f_alt[i];
}
void stream_done(fpb_handle* h, size_t b) {
  cudaEventRecord(h->ev_done[b & 1], h->stream);
  h->launches += h->kids[b]->launches;
  h->kids[b]->launches = 0;
}

void streaming_perform_op(fpb_handle* h, const double* d_x, double* d_y) {
  const uint32_t gb = (uint32_t)((h->n + 255) / 256);
  for (size_t b = 0; b < h->kids.size(); b++) {
    stream_in(h, b);
    launch_perform_op(h->kids[b], d_x, b == 0 ? d_y : h->d_ytmp);
    if (b > 0) {
      k_axpy1<<<gb, 256, 0, h->stream>>>(d_y, h->d_ytmp, h->n);  // block order, like upstream
      h->launches++;
    }
    stream_done(h, b);
  }
}
void streaming_crossprod(fpb_handle* h, const double* d_x, double* d_t) {
  for (size_t b = 0; b < h->kids.size(); b++) {
    stream_in(h, b);
    launch_crossprod(h->kids[b], d_x, d_t + h->kid_off[b]);
    stream_done(h, b);
  }
}
void streaming_prod(fpb_handle* h, const double* d_v, double* d_y) {
  const uint32_t gb = (uint32_t)((h->n + 255) / 256);
  for (size_t b = 0; b < h->kids.size(); b++) {
    stream_in(h, b);
    launch_prod(h->kids[b], d_v + h->kid_off[b], b == 0 ? d_y : h->d_ytmp);
    if (b > 0) {
      k_axpy1<<<gb, 256, 0, h->stream>>>(d_y, h->d_ytmp, h->n);
      h->launches++;
    }
    stream_done(h, b);
  }
}
