// Host side of the peer-memory shard sum (fpb_peer.cuh); textual include of fpb_capi.cu (inside its
// anonymous namespace).
//
// Every rank allocates one region  [loc: cap doubles][res: cap doubles][flag_a][flag_b][epoch]
// and maps the regions of the other ranks: through CUDA IPC handles when the ranks are processes
// (fpb_comm_init: the 64-byte handles travel over the NCCL communicator that was just created), or
// by plain pointers when they are handles of one process (fpb_comm_link_local: the shards of a
// test on a single GPU, or several GPUs driven by one process).  If a region cannot be mapped
// (no peer access between the devices, FPB_PEER=0) the handle keeps NCCL's all-reduce.

size_t peer_flags_bytes() { return sizeof(uint32_t) * fpb::kPeerCtas * fpb::kPeerMax; }
size_t peer_region_bytes(uint64_t cap) {
  return 2 * sizeof(double) * cap + 2 * peer_flags_bytes() + sizeof(uint32_t) * fpb::kPeerCtas + 256;
}

// pointers into a region with base address `base`
void peer_slot(fpb::PeerView& v, int g, unsigned char* base, uint64_t cap) {
  v.loc[g] = reinterpret_cast<double*>(base);
  v.res[g] = reinterpret_cast<double*>(base) + cap;
  unsigned char* f = base + 2 * sizeof(double) * cap;
  v.flag_a[g] = reinterpret_cast<uint32_t*>(f);
  v.flag_b[g] = reinterpret_cast<uint32_t*>(f + peer_flags_bytes());
}

int peer_alloc(fpb_handle* h) {
  fpb_handle::Peer& P = h->P;
  P.cap = h->n * kUmmaLanes;
  const size_t bytes = peer_region_bytes(P.cap);
  FPB_CUDA(h, cudaMalloc(&P.region, bytes));
  FPB_CUDA(h, cudaMemsetAsync(P.region, 0, bytes, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  unsigned char* base = static_cast<unsigned char*>(P.region);
  unsigned char* tail = base + 2 * sizeof(double) * P.cap + 2 * peer_flags_bytes();
  P.view.epoch = reinterpret_cast<uint32_t*>(tail);
  // the watchdog word lives in mapped pinned host memory: the host reads it after any sync
  // without another copy on the op's critical path
  FPB_CUDA(h, cudaHostAlloc(&P.h_err, sizeof(uint32_t), cudaHostAllocMapped));
  *P.h_err = 0;
  FPB_CUDA(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&P.view.err), P.h_err, 0));
  const char* to = getenv("FPB_PEER_TIMEOUT_S");
  P.view.timeout_ns = (unsigned long long)((to && atof(to) > 0 ? atof(to) : 60.0) * 1e9);
  return 0;
}

void peer_release(fpb_handle* h) {
  fpb_handle::Peer& P = h->P;
  for (int g = 0; g < fpb::kPeerMax; g++)
    if (P.mapped[g]) cudaIpcCloseMemHandle(P.mapped[g]);
  cudaFree(P.region);
  if (P.h_err) cudaFreeHost(P.h_err);
  P = fpb_handle::Peer();
}

// ranks = processes: exchange IPC handles over the (fresh) NCCL communicator
int peer_setup_ipc(fpb_handle* h) {
  static const bool off = getenv("FPB_PEER") && atoi(getenv("FPB_PEER")) == 0;
  fpb_handle::Peer& P = h->P;
  if (off || h->nranks < 2 || h->nranks > fpb::kPeerMax || !g_nccl.AllGather) return 0;
  if (peer_alloc(h)) return 1;
  const int W = h->nranks;
  cudaIpcMemHandle_t mine;
  unsigned char* d_hdl = nullptr;
  std::vector<cudaIpcMemHandle_t> all(W);
  // every rank must take part in the all-gather whatever happened locally: a rank that cannot
  // export sends zeros and everybody falls back together
  bool exported = cudaIpcGetMemHandle(&mine, P.region) == cudaSuccess;
  if (!exported) {
    cudaGetLastError();
    memset(&mine, 0, sizeof mine);
  }
  FPB_CUDA(h, cudaMalloc(&d_hdl, sizeof(mine) * W));
  FPB_CUDA(h, cudaMemcpyAsync(d_hdl + sizeof(mine) * h->rank, &mine, sizeof(mine),
                              cudaMemcpyHostToDevice, h->stream));
  int rc = g_nccl.AllGather(d_hdl + sizeof(mine) * h->rank, d_hdl, sizeof(mine), /*ncclUint8*/ 1,
                            h->comm, h->stream);
  if (rc != 0) {
    cudaFree(d_hdl);
    FPB_FAIL(h, "ncclAllGather of the peer-memory handles failed");
  }
  FPB_CUDA(h, cudaMemcpyAsync(all.data(), d_hdl, sizeof(mine) * W, cudaMemcpyDeviceToHost,
                              h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_hdl);
  bool ok = true;
  static const cudaIpcMemHandle_t zero = {};
  for (int g = 0; g < W; g++)
    if (memcmp(&all[g], &zero, sizeof zero) == 0) ok = false;
  for (int g = 0; ok && g < W; g++) {
    if (g == h->rank) {
      peer_slot(P.view, g, static_cast<unsigned char*>(P.region), P.cap);
      continue;
    }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;
      break;
    }
    P.mapped[g] = p;
    peer_slot(P.view, g, static_cast<unsigned char*>(p), P.cap);
  }
  // all ranks must agree (one that failed to map would wait on NCCL while the others spin on
  // flags): min over ranks of `ok`, again over the communicator
  int* d_ok = nullptr;
  FPB_CUDA(h, cudaMalloc(&d_ok, sizeof(int)));
  const int mine_ok = ok ? 1 : 0;
  int all_ok = 0;
  FPB_CUDA(h, cudaMemcpyAsync(d_ok, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  rc = g_nccl.AllReduce(d_ok, d_ok, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, h->comm, h->stream);
  if (rc != 0) {
    cudaFree(d_ok);
    FPB_FAIL(h, "ncclAllReduce of the peer-memory status failed");
  }
  FPB_CUDA(h, cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FPB_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_ok);
  if (!all_ok) {
    peer_release(h);
    return 0;
  }
  P.view.rank = h->rank;
  P.view.world = W;
  P.grid = fpb::kPeerCtas;
  P.ok = true;
  return 0;
}

uint32_t peer_grid(const fpb_handle* h) { return h->P.grid; }

// in-place sum of d_buf[count] over the ranks
int peer_allreduce(fpb_handle* h, double* d_buf, size_t count) {
  fpb_handle::Peer& P = h->P;
  for (size_t off = 0; off < count; off += P.cap) {
    const uint64_t len = std::min<uint64_t>(P.cap, count - off);
    fpb::k_allreduce_peer<fpb::kPeerSum><<<peer_grid(h), fpb::kPeerThreads, 0, h->stream>>>(
        P.view, d_buf + off, nullptr, 0, 0, nullptr, nullptr, 0, 0, len, d_buf + off);
    h->launches++;
  }
  P.used = true;
  return 0;
}

// k_finalize_prod + shard sum in one launch (single-vector second half)
void peer_finalize_prod(fpb_handle* h, uint32_t nsplits, double* d_y) {
  fpb_handle::Peer& P = h->P;
  fpb::k_allreduce_peer<fpb::kPeerFinalize><<<peer_grid(h), fpb::kPeerThreads, 0, h->stream>>>(
      P.view, nullptr, h->d_part, nsplits, h->part_stride, h->d_sc + 1,
      h->nmissing ? h->d_mc : nullptr, h->gtiles_i, h->n, h->n, d_y);
  h->launches++;
  P.used = true;
  P.summed = true;
}

// Host -> device copy of a vector (or N x k block) that every rank holds: each rank uploads only
// its slice and the slices are exchanged over NVLink (all-gather through the result buffers) --
// two GPUs share one PCIe uplink on an HGX board, NVLink is 15x wider.  d_dst receives all of it.
int peer_upload_replicated(fpb_handle* h, const double* host_src, double* d_dst, size_t count) {
  fpb_handle::Peer& P = h->P;
  const uint64_t W = (uint64_t)P.view.world, r = (uint64_t)P.view.rank;
  for (size_t off = 0; off < count; off += P.cap) {
    const uint64_t len = std::min<uint64_t>(P.cap, count - off);
    const uint64_t slice = (len + W - 1) / W;
    const uint64_t lo = std::min<uint64_t>(r * slice, len), hi = std::min<uint64_t>(lo + slice, len);
    if (hi > lo)
      FPB_CUDA(h, cudaMemcpyAsync(d_dst + off + lo, host_src + off + lo, sizeof(double) * (hi - lo),
                                  cudaMemcpyHostToDevice, h->stream));
    fpb::k_allreduce_peer<fpb::kPeerGather><<<peer_grid(h), fpb::kPeerThreads, 0, h->stream>>>(
        P.view, d_dst + off, nullptr, 0, 0, nullptr, nullptr, 0, 0, len, d_dst + off);
    h->launches++;
  }
  P.used = true;
  return 0;
}

int check_peer(fpb_handle* h) {  // after a synchronisation of the stream
  fpb_handle::Peer& P = h->P;
  if (!P.ok || !P.used) return 0;
  P.used = false;
  const uint32_t code = *reinterpret_cast<volatile uint32_t*>(P.h_err);
  if (code) {
    *P.h_err = 0;
    char buf[80];
    snprintf(buf, sizeof buf, "peer-memory shard sum: wait for a peer timed out (code %x)", code);
    FPB_FAIL(h, buf);
  }
  return 0;
}
