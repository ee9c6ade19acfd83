// Second-half finalize fused with the SNP-shard sum over NVLink peer memory.
//
// The sharded op ends with y = sum_g y_g over the GPUs of the box (the block sum
// of svdwide.cpp:48-59, distributed).  NCCL's all-reduce of the 4 MB vector costs 54 us on 8 GPUs,
// half of it latency (profiles/r02_allreduce_probe_8gpu.txt).  k_allreduce_peer<kPeerFinalize> does the
// finalize step of the second half *and* the exchange in one launch:
//
//   phase 0  CTA b turns its share of the split partials into the local y_g (k_finalize_prod's
//            formula) and stores it in this GPU's exchange buffer `loc`,
//   flag A   tells CTA b of every peer that the share is there (st.release.sys into the peer's
//            flag word, one word per (CTA, source rank)),
//   phase 1  CTA b of rank r sums sub-slice (r, b) over all ranks *in rank order* (P2P loads through
//            NVSwitch) and stores the sum into every GPU's result buffer `res` (P2P stores),
//   flag B   tells the peers the sub-slice has landed,
//   phase 2  copies the sub-slices (g, b), all g, from `res` to the caller's y.
//
// (kPeerSum: phase 0 copies an existing vector instead -- block forms, other paths.  kPeerGather:
// no phase 0 and phase 1 stores the rank's own slice instead of a sum -- the all-gather behind the
// host-pointer perform_op, whose input every rank holds: each rank uploads 1/G of it over PCIe.)
//
// Every element is summed by exactly one rank in a fixed order, so all ranks hold bit-identical y
// (the replicated Lanczos drivers must take identical decisions) and the result does not depend on
// timing.  A CTA only ever waits for the CTA of the same index on the other GPUs, so there is no
// dependency between the CTAs of one GPU; the grid (128 CTAs, one per SM) is always resident.  Flags carry a per-CTA epoch that the kernel advances itself (device memory), so
// the launch is a plain kernel node and the whole op replays as a CUDA graph.  Every wait is bounded
// (FPB_PEER_TIMEOUT_S, default 60 s: ranks may be skewed by host work between two ops) and reports
// through the handle's error word instead of hanging the GPU.
#pragma once

#include <cstdint>

#include "fpb_imma.cuh"

namespace fpb {

constexpr int kPeerMax = 8;          // GPUs of one box
constexpr int kPeerCtas = 128;       // CTAs of the exchange kernel between GPUs: one per SM, all resident
constexpr int kPeerCtasShared = 32;  // ... when the linked shards share one GPU (tests): the grids of
                                     // up to 4 ranks must be resident together
constexpr int kPeerThreads = 1024;

struct PeerView {
  double* loc[kPeerMax];      // exchange buffers (partial sums), by rank; [rank] is local
  double* res[kPeerMax];      // result buffers, by rank
  uint32_t* flag_a[kPeerMax]; // flag words of each rank: [cta * kPeerMax + source rank]
  uint32_t* flag_b[kPeerMax];
  uint32_t* epoch;            // local, one word per CTA
  uint32_t* err;              // watchdog word (mapped host memory)
  unsigned long long timeout_ns;  // bound of every wait
  int rank, world;
};

__device__ __forceinline__ void peer_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t peer_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double peer_ld(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// signal `flags` word (blockIdx, my rank) on every peer, then wait for all peers' words here
__device__ __forceinline__ bool peer_barrier(const PeerView& pv, uint32_t* const* flags, uint32_t e) {
  __syncthreads();
  bool ok = true;
  if (threadIdx.x < (unsigned)pv.world) {
    const int g = threadIdx.x;
    if (g != pv.rank) {
      peer_st_release(flags[g] + blockIdx.x * kPeerMax + pv.rank, e);
      const uint32_t* mine = flags[pv.rank] + blockIdx.x * kPeerMax + g;
      const unsigned long long t0 = peer_now();
      while ((int32_t)(peer_ld_acquire(mine) - e) < 0) {
        if (peer_now() - t0 > pv.timeout_ns) {
          *reinterpret_cast<volatile uint32_t*>(pv.err) = 0x50000000u | (blockIdx.x << 8) | g;
          __threadfence_system();
          ok = false;
          break;
        }
      }
    }
  }
  return __syncthreads_and(ok);
}

// MODE kPeerSum:      phase 0 copies src into the exchange buffer, phase 1 sums over the ranks.
// MODE kPeerFinalize: phase 0 is k_finalize_prod (part / sc_ab / mcv as there), phase 1 sums.
// MODE kPeerGather:   all-gather -- src holds this rank's slice only (the host uploaded just that
//                     part of a replicated vector); phase 1 stores it into every rank's result buffer.
constexpr int kPeerSum = 0, kPeerFinalize = 1, kPeerGather = 2;

template <int MODE>
__global__ void __launch_bounds__(kPeerThreads, 1)
k_allreduce_peer(PeerView pv, const double* __restrict__ src, const double* __restrict__ part,
                 uint32_t nsplits, uint64_t stride, const VecScale* __restrict__ sc_ab,
                 const double* __restrict__ mcv, uint32_t mc_tiles, uint64_t mc_stride,
                 uint64_t count, double* __restrict__ y) {
  constexpr bool FUSED = MODE == kPeerFinalize;
  const int W = pv.world, r = pv.rank;
  const uint64_t slice = (count + W - 1) / W;                     // elements per rank
  const uint64_t sub = (slice + gridDim.x - 1) / gridDim.x;       // elements per (rank, CTA)
  const uint32_t e = pv.epoch[blockIdx.x] + 1;
  double* loc = pv.loc[r];
  double delta = 0.0, sumb = 0.0;
  if (FUSED) {
    delta = sc_ab->delta;
    sumb = sc_ab->sum;
  }
  // phase 0: the sub-slices (g, b) of the local partial sum
  for (int g = 0; MODE != kPeerGather && g < W; g++) {
    const uint64_t lo = g * slice + blockIdx.x * sub;
    const uint64_t hi = min(min(lo + sub, (g + 1) * slice), count);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) {
      double v;
      if (FUSED) {
        double f = 0.0;
#pragma unroll 4
        for (uint32_t s = 0; s < nsplits; s++) f += part[(uint64_t)s * stride + i];
        double mc = 0.0;
        if (mcv) {
#pragma unroll 4
          for (uint32_t tt = 0; tt < mc_tiles; tt++) mc += mcv[(uint64_t)tt * mc_stride + i];
        }
        v = f * delta - sumb + mc;
      } else {
        v = src[i];
      }
      loc[i] = v;
    }
  }
  if (!peer_barrier(pv, pv.flag_a, e)) return;
  // phase 1: sum sub-slice (r, b) over the ranks in rank order (gather: take it from src), store it
  // everywhere
  {
    const uint64_t lo = r * slice + blockIdx.x * sub;
    const uint64_t hi = min(min(lo + sub, (r + 1) * slice), count);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) {
      double s;
      if (MODE == kPeerGather) {
        s = src[i];
      } else {
        double v[kPeerMax];
#pragma unroll
        for (int g = 0; g < kPeerMax; g++)
          if (g < W) v[g] = peer_ld(pv.loc[g] + i);
        s = v[0];
#pragma unroll
        for (int g = 1; g < kPeerMax; g++)
          if (g < W) s += v[g];
      }
#pragma unroll
      for (int g = 0; g < kPeerMax; g++)
        if (g < W) pv.res[g][i] = s;
    }
  }
  if (!peer_barrier(pv, pv.flag_b, e)) return;
  // phase 2: the sub-slices (g, b) of the sum
  const double* res = pv.res[r];
  for (int g = 0; g < W; g++) {
    const uint64_t lo = g * slice + blockIdx.x * sub;
    const uint64_t hi = min(min(lo + sub, (g + 1) * slice), count);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += kPeerThreads) y[i] = __ldcg(res + i);
  }
  if (threadIdx.x == 0) pv.epoch[blockIdx.x] = e;
}

}  // namespace fpb
