// fpb_fused.cuh -- perform_op with ONE HBM pass over the packed genotypes.
//
// y = X (X' x) needs every genotype twice with a full reduction over the
// individuals in between (svdwide.cpp:21-68 does it block by block).  The
// two-kernel path (fpb_imma.cuh) streams the 2-bit matrix from HBM once per
// half.  Here a persistent grid (one CTA per SM) walks the matrix in slabs of
// kFRows SNPs (16 MB at N = 500k): the first half of a slab is contracted from
// HBM, the per-SNP sums are exchanged between the CTAs through L2, and the
// second half re-reads the slab a few microseconds later while it is still
// resident in the 126 MB L2.  HBM traffic per op = the packed matrix, once.
//
// Work split.  CTA c owns nk <= kFSpc consecutive 128-byte column stripes (512
// individuals each) for the whole launch:
//   * first half  E_j = sum_{i in stripes} e_ij x_i   (partial over the CTA's
//     individuals, one value per SNP of the slab), same fragment mapping as
//     k_imma_gemv_tma; the digit slices of x for the CTA's stripes stay in
//     shared memory for the whole launch;
//   * second half F_i += sum_{j in slab} e_ij a_j for the CTA's individuals,
//     same LDSM.8.MT1616 mapping as k_imma_gemv_tma_t; the int32 accumulators
//     of all stripes (nk x 16 registers per thread) live in registers for the
//     whole launch.
// Roles (512 threads, register budgets set with setmaxnreg):
//   warp 0      producer 1: TMA loads of the first-half tiles (HBM -> ring 1)
//   warp 1      a-slice builder: waits for the slab's a_j and cuts them into int8
//               digit slices in shared memory
//   warp 3      producer 2: TMA loads of the second-half tiles (L2 -> ring 2)
//   warp 2      reducer: sums the G per-CTA partials of "its" SNP of the slab
//               (SNP jl of slab s belongs to CTA (jl + s) mod G), applies the
//               standardisation (k_finalize_crossprod's formula) and publishes
//               a_j, corr_j
//   warps 4-11  second-half consumers (16-byte chunk = 64 individuals each)
//   warps 12-15 first-half consumers (32 SNP rows each)
// Cross-CTA protocol per slab s.  The exchanged words carry their own readiness:
// a 64-bit store is atomic, and an all-ones word (a NaN no computation here
// produces) means "empty".
//   * part[s % ring][row][cta] is a one-word mailbox from CTA `cta` to the reducer
//     of `row`: the first-half warp stores its partial sum once it has seen the
//     word empty (the load is issued a slab ahead, so it costs nothing unless the
//     ring really is full); the reducer polls the row's G words until none is
//     empty, sums them in CTA order and stores all-ones back.  Same-address
//     coherence orders the two parties: no fence, no atomic.
//   * the reducer stores a_j and corr_j; the builders poll the slab's kFRows words
//     of a (set to all-ones by the host before the launch).
// Every sum has a fixed order (no floating-point atomics): results are
// bit-reproducible call to call.
//
// Scale of the second half.  The digits of a_j need a common power-of-two step
// inside one int32 accumulation, but max|a| is only known when every slab is
// done.  Each builder keeps the running maximum exponent ex_s = max(ex_{s-1},
// exponent(max|a| of slab s)) -- the same deterministic sequence in every CTA
// -- and the consumers drain their accumulators to FP64 whenever it grows (a
// handful of times per op).  The step is never coarser than the global one
// used by the two-kernel path, so the error bound of fpb_imma.cuh holds.
//
// First-half tiles are not allowed to run more than `window` slabs ahead of
// the CTA's own second half, which bounds the L2 footprint between the two
// reads of a slab to (window + 1) x 16 MB.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fpb_imma.cuh"

namespace fpb {

constexpr int kFRows = 128;                       // SNP rows per slab = per tile
constexpr int kFTileBytes = kFRows * 128;         // 16 KB: [128 rows x 128 B], SWIZZLE_128B
constexpr int kFSpc = 7;                          // max stripes per CTA (accumulator registers)
constexpr int kFStages1 = 6, kFStages2 = 5;       // tile rings of the two halves
constexpr int kFASlots = 8;                       // a-slice ring, partial ring, max window
constexpr int kFASlotBytes = 16 + (kFRows / 32) * 256;  // header (ex) + 1 KB of digits
constexpr int kFXsBytes = kFSpc * 4096;           // resident digit slices of x
#ifndef FPB_FUSED_P1_GROUPS
#define FPB_FUSED_P1_GROUPS 1
#endif
constexpr int kFP1Groups = FPB_FUSED_P1_GROUPS;   // first-half warpgroups (4 warps each): 1 or 2
constexpr int kFThreads = 128 + 128 * kFP1Groups + 256;
// the scheduler favours the highest warp ids: the first-half warps (one per scheduler) get them
constexpr int kFP2Warp0 = 4, kFP2Warps = 8, kFP1Warp0 = 12, kFP1Warps = 4 * kFP1Groups;
// register budgets per role (setmaxnreg): 512 threads start at 128, 640 threads at 96
constexpr int kFRegCtl = kFP1Groups == 1 ? 56 : 40, kFRegP1 = kFP1Groups == 1 ? 104 : 64,
              kFRegP2 = kFP1Groups == 1 ? 176 : 152;
static_assert(128 * kFRegCtl + 128 * kFP1Groups * kFRegP1 + 256 * kFRegP2 <=
                  kFThreads * (kFP1Groups == 1 ? 128 : 96), "register pool");
constexpr int kFFlushSlabs = 512;                 // 65536 SNPs x 255 x 64 < 2^31

constexpr uint32_t kFOffRing1 = 0;
constexpr uint32_t kFOffRing2 = kFOffRing1 + kFStages1 * kFTileBytes;
constexpr uint32_t kFOffXs = kFOffRing2 + kFStages2 * kFTileBytes;
constexpr uint32_t kFOffA = kFOffXs + kFXsBytes;
constexpr uint32_t kFOffBars = kFOffA + kFASlots * kFASlotBytes;
constexpr int kFNumBars = 2 * kFStages1 + 2 * kFStages2 + 2 * kFASlots + 1;
constexpr uint32_t kFOffMisc = kFOffBars + 8 * kFNumBars;
constexpr int kFSmemBytes = 1024 + kFOffMisc + 16;
static_assert(kFSmemBytes <= 232448, "fused kernel shared memory does not fit");
static_assert(kFOffBars % 8 == 0 && kFOffA % 16 == 0, "alignment");

constexpr int kFExZero = -(1 << 30);   // nothing non-zero seen so far
constexpr int kFExNan = (1 << 30);     // a non-finite value was seen: result is NaN
constexpr long long kFNotYet = -1LL;   // all-ones word: "not written yet"
constexpr int kFDbgSlabs = 256, kFDbgEvents = 8;
constexpr int kFReplicas = 8;           // copies of a the builders read (spreads the L2 hot spot)

struct FusedArgs {
  uint32_t n, nsnps, nslabs, nstripes;
  uint32_t window;             // first half may lead the second by this many slabs (2..kFASlots)
  uint32_t gpad;               // row pitch (doubles) of the partial ring, >= kFP1Groups * gridDim.x
  uint32_t mx_tiles;           // tiles of the Mx partial sums (0: nothing missing)
  uint32_t pol1, pol2;         // L2 hints of the two tile streams: 0 none, 1 evict_first, 2 evict_last
  uint32_t prefetch;           // slabs the L2 prefetch of the first-half tiles runs ahead (0: off)
  uint32_t dbg_mode;           // timing experiments only (results invalid): 1 = the second half
                               // does not wait for the a-slices and the first half is not throttled
  const uint4* xslices;        // k_slice_vec output for x
  const VecScale* sc_x;        // step and sum of x
  const double2* scale;        // per SNP (mean, 1/sd or 0)
  const double* mxv;           // mx_tiles x nsnps partial sums of Mx
  double* part;                // [kFASlots][kFRows][gpad] first-half partial mailboxes (all-ones = empty)
  double* a_out;               // [nsnps] a_j for the kernels that follow
  double* a_rep;               // [kFReplicas][rep_stride] copies the builders poll, all-ones at launch
  uint64_t rep_stride;
  double* corr_out;            // [nsnps]
  double* ybuf;                // [n] drained second-half sums
  double* f_out;               // [n] F_i = sum_j e_ij a_j
  uint32_t* err;               // watchdog: non-zero when a wait timed out
  unsigned long long* dbg;     // optional [2 CTAs][kFDbgSlabs][kFDbgEvents] globaltimer stamps
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_s64(const double* p) {
  long long v;
  asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ld_relaxed_v2_s64(const double* p, long long& a, long long& b) {
  asm volatile("ld.relaxed.gpu.global.v2.s64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}

// Every wait of the kernel goes through one of these: a wait that lasts longer
// than kFTimeoutNs raises the CTA's abort flag and the global error word, and
// every other wait of the CTA gives up when it sees the flag -- a protocol bug
// ends the launch with an error instead of hanging the GPU.
constexpr uint64_t kFTimeoutNs = 2000000000ull;
struct Watch {
  volatile uint32_t* abort_s;
  uint32_t* gerr;
  __device__ __forceinline__ bool expired(uint64_t& t0, uint32_t code) const {
    if (*abort_s) return true;
    if (*reinterpret_cast<volatile uint32_t*>(gerr) != 0) {  // another CTA gave up
      *abort_s = 1;
      return true;
    }
    const uint64_t t = global_timer_ns();
    if (t0 == 0) {
      t0 = t;
    } else if (t - t0 > kFTimeoutNs) {
      *abort_s = 1;
      atomicCAS(gerr, 0u, code);
      return true;
    }
    return false;
  }
};
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity, const Watch& w,
                                         uint32_t code) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try(bar, parity)) {
    if (((++spins) & 0x3FFu) == 0 && w.expired(t0, code)) return false;
  }
  return true;
}
__device__ __forceinline__ bool wait_smem_counter(volatile uint32_t* p, uint32_t target,
                                                  const Watch& w, uint32_t code) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (*p < target) {
    if (((++spins) & 0x3FFu) == 0 && w.expired(t0, code)) return false;
  }
  return true;
}

__device__ __forceinline__ uint64_t l2_policy(uint32_t kind) {
  uint64_t p;
  if (kind == 1) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  } else if (kind == 2) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  } else {
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  }
  return p;
}

template <int R>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}
template <int R>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}

__global__ void __launch_bounds__(kFThreads, 1)
k_fused_op(const __grid_constant__ TmaDesc tmap, const __grid_constant__ FusedArgs A) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles: 1 KB aligned
  uint8_t* const gbase = smem_raw + (base - raw);
  const uint32_t bars = base + kFOffBars;
  constexpr int B_FULL1 = 0, B_EMPTY1 = kFStages1, B_FULL2 = 2 * kFStages1,
                B_EMPTY2 = 2 * kFStages1 + kFStages2, B_AFULL = 2 * kFStages1 + 2 * kFStages2,
                B_AEMPTY = B_AFULL + kFASlots, B_XS = B_AEMPTY + kFASlots;
  volatile uint32_t* const misc = reinterpret_cast<volatile uint32_t*>(gbase + kFOffMisc);
  // misc[0] = abort flag; hints that tell the pollers when polling becomes worthwhile:
  // misc[1] = slabs whose first-half partials warp 4 has stored, misc[2] = slabs reduced
  const Watch watch{misc, A.err};

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x, cta = blockIdx.x;
  const uint32_t basec = A.nstripes / G, rem = A.nstripes % G;
  const uint32_t nk = basec + (cta < rem ? 1u : 0u);              // stripes of this CTA (<= kFSpc)
  const uint32_t stripe0 = cta * basec + (cta < rem ? cta : rem);
  const uint32_t nslabs = A.nslabs;
  // optional time stamps of CTA 0 and CTA G-1 (protocol latency analysis)
  unsigned long long* const dbg =
      (A.dbg && (cta == 0 || cta == G - 1))
          ? A.dbg + (size_t)(cta == 0 ? 0 : 1) * kFDbgSlabs * kFDbgEvents
          : nullptr;
  auto stamp = [&](uint32_t s, int ev) {
    if (dbg && s < (uint32_t)kFDbgSlabs) dbg[s * kFDbgEvents + ev] = global_timer_ns();
  };

  if (tid == 0) {
    for (int i = 0; i < kFStages1; i++) {
      mbar_init(bars + 8 * (B_FULL1 + i), 1);
      mbar_init(bars + 8 * (B_EMPTY1 + i), 4);  // a tile belongs to one warpgroup
    }
    for (int i = 0; i < kFStages2; i++) {
      mbar_init(bars + 8 * (B_FULL2 + i), 1);
      mbar_init(bars + 8 * (B_EMPTY2 + i), kFP2Warps);
    }
    for (int i = 0; i < kFASlots; i++) {
      mbar_init(bars + 8 * (B_AFULL + i), 1);
      mbar_init(bars + 8 * (B_AEMPTY + i), kFP2Warps);
    }
    mbar_init(bars + 8 * B_XS, 1);
    misc[0] = 0;
    misc[1] = 0;
    misc[2] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp < 4) {
    reg_dec<kFRegCtl>();
    if (warp == 0) {
      // ------------------------- producer 1: HBM -> ring 1 -------------------------
      if (lane != 0) return;
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol = l2_policy(A.pol1), pol_keep = l2_policy_evict_last();
      mbar_expect_tx(bars + 8 * B_XS, nk * 4096u);
      bulk_load(base + kFOffXs, A.xslices + (uint64_t)stripe0 * 256u, nk * 4096u, bars + 8 * B_XS,
                pol_keep);
      // The ring holds 96 KB per SM, too little to cover the HBM latency at full bandwidth
      // (~3 us x 7 TB/s / 148 SMs = 140 KB): the tiles are therefore prefetched into L2
      // `prefetch` slabs ahead, and the ring loads are L2 hits.
      auto prefetch_slab = [&](uint32_t sp) {
        for (uint32_t k = 0; k < nk; k++)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap),
                       "r"((int)((stripe0 + k) * 128u)), "r"((int)(sp * kFRows))
                       : "memory");
      };
      for (uint32_t sp = 0; sp < A.prefetch && sp < nslabs; sp++) prefetch_slab(sp);
      uint32_t slot = 0, round = 0;
      for (uint32_t s = 0; s < nslabs; s++) {
        if (A.prefetch && s + A.prefetch < nslabs) prefetch_slab(s + A.prefetch);
        if (s >= A.window && A.dbg_mode != 1) {  // throttle: own second half must have finished slab s - window
          const uint32_t sw = s - A.window;
          if (!wait_bar(bars + 8 * (B_AEMPTY + (sw % kFASlots)), (sw / kFASlots) & 1u, watch, 11))
            return;
        }
        stamp(s, 0);
        for (uint32_t k = 0; k < nk; k++) {
          if (round > 0 && !wait_bar(bars + 8 * (B_EMPTY1 + slot), (round - 1) & 1u, watch, 12))
            return;
          const uint32_t full = bars + 8 * (B_FULL1 + slot);
          mbar_expect_tx(full, kFTileBytes);
          tma_load_2d(base + kFOffRing1 + slot * kFTileBytes, &tmap, (int)((stripe0 + k) * 128u),
                      (int)(s * kFRows), full, pol);
          if (++slot == kFStages1) {
            slot = 0;
            round++;
          }
        }
      }
      return;
    }
    if (warp == 3) {
      // ------------------------- producer 2: L2 -> ring 2 -------------------------
      // Runs at most kFStages2 tiles ahead of the second-half consumers, which trail the
      // first half by at least one slab: these loads find their tiles in L2.
      if (lane != 0) return;
      const uint64_t pol = l2_policy(A.pol2);
      uint32_t slot = 0, round = 0;
      for (uint32_t s = 0; s < nslabs; s++) {
        // do not overtake the first half (the tile would be fetched from HBM twice)
        if (!wait_smem_counter(misc + 1, s + 1, watch, 61)) return;
        for (uint32_t k = 0; k < nk; k++) {
          if (round > 0 && !wait_bar(bars + 8 * (B_EMPTY2 + slot), (round - 1) & 1u, watch, 62))
            return;
          const uint32_t full = bars + 8 * (B_FULL2 + slot);
          mbar_expect_tx(full, kFTileBytes);
          tma_load_2d(base + kFOffRing2 + slot * kFTileBytes, &tmap, (int)((stripe0 + k) * 128u),
                      (int)(s * kFRows), full, pol);
          if (++slot == kFStages2) {
            slot = 0;
            round++;
          }
        }
      }
      return;
    }
    if (warp == 1) {
      // ----------------------------- a-slice builder -----------------------------
      // One warp serves every slab in turn, so a slab must cost it well under the slab period
      // (~2.7 us): the loads of the next slab's a_j are issued before the current slab is
      // quantised, and a slab whose values have all arrived costs no further round trip.
      int ex_run = kFExZero;
      long long nxt[4];
      // Every builder of the grid needs the same 1 KB per slab: the reducers keep kFReplicas
      // copies (different L2 lines), CTA c reads copy c % kFReplicas with 16-byte loads.
      const double* const arep = A.a_rep + (uint64_t)(cta % kFReplicas) * A.rep_stride;
      auto fetch = [&](uint32_t sf) {  // 4 consecutive SNPs per lane (zero behind the last SNP)
        const uint32_t j = sf * kFRows + 4u * lane;   // rep_stride covers whole slabs
        if (sf < nslabs) {
          ld_relaxed_v2_s64(arep + j, nxt[0], nxt[1]);
          ld_relaxed_v2_s64(arep + j + 2, nxt[2], nxt[3]);
#pragma unroll
          for (int b = 0; b < 4; b++)
            if (j + b >= A.nsnps) nxt[b] = 0;
        } else {
          nxt[0] = nxt[1] = nxt[2] = nxt[3] = 0;
        }
      };
      fetch(0);
      for (uint32_t s = 0; s < nslabs; s++) {
        const uint32_t aslot = s % kFASlots, around = s / kFASlots;
        int ok = 1;
        if (lane == 0 && around > 0 && A.dbg_mode != 1 &&
            !wait_bar(bars + 8 * (B_AEMPTY + aslot), (around - 1) & 1u, watch, 21))
          ok = 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return;
        bool ready = nxt[0] != kFNotYet && nxt[1] != kFNotYet && nxt[2] != kFNotYet &&
                     nxt[3] != kFNotYet;
        if (!__all_sync(0xffffffffu, ready)) {
          // All builders of the grid read the same 1 KB: poll only once this CTA's own reducer
          // is through the slab (the others finish within a microsecond of it), with a pause.
          if (lane == 0 && !wait_smem_counter(misc + 2, s + 1, watch, 23)) ok = 0;
          ok = __shfl_sync(0xffffffffu, ok, 0);
          if (!ok) return;
          uint32_t spins = 0;
          uint64_t t0 = 0;
          for (;;) {
            fetch(s);
            ready = nxt[0] != kFNotYet && nxt[1] != kFNotYet && nxt[2] != kFNotYet &&
                    nxt[3] != kFNotYet;
            if (__all_sync(0xffffffffu, ready)) break;
            __nanosleep(100);
            if (((++spins) & 0xFFu) == 0) {
              int dead = (lane == 0 && watch.expired(t0, 22)) ? 1 : 0;
              if (__shfl_sync(0xffffffffu, dead, 0)) return;
            }
          }
        }
        double av[4];
#pragma unroll
        for (int b = 0; b < 4; b++) av[b] = __longlong_as_double(nxt[b]);
        fetch(s + 1);  // in flight while this slab is quantised
        stamp(s, 5);
        double m = 0.0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double t = fabs(av[b]);
          m = (t > m || t != t) ? t : m;  // NaN propagates
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double om = __shfl_xor_sync(0xffffffffu, m, o);
          m = (om > m || om != om) ? om : m;
        }
        int ex_own;
        if (m == 0.0) {
          ex_own = kFExZero;
        } else if (!(m < 1.79e308)) {
          ex_own = kFExNan;
        } else {
          frexp(m, &ex_own);  // |a| < 2^ex_own
        }
        ex_run = max(ex_run, ex_own);
        const bool live = ex_run != kFExZero && ex_run != kFExNan;
        // 2^(54 - ex_run) as a double (exact; |a| < 2^ex_run so the product is < 2^54)
        const double qs =
            live ? __longlong_as_double((long long)(1023 + kSliceBits - ex_run) << 52) : 0.0;
        const bool qs_ok = live && (1023 + kSliceBits - ex_run) > 0 && (1023 + kSliceBits - ex_run) < 2047;
        uint32_t dig[8] = {};
#pragma unroll
        for (int b = 0; b < 4; b++) {
          long long q = 0;
          if (qs_ok) q = __double2ll_rn(av[b] * qs);
          else if (live) q = __double2ll_rn(ldexp(av[b], kSliceBits - ex_run));
#pragma unroll
          for (int sdx = 0; sdx < 8; sdx++) {
            const long long d = (sdx < 7) ? (((q + 64) & 127) - 64) : q;
            q = (q - d) >> 7;
            dig[sdx] |= ((uint32_t)(d & 0xFF)) << (8 * b);
          }
        }
        // K-major B fragments (k_slice_vec_k layout): lane = group of 4 SNPs
        uint32_t* const dst = reinterpret_cast<uint32_t*>(gbase + kFOffA + aslot * kFASlotBytes);
        const uint32_t grp = lane >> 3, kk4 = lane & 7, hh = kk4 >> 2, q4 = kk4 & 3;
#pragma unroll
        for (int sdx = 0; sdx < 8; sdx++) dst[4 + ((grp * 8 + sdx) * 4 + q4) * 2 + hh] = dig[sdx];
        if (lane == 0) dst[0] = (uint32_t)ex_run;
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (B_AFULL + aslot));
        stamp(s, 6);
      }
      return;
    }
    {
      // ------------------------------ reducer (warp 2) ------------------------------
      const VecScale scx = *A.sc_x;
      for (uint32_t s = 0; s < nslabs; s++) {
        const uint32_t r = s % G;
        const uint32_t first = cta >= r ? cta - r : cta + G - r;  // (cta - s) mod G
        // values that do not depend on the first half, fetched ahead of the waits
        double mx_pre = 0.0;
        double2 ms_pre = make_double2(0.0, 0.0);
        const uint32_t jf = s * kFRows + first;
        const bool have_first = first < (uint32_t)kFRows && jf < A.nsnps;
        if (have_first) {
          if (A.mx_tiles)
            for (uint32_t tt = lane; tt < A.mx_tiles; tt += 32)
              mx_pre += __ldcg(A.mxv + (uint64_t)tt * A.nsnps + jf);
          ms_pre = A.scale[jf];
        }
        int ok = 1;
        if (lane == 0 && !wait_smem_counter(misc + 1, s + 1, watch, 31)) ok = 0;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) return;
        for (uint32_t jl = first; jl < (uint32_t)kFRows; jl += G) {
          const uint32_t j = s * kFRows + jl;
          if (j >= A.nsnps) break;
          double* const pp = A.part + ((uint64_t)(s % kFASlots) * kFRows + jl) * A.gpad;
          // poll the row until both partials of every CTA have landed (2G mailboxes)
          const uint32_t G2 = kFP1Groups * G;
          double v[10];
          {
            uint32_t spins = 0;
            uint64_t t0 = 0;
            for (;;) {
              bool ready = true;
#pragma unroll
              for (int mm = 0; mm < 10; mm++) {
                const uint32_t c = lane + 32u * mm;
                long long bits = 0;
                if (c < G2) bits = ld_relaxed_s64(pp + c);
                ready = ready && bits != kFNotYet;
                v[mm] = __longlong_as_double(bits);
              }
              for (uint32_t c = lane + 320u; c < G2; c += 32)  // more than 160 SMs (not B200)
                ready = ready && ld_relaxed_s64(pp + c) != kFNotYet;
              if (__all_sync(0xffffffffu, ready)) break;
              if (((++spins) & 0xFFu) == 0) {
                int dead = (lane == 0 && watch.expired(t0, 32)) ? 1 : 0;
                if (__shfl_sync(0xffffffffu, dead, 0)) return;
              }
            }
          }
          if (jl == first) stamp(s, 3);
          double e = v[0];
#pragma unroll
          for (int mm = 1; mm < 10; mm++) e += v[mm];
          for (uint32_t c = lane + 320u; c < G2; c += 32) e += __ldcg(pp + c);
          // hand the mailboxes back (all-ones) for slab s + kFASlots
#pragma unroll
          for (int mm = 0; mm < 10; mm++) {
            const uint32_t c = lane + 32u * mm;
            if (c < G2) st_relaxed_f64(pp + c, __longlong_as_double(kFNotYet));
          }
          for (uint32_t c = lane + 320u; c < G2; c += 32)
            st_relaxed_f64(pp + c, __longlong_as_double(kFNotYet));
          double mx = 0.0;
          double2 ms;
          if (jl == first) {
            mx = mx_pre;
            ms = ms_pre;
          } else {
            if (A.mx_tiles)
              for (uint32_t tt = lane; tt < A.mx_tiles; tt += 32)
                mx += __ldcg(A.mxv + (uint64_t)tt * A.nsnps + j);
            ms = A.scale[j];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_xor_sync(0xffffffffu, e, o);
            mx += __shfl_xor_sync(0xffffffffu, mx, o);
          }
          {  // every lane computes the same a_j (e and mx are warp-uniform after the shuffles)
            e *= scx.delta;
            const double t = ((e - 3.0 * mx) - ms.x * (scx.sum - mx)) * ms.y;
            const bool dead = ms.y == 0.0;  // monomorphic SNP: zero column (data.cpp:300)
            double a = dead ? 0.0 : t * ms.y;
            if (a != a) a = __longlong_as_double(0x7FF8000000000000LL);  // never the all-ones word
            const double b = dead ? 0.0 : ms.x * a;
            if (lane < kFReplicas) st_relaxed_f64(A.a_rep + (uint64_t)lane * A.rep_stride + j, a);
            if (lane == 0) {
              A.corr_out[j] = b - 3.0 * a;
              A.a_out[j] = a;
            }
          }
        }
        stamp(s, 4);
        if (lane == 0) misc[2] = s + 1;
      }
      return;
    }
  }

  if (warp >= kFP1Warp0) {
    // ------------------------ first-half consumers (4 warps) ------------------------
    reg_dec<kFRegP1>();
    const int w = (warp - kFP1Warp0) & 3, wg = (warp - kFP1Warp0) >> 2;
    const int g = lane >> 2, q = lane & 3;
    const int rho = (g >> 1) | ((g & 1) << 2);
    uint32_t roff[2][2];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
      for (int hf = 0; hf < 2; hf++) roff[t][hf] = (uint32_t)(w * 32 + t * 16 + hf * 8 + rho) * 128u;
    const uint32_t rx = (uint32_t)(rho & 7);
    const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
    if (!wait_bar(bars + 8 * B_XS, 0, watch, 41)) return;
    // Stage release is lagged by one tile: the slot of tile t is handed back after the wait
    // for tile t+1.  By then every shared-memory read of tile t has been consumed by an IMMA
    // issued before that wait loop, so the producer's next TMA write cannot overtake a read
    // -- without the per-tile fence.proxy.async + membar k_imma_gemv_tma needs.
    uint32_t prev_slot = kFStages1;  // none yet
    for (uint32_t s = 0; s < nslabs; s++) {
      // This thread's four mailboxes of ring slot s % kFASlots (rows rho, rho+8 of both m16
      // tiles) were last used by slab s - kFASlots and must be empty again.  The loads are
      // issued now and only looked at after the slab's MMAs.
      double* const po = A.part + (uint64_t)(s % kFASlots) * kFRows * A.gpad + kFP1Groups * cta + wg;
      long long old[4] = {kFNotYet, kFNotYet, kFNotYet, kFNotYet};
      if (q == 0) {
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const uint32_t rr = (uint32_t)(w * 32 + t * 16 + rho);
          old[2 * t] = ld_relaxed_s64(po + (uint64_t)rr * A.gpad);
          old[2 * t + 1] = ld_relaxed_s64(po + (uint64_t)(rr + 8) * A.gpad);
        }
      }
      int acc[2][2][2][4] = {};  // [u][tile][chain]: 8 independent IMMA chains
      // the two first-half warpgroups take alternate tiles of the CTA's tile stream
      const uint32_t it0 = s * nk;
      for (uint32_t k = kFP1Groups == 1 ? 0u : ((it0 + wg) & 1u); k < nk; k += kFP1Groups) {
        const uint32_t it = it0 + k, slot = it % kFStages1, round = it / kFStages1;
        if (!wait_bar(bars + 8 * (B_FULL1 + slot), round & 1u, watch, 42)) return;
        __syncwarp();
        if (lane == 0 && prev_slot < (uint32_t)kFStages1) mbar_arrive(bars + 8 * (B_EMPTY1 + prev_slot));
        prev_slot = slot;
        const uint32_t tile = base + kFOffRing1 + slot * kFTileBytes;
        const uint32_t sl = base + kFOffXs + k * 4096u;
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const uint32_t chunk = ((uint32_t)(u * 4 + q) ^ rx) << 4;
          uint4 wv[2][2];
#pragma unroll
          for (int t = 0; t < 2; t++)
#pragma unroll
            for (int hf = 0; hf < 2; hf++)
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(wv[t][hf].x), "=r"(wv[t][hf].y), "=r"(wv[t][hf].z),
                             "=r"(wv[t][hf].w)
                           : "r"(tile + roff[t][hf] + chunk));
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int wl = u * 16 + q * 4 + j;
            uint4 bv;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(bv.x), "=r"(bv.y), "=r"(bv.z), "=r"(bv.w)
                         : "r"(sl + (uint32_t)slice_slot(wl, g) * 16u));
#pragma unroll
            for (int t = 0; t < 2; t++) {
              const uint32_t xa =
                  j == 0 ? wv[t][0].x : j == 1 ? wv[t][0].y : j == 2 ? wv[t][0].z : wv[t][0].w;
              const uint32_t xb =
                  j == 0 ? wv[t][1].x : j == 1 ? wv[t][1].y : j == 2 ? wv[t][1].z : wv[t][1].w;
              mma_u8s8(acc[u][t][0], xa & 0x03030303u, xb & 0x03030303u, xa & 0x0F0F0F0Fu,
                       xb & 0x0F0F0F0Fu, bv.x, bv.y);
              mma_u8s8(acc[u][t][1], xa & 0x3F3F3F3Fu, xb & 0x3F3F3F3Fu, xa, xb, bv.z, bv.w);
            }
          }
        }
      }
      if (w == 0 && wg == 0 && lane == 0) stamp(s, 2);
      // the CTA's partial E of the slab's SNPs -> mailboxes part[s % ring][row][cta]
#pragma unroll
      for (int t = 0; t < 2; t++) {
        int sum4[4];  // exact: |sum| <= 3584 x 255 x 127
#pragma unroll
        for (int c = 0; c < 4; c++)
          sum4[c] = (acc[0][t][0][c] + acc[0][t][1][c]) + (acc[1][t][0][c] + acc[1][t][1][c]);
        double ra = (double)sum4[0] * w0 + (double)sum4[1] * w1;
        double rb = (double)sum4[2] * w0 + (double)sum4[3] * w1;
        ra += __shfl_xor_sync(0xffffffffu, ra, 1);
        rb += __shfl_xor_sync(0xffffffffu, rb, 1);
        ra += __shfl_xor_sync(0xffffffffu, ra, 2);
        rb += __shfl_xor_sync(0xffffffffu, rb, 2);
        if (q == 0) {
#pragma unroll
          for (int hf = 0; hf < 2; hf++) {
            const uint32_t rr = (uint32_t)(w * 32 + t * 16 + hf * 8 + rho);
            if (s * kFRows + rr >= A.nsnps) continue;  // rows behind the last SNP have no reducer
            double* const slotp = po + (uint64_t)rr * A.gpad;
            long long o = old[2 * t + hf];
            uint32_t spins = 0;
            uint64_t t0 = 0;
            while (o != kFNotYet) {  // rare: the exchange ring is full
              if (((++spins) & 0xFFu) == 0 && watch.expired(t0, 43)) break;
              o = ld_relaxed_s64(slotp);
            }
            st_relaxed_f64(slotp, hf == 0 ? ra : rb);
          }
        }
      }
      if (w == 0 && wg == 0 && lane == 0) {
        misc[1] = s + 1;
        stamp(s, 1);
      }
    }
    return;
  }

  // --------------------------- second-half consumers (8 warps) ---------------------------
  reg_inc<kFRegP2>();
  {
    const int w = warp - kFP2Warp0;
    const int g = lane >> 2, q = lane & 3;
    const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
    int acc[kFSpc][4][4] = {};
    int cur_ex = kFExZero;
    bool drained = false;
    uint32_t since = 0;
    // drain: F (or the running FP64 sums) += step * accumulators; zero them
    auto drain = [&](double* __restrict__ dst) {
      double delta;
      if (cur_ex == kFExNan) delta = nan("");
      else if (cur_ex == kFExZero) delta = 0.0;
      else delta = ldexp(1.0, cur_ex - kSliceBits);
#pragma unroll
      for (int k = 0; k < kFSpc; k++) {
        if ((uint32_t)k < nk) {
          const uint64_t byte_a = (uint64_t)(stripe0 + k) * 128u + w * 16 + g;
#pragma unroll
          for (int f = 0; f < 4; f++) {
            const double sf = ldexp(1.0, -2 * f);  // the field carried e * 4^f
            // accumulator f holds the fields <= f (cumulative masks): difference = field f,
            // exact in int32 (|field sum| < 65536 x 192 x 64)
            int fd[4];
#pragma unroll
            for (int c = 0; c < 4; c++) fd[c] = f == 0 ? acc[k][0][c] : acc[k][f][c] - acc[k][f - 1][c];
            double ra = ((double)fd[0] * w0 + (double)fd[1] * w1) * sf;
            double rb = ((double)fd[2] * w0 + (double)fd[3] * w1) * sf;
            ra += __shfl_xor_sync(0xffffffffu, ra, 1);
            rb += __shfl_xor_sync(0xffffffffu, rb, 1);
            ra += __shfl_xor_sync(0xffffffffu, ra, 2);
            rb += __shfl_xor_sync(0xffffffffu, rb, 2);
            if (q == 0) {
              const uint64_t ia = byte_a * 4 + f, ib = (byte_a + 8) * 4 + f;
              if (ia < A.n) dst[ia] = ra * delta + (drained ? A.ybuf[ia] : 0.0);
              if (ib < A.n) dst[ib] = rb * delta + (drained ? A.ybuf[ib] : 0.0);
            }
          }
#pragma unroll
          for (int f = 0; f < 4; f++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[k][f][c] = 0;
        }
      }
    };
    uint32_t slot = 0, round = 0;
    uint32_t prev_slot = kFStages2, prev_aslot = kFASlots;  // lagged releases (see the first half)
    for (uint32_t s = 0; s < nslabs; s++) {
      const uint32_t aslot = s % kFASlots, around = s / kFASlots;
      if (A.dbg_mode != 1 && !wait_bar(bars + 8 * (B_AFULL + aslot), around & 1u, watch, 51)) return;
      const uint32_t asl = base + kFOffA + aslot * kFASlotBytes;
      int ex_s;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ex_s) : "r"(asl));
      if (ex_s != cur_ex || since >= (uint32_t)kFFlushSlabs) {
        if (cur_ex != kFExZero) {
          drain(A.ybuf);
          drained = true;
        }
        cur_ex = ex_s;
        since = 0;
      }
      since++;
#pragma unroll
      for (int k = 0; k < kFSpc; k++) {
        if ((uint32_t)k < nk) {
          if (!wait_bar(bars + 8 * (B_FULL2 + slot), round & 1u, watch, 52)) return;
          __syncwarp();
          if (lane == 0) {
            if (prev_slot < (uint32_t)kFStages2) mbar_arrive(bars + 8 * (B_EMPTY2 + prev_slot));
            // the previous slab is done: frees its a-slice slot, lifts the first half's throttle
            if (prev_aslot < (uint32_t)kFASlots) mbar_arrive(bars + 8 * (B_AEMPTY + prev_aslot));
          }
          prev_slot = slot;
          prev_aslot = kFASlots;
          const uint32_t tile = base + kFOffRing2 + slot * kFTileBytes;
#pragma unroll
          for (int ks = 0; ks < kFRows / 32; ks++) {
            const uint32_t row = (uint32_t)(ks * 32 + lane);
            const uint32_t addr = tile + row * 128u + ((((uint32_t)w) ^ (row & 7u)) << 4);
            uint32_t a0, a1, a2, a3, b0, b1;
            asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                         : "r"(addr));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                         : "=r"(b0), "=r"(b1)
                         : "r"(asl + 16u + (uint32_t)(((ks * 8 + g) * 4 + q) * 8)));
            mma_u8s8(acc[k][0], a0 & 0x03030303u, a1 & 0x03030303u, a2 & 0x03030303u,
                     a3 & 0x03030303u, b0, b1);
            // cumulative masks: accumulator f holds sum_{g <= f} 4^g e_g a (12 LOP3 per 4 IMMAs
            // instead of 16); the drain takes differences
            mma_u8s8(acc[k][1], a0 & 0x0F0F0F0Fu, a1 & 0x0F0F0F0Fu, a2 & 0x0F0F0F0Fu,
                     a3 & 0x0F0F0F0Fu, b0, b1);
            mma_u8s8(acc[k][2], a0 & 0x3F3F3F3Fu, a1 & 0x3F3F3F3Fu, a2 & 0x3F3F3F3Fu,
                     a3 & 0x3F3F3F3Fu, b0, b1);
            mma_u8s8(acc[k][3], a0, a1, a2, a3, b0, b1);
          }
          if (++slot == kFStages2) {
            slot = 0;
            round++;
          }
        }
      }
      prev_aslot = aslot;  // released after the next tile wait
      if (w == 0 && lane == 0) stamp(s, 7);
    }
    drain(A.f_out);
  }
}

// Sb = sum_j mean_j a_j over the live SNPs (fixed order) -> VecScale{sum = Sb, delta = 1}
// for k_finalize_prod
__global__ void __launch_bounds__(1024)
k_fused_sum_b(const double* __restrict__ a, const double2* __restrict__ scale, uint32_t nsnps,
              VecScale* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (uint32_t j = threadIdx.x; j < nsnps; j += 1024) {
    const double2 ms = scale[j];
    if (ms.y != 0.0) s += ms.x * a[j];  // dead SNPs have a = 0 and possibly a NaN mean
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      VecScale r;
      r.sum = s;
      r.delta = 1.0;
      r.ex = 0;
      r.pad = 0;
      *out = r;
    }
  }
}

}  // namespace fpb
