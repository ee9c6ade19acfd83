// fpb_umma.cuh -- the k-column block contraction on the 5th-generation tensor cores.
//
// The block variants of the operator (perform_op_mat / perform_op_multi, crossprod2, prod3:
// svdwide.cpp:71-118, 157-188, 229-275, 312-343) contract the packed matrix against b vectors at
// once.  With b vectors the B operand has N = 8 b columns (8 int8 digit slices per vector,
// fpb_imma.cuh), a shape tcgen05.mma kind::i8 handles at full rate, and the decode of the packed
// genotypes is shared by all b vectors:
//
//   HBM --TMA--> shared memory tile [128 rows x 128 B, SWIZZLE_128B]
//       --decoder warps: LDS / LDSM.trans, 3 LOP3 per packed word (cumulative field masks),
//         tcgen05.st--> A operand in TMEM (u8, 128 lanes x 8 columns per K = 32 block)
//   digit slices of the b vectors --bulk copy--> shared memory (s8, K-major, no swizzle) = B
//   one thread issues tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = 8 b, K = 32),
//   D (int32) accumulates in TMEM; the epilogue reads it back with tcgen05.ld, recombines the
//   slices in FP64 (sum_s 128^s D_s, every term exact) and writes per-split partial sums.
//
// The integer sums are the ones the mma.sync kernels of fpb_imma.cuh compute (same digits, same
// cumulative masks, same field separation), so the two paths agree to the FP64 recombination
// order of the split partials.
//
// Verified conventions (tools/tcgen05_probe.cu on B200): A in TMEM, lane = matrix row, column c
// holds K bytes 4c..4c+3 (little endian); B in shared memory, K-major SWIZZLE_NONE, 8 x 16-byte
// core matrices, LBO = 128 (next 16 K-bytes), SBO = 256 (next 8 N-rows); instruction descriptor
// bits per cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fpb_fused.cuh"  // Watch, wait_bar (bounded waits), mbar_try
#include "fpb_imma.cuh"

namespace fpb {

constexpr int kUBoxRows = 128;                    // rows of one TMA box = M of one MMA
constexpr int kUBoxBytes = kUBoxRows * 128;       // 16 KB
constexpr int kUSmemBytes = 232448;               // whole opt-in shared memory: one CTA per SM
constexpr uint32_t kUErrTimeout = 0x55AA0000u;

// kind::i8 instruction descriptor: D = S32, A = u8, B = s8, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// shared-memory descriptor of a K-major SWIZZLE_NONE operand whose K = 32 block is
// [n / 8][2 k-halves][8 rows][16 B]: LBO = 128, SBO = 256, version 1
__device__ __forceinline__ uint64_t umma_bdesc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) |
         ((uint64_t)(256u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&d)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15,%16};" ::"r"(taddr),
      "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
      "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&d)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15,%16};" ::"r"(taddr),
      "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
      "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&d)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, int (&d)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]),
                 "=r"(d[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// sum_s 128^s d_s (every term exact in FP64)
__device__ __forceinline__ double recombine8(const double (&d)[8]) {
  double r = 0.0;
#pragma unroll
  for (int s = 7; s >= 0; s--) r += d[s] * (double)(1ull << (7 * s));
  return r;
}

// ---------------------------------------------------------------------------
// Digit slices, first half (contraction over individuals).  One thread per K = 32 block kb
// (8 packed bytes = 32 individuals) of vector v.  Block kb of the B operand is NV x 256 bytes:
//   [v][k-half jl][slice s][16 B: 4 p + i] = p_p digit of slice s for packed byte 8 kb + 4 jl + i,
// where p_p = d_p - d_{p+1} are the differences the cumulative masks need (fpb_imma.cuh) and
// d_f is digit s of rint(x_{4 byte + f} / (4^f delta)).  Matches the decoder's A layout: the
// packed word j of a 16-byte group becomes the 4 TMEM columns [word & 0x03.., & 0x0F.., & 0x3F..,
// word].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_slice_umma_i(const double* __restrict__ v, uint64_t n, uint32_t nkb, uint32_t nv, uint32_t vi,
               const double* __restrict__ pmax, const double* __restrict__ psum, uint32_t nparts,
               VecScale* __restrict__ sc_out, uint4* __restrict__ out) {
  const VecScale sc = scale_from_partials(pmax, psum, nparts);
  if (blockIdx.x == 0 && threadIdx.x == 0) *sc_out = sc;
  const uint32_t kb = blockIdx.x * blockDim.x + threadIdx.x;
  if (kb >= nkb) return;
  const int ex = sc.ex;
  const bool live = sc.delta > 0.0;
  uint4* o = out + ((uint64_t)kb * nv + vi) * 16;  // 256 bytes
#pragma unroll
  for (int jl = 0; jl < 2; jl++) {
    uint32_t wd[8][4] = {};  // [slice][p], bytes over i
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int dg[4][8];
#pragma unroll
      for (int f = 0; f < 4; f++) {
        const uint64_t idx = ((uint64_t)kb * 8 + jl * 4 + i) * 4 + f;
        long long q = 0;
        if (live && idx < n) q = __double2ll_rn(ldexp(v[idx], kSliceBits - ex - 2 * f));
#pragma unroll
        for (int s = 0; s < 8; s++) {
          long long d = (s < 7) ? (((q + 64) & 127) - 64) : q;
          q = (q - d) >> 7;
          dg[f][s] = (int)d;
        }
      }
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int pd = dg[p][s] - (p < 3 ? dg[p + 1][s] : 0);
          wd[s][p] |= ((uint32_t)(pd & 0xFF)) << (8 * i);
        }
    }
#pragma unroll
    for (int s = 0; s < 8; s++) o[jl * 8 + s] = make_uint4(wd[s][0], wd[s][1], wd[s][2], wd[s][3]);
  }
}

// ---------------------------------------------------------------------------
// Digit slices, second half (contraction over SNPs).  One thread per K = 32 block (32 SNPs) of
// vector v; plain balanced digits of rint(a_j / delta).  The K order inside a block follows the
// LDSM.MT1616 fragment as the decoder stores it (a0, a2, a1, a3 with tcgen05.st.16x256b):
// TMEM column c of a row holds SNPs 4 (c / 2) + i for even c and 16 + 4 (c / 2) + i for odd c.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_slice_umma_j(const double* __restrict__ v, uint64_t n, uint32_t nkb, uint32_t nv, uint32_t vi,
               const double* __restrict__ pmax, const double* __restrict__ psum, uint32_t nparts,
               VecScale* __restrict__ sc_out, uint4* __restrict__ out) {
  const VecScale sc = scale_from_partials(pmax, psum, nparts);
  if (blockIdx.x == 0 && threadIdx.x == 0) *sc_out = sc;
  const uint32_t kb = blockIdx.x * blockDim.x + threadIdx.x;
  if (kb >= nkb) return;
  const int ex = sc.ex;
  const bool live = sc.delta > 0.0;
  uint4* o = out + ((uint64_t)kb * nv + vi) * 16;
#pragma unroll
  for (int jl = 0; jl < 2; jl++) {
    uint32_t wd[8][4] = {};  // [slice][column within the k-half], bytes over i
#pragma unroll
    for (int cc = 0; cc < 4; cc++) {
      const int c = jl * 4 + cc;  // TMEM column 0..7
      const int snp0 = (c & 1) ? 16 + 4 * (c >> 1) : 4 * (c >> 1);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint64_t idx = (uint64_t)kb * 32 + snp0 + i;
        long long q = 0;
        if (live && idx < n) q = __double2ll_rn(ldexp(v[idx], kSliceBits - ex));
#pragma unroll
        for (int s = 0; s < 8; s++) {
          long long d = (s < 7) ? (((q + 64) & 127) - 64) : q;
          q = (q - d) >> 7;
          wd[s][cc] |= ((uint32_t)(d & 0xFF)) << (8 * i);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 8; s++) o[jl * 8 + s] = make_uint4(wd[s][0], wd[s][1], wd[s][2], wd[s][3]);
  }
}

// ---------------------------------------------------------------------------
// First half: E[j][v] = sum_i e_ij x_i^(v) for NV vectors.
// CTA = RG row groups of 128 SNP rows x one split of the 128-byte column stages.  Row group r is
// an independent lane: decoder warpgroup r (thread = row = TMEM lane) and issuer warp r with its
// own accumulators D[r][chain] and its own ring of quarter-box staging slots (32 TMEM columns =
// 2 of the 8 sixteen-byte chunks of the row = 4 MMAs), so every mbarrier has one sequential
// waiter.  (A ring shared between warpgroups lets a fast one test a phase parity two phases ahead
// of the barrier, which try_wait.parity cannot tell from "done".)  Dependent MMAs into one
// accumulator are latency-bound at small N, hence NCH interleaved accumulator chains per lane.
// Warps: 4 RG decoders, RG issuers, 1 TMA producer.
// out[v * vstride + split * sstride + row] = sum_s 128^s D[row][8 v + s].
// ---------------------------------------------------------------------------
template <int NV, int RG, int NCH>
struct UmmaXtCfg {
  static constexpr int N = 8 * NV;
  static constexpr int DCols = RG * NCH * N;
  static constexpr int NTLmax = ((512 - DCols) / 32) / RG;   // quarter-box slots per lane
  static constexpr int NTL = NTLmax > 4 ? 4 : NTLmax;
  static constexpr int BBytes = 4096 * NV;  // digit slices of one 128-byte stage (16 K-blocks)
  static constexpr int NB = 2;
  static constexpr int NALmax = ((kUSmemBytes - 2048 - NB * BBytes) / kUBoxBytes) / RG;
  static constexpr int NAL = NALmax > 3 ? 3 : NALmax;       // boxes in flight per lane
  static constexpr int Threads = (5 * RG + 1) * 32;
  static_assert(NTL >= 2, "TMEM: not enough columns for two staging slots per lane");
  static_assert(NAL >= 2, "shared memory: A ring too small");
  static_assert(16 % NCH == 0, "chains must divide the 16 K-blocks of a box");
};

template <int NV, int RG, int NCH>
__global__ void __launch_bounds__((5 * RG + 1) * 32, 1)
k_umma_xt(const __grid_constant__ TmaDesc tmap, uint32_t R, const uint8_t* __restrict__ S,
          uint32_t nstages, uint32_t stages_per_split, double* __restrict__ out, uint64_t vstride,
          uint64_t sstride, uint32_t* __restrict__ gerr) {
  using C = UmmaXtCfg<NV, RG, NCH>;
  constexpr int N = C::N, NTL = C::NTL, NAL = C::NAL, NB = C::NB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = base, b_ring = base + RG * NAL * kUBoxBytes;
  const uint32_t bars = b_ring + NB * C::BBytes;
  constexpr int A_FULL = 0, A_EMPTY = RG * NAL, B_FULL = 2 * RG * NAL, B_EMPTY = B_FULL + NB,
                T_FULL = B_EMPTY + NB, T_EMPTY = T_FULL + RG * NTL, D_FULL = T_EMPTY + RG * NTL,
                NBARS = D_FULL + RG;
  const uint32_t misc = bars + 8 * NBARS;  // [0] tmem base, [1] abort flag
  volatile uint32_t* misc_p =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (misc - smem_u32(smem_raw)));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s_begin = blockIdx.y * stages_per_split;
  const uint32_t s_end = min(nstages, s_begin + stages_per_split);
  const uint32_t nst = s_end > s_begin ? s_end - s_begin : 0;
  const uint32_t row0 = blockIdx.x * (RG * kUBoxRows);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  if (tid == 0) {
    for (int i = 0; i < RG * NAL; i++) {
      mbar_init(bar(A_FULL + i), 1);
      mbar_init(bar(A_EMPTY + i), 4);  // one arrive per decoder warp of the lane
    }
    for (int i = 0; i < NB; i++) {
      mbar_init(bar(B_FULL + i), 1);
      mbar_init(bar(B_EMPTY + i), RG);  // one tcgen05.commit per issuer
    }
    for (int i = 0; i < RG * NTL; i++) {
      mbar_init(bar(T_FULL + i), 4);
      mbar_init(bar(T_EMPTY + i), 1);
    }
    for (int i = 0; i < RG; i++) mbar_init(bar(D_FULL + i), 1);
    misc_p[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5 * RG) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(misc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc_p[0];
  const uint32_t tA = tmem + (uint32_t)C::DCols;  // staging: lane r, slot t at (r NTL + t) 32
  const Watch watch{misc_p + 1, gerr};

  if (warp == 5 * RG) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0 && nst > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      bool ok = true;
      for (uint32_t c = 0; c < nst && ok; c++) {
        const uint32_t bs = c % NB, al = c % NAL;
        if (c >= NB) ok = wait_bar(bar(B_EMPTY + bs), ((c / NB) - 1) & 1, watch, kUErrTimeout | 1);
        if (!ok) break;
        mbar_expect_tx(bar(B_FULL + bs), C::BBytes);
        bulk_load(b_ring + bs * C::BBytes, S + (uint64_t)(s_begin + c) * C::BBytes, C::BBytes,
                  bar(B_FULL + bs), pol_keep);
#pragma unroll 1
        for (uint32_t r = 0; r < RG; r++) {
          const uint32_t as = r * NAL + al;
          if (c >= NAL) ok = wait_bar(bar(A_EMPTY + as), ((c / NAL) - 1) & 1, watch, kUErrTimeout | 2);
          if (!ok) break;
          mbar_expect_tx(bar(A_FULL + as), kUBoxBytes);
          tma_load_2d(a_ring + as * kUBoxBytes, &tmap, (int)((s_begin + c) * 128),
                      (int)(row0 + r * kUBoxRows), bar(A_FULL + as), pol_stream);
        }
      }
    }
  } else if (warp >= 4 * RG) {
    // ------------------------- MMA issuer of lane r ---------------------------
    const uint32_t r = (uint32_t)(warp - 4 * RG);
    if (lane == 0 && nst > 0) {
      const uint32_t idesc = umma_idesc(N);
      const uint32_t tD = tmem + r * (NCH * N), tS = tA + r * (NTL * 32);
      bool ok = true;
      for (uint32_t c = 0; c < nst && ok; c++) {
        const uint32_t bs = c % NB;
        ok = wait_bar(bar(B_FULL + bs), (c / NB) & 1, watch, kUErrTimeout | 3);
        const uint64_t bd0 = umma_bdesc(b_ring + bs * C::BBytes);
#pragma unroll 1
        for (uint32_t qd = 0; qd < 4 && ok; qd++) {
          const uint32_t m = c * 4 + qd, ts = m % NTL;
          ok = wait_bar(bar(T_FULL + r * NTL + ts), (m / NTL) & 1, watch, kUErrTimeout | 4);
          if (!ok) break;
          tc_fence_after();
#pragma unroll
          for (uint32_t k4 = 0; k4 < 4; k4++) {
            const uint32_t kc = qd * 4 + k4, ch = kc % NCH;
            umma_i8(tD + ch * N, tS + ts * 32 + k4 * 8, bd0 + (uint64_t)((kc * NV * 256) >> 4), idesc,
                    (c > 0 || kc >= NCH) ? 1u : 0u);
          }
          umma_commit(bar(T_EMPTY + r * NTL + ts));
        }
        if (ok) umma_commit(bar(B_EMPTY + bs));
      }
      if (ok) umma_commit(bar(D_FULL + r));
    }
  } else {
    // ----------------------------- decoders of lane r --------------------------
    const uint32_t r = (uint32_t)(warp >> 2), q = (uint32_t)(warp & 3);
    const uint32_t lane_base = (q * 32u) << 16;
    const uint32_t rib = q * 32u + (uint32_t)lane;  // row in box = TMEM lane
    const uint32_t roff = rib * 128u, rx = rib & 7u;
    const uint32_t tS = tA + r * (NTL * 32) + lane_base;
    bool ok = true;
    for (uint32_t c = 0; c < nst && ok; c++) {
      const uint32_t as = r * NAL + (c % NAL);
      ok = wait_bar(bar(A_FULL + as), (c / NAL) & 1, watch, kUErrTimeout | 5);
      if (!ok) break;
      const uint32_t tile = a_ring + as * kUBoxBytes + roff;
      uint4 w[8];
#pragma unroll
      for (uint32_t u = 0; u < 8; u++)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w)
                     : "r"(tile + ((u ^ rx) << 4)));
#pragma unroll
      for (uint32_t qd = 0; qd < 4; qd++) {
        const uint32_t m = c * 4 + qd, ts = m % NTL;
        if (m >= NTL && ok)
          ok = wait_bar(bar(T_EMPTY + r * NTL + ts), ((m / NTL) - 1) & 1, watch, kUErrTimeout | 6);
        tc_fence_after();
#pragma unroll
        for (uint32_t h2 = 0; h2 < 2; h2++) {
          const uint4 ww = w[qd * 2 + h2];
          const uint32_t x[4] = {ww.x, ww.y, ww.z, ww.w};
          uint32_t d[16];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            d[4 * j + 0] = x[j] & 0x03030303u;
            d[4 * j + 1] = x[j] & 0x0F0F0F0Fu;
            d[4 * j + 2] = x[j] & 0x3F3F3F3Fu;
            d[4 * j + 3] = x[j];
          }
          tmem_st_32x32b_x16(tS + ts * 32 + h2 * 16, d);
        }
        if (qd == 3) {
          // the box has been consumed into registers (the stores depend on every load)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(A_EMPTY + as));
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(T_FULL + r * NTL + ts));
      }
    }
    // -------------------------------- epilogue --------------------------------
    if (nst > 0 && ok) ok = wait_bar(bar(D_FULL + r), 0, watch, kUErrTimeout | 7);
    tc_fence_after();
    const uint32_t row = row0 + r * kUBoxRows + rib;
#pragma unroll 1
    for (int v = 0; v < NV; v++) {
      double dd[8] = {};
      if (nst > 0 && ok) {
#pragma unroll
        for (int ch = 0; ch < NCH; ch++) {
          int di[8];
          tmem_ld_32x32b_x8(tmem + r * (NCH * N) + ch * N + v * 8 + lane_base, di);
#pragma unroll
          for (int s = 0; s < 8; s++) dd[s] += (double)di[s];
        }
      }
      if (row < R) out[(uint64_t)v * vstride + (uint64_t)blockIdx.y * sstride + row] = recombine8(dd);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5 * RG)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------
// Second half: F[i][v] = sum_j e_ij a_j^(v) from the same SNP-major copy.
// CTA = one 128-byte column stripe (512 individuals) x one split of the 128-row SNP boxes.
// A box is decoded by 8 warps (warp = one 16-byte chunk of the stripe): LDSM.MT1616 transposes
// bytes, so lane = packed byte column and K = SNPs.  The four cumulative mask planes of a byte
// column are separate MMA rows (plane pair pp = p >> 1 selects the accumulator, p & 1 the upper
// or lower 16 lanes of the warp's TMEM quadrant), and the epilogue differences them into the four
// individuals of the byte.
// TMEM: D[hp] (hp = 2 h + pp, h = half of the stripe: 64 byte columns) at columns hp N: four
// independent accumulators, one issuer warp each (NISS = 4) -- a box is one handshake and 4 MMAs
// per issuer.  Staging: a ring of whole-box slots (128 columns: (hp 4 + kc) 8 for the K = 32 block
// kc of the box) with one sequential producer (the 8 decoder warps move box by box together).
// out[v * vstride + split * sstride + individual].
// ---------------------------------------------------------------------------
template <int NV>
struct UmmaXvCfg {
  static constexpr int N = 8 * NV;
  static constexpr int DCols = 4 * N;
  static constexpr int NTmax = (512 - DCols) / 128;  // whole-box staging slots
  static constexpr int NT = NTmax > 3 ? 3 : NTmax;
  static constexpr int BBytes = 1024 * NV;  // digit slices of one 128-row box (4 K-blocks)
  static constexpr int NB = 4;
  static constexpr int NAmax = (kUSmemBytes - 2048 - NB * BBytes) / kUBoxBytes;
  static constexpr int NA = NAmax > 8 ? 8 : NAmax;  // boxes in flight
  static_assert(NT >= 2, "TMEM: not enough columns for two staging slots");
};

template <int NV, int NISS>
__global__ void __launch_bounds__((8 + NISS + 1) * 32, 1)
k_umma_xv(const __grid_constant__ TmaDesc tmap, uint32_t Cn /* output length */,
          const uint8_t* __restrict__ S, uint32_t nboxes_total, uint32_t boxes_per_split,
          double* __restrict__ out, uint64_t vstride, uint64_t sstride, uint32_t* __restrict__ gerr) {
  using C = UmmaXvCfg<NV>;
  constexpr int N = C::N, NT = C::NT, NA = C::NA, NB = C::NB;
  constexpr int NW = 8;  // decoder warps
  static_assert(4 % NISS == 0, "issuers split the four accumulators");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_ring = base, b_ring = base + NA * kUBoxBytes;
  const uint32_t bars = b_ring + NB * C::BBytes;
  constexpr int A_FULL = 0, A_EMPTY = NA, B_FULL = 2 * NA, B_EMPTY = B_FULL + NB,
                T_FULL = B_EMPTY + NB, T_EMPTY = T_FULL + NT, D_FULL = T_EMPTY + NT,
                NBARS = D_FULL + 1;
  const uint32_t misc = bars + 8 * NBARS;
  volatile uint32_t* misc_p =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (misc - smem_u32(smem_raw)));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t_begin = blockIdx.y * boxes_per_split;
  const uint32_t t_end = min(nboxes_total, t_begin + boxes_per_split);
  const uint32_t nbx = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t xbyte0 = blockIdx.x * 128u;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  if (tid == 0) {
    for (int i = 0; i < NA; i++) {
      mbar_init(bar(A_FULL + i), 1);
      mbar_init(bar(A_EMPTY + i), NW);
    }
    for (int i = 0; i < NB; i++) {
      mbar_init(bar(B_FULL + i), 1);
      mbar_init(bar(B_EMPTY + i), NISS);
    }
    for (int i = 0; i < NT; i++) {
      mbar_init(bar(T_FULL + i), NW);
      mbar_init(bar(T_EMPTY + i), NISS);
    }
    mbar_init(bar(D_FULL), NISS);
    misc_p[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == NW + NISS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(misc)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc_p[0];
  const uint32_t tA = tmem + (uint32_t)C::DCols;  // staging slot t at t 128
  const Watch watch{misc_p + 1, gerr};

  if (warp == NW + NISS) {
    if (lane == 0 && nbx > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      bool ok = true;
      for (uint32_t i = 0; i < nbx && ok; i++) {
        const uint32_t bs = i % NB, as = i % NA;
        if (i >= NB) ok = wait_bar(bar(B_EMPTY + bs), ((i / NB) - 1) & 1, watch, kUErrTimeout | 0x11);
        if (!ok) break;
        mbar_expect_tx(bar(B_FULL + bs), C::BBytes);
        bulk_load(b_ring + bs * C::BBytes, S + (uint64_t)(t_begin + i) * C::BBytes, C::BBytes,
                  bar(B_FULL + bs), pol_keep);
        if (i >= NA) ok = wait_bar(bar(A_EMPTY + as), ((i / NA) - 1) & 1, watch, kUErrTimeout | 0x12);
        if (!ok) break;
        mbar_expect_tx(bar(A_FULL + as), kUBoxBytes);
        tma_load_2d(a_ring + as * kUBoxBytes, &tmap, (int)xbyte0, (int)((t_begin + i) * kUBoxRows),
                    bar(A_FULL + as), pol_stream);
      }
    }
  } else if (warp >= NW) {
    const uint32_t iss = (uint32_t)(warp - NW);
    constexpr uint32_t HPI = 4 / NISS;  // accumulators per issuer
    if (lane == 0 && nbx > 0) {
      const uint32_t idesc = umma_idesc(N);
      bool ok = true;
      for (uint32_t i = 0; i < nbx && ok; i++) {
        const uint32_t bs = i % NB, ts = i % NT;
        ok = wait_bar(bar(B_FULL + bs), (i / NB) & 1, watch, kUErrTimeout | 0x13);
        if (ok) ok = wait_bar(bar(T_FULL + ts), (i / NT) & 1, watch, kUErrTimeout | 0x14);
        if (!ok) break;
        tc_fence_after();
        const uint64_t bd0 = umma_bdesc(b_ring + bs * C::BBytes);
        const uint32_t slot = tA + ts * 128;
#pragma unroll
        for (uint32_t hh = 0; hh < HPI; hh++)
#pragma unroll
          for (uint32_t kc = 0; kc < 4; kc++) {
            const uint32_t hp = iss * HPI + hh;
            umma_i8(tmem + hp * N, slot + (hp * 4 + kc) * 8, bd0 + (uint64_t)((kc * NV * 256) >> 4),
                    idesc, (i > 0 || kc > 0) ? 1u : 0u);
          }
        umma_commit(bar(T_EMPTY + ts));
        umma_commit(bar(B_EMPTY + bs));
      }
      if (ok) umma_commit(bar(D_FULL));
    }
  } else {
    const uint32_t h = (uint32_t)(warp >> 2), q = (uint32_t)(warp & 3);
    const uint32_t ch = 4 * h + q;  // 16-byte chunk of the stripe
    const uint32_t lane_base = (q * 32u) << 16;
    bool ok = true;
    for (uint32_t i = 0; i < nbx && ok; i++) {
      const uint32_t as = i % NA, ts = i % NT;
      ok = wait_bar(bar(A_FULL + as), (i / NA) & 1, watch, kUErrTimeout | 0x15);
      if (!ok) break;
      const uint32_t tile = a_ring + as * kUBoxBytes;
      uint32_t a[4][4];
#pragma unroll
      for (uint32_t kc = 0; kc < 4; kc++) {
        const uint32_t row = kc * 32 + (uint32_t)lane;
        const uint32_t addr = tile + row * 128u + ((ch ^ (row & 7u)) << 4);
        asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a[kc][0]), "=r"(a[kc][1]), "=r"(a[kc][2]), "=r"(a[kc][3])
                     : "r"(addr));
      }
      if (i >= NT) ok = wait_bar(bar(T_EMPTY + ts), ((i / NT) - 1) & 1, watch, kUErrTimeout | 0x16);
      if (!ok) break;
      tc_fence_after();
      const uint32_t slot = tA + ts * 128 + lane_base;
#pragma unroll
      for (uint32_t p = 0; p < 4; p++) {
        const uint32_t mk = p == 0 ? 0x03030303u : p == 1 ? 0x0F0F0F0Fu : p == 2 ? 0x3F3F3F3Fu : 0xFFFFFFFFu;
        uint32_t d[16];
#pragma unroll
        for (int kc = 0; kc < 4; kc++) {  // (a0, a2 | a1, a3): lanes g | g + 8, columns 2 q', 2 q' + 1
          d[4 * kc + 0] = a[kc][0] & mk;
          d[4 * kc + 1] = a[kc][2] & mk;
          d[4 * kc + 2] = a[kc][1] & mk;
          d[4 * kc + 3] = a[kc][3] & mk;
        }
        const uint32_t hp = 2 * h + (p >> 1);
        tmem_st_16x256b_x4(slot + hp * 32 + (((p & 1u) * 16u) << 16), d);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(A_EMPTY + as));
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(T_FULL + ts));
    }
    // epilogue: read the accumulators back
    if (nbx > 0 && ok) ok = wait_bar(bar(D_FULL), 0, watch, kUErrTimeout | 0x17);
    tc_fence_after();
    {
      const int hi = lane >> 4, bc = lane & 15;
      const uint64_t bytecol = (uint64_t)xbyte0 + ch * 16 + bc;
#pragma unroll 1
      for (int v = 0; v < NV; v++) {
        int c0[8] = {}, c1[8] = {};  // this thread's cumulative plane sums: pp = 0 and pp = 1
        if (nbx > 0 && ok) {
          tmem_ld_32x32b_x8(tmem + (2 * h + 0) * N + v * 8 + lane_base, c0);
          tmem_ld_32x32b_x8(tmem + (2 * h + 1) * N + v * 8 + lane_base, c1);
        }
        // lanes < 16 hold planes 0 and 2, lanes >= 16 planes 1 and 3 of byte column bc
        double fa[8], fb[8];
#pragma unroll
        for (int s = 0; s < 8; s++) {
          const int o0 = __shfl_xor_sync(0xffffffffu, c0[s], 16);
          const int o1 = __shfl_xor_sync(0xffffffffu, c1[s], 16);
          // low lane:  field 0 = P0,       field 2 = P2 - P1
          // high lane: field 1 = P1 - P0,  field 3 = P3 - P2
          fa[s] = hi ? (double)c0[s] - (double)o0 : (double)c0[s];
          fb[s] = hi ? (double)c1[s] - (double)o1 : (double)c1[s] - (double)o0;
        }
        const int f0 = hi ? 1 : 0, f1 = hi ? 3 : 2;
        const double ya = recombine8(fa) * (hi ? 0.25 : 1.0);
        const double yb = recombine8(fb) * (hi ? 0.015625 : 0.0625);
        double* o = out + (uint64_t)v * vstride + (uint64_t)blockIdx.y * sstride;
        const uint64_t ia = bytecol * 4 + f0, ib = bytecol * 4 + f1;
        if (ia < Cn) o[ia] = ya;
        if (ib < Cn) o[ib] = yb;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NW + NISS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace fpb
