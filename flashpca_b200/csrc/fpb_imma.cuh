// fpb_imma.cuh -- the tensor-core perform_op path: exact fixed-point
// 2-bit x int8-slice contraction.
//
// Why not FP64 on the CUDA cores: B200 issues ~62 FP64 adds/clk/SM (measured,
// profiles/r01_microbench_b200.txt); y = X X' x needs >= 2 of them per genotype,
// which caps a CUDA-core kernel near 15 genotypes/clk/SM, while HBM delivers
// ~100 genotypes/clk/SM of packed input.  The int8 tensor path (mma.sync
// m16n8k32, measured 2035 MAC/clk/SM) contracts 512 genotypes x 8 slices per
// instruction straight from registers, so the kernel becomes HBM-bound.
//
// Maths.  With dosage codes e_ij in {0,1,2,3=missing} (fpb_kernels.cuh),
// Mx_j = sum of x_i over the missing genotypes of SNP j, Sx = sum_i x_i:
//   E_j  = sum_i e_ij x_i                                   (tensor cores, exact)
//   t_j  = (X'x)_j = [ (E_j - 3 Mx_j) - mu_j (Sx - Mx_j) ] / sd_j
// and for y = X v with a_j = v_j / sd_j, b_j = mu_j a_j, Sb = sum_j b_j:
//   F_i  = sum_j e_ij a_j                                   (tensor cores, exact)
//   y_i  = F_i - Sb + sum_{j missing for i} (b_j - 3 a_j)
// The missing-genotype terms come from two CSR index lists built at staging
// (missing rate is ~0.15 % in real data; above ~3 % the generic FP64 path is
// used instead).
//
// Exactness.  The input vector is converted to fixed point with a power-of-two
// step delta = 2^(ex-54), ex = exponent of max|x|: Q_i = rint(x_i / (4^f delta))
// where f = (column & 3) is the position of the genotype inside its byte, and
// Q_i is cut into 8 balanced base-128 digits (int8).  The A operand is the packed
// word ANDed with a byte mask -- at most one LOP3 per 4 genotypes, no shifts:
// field f carries e * 4^f, which the 4^-f in Q_i undoes.  The masks are
// cumulative (0x03, 0x0F, 0x3F, none: the fourth operand is the raw word, 3 LOP3
// per 4 operands); the fields are separated again by storing digit DIFFERENCES
// on the B side (first half) or by differencing the accumulators (second half).  Products and int32 accumulation are exact; slices are recombined in
// FP64 (sum_s 128^s D_s, every term exact).  The only rounding is the input
// quantisation, <= 2^-48 max|x| per element.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fpb_kernels.cuh"

namespace fpb {

constexpr int kSliceBits = 54;   // |Q| < 2^54
constexpr int kChunkBytes = 512; // packed bytes of one row per pipeline chunk (2048 columns)
constexpr int kChunkWords = kChunkBytes / 4;     // 128 word-columns
constexpr int kFlushChunks = 24;                 // int32 accumulators -> FP64 every 49152 columns
                                                 // (49152 x 255 x 127 < 2^31)

struct VecScale {   // written by k_slice_vec, read by the finalize kernels
  double sum;       // sum of a vector (deterministic order)
  double delta;     // fixed-point step 2^(ex - 54), 0 for an all-zero vector, NaN if non-finite
  int ex;
  int pad;
};

// ---------------------------------------------------------------------------
// Vector preparation: max|v| and sum(v) with a fixed-order two-stage reduction.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_vec_partial(const double* __restrict__ v, uint64_t n, double* __restrict__ pmax,
              double* __restrict__ psum) {
  __shared__ double smax[8], ssum[8];
  uint64_t per = (n + gridDim.x - 1) / gridDim.x;
  uint64_t b = blockIdx.x * per, e = min(n, b + per);
  double m = 0.0, s = 0.0;
  for (uint64_t i = b + threadIdx.x; i < e; i += 256) {
    double a = v[i];
    s += a;
    a = fabs(a);
    m = (a > m || a != a) ? a : m;  // NaN propagates
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double om = __shfl_xor_sync(0xffffffffu, m, o);
    m = (om > m || om != om) ? om : m;
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = m;
    ssum[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      m = (smax[w] > m || smax[w] != smax[w]) ? smax[w] : m;
      s += ssum[w];
    }
    // thread 0 holds warp 0's values in m, s
    pmax[blockIdx.x] = m;
    psum[blockIdx.x] = s;
  }
}

// Combine per-block (max|v|, sum) partials into a VecScale.  Called by every
// thread of a 128-thread block; every block that calls it on the same partials
// gets the same bits (fixed thread->partial assignment, fixed reduction tree).
__device__ __forceinline__ VecScale scale_from_partials(const double* __restrict__ pmax,
                                                        const double* __restrict__ psum,
                                                        uint32_t nparts) {
  __shared__ double smx[4], ssm[4];
  __shared__ VecScale sres;
  double m = 0.0, s = 0.0;
  for (uint32_t g = threadIdx.x; g < nparts; g += 128) {
    const double pm = pmax[g];
    m = (pm > m || pm != pm) ? pm : m;
    s += psum[g];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double om = __shfl_xor_sync(0xffffffffu, m, o);
    m = (om > m || om != om) ? om : m;
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smx[threadIdx.x >> 5] = m;
    ssm[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; w++) {
      m = (smx[w] > m || smx[w] != smx[w]) ? smx[w] : m;
      s += ssm[w];
    }
    VecScale r;
    r.sum = s;
    r.pad = 0;
    if (m == 0.0) {
      r.ex = 0;
      r.delta = 0.0;
    } else if (!(m < 1.79e308)) {  // NaN or Inf
      r.ex = 0;
      r.delta = nan("");
    } else {
      int ex;
      frexp(m, &ex);  // m = f * 2^ex, 0.5 <= f < 1  ->  |v| < 2^ex
      r.ex = ex;
      r.delta = ldexp(1.0, ex - kSliceBits);
    }
    sres = r;
  }
  __syncthreads();
  return sres;
}

// block-level (max|.|, sum) of one value per thread -> pmax[blockIdx.x], psum[blockIdx.x]
// (256-thread blocks; used by the kernels that produce the next vector to be sliced)
__device__ __forceinline__ void emit_block_partials(double absval, double addend,
                                                    double* __restrict__ pmax,
                                                    double* __restrict__ psum) {
  __shared__ double emx[8], esm[8];
  double m = absval, s = addend;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double om = __shfl_xor_sync(0xffffffffu, m, o);
    m = (om > m || om != om) ? om : m;
    s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  if ((threadIdx.x & 31) == 0) {
    emx[threadIdx.x >> 5] = m;
    esm[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) {
      m = (emx[w] > m || emx[w] != emx[w]) ? emx[w] : m;
      s += esm[w];
    }
    pmax[blockIdx.x] = m;
    psum[blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------
// Slicing: one thread per word-column (16 consecutive vector elements).
// out[wq * 8 + slice_slot(wq, s)] is a uint4 = 16 int8 digits ordered [f][b]:
// digit s of element 16 wq + 4 b + f, pre-divided by 4^f (see file header).  The
// 8 slices of a word-column are rotated by 2*((wq>>2)&3) so that the B-fragment
// LDS.128 of the contraction kernels (lanes g = slice, q = word-column group)
// are bank-conflict free when the block is copied verbatim to shared memory.
// Elements >= n are zero.  nwq = number of word-columns to write.  The scale
// comes from the (max|v|, sum) block partials written by the kernel that
// produced v (k_vec_partial, k_finalize_crossprod or k_prod_inputs).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_slice_vec(const double* __restrict__ v, uint64_t n, uint32_t nwq,
            const double* __restrict__ pmax, const double* __restrict__ psum, uint32_t nparts,
            VecScale* __restrict__ sc_out, uint4* __restrict__ out) {
  const VecScale sc = scale_from_partials(pmax, psum, nparts);
  if (blockIdx.x == 0 && threadIdx.x == 0) *sc_out = sc;  // for the finalize kernel
  uint32_t wq = blockIdx.x * blockDim.x + threadIdx.x;
  if (wq >= nwq) return;
  const int ex = sc.ex;
  const bool live = sc.delta > 0.0;  // false for zero / non-finite vectors
  // The contraction kernels use CUMULATIVE field masks (operand f = word & (4^(f+1) - 1), the
  // last one the raw word): sum_f A_f p_f = sum_g 4^g e_g d_g when p_f = d_f - d_{f+1} (d_4 = 0).
  // The stored digits are those differences, |p| <= 127.
  uint32_t dig[8][4] = {};
#pragma unroll
  for (int b = 0; b < 4; b++) {
    int dg[4][8];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      uint64_t i = (uint64_t)wq * 16 + 4 * b + f;
      long long q = 0;
      if (live && i < n) q = __double2ll_rn(ldexp(v[i], kSliceBits - ex - 2 * f));
#pragma unroll
      for (int s = 0; s < 8; s++) {
        long long d = (s < 7) ? (((q + 64) & 127) - 64) : q;
        q = (q - d) >> 7;
        dg[f][s] = (int)d;
      }
    }
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int s = 0; s < 8; s++) {
        const int pd = dg[f][s] - (f < 3 ? dg[f + 1][s] : 0);
        dig[s][f] |= ((uint32_t)(pd & 0xFF)) << (8 * b);
      }
  }
#pragma unroll
  for (int s = 0; s < 8; s++)
    out[(uint64_t)wq * 8 + ((s + 2 * ((wq >> 2) & 3)) & 7)] =
        make_uint4(dig[s][0], dig[s][1], dig[s][2], dig[s][3]);
}

// ---------------------------------------------------------------------------
// The hot kernel.  G: R rows x pitch bytes of packed dosage codes; S: digit
// slices of the input vector over G's columns.  A CTA of WARPS warps owns
// 16*WARPS consecutive rows and one split of the column chunks; warp w owns
// rows [16w, 16w+16) of the CTA tile.  Per chunk (2048 columns):
//   * the 16 KB slice block of the chunk is brought to shared memory with
//     cp.async (double buffered, XOR-swizzled so the B-fragment LDS.128 are
//     bank-conflict free);
//   * every thread streams its two rows' packed words from HBM with 16-byte
//     non-allocating loads, prefetched one step ahead;
//   * 4 LOP3 + 1 LDS.128 feed two mma.sync.m16n8k32.u8.s8 per packed word.
// out[split * out_stride + row] = sum_s 128^s * D[row][s]  (FP64, exact terms).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// slot (in uint4 units) of slice s of local word-column wl inside a slice block
__device__ __forceinline__ int slice_slot(int wl, int s) {
  return wl * 8 + ((s + 2 * ((wl >> 2) & 3)) & 7);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
k_imma_gemv(const uint8_t* __restrict__ G, uint64_t pitch, uint32_t R,
            const uint4* __restrict__ S, uint32_t nchunks, uint32_t chunks_per_split,
            double* __restrict__ out, uint64_t out_stride) {
  __shared__ uint4 sb[3][kChunkWords * 8];  // 3 x 16 KB ring: one barrier per chunk
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t row_a = min(R - 1, (blockIdx.x * WARPS + warp) * 16 + g);
  const uint32_t row_b = min(R - 1, (blockIdx.x * WARPS + warp) * 16 + g + 8);
  const uint32_t c_begin = blockIdx.y * chunks_per_split;
  const uint32_t c_end = min(nchunks, c_begin + chunks_per_split);
  const uint8_t* pa = G + (uint64_t)row_a * pitch + q * 16;
  const uint8_t* pb = G + (uint64_t)row_b * pitch + q * 16;

  auto issue_slices = [&](uint32_t chunk, int buf) {
    const uint4* src = S + (uint64_t)chunk * (kChunkWords * 8);
    for (int e = tid; e < kChunkWords * 8; e += WARPS * 32)
      cp_async16(&sb[buf][e], src + e);
    cp_async_commit();
  };
  // One batch = 4 steps = 256 contiguous bytes of each of the thread's two rows
  // (the quad's four lanes cover 64 B per step): 8 independent 16-byte loads in
  // flight per thread, and DRAM sees 256-byte bursts per row.
  auto load_batch = [&](uint32_t half, uint4 (&wa)[4], uint4 (&wb)[4]) {
    uint64_t off = (uint64_t)half * 256;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (off + u * 64 + q * 16 < pitch) {
        wa[u] = ld_stream_u128(reinterpret_cast<const uint4*>(pa + off + u * 64));
        wb[u] = ld_stream_u128(reinterpret_cast<const uint4*>(pb + off + u * 64));
      } else {
        wa[u] = make_uint4(0, 0, 0, 0);
        wb[u] = wa[u];
      }
    }
  };

  int acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};  // two independent IMMA chains
  double dacc[4] = {0.0, 0.0, 0.0, 0.0};
  if (c_begin < c_end) {
    issue_slices(c_begin, 0);
    if (c_begin + 1 < c_end) issue_slices(c_begin + 1, 1);
    uint4 na[4], nb[4];
    load_batch(2 * c_begin, na, nb);
    for (uint32_t c = c_begin; c < c_end; c++) {
      const int buf = (c - c_begin) % 3;
      if (c + 1 < c_end) cp_async_wait<1>();
      else cp_async_wait<0>();
      // chunk c's slices have landed for every thread, and every thread has
      // finished chunk c-1, so ring slot (c+2)%3 == (c-1)%3 may be refilled
      __syncthreads();
      if (c + 2 < c_end) issue_slices(c + 2, (buf + 2) % 3);
      const uint4* sbuf = sb[buf];
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        uint4 wa[4], wb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          wa[u] = na[u];
          wb[u] = nb[u];
        }
        if (hb == 0 || c + 1 < c_end) load_batch(2 * c + hb + 1, na, nb);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const uint32_t xa[4] = {wa[u].x, wa[u].y, wa[u].z, wa[u].w};
          const uint32_t xb[4] = {wb[u].x, wb[u].y, wb[u].z, wb[u].w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int wl = (hb * 4 + u) * 16 + q * 4 + j;
            const uint4 bv = sbuf[slice_slot(wl, g)];
            mma_u8s8(acc0, xa[j] & 0x03030303u, xb[j] & 0x03030303u, xa[j] & 0x0F0F0F0Fu,
                     xb[j] & 0x0F0F0F0Fu, bv.x, bv.y);
            mma_u8s8(acc1, xa[j] & 0x3F3F3F3Fu, xb[j] & 0x3F3F3F3Fu, xa[j], xb[j], bv.z, bv.w);
          }
        }
      }
      if (((c - c_begin) % kFlushChunks) == kFlushChunks - 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          dacc[k] += (double)acc0[k] + (double)acc1[k];
          acc0[k] = 0;
          acc1[k] = 0;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) dacc[k] += (double)acc0[k] + (double)acc1[k];
  // c0,c1: row g, slices 2q, 2q+1;  c2,c3: row g+8.  Weight slice s by 128^s.
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  double ra = dacc[0] * w0 + dacc[1] * w1;
  double rb = dacc[2] * w0 + dacc[3] * w1;
  ra += __shfl_xor_sync(0xffffffffu, ra, 1);
  rb += __shfl_xor_sync(0xffffffffu, rb, 1);
  ra += __shfl_xor_sync(0xffffffffu, ra, 2);
  rb += __shfl_xor_sync(0xffffffffu, rb, 2);
  if (q == 0) {
    uint32_t r0 = (blockIdx.x * WARPS + warp) * 16 + g;
    double* o = out + (uint64_t)blockIdx.y * out_stride;
    if (r0 < R) o[r0] = ra;
    if (r0 + 8 < R) o[r0 + 8] = rb;
  }
}

// ---------------------------------------------------------------------------
// The hot kernel, TMA version.  Same maths and fragment mapping as k_imma_gemv;
// the packed genotypes and the digit slices reach shared memory through the
// async proxy instead of registers:
//   * one producer thread issues, per stage, a 2-D cp.async.bulk.tensor of a
//     [256 rows x 128 B] genotype tile (SWIZZLE_128B) and a 4 KB bulk copy of
//     the matching slice block, both completing on the stage's "full" mbarrier;
//   * 8 consumer warps (32 rows each, i.e. two m16 tiles sharing every
//     B-fragment) wait on "full", run 32 IMMAs per stage and release the stage
//     through its "empty" mbarrier -- no CTA-wide barrier in the main loop;
//   * kTmaStages x 36 KB are in flight per SM without holding registers.
// MMA row g of a tile is matrix row rho(g) = (g>>1) | ((g&1)<<2) of the tile so
// that the two rows read by a quarter-warp differ in address bit 9..7 ^ bit 2
// and the swizzled LDS.128 are bank-conflict free.
// ---------------------------------------------------------------------------
constexpr int kTmaStages = 6;
constexpr int kTmaRows = 256;                       // rows per CTA tile
constexpr int kTmaStageCols = 128;                  // packed bytes per row per stage (512 columns)
constexpr int kTmaTileBytes = kTmaRows * kTmaStageCols;          // 32 KB
constexpr int kTmaSliceBytes = (kTmaStageCols / 4) * 8 * 16;     // 32 word-columns x 8 slices = 4 KB
constexpr int kTmaStageBytes = kTmaTileBytes + kTmaSliceBytes;   // 36 KB
// The kernel asks for the whole opt-in shared memory of the SM (232448 B), more
// than the 6 x 36 KB ring + alignment slack + barriers need.  Together with the
// 1 KB the hardware reserves per resident CTA this leaves no room for a CTA of
// any other kernel on the same SM: results were observed to be corrupted
// (per consumer warp, nondeterministically) whenever blocks of another kernel
// were co-resident with this kernel's CTA, so co-residency is ruled out by
// construction (profiles/r01_notes.md).
constexpr int kTmaSmemUsed = kTmaStages * kTmaStageBytes + 1024 + 128;
constexpr int kTmaSmemBytes = 232448;
static_assert(kTmaSmemUsed <= kTmaSmemBytes, "TMA ring does not fit in shared memory");
constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaFlushStages = 96;                 // int32 -> FP64 every 49152 columns / 24576 SNP rows
                                                    // (24576 x 255 x 64 < 2^31 with the raw-word operand)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// The genotype tiles are read exactly once per launch (evict-first in L2); the
// slice blocks are re-read by every row tile of the grid (evict-last).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int x, int y,
                                            uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
      "l"(tmap), "r"(x), "r"(y), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes,
                                          uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

struct alignas(64) TmaDesc {  // CUtensorMap is an opaque 128-byte, 64-byte aligned blob
  unsigned long long opaque[16];
};

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma(const __grid_constant__ TmaDesc tmap, uint32_t R, const uint4* __restrict__ S,
                uint32_t nstages, uint32_t stages_per_split, double* __restrict__ out,
                uint64_t out_stride, uint32_t keep_row = 0xFFFFFFFFu) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B: 1 KB aligned
  const uint32_t bars = base + kTmaStages * kTmaStageBytes;       // full[], then empty[]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s_begin = blockIdx.y * stages_per_split;
  const uint32_t s_end = min(nstages, s_begin + stages_per_split);
  const uint32_t nst = s_end > s_begin ? s_end - s_begin : 0;
  const uint32_t row0 = blockIdx.x * kTmaRows;

  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);                                 // full: producer + tx bytes
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);  // empty: one arrive per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      for (uint32_t it = 0; it < nst; it++) {
        const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
        const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
        if (round > 0) mbar_wait(empty, (round - 1) & 1);
        const uint32_t dst = base + slot * kTmaStageBytes;
        mbar_expect_tx(full, kTmaStageBytes);
        // rows from keep_row on stay in L2 for the second half of the op (evict_last); the rest
        // of the matrix streams through (evict_first)
        tma_load_2d(dst, &tmap, (int)((s_begin + it) * kTmaStageCols), (int)row0, full,
                    row0 >= keep_row ? pol_keep : pol_stream);
        bulk_load(dst + kTmaTileBytes, S + (uint64_t)(s_begin + it) * (kTmaSliceBytes / 16),
                  kTmaSliceBytes, full, pol_keep);
      }
    }
    return;
  }

  // -------------------------------- consumers -------------------------------
  const int g = lane >> 2, q = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);
  // byte offsets inside a stage tile of this thread's four rows (tile t, half h)
  uint32_t roff[2][2];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) roff[t][hf] = (uint32_t)(warp * 32 + t * 16 + hf * 8 + rho) * 128u;
  const uint32_t rx = (uint32_t)(rho & 7);  // (row & 7) of all four rows

  int acc[2][2][4] = {};      // [tile][chain][frag]
  double dacc[2][4] = {};
  for (uint32_t it = 0; it < nst; it++) {
    const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
    mbar_wait(bars + 8 * slot, round & 1);
    const uint32_t tile = base + slot * kTmaStageBytes;
    const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t chunk = ((uint32_t)(u * 4 + q) ^ rx) << 4;
      uint4 w[2][2];
#pragma unroll
      for (int t = 0; t < 2; t++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w[t][hf].x), "=r"(w[t][hf].y), "=r"(w[t][hf].z), "=r"(w[t][hf].w)
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
          const uint32_t xa = j == 0 ? w[t][0].x : j == 1 ? w[t][0].y : j == 2 ? w[t][0].z : w[t][0].w;
          const uint32_t xb = j == 0 ? w[t][1].x : j == 1 ? w[t][1].y : j == 2 ? w[t][1].z : w[t][1].w;
          mma_u8s8(acc[t][0], xa & 0x03030303u, xb & 0x03030303u, xa & 0x0F0F0F0Fu,
                   xb & 0x0F0F0F0Fu, bv.x, bv.y);
          mma_u8s8(acc[t][1], xa & 0x3F3F3F3Fu, xb & 0x3F3F3F3Fu, xa, xb, bv.z, bv.w);
        }
      }
    }
    // Release the stage.  The fence makes every lane's shared-memory reads of
    // this stage complete before the arrive can be observed: without it ptxas
    // schedules the SYNCS.ARRIVE right behind the *issue* of the stage's last
    // LDS.128 (ahead of the IMMAs that consume it), and the producer's next TMA
    // write into the slot can overtake that load when the LSU is congested
    // (observed as per-warp corruption, profiles/r01_notes.md).
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
    if ((it % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
      for (int t = 0; t < 2; t++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          dacc[t][k] += (double)acc[t][0][k] + (double)acc[t][1][k];
          acc[t][0][k] = 0;
          acc[t][1][k] = 0;
        }
    }
  }
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  double* o = out + (uint64_t)blockIdx.y * out_stride;
#pragma unroll
  for (int t = 0; t < 2; t++) {
#pragma unroll
    for (int k = 0; k < 4; k++) dacc[t][k] += (double)acc[t][0][k] + (double)acc[t][1][k];
    double ra = dacc[t][0] * w0 + dacc[t][1] * w1;
    double rb = dacc[t][2] * w0 + dacc[t][3] * w1;
    ra += __shfl_xor_sync(0xffffffffu, ra, 1);
    rb += __shfl_xor_sync(0xffffffffu, rb, 1);
    ra += __shfl_xor_sync(0xffffffffu, ra, 2);
    rb += __shfl_xor_sync(0xffffffffu, rb, 2);
    if (q == 0) {
      uint32_t r = row0 + warp * 32 + t * 16 + rho;
      if (r < R) o[r] = ra;
      if (r + 8 < R) o[r + 8] = rb;
    }
  }
}

// ---------------------------------------------------------------------------
// Second half from the SAME SNP-major copy:  F_i = sum_j e_ij a_j.
//
// The reduction now runs over SNPs (tile rows), while a packed byte still holds
// four individuals.  ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 (SASS
// LDSM.8.MT1616, sm_100a) transposes BYTES on the way to registers and returns
// exactly the m16n8k32 A fragment: reg0..3 = a0..a3 with M = 16 byte-columns
// (= 64 individuals) and K = 32 consecutive SNPs (tools/ldsm_probe.cu).  The
// field masks 0x03/0x0C/0x30/0xC0 then select individual 4*bytecol + f for a
// whole MMA, so the 4^f factor is uniform per accumulator and is removed in
// the epilogue; the B operand is the plain digit slices of a_j (K-major).
//
// A CTA owns one 128-byte column stripe (512 individuals) and walks a split of
// the 256-row SNP tiles with the same TMA box / mbarrier ring as above; consumer
// warp w owns the 16-byte chunk w of the stripe (64 individuals), 4 accumulator
// sets (one per field).  Per stage and warp: 8 LDSM.x2, 8 LDS.64, 128 LOP3,
// 32 IMMA.  out[split * out_stride + individual].
//
// Slices for this kernel (k_slice_vec_k): per 32 SNPs 256 bytes laid out
// [slice g][q][b0 (4 digits of SNPs 4q..4q+3) | b1 (SNPs 16+4q..)], i.e. the B
// fragment of lane (g, q) is one 8-byte load.
// ---------------------------------------------------------------------------
constexpr int kTmaTSliceBytes = (kTmaRows / 32) * 256;          // 2 KB per stage
constexpr int kTmaTStageBytes = kTmaTileBytes + kTmaTSliceBytes;  // 34 KB
static_assert(kTmaStages * kTmaTStageBytes + 1024 + 128 <= kTmaSmemBytes, "ring too large");

__global__ void __launch_bounds__(128)
k_slice_vec_k(const double* __restrict__ v, uint64_t n, uint32_t ngroups4,
              const double* __restrict__ pmax, const double* __restrict__ psum, uint32_t nparts,
              VecScale* __restrict__ sc_out, uint32_t* __restrict__ out) {
  const VecScale sc = scale_from_partials(pmax, psum, nparts);
  if (blockIdx.x == 0 && threadIdx.x == 0) *sc_out = sc;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // 4 consecutive elements
  if (t >= ngroups4) return;
  const int ex = sc.ex;
  const bool live = sc.delta > 0.0;
  uint32_t dig[8] = {};
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const uint64_t i = (uint64_t)t * 4 + b;
    long long q = 0;
    if (live && i < n) q = __double2ll_rn(ldexp(v[i], kSliceBits - ex));
#pragma unroll
    for (int s = 0; s < 8; s++) {
      long long d = (s < 7) ? (((q + 64) & 127) - 64) : q;
      q = (q - d) >> 7;
      dig[s] |= ((uint32_t)(d & 0xFF)) << (8 * b);
    }
  }
  const uint32_t grp = t >> 3, kk4 = t & 7;       // 8 groups of 4 per 32 elements
  const uint32_t h = kk4 >> 2, q4 = kk4 & 3;      // b0 (k < 16) or b1, and lane q
#pragma unroll
  for (int s = 0; s < 8; s++) out[(((uint64_t)grp * 8 + s) * 4 + q4) * 2 + h] = dig[s];
}

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_t(const __grid_constant__ TmaDesc tmap, uint32_t C /* output length */,
                  const uint32_t* __restrict__ S, uint32_t ntiles, uint32_t tiles_per_split,
                  double* __restrict__ out, uint64_t out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTmaStages * kTmaTStageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t_begin = blockIdx.y * tiles_per_split;
  const uint32_t t_end = min(ntiles, t_begin + tiles_per_split);
  const uint32_t nst = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t xbyte0 = blockIdx.x * kTmaStageCols;

  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      for (uint32_t it = 0; it < nst; it++) {
        const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
        const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
        if (round > 0) mbar_wait(empty, (round - 1) & 1);
        const uint32_t dst = base + slot * kTmaTStageBytes;
        mbar_expect_tx(full, kTmaTStageBytes);
        tma_load_2d(dst, &tmap, (int)xbyte0, (int)((t_begin + it) * kTmaRows), full, pol_stream);
        bulk_load(dst + kTmaTileBytes, S + (uint64_t)(t_begin + it) * (kTmaTSliceBytes / 4),
                  kTmaTSliceBytes, full, pol_keep);
      }
    }
    return;
  }

  const int g = lane >> 2, q = lane & 3;
  int acc[4][4] = {};       // [field][frag]
  double dacc[4][4] = {};
  for (uint32_t it = 0; it < nst; it++) {
    const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
    mbar_wait(bars + 8 * slot, round & 1);
    const uint32_t tile = base + slot * kTmaTStageBytes;
    const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
    for (int ks = 0; ks < kTmaRows / 32; ks++) {
      const uint32_t row = (uint32_t)(ks * 32 + lane);  // lane l supplies the address of row l
      const uint32_t addr = tile + row * 128u + ((((uint32_t)warp) ^ (row & 7u)) << 4);
      uint32_t a0, a1, a2, a3, b0, b1;
      asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                   : "r"(addr));
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                   : "=r"(b0), "=r"(b1)
                   : "r"(sl + (uint32_t)(((ks * 8 + g) * 4 + q) * 8)));
      // cumulative masks: accumulator f holds sum_{g <= f} 4^g e_g a (12 LOP3 per 4 IMMAs instead
      // of 16, the last operand is the raw packed word); the epilogue takes differences
      mma_u8s8(acc[0], a0 & 0x03030303u, a1 & 0x03030303u, a2 & 0x03030303u, a3 & 0x03030303u, b0, b1);
      mma_u8s8(acc[1], a0 & 0x0F0F0F0Fu, a1 & 0x0F0F0F0Fu, a2 & 0x0F0F0F0Fu, a3 & 0x0F0F0F0Fu, b0, b1);
      mma_u8s8(acc[2], a0 & 0x3F3F3F3Fu, a1 & 0x3F3F3F3Fu, a2 & 0x3F3F3F3Fu, a3 & 0x3F3F3F3Fu, b0, b1);
      mma_u8s8(acc[3], a0, a1, a2, a3, b0, b1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
    if ((it % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          dacc[f][k] += (double)acc[f][k];
          acc[f][k] = 0;
        }
    }
  }
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  double* o = out + (uint64_t)blockIdx.y * out_stride;
  const uint64_t byte_a = (uint64_t)xbyte0 + warp * 16 + g;  // packed byte of MMA row g; +8 for g+8
#pragma unroll
  for (int f = 0; f < 4; f++)
#pragma unroll
    for (int k = 0; k < 4; k++) dacc[f][k] += (double)acc[f][k];
#pragma unroll
  for (int f = 3; f > 0; f--)  // cumulative -> per field (integers below 2^53: exact)
#pragma unroll
    for (int k = 0; k < 4; k++) dacc[f][k] -= dacc[f - 1][k];
#pragma unroll
  for (int f = 0; f < 4; f++) {
    const double sf = ldexp(1.0, -2 * f);  // the field carried e * 4^f
    double ra = (dacc[f][0] * w0 + dacc[f][1] * w1) * sf;
    double rb = (dacc[f][2] * w0 + dacc[f][3] * w1) * sf;
    ra += __shfl_xor_sync(0xffffffffu, ra, 1);
    rb += __shfl_xor_sync(0xffffffffu, rb, 1);
    ra += __shfl_xor_sync(0xffffffffu, ra, 2);
    rb += __shfl_xor_sync(0xffffffffu, rb, 2);
    if (q == 0) {
      const uint64_t ia = byte_a * 4 + f, ib = (byte_a + 8) * 4 + f;
      if (ia < C) o[ia] = ra;
      if (ib < C) o[ib] = rb;
    }
  }
}

// ---------------------------------------------------------------------------
// Two-vector forms (the block variants perform_op_mat / crossprod2 / prod3,
// svdwide.cpp:71-118, 157-188, 312-343): one pass over the packed matrix serves
// two input vectors.  The masked A fragments are built once and fed to two sets
// of IMMAs (the digit slices of both vectors travel with every stage), so the
// LOP3 decode and the HBM traffic are shared; the tensor pipe, at ~40 % for one
// vector, has room for the second.
// ---------------------------------------------------------------------------
constexpr int kTma2Stages = 5;
constexpr int kTma2StageBytes = kTmaTileBytes + 2 * kTmaSliceBytes;          // 40 KB
static_assert(kTma2Stages * kTma2StageBytes + 1024 + 128 <= kTmaSmemBytes, "ring too large");

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_2v(const __grid_constant__ TmaDesc tmap, uint32_t R, const uint4* __restrict__ S0,
                   const uint4* __restrict__ S1, uint32_t nstages, uint32_t stages_per_split,
                   double* __restrict__ out0, double* __restrict__ out1, uint64_t out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTma2Stages * kTma2StageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t s_begin = blockIdx.y * stages_per_split;
  const uint32_t s_end = min(nstages, s_begin + stages_per_split);
  const uint32_t nst = s_end > s_begin ? s_end - s_begin : 0;
  const uint32_t row0 = blockIdx.x * kTmaRows;
  if (tid == 0) {
    for (int i = 0; i < kTma2Stages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTma2Stages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      for (uint32_t it = 0; it < nst; it++) {
        const uint32_t slot = it % kTma2Stages, round = it / kTma2Stages;
        const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTma2Stages + slot);
        if (round > 0) mbar_wait(empty, (round - 1) & 1);
        const uint32_t dst = base + slot * kTma2StageBytes;
        mbar_expect_tx(full, kTma2StageBytes);
        tma_load_2d(dst, &tmap, (int)((s_begin + it) * kTmaStageCols), (int)row0, full, pol_stream);
        const uint64_t so = (uint64_t)(s_begin + it) * (kTmaSliceBytes / 16);
        bulk_load(dst + kTmaTileBytes, S0 + so, kTmaSliceBytes, full, pol_keep);
        bulk_load(dst + kTmaTileBytes + kTmaSliceBytes, S1 + so, kTmaSliceBytes, full, pol_keep);
      }
    }
    return;
  }
  const int g = lane >> 2, q = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);
  uint32_t roff[2][2];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) roff[t][hf] = (uint32_t)(warp * 32 + t * 16 + hf * 8 + rho) * 128u;
  const uint32_t rx = (uint32_t)(rho & 7);
  int acc[2][2][2][4] = {};   // [vector][tile][chain][frag]
  double dacc[2][2][4] = {};
  for (uint32_t it = 0; it < nst; it++) {
    const uint32_t slot = it % kTma2Stages, round = it / kTma2Stages;
    mbar_wait(bars + 8 * slot, round & 1);
    const uint32_t tile = base + slot * kTma2StageBytes;
    const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const uint32_t chunk = ((uint32_t)(u * 4 + q) ^ rx) << 4;
      uint4 w[2][2];
#pragma unroll
      for (int t = 0; t < 2; t++)
#pragma unroll
        for (int hf = 0; hf < 2; hf++)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(w[t][hf].x), "=r"(w[t][hf].y), "=r"(w[t][hf].z), "=r"(w[t][hf].w)
                       : "r"(tile + roff[t][hf] + chunk));
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int wl = u * 16 + q * 4 + j;
        uint4 bv[2];
#pragma unroll
        for (int v = 0; v < 2; v++)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(bv[v].x), "=r"(bv[v].y), "=r"(bv[v].z), "=r"(bv[v].w)
                       : "r"(sl + (uint32_t)v * kTmaSliceBytes + (uint32_t)slice_slot(wl, g) * 16u));
#pragma unroll
        for (int t = 0; t < 2; t++) {
          const uint32_t xa = j == 0 ? w[t][0].x : j == 1 ? w[t][0].y : j == 2 ? w[t][0].z : w[t][0].w;
          const uint32_t xb = j == 0 ? w[t][1].x : j == 1 ? w[t][1].y : j == 2 ? w[t][1].z : w[t][1].w;
          const uint32_t a0 = xa & 0x03030303u, a1 = xb & 0x03030303u, a2 = xa & 0x0F0F0F0Fu,
                         a3 = xb & 0x0F0F0F0Fu, c0 = xa & 0x3F3F3F3Fu, c1 = xb & 0x3F3F3F3Fu;
#pragma unroll
          for (int v = 0; v < 2; v++) {
            mma_u8s8(acc[v][t][0], a0, a1, a2, a3, bv[v].x, bv[v].y);
            mma_u8s8(acc[v][t][1], c0, c1, xa, xb, bv[v].z, bv[v].w);
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (kTma2Stages + slot));
    if ((it % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
      for (int v = 0; v < 2; v++)
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dacc[v][t][k] += (double)acc[v][t][0][k] + (double)acc[v][t][1][k];
            acc[v][t][0][k] = 0;
            acc[v][t][1][k] = 0;
          }
    }
  }
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
#pragma unroll
  for (int v = 0; v < 2; v++) {
    double* o = (v == 0 ? out0 : out1) + (uint64_t)blockIdx.y * out_stride;
#pragma unroll
    for (int t = 0; t < 2; t++) {
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[v][t][k] += (double)acc[v][t][0][k] + (double)acc[v][t][1][k];
      double ra = dacc[v][t][0] * w0 + dacc[v][t][1] * w1;
      double rb = dacc[v][t][2] * w0 + dacc[v][t][3] * w1;
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      if (q == 0) {
        const uint32_t r = row0 + warp * 32 + t * 16 + rho;
        if (r < R) o[r] = ra;
        if (r + 8 < R) o[r + 8] = rb;
      }
    }
  }
}

constexpr int kTmaT2StageBytes = kTmaTileBytes + 2 * kTmaTSliceBytes;          // 36 KB
static_assert(kTmaStages * kTmaT2StageBytes + 1024 + 128 <= kTmaSmemBytes, "ring too large");

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_t_2v(const __grid_constant__ TmaDesc tmap, uint32_t C, const uint32_t* __restrict__ S0,
                     const uint32_t* __restrict__ S1, uint32_t ntiles, uint32_t tiles_per_split,
                     double* __restrict__ out0, double* __restrict__ out1, uint64_t out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTmaStages * kTmaT2StageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t_begin = blockIdx.y * tiles_per_split;
  const uint32_t t_end = min(ntiles, t_begin + tiles_per_split);
  const uint32_t nst = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t xbyte0 = blockIdx.x * kTmaStageCols;
  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      for (uint32_t it = 0; it < nst; it++) {
        const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
        const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
        if (round > 0) mbar_wait(empty, (round - 1) & 1);
        const uint32_t dst = base + slot * kTmaT2StageBytes;
        mbar_expect_tx(full, kTmaT2StageBytes);
        tma_load_2d(dst, &tmap, (int)xbyte0, (int)((t_begin + it) * kTmaRows), full, pol_stream);
        const uint64_t so = (uint64_t)(t_begin + it) * (kTmaTSliceBytes / 4);
        bulk_load(dst + kTmaTileBytes, S0 + so, kTmaTSliceBytes, full, pol_keep);
        bulk_load(dst + kTmaTileBytes + kTmaTSliceBytes, S1 + so, kTmaTSliceBytes, full, pol_keep);
      }
    }
    return;
  }
  const int g = lane >> 2, q = lane & 3;
  int acc[2][4][4] = {};       // [vector][field (cumulative)][frag]
  double dacc[2][4][4] = {};
  for (uint32_t it = 0; it < nst; it++) {
    const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
    mbar_wait(bars + 8 * slot, round & 1);
    const uint32_t tile = base + slot * kTmaT2StageBytes;
    const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
    for (int ks = 0; ks < kTmaRows / 32; ks++) {
      const uint32_t row = (uint32_t)(ks * 32 + lane);
      const uint32_t addr = tile + row * 128u + ((((uint32_t)warp) ^ (row & 7u)) << 4);
      uint32_t a0, a1, a2, a3;
      asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                   : "r"(addr));
      const uint32_t m00 = a0 & 0x03030303u, m01 = a1 & 0x03030303u, m02 = a2 & 0x03030303u,
                     m03 = a3 & 0x03030303u, m10 = a0 & 0x0F0F0F0Fu, m11 = a1 & 0x0F0F0F0Fu,
                     m12 = a2 & 0x0F0F0F0Fu, m13 = a3 & 0x0F0F0F0Fu, m20 = a0 & 0x3F3F3F3Fu,
                     m21 = a1 & 0x3F3F3F3Fu, m22 = a2 & 0x3F3F3F3Fu, m23 = a3 & 0x3F3F3F3Fu;
#pragma unroll
      for (int v = 0; v < 2; v++) {
        uint32_t b0, b1;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                     : "=r"(b0), "=r"(b1)
                     : "r"(sl + (uint32_t)v * kTmaTSliceBytes + (uint32_t)(((ks * 8 + g) * 4 + q) * 8)));
        mma_u8s8(acc[v][0], m00, m01, m02, m03, b0, b1);
        mma_u8s8(acc[v][1], m10, m11, m12, m13, b0, b1);
        mma_u8s8(acc[v][2], m20, m21, m22, m23, b0, b1);
        mma_u8s8(acc[v][3], a0, a1, a2, a3, b0, b1);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
    if ((it % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
      for (int v = 0; v < 2; v++)
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dacc[v][f][k] += (double)acc[v][f][k];
            acc[v][f][k] = 0;
          }
    }
  }
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  const uint64_t byte_a = (uint64_t)xbyte0 + warp * 16 + g;
#pragma unroll
  for (int v = 0; v < 2; v++) {
    double* o = (v == 0 ? out0 : out1) + (uint64_t)blockIdx.y * out_stride;
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[v][f][k] += (double)acc[v][f][k];
#pragma unroll
    for (int f = 3; f > 0; f--)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[v][f][k] -= dacc[v][f - 1][k];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const double sf = ldexp(1.0, -2 * f);
      double ra = (dacc[v][f][0] * w0 + dacc[v][f][1] * w1) * sf;
      double rb = (dacc[v][f][2] * w0 + dacc[v][f][3] * w1) * sf;
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      if (q == 0) {
        const uint64_t ia = byte_a * 4 + f, ib = (byte_a + 8) * 4 + f;
        if (ia < C) o[ia] = ra;
        if (ib < C) o[ib] = rb;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Second half, wide-stripe form.  A CTA owns a 256-byte column stripe (1024
// individuals): each stage holds two [128 rows x 128 B] boxes that are adjacent
// in memory, so every SNP row is read in 256-byte runs (k_imma_gemv_tma_t reads
// 128-byte runs and relies on the neighbouring CTA for the other half of the
// DRAM burst), the B fragments are shared by the two boxes, and a CTA walks
// twice as many stages (half the pipeline-fill overhead per byte).
// ---------------------------------------------------------------------------
constexpr int kTmaWRows = 128;
constexpr int kTmaWBoxBytes = kTmaWRows * 128;                           // 16 KB
constexpr int kTmaWSliceBytes = (kTmaWRows / 32) * 256;                  // 1 KB
constexpr int kTmaWStageBytes = 2 * kTmaWBoxBytes + kTmaWSliceBytes;     // 33 KB
static_assert(kTmaStages * kTmaWStageBytes + 1024 + 128 <= kTmaSmemBytes, "ring too large");
constexpr int kTmaWFlushStages = 192;                                    // 24576 SNP rows

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_tw(const __grid_constant__ TmaDesc tmap /* box 128 B x 128 rows */, uint32_t C,
                   const uint32_t* __restrict__ S, uint32_t ntiles, uint32_t tiles_per_split,
                   double* __restrict__ out, uint64_t out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTmaStages * kTmaWStageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t_begin = blockIdx.y * tiles_per_split;
  const uint32_t t_end = min(ntiles, t_begin + tiles_per_split);
  const uint32_t nst = t_end > t_begin ? t_end - t_begin : 0;
  const uint32_t xbyte0 = blockIdx.x * 256u;

  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      for (uint32_t it = 0; it < nst; it++) {
        const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
        const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
        if (round > 0) mbar_wait(empty, (round - 1) & 1);
        const uint32_t dst = base + slot * kTmaWStageBytes;
        mbar_expect_tx(full, kTmaWStageBytes);
        const int y = (int)((t_begin + it) * kTmaWRows);
        tma_load_2d(dst, &tmap, (int)xbyte0, y, full, pol_stream);
        tma_load_2d(dst + kTmaWBoxBytes, &tmap, (int)xbyte0 + 128, y, full, pol_stream);
        bulk_load(dst + 2 * kTmaWBoxBytes, S + (uint64_t)(t_begin + it) * (kTmaWSliceBytes / 4),
                  kTmaWSliceBytes, full, pol_keep);
      }
    }
    return;
  }

  const int g = lane >> 2, q = lane & 3;
  int acc[2][4][4] = {};       // [box][field (cumulative)][frag]
  double dacc[2][4][4] = {};
  for (uint32_t it = 0; it < nst; it++) {
    const uint32_t slot = it % kTmaStages, round = it / kTmaStages;
    mbar_wait(bars + 8 * slot, round & 1);
    const uint32_t tile = base + slot * kTmaWStageBytes;
    const uint32_t sl = tile + 2 * kTmaWBoxBytes;
#pragma unroll
    for (int ks = 0; ks < kTmaWRows / 32; ks++) {
      const uint32_t row = (uint32_t)(ks * 32 + lane);
      const uint32_t addr = tile + row * 128u + ((((uint32_t)warp) ^ (row & 7u)) << 4);
      uint32_t b0, b1;
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                   : "=r"(b0), "=r"(b1)
                   : "r"(sl + (uint32_t)(((ks * 8 + g) * 4 + q) * 8)));
#pragma unroll
      for (int bx = 0; bx < 2; bx++) {
        uint32_t a0, a1, a2, a3;
        asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(addr + (uint32_t)bx * kTmaWBoxBytes));
        mma_u8s8(acc[bx][0], a0 & 0x03030303u, a1 & 0x03030303u, a2 & 0x03030303u, a3 & 0x03030303u, b0, b1);
        mma_u8s8(acc[bx][1], a0 & 0x0F0F0F0Fu, a1 & 0x0F0F0F0Fu, a2 & 0x0F0F0F0Fu, a3 & 0x0F0F0F0Fu, b0, b1);
        mma_u8s8(acc[bx][2], a0 & 0x3F3F3F3Fu, a1 & 0x3F3F3F3Fu, a2 & 0x3F3F3F3Fu, a3 & 0x3F3F3F3Fu, b0, b1);
        mma_u8s8(acc[bx][3], a0, a1, a2, a3, b0, b1);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
    __threadfence_block();
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
    if ((it % kTmaWFlushStages) == kTmaWFlushStages - 1) {
#pragma unroll
      for (int bx = 0; bx < 2; bx++)
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dacc[bx][f][k] += (double)acc[bx][f][k];
            acc[bx][f][k] = 0;
          }
    }
  }
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  double* o = out + (uint64_t)blockIdx.y * out_stride;
#pragma unroll
  for (int bx = 0; bx < 2; bx++) {
    const uint64_t byte_a = (uint64_t)xbyte0 + bx * 128 + warp * 16 + g;
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[bx][f][k] += (double)acc[bx][f][k];
#pragma unroll
    for (int f = 3; f > 0; f--)  // cumulative -> per field (integers below 2^53: exact)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[bx][f][k] -= dacc[bx][f - 1][k];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const double sf = ldexp(1.0, -2 * f);
      double ra = (dacc[bx][f][0] * w0 + dacc[bx][f][1] * w1) * sf;
      double rb = (dacc[bx][f][2] * w0 + dacc[bx][f][3] * w1) * sf;
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      if (q == 0) {
        const uint64_t ia = byte_a * 4 + f, ib = (byte_a + 8) * 4 + f;
        if (ia < C) o[ia] = ra;
        if (ib < C) o[ib] = rb;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Persistent forms of the two TMA contraction kernels.  One CTA per SM walks a
// static round-robin list of work items (item = one tile of the output x one
// split of the reduction axis); the TMA ring keeps running across items, so the
// pipeline is filled once per launch instead of once per CTA and there is no
// partial last wave (the one-item-per-CTA grids above lose ~4 % to each).
// Same fragment mapping, fences and output layout as the kernels above.
// Items are numbered tile-fastest (item = split x ntiles + tile), like blockIdx.x of the grids
// above: CTAs that run side by side then read neighbouring 128-byte columns of the same rows
// (second half) -- whole DRAM pages between them; split-fastest numbering cost 45 % there.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_p(const __grid_constant__ TmaDesc tmap, uint32_t R, const uint4* __restrict__ S,
                  uint32_t nstages, uint32_t stages_per_split, uint32_t nsplits, uint32_t nitems,
                  double* __restrict__ out, uint64_t out_stride, uint32_t keep_row = 0xFFFFFFFFu) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTmaStages * kTmaStageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      uint32_t slot = 0, round = 0;
      for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint32_t rt = item % (nitems / nsplits), sp = item / (nitems / nsplits);
        const uint32_t s_begin = sp * stages_per_split;
        const uint32_t s_end = min(nstages, s_begin + stages_per_split);
        for (uint32_t st = s_begin; st < s_end; st++) {
          const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
          if (round > 0) mbar_wait(empty, (round - 1) & 1);
          const uint32_t dst = base + slot * kTmaStageBytes;
          mbar_expect_tx(full, kTmaStageBytes);
          tma_load_2d(dst, &tmap, (int)(st * kTmaStageCols), (int)(rt * kTmaRows), full,
                      rt * kTmaRows >= keep_row ? pol_keep : pol_stream);
          bulk_load(dst + kTmaTileBytes, S + (uint64_t)st * (kTmaSliceBytes / 16), kTmaSliceBytes,
                    full, pol_keep);
          if (++slot == kTmaStages) {
            slot = 0;
            round++;
          }
        }
      }
    }
    return;
  }

  const int g = lane >> 2, q = lane & 3;
  const int rho = (g >> 1) | ((g & 1) << 2);
  uint32_t roff[2][2];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int hf = 0; hf < 2; hf++) roff[t][hf] = (uint32_t)(warp * 32 + t * 16 + hf * 8 + rho) * 128u;
  const uint32_t rx = (uint32_t)(rho & 7);
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  uint32_t slot = 0, round = 0;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t rt = item % (nitems / nsplits), sp = item / (nitems / nsplits);
    const uint32_t s_begin = sp * stages_per_split;
    const uint32_t s_end = min(nstages, s_begin + stages_per_split);
    int acc[2][2][4] = {};
    double dacc[2][4] = {};
    for (uint32_t st = s_begin; st < s_end; st++) {
      mbar_wait(bars + 8 * slot, round & 1);
      const uint32_t tile = base + slot * kTmaStageBytes;
      const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t chunk = ((uint32_t)(u * 4 + q) ^ rx) << 4;
        uint4 w[2][2];
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
          for (int hf = 0; hf < 2; hf++)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(w[t][hf].x), "=r"(w[t][hf].y), "=r"(w[t][hf].z), "=r"(w[t][hf].w)
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
            const uint32_t xa = j == 0 ? w[t][0].x : j == 1 ? w[t][0].y : j == 2 ? w[t][0].z : w[t][0].w;
            const uint32_t xb = j == 0 ? w[t][1].x : j == 1 ? w[t][1].y : j == 2 ? w[t][1].z : w[t][1].w;
            mma_u8s8(acc[t][0], xa & 0x03030303u, xb & 0x03030303u, xa & 0x0F0F0F0Fu,
                     xb & 0x0F0F0F0Fu, bv.x, bv.y);
            mma_u8s8(acc[t][1], xa & 0x3F3F3F3Fu, xb & 0x3F3F3F3Fu, xa, xb, bv.z, bv.w);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
      if (++slot == kTmaStages) {
        slot = 0;
        round++;
      }
      if (((st - s_begin) % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dacc[t][k] += (double)acc[t][0][k] + (double)acc[t][1][k];
            acc[t][0][k] = 0;
            acc[t][1][k] = 0;
          }
      }
    }
    double* o = out + (uint64_t)sp * out_stride;
#pragma unroll
    for (int t = 0; t < 2; t++) {
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[t][k] += (double)acc[t][0][k] + (double)acc[t][1][k];
      double ra = dacc[t][0] * w0 + dacc[t][1] * w1;
      double rb = dacc[t][2] * w0 + dacc[t][3] * w1;
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      if (q == 0) {
        const uint32_t r = rt * kTmaRows + warp * 32 + t * 16 + rho;
        if (r < R) o[r] = ra;
        if (r + 8 < R) o[r + 8] = rb;
      }
    }
  }
}

__global__ void __launch_bounds__((kTmaConsumerWarps + 1) * 32, 1)
k_imma_gemv_tma_t_p(const __grid_constant__ TmaDesc tmap, uint32_t C /* output length */,
                    const uint32_t* __restrict__ S, uint32_t ntiles, uint32_t tiles_per_split,
                    uint32_t nsplits, uint32_t nitems, double* __restrict__ out,
                    uint64_t out_stride) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTmaStages * kTmaTStageBytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < kTmaStages; i++) {
      mbar_init(bars + 8 * i, 1);
      mbar_init(bars + 8 * (kTmaStages + i), kTmaConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumerWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      uint32_t slot = 0, round = 0;
      for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const uint32_t cs = item % (nitems / nsplits), sp = item / (nitems / nsplits);
        const uint32_t t_begin = sp * tiles_per_split;
        const uint32_t t_end = min(ntiles, t_begin + tiles_per_split);
        for (uint32_t tt = t_begin; tt < t_end; tt++) {
          const uint32_t full = bars + 8 * slot, empty = bars + 8 * (kTmaStages + slot);
          if (round > 0) mbar_wait(empty, (round - 1) & 1);
          const uint32_t dst = base + slot * kTmaTStageBytes;
          mbar_expect_tx(full, kTmaTStageBytes);
          tma_load_2d(dst, &tmap, (int)(cs * kTmaStageCols), (int)(tt * kTmaRows), full, pol_stream);
          bulk_load(dst + kTmaTileBytes, S + (uint64_t)tt * (kTmaTSliceBytes / 4), kTmaTSliceBytes,
                    full, pol_keep);
          if (++slot == kTmaStages) {
            slot = 0;
            round++;
          }
        }
      }
    }
    return;
  }

  const int g = lane >> 2, q = lane & 3;
  const double w0 = ldexp(1.0, 14 * q), w1 = ldexp(1.0, 14 * q + 7);
  uint32_t slot = 0, round = 0;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t cs = item % (nitems / nsplits), sp = item / (nitems / nsplits);
    const uint32_t t_begin = sp * tiles_per_split;
    const uint32_t t_end = min(ntiles, t_begin + tiles_per_split);
    int acc[4][4] = {};
    double dacc[4][4] = {};
    for (uint32_t tt = t_begin; tt < t_end; tt++) {
      mbar_wait(bars + 8 * slot, round & 1);
      const uint32_t tile = base + slot * kTmaTStageBytes;
      const uint32_t sl = tile + kTmaTileBytes;
#pragma unroll
      for (int ks = 0; ks < kTmaRows / 32; ks++) {
        const uint32_t row = (uint32_t)(ks * 32 + lane);
        const uint32_t addr = tile + row * 128u + ((((uint32_t)warp) ^ (row & 7u)) << 4);
        uint32_t a0, a1, a2, a3, b0, b1;
        asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                     : "r"(addr));
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                     : "=r"(b0), "=r"(b1)
                     : "r"(sl + (uint32_t)(((ks * 8 + g) * 4 + q) * 8)));
        mma_u8s8(acc[0], a0 & 0x03030303u, a1 & 0x03030303u, a2 & 0x03030303u, a3 & 0x03030303u, b0, b1);
        mma_u8s8(acc[1], a0 & 0x0F0F0F0Fu, a1 & 0x0F0F0F0Fu, a2 & 0x0F0F0F0Fu, a3 & 0x0F0F0F0Fu, b0, b1);
        mma_u8s8(acc[2], a0 & 0x3F3F3F3Fu, a1 & 0x3F3F3F3Fu, a2 & 0x3F3F3F3Fu, a3 & 0x3F3F3F3Fu, b0, b1);
        mma_u8s8(acc[3], a0, a1, a2, a3, b0, b1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // see k_imma_gemv_tma
      __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (kTmaStages + slot));
      if (++slot == kTmaStages) {
        slot = 0;
        round++;
      }
      if (((tt - t_begin) % kTmaFlushStages) == kTmaFlushStages - 1) {
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            dacc[f][k] += (double)acc[f][k];
            acc[f][k] = 0;
          }
      }
    }
    double* o = out + (uint64_t)sp * out_stride;
    const uint64_t byte_a = (uint64_t)cs * kTmaStageCols + warp * 16 + g;
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[f][k] += (double)acc[f][k];
#pragma unroll
    for (int f = 3; f > 0; f--)  // cumulative -> per field (integers below 2^53: exact)
#pragma unroll
      for (int k = 0; k < 4; k++) dacc[f][k] -= dacc[f - 1][k];
#pragma unroll
    for (int f = 0; f < 4; f++) {
      const double sf = ldexp(1.0, -2 * f);
      double ra = (dacc[f][0] * w0 + dacc[f][1] * w1) * sf;
      double rb = (dacc[f][2] * w0 + dacc[f][3] * w1) * sf;
      ra += __shfl_xor_sync(0xffffffffu, ra, 1);
      rb += __shfl_xor_sync(0xffffffffu, rb, 1);
      ra += __shfl_xor_sync(0xffffffffu, ra, 2);
      rb += __shfl_xor_sync(0xffffffffu, rb, 2);
      if (q == 0) {
        const uint64_t ia = byte_a * 4 + f, ib = (byte_a + 8) * 4 + f;
        if (ia < C) o[ia] = ra;
        if (ib < C) o[ib] = rb;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Sparse missing-genotype sums  out[r] = sum_{c in row r} vec[c].
//
// A plain CSR gather is bound by L2 sector traffic (every 8-byte read of `vec`
// pulls a 32-byte sector: 0.28 ms per pass at 7.5e7 entries).  The lists are
// therefore stored column-blocked and row-sliced (SELL-32 per tile):
//   * the gathered vector is cut into tiles of kGatherTile elements that fit in
//     shared memory;
//   * inside a tile, rows are grouped in blocks of 32; a block stores its
//     entries interleaved, slot (i, lane) = i-th entry of row 32*blk + lane, as
//     16-bit in-tile column offsets, padded to the longest row of the block
//     with an index that points at a zero in shared memory.
// A CTA loads one tile of the vector into shared memory; a warp walks a block
// with one coalesced 64-byte load, one LDS and one DADD per 32 entries, each
// lane owning one row's sum (no shuffles, no atomics).  Per-tile partial sums
// are added in tile order by the finalize kernels => bit-reproducible.
// ---------------------------------------------------------------------------
constexpr int kGatherTile = 12288;                        // 96 KB of doubles: 2 CTAs per SM
constexpr int kGatherPad = 64;                            // zero slots behind the tile
constexpr int kGatherSentinel = 0x3030;                   // = memset byte 0x30 twice, inside the pad
constexpr int kGatherSmem = (kGatherTile + kGatherPad) * 8;
constexpr int kGatherThreads = 512;
static_assert(kGatherSentinel >= kGatherTile && kGatherSentinel < kGatherTile + kGatherPad,
              "padding sentinel must point into the zero pad");

// counts[t * nrows + r] = entries of row r whose column falls in tile t (one warp per row)
__global__ void __launch_bounds__(256)
k_bcsr_count(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
             uint64_t nrows, uint32_t* __restrict__ counts) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  for (uint64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32)
    atomicAdd(counts + (uint64_t)(colidx[k] / kGatherTile) * nrows + r, 1u);
}

// sizes[t * nblk + b] = 32 * (longest row of block b in tile t, rounded up to 8 entries -- a lane
// reads its row 8 entries = one 16-byte load at a time); one warp per (t, b)
__global__ void __launch_bounds__(256)
k_sell_sizes(const uint32_t* __restrict__ counts, uint64_t nrows, uint32_t nblk, uint32_t ntiles,
             uint32_t* __restrict__ sizes) {
  uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= (uint64_t)ntiles * nblk) return;
  uint32_t t = (uint32_t)(w / nblk), b = (uint32_t)(w % nblk);
  uint64_t r = (uint64_t)b * 32 + lane;
  uint32_t c = r < nrows ? counts[(uint64_t)t * nrows + r] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
  if (lane == 0) sizes[w] = ((c + 7u) & ~7u) * 32u;
}

// scatter the CSR entries into their SELL-32 slots (col16 pre-filled with the sentinel)
__global__ void __launch_bounds__(256)
k_sell_fill(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
            uint64_t nrows, uint32_t nblk, const uint64_t* __restrict__ blkoff,
            const uint32_t* __restrict__ counts, uint16_t* __restrict__ col16) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const uint64_t b = rowptr[r], e = rowptr[r + 1];
  for (uint64_t k = b + lane; k < e; k += 32) {
    uint32_t c = colidx[k], t = c / kGatherTile;
    uint64_t before = 0;  // entries of this row in earlier tiles (columns ascend within a row)
    for (uint32_t tt = 0; tt < t; tt++) before += counts[(uint64_t)tt * nrows + r];
    const uint64_t rank = k - b - before;
    // group of 8 entries of lane (r & 31): [group][lane][8]
    col16[blkoff[(uint64_t)t * nblk + (r >> 5)] + (rank >> 3) * 256 + (uint64_t)(r & 31) * 8 +
          (rank & 7)] = (uint16_t)(c - t * kGatherTile);
  }
}

// tile t of the gathered vector -> shared memory (THREADS threads; caller synchronises)
template <int THREADS>
__device__ __forceinline__ void sell_load_tile(double* xs, const double* __restrict__ vec,
                                               uint64_t veclen, uint32_t t) {
  const uint64_t base = (uint64_t)t * kGatherTile;
  constexpr int PER = kGatherTile / THREADS;  // loads per thread, issued together
  double tmp[PER];
#pragma unroll
  for (int m = 0; m < PER; m++) {
    const uint32_t i = threadIdx.x + m * THREADS;
    tmp[m] = (base + i < veclen) ? vec[base + i] : 0.0;
  }
#pragma unroll
  for (int m = 0; m < PER; m++) xs[threadIdx.x + m * THREADS] = tmp[m];
  if (threadIdx.x < kGatherPad) xs[kGatherTile + threadIdx.x] = 0.0;
}

// row blocks [b0, b1) of tile t: out[r] = sum of the row's entries in the tile
template <int THREADS>
__device__ __forceinline__ void sell_walk_blocks(const double* xs, const uint64_t* __restrict__ bo,
                                                 const uint16_t* __restrict__ col16, uint32_t b0,
                                                 uint32_t b1, uint64_t nrows,
                                                 double* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = THREADS / 32;
  uint32_t b = b0 + warp;
  uint64_t o = 0, onext = 0;
  if (b < b1) {
    o = bo[b];
    onext = bo[b + 1];
  }
  for (; b < b1; b += NW) {
    const uint32_t width = (uint32_t)((onext - o) >> 5);
    // pointers of the warp's next block, fetched while this one is processed
    uint64_t o2 = 0, o2next = 0;
    if (b + NW < b1) {
      o2 = bo[b + NW];
      o2next = bo[b + NW + 1];
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    // width is a multiple of 8: groups of 8 entries per lane, one 16-byte load each (512 B per
    // warp and load, four loads in flight -- the u16 loads of round 1 kept ~1 KB per warp in
    // flight and left the kernel latency-bound at 2.4 TB/s)
    const uint4* gp = reinterpret_cast<const uint4*>(col16 + o) + lane;
    const uint32_t ngroups = width >> 3;
    for (uint32_t gi = 0; gi < ngroups; gi += 4) {
      uint4 v[4];
#pragma unroll
      for (int m = 0; m < 4; m++)
        v[m] = (gi + m < ngroups) ? gp[(uint64_t)(gi + m) * 32]
                                  : make_uint4(0x30303030u, 0x30303030u, 0x30303030u, 0x30303030u);
#pragma unroll
      for (int m = 0; m < 4; m++) {
        s0 += xs[v[m].x & 0xFFFFu];
        s1 += xs[v[m].x >> 16];
        s2 += xs[v[m].y & 0xFFFFu];
        s3 += xs[v[m].y >> 16];
        s0 += xs[v[m].z & 0xFFFFu];
        s1 += xs[v[m].z >> 16];
        s2 += xs[v[m].w & 0xFFFFu];
        s3 += xs[v[m].w >> 16];
      }
    }
    const uint64_t r = (uint64_t)b * 32 + lane;
    if (r < nrows) out[r] = (s0 + s1) + (s2 + s3);
    o = o2;
    onext = o2next;
  }
}

// grid (ntiles, chunks of row blocks); partial[t * nrows + r] = sum over tile t of row r
__global__ void __launch_bounds__(kGatherThreads, 2)
k_sell_gather(const uint64_t* __restrict__ blkoff, const uint16_t* __restrict__ col16,
              const double* __restrict__ vec, uint64_t veclen, uint64_t nrows, uint32_t nblk,
              uint32_t blocks_per_cta, double* __restrict__ partial) {
  extern __shared__ double xs[];
  const uint32_t t = blockIdx.x;
  sell_load_tile<kGatherThreads>(xs, vec, veclen, t);
  __syncthreads();
  const uint32_t b0 = blockIdx.y * blocks_per_cta, b1 = min(nblk, b0 + blocks_per_cta);
  sell_walk_blocks<kGatherThreads>(xs, blkoff + (uint64_t)t * nblk, col16, b0, b1, nrows,
                                   partial + (uint64_t)t * nrows);
}

// The same sums on a few dedicated SMs: `gridDim.x` CTAs of 1024 threads (launched with the SM's
// whole shared memory, so nothing shares the SM) walk contiguous ranges of the (tile, chunk)
// items while the contraction kernel -- whose CTAs cannot share an SM with a gather CTA -- runs on
// all the others.  The op's critical path then holds no gather at all; same summation order as
// k_sell_gather (bit-identical results).
constexpr int kGatherThreadsP = 1024;
__global__ void __launch_bounds__(kGatherThreadsP, 1)
k_sell_gather_p(const uint64_t* __restrict__ blkoff, const uint16_t* __restrict__ col16,
                const double* __restrict__ vec, uint64_t veclen, uint64_t nrows, uint32_t nblk,
                uint32_t blocks_per_item, uint32_t chunks, uint32_t nitems,
                double* __restrict__ partial) {
  extern __shared__ double xs[];
  const uint32_t per = (nitems + gridDim.x - 1) / gridDim.x;
  const uint32_t i0 = blockIdx.x * per, i1 = min(nitems, i0 + per);
  uint32_t cur = ~0u;
  for (uint32_t item = i0; item < i1; item++) {
    const uint32_t t = item / chunks, c = item - t * chunks;
    if (t != cur) {
      __syncthreads();  // the previous item's reads of the tile
      sell_load_tile<kGatherThreadsP>(xs, vec, veclen, t);
      __syncthreads();
      cur = t;
    }
    const uint32_t b0 = c * blocks_per_item, b1 = min(nblk, b0 + blocks_per_item);
    sell_walk_blocks<kGatherThreadsP>(xs, blkoff + (uint64_t)t * nblk, col16, b0, b1, nrows,
                                      partial + (uint64_t)t * nrows);
  }
}

// Finalise X'x for SNP j:
//   E_j = delta * sum_splits part[s][j];   t_j = [(E_j - 3 Mx_j) - mu_j (Sx - Mx_j)] * inv_sd_j
// t_out (optional) receives t_j; a_out (optional) receives a_j = t_j * inv_sd_j and
// corr_j = b_j - 3 a_j with b_j = mu_j a_j (inputs of the second half of perform_op),
// plus per-block partials of max|a| and sum b.
// mxv: mx_tiles x nsnps per-tile partial sums of Mx (null when nothing is missing).
__global__ void __launch_bounds__(256)
k_finalize_crossprod(const double* __restrict__ part, uint32_t nsplits, uint64_t stride,
                     uint32_t nsnps, const VecScale* __restrict__ sc,
                     const double2* __restrict__ scale, const double* __restrict__ mxv,
                     uint32_t mx_tiles, double* __restrict__ t_out, double* __restrict__ a_out,
                     double* __restrict__ corr_out, double* __restrict__ pmax_a,
                     double* __restrict__ psum_b) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  double abs_a = 0.0, bj = 0.0;
  if (j < nsnps) {
  double mx = 0.0;
  if (mxv)
    for (uint32_t tt = 0; tt < mx_tiles; tt++) mx += mxv[(uint64_t)tt * nsnps + j];
  double e = 0.0;
  for (uint32_t s = 0; s < nsplits; s++) e += part[(uint64_t)s * stride + j];
  e *= sc->delta;
  double2 ms = scale[j];
  double t = ((e - 3.0 * mx) - ms.x * (sc->sum - mx)) * ms.y;
  const bool dead = ms.y == 0.0;  // monomorphic / undefined SNP: zero column (data.cpp:300)
  if (dead) t = 0.0;
  if (t_out) t_out[j] = t;
  if (a_out) {
    double a = dead ? 0.0 : t * ms.y, b = dead ? 0.0 : ms.x * a;
    a_out[j] = a;
    corr_out[j] = b - 3.0 * a;
    abs_a = fabs(a);
    bj = b;
  }
  }
  // (max|a|, sum b) block partials: the scale of the second half's input vector
  if (a_out) emit_block_partials(abs_a, bj, pmax_a, psum_b);
}

// Inputs of prod from a user vector v:  a_j = v_j inv_sd_j, b_j = mu_j a_j, corr_j = b_j - 3 a_j
__global__ void __launch_bounds__(256)
k_prod_inputs(const double* __restrict__ v, const double2* __restrict__ scale, uint32_t nsnps,
              double* __restrict__ a_out, double* __restrict__ corr_out,
              double* __restrict__ pmax_a, double* __restrict__ psum_b) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (j < nsnps) {
    double2 ms = scale[j];
    a = v[j] * ms.y;
    b = ms.x * a;
    if (ms.y == 0.0) { a = 0.0; b = 0.0; }
    a_out[j] = a;
    corr_out[j] = b - 3.0 * a;
  }
  emit_block_partials(fabs(a), b, pmax_a, psum_b);
}

// Finalise X v for individual i:
//   y_i = delta_a * sum_splits part[s][i] - Sb + mc_i,  mc_i = sum_{j missing} corr_j
// (sc_ab holds the step of a and the sum of b)
__global__ void __launch_bounds__(256)
k_finalize_prod(const double* __restrict__ part, uint32_t nsplits, uint64_t stride, uint64_t n,
                const VecScale* __restrict__ sc_ab, const double* __restrict__ mcv,
                uint32_t mc_tiles, double* __restrict__ y) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double f = 0.0;
  for (uint32_t s = 0; s < nsplits; s++) f += part[(uint64_t)s * stride + i];
  double mc = 0.0;
  if (mcv)
    for (uint32_t tt = 0; tt < mc_tiles; tt++) mc += mcv[(uint64_t)tt * n + i];
  y[i] = f * sc_ab->delta - sc_ab->sum + mc;  // delta of a, sum of b
}

}  // namespace fpb
