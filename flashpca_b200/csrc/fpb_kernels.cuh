// fpb_kernels.cuh -- staging, statistics and the generic FP64 kernels.
//
// HBM layout ("staged genotypes").  Two packed 2-bit copies of the same matrix:
//   gs : SNP-major,        nsnps rows x pitch_s bytes (16 individuals per 32-bit word)
//   gi : individual-major, N rows     x pitch_i bytes (16 SNPs per 32-bit word)
// Pitches are multiples of 64 B.  Fields hold DOSAGE codes, a bijective recode
// of the PLINK .bed codes done once at staging (decode_plink, data.cpp:65-126):
//   PLINK 00 -> e=2   PLINK 10 -> e=1   PLINK 11 -> e=0   PLINK 01 (missing) -> e=3
// i.e. e = copies of the minor allele, 3 = missing.  Slots beyond the matrix
// edge (bed pad bits, row padding) hold e=0 and are neutralised by zero inputs.
//
// Standardised value of a code for SNP j with mean mu, sd s (data.cpp:316-319):
//   e=0 -> (0-mu)/s   e=1 -> (1-mu)/s   e=2 -> (2-mu)/s   e=3 -> 0
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fpb {

constexpr double kVarTol = 1e-9;  // util.h:33 VAR_TOL

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream_u64(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u128(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// PLINK code -> dosage code, all 16 fields of a word at once:
//   e_hi = ~c_hi,  e_lo = c_lo ^ c_hi
__host__ __device__ __forceinline__ uint32_t plink_to_dosage(uint32_t w) {
  uint32_t hi = w & 0xAAAAAAAAu, lo = w & 0x55555555u;
  return (~hi & 0xAAAAAAAAu) | ((lo ^ (hi >> 1)) & 0x55555555u);
}
// inverse:  c_hi = ~e_hi,  c_lo = e_lo ^ c_hi
__host__ __device__ __forceinline__ uint32_t dosage_to_plink(uint32_t e) {
  uint32_t chi = ~e & 0xAAAAAAAAu;
  return chi | ((e ^ (chi >> 1)) & 0x55555555u);
}

// ---------------------------------------------------------------------------
// Staging step 1: recode raw bed rows (pitch_s bytes, first np bytes valid) in
// place to dosage codes and zero every slot with individual index >= N.
// One thread per 32-bit word.
// ---------------------------------------------------------------------------
__global__ void k_recode_rows(uint8_t* __restrict__ g, uint64_t nrows, uint64_t n,
                              uint64_t pitch) {
  uint64_t words_per_row = pitch / 4;
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= nrows * words_per_row) return;
  uint64_t w = idx % words_per_row;
  uint32_t* p = reinterpret_cast<uint32_t*>(g) + idx;
  uint64_t first = w * 16;  // first individual in this word
  uint32_t v = 0;
  if (first < n) {
    v = plink_to_dosage(*p);
    uint64_t valid = n - first;
    if (valid < 16) v &= (1u << (2 * valid)) - 1u;
  }
  *p = v;
}

__global__ void k_decode_rows(const uint8_t* __restrict__ g, uint8_t* __restrict__ out,
                              uint64_t nrows, uint64_t np, uint64_t pitch) {
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= nrows * np) return;
  uint64_t r = idx / np, b = idx - r * np;
  out[idx] = (uint8_t)dosage_to_plink(g[r * pitch + b]);
}

// ---------------------------------------------------------------------------
// Staging step 2: 2-bit transpose, gs (rows = SNPs) -> gi (rows = individuals).
// A CTA moves a 128 x 128 tile through shared memory (one byte per genotype).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_transpose_2bit(const uint8_t* __restrict__ src, uint64_t src_rows, uint64_t src_pitch,
                 uint8_t* __restrict__ dst, uint64_t dst_rows, uint64_t dst_pitch) {
  __shared__ uint8_t tile[128][128 + 4];  // [src row][src col]
  const uint64_t r0 = (uint64_t)blockIdx.y * 128, c0 = (uint64_t)blockIdx.x * 128;
  // load: 128 rows x 32 bytes
  for (int k = 0; k < 16; k++) {
    int rr = (threadIdx.x >> 5) + 8 * k, bb = threadIdx.x & 31;
    uint64_t r = r0 + rr, byte = c0 / 4 + bb;
    uint8_t v = (r < src_rows && byte < src_pitch) ? src[r * src_pitch + byte] : 0;
#pragma unroll
    for (int q = 0; q < 4; q++) tile[rr][4 * bb + q] = (v >> (2 * q)) & 3;
  }
  __syncthreads();
  // store: dst row = src col; 128 dst rows x 32 bytes
  for (int k = 0; k < 16; k++) {
    int cc = (threadIdx.x >> 5) + 8 * k, bb = threadIdx.x & 31;
    uint64_t r = c0 + cc, byte = r0 / 4 + bb;
    if (r < dst_rows && byte < dst_pitch) {
      uint8_t v = 0;
#pragma unroll
      for (int q = 0; q < 4; q++) v |= (uint8_t)(tile[4 * bb + q][cc] << (2 * q));
      dst[r * dst_pitch + byte] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Synthetic genotypes generated in place as RAW PLINK bytes (recoded afterwards
// like a file); mirrored bit for bit by flashpca_b200/synth.py.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void k_synth_bed(uint8_t* __restrict__ bed, uint64_t nsnps, uint64_t n, uint64_t pitch,
                            uint64_t snp_offset, const uint8_t* __restrict__ pop,
                            const uint32_t* __restrict__ thr, uint32_t miss_thr, uint64_t seed) {
  uint64_t np = (n + 3) / 4;
  uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= nsnps * np) return;
  uint64_t j = idx / np, b = idx - j * np;
  uint64_t gj = j + snp_offset;
  uint8_t out = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint64_t i = 4 * b + q;
    uint8_t code = 0;  // pad bits are 0, as in a PLINK-written bed
    if (i < n) {
      uint64_t h = mix64(seed ^ mix64(gj * 0x100000001B3ull + i));
      uint32_t u1 = (uint32_t)h, u2 = (uint32_t)(h >> 32);
      uint64_t h2 = mix64(h ^ 0xD6E8FEB86659FD93ull);
      uint32_t um = (uint32_t)h2;
      uint32_t t = thr[(uint64_t)pop[i] * nsnps + j];
      int g = (u1 < t) + (u2 < t);  // copies of the minor allele
      // dosage 2 -> 00, 1 -> 10 (binary), 0 -> 11, missing -> 01 (data.cpp:41-45)
      code = (g == 2) ? 0 : (g == 1 ? 2 : 3);
      if (um < miss_thr) code = 1;
    }
    out |= (uint8_t)(code << (2 * q));
  }
  bed[j * pitch + b] = out;
}

// ---------------------------------------------------------------------------
// First-visit statistics (data.cpp:257-322), one warp per SNP row of gs.
// Codes are counted with popcounts; mean = (n1 + 2 n2) / (n0+n1+n2) is the same
// double the reference gets from summing dosages (every partial sum is an
// exact integer).  snp[j] = (mean, 1/sd or 0, l0, l1); lut[j] = (l0, l1, l2, 0)
// indexed by dosage code.  tracej[j] = sum_i X_ij^2 (svdwide.cpp:44-45).
// nmiss[j] = number of missing genotypes of the row (for the CSR build).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_snp_stats(const uint8_t* __restrict__ gs, uint64_t nsnps, uint64_t n, uint64_t pitch,
            int stand_method, int use_preloaded, double* __restrict__ meansd,
            double4* __restrict__ lut, double2* __restrict__ scale,
            double* __restrict__ tracej, uint32_t* __restrict__ nmiss) {
  uint64_t j = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (j >= nsnps) return;
  const uint4* row = reinterpret_cast<const uint4*>(gs + j * pitch);
  uint32_t nvec = (uint32_t)(pitch / 16);
  uint32_t n1 = 0, n2 = 0, n3 = 0;
  for (uint32_t v = lane; v < nvec; v += 32) {
    uint4 q = ld_stream_u128(row + v);
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t lo = w[k] & 0x55555555u, hi = (w[k] >> 1) & 0x55555555u;
      n3 += __popc(lo & hi);
      n2 += __popc(hi & ~lo);
      n1 += __popc(lo & ~hi);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    n3 += __shfl_xor_sync(0xffffffffu, n3, o);
  }
  if (lane != 0) return;
  uint64_t n0 = n - n1 - n2 - n3;  // padding slots are e=0 and not counted
  double mean, sd;
  if (!use_preloaded) {
    uint64_t ngood = n0 + n1 + n2;
    mean = __ddiv_rn((double)(n1 + 2ull * n2), (double)ngood);
    double pf = __ddiv_rn(mean, 2.0);
    double v = __dmul_rn(pf, __dsub_rn(1.0, pf));
    if (stand_method == 3) v = __dmul_rn(__dmul_rn(2.0, pf), __dsub_rn(1.0, pf));
    sd = __dsqrt_rn(v);
    meansd[j] = mean;
    meansd[nsnps + j] = sd;
  } else {
    mean = meansd[j];
    sd = meansd[nsnps + j];
  }
  double4 l = make_double4(0.0, 0.0, 0.0, 0.0);
  double inv = 0.0;
  if (sd > kVarTol) {
    l.x = __ddiv_rn(__dsub_rn(0.0, mean), sd);
    l.y = __ddiv_rn(__dsub_rn(1.0, mean), sd);
    l.z = __ddiv_rn(__dsub_rn(2.0, mean), sd);
    inv = __ddiv_rn(1.0, sd);
  }
  lut[j] = l;
  scale[j] = make_double2(mean, inv);
  tracej[j] = (double)n0 * l.x * l.x + (double)n1 * l.y * l.y + (double)n2 * l.z * l.z;
  nmiss[j] = n3;
}

// CSR fill: column indices of the missing entries of every row, ascending.
// One warp per row; words are scanned in order with a warp prefix sum.
__global__ void __launch_bounds__(256)
k_fill_missing_csr(const uint8_t* __restrict__ g, uint64_t nrows, uint64_t pitch,
                   const uint64_t* __restrict__ rowptr, uint32_t* __restrict__ colidx) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  uint64_t base = rowptr[r];
  if (rowptr[r + 1] == base) return;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(g + r * pitch);
  uint32_t nwords = (uint32_t)(pitch / 4);
  for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
    uint32_t w = w0 + lane;
    uint32_t m = 0;
    if (w < nwords) {
      uint32_t v = row[w];
      m = v & (v >> 1) & 0x55555555u;
    }
    uint32_t cnt = __popc(m), incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    uint64_t off = base + incl - cnt;
    while (m) {
      int bit = __ffs(m) - 1;
      m &= m - 1;
      colidx[off++] = w * 16 + (bit >> 1);
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// CSR transposition helpers (staging): keys = (column << 32 | row) of every entry,
// sorted by a stable radix sort, give the transposed lists in ascending order.
__global__ void __launch_bounds__(256)
k_csr_to_keys(const uint64_t* __restrict__ rowptr, const uint32_t* __restrict__ colidx,
              uint64_t nrows, uint64_t* __restrict__ keys) {
  uint64_t r = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  for (uint64_t k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32)
    keys[k] = ((uint64_t)colidx[k] << 32) | r;
}
__global__ void k_keys_to_csr(const uint64_t* __restrict__ keys, uint64_t nnz,
                              uint32_t* __restrict__ colidx, uint32_t* __restrict__ counts) {
  uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  uint64_t key = keys[k];
  colidx[k] = (uint32_t)key;
  atomicAdd(counts + (key >> 32), 1u);
}

// ---------------------------------------------------------------------------
// Generic FP64 path (any missingness; also the cross-check of the tensor path).
//
// crossprod:  t_j = sum_i X_ij x_i   (svdwide.cpp:122-153; first half of :42)
// Grid (chunks, splits): a CTA owns blockDim.x * 16W consecutive individuals
// (their x values stay in registers) and walks the SNP range of its split.
// Per genotype two bit-masked adds build H = sum x_i [hi bit], L = sum x_i [lo
// bit]; a rare slow path adds M = sum x_i [missing].  With S_e = sum of x over
// genotypes with code e:  S1 = L - M, S2 = H - M, S0 = Stot - L - H + M and
//   t_j partial = l0*S0 + l1*S1 + l2*S2,
// warp-reduced by shuffles and written to part[(chunk, warp)][j]; k_sum_rows adds the
// (chunk, warp) partials of a SNP in a fixed order (no atomics: bit-reproducible).
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_crossprod(const uint8_t* __restrict__ gs, uint64_t pitch, uint64_t n, uint32_t nsnps,
            uint32_t snps_per_split, const double* __restrict__ x,
            const double4* __restrict__ lut, double* __restrict__ part) {
  const uint32_t words_per_row = (uint32_t)(pitch / 4);
  const uint32_t widx = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  const bool active = widx < words_per_row;
  const uint32_t j0 = blockIdx.y * snps_per_split;
  const uint32_t j1 = min(nsnps, j0 + snps_per_split);

  double xr[16 * W];
  double stot = 0.0;
#pragma unroll
  for (int k = 0; k < 16 * W; k++) {
    uint64_t i = (uint64_t)widx * 16 + k;
    xr[k] = (active && i < n) ? x[i] : 0.0;
    stot += xr[k];
  }
  const uint8_t* base = gs + (uint64_t)widx * 4;

  for (uint32_t j = j0; j < j1; j++) {
    uint32_t w[W];
    if (active) {
      if (W == 1) {
        w[0] = ld_stream_u32(reinterpret_cast<const uint32_t*>(base + (uint64_t)j * pitch));
      } else {
        uint2 v = ld_stream_u64(reinterpret_cast<const uint2*>(base + (uint64_t)j * pitch));
        w[0] = v.x;
        w[W - 1] = v.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < W; q++) w[q] = 0u;
    }
    double hs = 0.0, ls = 0.0, ms = 0.0;
#pragma unroll
    for (int q = 0; q < W; q++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if (w[q] & (2u << (2 * k))) hs += xr[16 * q + k];
        if (w[q] & (1u << (2 * k))) ls += xr[16 * q + k];
      }
      uint32_t m = w[q] & (w[q] >> 1) & 0x55555555u;
      if (m) {
#pragma unroll
        for (int k = 0; k < 16; k++)
          if (m & (1u << (2 * k))) ms += xr[16 * q + k];
      }
    }
    const double4 l = lut[j];
    double tj = l.x * (stot - ls - hs + ms) + l.y * (ls - ms) + l.z * (hs - ms);
    tj = warp_sum(tj);
    if ((threadIdx.x & 31) == 0)
      part[((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * nsnps + j] = tj;
  }
}

// out[c] = sum_r part[r * ncols + c], r ascending (fixed order)
__global__ void k_sum_rows(const double* __restrict__ part, uint32_t nrows, uint64_t ncols,
                           double* __restrict__ out) {
  uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (c >= ncols) return;
  double s = 0.0;
  for (uint32_t r = 0; r < nrows; r++) s += part[(uint64_t)r * ncols + c];
  out[c] = s;
}

// Per-SNP coefficients for the generic prod:  with a_e = l_e * v_j,
//   value(e) = a0 + [lo](a1 - a0) + [hi](a2 - a0)        for e = 0,1,2
// and e = 3 (missing, value 0) is corrected by subtracting (a1 + a2 - a0).
// coef[j] = (a1-a0, a2-a0, a1+a2-a0, a0).
__global__ void k_prod_coef(const double4* __restrict__ lut, const double* __restrict__ v,
                            uint32_t nsnps, double4* __restrict__ coef) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsnps) return;
  double4 l = lut[j];
  double vj = v[j];
  double a0 = l.x * vj, a1 = l.y * vj, a2 = l.z * vj;
  coef[j] = make_double4(a1 - a0, a2 - a0, a1 + a2 - a0, a0);
}

// Deterministic single-block reduction of coef[].w into *c0.
__global__ void __launch_bounds__(1024)
k_sum_a0(const double4* __restrict__ coef, uint32_t nsnps, double* __restrict__ c0) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (uint32_t j = threadIdx.x; j < nsnps; j += 1024) s += coef[j].w;
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *c0 = sh[0];
}

// prod:  y_i = sum_j X_ij v_j   (svdwide.cpp:193-226; second half of :43)
template <int W>
__global__ void __launch_bounds__(256)
k_prod(const uint8_t* __restrict__ gs, uint64_t pitch, uint64_t n, uint32_t nsnps,
       uint32_t snps_per_split, const double4* __restrict__ coef,
       const double* __restrict__ c0, double* __restrict__ y /* [split][n] when gridDim.y > 1 */) {
  const uint32_t words_per_row = (uint32_t)(pitch / 4);
  const uint32_t widx = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  if (widx >= words_per_row) return;
  const uint32_t j0 = blockIdx.y * snps_per_split;
  const uint32_t j1 = min(nsnps, j0 + snps_per_split);

  double acc[16 * W];
  const double init = (blockIdx.y == 0) ? *c0 : 0.0;
#pragma unroll
  for (int k = 0; k < 16 * W; k++) acc[k] = init;
  const uint8_t* base = gs + (uint64_t)widx * 4;

  for (uint32_t j = j0; j < j1; j++) {
    uint32_t w[W];
    if (W == 1) {
      w[0] = ld_stream_u32(reinterpret_cast<const uint32_t*>(base + (uint64_t)j * pitch));
    } else {
      uint2 v = ld_stream_u64(reinterpret_cast<const uint2*>(base + (uint64_t)j * pitch));
      w[0] = v.x;
      w[W - 1] = v.y;
    }
    const double4 c = coef[j];
#pragma unroll
    for (int q = 0; q < W; q++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if (w[q] & (1u << (2 * k))) acc[16 * q + k] += c.x;
        if (w[q] & (2u << (2 * k))) acc[16 * q + k] += c.y;
      }
      uint32_t m = w[q] & (w[q] >> 1) & 0x55555555u;
      if (m) {
#pragma unroll
        for (int k = 0; k < 16; k++)
          if (m & (1u << (2 * k))) acc[16 * q + k] -= c.z;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16 * W; k++) {
    uint64_t i = (uint64_t)widx * 16 + k;
    if (i < n) y[(uint64_t)blockIdx.y * n + i] = acc[k];  // per-split partials, summed in split order
  }
}

}  // namespace fpb
