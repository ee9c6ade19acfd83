// fpb_kernels.cuh -- sm_100a kernels of the FlashPCA2 hot path.
//
// HBM layout ("staged bed"): nsnps rows, one per SNP, `pitch` bytes each
// (pitch = ceil(N/4) rounded up to 16 B so every row starts 16-byte aligned).
// Bytes are raw PLINK .bed bytes: 4 individuals per byte, individual q of a
// byte in bits [2q, 2q+1] (data.cpp:128-148).  Genotype slots at individual
// index >= N (bed pad bits and row padding) are rewritten at staging time to
// code 01 = missing, which standardises to 0 (data.cpp:319), so no hot kernel
// needs tail logic on the genotype side.
//
// Code -> standardised value (data.cpp:316-319), per SNP j with mean mu, sd s:
//   code 0 -> (2-mu)/s   code 1 -> 0 (missing)   code 2 -> (1-mu)/s   code 3 -> (0-mu)/s
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

// ---------------------------------------------------------------------------
// Staging: mark genotype slots with individual index >= N as missing (01).
// One thread per SNP row; touches at most 16 bytes.
// ---------------------------------------------------------------------------
__global__ void k_fix_padding(uint8_t* __restrict__ bed, uint64_t nsnps, uint64_t n,
                              uint64_t pitch) {
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= nsnps) return;
  uint8_t* row = bed + j * pitch;
  uint64_t full = n / 4;  // bytes that hold 4 real individuals
  uint32_t rem = (uint32_t)(n & 3);
  uint64_t b = full;
  if (rem) {
    uint8_t keep = (uint8_t)((1u << (2 * rem)) - 1u);
    row[b] = (uint8_t)((row[b] & keep) | (0x55u & ~keep));
    b++;
  }
  for (; b < pitch; b++) row[b] = 0x55;
}

// ---------------------------------------------------------------------------
// Synthetic genotypes generated in place (bench input path; mirrored bit for
// bit by flashpca_b200/synth.py).  One thread per packed byte.
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
    uint8_t code = 0;  // pad bits stay 0 here, as in a real bed; fixed by k_fix_padding
    if (i < n) {
      uint64_t h = mix64(seed ^ mix64(gj * 0x100000001B3ull + i));
      uint32_t u1 = (uint32_t)h, u2 = (uint32_t)(h >> 32);
      uint64_t h2 = mix64(h ^ 0xD6E8FEB86659FD93ull);
      uint32_t um = (uint32_t)h2;
      uint32_t t = thr[(uint64_t)pop[i] * nsnps + j];
      int g = (u1 < t) + (u2 < t);  // copies of the minor allele
      // dosage 2 -> 00, 1 -> 10 (binary, =2), 0 -> 11, missing -> 01 (data.cpp:41-45)
      code = (g == 2) ? 0 : (g == 1 ? 2 : 3);
      if (um < miss_thr) code = 1;
    }
    out |= (uint8_t)(code << (2 * q));
  }
  bed[j * pitch + b] = out;
}

// ---------------------------------------------------------------------------
// First-visit statistics (data.cpp:257-322), one warp per SNP.
// Counts codes by popcount; mean = (2*n0 + n2) / (n0+n2+n3) is the same
// double the reference gets from summing dosages (all partial sums are exact
// integers).  lut[j] = (l0, l1=0, l2, l3) indexed by raw code.
// tracej[j] = sum_i X_ij^2 (svdwide.cpp:44-45) from the counts.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_snp_stats(const uint8_t* __restrict__ bed, uint64_t nsnps, uint64_t pitch, int stand_method,
            int use_preloaded, double* __restrict__ meansd, double4* __restrict__ lut,
            double* __restrict__ tracej) {
  uint64_t j = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (j >= nsnps) return;
  const uint4* row = reinterpret_cast<const uint4*>(bed + j * pitch);
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
  uint64_t total = pitch * 4;
  uint64_t n0 = total - n1 - n2 - n3;
  double mean, sd;
  if (!use_preloaded) {
    uint64_t ngood = n0 + n2 + n3;
    mean = __ddiv_rn((double)(2 * n0 + n2), (double)ngood);
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
  if (sd > kVarTol) {
    l.w = __ddiv_rn(__dsub_rn(0.0, mean), sd);  // code 3
    l.z = __ddiv_rn(__dsub_rn(1.0, mean), sd);  // code 2
    l.x = __ddiv_rn(__dsub_rn(2.0, mean), sd);  // code 0
  }
  lut[j] = l;
  tracej[j] = (double)n0 * l.x * l.x + (double)n2 * l.z * l.z + (double)n3 * l.w * l.w;
}

// ---------------------------------------------------------------------------
// crossprod:  t_j = sum_i X_ij x_i   (svdwide.cpp:122-153; first half of :42)
//
// Grid (chunks, splits): a CTA owns blockDim.x * 16W consecutive individuals
// (their x values live in registers for the whole kernel) and walks the SNP
// range of its split, reading blockDim.x * 4W contiguous bytes per SNP row.
// Per genotype: two bit-predicated DADDs build H = sum x_i [hi bit] and
// L = sum x_i [lo bit]; a rare slow path adds M = sum x_i [missing].  Then
//   S3 = L - M, S2 = H - L + M, S0 = Stot - M - H
//   t_j partial = l0*S0 + l2*S2 + l3*S3
// is warp-reduced by shuffles and added to t[j] with one FP64 atomic per warp.
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_crossprod(const uint8_t* __restrict__ bed, uint64_t pitch, uint64_t n, uint32_t nsnps,
            uint32_t snps_per_split, const double* __restrict__ x,
            const double4* __restrict__ lut, double* __restrict__ t) {
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
  const uint8_t* base = bed + (uint64_t)widx * 4;

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
      for (int q = 0; q < W; q++) w[q] = 0x55555555u;
    }
    double hs = 0.0, ls = 0.0, ms = 0.0;
#pragma unroll
    for (int q = 0; q < W; q++) {
#pragma unroll
      for (int k = 0; k < 16; k++) {
        if (w[q] & (2u << (2 * k))) hs += xr[16 * q + k];
        if (w[q] & (1u << (2 * k))) ls += xr[16 * q + k];
      }
      uint32_t m = w[q] & ~(w[q] >> 1) & 0x55555555u;
      if (m) {
#pragma unroll
        for (int k = 0; k < 16; k++)
          if (m & (1u << (2 * k))) ms += xr[16 * q + k];
      }
    }
    const double4 l = lut[j];
    double tj = l.x * (stot - ms - hs) + l.z * (hs - ls + ms) + l.w * (ls - ms);
    tj = warp_sum(tj);
    if ((threadIdx.x & 31) == 0) atomicAdd(t + j, tj);
  }
}

// ---------------------------------------------------------------------------
// Per-SNP coefficients for prod:  with a_c = l_c * v_j (c = raw code),
//   value(code) = a0 + [hi](a2 - a0) + [lo](a3 - a2)        for codes 0,2,3
// and code 1 (missing, value 0) is corrected by subtracting (a0 + a3 - a2).
// coef[j] = (b, g, cm, a0);  c0 = sum_j a0 is reduced deterministically.
// ---------------------------------------------------------------------------
__global__ void k_prod_coef(const double4* __restrict__ lut, const double* __restrict__ v,
                            uint32_t nsnps, double4* __restrict__ coef) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nsnps) return;
  double4 l = lut[j];
  double vj = v[j];
  double a0 = l.x * vj, a2 = l.z * vj, a3 = l.w * vj;
  coef[j] = make_double4(a2 - a0, a3 - a2, a0 + a3 - a2, a0);
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

// ---------------------------------------------------------------------------
// prod:  y_i = sum_j X_ij v_j   (svdwide.cpp:193-226; second half of :43)
// Same tiling as crossprod; a thread keeps 16W y accumulators in registers
// and adds b_j / g_j under the hi / lo bit of each genotype.  Split 0 adds the
// constant c0; results are merged with FP64 atomics when splits > 1.
// ---------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256)
k_prod(const uint8_t* __restrict__ bed, uint64_t pitch, uint64_t n, uint32_t nsnps,
       uint32_t snps_per_split, const double4* __restrict__ coef,
       const double* __restrict__ c0, double* __restrict__ y) {
  const uint32_t words_per_row = (uint32_t)(pitch / 4);
  const uint32_t widx = (blockIdx.x * blockDim.x + threadIdx.x) * W;
  if (widx >= words_per_row) return;
  const uint32_t j0 = blockIdx.y * snps_per_split;
  const uint32_t j1 = min(nsnps, j0 + snps_per_split);

  double acc[16 * W];
  const double init = (blockIdx.y == 0) ? *c0 : 0.0;
#pragma unroll
  for (int k = 0; k < 16 * W; k++) acc[k] = init;
  const uint8_t* base = bed + (uint64_t)widx * 4;

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
        if (w[q] & (2u << (2 * k))) acc[16 * q + k] += c.x;
        if (w[q] & (1u << (2 * k))) acc[16 * q + k] += c.y;
      }
      uint32_t m = w[q] & ~(w[q] >> 1) & 0x55555555u;
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
    if (i < n) {
      if (gridDim.y == 1) y[i] = acc[k];
      else atomicAdd(y + i, acc[k]);
    }
  }
}

}  // namespace fpb
