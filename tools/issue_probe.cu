// issue_probe.cu -- probes for the open questions of the fused single-read kernel
// (DESIGN.md section 4.7 / 8).  Everything is per SM, measured with clock64()
// inside one CTA per SM; no memory traffic except where stated.
//
//   1. imma_chain   : latency of a dependent mma.sync.m16n8k32.u8.s8 chain, and throughput with
//                     C independent chains per warp and W warps per scheduler.
//   2. decode_mix   : the contraction kernels' instruction mix -- 3 LOP3 + 1 IMMA per A operand
//                     group (cumulative masks) -- with W warps per scheduler: the compute ceiling
//                     of any LOP3 + mma.sync design, in genotypes per clock per SM, for one half
//                     ("x1") and for both halves' worth of work on the same bytes ("x2").
//   3. lds_width    : shared-memory read throughput with 4-, 8- and 16-byte loads per lane
//                     (the second fused structure used LDS.32).
//   4. tmem_park    : tcgen05.alloc / st / ld round trip of 16 registers per thread (parking the
//                     second-half accumulators in TMEM between tiles): checks the data and times it.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_probe tools/issue_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x);              \
      return 1;                                                                 \
    }                                                                           \
  } while (0)

__device__ __forceinline__ void mma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                    uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- 1. IMMA chains ---------------------------------------------------------
template <int CHAINS>
__global__ void k_imma_chain(int* out, long long* cyc, int iters) {
  int acc[CHAINS][4] = {};
  const uint32_t a = threadIdx.x * 0x01010101u, b = 0x01020304u;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) mma(acc[c], a, a + 1, a + 2, a + 3, b, b + c);
  }
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 2. decode + IMMA mix ---------------------------------------------------
// One "group" = one packed word pair (rows g, g+8): 6 LOP3 + 2 IMMA = 32 genotypes x 16 rows...
// per warp: 2 IMMAs cover 16 rows x 64 genotype columns = 1024 genotypes.  REPS = how many IMMA
// sets consume the same masked operands (1 = one half, 2 = the two-vector kernels' ratio).
template <int REPS>
__global__ void k_decode_mix(int* out, long long* cyc, const uint32_t* words, int iters) {
  int acc[REPS][4][4] = {};
  uint32_t xa = words[threadIdx.x], xb = words[threadIdx.x + blockDim.x];
  const uint32_t b0 = 0x01020304u, b1 = 0x04030201u;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {  // 4 independent word pairs per iteration
      const uint32_t wa = xa + u * 0x9E3779B9u, wb = xb ^ (u * 0x85EBCA6Bu);
      const uint32_t m0 = wa & 0x03030303u, m1 = wb & 0x03030303u, m2 = wa & 0x0F0F0F0Fu,
                     m3 = wb & 0x0F0F0F0Fu, m4 = wa & 0x3F3F3F3Fu, m5 = wb & 0x3F3F3F3Fu;
#pragma unroll
      for (int r = 0; r < REPS; r++) {
        mma(acc[r][u], m0, m1, m2, m3, b0 + r, b1);
        mma(acc[r][u], m4, m5, wa, wb, b1 + r, b0);
      }
    }
    xa = xa * 1664525u + 1013904223u;
    xb = xb * 22695477u + 1u;
  }
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int r = 0; r < REPS; r++)
#pragma unroll
    for (int u = 0; u < 4; u++) s += acc[r][u][0] + acc[r][u][1] + acc[r][u][2] + acc[r][u][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 3. LDS width -----------------------------------------------------------
template <int BYTES>
__global__ void k_lds_width(uint32_t* out, long long* cyc, int iters) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t s = 0;
  const uint32_t lane_off = (threadIdx.x & 31) * BYTES + (threadIdx.x >> 5) * 2048;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t addr = base + ((lane_off + u * 32 * BYTES + i * 64) & 0xFFFF & ~(BYTES - 1));
      if (BYTES == 4) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
        s += v;
      } else if (BYTES == 8) {
        uint32_t v0, v1;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(addr));
        s += v0 ^ v1;
      } else {
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                     : "r"(addr));
        s += v0 ^ v1 ^ v2 ^ v3;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 4. TMEM parking --------------------------------------------------------
// 4 warps (one per TMEM lane quadrant): allocate 64 columns, then ITERS times store 16 registers
// per thread to 16 columns, load them back from the previous slot, accumulate.  Verifies that a
// thread reads back what it stored (lane = 32 * (warp % 4) + lane id, 32x32b shape).
__global__ void __launch_bounds__(128, 1)
k_tmem_park(int* out, int* errs, long long* cyc, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(dst)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant
  int r[16], bad = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) r[k] = threadIdx.x * 100 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    const uint32_t col = (uint32_t)((i & 3) * 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
        "%15,%16};" ::"r"(taddr + col),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    int q[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, "
        "[%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]),
          "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]),
          "=r"(q[14]), "=r"(q[15])
        : "r"(taddr + col)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 16; k++) {
      bad += q[k] != r[k];
      r[k] = q[k] + 1;
    }
  }
  const long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += r[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  atomicAdd(errs, bad);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
  (void)lane;
}

static double mean_cycles(const long long* h, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += (double)h[i];
  return s / n;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  int* d_out;
  uint32_t* d_words;
  long long* d_cyc;
  int* d_err;
  CK(cudaMalloc(&d_out, sizeof(int) * sms * 1024));
  CK(cudaMalloc(&d_words, sizeof(uint32_t) * 4096));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_words, 0x5A, sizeof(uint32_t) * 4096));
  long long* h_cyc = new long long[sms];
  const int iters = 4000;

  // 1. IMMA chains: tpb = 32 * 4 * W (W warps per scheduler)
  for (int W = 1; W <= 4; W++) {
    const int tpb = 128 * W;
#define RUN_CHAIN(C)                                                                        \
  {                                                                                         \
    k_imma_chain<C><<<sms, tpb>>>(d_out, d_cyc, iters);                                     \
    CK(cudaDeviceSynchronize());                                                            \
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));          \
    const double c = mean_cycles(h_cyc, sms);                                               \
    printf("imma_chain W=%d chains=%d: %.1f clk per IMMA per warp, %.0f MAC/clk/SM\n", W, C, \
           c / (iters * (double)C), 4096.0 * C * (tpb / 32) * iters / c);                   \
  }
    RUN_CHAIN(1) RUN_CHAIN(2) RUN_CHAIN(4) RUN_CHAIN(8)
  }
  // 2. decode mix
  for (int W = 1; W <= 4; W++) {
    const int tpb = 128 * W;
#define RUN_MIX(R)                                                                              \
  {                                                                                             \
    k_decode_mix<R><<<sms, tpb>>>(d_out, d_cyc, d_words, iters);                                \
    CK(cudaDeviceSynchronize());                                                                \
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));              \
    const double c = mean_cycles(h_cyc, sms);                                                   \
    /* per iteration and warp: 4 word pairs x 16 rows x 64 columns = 4096 genotypes decoded */ \
    printf("decode_mix W=%d imma-sets=%d: %.1f genotypes decoded/clk/SM, %.0f MAC/clk/SM "       \
           "(HBM delivers ~100 genotypes/clk/SM)\n",                                            \
           W, R, 4096.0 * (tpb / 32) * iters / c, 8.0 * R * 4096.0 * (tpb / 32) * iters / c);   \
  }
    RUN_MIX(1) RUN_MIX(2)
  }
  // 3. LDS width
  CK(cudaFuncSetAttribute(k_lds_width<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_width<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds_width<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
#define RUN_LDS(B)                                                                              \
  {                                                                                             \
    k_lds_width<B><<<sms, 256, 65536>>>((uint32_t*)d_out, d_cyc, iters);                        \
    CK(cudaDeviceSynchronize());                                                                \
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));              \
    const double c = mean_cycles(h_cyc, sms);                                                   \
    printf("lds_width %2d B/lane: %.1f B/clk/SM, %.2f warp-instructions/clk/SM\n", B,           \
           8.0 * 256 * B * iters / c, 8.0 * 8 * iters / c);                                     \
  }
  RUN_LDS(4) RUN_LDS(8) RUN_LDS(16)
  // 4. TMEM parking
  {
    CK(cudaMemset(d_err, 0, sizeof(int)));
    k_tmem_park<<<sms, 128>>>(d_out, d_err, d_cyc, 2000);
    CK(cudaDeviceSynchronize());
    int bad = 0;
    CK(cudaMemcpy(&bad, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    const double c = mean_cycles(h_cyc, sms);
    printf("tmem_park: %d mismatches; %.1f clk per st+wait+ld+wait round trip of 16 regs/thread "
           "(4 warps)\n", bad, c / 2000.0);
  }
  return 0;
}
