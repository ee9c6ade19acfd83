// tcgen05_mix.cu -- issue-rate ceiling of the tcgen05 formulation of the contraction
// (DESIGN.md section 8): every thread owns one matrix row, decodes packed words with the
// cumulative masks (3 LOP3 per word), stores the decoded words to the A operand in TMEM
// (tcgen05.st.32x32b.x16) and one thread per warpgroup issues tcgen05.mma.kind::i8
// (M = 128, N = 8, K = 32, A from TMEM) on them.  No memory traffic: the packed words are
// generated in registers.  WG warpgroups per CTA each run their own stream (own TMEM columns,
// own mbarriers; the A staging area is double buffered).  Prints genotypes per clock per SM;
// HBM delivers ~100, both halves of perform_op need 2 x that.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_mix tools/tcgen05_mix.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x)                                                      \
  do {                                                             \
    cudaError_t e_ = (x);                                          \
    if (e_ != cudaSuccess) {                                       \
      printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); \
      return 1;                                                    \
    }                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
constexpr uint32_t kIdesc = (2u << 4) | (1u << 10) | (1u << 17) | (8u << 24);  // S32, u8 x s8, N=8, M=128

// per warpgroup: columns [0, 8) D, [64, 96) and [96, 128) the two A staging buffers (32 columns = 4 MMAs)
template <int WG>
__global__ void __launch_bounds__(128 * WG, 1)
k_mix(int* out, long long* cyc, int iters) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long bars[WG][2];
  __shared__ __align__(128) int8_t sB[256];  // one K = 32 x N = 8 block, reused by every MMA
  const int tid = threadIdx.x, warp = tid >> 5, wg = warp >> 2, lane_id = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 2 * WG) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[tid >> 1][tid & 1])));
  }
  for (int i = tid; i < 256; i += blockDim.x) sB[i] = (int8_t)((i * 7) % 127 - 63);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s + (uint32_t)wg * 128u;               // this warpgroup's columns
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint64_t bdesc = 0;
  bdesc |= (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF);
  bdesc |= (uint64_t)(128 >> 4) << 16;
  bdesc |= (uint64_t)(256 >> 4) << 32;
  bdesc |= (uint64_t)1 << 46;
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; j++) w[j] = tid * 2654435761u + j * 40503u;
  const bool issuer = (tid & 127) == 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int buf = it & 1;
    // the MMAs that read this staging buffer two iterations ago must be done
    if (it >= 2) {
      const uint32_t bar = smem_u32(&bars[wg][buf]);
      const uint32_t parity = (uint32_t)((it >> 1) - 1) & 1u;
      for (uint32_t spin = 0; !try_wait(bar, parity) && spin < (1u << 20); spin++) {
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t tA = tbase + 64u + (uint32_t)buf * 32u + lane_base;
#pragma unroll
    for (int h = 0; h < 2; h++) {  // 4 packed words -> 16 decoded words -> one STTM.x16
      uint32_t d[16];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x = w[h * 4 + j];
        d[4 * j + 0] = x & 0x03030303u;
        d[4 * j + 1] = x & 0x0F0F0F0Fu;
        d[4 * j + 2] = x & 0x3F3F3F3Fu;
        d[4 * j + 3] = x;
        w[h * 4 + j] = x * 1664525u + 1013904223u;
      }
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
          "%15,%16};" ::"r"(tA + (uint32_t)h * 16u),
          "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
          "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // warpgroup-wide: every row of the 128 x 32-column block is stored
    asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
    if (issuer) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tAm = tbase + 64u + (uint32_t)buf * 32u;
#pragma unroll
      for (int k = 0; k < 4; k++) {  // 32 columns = 4 x (K = 32 bytes = 8 columns)
        const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tbase),
            "r"(tAm + (uint32_t)k * 8u), "l"(bdesc), "r"(kIdesc), "r"(acc), "r"(0u)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&bars[wg][buf]))
                   : "memory");
    }
  }
  // drain: wait for the last commit on each staging buffer
  {
    const int n0 = (iters + 1) / 2, n1 = iters / 2;  // commits on buffer 0 / 1
    if (n0 > 0)
      for (uint32_t spin = 0; !try_wait(smem_u32(&bars[wg][0]), (uint32_t)(n0 - 1) & 1u) && spin < (1u << 22); spin++) {
      }
    if (n1 > 0)
      for (uint32_t spin = 0; !try_wait(smem_u32(&bars[wg][1]), (uint32_t)(n1 - 1) & 1u) && spin < (1u << 22); spin++) {
      }
  }
  const long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  int dsum[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(dsum[0]), "=r"(dsum[1]), "=r"(dsum[2]), "=r"(dsum[3]), "=r"(dsum[4]),
                 "=r"(dsum[5]), "=r"(dsum[6]), "=r"(dsum[7])
               : "r"(tbase + lane_base)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  out[blockIdx.x * blockDim.x + tid] = dsum[0] ^ dsum[3] ^ dsum[7] ^ (int)w[0];
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s) : "memory");
  (void)lane_id;
}


// ---- deeper pipeline: ring of NBUF staging buffers (16 columns = one STTM.x16 = two MMAs each), no
// warpgroup-wide barrier: each decoding warp arrives on the buffer's "filled" mbarrier (count 4),
// a dedicated issuer warp (one lane, all warpgroups round-robin) issues the MMAs and commits them to
// the buffer's "free" mbarrier, which the decoders wait on before they overwrite the buffer.
template <int WG, int NBUF>
__global__ void __launch_bounds__(128 * WG + 32, 1)
k_mix_ring(int* out, long long* cyc, int iters /* buffers filled per warpgroup */) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long filled[WG][NBUF], freeb[WG][NBUF];
  __shared__ __align__(128) int8_t sB[256];
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool is_issuer = warp == 4 * WG;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < WG * NBUF) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&filled[tid / NBUF][tid % NBUF])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&freeb[tid / NBUF][tid % NBUF])));
  }
  for (int i = tid; i < 256; i += blockDim.x) sB[i] = (int8_t)((i * 7) % 127 - 63);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem0 = tmem_base_s;
  uint64_t bdesc = 0;
  bdesc |= (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF);
  bdesc |= (uint64_t)(128 >> 4) << 16;
  bdesc |= (uint64_t)(256 >> 4) << 32;
  bdesc |= (uint64_t)1 << 46;
  const long long t0 = clock64();
  if (is_issuer) {
    if ((tid & 31) == 0) {
      for (int it = 0; it < iters; it++) {
        const int b = it % NBUF;
        const uint32_t parity = (uint32_t)(it / NBUF) & 1u;
        for (int wg = 0; wg < WG; wg++) {
          for (uint32_t spin = 0; !try_wait(smem_u32(&filled[wg][b]), parity) && spin < (1u << 22); spin++) {
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tb = tmem0 + (uint32_t)wg * 128u;
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tb),
                "r"(tb + 16u + (uint32_t)b * 16u + (uint32_t)k * 8u), "l"(bdesc), "r"(kIdesc), "r"(acc),
                "r"(0u)
                : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&freeb[wg][b]))
                       : "memory");
        }
      }
      // wait for the last commit of every warpgroup
      for (int wg = 0; wg < WG; wg++) {
        const int it = iters - 1, b = it % NBUF;
        for (uint32_t spin = 0;
             !try_wait(smem_u32(&freeb[wg][b]), (uint32_t)(it / NBUF) & 1u) && spin < (1u << 22); spin++) {
        }
      }
    }
  } else {
    const int wg = warp >> 2;
    const uint32_t tb = tmem0 + (uint32_t)wg * 128u;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = tid * 2654435761u + j * 40503u;
    for (int it = 0; it < iters; it++) {
      const int b = it % NBUF;
      if (it >= NBUF) {  // the MMAs that read this buffer NBUF fills ago must be done
        const uint32_t parity = (uint32_t)(it / NBUF - 1) & 1u;
        for (uint32_t spin = 0; !try_wait(smem_u32(&freeb[wg][b]), parity) && spin < (1u << 22); spin++) {
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      uint32_t d[16];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t x = w[j];
        d[4 * j + 0] = x & 0x03030303u;
        d[4 * j + 1] = x & 0x0F0F0F0Fu;
        d[4 * j + 2] = x & 0x3F3F3F3Fu;
        d[4 * j + 3] = x;
        w[j] = x * 1664525u + 1013904223u;
      }
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
          "%15,%16};" ::"r"(tb + 16u + (uint32_t)b * 16u + lane_base),
          "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]),
          "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]), "r"(d[14]), "r"(d[15])
          : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&filled[wg][b])) : "memory");
    }
    out[blockIdx.x * blockDim.x + tid] = (int)w[0];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem0) : "memory");
}

template <int WG, int NBUF>
static int run_ring(int sms, int* d_out, long long* d_cyc, long long* h_cyc, int iters) {
  k_mix_ring<WG, NBUF><<<sms, 128 * WG + 32>>>(d_out, d_cyc, iters);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double c = 0;
  for (int i = 0; i < sms; i++) c += (double)h_cyc[i];
  c /= sms;
  const double geno = 128.0 * 4 * 16 * WG * iters;  // per fill: 128 rows x 4 packed words x 16 genotypes
  printf("tcgen05_mix ring WG=%d NBUF=%d: %.1f genotypes/clk/SM (%.0f MAC/clk/SM), %.1f clk per MMA\n", WG,
         NBUF, geno / c, 8.0 * geno / c, c / (2.0 * WG * iters));
  return 0;
}

template <int WG>
static int run(int sms, int* d_out, long long* d_cyc, long long* h_cyc, int iters) {
  k_mix<WG><<<sms, 128 * WG>>>(d_out, d_cyc, iters);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h_cyc, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double c = 0;
  for (int i = 0; i < sms; i++) c += (double)h_cyc[i];
  c /= sms;
  // per iteration and warpgroup: 128 rows x 8 packed words x 16 genotypes
  const double geno = 128.0 * 8 * 16 * WG * iters;
  printf("tcgen05_mix WG=%d: %.1f genotypes decoded+contracted /clk/SM (%.0f MAC/clk/SM), %.0f clk per iteration\n",
         WG, geno / c, 8.0 * geno / c, c / iters);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  int* d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(int) * sms * 512));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * sms));
  long long* h_cyc = new long long[sms];
  if (run<1>(sms, d_out, d_cyc, h_cyc, 2000)) return 1;
  if (run<2>(sms, d_out, d_cyc, h_cyc, 2000)) return 1;
  if (run<4>(sms, d_out, d_cyc, h_cyc, 2000)) return 1;
  CK(cudaFree(d_out));
  CK(cudaMalloc(&d_out, sizeof(int) * sms * 1024));
  if (run_ring<1, 6>(sms, d_out, d_cyc, h_cyc, 4000)) return 1;
  if (run_ring<2, 6>(sms, d_out, d_cyc, h_cyc, 4000)) return 1;
  if (run_ring<4, 6>(sms, d_out, d_cyc, h_cyc, 4000)) return 1;
  if (run_ring<4, 3>(sms, d_out, d_cyc, h_cyc, 4000)) return 1;
  return 0;
}
