// tcgen05_probe.cu -- first steps of the tcgen05 contraction sketched in DESIGN.md section 8:
//
//   A. tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 8, K = 32, A operand in TMEM (written by
//      the threads with tcgen05.st.32x32b: lane = matrix row, 8 columns = 32 K-bytes), B operand
//      (signed digits, K-major, no swizzle) in shared memory, D (int32) in TMEM.  The result is
//      checked against a host computation, so a wrong descriptor shows up as mismatches, not as a
//      silent layout error.
//   B. the lane / column mapping of tcgen05.st.16x256b (the store that would carry an
//      ldmatrix.trans fragment into the A operand of the second half): every thread stores 4
//      tagged registers, the block is read back with the plain 32x32b shape and printed.
//
// Descriptor fields follow cute/arch/mma_sm100_desc.hpp (UMMA::InstrDescriptor, SmemDescriptor).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define CK(x)                                                      \
  do {                                                             \
    cudaError_t e_ = (x);                                          \
    if (e_ != cudaSuccess) {                                       \
      printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); \
      return 1;                                                    \
    }                                                              \
  } while (0)

constexpr int M = 128, N = 8, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// instruction descriptor: dense, no saturate, D = S32 (2), A = unsigned 8 bit (0), B = signed 8 bit (1),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc() {
  return (2u << 4) | (0u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_NONE: 8-row x 16-byte core matrices; for a K-major
// operand LBO = byte distance between core matrices along K, SBO = along M/N; version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version_
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128, 1)
k_probe(const uint8_t* __restrict__ A /* M x K */, const int8_t* __restrict__ B /* K x N */,
        int* __restrict__ D /* M x N */, uint32_t* __restrict__ map /* 32 lanes x 8 columns */) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(128) int8_t sB[2 * 8 * 16];  // two core matrices: k 0..15 and k 16..31
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B: core matrix c holds k = 16 c .. 16 c + 15, row n at n * 16 bytes
  for (int i = tid; i < K * N; i += blockDim.x) {
    const int k = i / N, n = i % N;
    sB[(k >> 4) * 128 + n * 16 + (k & 15)] = B[k * N + n];
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;  // this warp's TMEM lane quadrant
  const uint32_t tD = tbase, tA = tbase + 16, tX = tbase + 32;

  // ---- A: row = thread, 8 columns of 4 K-bytes each
  {
    uint32_t a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint8_t* p = A + tid * K + 4 * c;
      a[c] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                     tA + lane_base),
                 "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // sB was written by the generic proxy
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 128, 256);
    const uint32_t idesc = make_idesc(), zero = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tD),
        "r"(tA), "l"(bdesc), "r"(idesc), "r"(zero), "r"(zero)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bar))
                 : "memory");
  }
  {  // wait for the MMA
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 22); spin++)  // bounded: never hang the box
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    int d[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]),
                   "=r"(d[7])
                 : "r"(tD + lane_base)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int n = 0; n < 8; n++) D[tid * N + n] = d[n];
  }
  // ---- B. mapping of tcgen05.st.16x256b (warp 0 only): register r of thread t carries (t << 8) | r
  if (warp == 0) {
    const uint32_t tag = 0x10000u | ((uint32_t)lane << 8);  // bit 16 marks a written cell
    const uint32_t v0 = tag | 0, v1 = tag | 1, v2 = tag | 2, v3 = tag | 3;
    uint32_t z[8] = {};
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tX),
                 "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(tX), "r"(v0),
                 "r"(v1), "r"(v2), "r"(v3)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t m[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]), "=r"(m[4]), "=r"(m[5]), "=r"(m[6]),
                   "=r"(m[7])
                 : "r"(tX)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 8; c++) map[lane * 8 + c] = m[c];
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
  uint8_t hA[M * K];
  int8_t hB[K * N];
  for (int m = 0; m < M; m++)
    for (int k = 0; k < K; k++) hA[m * K + k] = (uint8_t)((m * 7 + k * 3) % 200);
  for (int k = 0; k < K; k++)
    for (int n = 0; n < N; n++) hB[k * N + n] = (int8_t)(((k * 5 + n * 11) % 127) - 63);
  uint8_t* dA;
  int8_t* dB;
  int* dD;
  uint32_t* dmap;
  CK(cudaMalloc(&dA, sizeof(hA)));
  CK(cudaMalloc(&dB, sizeof(hB)));
  CK(cudaMalloc(&dD, sizeof(int) * M * N));
  CK(cudaMalloc(&dmap, sizeof(uint32_t) * 32 * 8));
  CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, sizeof(int) * M * N));
  k_probe<<<1, 128>>>(dA, dB, dD, dmap);
  CK(cudaDeviceSynchronize());
  int hD[M * N];
  uint32_t hmap[32 * 8];
  CK(cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hmap, dmap, sizeof(hmap), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      int ref = 0;
      for (int k = 0; k < K; k++) ref += (int)hA[m * K + k] * (int)hB[k * N + n];
      if (ref != hD[m * N + n]) {
        if (bad < 8) printf("D[%d][%d] = %d, expected %d\n", m, n, hD[m * N + n], ref);
        bad++;
      }
    }
  printf("tcgen05.mma kind::i8 M=128 N=8 K=32, A in TMEM: %d mismatches of %d\n", bad, M * N);
  printf("tcgen05.st.16x256b.x1 mapping (TMEM lane, column) <- (thread, register):\n");
  for (int l = 0; l < 32; l++) {
    printf("lane %2d:", l);
    for (int c = 0; c < 8; c++) {
      const uint32_t v = hmap[l * 8 + c];
      if (v == 0) printf("    .   ");
      else printf(" t%02u.r%u ", (v >> 8) & 0xFF, v & 0xFF);
    }
    printf("\n");
  }
  return 0;
}
