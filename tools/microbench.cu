// microbench.cu -- B200 pipe-rate probes that size the perform_op kernel design
// (FP64 add/FMA issue rate, shared-memory table gathers, legacy IMMA / DMMA
// rates, HBM streaming read).  Prints per-SM-per-clock rates measured with
// clock64() inside the kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %s\n", cudaGetErrorString(e), #x); return 1; } } while (0)

__global__ void k_dadd(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  double x = 1e-9 * blockIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    a0 += x; a1 += x; a2 += x; a3 += x; a4 += x; a5 += x; a6 += x; a7 += x;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dfma(double* out, long long* cyc, int iters) {
  double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
  double x = 1e-9 * blockIdx.x, y = 1.0000001;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, y, x); a1 = fma(a1, y, x); a2 = fma(a2, y, x); a3 = fma(a3, y, x);
    a4 = fma(a4, y, x); a5 = fma(a5, y, x); a6 = fma(a6, y, x); a7 = fma(a7, y, x);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// predicated-add pattern of the v1 kernels: 32 genotypes per word, 2 masked sums
__global__ void k_maskadd(double* out, long long* cyc, const uint32_t* words, int iters) {
  double xr[16];
  for (int k = 0; k < 16; k++) xr[k] = threadIdx.x * 0.01 + k;
  double hs = 0, ls = 0;
  uint32_t w = words[threadIdx.x];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (w & (2u << (2 * k))) hs += xr[k];
      if (w & (1u << (2 * k))) ls += xr[k];
    }
    w = w * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = hs + ls;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// conflict-free byte-table gather: lane l owns bank pair (l & 15): entry v of 16
// interleaved tables at ((v * 16) + (l & 15)) * 8 bytes; two half-warps use two table sets.
__global__ void k_lds_gather(double* out, long long* cyc, const uint32_t* words, int iters) {
  extern __shared__ double tab[];  // 2 sets * 256 entries * 16 lanes
  for (int i = threadIdx.x; i < 2 * 256 * 16; i += blockDim.x) tab[i] = i * 1e-3;
  __syncthreads();
  int lane = threadIdx.x & 31;
  const double* my = tab + (lane >> 4) * 4096 + (lane & 15);
  uint32_t w = words[threadIdx.x];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    a0 += my[(w & 0xFF) * 16];
    a1 += my[((w >> 8) & 0xFF) * 16];
    a2 += my[((w >> 16) & 0xFF) * 16];
    a3 += my[(w >> 24) * 16];
    w = w * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// legacy tensor path: mma.sync m16n8k32 u8 x s8 -> s32
__global__ void k_imma(int* out, long long* cyc, int iters) {
  uint32_t a0 = threadIdx.x, a1 = 0x01020102, a2 = 0x02010201, a3 = 0x01010101;
  uint32_t b0 = 0x7f807f80 ^ threadIdx.x, b1 = 0x11223344;
  int c[4][4] = {};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[q][0]), "+r"(c[q][1]), "+r"(c[q][2]), "+r"(c[q][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  int s = 0;
  for (int q = 0; q < 4; q++) for (int k = 0; k < 4; k++) s += c[q][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FP64 tensor path: mma.sync m8n8k4 f64
__global__ void k_dmma(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c[4][2] = {};
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int q = 0; q < 4; q++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 4; q++) s += c[q][0] + c[q][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_stream(const uint4* __restrict__ in, size_t nvec, unsigned long long* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i + 3 * stride < nvec; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
  }
  for (; i < nvec; i += stride) { uint4 a = in[i]; acc += a.x ^ a.y ^ a.z ^ a.w; }
  if (acc == 0x12345678u) atomicAdd(out, 1ull);
}

static double avg_cycles(long long* d_cyc, int nb) {
  static long long h[4096];
  cudaMemcpy(h, d_cyc, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < nb; i++) s += (double)h[i];
  return s / nb;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  double* d_out; long long* d_cyc; uint32_t* d_words; int* d_iout;
  CK(cudaMalloc(&d_out, sizeof(double) * 4096 * 1024));
  CK(cudaMalloc(&d_iout, sizeof(int) * 4096 * 1024));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * 4096));
  CK(cudaMalloc(&d_words, sizeof(uint32_t) * 1024));
  uint32_t hw[1024];
  for (int i = 0; i < 1024; i++) hw[i] = 2654435761u * (i + 1);
  CK(cudaMemcpy(d_words, hw, sizeof(hw), cudaMemcpyHostToDevice));
  const int iters = 4096;
  for (int tpb : {256, 512, 1024}) {
    int nb = sms;  // one CTA per SM
    k_dadd<<<nb, tpb>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    printf("DADD  tpb=%4d: %.1f lane-ops/clk/SM\n", tpb, (double)tpb * iters * 8 / avg_cycles(d_cyc, nb));
    k_dfma<<<nb, tpb>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    printf("DFMA  tpb=%4d: %.1f lane-ops/clk/SM\n", tpb, (double)tpb * iters * 8 / avg_cycles(d_cyc, nb));
    k_maskadd<<<nb, tpb>>>(d_out, d_cyc, d_words, iters / 4); CK(cudaDeviceSynchronize());
    printf("MASKADD tpb=%4d: %.1f masked-adds/clk/SM (2 per genotype)\n", tpb, (double)tpb * (iters / 4) * 32 / avg_cycles(d_cyc, nb));
    CK(cudaFuncSetAttribute(k_lds_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k_lds_gather<<<nb, tpb, 65536>>>(d_out, d_cyc, d_words, iters); CK(cudaDeviceSynchronize());
    printf("LDS64 gather tpb=%4d: %.1f lookups/clk/SM (x4 genotypes each)\n", tpb, (double)tpb * iters * 4 / avg_cycles(d_cyc, nb));
    k_imma<<<nb, tpb>>>(d_iout, d_cyc, iters); CK(cudaDeviceSynchronize());
    printf("IMMA m16n8k32 tpb=%4d: %.1f MACs/clk/SM\n", tpb, (double)(tpb / 32) * iters * 4 * 4096 / avg_cycles(d_cyc, nb));
    k_dmma<<<nb, tpb>>>(d_out, d_cyc, iters); CK(cudaDeviceSynchronize());
    printf("DMMA m8n8k4 tpb=%4d: %.1f FMAs/clk/SM\n", tpb, (double)(tpb / 32) * iters * 4 * 256 / avg_cycles(d_cyc, nb));
  }
  // HBM streaming read
  size_t bytes = 8ull << 30;
  uint4* d_in; unsigned long long* d_cnt;
  CK(cudaMalloc(&d_in, bytes)); CK(cudaMalloc(&d_cnt, 8));
  CK(cudaMemset(d_in, 1, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mult : {4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      k_stream<<<sms * mult, 512>>>(d_in, bytes / 16, d_cnt);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("HBM read LDG.128 grid=%dxSMs: %.0f GB/s\n", mult, bytes / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
