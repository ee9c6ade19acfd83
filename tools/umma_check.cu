// umma_check.cu -- bring-up harness of the tcgen05 block-contraction kernels (fpb_umma.cuh):
//   1. register -> TMEM mapping of tcgen05.st.16x256b.x4 (the store the second half relies on);
//   2. correctness of k_umma_xt / k_umma_xv against brute-force FP64 kernels on ragged shapes;
//   3. throughput at the 500,000 x 100,000 headline shape (pseudo-random packed bytes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/umma_check tools/umma_check.cu
// Run:   tools/umma_check [quick]
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../flashpca_b200/csrc/fpb_umma.cuh"

#define CK(x)                                                                \
  do {                                                                       \
    cudaError_t e_ = (x);                                                    \
    if (e_ != cudaSuccess) {                                                 \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), #x, __LINE__); \
      exit(1);                                                               \
    }                                                                        \
  } while (0)

using namespace fpb;

static int make_map(const uint8_t* base, uint64_t pitch, uint64_t rows, TmaDesc* out, uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    encode = (EncodeFn)fn;
  }
  cuuint64_t dims[2] = {pitch, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                       const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)rc);
    return 1;
  }
  return 0;
}

// ---- 1. mapping probe ------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_map_probe(uint32_t* map /* 32 lanes x 32 columns */) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
                     smem_u32(&tmem_base_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    uint32_t z[16] = {};
    tmem_st_32x32b_x16(tb, z);
    tmem_st_32x32b_x16(tb + 16, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t d[16];
    for (int r = 0; r < 16; r++) d[r] = 0x10000u | ((uint32_t)lane << 8) | (uint32_t)r;
    tmem_st_16x256b_x4(tb, d);  // lanes 0..15
    for (int r = 0; r < 16; r++) d[r] |= 0x20000u;
    tmem_st_16x256b_x4(tb + (16u << 16), d);  // lanes 16..31
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; c += 8) {
      int m[8];
      tmem_ld_32x32b_x8(tb + c, m);
      for (int k = 0; k < 8; k++) map[lane * 32 + c + k] = (uint32_t)m[k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}

// ---- test data and references ------------------------------------------------------------------
__host__ __device__ inline uint32_t hash32(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
// packed dosage codes (0..3, 3 rare), zero beyond column n
__global__ void k_fill(uint8_t* g, uint64_t rows, uint64_t pitch, uint64_t n, uint64_t seed) {
  const uint64_t nw = rows * (pitch / 4);
  for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < nw;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w = idx % (pitch / 4);
    uint32_t v = 0;
    for (int f = 0; f < 16; f++) {
      const uint64_t col = w * 16 + f;
      if (col >= n) break;
      const uint32_t hh = hash32(idx * 16 + f + seed);
      uint32_t e = hh % 3u;
      if ((hh >> 8) % 97u == 0) e = 3;
      v |= e << (2 * f);
    }
    reinterpret_cast<uint32_t*>(g)[idx] = v;
  }
}
__global__ void k_fill_vec(double* v, uint64_t len, uint64_t seed, double scale) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < len) v[i] = scale * ((double)hash32(i + seed) / 4294967296.0 - 0.5) *
                      (1.0 + (double)(hash32(i * 7 + seed) & 1023));
}
// ref_xt[row] = sum_i e[row][i] x[i]   (one warp per row)
__global__ void k_ref_xt(const uint8_t* g, uint64_t pitch, uint32_t rows, uint64_t n, const double* x,
                         double* out) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s = 0.0;
  for (uint64_t i = lane; i < n; i += 32) {
    const uint32_t e = (g[(uint64_t)row * pitch + (i >> 2)] >> (2 * (i & 3))) & 3u;
    s += (double)e * x[i];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}
// ref_xv[i] = sum_j e[j][i] a[j]   (one thread per individual)
__global__ void k_ref_xv(const uint8_t* g, uint64_t pitch, uint32_t rows, uint64_t n, const double* a,
                         double* out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (uint32_t j = 0; j < rows; j++) {
    const uint32_t e = (g[(uint64_t)j * pitch + (i >> 2)] >> (2 * (i & 3))) & 3u;
    s += (double)e * a[j];
  }
  out[i] = s;
}

struct Problem {
  uint64_t n, pitch;
  uint32_t rows;
  uint8_t* g;
  TmaDesc tm;
};

static uint32_t* d_err;
static uint32_t read_err() {
  uint32_t e;
  CK(cudaMemcpy(&e, d_err, 4, cudaMemcpyDeviceToHost));
  return e;
}

template <int NV, int RG, int NCH>
static float run_xt(const Problem& P, const uint8_t* S, uint32_t splits, double* part, uint64_t vstride,
                    uint64_t sstride, int reps) {
  const uint32_t nstages = (uint32_t)((P.pitch + 127) / 128);
  const uint32_t sps = (nstages + splits - 1) / splits;
  auto kern = k_umma_xt<NV, RG, NCH>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
  dim3 grid((P.rows + RG * 128 - 1) / (RG * 128), splits);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kern<<<grid, (5 * RG + 1) * 32, kUSmemBytes>>>(P.tm, P.rows, S, nstages, sps, part, vstride, sstride, d_err);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++)
    kern<<<grid, (5 * RG + 1) * 32, kUSmemBytes>>>(P.tm, P.rows, S, nstages, sps, part, vstride, sstride, d_err);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return reps ? ms / reps : 0.f;
}

template <int NV, int NPAIR, int NISS>
static float run_xv(const Problem& P, const uint8_t* S, uint32_t splits, double* part, uint64_t vstride,
                    uint64_t sstride, int reps) {
  const uint32_t nboxes = (P.rows + 127) / 128;
  const uint32_t bps = (nboxes + splits - 1) / splits;
  auto kern = k_umma_xv<NV, NISS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
  dim3 grid((uint32_t)((P.pitch + 127) / 128), splits);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kern<<<grid, (8 + NISS + 1) * 32, kUSmemBytes>>>(P.tm, (uint32_t)P.n, S, nboxes, bps, part, vstride, sstride, d_err);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++)
    kern<<<grid, (8 + NISS + 1) * 32, kUSmemBytes>>>(P.tm, (uint32_t)P.n, S, nboxes, bps, part, vstride, sstride, d_err);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return reps ? ms / reps : 0.f;
}

// vectors, slices, references for one problem
struct Vecs {
  int nv;
  double *x, *a;          // nv x n, nv x rows
  double *ref_t, *ref_y;  // nv x rows, nv x n
  uint8_t *s_i, *s_j;     // slices
  VecScale* sc;           // [2 nv]
  double *pmax, *psum;
  uint32_t nkb_i, nkb_j;
};

static void prep_vecs(const Problem& P, Vecs& V, int nv, bool refs) {
  V.nv = nv;
  const uint32_t nstages = (uint32_t)((P.pitch + 127) / 128), nboxes = (P.rows + 127) / 128;
  V.nkb_i = nstages * 16;
  V.nkb_j = nboxes * 4;
  CK(cudaMalloc(&V.x, sizeof(double) * nv * P.n));
  CK(cudaMalloc(&V.a, sizeof(double) * nv * P.rows));
  CK(cudaMalloc(&V.s_i, (size_t)V.nkb_i * nv * 256));
  CK(cudaMalloc(&V.s_j, (size_t)V.nkb_j * nv * 256));
  CK(cudaMalloc(&V.sc, sizeof(VecScale) * 2 * nv));
  CK(cudaMalloc(&V.pmax, sizeof(double) * 256));
  CK(cudaMalloc(&V.psum, sizeof(double) * 256));
  V.ref_t = V.ref_y = nullptr;
  if (refs) {
    CK(cudaMalloc(&V.ref_t, sizeof(double) * nv * P.rows));
    CK(cudaMalloc(&V.ref_y, sizeof(double) * nv * P.n));
  }
  for (int v = 0; v < nv; v++) {
    double* xv = V.x + (uint64_t)v * P.n;
    double* av = V.a + (uint64_t)v * P.rows;
    k_fill_vec<<<(uint32_t)((P.n + 255) / 256), 256>>>(xv, P.n, 1000 + v, ldexp(1.0, 3 * v - 5));
    k_fill_vec<<<(P.rows + 255) / 256, 256>>>(av, P.rows, 5000 + v, ldexp(1.0, 7 - 2 * v));
    k_vec_partial<<<256, 256>>>(xv, P.n, V.pmax, V.psum);
    k_slice_umma_i<<<(V.nkb_i + 127) / 128, 128>>>(xv, P.n, V.nkb_i, nv, v, V.pmax, V.psum, 256, V.sc + v,
                                                  reinterpret_cast<uint4*>(V.s_i));
    k_vec_partial<<<256, 256>>>(av, P.rows, V.pmax, V.psum);
    k_slice_umma_j<<<(V.nkb_j + 127) / 128, 128>>>(av, P.rows, V.nkb_j, nv, v, V.pmax, V.psum, 256,
                                                  V.sc + nv + v, reinterpret_cast<uint4*>(V.s_j));
    if (refs) {
      k_ref_xt<<<(P.rows * 32 + 255) / 256, 256>>>(P.g, P.pitch, P.rows, P.n, xv, V.ref_t + (uint64_t)v * P.rows);
      k_ref_xv<<<(uint32_t)((P.n + 255) / 256), 256>>>(P.g, P.pitch, P.rows, P.n, av, V.ref_y + (uint64_t)v * P.n);
    }
  }
  CK(cudaDeviceSynchronize());
}
static void free_vecs(Vecs& V) {
  cudaFree(V.x); cudaFree(V.a); cudaFree(V.s_i); cudaFree(V.s_j); cudaFree(V.sc);
  cudaFree(V.pmax); cudaFree(V.psum); cudaFree(V.ref_t); cudaFree(V.ref_y);
}

// max relative error of delta * sum_splits part vs ref
static double check(const double* d_part, uint32_t splits, uint64_t sstride, uint64_t vstride, int nv,
                    uint64_t len, const VecScale* d_sc, const double* d_ref) {
  std::vector<double> part((size_t)nv * vstride), ref((size_t)nv * len);
  std::vector<VecScale> sc(nv);
  CK(cudaMemcpy(part.data(), d_part, sizeof(double) * nv * vstride, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ref.data(), d_ref, sizeof(double) * nv * len, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sc.data(), d_sc, sizeof(VecScale) * nv, cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (int v = 0; v < nv; v++) {
    double mx = 0.0, err = 0.0;
    uint64_t at = 0;
    for (uint64_t i = 0; i < len; i++) {
      double s = 0.0;
      for (uint32_t sp = 0; sp < splits; sp++) s += part[(size_t)v * vstride + (size_t)sp * sstride + i];
      s *= sc[v].delta;
      const double r = ref[(size_t)v * len + i];
      mx = fmax(mx, fabs(r));
      if (!(fabs(s - r) <= 1e300)) {  // NaN / Inf: unwritten or broken output
        err = INFINITY;
        at = i;
        break;
      }
      if (fabs(s - r) > err) {
        err = fabs(s - r);
        at = i;
      }
    }
    const double rel = mx > 0 ? err / mx : err;
    if (rel > 1e-12) printf("    vector %d: rel err %.3e at %llu (max |ref| %.3e)\n", v, rel, (unsigned long long)at, mx);
    worst = fmax(worst, rel);
  }
  return worst;
}

template <int NV, int RG, int NCH, int NPAIR, int NISS>
static int correctness(uint32_t rows, uint64_t n, uint32_t splits_t, uint32_t splits_v) {
  Problem P;
  P.n = n;
  P.rows = rows;
  P.pitch = (((n + 3) / 4) + 63) / 64 * 64;
  CK(cudaMalloc(&P.g, (size_t)rows * P.pitch));
  k_fill<<<1024, 256>>>(P.g, rows, P.pitch, n, 77);
  if (make_map(P.g, P.pitch, rows, &P.tm, 128)) return 1;
  Vecs V;
  prep_vecs(P, V, NV, true);
  const uint64_t len = n > rows ? n : rows;
  const uint32_t smax = splits_t > splits_v ? splits_t : splits_v;
  const uint64_t sstride = (len + 63) / 64 * 64, vstride = sstride * smax;
  double* part;
  CK(cudaMalloc(&part, sizeof(double) * NV * vstride));
  CK(cudaMemset(part, 0xFF, sizeof(double) * NV * vstride));
  run_xt<NV, RG, NCH>(P, V.s_i, splits_t, part, vstride, sstride, 0);
  const double e1 = check(part, splits_t, sstride, vstride, NV, rows, V.sc, V.ref_t);
  const uint32_t err1 = read_err();
  CK(cudaMemset(d_err, 0, 4));
  CK(cudaMemset(part, 0xFF, sizeof(double) * NV * vstride));
  run_xv<NV, NPAIR, NISS>(P, V.s_j, splits_v, part, vstride, sstride, 0);
  const double e2 = check(part, splits_v, sstride, vstride, NV, n, V.sc + NV, V.ref_y);
  const uint32_t err2 = read_err();
  printf("correctness NV=%d RG=%d NCH=%d NPAIR=%d NISS=%d rows=%u n=%llu splits=%u/%u: xt rel err %.2e (gerr %x), "
         "xv rel err %.2e (gerr %x) %s\n",
         NV, RG, NCH, NPAIR, NISS, rows, (unsigned long long)n, splits_t, splits_v, e1, err1, e2, err2,
         (e1 < 1e-12 && e2 < 1e-12 && !err1 && !err2) ? "OK" : "FAIL");
  CK(cudaMemset(d_err, 0, 4));
  free_vecs(V);
  cudaFree(part);
  cudaFree(P.g);
  return 0;
}

template <int NV, int RG, int NCH, int NPAIR, int NISS>
static void timing(const Problem& P, uint32_t splits_t, uint32_t splits_v, int reps) {
  Vecs V;
  prep_vecs(P, V, NV, false);
  const uint64_t len = P.n > P.rows ? P.n : P.rows;
  const uint32_t smax = splits_t > splits_v ? splits_t : splits_v;
  const uint64_t sstride = (len + 63) / 64 * 64, vstride = sstride * smax;
  double* part;
  CK(cudaMalloc(&part, sizeof(double) * NV * vstride));
  const double bytes = (double)P.rows * (double)((P.n + 3) / 4);
  const float t1 = run_xt<NV, RG, NCH>(P, V.s_i, splits_t, part, vstride, sstride, reps);
  const uint32_t err1 = read_err();
  CK(cudaMemset(d_err, 0, 4));
  const float t2 = run_xv<NV, NPAIR, NISS>(P, V.s_j, splits_v, part, vstride, sstride, reps);
  const uint32_t err2 = read_err();
  printf("timing NV=%d RG=%d NCH=%d NPAIR=%d NISS=%d splits=%u/%u: xt %.3f ms (%.0f GB/s, gerr %x)  xv %.3f ms (%.0f GB/s, gerr %x)"
         "  per-vector op %.3f ms\n",
         NV, RG, NCH, NPAIR, NISS, splits_t, splits_v, t1, bytes / t1 * 1e-6, err1, t2, bytes / t2 * 1e-6, err2,
         (t1 + t2) / NV);
  fflush(stdout);
  CK(cudaMemset(d_err, 0, 4));
  free_vecs(V);
  cudaFree(part);
}

int main(int argc, char** argv) {
  const bool quick = argc > 1;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  CK(cudaMalloc(&d_err, 4));
  CK(cudaMemset(d_err, 0, 4));
  {  // 1. mapping
    uint32_t* d_map;
    CK(cudaMalloc(&d_map, sizeof(uint32_t) * 32 * 32));
    CK(cudaMemset(d_map, 0, sizeof(uint32_t) * 32 * 32));
    k_map_probe<<<1, 128>>>(d_map);
    CK(cudaDeviceSynchronize());
    uint32_t hm[32 * 32];
    CK(cudaMemcpy(hm, d_map, sizeof(hm), cudaMemcpyDeviceToHost));
    printf("tcgen05.st.16x256b.x4 mapping, (lane, column) <- thread.register (second store tagged '):\n");
    int expected = 1;
    for (int l = 0; l < 32; l++) {
      printf("lane %2d:", l);
      for (int c = 0; c < 32; c++) {
        const uint32_t v = hm[l * 32 + c];
        const int t = (v >> 8) & 0xFF, r = v & 0xFF, second = (v >> 17) & 1;
        printf(" %02d.%02d%c", t, r, second ? '\'' : ' ');
        // expectation: lane = g (+8 for registers 2,3 of a group of 4), column = 8 k + 2 q + (r & 1)
        const int g = (l & 15) & 7, up = ((l & 15) >> 3), k = c >> 3, qq = (c & 7) >> 1;
        const int et = g * 4 + qq, er = 4 * k + 2 * up + (c & 1);
        if (t != et || r != er || second != (l >> 4)) expected = 0;
      }
      printf("\n");
    }
    printf("mapping %s the expected pattern\n", expected ? "MATCHES" : "DOES NOT MATCH");
    cudaFree(d_map);
  }
  // 2. correctness on ragged shapes
  correctness<1, 2, 1, 1, 1>(300, 1000, 1, 1);
  correctness<1, 2, 2, 1, 2>(1000, 5003, 3, 2);
  correctness<1, 4, 4, 2, 4>(1000, 5003, 3, 2);
  correctness<2, 2, 2, 2, 1>(1000, 5003, 2, 3);
  correctness<4, 4, 1, 2, 2>(1531, 9001, 4, 2);
  correctness<8, 4, 1, 2, 1>(1531, 9001, 2, 4);
  correctness<8, 2, 1, 1, 4>(777, 20011, 5, 1);
  correctness<3, 2, 2, 1, 1>(129, 517, 1, 1);
  if (quick) return 0;
  // 3. throughput at 500,000 x 100,000
  Problem P;
  P.n = 500000;
  P.rows = 100000;
  P.pitch = (((P.n + 3) / 4) + 63) / 64 * 64;
  CK(cudaMalloc(&P.g, (size_t)P.rows * P.pitch));
  k_fill<<<148 * 8, 256>>>(P.g, P.rows, P.pitch, P.n, 99);
  CK(cudaDeviceSynchronize());
  if (make_map(P.g, P.pitch, P.rows, &P.tm, 128)) return 1;
  timing<1, 4, 1, 1, 4>(P, 8, 3, 5);
  timing<4, 4, 1, 1, 4>(P, 8, 3, 5);
  timing<4, 4, 1, 1, 2>(P, 8, 3, 5);
  timing<8, 4, 1, 1, 4>(P, 8, 3, 5);
  timing<8, 4, 1, 1, 2>(P, 8, 3, 5);
  timing<8, 4, 1, 1, 1>(P, 8, 3, 5);
  timing<8, 4, 1, 1, 4>(P, 8, 6, 5);
  return 0;
}
