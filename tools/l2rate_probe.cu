// l2rate_probe.cu -- how fast do the two HBM-bound contraction kernels (fpb_imma.cuh) run when the
// packed matrix comes from L2 instead of HBM?  The tensor map aliases the 500,000 x 100,000 matrix
// onto a 26 MB buffer (row stride 256 B instead of 125,056 B): same instruction stream, same
// shared-memory traffic, TMA loads that hit L2.  The answer is the compute-bound rate of the
// loops, i.e. the headroom a design that splits the SMs between the two halves could use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2rate_probe tools/l2rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../flashpca_b200/csrc/fpb_imma.cuh"

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), #x, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)
using namespace fpb;

static int make_map(const uint8_t* base, uint64_t width, uint64_t rows, uint64_t stride, TmaDesc* out,
                    uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  cuuint64_t dims[2] = {width, rows};
  cuuint64_t strides[1] = {stride};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = ((EncodeFn)fn)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                               const_cast<uint8_t*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed %d (stride %llu)\n", (int)rc, (unsigned long long)stride);
  return rc != CUDA_SUCCESS;
}

int main() {
  const uint64_t n = 500000, rows = 100000, pitch = 125056;
  const uint32_t nstages = (uint32_t)((pitch + 127) / 128);
  for (int pass = 0; pass < 2; pass++) {
    const uint64_t stride = pass == 0 ? pitch : 256;     // pass 0: the real layout (HBM), pass 1: aliased (L2)
    const size_t bytes = (size_t)rows * stride + pitch;
    uint8_t* g;
    CK(cudaMalloc(&g, bytes));
    CK(cudaMemset(g, 0x5A, bytes));
    TmaDesc tm;
    if (make_map(g, pitch, rows, stride, &tm, kTmaRows)) return 1;
    uint4* S;
    CK(cudaMalloc(&S, (size_t)nstages * kTmaSliceBytes + (size_t)(rows / 256 + 2) * kTmaTSliceBytes));
    CK(cudaMemset(S, 0x11, (size_t)nstages * kTmaSliceBytes + (size_t)(rows / 256 + 2) * kTmaTSliceBytes));
    double* out;
    const uint32_t splits1 = 60, sps = (nstages + splits1 - 1) / splits1;
    const uint32_t ntiles = (uint32_t)((rows + 255) / 256), splits2 = 8, tps = (ntiles + splits2 - 1) / splits2;
    CK(cudaMalloc(&out, sizeof(double) * 64 * 500000));
    CK(cudaFuncSetAttribute(k_imma_gemv_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
    CK(cudaFuncSetAttribute(k_imma_gemv_tma_t, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; which++) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        if (which == 0) {
          dim3 grid((uint32_t)((rows + kTmaRows - 1) / kTmaRows), (nstages + sps - 1) / sps);
          k_imma_gemv_tma<<<grid, (kTmaConsumerWarps + 1) * 32, kTmaSmemBytes>>>(tm, (uint32_t)rows, S, nstages, sps,
                                                                                out, 500000);
        } else {
          dim3 grid(nstages, (ntiles + tps - 1) / tps);
          k_imma_gemv_tma_t<<<grid, (kTmaConsumerWarps + 1) * 32, kTmaSmemBytes>>>(
              tm, (uint32_t)n, reinterpret_cast<const uint32_t*>(S), ntiles, tps, out, 500000);
        }
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
      printf("%s matrix, %s: %.3f ms per pass over 12.5 GB (%.0f GB/s equivalent, %.1f genotypes/clk/SM at 1.9 GHz)\n",
             pass == 0 ? "HBM-resident" : "L2-aliased  ", which == 0 ? "k_imma_gemv_tma  " : "k_imma_gemv_tma_t", best,
             12.5e9 / best * 1e-6, 5e10 / (best * 1e-3) / 148 / 1.9e9);
    }
    cudaFree(g); cudaFree(S); cudaFree(out);
  }
  return 0;
}
