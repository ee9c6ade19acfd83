// fpb_dense.cuh -- the in-memory matrix path: upstream's SVDWide operator on a
// pre-standardised N x P double matrix (svdwide.cpp:4-12) and the in-place
// standardiser that feeds it (util.cpp:24-192), used by
// RandomPCA::pca_fast(MatrixXd&, ...) (randompca.cpp:121-166), i.e. by
// `flashpca --batch` and flashpcaR's numeric-matrix entry point.  The matrix
// lives in HBM as column-major doubles (like Eigen::MatrixXd), so this path is
// for matrices that fit there at 8 B per genotype; the packed 2-bit path is the
// product path for real bed files.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "fpb_kernels.cuh"

namespace fpb {

// fixed-order block reduction of three doubles (256 threads)
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*sh)[3]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp][0] = a;
    sh[warp][1] = b;
    sh[warp][2] = c;
  }
  __syncthreads();
  a = b = c = 0.0;
  for (int w = 0; w < 8; w++) {
    a += sh[w][0];
    b += sh[w][1];
    c += sh[w][2];
  }
  __syncthreads();
}

// util.cpp:24-192, one block per column.  method: 0 none, 1 sd, 2 binom, 3 binom2,
// 4 center (util.h:34-38).  NaN = missing.  meansd: p x 2 (mean, sd), colsq[j] =
// sum of squares of the standardised column (trace, randompca.cpp:154).
__global__ void __launch_bounds__(256)
k_dense_standardise(double* __restrict__ X, uint64_t n, uint32_t p, int method,
                    double* __restrict__ meansd, double* __restrict__ colsq) {
  __shared__ double sh[8][3];
  const uint32_t j = blockIdx.x;
  double* col = X + (uint64_t)j * n;
  const double K = (method == 1) ? 1.0 : 0.0;  // shifted-data variance, util.cpp:84
  double sum = 0.0, sq = 0.0, cnt = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += 256) {
    double v = col[i];
    if (v == v) {
      sum += v - K;
      sq += (v - K) * (v - K);
      cnt += 1.0;
    }
  }
  block_sum3(sum, sq, cnt, sh);
  double mean, sd = 1.0;
  if (method == 1) {
    double var = (sq - (sum * sum) / cnt) / (cnt - 1.0);
    mean = (sum + K * cnt) / cnt;
    sd = sqrt(var);
  } else {
    mean = sum / cnt;
    if (method == 2 || method == 3) {
      double r = mean / 2.0;
      sd = sqrt((method == 2 ? 1.0 : 2.0) * r * (1.0 - r));
    }
  }
  double s2 = 0.0, d0 = 0.0, d1 = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += 256) {
    double v = col[i], o;
    bool na = !(v == v);
    if (method == 0) o = na ? mean : v;
    else if (method == 4) o = na ? 0.0 : v - mean;
    else o = na ? 0.0 : (sd > kVarTol ? (v - mean) / sd : mean);  // util.cpp:139-146 quirk kept
    col[i] = o;
    s2 += o * o;
  }
  block_sum3(s2, d0, d1, sh);
  if (threadIdx.x == 0) {
    meansd[j] = mean;
    meansd[p + j] = sd;
    colsq[j] = s2;
  }
}

// t_j = sum_i X_ij x_i, one block per column, fixed order.  Four independent 8-byte loads of the
// column per thread and iteration (8 KB of the column in flight per block) keep the 8 N P byte
// stream of svdwide.cpp:10 (`mat.transpose() * x`) at HBM speed.
__global__ void __launch_bounds__(256)
k_dense_gemv_t(const double* __restrict__ X, uint64_t n, const double* __restrict__ x,
               double* __restrict__ t) {
  __shared__ double sh[8][3];
  const double* col = X + (uint64_t)blockIdx.x * n;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  uint64_t i = threadIdx.x;
  for (; i + 768 < n; i += 1024) {
    const double c0 = col[i], c1 = col[i + 256], c2 = col[i + 512], c3 = col[i + 768];
    s0 += c0 * x[i];
    s1 += c1 * x[i + 256];
    s2 += c2 * x[i + 512];
    s3 += c3 * x[i + 768];
  }
  for (; i < n; i += 256) s0 += col[i] * x[i];
  double s = (s0 + s1) + (s2 + s3), d0 = 0.0, d1 = 0.0;
  block_sum3(s, d0, d1, sh);
  if (threadIdx.x == 0) t[blockIdx.x] = s;
}

// partial[split * n + i] = sum_{j in split} X_ij t_j   (thread per row, coalesced columns, four
// columns in flight per thread)
__global__ void __launch_bounds__(256)
k_dense_gemv_n(const double* __restrict__ X, uint64_t n, uint32_t p, uint32_t cols_per_split,
               const double* __restrict__ t, double* __restrict__ partial) {
  uint64_t i = blockIdx.x * (uint64_t)256 + threadIdx.x;
  if (i >= n) return;
  uint32_t j0 = blockIdx.y * cols_per_split, j1 = min(p, j0 + cols_per_split);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  uint32_t j = j0;
  for (; j + 3 < j1; j += 4) {
    const double c0 = X[i + (uint64_t)j * n], c1 = X[i + (uint64_t)(j + 1) * n],
                 c2 = X[i + (uint64_t)(j + 2) * n], c3 = X[i + (uint64_t)(j + 3) * n];
    s0 += c0 * t[j];
    s1 += c1 * t[j + 1];
    s2 += c2 * t[j + 2];
    s3 += c3 * t[j + 3];
  }
  for (; j < j1; j++) s0 += X[i + (uint64_t)j * n] * t[j];
  partial[(uint64_t)blockIdx.y * n + i] = (s0 + s1) + (s2 + s3);
}

__global__ void k_sum_splits(const double* __restrict__ partial, uint32_t nsplits, uint64_t n,
                             double* __restrict__ y) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (uint32_t k = 0; k < nsplits; k++) s += partial[(uint64_t)k * n + i];
  y[i] = s;
}

}  // namespace fpb
