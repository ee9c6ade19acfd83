// fpb_block.cuh -- block Krylov eigensolver on the tcgen05 block operator (extension).
//
// Upstream drives Spectra's single-vector implicitly restarted Lanczos (randompca.cpp:174-190),
// which fpb_irlm.cuh follows step for step.  On B200 the operator is 4.5 times cheaper per column
// when 8 columns share one pass over the packed matrix (fpb_umma.cuh: 0.75 ms per column instead of
// 3.9 ms), so the solver the hardware wants is a block method: block Lanczos with full
// re-orthogonalisation and Rayleigh-Ritz on the accumulated block Krylov space
//   K_j = span{V_0, A V_0, ..., A^j V_0},  V_0 = orth(random N x b),
// no restart (k = 20 converges in 12-26 passes of b = 8 columns on the matrices studied in
// profiles/r01_block_solver_study.txt, i.e. a basis of 96-208 vectors).  Convergence test and
// tolerance are Spectra's: ||A u - theta u|| < tol max(eps^(2/3), |theta|) for every wanted pair.
// The iteration trajectory differs from upstream's by construction; parity is judged on the
// converged eigenpairs, and fpb_pca (Spectra's schedule) stays the default.
//
// All tall-skinny algebra is on the device with fixed-order reductions (SNP-sharded ranks run this
// driver redundantly on identical all-reduced blocks and must take identical decisions); the
// m x m projected problem (m <= b x passes) is solved on the host (Householder + implicit QL).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cfloat>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include "fpb_irlm.cuh"
#include "fpb_kernels.cuh"

namespace fpb {

constexpr int kBlkMaxB = 8;        // block width (columns per operator pass)
constexpr int kBlkRows = 512;      // rows per thread block of the V'W kernel

// partial[blk][c * b + t] = sum over the block's rows of V[r, c] W[r, t]   (c < m, t < b <= 8)
// The W rows of the block sit in shared memory; warp w takes columns w, w + 8, ... of V.
__global__ void __launch_bounds__(256)
k_vt_w_partial(const double* __restrict__ V, uint64_t ldv, uint32_t m, const double* __restrict__ W,
               uint64_t ldw, uint32_t b, uint64_t n, double* __restrict__ partial) {
  __shared__ double ws[kBlkRows * kBlkMaxB];
  const uint64_t r0 = (uint64_t)blockIdx.x * kBlkRows;
  for (uint32_t e = threadIdx.x; e < kBlkRows * b; e += 256) {
    const uint32_t t = e / kBlkRows, r = e - t * kBlkRows;
    ws[r * kBlkMaxB + t] = (r0 + r < n) ? W[r0 + r + (uint64_t)t * ldw] : 0.0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t c = warp; c < m; c += 8) {
    const double* col = V + (uint64_t)c * ldv + r0;
    double s[kBlkMaxB] = {};
#pragma unroll 4
    for (int i = 0; i < kBlkRows / 32; i++) {
      const uint32_t r = (uint32_t)lane + 32u * i;
      const double v = (r0 + r < n) ? col[r] : 0.0;
#pragma unroll
      for (int t = 0; t < kBlkMaxB; t++) s[t] += v * ws[r * kBlkMaxB + t];
    }
#pragma unroll
    for (int t = 0; t < kBlkMaxB; t++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
    }
    if (lane == 0)
      for (uint32_t t = 0; t < b; t++) partial[((uint64_t)blockIdx.x * m + c) * b + t] = s[t];
  }
}

// W[r, t] = uniform(-0.5, 0.5) from a counter-based hash of (r, t, seed): the start block
__global__ void k_random_block(double* __restrict__ W, uint64_t n, uint32_t b, uint64_t seed) {
  const uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (idx >= n * b) return;
  uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  W[idx] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// ------------------------------- host algebra -------------------------------

// Eigen-decomposition of a real symmetric n x n matrix (column-major, full storage): Householder
// reduction to tridiagonal form with the transformations accumulated, then implicit-shift QL.
// w: eigenvalues, ascending order NOT guaranteed; z: eigenvectors in columns (column-major).
inline bool sym_eigen(int n, const std::vector<double>& a_in, std::vector<double>& w,
                      std::vector<double>& z) {
  std::vector<double> a(a_in);  // a(i, j) = a[j * n + i]
  auto A = [&](int i, int j) -> double& { return a[(size_t)j * n + i]; };
  std::vector<double> d(n, 0.0), e(n, 0.0);
  for (int i = n - 1; i > 0; i--) {
    const int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; k++) scale += fabs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        for (int k = 0; k <= l; k++) {
          A(i, k) /= scale;
          h += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; j++) {
          A(j, i) = A(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; k++) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; k++) g += A(k, j) * A(i, k);
          e[j] = g / h;
          f += e[j] * A(i, j);
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; j++) {
          f = A(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; k++) A(j, k) -= (f * e[k] + g * A(i, k));
        }
      }
    } else {
      e[i] = A(i, l);
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; i++) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; j++) {
        double g = 0.0;
        for (int k = 0; k <= l; k++) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; k++) A(k, j) -= g * A(k, i);
      }
    }
    d[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; j++) A(j, i) = A(i, j) = 0.0;
  }
  // implicit QL on (d, e) with z = the accumulated Householder matrix
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  z = a;  // column-major: z(k, i) = z[i * n + k]
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= DBL_EPSILON * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; i--) {
          double f = s * e[i];
          const double bb = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * bb;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - bb;
          double* zi = &z[(size_t)i * n];
          double* zi1 = &z[(size_t)(i + 1) * n];
          for (int k = 0; k < n; k++) {
            f = zi1[k];
            zi1[k] = s * zi[k] + c * f;
            zi[k] = c * zi[k] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  w = d;
  return true;
}

struct BlockResult {
  std::vector<double> evals;  // nev, descending
  uint32_t nconv = 0, npasses = 0, nvecprod = 0;
};

// `op(d_in, k, d_out)` enqueues A * in (N x k, column-major, leading dimension N) -> out on `stream`.
class BlockKrylov {
 public:
  BlockKrylov(uint64_t n, uint32_t nev, uint32_t b, uint32_t max_passes, cudaStream_t stream,
              std::function<int(const double*, uint32_t, double*)> op)
      : n_(n), nev_(nev), b_(b), maxp_(max_passes), stream_(stream), op_(std::move(op)) {}
  ~BlockKrylov() { release(); }
  bool matches(uint64_t n, uint32_t nev, uint32_t b, uint32_t mp) const {
    return n == n_ && nev == nev_ && b == b_ && mp == maxp_;
  }
  void set_op(std::function<int(const double*, uint32_t, double*)> op) { op_ = std::move(op); }
  int run(double tol, BlockResult& res);
  const double* eigenvectors() const { return dU_; }  // N x nev, device, sorted like evals
  std::string error;
  double t_op = 0, t_orth = 0, t_ritz = 0;

 private:
  int alloc();
  void release();
  // host C (m x b, column-major) = V[:, :m]' W
  void vt_w(const double* V, uint32_t m, const double* W, std::vector<double>& C);
  // out (N x nc) (+)= B[:, :m] * Q (m x nc, host, column-major)
  void times_small(const double* B, uint32_t m, const std::vector<double>& Q, uint32_t nc, double* out,
                   bool accumulate);
  // W <- orth(W) by two rounds of Cholesky QR; false when W is (numerically) rank deficient
  bool cholqr2(double* W);

  uint64_t n_;
  uint32_t nev_, b_, maxp_;
  cudaStream_t stream_;
  std::function<int(const double*, uint32_t, double*)> op_;
  double *dV_ = nullptr, *dAV_ = nullptr, *dW_ = nullptr, *dT_ = nullptr, *dU_ = nullptr, *dAU_ = nullptr,
         *dPart_ = nullptr, *dSmall_ = nullptr, *dQ_ = nullptr;
  uint32_t nblk_ = 0, mmax_ = 0;
};

inline int BlockKrylov::alloc() {
  if (dV_) return 0;
  mmax_ = b_ * maxp_;
  nblk_ = (uint32_t)((n_ + kBlkRows - 1) / kBlkRows);
  const size_t nm = (size_t)n_ * mmax_;
  if (cudaMalloc(&dV_, sizeof(double) * nm) != cudaSuccess ||
      cudaMalloc(&dAV_, sizeof(double) * nm) != cudaSuccess ||
      cudaMalloc(&dW_, sizeof(double) * n_ * b_) != cudaSuccess ||
      cudaMalloc(&dT_, sizeof(double) * n_ * b_) != cudaSuccess ||
      cudaMalloc(&dU_, sizeof(double) * n_ * nev_) != cudaSuccess ||
      cudaMalloc(&dAU_, sizeof(double) * n_ * nev_) != cudaSuccess ||
      cudaMalloc(&dPart_, sizeof(double) * (size_t)nblk_ * mmax_ * b_) != cudaSuccess ||
      cudaMalloc(&dSmall_, sizeof(double) * (size_t)mmax_ * std::max<uint32_t>(b_, nev_)) != cudaSuccess ||
      cudaMalloc(&dQ_, sizeof(double) * (size_t)mmax_ * std::max<uint32_t>(b_, nev_)) != cudaSuccess) {
    cudaGetLastError();
    error = "out of device memory for the block Krylov basis";
    release();
    return 1;
  }
  return 0;
}

inline void BlockKrylov::release() {
  cudaFree(dV_); cudaFree(dAV_); cudaFree(dW_); cudaFree(dT_); cudaFree(dU_); cudaFree(dAU_);
  cudaFree(dPart_); cudaFree(dSmall_); cudaFree(dQ_);
  dV_ = dAV_ = dW_ = dT_ = dU_ = dAU_ = dPart_ = dSmall_ = dQ_ = nullptr;
}

inline void BlockKrylov::vt_w(const double* V, uint32_t m, const double* W, std::vector<double>& C) {
  k_vt_w_partial<<<nblk_, 256, 0, stream_>>>(V, n_, m, W, n_, b_, n_, dPart_);
  const uint64_t ncols = (uint64_t)m * b_;
  k_sum_rows<<<(uint32_t)((ncols + 255) / 256), 256, 0, stream_>>>(dPart_, nblk_, ncols, dSmall_);
  std::vector<double> tmp(ncols);
  cudaMemcpyAsync(tmp.data(), dSmall_, sizeof(double) * ncols, cudaMemcpyDeviceToHost, stream_);
  cudaStreamSynchronize(stream_);
  C.assign(ncols, 0.0);  // device layout [c][t] -> column-major m x b
  for (uint32_t c = 0; c < m; c++)
    for (uint32_t t = 0; t < b_; t++) C[(size_t)t * m + c] = tmp[(size_t)c * b_ + t];
}

inline void BlockKrylov::times_small(const double* B, uint32_t m, const std::vector<double>& Q,
                                     uint32_t nc, double* out, bool accumulate) {
  cudaMemcpyAsync(dQ_, Q.data(), sizeof(double) * (size_t)m * nc, cudaMemcpyHostToDevice, stream_);
  constexpr uint32_t KT = 64, CT = 64;
  const uint32_t grid = (uint32_t)((n_ + 127) / 128);
  for (uint32_t c0 = 0; c0 < nc; c0 += CT)
    for (uint32_t k0 = 0; k0 < m; k0 += KT) {
      const uint32_t kt = std::min(KT, m - k0), ct = std::min(CT, nc - c0);
      k_tall_times_small_tile<KT><<<grid, 128, sizeof(double) * kt * ct, stream_>>>(
          B, n_, k0, kt, dQ_, m, c0, ct, out, n_, n_, (accumulate || k0 > 0) ? 1 : 0);
    }
  cudaStreamSynchronize(stream_);  // Q's host storage may go away; dQ_ is reused by the next call
}

inline bool BlockKrylov::cholqr2(double* W) {
  for (int round = 0; round < 2; round++) {
    std::vector<double> G;
    vt_w(W, b_, W, G);  // b x b Gram matrix
    // Cholesky G = R' R (upper), then W <- W R^-1
    std::vector<double> R((size_t)b_ * b_, 0.0);
    for (uint32_t j = 0; j < b_; j++) {
      double s = G[(size_t)j * b_ + j];
      for (uint32_t k = 0; k < j; k++) s -= R[(size_t)j * b_ + k] * R[(size_t)j * b_ + k];
      if (!(s > 1e-24 * std::max(G[(size_t)j * b_ + j], DBL_MIN))) return false;
      const double rjj = sqrt(s);
      R[(size_t)j * b_ + j] = rjj;  // R(j, j) stored at column j, row j: R(k, j) = R[j * b + k]
      for (uint32_t i = j + 1; i < b_; i++) {
        double t = G[(size_t)i * b_ + j];
        for (uint32_t k = 0; k < j; k++) t -= R[(size_t)j * b_ + k] * R[(size_t)i * b_ + k];
        R[(size_t)i * b_ + j] = t / rjj;  // R(j, i)
      }
    }
    // Rinv (upper triangular), column-major b x b
    std::vector<double> Ri((size_t)b_ * b_, 0.0);
    for (uint32_t j = 0; j < b_; j++) {
      Ri[(size_t)j * b_ + j] = 1.0 / R[(size_t)j * b_ + j];
      for (int i = (int)j - 1; i >= 0; i--) {
        double t = 0.0;
        for (uint32_t k = (uint32_t)i + 1; k <= j; k++) t += R[(size_t)k * b_ + i] * Ri[(size_t)j * b_ + k];
        Ri[(size_t)j * b_ + i] = -t / R[(size_t)i * b_ + i];
      }
    }
    times_small(W, b_, Ri, b_, dT_, false);
    cudaMemcpyAsync(W, dT_, sizeof(double) * n_ * b_, cudaMemcpyDeviceToDevice, stream_);
  }
  return true;
}

inline int BlockKrylov::run(double tol, BlockResult& res) {
  auto now_s = [] {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  if (alloc()) return 1;
  error.clear();
  t_op = t_orth = t_ritz = 0;
  const double eps23 = pow(DBL_EPSILON, 2.0 / 3.0);
  // V_0 = orth(random block), generated on the device (counter-based hash: the same block on every
  // rank and every run)
  k_random_block<<<(uint32_t)((n_ * b_ + 255) / 256), 256, 0, stream_>>>(dV_, n_, b_, 20240601ull);
  if (!cholqr2(dV_)) {
    error = "block Krylov: start block is rank deficient";
    return 1;
  }
  std::vector<double> H((size_t)mmax_ * mmax_, 0.0);  // column-major, leading dimension mmax_
  std::vector<double> C, C2, theta, Sm;
  res = BlockResult();
  res.evals.assign(nev_, 0.0);
  uint32_t m = 0;
  bool done = false;
  for (uint32_t j = 0; j < maxp_ && !done; j++) {
    double* Vj = dV_ + (uint64_t)j * b_ * n_;
    double* AVj = dAV_ + (uint64_t)j * b_ * n_;
    double ta = now_s();
    if (op_(Vj, b_, AVj)) {
      error = "block Krylov: operator failed";
      return 1;
    }
    m = (j + 1) * b_;
    res.npasses = j + 1;
    res.nvecprod = m;
    // H[:, j-th block] = V[:, :m]' (A V_j), and its mirror image
    vt_w(dV_, m, AVj, C);
    double tb = now_s();
    t_op += tb - ta;
    for (uint32_t t = 0; t < b_; t++)
      for (uint32_t c = 0; c < m; c++) {
        const double v = C[(size_t)t * m + c];
        H[(size_t)(j * b_ + t) * mmax_ + c] = v;
        H[(size_t)c * mmax_ + (j * b_ + t)] = v;
      }
    // Rayleigh-Ritz once the space can hold the wanted pairs
    if (m >= nev_ || m + b_ > n_) {
      const uint32_t kk = std::min<uint32_t>(nev_, m);
      std::vector<double> Hm((size_t)m * m);
      for (uint32_t c = 0; c < m; c++)
        for (uint32_t r = 0; r < m; r++)
          Hm[(size_t)c * m + r] = 0.5 * (H[(size_t)c * mmax_ + r] + H[(size_t)r * mmax_ + c]);
      std::vector<double> w, Z;
      if (!sym_eigen((int)m, Hm, w, Z)) {
        error = "block Krylov: projected eigenproblem failed";
        return 1;
      }
      std::vector<uint32_t> ord(m);
      std::iota(ord.begin(), ord.end(), 0u);
      std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t bq) { return w[a] > w[bq]; });
      theta.assign(kk, 0.0);
      Sm.assign((size_t)m * kk, 0.0);
      for (uint32_t i = 0; i < kk; i++) {
        theta[i] = w[ord[i]];
        for (uint32_t r = 0; r < m; r++) Sm[(size_t)i * m + r] = Z[(size_t)ord[i] * m + r];
      }
      times_small(dV_, m, Sm, kk, dU_, false);    // U = V S
      times_small(dAV_, m, Sm, kk, dAU_, false);  // A U = (A V) S
      cudaMemcpyAsync(dSmall_, theta.data(), sizeof(double) * kk, cudaMemcpyHostToDevice, stream_);
      k_check_resid<<<kk, 1024, 0, stream_>>>(dAU_, dU_, dSmall_, n_, 1.0, dSmall_ + kk);
      std::vector<double> r2(kk);
      cudaMemcpyAsync(r2.data(), dSmall_ + kk, sizeof(double) * kk, cudaMemcpyDeviceToHost, stream_);
      cudaStreamSynchronize(stream_);
      uint32_t nconv = 0;
      for (uint32_t i = 0; i < kk; i++)
        if (sqrt(r2[i]) < tol * std::max(eps23, fabs(theta[i]))) nconv++;
      res.nconv = nconv;
      for (uint32_t i = 0; i < kk; i++) res.evals[i] = theta[i];
      if ((nconv >= nev_ && kk == nev_) || m + b_ > n_) done = true;
      t_ritz += now_s() - tb;
      tb = now_s();
    }
    if (done || j + 1 >= maxp_ || m + b_ > n_) break;
    // next block: W = A V_j orthogonalised against the whole basis (twice), then orth(W)
    cudaMemcpyAsync(dW_, AVj, sizeof(double) * n_ * b_, cudaMemcpyDeviceToDevice, stream_);
    for (auto& v : C) v = -v;
    times_small(dV_, m, C, b_, dW_, true);
    vt_w(dV_, m, dW_, C2);
    for (auto& v : C2) v = -v;
    times_small(dV_, m, C2, b_, dW_, true);
    if (!cholqr2(dW_)) {  // invariant subspace: the space built so far holds everything reachable
      done = true;
      break;
    }
    cudaMemcpyAsync(dV_ + (uint64_t)m * n_, dW_, sizeof(double) * n_ * b_, cudaMemcpyDeviceToDevice,
                    stream_);
    t_orth += now_s() - tb;
  }
  cudaStreamSynchronize(stream_);
  return 0;
}

}  // namespace fpb
