// fpb_irlm.cuh -- device-resident implicitly restarted Lanczos.
//
// Follows the schedule of Spectra 0.8.1 SymEigsSolver<double, LARGEST_ALGE, Op>
// as flashpca drives it (randompca.cpp:174-190: nev = ndim, ncv = 2*ndim+1,
// init() then compute(maxiter, tol)).  Spectra is a third-party header library
// that is not part of the flashpca tree; the steps below restate its published
// algorithm (ARPACK dsaup2-style IRLM with full re-orthogonalisation).
//
// The Krylov basis V (N x ncv), residual f and work vector w stay in HBM; the
// ncv x ncv projected matrix H lives on the host.  All reductions are two-stage
// with a fixed order (no atomics) so that SNP-sharded ranks, which run this
// driver redundantly on identical all-reduced vectors, stay bit-identical and
// take the same convergence decisions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

namespace fpb {

constexpr int kMaxNcv = 128;      // columns one launch of the tall-skinny kernels handles (host loops)
constexpr int kRowsPerBlock = 2048;

// partial[g * m + c] = sum over rows of block g of V[r + c*ld] * f[r]
__global__ void __launch_bounds__(256)
k_gemv_t_partial(const double* __restrict__ V, uint64_t ld, uint32_t m,
                 const double* __restrict__ f, uint64_t n, double* __restrict__ partial) {
  __shared__ double sh[8][kMaxNcv];
  const int RPT = kRowsPerBlock / 256;
  uint64_t r0 = (uint64_t)blockIdx.x * kRowsPerBlock + threadIdx.x;
  double fr[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    uint64_t r = r0 + (uint64_t)q * 256;
    fr[q] = r < n ? f[r] : 0.0;
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t c = 0; c < m; c++) {
    const double* col = V + (uint64_t)c * ld;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < RPT; q++) {
      uint64_t r = r0 + (uint64_t)q * 256;
      if (r < n) s += col[r] * fr[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sh[warp][c] = s;
  }
  __syncthreads();
  for (uint32_t c = threadIdx.x; c < m; c += 256) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) s += sh[w][c];
    partial[(uint64_t)blockIdx.x * m + c] = s;
  }
}

// out[c] = sum_g partial[g * m + c]; one warp per column, fixed summation order
__global__ void __launch_bounds__(256)
k_gemv_t_final(const double* __restrict__ partial, uint32_t nblocks, uint32_t m,
               double* __restrict__ out) {
  uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (c >= m) return;
  double s = 0.0;
  for (uint32_t g = lane; g < nblocks; g += 32) s += partial[(uint64_t)g * m + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[c] = s;
}

// One pass over V[:, :m] per re-orthogonalisation step (256 rows per block, a
// row per thread kept in registers):
//   f[r]  = fin[r] - sum_c V[r,c] h[c]          (h may be null: f = fin)
//   partial[g][c] = sum_{r in block g} V[r,c] f[r]   (c < m),  partial[g][m] = sum f[r]^2
template <int MAXM>
__global__ void __launch_bounds__(256)
k_fused_reorth(const double* __restrict__ V, uint64_t ld, uint32_t m, const double* __restrict__ h,
               const double* fin, double* f, uint64_t n, double* __restrict__ partial) {
  __shared__ double hs[MAXM];
  __shared__ double red[8][MAXM + 1];
  if (h)
    for (uint32_t c = threadIdx.x; c < m; c += 256) hs[c] = h[c];
  __syncthreads();
  const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const bool valid = r < n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double vr[MAXM];
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < MAXM; c++) {
    vr[c] = (valid && (uint32_t)c < m) ? V[r + (uint64_t)c * ld] : 0.0;
    if (h && (uint32_t)c < m) s += vr[c] * hs[c];
  }
  const double fn = valid ? fin[r] - s : 0.0;
  if (valid) f[r] = fn;
#pragma unroll
  for (int c = 0; c < MAXM; c++) {
    if ((uint32_t)c < m) {
      double p = vr[c] * fn;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
      if (lane == 0) red[warp][c] = p;
    }
  }
  double nn = fn * fn;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  if (lane == 0) red[warp][m] = nn;
  __syncthreads();
  for (uint32_t c = threadIdx.x; c <= m; c += 256) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += red[w][c];
    partial[(uint64_t)blockIdx.x * (m + 1) + c] = t;
  }
}

// f[r] -= sum_c V[r + c*ld] * h[c]
__global__ void __launch_bounds__(256)
k_gemv_n_sub(const double* __restrict__ V, uint64_t ld, uint32_t m, const double* __restrict__ h,
             double* __restrict__ f, uint64_t n) {
  __shared__ double hs[kMaxNcv];
  for (uint32_t c = threadIdx.x; c < m; c += blockDim.x) hs[c] = h[c];
  __syncthreads();
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (uint32_t c = 0; c < m; c++) s += V[r + (uint64_t)c * ld] * hs[c];
  f[r] -= s;
}

// f = w - b * vprev - a * vcur   (vprev may be null when b is unused)
__global__ void k_resid(const double* __restrict__ w, const double* __restrict__ vprev, double b,
                        const double* __restrict__ vcur, double a, double* __restrict__ f,
                        uint64_t n) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double v = w[r] - a * vcur[r];
  if (vprev) v -= b * vprev[r];
  f[r] = v;
}

// out = out2 = a * x: the new Lanczos vector goes to its column of V and to the fixed buffer the
// operator reads (one (x, y) pointer pair for every op of the solve: one CUDA graph, fpb_capi.cu)
__global__ void k_scale2(double a, const double* __restrict__ x, double* __restrict__ out,
                         double* __restrict__ out2, uint64_t n) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const double v = a * x[r];
  out[r] = v;
  out2[r] = v;
}

// out = a * x + b * y  (y may be null)
__global__ void k_axpby(double a, const double* x, double b, const double* y, double* out,
                        uint64_t n) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double v = a * x[r];
  if (y) v += b * y[r];
  out[r] = v;
}

// out[r, c] = sum_k V[r, k] * Q[k + c*ldq], c < nc, k < m.  The small factor
// (m x nc <= ncv x ncv) sits in shared memory; one thread per row keeps its V
// row in registers.
template <int MAXM>
__global__ void __launch_bounds__(128)
k_tall_times_small(const double* __restrict__ V, uint64_t ld, uint32_t m,
                   const double* __restrict__ Q, uint32_t ldq, uint32_t nc,
                   double* __restrict__ out, uint64_t ldo, uint64_t n) {
  extern __shared__ double qs[];
  for (uint32_t e = threadIdx.x; e < m * nc; e += blockDim.x) {
    uint32_t c = e / m, k = e - c * m;
    qs[e] = Q[k + (uint64_t)c * ldq];
  }
  __syncthreads();
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double vr[MAXM];
#pragma unroll
  for (int k = 0; k < MAXM; k++) vr[k] = (uint32_t)k < m ? V[r + (uint64_t)k * ld] : 0.0;
  for (uint32_t c = 0; c < nc; c++) {
    const double* qc = qs + c * m;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < MAXM; k++)
      if ((uint32_t)k < m) s += vr[k] * qc[k];
    out[r + (uint64_t)c * ldo] = s;
  }
}

// m > MAXM of the register-cached kernel above: the product is accumulated over K-tiles of KT
// columns of V (one launch per tile).  out (+)= V[:, k0:k0+kt] * Q[k0:k0+kt, c0:c0+nc]; the Q tile
// sits in shared memory (kt x nc doubles), a thread keeps its KT values of the V row in registers.
template <int KT>
__global__ void __launch_bounds__(128)
k_tall_times_small_tile(const double* __restrict__ V, uint64_t ld, uint32_t k0, uint32_t kt,
                        const double* __restrict__ Q, uint32_t ldq, uint32_t c0, uint32_t nc,
                        double* __restrict__ out, uint64_t ldo, uint64_t n, int accumulate) {
  extern __shared__ double qs[];
  for (uint32_t e = threadIdx.x; e < kt * nc; e += blockDim.x) {
    const uint32_t c = e / kt, k = e - c * kt;
    qs[e] = Q[(k0 + k) + (uint64_t)(c0 + c) * ldq];
  }
  __syncthreads();
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double vr[KT];
#pragma unroll
  for (int k = 0; k < KT; k++) vr[k] = (uint32_t)k < kt ? V[r + (uint64_t)(k0 + k) * ld] : 0.0;
  for (uint32_t c = 0; c < nc; c++) {
    const double* qc = qs + c * kt;
    double s = accumulate ? out[r + (uint64_t)(c0 + c) * ldo] : 0.0;
#pragma unroll
    for (int k = 0; k < KT; k++)
      if ((uint32_t)k < kt) s += vr[k] * qc[k];
    out[r + (uint64_t)(c0 + c) * ldo] = s;
  }
}

// Reference form (kept for cross-checks).
__global__ void __launch_bounds__(128)
k_tall_times_small_generic(const double* __restrict__ V, uint64_t ld, uint32_t m,
                           const double* __restrict__ Q, uint32_t ldq, uint32_t nc,
                           double* __restrict__ out, uint64_t ldo, uint64_t n) {
  uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (uint32_t c = 0; c < nc; c++) {
    double s = 0.0;
    for (uint32_t k = 0; k < m; k++) s += V[r + (uint64_t)k * ld] * Q[k + (uint64_t)c * ldq];
    out[r + (uint64_t)c * ldo] = s;
  }
}

// err[c] = sum_r (Y[r, c] / div - U[r, c] lambda[c] / div)^2: the per-eigenvector squared residual
// of RandomPCA::check (randompca.cpp:663-703).  One block per column, fixed summation order.
__global__ void __launch_bounds__(1024)
k_check_resid(const double* __restrict__ Y, const double* __restrict__ U,
              const double* __restrict__ lambda, uint64_t n, double div, double* __restrict__ err) {
  __shared__ double sh[32];
  const uint32_t c = blockIdx.x;
  const double d = lambda[c] / div;
  const double* y = Y + (uint64_t)c * n;
  const double* u = U + (uint64_t)c * n;
  double s = 0.0;
  for (uint64_t r = threadIdx.x; r < n; r += 1024) {
    const double e = y[r] / div - u[r] * d;
    s += e * e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = sh[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) err[c] = s;
  }
}

// ------------------------------- host algebra -------------------------------

// Spectra SimpleRandom<double>: Park-Miller LCG, a = 16807, m = 2^31 - 1,
// seed 0 -> 1, uniform(-0.5, 0.5).
inline void simple_random_vec(std::vector<double>& out, uint64_t n, unsigned long seed) {
  const unsigned long m = 2147483647UL, a = 16807UL;
  unsigned long r = seed ? (seed & m) : 1UL;
  out.resize(n);
  for (uint64_t i = 0; i < n; i++) {
    r = (a * r) % m;
    out[i] = (double)r / (double)m - 0.5;
  }
}

inline unsigned long long simple_random_r0(unsigned long seed) {
  return seed ? (seed & 2147483647UL) : 1UL;
}

// The same vector on the device: element i is r0 a^(i+1) mod m, so every thread jumps ahead on its
// own (31 squarings); bit-identical to simple_random_vec.  (The host loop plus the pageable 8 N
// byte upload cost ~5 ms per solve at N = 500,000: 2 % of a 1-GPU solve, 9 % of an 8-GPU one.)
__global__ void __launch_bounds__(256)
k_simple_random(double* __restrict__ out, uint64_t n, unsigned long long r0) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long m = 2147483647ull;
  unsigned long long e = i + 1, base = 16807ull, acc = r0 % m;
  while (e) {
    if (e & 1) acc = (acc * base) % m;
    base = (base * base) % m;
    e >>= 1;
  }
  out[i] = (double)acc / (double)m - 0.5;
}

// coefficient vector of a Lanczos step's first pass, built on the device:
// h = (0, ..., 0, beta_prev, Hii) with Hii = v_i . w read from *hii
__global__ void k_first_coeffs(double* __restrict__ h, uint32_t m, double beta_prev,
                               const double* __restrict__ hii) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  h[c] = c + 1 == m ? *hii : (c + 2 == m ? beta_prev : 0.0);
}

// Eigen-decomposition of a symmetric tridiagonal matrix by implicit-shift QL.
// d: diagonal (n), e: sub-diagonal (e[i] couples i and i+1; e[n-1] unused),
// z: n x n column-major, returns eigenvectors in columns.  false on failure.
inline bool tridiag_eigen(int n, std::vector<double>& d, std::vector<double>& e,
                          std::vector<double>& z) {
  z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) z[(size_t)i * n + i] = 1.0;
  e[n - 1] = 0.0;
  for (int l = 0; l < n; l++) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; m++) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
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
          double f = s * e[i], b = c * e[i];
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
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          for (int k = 0; k < n; k++) {
            double* zi = &z[(size_t)i * n + k];
            double* zi1 = &z[(size_t)(i + 1) * n + k];
            f = *zi1;
            *zi1 = s * (*zi) + c * f;
            *zi = c * (*zi) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// One shifted QR step on the (tridiagonal, symmetric) projected matrix:
// H - mu I = QR;  H <- RQ + mu I;  Qacc <- Qacc * Q.   (Spectra restart())
inline void tridiag_qr_step(int n, std::vector<double>& H, double mu, std::vector<double>& Qacc) {
  auto A = [&](int r, int c) -> double& { return H[(size_t)c * n + r]; };
  std::vector<double> cs(n), sn(n);
  for (int i = 0; i < n; i++) A(i, i) -= mu;
  for (int i = 0; i < n - 1; i++) {
    double x = A(i, i), y = A(i + 1, i);
    double r = hypot(x, y), c = 1.0, s = 0.0;
    if (r > 0.0) {
      c = x / r;
      s = y / r;
    }
    cs[i] = c;
    sn[i] = s;
    for (int k = 0; k < n; k++) {
      double a = A(i, k), b = A(i + 1, k);
      A(i, k) = c * a + s * b;
      A(i + 1, k) = -s * a + c * b;
    }
  }
  for (int i = 0; i < n - 1; i++) {
    double c = cs[i], s = sn[i];
    for (int k = 0; k < n; k++) {
      double a = A(k, i), b = A(k, i + 1);
      A(k, i) = c * a + s * b;
      A(k, i + 1) = -s * a + c * b;
      double qa = Qacc[(size_t)i * n + k], qb = Qacc[(size_t)(i + 1) * n + k];
      Qacc[(size_t)i * n + k] = c * qa + s * qb;
      Qacc[(size_t)(i + 1) * n + k] = -s * qa + c * qb;
    }
  }
  // keep H exactly symmetric tridiagonal, as Spectra's TridiagQR::matrix_RQ does
  std::vector<double> dg(n), sb(n);
  for (int i = 0; i < n; i++) dg[i] = A(i, i) + mu;
  for (int i = 0; i + 1 < n; i++) sb[i] = A(i + 1, i);
  std::fill(H.begin(), H.end(), 0.0);
  for (int i = 0; i < n; i++) A(i, i) = dg[i];
  for (int i = 0; i + 1 < n; i++) {
    A(i + 1, i) = sb[i];
    A(i, i + 1) = sb[i];
  }
}

struct IrlmResult {
  std::vector<double> evals;  // nev, descending, only the first nconv_sorted are converged-flagged
  std::vector<char> conv;     // nev flags after sorting
  uint32_t nconv = 0, nops = 0, niter = 0;
};

// Device-vector IRLM.  `op(d_in, d_out)` enqueues A * in -> out on `stream`.
class Irlm {
 public:
  Irlm(uint64_t n, uint32_t nev, uint32_t ncv, cudaStream_t stream,
       std::function<void(const double*, double*)> op)
      : n_(n), nev_(nev), ncv_(ncv), stream_(stream), op_(std::move(op)) {}
  ~Irlm() { release(); }
  void set_op(std::function<void(const double*, double*)> op) { op_ = std::move(op); }
  bool matches(uint64_t n, uint32_t nev, uint32_t ncv) const {
    return n == n_ && nev == nev_ && ncv == ncv_;
  }

  // Runs init() + compute(maxit, tol).  On return d_V()/ritz vectors can be
  // combined with eigenvectors().
  void run(uint32_t maxit, double tol, IrlmResult& res);
  // evecs (N x nev, column-major, device) = V * ritz_vec, sorted like evals.
  void eigenvectors(double* d_out);
  std::string error;
  double t_op = 0, t_hii = 0, t_reorth = 0, t_other = 0;  // host seconds by phase (trace)
  uint32_t n_reorth = 0;

 private:
  static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
  }
  void alloc();
  void release();
  void gemv_t(const double* V, uint32_t m, const double* f, double* host_out);
  double norm(const double* f) {
    double s;
    gemv_t(f, 1, f, &s);
    return sqrt(s);
  }
  // fused pass: f = fin - V[:, :m] h (h_host may be null), returns V'f in vf[0..m) and ||f||
  double reorth_pass(uint32_t m, const double* h_host, const double* fin, double* vf);
  double reorth_first(uint32_t i, double beta_prev, double* vf, double* hii_out);
  void factorize_from(uint32_t from_k, uint32_t to_m);
  void retrieve_ritzpair();
  // d_out (N x nc) = V (N x ncv) * dQ (ncv x nc)
  void tall_times_small(const double* dQ, uint32_t nc, double* d_out) {
    if (ncv_ <= 48) {
      k_tall_times_small<48><<<grid1d(128), 128, sizeof(double) * ncv_ * nc, stream_>>>(
          dV_, n_, ncv_, dQ, ncv_, nc, d_out, n_, n_);
      return;
    }
    // any ncv (ndim up to (min(N, P) - 1) / 2, flashpca.cpp:623-633): K-tiles of 64 columns of V,
    // column tiles of 64 outputs (32 KB of shared memory per launch)
    constexpr uint32_t KT = 64, CT = 64;
    for (uint32_t c0 = 0; c0 < nc; c0 += CT)
      for (uint32_t k0 = 0; k0 < ncv_; k0 += KT) {
        const uint32_t kt = std::min(KT, ncv_ - k0), ct = std::min(CT, nc - c0);
        k_tall_times_small_tile<KT><<<grid1d(128), 128, sizeof(double) * kt * ct, stream_>>>(
            dV_, n_, k0, kt, dQ, ncv_, c0, ct, d_out, n_, n_, k0 > 0);
      }
  }
  // f -= V[:, :m] h for any m (chunks of kMaxNcv columns; h on the device)
  void gemv_n_sub(uint32_t m, const double* d_h) {
    for (uint32_t c0 = 0; c0 < m; c0 += kMaxNcv)
      k_gemv_n_sub<<<grid1d(256), 256, 0, stream_>>>(dV_ + (uint64_t)c0 * n_, n_,
                                                      std::min<uint32_t>(kMaxNcv, m - c0), d_h + c0,
                                                      dF_, n_);
  }
  double& H(uint32_t r, uint32_t c) { return H_[(size_t)c * ncv_ + r]; }
  double* col(uint32_t i) { return dV_ + (uint64_t)i * n_; }
  uint32_t grid1d(uint32_t bs) const { return (uint32_t)((n_ + bs - 1) / bs); }

  uint64_t n_;
  uint32_t nev_, ncv_;
  cudaStream_t stream_;
  std::function<void(const double*, double*)> op_;
  double *dV_ = nullptr, *dF_ = nullptr, *dW_ = nullptr, *dVs_ = nullptr, *dPartial_ = nullptr,
         *dSmall_ = nullptr, *dQ_ = nullptr, *dX_ = nullptr;
  uint32_t nblocks_ = 0, nblocks256_ = 0;
  double *dPartial2_ = nullptr, *dH_ = nullptr, *hPinned_ = nullptr;
  bool fused_ = false;
  std::vector<double> H_, ritz_val_, ritz_est_, ritz_vec_;  // ritz_vec_: ncv x nev
  double beta_ = 0.0;
  uint32_t nops_ = 0;
  std::vector<uint32_t> order_;
};

inline void Irlm::alloc() {
  if (dV_) {  // workspace kept from a previous solve on the same handle
    cudaMemsetAsync(dV_, 0, sizeof(double) * n_ * ncv_, stream_);
    H_.assign((size_t)ncv_ * ncv_, 0.0);
    error.clear();
    t_op = t_hii = t_reorth = t_other = 0;
    n_reorth = 0;
    return;
  }
  nblocks_ = (uint32_t)((n_ + kRowsPerBlock - 1) / kRowsPerBlock);
  cudaMalloc(&dV_, sizeof(double) * n_ * ncv_);
  cudaMalloc(&dVs_, sizeof(double) * n_ * ncv_);
  cudaMalloc(&dF_, sizeof(double) * n_);
  cudaMalloc(&dW_, sizeof(double) * n_);
  cudaMalloc(&dX_, sizeof(double) * n_);
  cudaMalloc(&dPartial_, sizeof(double) * (size_t)nblocks_ * ncv_);
  cudaMalloc(&dSmall_, sizeof(double) * (ncv_ + 2));
  cudaMalloc(&dQ_, sizeof(double) * ncv_ * ncv_);
  fused_ = ncv_ <= 48;
  nblocks256_ = (uint32_t)((n_ + 255) / 256);
  cudaMalloc(&dPartial2_, sizeof(double) * (size_t)nblocks256_ * (ncv_ + 1));
  cudaMalloc(&dH_, sizeof(double) * (ncv_ + 1));
  cudaMallocHost(&hPinned_, sizeof(double) * (ncv_ + 2));
  cudaMemsetAsync(dV_, 0, sizeof(double) * n_ * ncv_, stream_);
  H_.assign((size_t)ncv_ * ncv_, 0.0);
}

inline void Irlm::release() {
  cudaFree(dV_); cudaFree(dVs_); cudaFree(dF_); cudaFree(dW_); cudaFree(dX_);
  cudaFree(dPartial_); cudaFree(dSmall_); cudaFree(dQ_);
  cudaFree(dPartial2_); cudaFree(dH_);
  if (hPinned_) cudaFreeHost(hPinned_);
  dV_ = dVs_ = dF_ = dW_ = dX_ = dPartial_ = dSmall_ = dQ_ = dPartial2_ = dH_ = hPinned_ = nullptr;
}

inline void Irlm::gemv_t(const double* V, uint32_t m, const double* f, double* host_out) {
  for (uint32_t c0 = 0; c0 < m; c0 += kMaxNcv) {  // the partial kernel handles kMaxNcv columns
    const uint32_t mc = std::min<uint32_t>(kMaxNcv, m - c0);
    double* part = dPartial_ + (uint64_t)nblocks_ * c0;
    k_gemv_t_partial<<<nblocks_, 256, 0, stream_>>>(V + (uint64_t)c0 * n_, n_, mc, f, n_, part);
    k_gemv_t_final<<<(mc * 32 + 255) / 256, 256, 0, stream_>>>(part, nblocks_, mc, dSmall_ + c0);
  }
  cudaMemcpyAsync(host_out, dSmall_, sizeof(double) * m, cudaMemcpyDeviceToHost, stream_);
  cudaStreamSynchronize(stream_);
}

inline double Irlm::reorth_pass(uint32_t m, const double* h_host, const double* fin, double* vf) {
  if (h_host) {
    memcpy(hPinned_, h_host, sizeof(double) * m);
    cudaMemcpyAsync(dH_, hPinned_, sizeof(double) * m, cudaMemcpyHostToDevice, stream_);
  }
  k_fused_reorth<48><<<nblocks256_, 256, 0, stream_>>>(dV_, n_, m, h_host ? dH_ : nullptr, fin,
                                                        dF_, n_, dPartial2_);
  k_gemv_t_final<<<((m + 1) * 32 + 255) / 256, 256, 0, stream_>>>(dPartial2_, nblocks256_, m + 1,
                                                                   dSmall_);
  cudaMemcpyAsync(hPinned_, dSmall_, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, stream_);
  cudaStreamSynchronize(stream_);
  memcpy(vf, hPinned_, sizeof(double) * m);
  return sqrt(hPinned_[m]);
}

// First pass of Lanczos step i without a host round trip for Hii: Hii = v_i . w goes straight
// into the coefficient vector on the device, f = w - beta v_{i-1} - Hii v_i, V'f and ||f|| follow
// in the same stream, and one synchronisation returns all of them (same arithmetic as
// gemv_t + reorth_pass: Hii is the same double either way).
inline double Irlm::reorth_first(uint32_t i, double beta_prev, double* vf, double* hii_out) {
  const uint32_t m = i + 1;
  k_gemv_t_partial<<<nblocks_, 256, 0, stream_>>>(col(i), n_, 1, dW_, n_, dPartial_);
  k_gemv_t_final<<<1, 256, 0, stream_>>>(dPartial_, nblocks_, 1, dSmall_ + m + 1);
  k_first_coeffs<<<(m + 63) / 64, 64, 0, stream_>>>(dH_, m, beta_prev, dSmall_ + m + 1);
  k_fused_reorth<48><<<nblocks256_, 256, 0, stream_>>>(dV_, n_, m, dH_, dW_, dF_, n_, dPartial2_);
  k_gemv_t_final<<<((m + 1) * 32 + 255) / 256, 256, 0, stream_>>>(dPartial2_, nblocks256_, m + 1,
                                                                   dSmall_);
  cudaMemcpyAsync(hPinned_, dSmall_, sizeof(double) * (m + 2), cudaMemcpyDeviceToHost, stream_);
  cudaStreamSynchronize(stream_);
  memcpy(vf, hPinned_, sizeof(double) * m);
  *hii_out = hPinned_[m + 1];
  return sqrt(hPinned_[m]);
}

inline void Irlm::factorize_from(uint32_t from_k, uint32_t to_m) {
  if (to_m <= from_k) return;
  const double eps = DBL_EPSILON, near0 = DBL_MIN * 10.0;
  double beta = norm(dF_);
  for (uint32_t c = from_k; c < ncv_; c++)
    for (uint32_t r = 0; r < ncv_; r++) H(r, c) = 0.0;
  for (uint32_t r = from_k; r < ncv_; r++)
    for (uint32_t c = 0; c < from_k; c++) H(r, c) = 0.0;
  std::vector<double> Vf(ncv_);
  for (uint32_t i = from_k; i < to_m; i++) {
    bool restart = false;
    if (beta < near0) {
      // invariant subspace: new random direction orthogonal to V[:, :i]
      k_simple_random<<<grid1d(256), 256, 0, stream_>>>(dF_, n_, simple_random_r0(2UL * i));
      gemv_t(dV_, i, dF_, Vf.data());
      cudaMemcpyAsync(dSmall_, Vf.data(), sizeof(double) * i, cudaMemcpyHostToDevice, stream_);
      gemv_n_sub(i, dSmall_);
      beta = norm(dF_);
      restart = true;
    }
    double ta = now_s();
    k_scale2<<<grid1d(256), 256, 0, stream_>>>(1.0 / beta, dF_, col(i), dX_, n_);
    H(i, i - 1) = restart ? 0.0 : beta;
    op_(dX_, dW_);
    nops_++;
    double tb = now_s();
    double Hii;
    const uint32_t i1 = i + 1;
    double tc = tb;
    if (fused_) {
      // Hii, f = w - H(i,i-1) v_{i-1} - Hii v_i, V'f and ||f||: one pass over V[:, :i+1], one sync
      beta = reorth_first(i, restart ? 0.0 : H(i, i - 1), Vf.data(), &Hii);
      tc = now_s();
    } else {
      gemv_t(col(i), 1, dW_, &Hii);
      tc = now_s();
    }
    t_op += tb - ta;
    t_hii += tc - tb;
    H(i - 1, i) = H(i, i - 1);
    H(i, i) = Hii;
    if (!fused_) {
      k_resid<<<grid1d(256), 256, 0, stream_>>>(dW_, restart ? nullptr : col(i - 1), H(i, i - 1),
                                                col(i), Hii, dF_, n_);
      beta = norm(dF_);
      gemv_t(dV_, i1, dF_, Vf.data());
    }
    auto maxabs = [&]() {
      double m = 0.0;
      for (uint32_t c = 0; c < i1; c++) m = std::max(m, fabs(Vf[c]));
      return m;
    };
    double ortho_err = maxabs();
    int count = 0;
    while (count < 5 && ortho_err > eps * beta) {
      if (beta < near0) {
        cudaMemsetAsync(dF_, 0, sizeof(double) * n_, stream_);
        beta = 0.0;
        break;
      }
      std::vector<double> hprev(Vf.begin(), Vf.begin() + i1);
      H(i - 1, i) += hprev[i - 1];
      H(i, i - 1) = H(i - 1, i);
      H(i, i) += hprev[i];
      if (fused_) {
        beta = reorth_pass(i1, hprev.data(), dF_, Vf.data());
      } else {
        cudaMemcpyAsync(dSmall_, hprev.data(), sizeof(double) * i1, cudaMemcpyHostToDevice,
                        stream_);
        gemv_n_sub(i1, dSmall_);
        cudaStreamSynchronize(stream_);
        beta = norm(dF_);
        gemv_t(dV_, i1, dF_, Vf.data());
      }
      ortho_err = maxabs();
      count++;
      n_reorth++;
    }
    n_reorth++;
    t_reorth += now_s() - tc;
  }
  beta_ = beta;
}

inline void Irlm::retrieve_ritzpair() {
  std::vector<double> d(ncv_), e(ncv_, 0.0), z;
  for (uint32_t i = 0; i < ncv_; i++) d[i] = H(i, i);
  for (uint32_t i = 0; i + 1 < ncv_; i++) e[i] = H(i + 1, i);
  if (!tridiag_eigen((int)ncv_, d, e, z)) error = "tridiagonal eigen-decomposition failed";
  std::vector<uint32_t> ind(ncv_);
  std::iota(ind.begin(), ind.end(), 0u);
  std::stable_sort(ind.begin(), ind.end(), [&](uint32_t a, uint32_t b) { return d[a] > d[b]; });
  ritz_val_.resize(ncv_);
  ritz_est_.resize(ncv_);
  ritz_vec_.assign((size_t)ncv_ * nev_, 0.0);
  for (uint32_t i = 0; i < ncv_; i++) {
    ritz_val_[i] = d[ind[i]];
    ritz_est_[i] = z[(size_t)ind[i] * ncv_ + (ncv_ - 1)];
  }
  for (uint32_t i = 0; i < nev_; i++)
    for (uint32_t r = 0; r < ncv_; r++) ritz_vec_[(size_t)i * ncv_ + r] = z[(size_t)ind[i] * ncv_ + r];
}

inline void Irlm::run(uint32_t maxit, double tol, IrlmResult& res) {
  const double eps = DBL_EPSILON, near0 = DBL_MIN * 10.0, eps23 = pow(eps, 2.0 / 3.0);
  alloc();
  nops_ = 0;
  // init(): v0 = SimpleRandom(0) normalised; w = A v0; H00 = v0.w; f = w - H00 v0
  k_simple_random<<<grid1d(256), 256, 0, stream_>>>(dF_, n_, simple_random_r0(0));
  double vnorm = norm(dF_);
  k_scale2<<<grid1d(256), 256, 0, stream_>>>(1.0 / vnorm, dF_, col(0), dX_, n_);
  op_(dX_, dW_);
  nops_++;
  double h00;
  gemv_t(col(0), 1, dW_, &h00);
  H(0, 0) = h00;
  k_resid<<<grid1d(256), 256, 0, stream_>>>(dW_, nullptr, 0.0, col(0), h00, dF_, n_);

  factorize_from(1, ncv_);
  retrieve_ritzpair();

  uint32_t nconv = 0, it = 0;
  std::vector<char> conv(nev_, 0);
  for (it = 0; it < maxit; it++) {
    nconv = 0;
    for (uint32_t i = 0; i < nev_; i++) {
      double thresh = tol * std::max(eps23, fabs(ritz_val_[i]));
      double resid = fabs(ritz_est_[i]) * beta_;
      conv[i] = resid < thresh;
      nconv += conv[i];
    }
    if (nconv >= nev_) break;
    // nev_adjusted(): ARPACK dsaup2 rule
    uint32_t nev_new = nev_;
    for (uint32_t i = nev_; i < ncv_; i++)
      if (fabs(ritz_est_[i]) < near0) nev_new++;
    nev_new += std::min(nconv, (ncv_ - nev_new) / 2);
    if (nev_new == 1 && ncv_ >= 6) nev_new = ncv_ / 2;
    else if (nev_new == 1 && ncv_ > 2) nev_new = 2;
    if (nev_new > ncv_ - 1) nev_new = ncv_ - 1;
    const uint32_t k = nev_new;
    // restart(k): shifts = unwanted Ritz values
    std::vector<double> Q((size_t)ncv_ * ncv_, 0.0);
    for (uint32_t i = 0; i < ncv_; i++) Q[(size_t)i * ncv_ + i] = 1.0;
    for (uint32_t i = k; i < ncv_; i++) tridiag_qr_step((int)ncv_, H_, ritz_val_[i], Q);
    // V[:, :k+1] <- V Q[:, :k+1]
    cudaMemcpyAsync(dQ_, Q.data(), sizeof(double) * ncv_ * ncv_, cudaMemcpyHostToDevice, stream_);
    tall_times_small(dQ_, k + 1, dVs_);
    cudaMemcpyAsync(dV_, dVs_, sizeof(double) * n_ * (k + 1), cudaMemcpyDeviceToDevice, stream_);
    // f <- f * Q(ncv-1, k-1) + V[:, k] * H(k, k-1)
    k_axpby<<<grid1d(256), 256, 0, stream_>>>(Q[(size_t)(k - 1) * ncv_ + (ncv_ - 1)], dF_,
                                              H(k, k - 1), col(k), dF_, n_);
    factorize_from(k, ncv_);
    retrieve_ritzpair();
    if (!error.empty()) break;
  }
  // sort_ritzpair(LARGEST_ALGE) over the first nev
  order_.resize(nev_);
  std::iota(order_.begin(), order_.end(), 0u);
  std::stable_sort(order_.begin(), order_.end(),
                   [&](uint32_t a, uint32_t b) { return ritz_val_[a] > ritz_val_[b]; });
  res.evals.resize(nev_);
  res.conv.resize(nev_);
  for (uint32_t i = 0; i < nev_; i++) {
    res.evals[i] = ritz_val_[order_[i]];
    res.conv[i] = conv[order_[i]];
  }
  if (getenv("FPB_IRLM_TRACE"))
    fprintf(stderr, "[irlm] ops %u: enqueue-op %.1f ms, wait-op + Hii + first pass %.1f ms, further passes %.1f ms (%u passes in all)\n",
            nops_, t_op * 1e3, t_hii * 1e3, t_reorth * 1e3, n_reorth);
  res.nconv = nconv;
  res.nops = nops_;
  res.niter = it + 1;
}

inline void Irlm::eigenvectors(double* d_out) {
  std::vector<double> R((size_t)ncv_ * nev_);
  for (uint32_t i = 0; i < nev_; i++)
    for (uint32_t r = 0; r < ncv_; r++)
      R[(size_t)i * ncv_ + r] = ritz_vec_[(size_t)order_[i] * ncv_ + r];
  cudaMemcpyAsync(dQ_, R.data(), sizeof(double) * ncv_ * nev_, cudaMemcpyHostToDevice, stream_);
  tall_times_small(dQ_, nev_, d_out);
  cudaStreamSynchronize(stream_);
}

}  // namespace fpb
