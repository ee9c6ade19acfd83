#include "randompca.hpp"

#include <cstdlib>

#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>

#include "svdwide.hpp"
#include "util.hpp"

namespace flashpca {

static double divisor_value(int divisor, double n, double p) {
  if (divisor == DIVISOR_N1) return n - 1;
  if (divisor == DIVISOR_P) return p;
  return 1;
}

// shared tail of both pca_fast overloads (randompca.cpp:138-157, 180-210)
template <class Op>
static void finish_pca(RandomPCA& r, Op& op, unsigned int N, unsigned int p, unsigned int ndim,
                       unsigned int maxiter, double tol, bool do_loadings) {
  r.U = Matrix(N, ndim);
  Vector evals(ndim);
  uint32_t nconv = 0, nops_ = 0, niter = 0;
  // FPB_SOLVER=block selects the block Krylov solver (an extension: 8 columns per pass over the
  // matrix on the tcgen05 kernels, Spectra's convergence criterion); default = Spectra's schedule
  const char* sv = getenv("FPB_SOLVER");
  if (sv && std::string(sv) == "block" && (uint64_t)ndim + 8 <= N) {
    uint32_t npasses = 0;
    if (fpb_pca_block(op.handle(), ndim, 8, 0, tol, evals.data(), r.U.data(), &nconv, &npasses))
      throw std::runtime_error(fpb_last_error(op.handle()));
    nops_ = 8 * npasses;
    r.verbose&& std::cout << timestamp() << "Block Krylov solver: " << npasses
                          << " passes of 8 columns" << std::endl;
  } else {
    if (fpb_pca(op.handle(), ndim, ndim * 2 + 1, maxiter, tol, evals.data(), r.U.data(), &nconv,
                &nops_, &niter))
      throw std::runtime_error(fpb_last_error(op.handle()));
    r.verbose&& std::cout << timestamp() << "Matrix operations: " << nops_
                          << ", restarts: " << (niter - 1) << std::endl;
  }
  r.nops = nops_;
  if (nconv < ndim)
    // upstream throws a *pointer* here (randompca.cpp:160-165, 212-217) and main's catch(...)
    // reports an unknown exception; a value is thrown instead so the message survives.
    throw std::runtime_error(std::string("Spectra eigen-decomposition was not successful") +
                             ", status: 1 (converged " + std::to_string(nconv) + " of " +
                             std::to_string(ndim) + ")");
  double div = divisor_value(r.divisor, N, p);
  r.d.resize(ndim);
  for (unsigned int j = 0; j < ndim; j++) r.d[j] = evals[j] / div;  // eigenvalues, not singular values
  if (do_loadings) {
    r.verbose&& std::cout << "Computing loadings" << std::endl;
    r.V = op.crossprod2(r.U);
    for (unsigned int j = 0; j < ndim; j++) {
      double s = r.d[j];
      for (unsigned int i = 0; i < p; i++) r.V(i, j) = r.V(i, j) * (1.0 / sqrt(s)) / sqrt(div);
    }
  }
  r.trace = op.trace / div;
  r.pve.resize(ndim);
  r.Px = Matrix(N, ndim);
  for (unsigned int j = 0; j < ndim; j++) {
    r.pve[j] = r.d[j] / r.trace;
    double sd = sqrt(r.d[j]);
    for (unsigned int i = 0; i < N; i++) r.Px(i, j) = r.U(i, j) * sd;
  }
  r.verbose&& std::cout << timestamp() << "GRM trace: " << r.trace << std::endl;
}

void RandomPCA::pca_fast(Matrix& X, unsigned int block_size, unsigned int ndim,
                         unsigned int maxiter, double tol, long seed, bool do_loadings) {
  (void)block_size;
  (void)seed;
  // X_meansd = standardise(X, stand_method_x); SVDWide op(X)   (randompca.cpp:127-130)
  SVDWide op(X, stand_method_x, verbose, device);
  X_meansd = op.meansd();
  finish_pca(*this, op, (unsigned int)X.rows(), (unsigned int)X.cols(), ndim, maxiter, tol,
             do_loadings);
}

void RandomPCA::pca_fast(Data& dat, unsigned int block_size, unsigned int ndim,
                         unsigned int maxiter, double tol, long seed, bool do_loadings) {
  (void)seed;  // upstream never uses it for PCA either (randompca.cpp:168-178)
  unsigned int N = dat.N, p = dat.nsnps;
  SVDWideOnline op(dat, block_size, stand_method_x, verbose, device);

  // Spectra::SymEigsSolver<double, LARGEST_ALGE, SVDWideOnline> eigs(&op, ndim, 2*ndim+1);
  // eigs.init(); eigs.compute(maxiter, tol);   -- run with the basis resident in HBM
  finish_pca(*this, op, N, p, ndim, maxiter, tol, do_loadings);
  X_meansd = dat.X_meansd;
}

void RandomPCA::check(Data& dat, unsigned int block_size, std::string evec_file,
                      std::string eval_file) {
  verbose&& std::cout << timestamp() << "Loading eigenvalue file '" << eval_file << "'" << std::endl;
  Matrix ev = read_text(eval_file.c_str(), 1, 0);
  if (ev.rows() == 0) throw std::runtime_error("No eigenvalues found in file");
  Vector eval(ev.rows());
  for (size_t i = 0; i < ev.rows(); i++) eval[i] = ev(i, 0);

  verbose&& std::cout << timestamp() << "Loading eigenvector file '" << evec_file << "'" << std::endl;
  Matrix evec = read_text(evec_file.c_str(), 3, 1);
  if (evec.rows() != dat.N)
    throw std::runtime_error(std::string("Eigenvector dimension doesn't match data dimension") +
                             " (evec.rows = " + std::to_string(evec.rows()) +
                             "; dat.N = " + std::to_string(dat.N) + ")");
  if (eval.size() != evec.cols())
    throw std::runtime_error("Eigenvector dimension doesn't match the number of eigenvalues");
  check(dat, block_size, evec, eval);
}

void RandomPCA::check(Data& dat, unsigned int block_size, Matrix& evec, Vector& eval) {
  SVDWideOnline op(dat, block_size, 1, verbose, device);
  unsigned int K = (unsigned int)std::min(evec.cols(), eval.size());
  verbose&& std::cout << timestamp()
                      << "Checking mean square error between (X X' U) / div and (U D^2)"
                      << " for " << K << " dimensions" << std::endl;
  double div = divisor_value(divisor, dat.N, dat.nsnps);
  Matrix XXU = op.perform_op_mat(evec);
  err.assign(evec.cols(), 0.0);
  for (size_t j = 0; j < evec.cols(); j++) {
    double s = 0;
    for (size_t i = 0; i < evec.rows(); i++) {
      double e = XXU(i, j) / div - evec(i, j) * eval[j];
      s += e * e;
    }
    err[j] = s;
  }
  double tot = 0;
  for (unsigned int j = 0; j < K; j++) {
    verbose&& std::cout << timestamp() << "eval(" << (j + 1) << "): " << eval[j]
                        << ", sum squared error: " << err[j] << std::endl;
  }
  for (double e : err) tot += e;
  mse = tot / ((double)dat.N * K);
  rmse = std::sqrt(mse);
  verbose&& std::cout << timestamp() << "Mean squared error: " << mse
                      << ", Root mean squared error: " << rmse << " (n=" << dat.N << ")"
                      << std::endl;
}

Matrix maf2meansd(const Matrix& maf) {
  Matrix ms(maf.rows(), 2);
  for (size_t i = 0; i < maf.rows(); i++) {
    ms(i, 0) = maf(i, 0) * 2.0;
    ms(i, 1) = maf(i, 0) * 2.0 * (1.0 - maf(i, 0));  // no sqrt, as upstream
  }
  return ms;
}

void RandomPCA::project(Data& dat, unsigned int block_size, std::string loadings_file,
                        std::string maf_file, std::string meansd_file) {
  V = read_text(loadings_file.c_str(), 3, 1);
  if (maf_file != "") {
    verbose&& std::cout << timestamp() << "Reading MAF file " << maf_file << std::endl;
    Matrix maf = read_MAF(maf_file.c_str(), dat.snp_ids, verbose);
    dat.X_meansd = maf2meansd(maf);
    dat.use_preloaded_maf = true;
  } else if (meansd_file != "") {
    verbose&& std::cout << timestamp() << " Reading mean/stdev file " << meansd_file << std::endl;
    dat.X_meansd = read_text(meansd_file.c_str(), 3, 1);
    dat.use_preloaded_maf = true;
  } else {
    verbose&& std::cout << timestamp() << " Using MAF from the data" << std::endl;
    dat.use_preloaded_maf = false;
  }
  project(dat, block_size);
}

void RandomPCA::project(Data& dat, unsigned int block_size) {
  if (V.rows() != dat.nsnps)
    throw std::runtime_error("The number of SNPs in the loadings doesn't match the bed file (" +
                             std::to_string(V.rows()) + " vs " + std::to_string(dat.nsnps) + ")");
  SVDWideOnline op(dat, block_size, 1, verbose, device);
  unsigned int k = (unsigned int)V.cols();
  double div = 1;
  if (divisor == DIVISOR_N1) div = dat.N - 1;
  else if (divisor == DIVISOR_P) div = (double)V.rows();
  Px = op.prod3(V);  // the k single-vector prod calls of randompca.cpp:813-819, batched
  double s = sqrt(div);
  for (unsigned int j = 0; j < k; j++)
    for (unsigned int i = 0; i < dat.N; i++) Px(i, j) = Px(i, j) / s;  // X V = U D
}

}  // namespace flashpca
