// svdwide.hpp -- drop-in for upstream class SVDWideOnline (svdwide.h:32-107):
// the matrix-free operator that flashpca passes to Spectra::SymEigsSolver
// (randompca.cpp:173-175).  Same public surface -- rows(), cols(),
// perform_op(const double*, double*), the block variants and `trace` -- with
// every body forwarding to the C ABI in include/flashpca_b200.h.  The genotype
// matrix is staged into HBM once, in the constructor; block_size is accepted
// and ignored (there is no host-side N x block buffer any more).
//
// Errors surface as std::runtime_error, as upstream's Data I/O does
// (data.cpp:160,188,287).
#pragma once
#include <cstdlib>
#include <algorithm>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>

#include "data.hpp"
#include "flashpca_b200.h"
#include "matrix.hpp"

namespace flashpca {

// owns an fpb_handle from the moment fpb_create* returns, so that a constructor that throws later
// (C++ does not run the destructor of a partially constructed object) still releases the HBM
struct HandleDeleter {
  void operator()(fpb_handle* h) const { fpb_destroy(h); }
};
using HandlePtr = std::unique_ptr<fpb_handle, HandleDeleter>;

// Drop-in for upstream class SVDWide (svdwide.h:9-30): the operator of the
// in-memory path.  Upstream standardises the matrix on the host first
// (randompca.cpp:127 -> util.cpp:24-192) and keeps a reference to it; here the
// raw dosage matrix (NaN = missing) is copied to HBM and standardised there,
// and X_meansd / trace are available from the operator.
class SVDWide {
 public:
  SVDWide(const Matrix& raw, int stand_method, bool verbose_ = false, int device = 0)
      : n((unsigned int)raw.rows()), p((unsigned int)raw.cols()) {
    verbose = verbose_;
    nops = 1;
    fpb_handle* raw_h = nullptr;
    if (fpb_create_dense(&raw_h, raw.data(), raw.rows(), raw.cols(), stand_method, device))
      throw std::runtime_error(fpb_last_error(nullptr));
    owner.reset(raw_h);
    h = raw_h;
    if (fpb_get_trace(h, &trace)) throw std::runtime_error(fpb_last_error(h));
  }
#ifdef EIGEN_CORE_H
  // upstream's own signature (svdwide.h:18): a matrix that standardise() has already processed
  // (randompca.cpp:127); stand_method 0 = "none" leaves finite values untouched
  SVDWide(const Eigen::MatrixXd& mat_, bool verbose_ = false)
      : SVDWide(Matrix::from_column_major(mat_.data(), (size_t)mat_.rows(), (size_t)mat_.cols()), 0,
                verbose_) {}
#endif
  SVDWide(const SVDWide&) = delete;
  SVDWide& operator=(const SVDWide&) = delete;

  inline unsigned int rows() const { return n; }
  inline unsigned int cols() const { return n; }
  // y = mat * (mat' x)   (svdwide.cpp:4-12)
  void perform_op(const double* x_in, double* y_out) {
    if (fpb_perform_op(h, x_in, y_out)) throw std::runtime_error(fpb_last_error(h));
    nops++;
  }
  Matrix crossprod2(const Matrix& x) {  // mat' * x, for the loadings (randompca.cpp:151-152)
    Matrix Y(p, x.cols());
    if (fpb_crossprod_multi(h, x.data(), (uint32_t)x.cols(), Y.data()))
      throw std::runtime_error(fpb_last_error(h));
    return Y;
  }
  Matrix meansd() {
    Matrix m(p, 2);
    if (fpb_get_meansd(h, m.data())) throw std::runtime_error(fpb_last_error(h));
    return m;
  }
  double trace = 0;  // sum of squares of the standardised matrix (randompca.cpp:154)
  fpb_handle* handle() { return h; }

 private:
  const unsigned int n, p;
  bool verbose;
  unsigned int nops;
  HandlePtr owner;
  fpb_handle* h = nullptr;
};

class SVDWideOnline {
 public:
  double trace = 0;  // svdwide.h:37 (known right after staging)

  SVDWideOnline(Data& dat_, unsigned int block_size_, int stand_method_, bool verbose_,
                int device = 0)
      : dat(dat_), n(dat_.N), p(dat_.nsnps) {
    verbose = verbose_;
    block_size = block_size_;
    stand_method = stand_method_;
    nops = 1;
    const double* pre = dat.use_preloaded_maf ? dat.X_meansd.data() : nullptr;
    if (dat.use_preloaded_maf && dat.X_meansd.rows() != dat.nsnps)
      throw std::runtime_error("preloaded mean/sd table does not match the number of SNPs");
    // A bed that fits in HBM is staged once and stays there.  One that does not is kept in pinned
    // host memory and streamed slab by slab on every op -- the role upstream's --memory / block_size
    // plays (flashpca.cpp:649-676), chosen here from the device's free memory.  FPB_STREAM_SLAB_SNPS
    // forces the streaming mode with that many SNPs per slab.
    uint64_t slab = 0;
    if (const char* sv = getenv("FPB_STREAM_SLAB_SNPS")) slab = strtoull(sv, nullptr, 10);
    if (!slab) {
      uint64_t free_b = 0, total_b = 0;
      if (fpb_device_memory(device, &free_b, &total_b))
        throw std::runtime_error(fpb_last_error(nullptr));
      const uint64_t np = ((uint64_t)dat.N + 3) / 4, bed = np * dat.nsnps;
      if (bed + bed / 16 > free_b / 10 * 9) slab = std::max<uint64_t>(1, free_b / 4 / np);
    }
    // the standardisation method is Data's, as in data.cpp:279-288
    fpb_handle* raw_h = nullptr;
    const int rc = slab ? fpb_create_streaming(&raw_h, dat.geno_filename.c_str(), dat.N, 0, dat.nsnps,
                                               slab, dat.stand_method_x, pre, device)
                        : fpb_create_from_file(&raw_h, dat.geno_filename.c_str(), dat.N, 0, dat.nsnps,
                                               dat.stand_method_x, pre, device);
    if (rc) throw std::runtime_error(fpb_last_error(nullptr));
    owner.reset(raw_h);  // from here on an exception releases the handle and its HBM
    h = raw_h;
    if (slab && verbose)
      std::cout << "bed streamed from host memory, " << slab << " SNPs per slab" << std::endl;
    check(fpb_get_trace(h, &trace));
    if (!dat.use_preloaded_maf) {
      dat.X_meansd = Matrix(p, 2);
      check(fpb_get_meansd(h, dat.X_meansd.data()));
    }
  }
  SVDWideOnline(const SVDWideOnline&) = delete;
  SVDWideOnline& operator=(const SVDWideOnline&) = delete;

  inline unsigned int rows() const { return n; }
  inline unsigned int cols() const { return n; }

  // y = X X' x (svdwide.cpp:21-68)
  void perform_op(const double* x_in, double* y_out) {
    check(fpb_perform_op(h, x_in, y_out));
    nops++;
  }
  // Y = X X' M (svdwide.cpp:71-118, 229-275)
  Matrix perform_op_mat(const Matrix& x) {
    need_rows(x, n);
    Matrix Y(n, x.cols());
    check(fpb_perform_op_multi(h, x.data(), (uint32_t)x.cols(), Y.data()));
    nops++;
    return Y;
  }
  Matrix perform_op_multi(const Matrix& x) { return perform_op_mat(x); }
  // y = X' x (svdwide.cpp:122-153)
  void crossprod(double* x_in, double* y_out) {
    check(fpb_crossprod(h, x_in, y_out));
    nops++;
  }
  Matrix crossprod2(const Matrix& x) {  // svdwide.cpp:157-188
    need_rows(x, n);
    Matrix Y(p, x.cols());
    check(fpb_crossprod_multi(h, x.data(), (uint32_t)x.cols(), Y.data()));
    nops++;
    return Y;
  }
  // y = X x (svdwide.cpp:193-226)
  void prod(double* x_in, double* y_out) {
    check(fpb_prod(h, x_in, y_out));
    nops++;
  }
  Matrix prod3(const Matrix& x) {  // svdwide.cpp:312-343
    need_rows(x, p);
    Matrix Y(n, x.cols());
    check(fpb_prod_multi(h, x.data(), (uint32_t)x.cols(), Y.data()));
    nops++;
    return Y;
  }
  Matrix prod2(const Matrix& x) {  // svdwide.cpp:278-309: Y = x' X
    Matrix T = crossprod2(x);
    Matrix Y(x.cols(), p);
    for (size_t c = 0; c < x.cols(); c++)
      for (size_t j = 0; j < p; j++) Y(c, j) = T(j, c);
    return Y;
  }

#ifdef EIGEN_CORE_H
  // Upstream's exact signatures (svdwide.h:84-106) for a host that keeps Eigen: include
  // <Eigen/Core> before this header and the block variants take / return Eigen::MatrixXd.  Eigen's
  // default storage is column-major, the C ABI's convention, so the data pointers pass straight through.
  Eigen::MatrixXd perform_op_mat(const Eigen::MatrixXd& x) { return eig(&SVDWideOnline::call_op, x, n, n); }
  Eigen::MatrixXd perform_op_multi(const Eigen::MatrixXd& x) { return perform_op_mat(x); }
  Eigen::MatrixXd crossprod2(const Eigen::MatrixXd& x) { return eig(&SVDWideOnline::call_cp, x, n, p); }
  Eigen::MatrixXd prod3(const Eigen::MatrixXd& x) { return eig(&SVDWideOnline::call_pr, x, p, n); }
  Eigen::MatrixXd prod2(const Eigen::MatrixXd& x) { return crossprod2(x).transpose(); }
#endif

  fpb_handle* handle() { return h; }

 private:
#ifdef EIGEN_CORE_H
  int call_op(const double* in, uint32_t k, double* out) { return fpb_perform_op_multi(h, in, k, out); }
  int call_cp(const double* in, uint32_t k, double* out) { return fpb_crossprod_multi(h, in, k, out); }
  int call_pr(const double* in, uint32_t k, double* out) { return fpb_prod_multi(h, in, k, out); }
  Eigen::MatrixXd eig(int (SVDWideOnline::*fn)(const double*, uint32_t, double*), const Eigen::MatrixXd& x,
                      size_t rows_in, size_t rows_out) {
    if ((size_t)x.rows() != rows_in)
      throw std::runtime_error("operator argument has the wrong number of rows");
    Eigen::MatrixXd Y(rows_out, x.cols());
    check((this->*fn)(x.data(), (uint32_t)x.cols(), Y.data()));
    nops++;
    return Y;
  }
#endif
  void check(int rc) {
    if (rc) throw std::runtime_error(fpb_last_error(h));
  }
  static void need_rows(const Matrix& m, size_t r) {
    if (m.rows() != r) throw std::runtime_error("operator argument has the wrong number of rows");
  }
  Data& dat;
  const unsigned int n, p;
  int stand_method;
  bool verbose;
  unsigned int nops;
  unsigned int block_size;
  HandlePtr owner;
  fpb_handle* h = nullptr;
};

}  // namespace flashpca
