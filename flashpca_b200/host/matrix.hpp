// matrix.hpp -- minimal column-major double matrix (stands in for the
// Eigen::MatrixXd / VectorXd members of upstream's Data and RandomPCA; Eigen is
// not a dependency of this host layer).
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

namespace flashpca {

struct Matrix {
  size_t nrow = 0, ncol = 0;
  std::vector<double> v;
  Matrix() = default;
  Matrix(size_t r, size_t c, double init = 0.0) : nrow(r), ncol(c), v(r * c, init) {}
  static Matrix from_column_major(const double* src, size_t r, size_t c) {
    Matrix m(r, c);
    std::copy(src, src + r * c, m.v.begin());
    return m;
  }
  double& operator()(size_t r, size_t c) { return v[c * nrow + r]; }
  double operator()(size_t r, size_t c) const { return v[c * nrow + r]; }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double* col(size_t c) { return v.data() + c * nrow; }
  const double* col(size_t c) const { return v.data() + c * nrow; }
  size_t rows() const { return nrow; }
  size_t cols() const { return ncol; }
  size_t size() const { return v.size(); }
};

using Vector = std::vector<double>;

}  // namespace flashpca
