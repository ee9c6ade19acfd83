// flashpcar.hpp -- the two PCA entry points of the flashpcaR package
// (flashpcaR/src/flashpca.cpp:17-93 flashpca_internal, :96-197
// flashpca_plink_internal) as plain C++ over the host mirror: identical
// parameter lists (the Eigen::Map argument becomes pointer + dimensions) and
// the same named result fields.  With -DRENV the Rcpp shims at the end of
// flashpcar.cpp export them under upstream's names, so R/flashpca.R works
// unchanged; R and Rcpp are not part of this image, the shims are not built here.
#pragma once
#include <string>
#include <vector>

#include "matrix.hpp"

namespace flashpca {

struct PcaResult {          // the Rcpp::List of flashpca.cpp:57-79 / 160-183
  Vector values;            // "values"     d
  Matrix vectors;           // "vectors"    U  (N x ndim)
  Matrix projection;        // "projection" Px (N x ndim)
  Matrix loadings;          // "loadings"   V  (p x ndim), only with do_loadings
  Vector center, scale;     // "center", "scale": empty unless return_scale && stand != 0
  Vector pve;               // "pve"
  bool has_loadings = false;
  std::vector<std::string> rownames;  // "FID:IID" (plink entry point only, flashpca.cpp:147-155)
};

PcaResult flashpca_internal(const double* X, size_t nrow, size_t ncol, int stand,
                            unsigned int ndim, unsigned int divisor, unsigned int maxiter,
                            double tol, long seed, bool verbose, bool do_loadings,
                            bool return_scale);

PcaResult flashpca_plink_internal(const std::string& fn, int stand, unsigned int ndim,
                                  unsigned int divisor, unsigned int maxiter,
                                  unsigned int block_size, double tol, long seed, bool verbose,
                                  bool do_loadings, bool return_scale);

}  // namespace flashpca
