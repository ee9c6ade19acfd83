#include "flashpcar.hpp"

#include "data.hpp"
#include "randompca.hpp"

namespace flashpca {

static void fill_result(PcaResult& res, RandomPCA& rpca, int stand, bool do_loadings,
                        bool return_scale) {
  res.vectors = rpca.U;
  res.projection = rpca.Px;
  res.values = rpca.d;
  res.pve = rpca.pve;
  if (return_scale && stand != 0) {  // STANDARDISE_NONE: 0 (flashpca.cpp:46-52, 158-164)
    const size_t p = rpca.X_meansd.rows();
    res.center.assign(rpca.X_meansd.col(0), rpca.X_meansd.col(0) + p);
    res.scale.assign(rpca.X_meansd.col(1), rpca.X_meansd.col(1) + p);
  }
  res.has_loadings = do_loadings;
  if (do_loadings) res.loadings = rpca.V;
}

// flashpcaR/src/flashpca.cpp:17-93
PcaResult flashpca_internal(const double* X, size_t nrow, size_t ncol, int stand,
                            unsigned int ndim, unsigned int divisor, unsigned int maxiter,
                            double tol, long seed, bool verbose, bool do_loadings,
                            bool return_scale) {
  Matrix Xm(nrow, ncol);  // Eigen::MatrixXd Xm(X): the caller's matrix is copied, not modified
  std::copy(X, X + nrow * ncol, Xm.data());
  RandomPCA rpca;
  rpca.stand_method_x = stand;
  rpca.divisor = (int)divisor;
  rpca.verbose = verbose;
  rpca.pca_fast(Xm, 0, ndim, maxiter, tol, seed, do_loadings);
  PcaResult res;
  fill_result(res, rpca, stand, do_loadings, return_scale);
  return res;
}

// flashpcaR/src/flashpca.cpp:96-197
PcaResult flashpca_plink_internal(const std::string& fn, int stand, unsigned int ndim,
                                  unsigned int divisor, unsigned int maxiter,
                                  unsigned int block_size, double tol, long seed, bool verbose,
                                  bool do_loadings, bool return_scale) {
  RandomPCA rpca;
  rpca.stand_method_x = stand;
  rpca.divisor = (int)divisor;
  rpca.verbose = verbose;
  const std::string geno_file = fn + ".bed", bim_file = fn + ".bim", fam_file = fn + ".fam";
  Data data;
  data.verbose = verbose;
  data.stand_method_x = stand;
  data.read_pheno(fam_file.c_str(), 6);
  data.read_plink_fam(fam_file.c_str());
  data.read_plink_bim(bim_file.c_str());
  data.geno_filename = geno_file;
  data.get_size();
  data.prepare();
  rpca.pca_fast(data, block_size, ndim, maxiter, tol, seed, do_loadings);
  PcaResult res;
  fill_result(res, rpca, stand, do_loadings, return_scale);
  res.rownames.resize(data.fam_ids.size());
  for (size_t i = 0; i < data.fam_ids.size(); i++)
    res.rownames[i] = data.fam_ids[i] + ":" + data.indiv_ids[i];
  return res;
}

}  // namespace flashpca

#ifdef RENV
// Rcpp shims under upstream's exported names (R/RcppExports.R:4-10 calls them).
#include <Rcpp.h>

static Rcpp::NumericMatrix to_r(const flashpca::Matrix& m) {
  Rcpp::NumericMatrix out((int)m.rows(), (int)m.cols());
  std::copy(m.data(), m.data() + m.size(), out.begin());
  return out;
}
static Rcpp::List to_list(const flashpca::PcaResult& r) {
  Rcpp::NumericMatrix U = to_r(r.vectors), P = to_r(r.projection);
  if (!r.rownames.empty()) {
    Rcpp::CharacterVector rn(r.rownames.begin(), r.rownames.end());
    Rcpp::rownames(U) = rn;
    Rcpp::rownames(P) = rn;
  }
  Rcpp::NumericVector d(r.values.begin(), r.values.end()), c(r.center.begin(), r.center.end()),
      s(r.scale.begin(), r.scale.end()), pve(r.pve.begin(), r.pve.end());
  if (r.has_loadings)
    return Rcpp::List::create(Rcpp::Named("values") = d, Rcpp::Named("vectors") = U,
                              Rcpp::Named("projection") = P,
                              Rcpp::Named("loadings") = to_r(r.loadings),
                              Rcpp::Named("center") = c, Rcpp::Named("scale") = s,
                              Rcpp::Named("pve") = pve);
  return Rcpp::List::create(Rcpp::Named("values") = d, Rcpp::Named("vectors") = U,
                            Rcpp::Named("projection") = P, Rcpp::Named("center") = c,
                            Rcpp::Named("scale") = s, Rcpp::Named("pve") = pve);
}

// [[Rcpp::export]]
Rcpp::List flashpca_internal(const Rcpp::NumericMatrix X, const int stand, const unsigned int ndim,
                             const unsigned int divisor, const unsigned int maxiter,
                             const double tol, const long seed, const bool verbose,
                             const bool do_loadings, const bool return_scale) {
  try {
    return to_list(flashpca::flashpca_internal(&X[0], X.nrow(), X.ncol(), stand, ndim, divisor,
                                               maxiter, tol, seed, verbose, do_loadings,
                                               return_scale));
  } catch (std::exception& ex) {
    forward_exception_to_r(ex);
  } catch (...) {
    ::Rf_error("flashpca_internal: unknown c++ exception");
  }
  return Rcpp::List();
}

// [[Rcpp::export]]
Rcpp::List flashpca_plink_internal(const std::string fn, const int stand, const unsigned int ndim,
                                   const unsigned int divisor, const unsigned int maxiter,
                                   const unsigned int block_size, const double tol,
                                   const long seed, const bool verbose, const bool do_loadings,
                                   const bool return_scale) {
  try {
    return to_list(flashpca::flashpca_plink_internal(fn, stand, ndim, divisor, maxiter, block_size,
                                                     tol, seed, verbose, do_loadings,
                                                     return_scale));
  } catch (std::exception& ex) {
    forward_exception_to_r(ex);
  } catch (...) {
    ::Rf_error("flashpca_plink_internal: unknown c++ exception");
  }
  return Rcpp::List();
}
#endif
