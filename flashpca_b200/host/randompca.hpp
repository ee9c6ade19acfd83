// randompca.hpp -- host mirror of upstream class RandomPCA for the PCA path
// (randompca.h:54-108): pca_fast over a Data (online mode), check, project.
#pragma once
#include <string>

#include "data.hpp"
#include "matrix.hpp"

#define DIVISOR_NONE 0
#define DIVISOR_N1 1
#define DIVISOR_P 2

namespace flashpca {

class RandomPCA {
 public:
  Matrix U, V, Px;
  Vector d, pve, err;
  double trace = 0, mse = 0, rmse = 0;
  Matrix X_meansd;
  int stand_method_x = 3, stand_method_y = 1;
  bool verbose = false, debug = false;
  int divisor = DIVISOR_P;
  int device = 0;
  unsigned int nops = 0;

  // randompca.cpp:121-166 (in-memory matrix of dosages, NaN = missing)
  void pca_fast(Matrix& X, unsigned int block_size, unsigned int ndim, unsigned int maxiter,
                double tol, long seed, bool do_loadings);
  // randompca.cpp:168-218
  void pca_fast(Data& dat, unsigned int block_size, unsigned int ndim, unsigned int maxiter,
                double tol, long seed, bool do_loadings);
  // randompca.cpp:627-703
  void check(Data& dat, unsigned int block_size, std::string evec_file, std::string eval_file);
  void check(Data& dat, unsigned int block_size, Matrix& evec, Vector& eval);
  // randompca.cpp:753-820
  void project(Data& dat, unsigned int block_size, std::string loadings_file,
               std::string maf_file, std::string meansd_file);
  void project(Data& dat, unsigned int block_size);
};

Matrix maf2meansd(const Matrix& maf);  // randompca.cpp:745-751

}  // namespace flashpca
