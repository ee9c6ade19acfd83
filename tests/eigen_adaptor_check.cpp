// Drives the Eigen-typed surface of flashpca_b200/host/svdwide.hpp (svdwide.h:18, 84-106) with the
// stub of tests/eigen_stub: SVDWideOnline::perform_op_mat / crossprod2 / prod3 / prod2 on
// Eigen::MatrixXd must equal the Matrix-typed calls.  Usage: eigen_adaptor_check <bfile stem>
#include <Eigen/Core>

#include <cmath>
#include <cstdio>

#include "data.hpp"
#include "svdwide.hpp"

using namespace flashpca;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    Data d;
    d.verbose = false;
    d.stand_method_x = 3;
    const std::string stem = argv[1];
    d.read_pheno((stem + ".fam").c_str(), 6);
    d.read_plink_bim((stem + ".bim").c_str());
    d.read_plink_fam((stem + ".fam").c_str());
    d.geno_filename = stem + ".bed";
    d.get_size();
    d.prepare();
    SVDWideOnline op(d, 100, 3, false);
    const size_t n = op.rows(), p = d.nsnps, k = 3;
    Eigen::MatrixXd xe(n, k), ve(p, k);
    Matrix xm(n, k), vm(p, k);
    for (size_t c = 0; c < k; c++) {
      for (size_t i = 0; i < n; i++) xe(i, c) = xm(i, c) = std::sin(0.37 * (double)(i + 1) * (double)(c + 1));
      for (size_t j = 0; j < p; j++) ve(j, c) = vm(j, c) = std::cos(0.11 * (double)(j + 1) * (double)(c + 2));
    }
    double worst = 0.0;
    auto cmp = [&](const Eigen::MatrixXd& a, const Matrix& b) {
      if ((size_t)a.rows() != b.rows() || (size_t)a.cols() != b.cols()) worst = 1e300;
      else
        for (size_t i = 0; i < b.size(); i++) worst = std::fmax(worst, std::fabs(a.data()[i] - b.data()[i]));
    };
    cmp(op.perform_op_mat(xe), op.perform_op_mat(xm));
    cmp(op.perform_op_multi(xe), op.perform_op_multi(xm));
    cmp(op.crossprod2(xe), op.crossprod2(xm));
    cmp(op.prod3(ve), op.prod3(vm));
    cmp(op.prod2(xe), op.prod2(xm));
    std::printf("eigen adaptor max abs difference: %g\n", worst);
    return worst == 0.0 ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
