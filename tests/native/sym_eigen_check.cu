// Host-only check of fpb::sym_eigen (flashpca_b200/csrc/fpb_block.cuh): the projected eigenproblem of
// the block Krylov solver.  Compiled with nvcc (the header carries kernels) and run on the CPU.
#include <cstdio>
#include <cstdlib>

#include "fpb_block.cuh"

int main() {
  int bad = 0;
  for (int n : {1, 2, 3, 8, 37, 96, 208}) {
    std::vector<double> a((size_t)n * n);
    srand(n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) {
        const double v = (double)rand() / RAND_MAX - 0.5 + (i == j ? 3.0 * (i % 5) : 0.0);
        a[(size_t)j * n + i] = a[(size_t)i * n + j] = v;
      }
    std::vector<double> w, z;
    const bool ok = fpb::sym_eigen(n, a, w, z);
    double res = 0, orth = 0;
    for (int e = 0; e < n; e++) {
      for (int i = 0; i < n; i++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += a[(size_t)k * n + i] * z[(size_t)e * n + k];
        res = fmax(res, fabs(s - w[e] * z[(size_t)e * n + i]));
      }
      for (int f = 0; f < n; f++) {
        double s = 0;
        for (int k = 0; k < n; k++) s += z[(size_t)e * n + k] * z[(size_t)f * n + k];
        orth = fmax(orth, fabs(s - (e == f ? 1.0 : 0.0)));
      }
    }
    printf("n=%d ok=%d residual %.2e orthogonality %.2e\n", n, (int)ok, res, orth);
    if (!ok || res > 1e-12 || orth > 1e-12) bad++;
  }
  return bad;
}
