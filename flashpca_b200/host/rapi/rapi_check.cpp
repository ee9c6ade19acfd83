// rapi_check -- exercises the two flashpcaR entry points (flashpcar.hpp) the way
// R/flashpca.R does and prints the named result fields as JSON for the tests:
//   rapi_check plink  <bfile stem> <ndim> <stand> <divisor> <do_loadings> <return_scale>
//   rapi_check matrix <bfile stem> <ndim> <stand> <divisor> <do_loadings> <return_scale>
// ("matrix" loads the dosages of the bed on the host, missing -> NaN, and passes
// the numeric matrix like flashpca(X = <matrix>) does.)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>

#include "../data.hpp"
#include "../flashpcar.hpp"

using namespace flashpca;

static void put_vec(const char* name, const Vector& v, bool last = false) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); i++) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]%s\n", last ? "" : ",");
}
static void put_mat(const char* name, const Matrix& m) {
  printf("\"%s\": {\"nrow\": %zu, \"ncol\": %zu, \"data\": [", name, m.rows(), m.cols());
  for (size_t i = 0; i < m.size(); i++) printf("%s%.17g", i ? ", " : "", m.data()[i]);
  printf("]},\n");
}

int main(int argc, char** argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: rapi_check plink|matrix stem ndim stand divisor do_loadings return_scale\n");
    return 2;
  }
  const std::string mode = argv[1], stem = argv[2];
  const unsigned ndim = atoi(argv[3]);
  const int stand = atoi(argv[4]);
  const unsigned divisor = atoi(argv[5]);
  const bool do_loadings = atoi(argv[6]) != 0, return_scale = atoi(argv[7]) != 0;
  try {
    PcaResult r;
    if (mode == "plink") {
      r = flashpca_plink_internal(stem, stand, ndim, divisor, 500, 0, 1e-8, 1, false, do_loadings,
                                  return_scale);
    } else {
      Data d;
      d.read_pheno((stem + ".fam").c_str(), 6);
      d.geno_filename = stem + ".bed";
      d.get_size();
      // dosage matrix with NaN for missing (what an R user passes)
      std::ifstream in(d.geno_filename, std::ios::binary);
      in.seekg(3);
      Matrix X(d.N, d.nsnps);
      std::vector<unsigned char> tmp(d.np);
      for (unsigned j = 0; j < d.nsnps; j++) {
        in.read((char*)tmp.data(), d.np);
        for (unsigned i = 0; i < d.N; i++) {
          unsigned char g = (tmp[i >> 2] >> (2 * (i & 3))) & 3;
          X(i, j) = g == 1 ? std::numeric_limits<double>::quiet_NaN()
                           : (double)(!(g & 1) + !(g >> 1));
        }
      }
      r = flashpca_internal(X.data(), X.rows(), X.cols(), stand, ndim, divisor, 500, 1e-8, 1,
                            false, do_loadings, return_scale);
    }
    printf("{\n");
    put_vec("values", r.values);
    put_mat("vectors", r.vectors);
    put_mat("projection", r.projection);
    if (r.has_loadings) put_mat("loadings", r.loadings);
    put_vec("center", r.center);
    put_vec("scale", r.scale);
    printf("\"rownames\": [");
    for (size_t i = 0; i < r.rownames.size(); i++) printf("%s\"%s\"", i ? ", " : "", r.rownames[i].c_str());
    printf("],\n");
    put_vec("pve", r.pve, true);
    printf("}\n");
  } catch (std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
