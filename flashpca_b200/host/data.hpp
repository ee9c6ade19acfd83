// data.hpp -- host mirror of upstream class Data (data.h:81-100) for the PCA
// path.  Text parsing (fam / bim / MAF) follows data.cpp:408-672; the genotypes
// themselves are never decoded on the host: the bed is staged into HBM by the
// native library (fpb_create_from_file) when an operator is built on this Data.
#pragma once
#include <string>
#include <vector>

#include "matrix.hpp"

namespace flashpca {

class Data {
 public:
  Matrix X;         // N x nsnps dosages, only filled by read_bed (--batch)
  Matrix X_meansd;  // nsnps x 2 (mean, sd): data.cpp:198,290-291
  unsigned int N = 0, nsnps = 0;
  unsigned long long len = 0, np = 0;
  std::string geno_filename;
  int stand_method_x = 3;
  bool verbose = false;
  bool use_preloaded_maf = false;
  std::vector<std::string> snp_ids, ref_alleles, alt_alleles, fam_ids, indiv_ids;
  std::vector<unsigned long long> bp;

  void read_pheno(const char* filename, unsigned int firstcol);  // data.cpp:408-413
  void read_plink_bim(const char* filename);                    // data.cpp:589-637
  void read_plink_fam(const char* filename);                    // data.cpp:639-672
  void get_size();                                              // data.cpp:150-176
  void prepare();                                               // data.cpp:179-206
  void read_bed(bool transpose);                                // data.cpp:339-406
};

// data.cpp:419-496: plink .frq (CHR SNP A1 A2 MAF NCHROBS), SNP ids must match the bim.
Matrix read_MAF(const char* filename, const std::vector<std::string>& snp_ids, bool verbose);

}  // namespace flashpca
