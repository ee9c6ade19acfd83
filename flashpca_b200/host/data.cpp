#include "data.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "util.hpp"

namespace flashpca {

static std::vector<std::string> read_lines(std::ifstream& in, unsigned int skip = 0) {
  std::vector<std::string> lines;
  unsigned int line_num = 0;
  while (in) {
    std::string line;
    std::getline(in, line);
    if (!in.eof()) {  // a last line without '\n' is dropped, as upstream does
      if (line_num >= skip) lines.push_back(line);
      line_num++;
    }
  }
  return lines;
}

static std::vector<std::string> split_ws(const std::string& line) {
  std::stringstream ss(line);
  std::string s;
  std::vector<std::string> tokens;
  while (ss >> s) tokens.push_back(s);
  return tokens;
}

void Data::read_pheno(const char* filename, unsigned int firstcol) {
  Matrix m = read_text(filename, firstcol, 0);
  N = (unsigned int)m.rows();
}

void Data::read_plink_bim(const char* filename) {
  std::ifstream in(filename, std::ios::in);
  if (!in) throw std::runtime_error(std::string("Error reading file ") + filename);
  std::vector<std::string> lines = read_lines(in);
  if (verbose)
    std::cout << timestamp() << "Detected bim file " << filename << ", " << lines.size()
              << " SNPs" << std::endl;
  for (size_t i = 0; i < lines.size(); i++) {
    std::vector<std::string> tokens = split_ws(lines[i]);
    if (tokens.size() < 6)
      throw std::runtime_error(std::string("Error reading file '") + filename + "', line " +
                               std::to_string(i + 1) + ": expected 6 columns");
    snp_ids.push_back(tokens[1]);
    ref_alleles.push_back(tokens[4]);
    alt_alleles.push_back(tokens[5]);
    char* end;
    errno = 0;
    unsigned long long m = std::strtol(tokens[3].c_str(), &end, 10);
    if (*end != '\0' || errno != 0)
      throw std::runtime_error(std::string("Error reading file '") + filename + "', line " +
                               std::to_string(i + 1) + ": '" + tokens[3] +
                               "' cannot be parsed as a number");
    bp.push_back(m);
  }
}

void Data::read_plink_fam(const char* filename) {
  std::ifstream in(filename, std::ios::in);
  if (!in)
    throw std::runtime_error(std::string("[Data::read_plink_fam] Error reading file ") + filename);
  std::vector<std::string> lines = read_lines(in);
  for (size_t i = 0; i < lines.size(); i++) {
    std::vector<std::string> tokens = split_ws(lines[i]);
    if (tokens.size() < 2)
      throw std::runtime_error(std::string("[Data::read_plink_fam] Error reading file ") +
                               filename + ": line " + std::to_string(i + 1));
    fam_ids.push_back(tokens[0]);
    indiv_ids.push_back(tokens[1]);
  }
}

void Data::get_size() {
  if (verbose) std::cout << timestamp() << "Analyzing BED file '" << geno_filename << "'";
  std::ifstream in(geno_filename, std::ios::in | std::ios::binary);
  if (!in)
    throw std::runtime_error(std::string("[Data::read_bed] Error reading file ") + geno_filename +
                             ", error " + strerror(errno));
  in.seekg(0, std::ifstream::end);
  len = (unsigned long long)in.tellg() - 3;  // no magic-byte validation, as upstream
  np = ((unsigned long long)N + 3) / 4;
  nsnps = np ? (unsigned int)(len / np) : 0;
  if (verbose) std::cout << ", found " << (len + 3) << " bytes, " << nsnps << " SNPs" << std::endl;
}

void Data::prepare() {
  std::ifstream in(geno_filename, std::ios::in | std::ios::binary);
  if (!in) throw std::runtime_error(std::string("[Data::read_bed] Error reading file ") + geno_filename);
  if (!use_preloaded_maf) X_meansd = Matrix(nsnps, 2, 0.0);
  if (verbose)
    std::cout << timestamp() << "Detected BED file: " << geno_filename << " with " << (len + 3)
              << " bytes, " << N << " samples, " << nsnps << " SNPs." << std::endl;
}

// data.cpp:339-406: the whole bed as doubles (minor-allele dosage, decode_plink
// data.cpp:65-126), missing genotypes imputed to the per-SNP average of the
// non-missing ones.  Only the non-transposed form is used by the PCA path.
void Data::read_bed(bool transpose) {
  if (transpose) throw std::runtime_error("read_bed(transpose=true) is not used by the PCA path");
  std::ifstream in(geno_filename, std::ios::in | std::ios::binary);
  if (!in) throw std::runtime_error(std::string("[Data::read_bed] Error reading file ") + geno_filename);
  in.seekg(3, std::ifstream::beg);
  X = Matrix(N, nsnps);
  std::vector<unsigned char> tmp(np);
  for (unsigned int j = 0; j < nsnps; j++) {
    in.read((char*)tmp.data(), np);
    double avg = 0;
    unsigned int ngood = 0;
    double* col = X.col(j);
    for (unsigned int i = 0; i < N; i++) {
      unsigned char g = (tmp[i >> 2] >> (2 * (i & 3))) & 3;
      // 00 -> 2, 10 -> 1, 11 -> 0, 01 -> missing
      if (g == 1) {
        col[i] = -1.0;
      } else {
        double s = (double)(!(g & 1) + !(g >> 1));
        col[i] = s;
        avg += s;
        ngood++;
      }
    }
    avg /= ngood;
    for (unsigned int i = 0; i < N; i++)
      if (col[i] < 0) col[i] = avg;
  }
  if (verbose)
    std::cout << timestamp() << "Loaded genotypes: " << N << " samples, " << nsnps << " SNPs"
              << std::endl;
}

Matrix read_MAF(const char* filename, const std::vector<std::string>& snp_ids, bool verbose) {
  std::ifstream in(filename, std::ios::in);
  if (!in)
    throw std::runtime_error(std::string("Error reading file '") + filename +
                             "': " + strerror(errno));
  std::vector<std::string> lines = read_lines(in, 1);  // skip the .frq header
  if (verbose)
    std::cout << timestamp() << "Detected text file " << filename << ", " << lines.size()
              << " rows" << std::endl;
  if (lines.size() != snp_ids.size())
    throw std::runtime_error(std::string("Error number of SNPs in '") + filename +
                             "': different number of SNPs than in the bim file'");
  Matrix m(lines.size(), 1);
  for (size_t i = 0; i < lines.size(); i++) {
    std::vector<std::string> tokens = split_ws(lines[i]);
    if (tokens.size() != 6)
      throw std::runtime_error(std::string("Error reading file '") + filename +
                               "': inconsistent number of columns");
    if (tokens[1] != snp_ids[i])
      throw std::runtime_error(std::string("Error reading file '") + filename +
                               "': inconsistent SNP id at row':" + std::to_string(i));
    char* end;
    errno = 0;
    double val = std::strtod(tokens[4].c_str(), &end);
    if (*end != '\0' || errno != 0)
      throw std::runtime_error(std::string("Error reading file '") + filename + "', line " +
                               std::to_string(i + 1) + ": '" + tokens[4] +
                               "' cannot be parsed as a number");
    m(i, 0) = val;
  }
  return m;
}

}  // namespace flashpca
