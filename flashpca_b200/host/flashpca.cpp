// flashpca.cpp -- the flashpca command line for the PCA path (PCA, --check,
// --project) on the B200 library.  Option names, defaults, validation order,
// messages and output files follow upstream flashpca.cpp:40-892; Boost
// program_options is replaced by a small parser with the same surface
// (--opt value, --opt=value, and the short forms -p -m -b -n -d -s -v -f -c).
// --batch (all genotypes as doubles, the in-memory matrix path) is supported;
// SCCA / UCCA are outside this build's scope and are rejected with a clear message.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "data.hpp"
#include "randompca.hpp"
#include "util.hpp"

#ifndef VERSION
#define VERSION "2.1-b200"
#endif

#define MODE_PCA 1
#define MODE_CHECK_PCA 5
#define MODE_PREDICT_PCA 6

using namespace flashpca;

namespace {

struct OptSpec {
  const char* name;
  char shortname;
  bool has_value;
  const char* help;
};

const OptSpec kOptions[] = {
    {"help", 0, false, "produce help message"},
    {"scca", 0, false, "perform sparse canonical correlation analysis (SCCA) [not in this build]"},
    {"ucca", 0, false, "perform per-SNP canonical correlation analysis [not in this build]"},
    {"project", 'p', false, "project new samples onto existing principal components"},
    {"batch", 0, false, "load all genotypes into RAM at once"},
    {"memory", 'm', true, "size of block, in MB"},
    {"blocksize", 'b', true, "size of block for, in number of SNPs"},
    {"numthreads", 'n', true, "set number of OpenMP threads"},
    {"seed", 0, true, "set random seed"},
    {"bed", 0, true, "PLINK bed file"},
    {"bim", 0, true, "PLINK bim file"},
    {"fam", 0, true, "PLINK fam file"},
    {"pheno", 0, true, "PLINK phenotype file"},
    {"bfile", 0, true, "PLINK root name"},
    {"ndim", 'd', true, "number of PCs to output"},
    {"standx", 's', true, "standardization method for genotypes [binom2 | binom]"},
    {"standy", 0, true, "standardization method for phenotypes [sd | binom2 | binom | none | center]"},
    {"div", 0, true, "whether to divide the eigenvalues by p, n - 1, or don't divide [p | n1 | none]"},
    {"outpc", 0, true, "PC output file"},
    {"outpcx", 0, true, "X PC output file, for CCA"},
    {"outpcy", 0, true, "Y PC output file, for CCA"},
    {"outvec", 0, true, "eigenvector output file"},
    {"outload", 0, true, "SNP loadings"},
    {"outvecx", 0, true, "X eigenvector output file, for CCA"},
    {"outvecy", 0, true, "Y eigenvector output file, for CCA"},
    {"outval", 0, true, "Eigenvalue output file"},
    {"outpve", 0, true, "proportion of variance explained output file"},
    {"outmeansd", 0, true, "mean+SD (used to standardize SNPs) output file"},
    {"outproj", 0, true, "PCA projection output file"},
    {"inload", 0, true, "SNP loadings input file"},
    {"inmeansd", 0, true, "mean+SD (used to standardize SNPs) input file"},
    {"inmaf", 0, true, "MAF input file"},
    {"verbose", 'v', false, "verbose"},
    {"tol", 0, true, "tolerance for PCA iterations"},
    {"lambda1", 0, true, "1st penalty for CCA/SCCA"},
    {"lambda2", 0, true, "2nd penalty for CCA/SCCA"},
    {"maxiter", 0, true, "maximum number of SCCA iterations"},
    {"debug", 0, false, "debug, dumps all intermediate data (WARNING: slow, call only on small data)"},
    {"suffix", 'f', true, "suffix for all output files"},
    {"check", 'c', false, "check eigenvalues/eigenvectors"},
    {"precision", 0, true, "digits of precision for output"},
    {"notime", 0, false, "don't print timestamp in output"},
    {"save-vinit", 0, false, "saves the initial v eigenvector for SCCA"},
    {"version", 0, false, "version"},
    {"device", 0, true, "CUDA device ordinal (B200 build only; default 0)"},
};

struct VarMap {
  std::map<std::string, std::string> v;
  int count(const std::string& k) const { return (int)v.count(k); }
  const std::string& str(const std::string& k) const { return v.at(k); }
  long as_long(const std::string& k) const {
    char* end;
    const std::string& s = v.at(k);
    long r = strtol(s.c_str(), &end, 10);
    if (*end != '\0' || s.empty())
      throw std::invalid_argument("the argument ('" + s + "') for option '--" + k + "' is invalid");
    return r;
  }
  double as_double(const std::string& k) const {
    char* end;
    const std::string& s = v.at(k);
    double r = strtod(s.c_str(), &end);
    if (*end != '\0' || s.empty())
      throw std::invalid_argument("the argument ('" + s + "') for option '--" + k + "' is invalid");
    return r;
  }
};

const OptSpec* find_long(const std::string& name) {
  for (const OptSpec& o : kOptions)
    if (name == o.name) return &o;
  return nullptr;
}
const OptSpec* find_short(char c) {
  for (const OptSpec& o : kOptions)
    if (o.shortname && o.shortname == c) return &o;
  return nullptr;
}

void parse_command_line(int argc, char* argv[], VarMap& vm) {
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    const OptSpec* o = nullptr;
    std::string val;
    bool have_val = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      std::string name = a.substr(2);
      size_t eq = name.find('=');
      if (eq != std::string::npos) {
        val = name.substr(eq + 1);
        name = name.substr(0, eq);
        have_val = true;
      }
      o = find_long(name);
      if (!o) throw std::invalid_argument("unrecognised option '" + a + "'");
    } else if (a.size() >= 2 && a[0] == '-') {
      o = find_short(a[1]);
      if (!o) throw std::invalid_argument("unrecognised option '" + a + "'");
      if (a.size() > 2) {
        val = a.substr(2);
        have_val = true;
      }
    } else {
      throw std::invalid_argument("too many positional options have been specified on the command line");
    }
    if (o->has_value) {
      if (!have_val) {
        if (i + 1 >= argc)
          throw std::invalid_argument(std::string("the required argument for option '--") + o->name +
                                      "' is missing");
        val = argv[++i];
      }
      if (vm.count(o->name))
        throw std::invalid_argument(std::string("option '--") + o->name +
                                    "' cannot be specified more than once");
      vm.v[o->name] = val;
    } else {
      vm.v[o->name] = "";
    }
  }
}

void print_options(std::ostream& os) {
  os << "Options:" << std::endl;
  for (const OptSpec& o : kOptions) {
    std::string left = "  ";
    if (o.shortname) left += std::string("-") + o.shortname + " [ --" + o.name + " ]";
    else left += std::string("--") + o.name;
    if (o.has_value) left += " arg";
    os << left;
    for (size_t k = left.size(); k < 28; k++) os << ' ';
    os << o.help << std::endl;
  }
}

}  // namespace

int main(int argc, char* argv[]) {
  VarMap vm;
  try {
    parse_command_line(argc, argv, vm);
  } catch (std::exception& e) {
    std::cerr << e.what() << std::endl << "Use --help to get more help" << std::endl;
    return EXIT_SUCCESS;  // as upstream (flashpca.cpp:100-105)
  }

  show_timestamp = !vm.count("notime");
  bool verbose = vm.count("verbose");

  std::cout << timestamp() << "arguments: flashpca ";
  for (int i = 0; i < argc; i++) std::cout << argv[i] << " ";
  std::cout << std::endl;

  if (vm.count("version")) {
    std::cerr << "flashpca " << VERSION << std::endl;
    std::cerr << "B200-native build of the FlashPCA2 PCA path." << std::endl
              << "This is free software; see the source for copying conditions.  There is NO"
              << std::endl
              << "warranty; not even for MERCHANTABILITY or FITNESS FOR A PARTICULAR PURPOSE."
              << std::endl
              << std::endl;
    return EXIT_SUCCESS;
  }
  if (vm.count("help")) {
    std::cerr << "flashpca " << VERSION << std::endl;
    print_options(std::cerr);
    return EXIT_SUCCESS;
  }

  try {
    // ---- mode selection (flashpca.cpp:136-228)
    int mode = MODE_PCA;
    const std::vector<std::string> modes = {"cca", "ucca", "scca", "check", "project"};
    for (const char* unsupported : {"scca", "ucca"}) {
      if (vm.count(unsupported)) {
        for (const std::string& m : modes)
          if (m != unsupported && vm.count(m)) {
            std::cerr << "Error: conflicting modes requested: --" << unsupported << ", --" << m
                      << std::endl << "Use --help to get more help" << std::endl;
            return EXIT_FAILURE;
          }
        std::cerr << "Error: --" << unsupported
                  << " is not part of the B200 build (PCA, --check and --project only)"
                  << std::endl;
        return EXIT_FAILURE;
      }
    }
    if (vm.count("check")) {
      if (vm.count("project")) {
        std::cerr << "Error: conflicting modes requested: --check, --project" << std::endl
                  << "Use --help to get more help" << std::endl;
        return EXIT_FAILURE;
      }
      mode = MODE_CHECK_PCA;
    } else if (vm.count("project")) {
      mode = MODE_PREDICT_PCA;
      if (!vm.count("inload")) {
        std::cerr << "Error: SNP-loadings must be specified using --inload" << std::endl;
        return EXIT_FAILURE;
      }
      if (!vm.count("inmaf") && !vm.count("inmeansd")) {
        std::cerr << "Error: one of MAF or mean/stdev must be specified using "
                  << " --inmaf or --inmeansd, respectively" << std::endl;
        return EXIT_FAILURE;
      }
    }
    // --batch: all genotypes as doubles (flashpca.cpp:229-234, MEM_MODE_OFFLINE)
    bool batch = vm.count("batch") && mode == MODE_PCA;

    int memory = 2048;
    if (vm.count("memory")) {
      memory = (int)vm.as_long("memory");
      if (memory < 1) {
        std::cerr << "Error: memory (MB) must be >=1" << std::endl;
        return EXIT_FAILURE;
      }
    }
    unsigned int block_size = 0;
    if (vm.count("blocksize")) {
      if (vm.count("memory")) {
        std::cerr << "Error: cannot specify both --memory and --blocksize"
                  << " at the same time" << std::endl;
        return EXIT_FAILURE;
      }
      long bs = vm.as_long("blocksize");
      if (bs < 1) {
        std::cerr << "Error: blocksize must be >=1" << std::endl;
        return EXIT_FAILURE;
      }
      block_size = (unsigned int)bs;
    }
    long seed = 1L;
    if (vm.count("seed")) seed = vm.as_long("seed");

    std::string fam_file, geno_file, bim_file;
    if (vm.count("bfile")) {
      geno_file = vm.str("bfile") + std::string(".bed");
      bim_file = vm.str("bfile") + std::string(".bim");
      fam_file = vm.str("bfile") + std::string(".fam");
    } else {
      bool good = true;
      if (vm.count("bed")) geno_file = vm.str("bed");
      else good = false;
      if (good && vm.count("bim")) bim_file = vm.str("bim");
      else good = false;
      if (good && vm.count("fam")) fam_file = vm.str("fam");
      else good = false;
      if (!good) {
        std::cerr << "Error: you must specify either --bfile "
                  << "or --bed / --fam / --bim" << std::endl
                  << "Use --help to get more help" << std::endl;
        return EXIT_FAILURE;
      }
    }

    int n_dim = 10;
    if (vm.count("ndim")) {
      n_dim = (int)vm.as_long("ndim");
      if (n_dim < 1) {
        std::cerr << "Error: --ndim can't be less than 1" << std::endl;
        return EXIT_FAILURE;
      }
    }

    int stand_method_x = STANDARDISE_BINOM2;
    if (vm.count("standx")) {
      std::string m = vm.str("standx");
      if (m == "binom") stand_method_x = STANDARDISE_BINOM;
      else if (m == "binom2") stand_method_x = STANDARDISE_BINOM2;
      else {
        std::cerr << "Error: unknown standardization method (--standx): " << m << std::endl;
        return EXIT_FAILURE;
      }
    }

    std::string suffix = ".txt";
    if (vm.count("suffix")) suffix = vm.str("suffix");
    std::string pcfile = "pcs" + suffix;
    if (vm.count("outpc")) pcfile = vm.str("outpc");
    std::string eigvecfile = "eigenvectors" + suffix;
    if (vm.count("outvec")) eigvecfile = vm.str("outvec");
    std::string eigvalfile = "eigenvalues" + suffix;
    if (vm.count("outval")) eigvalfile = vm.str("outval");
    std::string eigpvefile = "pve" + suffix;
    if (vm.count("outpve")) eigpvefile = vm.str("outpve");
    std::string meansdfile = "meansd" + suffix;
    bool save_meansd = false;
    if (vm.count("outmeansd")) {
      meansdfile = vm.str("outmeansd");
      save_meansd = true;
    }
    std::string projfile = "projection" + suffix;
    if (vm.count("outproj")) projfile = vm.str("outproj");

    int maxiter = 500;
    bool debug = vm.count("debug");
    if (vm.count("maxiter")) {
      maxiter = (int)vm.as_long("maxiter");
      if (maxiter <= 0) {
        std::cerr << "Error: --maxiter can't be less than 1" << std::endl;
        return EXIT_FAILURE;
      }
    }
    double tol = 1e-6;
    if (vm.count("tol")) {
      tol = vm.as_double("tol");
      if (tol <= 0) {
        std::cerr << "Error: --tol can't be zero or negative" << std::endl;
        return EXIT_FAILURE;
      }
    }
    bool do_loadings = false;
    std::string loadingsfile = "";
    if (vm.count("outload")) {
      loadingsfile = vm.str("outload");
      do_loadings = true;
    }
    int divisor = DIVISOR_P;
    if (vm.count("div")) {
      std::string m = vm.str("div");
      if (m == "none") divisor = DIVISOR_NONE;
      else if (m == "n1") divisor = DIVISOR_N1;
      else if (m == "p") divisor = DIVISOR_P;
      else {
        std::cerr << "Error: unknown divisor (--div): " << m << std::endl;
        return EXIT_FAILURE;
      }
    }
    std::string in_meansd_file = "", in_maf_file = "";
    if (vm.count("inmeansd")) {
      if (vm.count("inmaf")) {
        std::cerr << "Error: conflicting options requested --inmeansd, --inmaf" << std::endl;
        return EXIT_FAILURE;
      }
      in_meansd_file = vm.str("inmeansd");
      if (in_meansd_file == "") {
        std::cerr << "Error: no file specified for --inmeansd" << std::endl;
        return EXIT_FAILURE;
      }
    } else if (vm.count("inmaf")) {
      in_maf_file = vm.str("inmaf");
      if (in_maf_file == "") {
        std::cerr << "Error: no file specified for --inmaf" << std::endl;
        return EXIT_FAILURE;
      }
    }
    std::string in_load_file = "";
    if (vm.count("inload")) {
      in_load_file = vm.str("inload");
      if (in_load_file == "") {
        std::cerr << "Error: no file specified for --inload" << std::endl;
        return EXIT_FAILURE;
      }
    }
    int precision = 7;
    if (vm.count("precision")) {
      precision = (int)vm.as_long("precision");
      if (precision <= 1) {
        std::cerr << "Error: output --precision too low" << std::endl;
        return EXIT_FAILURE;
      }
    }
    int device = 0;
    if (vm.count("device")) device = (int)vm.as_long("device");

    // ---- end of command line parsing
    std::cout << timestamp() << "Start flashpca (version " << VERSION << ")" << std::endl;

    Data data;
    data.verbose = verbose;
    data.stand_method_x = stand_method_x;
    verbose&& std::cout << timestamp() << "seed: " << seed << std::endl;

    data.read_pheno(fam_file.c_str(), 6);
    data.read_plink_bim(bim_file.c_str());
    data.read_plink_fam(fam_file.c_str());
    data.geno_filename = geno_file;
    data.get_size();
    data.prepare();
    if (batch) data.read_bed(false);  // flashpca.cpp:597-601

    RandomPCA rpca;
    rpca.verbose = verbose;
    rpca.debug = debug;
    rpca.stand_method_x = stand_method_x;
    rpca.divisor = divisor;
    rpca.device = device;

    // ncv = 2*ndim+1 --> ndim < (n-1)/2   (flashpca.cpp:623-633)
    unsigned int max_dim = (unsigned int)((fminl(data.N, data.nsnps) - 1) / 2.0);
    if ((unsigned int)n_dim > max_dim) {
      std::cerr << "Error: You asked for " << n_dim << " dimensions, but only " << max_dim
                << "allowed" << std::endl;
      return EXIT_FAILURE;
    }

    // --memory -> block_size (flashpca.cpp:636-688).  The value only feeds the
    // log line below: the genotypes are resident in HBM, not re-read in blocks.
    long long mem = (long long)memory * 1048576;
    if (block_size == 0) {
      long long mem_req_bytes = 2 * (long long)data.nsnps * 8 * 2 + 3 * (long long)data.nsnps * 8 +
                                (long long)data.N * n_dim * 8 +
                                (do_loadings ? (long long)data.nsnps * n_dim * 8 : 0) +
                                2 * (long long)data.N +
                                2 * (long long)(data.N + data.nsnps) * n_dim * 8 +
                                2 * 1024 * 1024 + (long long)data.N * 8;
      long long mem_remain_bytes = mem - mem_req_bytes;
      verbose&& std::cout << timestamp() << "mem: " << mem << " mem_req_bytes: " << mem_req_bytes
                          << " mem_remain_bytes: " << mem_remain_bytes << std::endl;
      if (mem_remain_bytes <= 0) {
        std::cerr << "The memory specified using --memory is not sufficient, try"
                  << " increasing it to at least " << (mem_req_bytes + data.N * 8) / 1048576
                  << " MB" << std::endl;
        return EXIT_FAILURE;
      }
      block_size = (unsigned int)floor(mem_remain_bytes / ((double)data.N * 8.0));
      if (block_size < 1) {
        std::cerr << "The memory specified using --memory is not sufficient, try"
                  << " increasing it" << std::endl;
        return EXIT_FAILURE;
      }
    }
    block_size = (unsigned int)fminl(block_size, data.nsnps);
    std::cout << timestamp() << "blocksize: " << block_size << " ("
              << (long long)block_size * 8 * data.N << " bytes per block)" << std::endl;

    // ---- the main analysis
    if (mode == MODE_PCA) {
      std::cout << timestamp() << "PCA begin" << std::endl;
      if (batch) rpca.pca_fast(data.X, block_size, n_dim, maxiter, tol, seed, do_loadings);
      else rpca.pca_fast(data, block_size, n_dim, maxiter, tol, seed, do_loadings);
      std::cout << timestamp() << "PCA done" << std::endl;
    } else if (mode == MODE_CHECK_PCA) {
      rpca.check(data, block_size, eigvecfile, eigvalfile);
    } else if (mode == MODE_PREDICT_PCA) {
      rpca.project(data, block_size, in_load_file, in_maf_file, in_meansd_file);
    }

    // ---- write out results (flashpca.cpp:755-878)
    const std::vector<std::string> none;
    if (mode == MODE_PCA) {
      std::cout << timestamp() << "Writing " << n_dim << " eigenvalues to file " << eigvalfile
                << std::endl;
      save_text(rpca.d, none, none, eigvalfile.c_str(), precision);

      std::cout << timestamp() << "Writing " << n_dim << " eigenvectors to file " << eigvecfile
                << std::endl;
      std::vector<std::string> rownames(rpca.Px.rows());
      for (size_t i = 0; i < rpca.Px.rows(); i++)
        rownames[i] = data.fam_ids[i] + TXT_SEP + data.indiv_ids[i];
      std::vector<std::string> colnames(rpca.Px.cols() + 1);
      colnames[0] = std::string("FID") + TXT_SEP + "IID";
      for (size_t i = 0; i < rpca.Px.cols(); i++) colnames[i + 1] = "U" + std::to_string(i + 1);
      save_text(rpca.U, colnames, rownames, eigvecfile.c_str(), precision);

      std::cout << timestamp() << "Writing " << n_dim << " PCs to file " << pcfile << std::endl;
      for (size_t i = 0; i < rpca.Px.cols(); i++) colnames[i + 1] = "PC" + std::to_string(i + 1);
      save_text(rpca.Px, colnames, rownames, pcfile.c_str(), precision);

      std::cout << timestamp() << "Writing " << n_dim << " proportion variance explained to file "
                << eigpvefile << std::endl;
      save_text(rpca.pve, none, none, eigpvefile.c_str(), precision);

      if (do_loadings) {
        std::cout << timestamp() << "Writing"
                  << " SNP loadings to file " << loadingsfile << std::endl;
        std::vector<std::string> lcol = {std::string("SNP") + TXT_SEP + "RefAllele"};
        for (size_t i = 0; i < rpca.V.cols(); i++)
          lcol.push_back(std::string("V") + std::to_string(i + 1));
        std::vector<std::string> lrow(data.snp_ids.size());
        for (size_t i = 0; i < lrow.size(); i++)
          lrow[i] = data.snp_ids[i] + TXT_SEP + data.ref_alleles[i];
        save_text(rpca.V, lcol, lrow, loadingsfile.c_str(), precision);
      }
    } else if (mode == MODE_PREDICT_PCA) {
      std::vector<std::string> rownames(rpca.Px.rows());
      for (size_t i = 0; i < rpca.Px.rows(); i++)
        rownames[i] = data.fam_ids[i] + TXT_SEP + data.indiv_ids[i];
      std::vector<std::string> colnames(rpca.Px.cols() + 1);
      colnames[0] = std::string("FID") + TXT_SEP + "IID";
      for (size_t i = 0; i < rpca.Px.cols(); i++) colnames[i + 1] = "PC" + std::to_string(i + 1);
      save_text(rpca.Px, colnames, rownames, projfile.c_str(), precision);
    } else if (mode == MODE_CHECK_PCA) {
      std::cout << timestamp() << "Mean squared error: " << rpca.mse
                << ", Root mean squared error: " << rpca.rmse << " (n=" << data.N << ")"
                << std::endl;
    }

    if (save_meansd) {
      std::cout << timestamp() << "Writing mean + sd file " << meansdfile << std::endl;
      std::vector<std::string> v = {std::string("SNP") + TXT_SEP + "RefAllele", "Mean", "SD"};
      std::vector<std::string> rownames(data.snp_ids.size());
      for (size_t i = 0; i < rownames.size(); i++)
        rownames[i] = data.snp_ids[i] + TXT_SEP + data.ref_alleles[i];
      save_text(rpca.X_meansd, v, rownames, meansdfile.c_str(), precision);
    }

    std::cout << timestamp() << "Goodbye!" << std::endl;
  } catch (std::exception& e) {
    std::cerr << timestamp() << "Exception: " << e.what() << std::endl;
    std::cerr << timestamp() << "Terminating" << std::endl;
    return EXIT_FAILURE;
  } catch (...) {
    std::cerr << timestamp() << "Caught unknown exception, terminating " << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
