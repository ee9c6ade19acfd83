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

// A failed option check: message for stderr + process exit status.
struct UsageError {
  std::string msg;
  int status;
};
[[noreturn]] void fail(const std::string& msg) { throw UsageError{msg, EXIT_FAILURE}; }

// Everything main() needs after the command line has been checked.
struct Config {
  int mode = MODE_PCA;
  bool batch = false, verbose = false, debug = false, do_loadings = false, save_meansd = false;
  int memory = 2048, n_dim = 10, maxiter = 500, precision = 7, device = 0;
  int stand_method_x = STANDARDISE_BINOM2, divisor = DIVISOR_P;
  unsigned int block_size = 0;
  long seed = 1L;
  double tol = 1e-6;
  std::string bed, bim, fam;
  std::map<std::string, std::string> out;  // logical output name -> file name
  std::string in_load, in_maf, in_meansd;
};

// keyword -> code, with upstream's "unknown ..." message on a miss
int lookup(const VarMap& vm, const char* opt, const std::map<std::string, int>& table,
           const char* what, int dflt) {
  if (!vm.count(opt)) return dflt;
  auto it = table.find(vm.str(opt));
  if (it == table.end()) fail(std::string("Error: unknown ") + what + " (--" + opt + "): " + vm.str(opt));
  return it->second;
}

// integer option with a lower bound and upstream's message when it is violated
long bounded(const VarMap& vm, const char* opt, long dflt, long lowest, const char* msg) {
  if (!vm.count(opt)) return dflt;
  long v = vm.as_long(opt);
  if (v < lowest) fail(msg);
  return v;
}

// The option checks of upstream flashpca.cpp:136-564, in upstream's order of
// precedence, expressed as data where they are uniform.
Config check_options(const VarMap& vm) {
  Config c;
  c.verbose = vm.count("verbose");
  c.debug = vm.count("debug");

  // -- analysis mode
  const char* all_modes[] = {"cca", "ucca", "scca", "check", "project"};
  auto conflicts = [&](const std::string& chosen) {
    for (const char* other : all_modes)
      if (chosen != other && vm.count(other))
        fail("Error: conflicting modes requested: --" + chosen + ", --" + other +
             "\nUse --help to get more help");
  };
  for (const char* gone : {"scca", "ucca"})
    if (vm.count(gone)) {
      conflicts(gone);
      fail(std::string("Error: --") + gone +
           " is not part of the B200 build (PCA, --check and --project only)");
    }
  if (vm.count("check")) {
    conflicts("check");
    c.mode = MODE_CHECK_PCA;
  } else if (vm.count("project")) {
    c.mode = MODE_PREDICT_PCA;
    if (!vm.count("inload")) fail("Error: SNP-loadings must be specified using --inload");
    if (!vm.count("inmaf") && !vm.count("inmeansd"))
      fail("Error: one of MAF or mean/stdev must be specified using "
           " --inmaf or --inmeansd, respectively");
  }
  c.batch = vm.count("batch") && c.mode == MODE_PCA;  // MEM_MODE_OFFLINE applies to PCA only

  // -- sizes
  c.memory = (int)bounded(vm, "memory", 2048, 1, "Error: memory (MB) must be >=1");
  if (vm.count("blocksize")) {
    if (vm.count("memory"))
      fail("Error: cannot specify both --memory and --blocksize at the same time");
    c.block_size = (unsigned int)bounded(vm, "blocksize", 0, 1, "Error: blocksize must be >=1");
  }
  if (vm.count("seed")) c.seed = vm.as_long("seed");

  // -- input fileset
  if (vm.count("bfile")) {
    const std::string& root = vm.str("bfile");
    c.bed = root + ".bed";
    c.bim = root + ".bim";
    c.fam = root + ".fam";
  } else if (vm.count("bed") && vm.count("bim") && vm.count("fam")) {
    c.bed = vm.str("bed");
    c.bim = vm.str("bim");
    c.fam = vm.str("fam");
  } else {
    fail("Error: you must specify either --bfile or --bed / --fam / --bim\n"
         "Use --help to get more help");
  }

  c.n_dim = (int)bounded(vm, "ndim", 10, 1, "Error: --ndim can't be less than 1");
  c.stand_method_x = lookup(vm, "standx", {{"binom", STANDARDISE_BINOM}, {"binom2", STANDARDISE_BINOM2}},
                            "standardization method", STANDARDISE_BINOM2);

  // -- output names: <stem><suffix> unless an explicit --out* option is given
  const std::string suffix = vm.count("suffix") ? vm.str("suffix") : ".txt";
  const struct { const char *key, *stem, *opt; } outs[] = {
      {"pcs", "pcs", "outpc"},           {"eigenvectors", "eigenvectors", "outvec"},
      {"eigenvalues", "eigenvalues", "outval"}, {"pve", "pve", "outpve"},
      {"meansd", "meansd", "outmeansd"}, {"projection", "projection", "outproj"}};
  for (const auto& o : outs) c.out[o.key] = vm.count(o.opt) ? vm.str(o.opt) : o.stem + suffix;
  c.save_meansd = vm.count("outmeansd");
  if (vm.count("outload")) {
    c.out["loadings"] = vm.str("outload");
    c.do_loadings = true;
  }

  // -- solver
  c.maxiter = (int)bounded(vm, "maxiter", 500, 1, "Error: --maxiter can't be less than 1");
  if (vm.count("tol")) {
    c.tol = vm.as_double("tol");
    if (c.tol <= 0) fail("Error: --tol can't be zero or negative");
  }
  c.divisor = lookup(vm, "div", {{"none", DIVISOR_NONE}, {"n1", DIVISOR_N1}, {"p", DIVISOR_P}},
                     "divisor", DIVISOR_P);

  // -- projection inputs
  if (vm.count("inmeansd") && vm.count("inmaf"))
    fail("Error: conflicting options requested --inmeansd, --inmaf");
  for (const auto& in : {std::make_pair("inmeansd", &c.in_meansd), std::make_pair("inmaf", &c.in_maf),
                         std::make_pair("inload", &c.in_load)}) {
    if (!vm.count(in.first)) continue;
    *in.second = vm.str(in.first);
    if (in.second->empty()) fail(std::string("Error: no file specified for --") + in.first);
  }

  c.precision = (int)bounded(vm, "precision", 7, 2, "Error: output --precision too low");
  if (vm.count("device")) c.device = (int)vm.as_long("device");
  return c;
}

// --memory (MB) -> number of SNPs per block, upstream flashpca.cpp:636-688.  The
// value only feeds the "blocksize" log line: the genotypes are resident in HBM.
unsigned int block_size_from_memory(const Config& c, const Data& data) {
  const long long n = data.N, p = data.nsnps, k = c.n_dim;
  const long long reserve = 2 * p * 8 * 2            // avg + stdev
                            + 3 * p * 8              // genotype table
                            + n * k * 8              // U
                            + (c.do_loadings ? p * k * 8 : 0)  // V
                            + 2 * n                  // PLINK buffers
                            + 2 * (n + p) * k * 8    // solver workspace
                            + 2 * 1024 * 1024 + n * 8;
  const long long budget = (long long)c.memory * 1048576, left = budget - reserve;
  if (c.verbose)
    std::cout << timestamp() << "mem: " << budget << " mem_req_bytes: " << reserve
              << " mem_remain_bytes: " << left << std::endl;
  if (left <= 0)
    fail("The memory specified using --memory is not sufficient, try increasing it to at least " +
         std::to_string((reserve + n * 8) / 1048576) + " MB");
  unsigned int bs = (unsigned int)floor(left / ((double)n * 8.0));
  if (bs < 1) fail("The memory specified using --memory is not sufficient, try increasing it");
  return bs;
}

std::vector<std::string> paired(const std::vector<std::string>& a, const std::vector<std::string>& b) {
  std::vector<std::string> r(a.size());
  for (size_t i = 0; i < a.size(); i++) r[i] = a[i] + TXT_SEP + b[i];
  return r;
}
std::vector<std::string> numbered(const std::string& first, const std::string& prefix, size_t n) {
  std::vector<std::string> r{first};
  for (size_t i = 1; i <= n; i++) r.push_back(prefix + std::to_string(i));
  return r;
}

// Result files of upstream flashpca.cpp:755-878.
void write_outputs(const Config& c, const Data& data, RandomPCA& rpca) {
  const std::vector<std::string> none;
  const std::string idhdr = std::string("FID") + TXT_SEP + "IID";
  const std::string snphdr = std::string("SNP") + TXT_SEP + "RefAllele";
  auto announce = [&](const std::string& what, const std::string& file) {
    std::cout << timestamp() << "Writing " << what << " to file " << file << std::endl;
  };
  const std::string k = std::to_string(c.n_dim);
  if (c.mode == MODE_PCA) {
    const std::vector<std::string> people = paired(data.fam_ids, data.indiv_ids);
    announce(k + " eigenvalues", c.out.at("eigenvalues"));
    save_text(rpca.d, none, none, c.out.at("eigenvalues").c_str(), c.precision);
    announce(k + " eigenvectors", c.out.at("eigenvectors"));
    save_text(rpca.U, numbered(idhdr, "U", rpca.U.cols()), people, c.out.at("eigenvectors").c_str(),
              c.precision);
    announce(k + " PCs", c.out.at("pcs"));
    save_text(rpca.Px, numbered(idhdr, "PC", rpca.Px.cols()), people, c.out.at("pcs").c_str(),
              c.precision);
    announce(k + " proportion variance explained", c.out.at("pve"));
    save_text(rpca.pve, none, none, c.out.at("pve").c_str(), c.precision);
    if (c.do_loadings) {
      std::cout << timestamp() << "Writing SNP loadings to file " << c.out.at("loadings") << std::endl;
      save_text(rpca.V, numbered(snphdr, "V", rpca.V.cols()), paired(data.snp_ids, data.ref_alleles),
                c.out.at("loadings").c_str(), c.precision);
    }
  } else if (c.mode == MODE_PREDICT_PCA) {
    save_text(rpca.Px, numbered(idhdr, "PC", rpca.Px.cols()), paired(data.fam_ids, data.indiv_ids),
              c.out.at("projection").c_str(), c.precision);
  } else {
    std::cout << timestamp() << "Mean squared error: " << rpca.mse
              << ", Root mean squared error: " << rpca.rmse << " (n=" << data.N << ")" << std::endl;
  }
  if (c.save_meansd) {
    std::cout << timestamp() << "Writing mean + sd file " << c.out.at("meansd") << std::endl;
    save_text(rpca.X_meansd, {snphdr, "Mean", "SD"}, paired(data.snp_ids, data.ref_alleles),
              c.out.at("meansd").c_str(), c.precision);
  }
}

}  // namespace

int main(int argc, char* argv[]) {
  VarMap vm;
  try {
    parse_command_line(argc, argv, vm);
  } catch (std::exception& e) {
    // upstream reports a malformed command line and still exits with success (flashpca.cpp:100-105)
    std::cerr << e.what() << std::endl << "Use --help to get more help" << std::endl;
    return EXIT_SUCCESS;
  }
  show_timestamp = !vm.count("notime");

  std::cout << timestamp() << "arguments: flashpca ";
  for (int i = 0; i < argc; i++) std::cout << argv[i] << " ";
  std::cout << std::endl;

  if (vm.count("version") || vm.count("help")) {
    std::cerr << "flashpca " << VERSION << std::endl;
    if (vm.count("help")) print_options(std::cerr);
    else
      std::cerr << "B200-native build of the FlashPCA2 PCA path." << std::endl
                << "This is free software; see the source for copying conditions.  There is NO"
                << std::endl
                << "warranty; not even for MERCHANTABILITY or FITNESS FOR A PARTICULAR PURPOSE."
                << std::endl << std::endl;
    return EXIT_SUCCESS;
  }

  try {
    const Config c = check_options(vm);
    std::cout << timestamp() << "Start flashpca (version " << VERSION << ")" << std::endl;

    Data data;
    data.verbose = c.verbose;
    data.stand_method_x = c.stand_method_x;
    if (c.verbose) std::cout << timestamp() << "seed: " << c.seed << std::endl;
    data.read_pheno(c.fam.c_str(), 6);  // N = number of fam lines; column 6 must be numeric
    data.read_plink_bim(c.bim.c_str());
    data.read_plink_fam(c.fam.c_str());
    data.geno_filename = c.bed;
    data.get_size();
    data.prepare();
    if (c.batch) data.read_bed(false);

    RandomPCA rpca;
    rpca.verbose = c.verbose;
    rpca.debug = c.debug;
    rpca.stand_method_x = c.stand_method_x;
    rpca.divisor = c.divisor;
    rpca.device = c.device;

    // ncv = 2*ndim+1 must stay below min(N, p): ndim <= (min(N, p) - 1) / 2
    const unsigned int max_dim = (unsigned int)((fminl(data.N, data.nsnps) - 1) / 2.0);
    if ((unsigned int)c.n_dim > max_dim)
      fail("Error: You asked for " + std::to_string(c.n_dim) + " dimensions, but only " +
           std::to_string(max_dim) + "allowed");

    unsigned int block_size = c.block_size ? c.block_size : block_size_from_memory(c, data);
    block_size = (unsigned int)fminl(block_size, data.nsnps);
    std::cout << timestamp() << "blocksize: " << block_size << " ("
              << (long long)block_size * 8 * data.N << " bytes per block)" << std::endl;

    switch (c.mode) {
      case MODE_PCA:
        std::cout << timestamp() << "PCA begin" << std::endl;
        if (c.batch)
          rpca.pca_fast(data.X, block_size, c.n_dim, c.maxiter, c.tol, c.seed, c.do_loadings);
        else
          rpca.pca_fast(data, block_size, c.n_dim, c.maxiter, c.tol, c.seed, c.do_loadings);
        std::cout << timestamp() << "PCA done" << std::endl;
        break;
      case MODE_CHECK_PCA:
        rpca.check(data, block_size, c.out.at("eigenvectors"), c.out.at("eigenvalues"));
        break;
      case MODE_PREDICT_PCA:
        rpca.project(data, block_size, c.in_load, c.in_maf, c.in_meansd);
        break;
    }
    write_outputs(c, data, rpca);
    std::cout << timestamp() << "Goodbye!" << std::endl;
  } catch (UsageError& u) {
    std::cerr << u.msg << std::endl;
    return u.status;
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
