// util.hpp -- text output and small helpers of the flashpca command line
// (upstream util.h:33-40 constants, :69-108 save_text, util.cpp:270-283 timestamp).
#pragma once
#include <string>
#include <vector>

#include "matrix.hpp"

#define VAR_TOL 1e-9
#define STANDARDISE_NONE 0
#define STANDARDISE_SD 1
#define STANDARDISE_BINOM 2
#define STANDARDISE_BINOM2 3
#define STANDARDISE_CENTER 4
#define TXT_SEP "\t"

namespace flashpca {

extern bool show_timestamp;
std::string timestamp();

// util.h:69-108: optional header line, optional row names, TAB separated,
// numbers in the stream's general format with `precision` significant digits.
bool save_text(const Matrix& m, const std::vector<std::string>& colnames,
               const std::vector<std::string>& rownames, const char* filename,
               unsigned int precision = 7);
bool save_text(const Vector& v, const std::vector<std::string>& colnames,
               const std::vector<std::string>& rownames, const char* filename,
               unsigned int precision = 7);

// data.cpp:504-586 read_text: whitespace separated numeric table; `firstcol` is
// one-based, `skip` header lines are dropped; a final unterminated line is
// ignored exactly as upstream does.
Matrix read_text(const char* filename, unsigned int firstcol, unsigned int skip);

}  // namespace flashpca
