#include "util.hpp"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>

namespace flashpca {

bool show_timestamp = true;

std::string timestamp() {
  if (!show_timestamp) return std::string("");
  time_t t = time(NULL);
  char buf[64];
  struct tm tmv;
  localtime_r(&t, &tmv);
  asctime_r(&tmv, buf);
  size_t len = strlen(buf);
  if (len && buf[len - 1] == '\n') buf[len - 1] = '\0';
  return std::string("[") + buf + "] ";
}

static bool save_impl(const double* data, size_t rows, size_t cols, size_t ld,
                      const std::vector<std::string>& colnames,
                      const std::vector<std::string>& rownames, const char* filename,
                      unsigned int precision) {
  std::ofstream out(filename, std::ofstream::out);
  out << std::setprecision(precision);
  if (!out) {
    std::cerr << "Error while saving to file " << filename << ":" << strerror(errno) << std::endl;
    return false;
  }
  for (size_t i = 0; i < colnames.size(); i++) {
    out << colnames[i];
    if (i == colnames.size() - 1) out << std::endl;
    else out << TXT_SEP;
  }
  for (size_t j = 0; j < rows; j++) {
    if (!rownames.empty()) out << rownames[j] << TXT_SEP;
    for (size_t c = 0; c < cols; c++) {
      if (c) out << TXT_SEP;
      out << data[c * ld + j];
    }
    out << std::endl;
  }
  out.close();
  return true;
}

bool save_text(const Matrix& m, const std::vector<std::string>& colnames,
               const std::vector<std::string>& rownames, const char* filename,
               unsigned int precision) {
  return save_impl(m.data(), m.rows(), m.cols(), m.rows(), colnames, rownames, filename, precision);
}

bool save_text(const Vector& v, const std::vector<std::string>& colnames,
               const std::vector<std::string>& rownames, const char* filename,
               unsigned int precision) {
  return save_impl(v.data(), v.size(), 1, v.size(), colnames, rownames, filename, precision);
}

Matrix read_text(const char* filename, unsigned int firstcol, unsigned int skip) {
  std::ifstream in(filename, std::ios::in);
  if (!in)
    throw std::runtime_error(std::string("Error reading file '") + filename +
                             "': " + strerror(errno));
  std::vector<std::string> lines;
  unsigned int line_num = 0;
  while (in) {
    std::string line;
    std::getline(in, line);
    if (!in.eof()) {
      if (line_num >= skip) lines.push_back(line);
      line_num++;
    }
  }
  Matrix m;
  size_t numfields_1st = 0;
  for (size_t i = 0; i < lines.size(); i++) {
    std::stringstream ss(lines[i]);
    std::string s;
    std::vector<std::string> tokens;
    while (ss >> s) tokens.push_back(s);
    size_t numfields = tokens.size() + 1 >= firstcol ? tokens.size() - firstcol + 1 : 0;
    if (i == 0) {
      m = Matrix(lines.size(), numfields);
      numfields_1st = numfields;
    } else if (numfields_1st != numfields) {
      throw std::runtime_error(std::string("Error reading file '") + filename +
                               "': inconsistent number of columns");
    }
    for (size_t j = 0; j < numfields; j++) {
      const std::string& tok = tokens[j + firstcol - 1];
      char* end;
      errno = 0;
      double val = std::strtod(tok.c_str(), &end);
      if (*end != '\0' || errno != 0)
        throw std::runtime_error(std::string("Error reading file '") + filename + "', line " +
                                 std::to_string(i + 1) + ": '" + tok +
                                 "' cannot be parsed as a number");
      m(i, j) = val;
    }
  }
  return m;
}

}  // namespace flashpca
