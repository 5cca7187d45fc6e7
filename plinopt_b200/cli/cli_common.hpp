// cli_common.hpp -- shared pieces of the drop-in CLIs (bin/sparsifier, bin/orbiter, bin/MMchecker).
#pragma once
#include <chrono>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/plinopt_b200.h"
#include "../csrc/host/matrix_io.hpp"

namespace cli {

using plo::host::Dense;
using plo::host::QField;
using plo::host::Rat;

struct NumDen {
  std::vector<int64_t> num, den;
  int rows = 0, cols = 0;
};

inline NumDen flatten(const Dense<QField>& M) {
  NumDen o;
  o.rows = (int)M.rows; o.cols = (int)M.cols;
  o.num.resize(M.v.size()); o.den.resize(M.v.size());
  for (size_t e = 0; e < M.v.size(); ++e) { o.num[e] = M.v[e].num; o.den[e] = M.v[e].den; }
  return o;
}
inline Dense<QField> unflatten(int rows, int cols, const std::vector<int64_t>& num, const std::vector<int64_t>& den) {
  QField Q;
  Dense<QField> M(Q, (size_t)rows, (size_t)cols);
  for (size_t e = 0; e < M.v.size(); ++e) M.v[e] = Rat::make(num[e], den[e]);
  return M;
}
inline bool read_file(const std::string& name, Dense<QField>& M) {
  std::ifstream in(name);
  if (!in) { std::cerr << "# \033[1;31m****** ERROR, cannot open " << name << " ******\033[0m" << std::endl; return false; }
  try {
    if (!plo::host::read_sms(in, M)) { std::cerr << "# \033[1;31m****** ERROR, malformed SMS file " << name << " ******\033[0m" << std::endl; return false; }
  } catch (const std::exception& e) {
    std::cerr << "# \033[1;31m****** ERROR, " << name << ": " << e.what() << " ******\033[0m" << std::endl;
    return false;
  }
  return true;
}
inline size_t profile(std::ostream& out, const Dense<QField>& M) {  // densityProfile, plinopt_sparsify.inl:118-126
  size_t ss = 0;
  for (size_t i = 0; i < M.rows; ++i) {
    size_t s = 0;
    for (size_t j = 0; j < M.cols; ++j) s += M.at(i, j).num != 0;
    ss += s;
    out << s << ' ';
  }
  out << '=' << ss;
  return ss;
}
struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
inline std::string replace_extension(const std::string& path, const std::string& ext) {  // std::filesystem::path::replace_extension
  const size_t slash = path.find_last_of('/');
  const size_t dot = path.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return path + ext;
  return path.substr(0, dot) + ext;
}

}  // namespace cli
