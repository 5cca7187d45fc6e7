// matrix_io.hpp -- SMS reader/writer and the other print formats of the CLIs.
// SMS (reference data/README.md:10-17): optional '#' comment lines, header `rows cols R|M`,
// 1-based `i j value` lines (value = integer or a/b, entries in any order), terminator `0 0 0`.
// Writers follow the shapes LinBox produces for FileFormat(5) SMS (15 `M`-typed files in the
// reference's data/ show it: header `r c M`, row-major entries, `0 0 0`), Maple (1) and a
// bracketed row listing for Pretty (8) / Linalg (12); LinBox's writer itself is not in the
// reference tree (parity unpinned for the pretty layouts).
#pragma once
#include <istream>
#include <ostream>
#include <sstream>
#include <string>

#include "exact.hpp"

namespace plo {
namespace host {

enum FileFormat { FF_MAPLE = 1, FF_SMS = 5, FF_PRETTY = 8, FF_LINALG = 12 };

inline Rat parse_rational(const std::string& tok) {
  const size_t slash = tok.find('/');
  try {
    size_t used = 0;
    const long long n = std::stoll(tok.substr(0, slash), &used);
    if (used != (slash == std::string::npos ? tok.size() : slash)) throw std::invalid_argument(tok);
    if (slash == std::string::npos) return Rat((int64_t)n);
    const std::string ds = tok.substr(slash + 1);
    const long long d = std::stoll(ds, &used);
    if (used != ds.size()) throw std::invalid_argument(tok);
    return Rat::make(n, d);
  } catch (const std::exception&) {
    throw RangeError("cannot parse matrix entry '" + tok + "' (polynomial entries are out of scope)");
  }
}

inline bool read_sms(std::istream& in, Dense<QField>& M) {
  QField Q;
  std::string line;
  bool have_header = false;
  while (std::getline(in, line)) {
    size_t b = line.find_first_not_of(" \t\r");
    if (b == std::string::npos || line[b] == '#' || line[b] == '%') continue;
    std::istringstream ls(line);
    if (!have_header) {
      long long r, c;
      if (!(ls >> r >> c) || r < 0 || c < 0) return false;
      M = Dense<QField>(Q, (size_t)r, (size_t)c);
      have_header = true;
      continue;
    }
    long long i, j;
    std::string val;
    if (!(ls >> i >> j >> val)) return false;
    if (i == 0 && j == 0) break;
    if (i < 1 || j < 1 || (size_t)i > M.rows || (size_t)j > M.cols) return false;
    M.at((size_t)i - 1, (size_t)j - 1) = parse_rational(val);
  }
  return have_header;
}

inline void print_elt(std::ostream& o, const Rat& r) { o << r.num; if (r.den != 1) o << '/' << r.den; }
inline void print_elt(std::ostream& o, const int64_t& r) { o << r; }

template <class F>
std::ostream& write_matrix(std::ostream& out, const F& f, const Dense<F>& M, int format) {
  if (format == FF_SMS) {
    out << M.rows << ' ' << M.cols << " M\n";
    for (size_t i = 0; i < M.rows; ++i)
      for (size_t j = 0; j < M.cols; ++j)
        if (!f.is_zero(M.at(i, j))) { out << i + 1 << ' ' << j + 1 << ' '; print_elt(out, M.at(i, j)); out << '\n'; }
    return out << "0 0 0\n";
  }
  if (format == FF_MAPLE) {
    out << "Matrix(" << M.rows << ',' << M.cols << ",[";
    for (size_t i = 0; i < M.rows; ++i) {
      out << (i ? ",[" : "[");
      for (size_t j = 0; j < M.cols; ++j) { if (j) out << ','; print_elt(out, M.at(i, j)); }
      out << ']';
    }
    return out << "])";
  }
  for (size_t i = 0; i < M.rows; ++i) {  // Pretty / Linalg
    out << "  [";
    for (size_t j = 0; j < M.cols; ++j) { out << (j ? " " : ""); print_elt(out, M.at(i, j)); }
    out << " ]\n";
  }
  return out;
}

}  // namespace host
}  // namespace plo
