// slp.hpp -- straight-line program -> matrix.
// Restates what the reference does with  programParser + parenthesisExpand + matrixBuilder
// (include/plinopt_programs.inl:618-686, 1615-1679, 1459-1608; driver src/SLPchecker.cpp:22-40,
// rule data/Makefile:31-32 `%.sms:%.slp`) as a direct evaluation: every variable is a sparse
// linear form in the inputs `i<N>`; the lines `x:=expr;` are evaluated in order with
//   expr := ['+'|'-'] term {('+'|'-') term} ;  term := factor {('*' number | '/' number)} ;
//   factor := variable | '(' expr ')' | 0 ;   number := integer ['/' integer]
// (an output may appear on its own right-hand side: its previous value is used, like :1490-1508).
// Row i of the result is the form of `o<i>` (outchar), column j the coefficient of `i<j>`.
// Needed once on the host to regenerate data/32x32x32_15096_{L,R,P}.sms, which the reference
// ships only as .slp (.MISSING_LARGE_BLOBS:1-3).
#pragma once
#include <cctype>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "exact.hpp"

namespace plo {
namespace host {

struct SparseRows {
  size_t rows = 0, cols = 0;
  std::vector<std::vector<std::pair<int, Rat>>> r;  // sorted by column
  size_t nnz() const { size_t s = 0; for (const auto& x : r) s += x.size(); return s; }
};

class SlpBuilder {
  typedef std::vector<std::pair<int, Rat>> Form;  // sorted by input index
  QField Q;
  std::unordered_map<std::string, Form> vars;
  const char* p = nullptr;
  const char* end = nullptr;
  std::string cur_output;
  int max_input = -1;

  void skip() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p; }
  [[noreturn]] void fail(const std::string& why) const { throw RangeError("SLP parse error: " + why); }

  static void axpy(Form& acc, const Form& x, const Rat& c, const QField& Q) {  // acc += c*x
    if (c.num == 0 || x.empty()) return;
    Form out;
    out.reserve(acc.size() + x.size());
    size_t a = 0, b = 0;
    while (a < acc.size() || b < x.size()) {
      if (b == x.size() || (a < acc.size() && acc[a].first < x[b].first)) out.push_back(acc[a++]);
      else if (a == acc.size() || x[b].first < acc[a].first) { out.emplace_back(x[b].first, Q.mul(c, x[b].second)); ++b; }
      else { const Rat s = Q.add(acc[a].second, Q.mul(c, x[b].second)); if (s.num != 0) out.emplace_back(acc[a].first, s); ++a; ++b; }
    }
    acc.swap(out);
  }
  long long integer() {
    skip();
    if (p >= end || !isdigit((unsigned char)*p)) fail("integer expected");
    long long v = 0;
    while (p < end && isdigit((unsigned char)*p)) { v = v * 10 + (*p - '0'); ++p; }
    return v;
  }
  Form factor() {
    skip();
    if (p >= end) fail("unexpected end of expression");
    if (*p == '(') {
      ++p;
      Form f = expr();
      skip();
      if (p >= end || *p != ')') fail("')' expected");
      ++p;
      return f;
    }
    if (isdigit((unsigned char)*p)) {  // a bare constant must be zero (:1510-1518)
      if (integer() != 0) fail("non-zero constant term");
      return Form();
    }
    if (!isalpha((unsigned char)*p)) fail(std::string("unexpected character '") + *p + "'");
    const char* s = p;
    while (p < end && (isalnum((unsigned char)*p) || *p == '_')) ++p;
    const std::string name(s, p);
    auto it = vars.find(name);
    if (it != vars.end()) return it->second;
    // unknown name: an input (:1519-1527); its column is its number
    size_t d = 0;
    while (d < name.size() && !isdigit((unsigned char)name[d])) ++d;
    if (d == name.size()) fail("input without index: " + name);
    const int j = std::stoi(name.substr(d));
    if (j > max_input) max_input = j;
    return Form{{j, Rat(1)}};
  }
  Form term() {
    Form f = factor();
    for (;;) {
      skip();
      if (p < end && (*p == '*' || *p == '/')) {
        const char op = *p++;
        const long long v = integer();
        if (v == 0 && op == '/') fail("division by zero");
        const Rat c = op == '*' ? Rat((int64_t)v) : Rat::make(1, v);
        for (auto& e : f) e.second = Q.mul(e.second, c);
        if (v == 0) f.clear();
      } else break;
    }
    return f;
  }
  Form expr() {
    Form acc;
    skip();
    int sign = 1;
    if (p < end && (*p == '+' || *p == '-')) { sign = *p == '-' ? -1 : 1; ++p; }
    for (;;) {
      const Form t = term();
      axpy(acc, t, Rat(sign), Q);
      skip();
      if (p < end && (*p == '+' || *p == '-')) { sign = *p == '-' ? -1 : 1; ++p; }
      else break;
    }
    return acc;
  }

 public:
  // Evaluates the whole program text; returns the matrix of the `outchar` variables.
  SparseRows build(const std::string& text, char outchar = 'o') {
    vars.clear();
    max_input = -1;
    size_t pos = 0;
    while (pos < text.size()) {
      size_t eol = text.find('\n', pos);
      if (eol == std::string::npos) eol = text.size();
      std::string line = text.substr(pos, eol - pos);
      pos = eol + 1;
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.resize(hash);
      size_t b = line.find_first_not_of(" \t\r");
      if (b == std::string::npos) continue;
      // a line may hold several `x:=...;` statements
      size_t s0 = b;
      while (s0 < line.size()) {
        size_t semi = line.find(';', s0);
        if (semi == std::string::npos) semi = line.size();
        const std::string st = line.substr(s0, semi - s0);
        s0 = semi + 1;
        const size_t asg = st.find(":=");
        if (asg == std::string::npos) { if (st.find_first_not_of(" \t\r") != std::string::npos) fail("statement without ':=': " + st); continue; }
        std::string name = st.substr(0, asg);
        name.erase(0, name.find_first_not_of(" \t"));
        name.erase(name.find_last_not_of(" \t") + 1);
        if (name.empty()) fail("empty left-hand side");
        cur_output = name;
        p = st.data() + asg + 2;
        end = st.data() + st.size();
        Form f = expr();
        skip();
        if (p != end) fail("trailing characters in: " + st);
        vars[name].swap(f);
      }
    }
    SparseRows M;
    int max_out = -1;
    for (const auto& kv : vars)
      if (kv.first[0] == outchar && kv.first.size() > 1 && isdigit((unsigned char)kv.first[1])) {
        const int i = std::stoi(kv.first.substr(1));
        if (i > max_out) max_out = i;
      }
    M.rows = (size_t)(max_out + 1);
    M.cols = (size_t)(max_input + 1);
    M.r.assign(M.rows, {});
    for (const auto& kv : vars)
      if (kv.first[0] == outchar && kv.first.size() > 1 && isdigit((unsigned char)kv.first[1]))
        M.r[(size_t)std::stoi(kv.first.substr(1))] = kv.second;
    return M;
  }
};

}  // namespace host
}  // namespace plo
