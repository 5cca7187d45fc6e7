// exact.hpp -- host-side exact arithmetic of the engine: rationals on machine
// words (stand-in for Givaro::Rational / QField, include/plinopt_library.h:52-60)
// and Z/pZ (Givaro::Modular<Integer>, src/sparsifier.cpp:71-76), plus the small
// dense linear algebra the host layer needs around the GPU sweeps.
// A rational leaving the int64 range throws plo::host::RangeError, which the
// C ABI turns into PLO_E_RANGE (never a silent wrap-around).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

namespace plo {
namespace host {

struct RangeError : std::runtime_error {
  explicit RangeError(const std::string& s) : std::runtime_error(s) {}
};

typedef __int128 wide;

inline wide wabs(wide a) { return a < 0 ? -a : a; }
inline uint64_t gcd64(uint64_t a, uint64_t b) {  // binary gcd: shifts and subtractions only
  if (a == 0) return b;
  if (b == 0) return a;
  const int sh = __builtin_ctzll(a | b);
  a >>= __builtin_ctzll(a);
  do {
    b >>= __builtin_ctzll(b);
    if (a > b) { const uint64_t t = a; a = b; b = t; }
    b -= a;
  } while (b != 0);
  return a << sh;
}
inline wide wgcd(wide a, wide b) {
  a = wabs(a); b = wabs(b);
  if ((a >> 64) == 0 && (b >> 64) == 0) return (wide)gcd64((uint64_t)a, (uint64_t)b);  // the usual case: no 128-bit division
  while (b != 0) { wide t = a % b; a = b; b = t; }
  return a;
}

struct Rat {
  int64_t num, den;  // den > 0, lowest terms
  Rat() : num(0), den(1) {}
  Rat(int64_t n) : num(n), den(1) {}
  Rat(int64_t n, int64_t d) { *this = make(n, d); }
  static Rat make(wide n, wide d) {
    if (d == 0) throw RangeError("rational with zero denominator");
    if (d < 0) { n = -n; d = -d; }
    Rat r;
    if (n == 0) return r;
    const wide an = wabs(n);
    if ((an >> 63) == 0 && (d >> 63) == 0) {  // both fit 63 bits: 64-bit arithmetic throughout
      const uint64_t g = gcd64((uint64_t)an, (uint64_t)d);
      const int64_t q = (int64_t)((uint64_t)an / g);
      r.num = n < 0 ? -q : q;
      r.den = (int64_t)((uint64_t)d / g);
      return r;
    }
    const wide g = wgcd(n, d);
    n /= g; d /= g;
    if (wabs(n) > (wide)INT64_MAX || d > (wide)INT64_MAX) throw RangeError("rational exceeds 64 bits");
    r.num = (int64_t)n; r.den = (int64_t)d;
    return r;
  }
  bool operator==(const Rat& o) const { return num == o.num && den == o.den; }
  bool operator!=(const Rat& o) const { return !(*this == o); }
  bool operator<(const Rat& o) const { return (wide)num * o.den < (wide)o.num * den; }
};

// Field over Q
struct QField {
  typedef Rat Elt;
  static constexpr bool modular = false;
  uint64_t characteristic() const { return 0; }
  Elt zero() const { return Rat(); }
  Elt one() const { return Rat(1); }
  Elt mone() const { return Rat(-1); }
  Elt from_int(int64_t i) const { return Rat(i); }
  Elt from_ratio(int64_t n, int64_t d) const { return Rat::make(n, d); }
  static Elt integer(wide v) {
    if (wabs(v) > (wide)INT64_MAX) throw RangeError("rational exceeds 64 bits");
    Rat r; r.num = (int64_t)v; return r;
  }
  Elt add(const Elt& a, const Elt& b) const {
    if (a.den == 1 && b.den == 1) return integer((wide)a.num + b.num);
    return Rat::make((wide)a.num * b.den + (wide)b.num * a.den, (wide)a.den * b.den);
  }
  Elt sub(const Elt& a, const Elt& b) const {
    if (a.den == 1 && b.den == 1) return integer((wide)a.num - b.num);
    return Rat::make((wide)a.num * b.den - (wide)b.num * a.den, (wide)a.den * b.den);
  }
  Elt mul(const Elt& a, const Elt& b) const {
    if (a.den == 1 && b.den == 1) return integer((wide)a.num * b.num);
    return Rat::make((wide)a.num * b.num, (wide)a.den * b.den);
  }
  Elt div(const Elt& a, const Elt& b) const { return Rat::make((wide)a.num * b.den, (wide)a.den * b.num); }
  Elt neg(const Elt& a) const { Rat r; r.num = -a.num; r.den = a.den; return r; }
  Elt raw_neg(const Elt& a) const { return neg(a); }
  Elt inv(const Elt& a) const { return Rat::make(a.den, a.num); }
  bool is_zero(const Elt& a) const { return a.num == 0; }
  bool is_one(const Elt& a) const { return a.num == 1 && a.den == 1; }
  bool is_mone(const Elt& a) const { return a.num == -1 && a.den == 1; }
  bool same_rep(const Elt& a, const Elt& b) const { return a == b; }
  bool less(const Elt& a, const Elt& b) const { return a < b; }
};

// Field Z/pZ.  Elements are int64 representatives; field operations return the
// canonical representative in [0,p) whereas raw_neg / from_int keep the
// un-reduced integer, mirroring what the reference does with Givaro::Integer
// elements.
struct ZpField {
  typedef int64_t Elt;
  static constexpr bool modular = true;
  int64_t p;
  explicit ZpField(int64_t p_) : p(p_) {}
  uint64_t characteristic() const { return (uint64_t)p; }
  Elt canon(Elt a) const {
    if ((uint64_t)a < (uint64_t)p) return a;  // already the canonical representative (the usual case)
    a %= p;
    return a < 0 ? a + p : a;
  }
  Elt zero() const { return 0; }
  Elt one() const { return 1 % p; }
  Elt mone() const { return canon(-1); }
  Elt from_int(int64_t i) const { return i; }
  Elt from_ratio(int64_t n, int64_t d) const { return div(canon(n), canon(d)); }
  Elt add(Elt a, Elt b) const { const uint64_t s = (uint64_t)canon(a) + (uint64_t)canon(b); return (Elt)(s >= (uint64_t)p ? s - (uint64_t)p : s); }  // p < 2^63
  Elt sub(Elt a, Elt b) const { const uint64_t x = (uint64_t)canon(a), y = (uint64_t)canon(b); return (Elt)(x >= y ? x - y : x + (uint64_t)p - y); }
  Elt mul(Elt a, Elt b) const {
    const uint64_t x = (uint64_t)canon(a), y = (uint64_t)canon(b);
    if ((uint64_t)p <= 0xFFFFFFFFull) return (Elt)((x * y) % (uint64_t)p);  // word-size moduli: one 64-bit division
    return (Elt)(((unsigned __int128)x * y) % (uint64_t)p);
  }
  Elt neg(Elt a) const { a = canon(a); return a ? p - a : 0; }
  Elt raw_neg(Elt a) const { return -a; }
  Elt inv(Elt a) const {
    int64_t r0 = p, r1 = canon(a), t0 = 0, t1 = 1;
    if (r1 == 0) throw RangeError("inverse of zero mod p");
    while (r1) { const int64_t q = r0 / r1; int64_t t = r0 - q * r1; r0 = r1; r1 = t; t = t0 - q * t1; t0 = t1; t1 = t; }
    if (r0 != 1) throw RangeError("non-invertible element mod p");
    return canon(t0);
  }
  Elt div(Elt a, Elt b) const { return mul(a, inv(b)); }
  bool is_zero(Elt a) const { return canon(a) == 0; }
  bool is_one(Elt a) const { return canon(a) == one(); }
  bool is_mone(Elt a) const { return canon(a) == mone(); }
  bool same_rep(Elt a, Elt b) const { return a == b; }
  bool less(Elt a, Elt b) const { return a < b; }
};

// Dense row-major matrix over a field.
template <class F>
struct Dense {
  typedef typename F::Elt Elt;
  size_t rows, cols;
  std::vector<Elt> v;
  Dense() : rows(0), cols(0) {}
  Dense(const F& f, size_t r, size_t c) : rows(r), cols(c), v(r * c, f.zero()) {}
  Elt& at(size_t i, size_t j) { return v[i * cols + j]; }
  const Elt& at(size_t i, size_t j) const { return v[i * cols + j]; }
};

// Reduced row echelon form in place; returns the pivot columns.
template <class F>
std::vector<size_t> rref(const F& f, Dense<F>& A) {
  std::vector<size_t> piv;
  size_t row = 0;
  for (size_t col = 0; col < A.cols && row < A.rows; ++col) {
    size_t sel = A.rows;
    for (size_t i = row; i < A.rows; ++i)
      if (!f.is_zero(A.at(i, col))) { sel = i; break; }
    if (sel == A.rows) continue;
    if (sel != row)
      for (size_t j = 0; j < A.cols; ++j) std::swap(A.at(sel, j), A.at(row, j));
    const typename F::Elt ip = f.inv(A.at(row, col));
    for (size_t j = 0; j < A.cols; ++j) A.at(row, j) = f.mul(A.at(row, j), ip);
    for (size_t i = 0; i < A.rows; ++i) {
      if (i == row || f.is_zero(A.at(i, col))) continue;
      const typename F::Elt mlt = A.at(i, col);
      for (size_t j = 0; j < A.cols; ++j) A.at(i, j) = f.sub(A.at(i, j), f.mul(mlt, A.at(row, j)));
    }
    piv.push_back(col);
    ++row;
  }
  return piv;
}

// Basis of the right nullspace {x : A x = 0}, one vector per free column.
template <class F>
std::vector<std::vector<typename F::Elt>> nullspace(const F& f, const Dense<F>& A0, size_t* rank_out = nullptr) {
  Dense<F> A = A0;
  const std::vector<size_t> piv = rref(f, A);
  if (rank_out) *rank_out = piv.size();
  std::vector<char> isp(A.cols, 0);
  for (size_t c : piv) isp[c] = 1;
  std::vector<std::vector<typename F::Elt>> basis;
  for (size_t fc = 0; fc < A.cols; ++fc) {
    if (isp[fc]) continue;
    std::vector<typename F::Elt> x(A.cols, f.zero());
    x[fc] = f.one();
    for (size_t k = 0; k < piv.size(); ++k) x[piv[k]] = f.neg(A.at(k, fc));
    basis.push_back(x);
  }
  return basis;
}

}  // namespace host
}  // namespace plo
