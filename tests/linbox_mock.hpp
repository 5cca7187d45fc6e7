// linbox_mock.hpp -- the part of LinBox::SparseMatrix<Field, SparseSeq> that include/plinopt_b200_linbox.hpp touches, restated
// in 50 lines so that the adapter is compile- and run-tested in a tree without LinBox: rows are std::vector<std::pair<size_t,
// Element>> reachable through operator[] (the reference accesses them as .first/.second, include/plinopt_library.inl:44-47),
// plus rowdim(), coldim(), field() and setEntry().  Two mock fields: rationals as (num, den) pairs and Z/pZ on int64.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace mock {

struct Rational { int64_t n = 0, d = 1; };
struct QField { typedef Rational Element; };
struct ZpField { typedef int64_t Element; int64_t p; };

template <class F>
class SparseMatrix {
 public:
  typedef F Field;
  typedef typename F::Element Element;
  typedef std::vector<std::pair<size_t, Element>> Row;
  SparseMatrix(const F& f, size_t r, size_t c) : f_(f), cols_(c), rows_(r) {}
  size_t rowdim() const { return rows_.size(); }
  size_t coldim() const { return cols_; }
  const F& field() const { return f_; }
  const Row& operator[](size_t i) const { return rows_[i]; }
  void setEntry(size_t i, size_t j, const Element& e) {
    Row& row = rows_[i];
    auto it = row.begin();
    while (it != row.end() && it->first < j) ++it;
    if (it != row.end() && it->first == j) it->second = e; else row.insert(it, std::make_pair(j, e));
  }
 private:
  F f_;
  size_t cols_;
  std::vector<Row> rows_;
};

}  // namespace mock

#include "../include/plinopt_b200_linbox.hpp"
namespace plo {
namespace adapter {
template <>
struct FieldTraits<mock::QField> {
  static uint32_t characteristic(const mock::QField&) { return 0; }
  static void num_den(const mock::QField&, const mock::Rational& e, int64_t& n, int64_t& d) { n = e.n; d = e.d; }
};
template <>
struct FieldTraits<mock::ZpField> {
  static uint32_t characteristic(const mock::ZpField& f) { return (uint32_t)f.p; }
  static void num_den(const mock::ZpField& f, const int64_t& e, int64_t& n, int64_t& d) { n = ((e % f.p) + f.p) % f.p; d = 1; }
};
}  // namespace adapter
}  // namespace plo
