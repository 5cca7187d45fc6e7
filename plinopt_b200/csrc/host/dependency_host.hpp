// dependency_host.hpp -- host side of `dependency`:
// Depender  src/dependency.cpp:106-169  around the GPU enumeration plo_dependency_explore
// (dependency_explore.cu): coefficient list (:118-146), product tables, and the text form of
// a hit (showOut / showLC :47-71).
#pragma once
#include <sstream>

#include "sparsify_host.hpp"

namespace plo {
namespace host {

// v is augmented by r, -r, 1/r, -1/r if r is new  (plinopt_sparsify.inl:20-35, over Q)
inline void augment_q(std::vector<Rat>& v, const Rat& r) {
  QField Q;
  if (std::find(v.begin(), v.end(), r) != v.end()) return;
  v.push_back(r);
  v.push_back(Q.neg(r));
  const Rat t = Q.inv(r);
  v.push_back(t);
  v.push_back(Q.neg(t));
}

// src/dependency.cpp:118-131: {1,-1} + user list, numerators and denominators of the stored
// entries (row-major), then 2,3,.. up to maxnumcoeff; truncated.
inline std::vector<Rat> dependency_coefficients(const Dense<QField>& B, const std::vector<Rat>& user, size_t maxnumcoeff) {
  std::vector<Rat> C{Rat(1), Rat(-1)};
  C.insert(C.end(), user.begin(), user.end());
  for (size_t e = 0; e < B.v.size(); ++e) {
    if (B.v[e].num == 0) continue;
    augment_q(C, Rat(B.v[e].num));
    augment_q(C, Rat(B.v[e].den));
  }
  for (int64_t i = 2; C.size() < maxnumcoeff; ++i) augment_q(C, Rat(i));
  if (C.size() > maxnumcoeff) C.resize(maxnumcoeff);
  return C;
}

// :133-143 images in the field, zero and repeated values dropped
template <class F>
std::vector<typename F::Elt> field_coefficients(const F& f, const std::vector<Rat>& C) {
  std::vector<typename F::Elt> out;
  for (const Rat& e : C) {
    typename F::Elt x;
    try { x = f.from_ratio(e.num, e.den); } catch (const RangeError&) { continue; }  // denominator not invertible mod p
    if (f.is_zero(x)) continue;
    bool seen = false;
    for (const auto& y : out) if (f.same_rep(f.add(y, f.zero()), f.add(x, f.zero()))) { seen = true; break; }
    if (!seen) out.push_back(x);
  }
  return out;
}

// showOut :47-66
inline void show_out(std::ostream& out, const QField&, char v, size_t i, const Rat& r) {
  out << (r.num < 0 ? '-' : '+') << v << i;
  const int64_t an = r.num < 0 ? -r.num : r.num;
  if (!(an == 1 && r.den == 1)) {
    if (an == 1) out << '/' << r.den;
    else { out << '*' << an; if (r.den != 1) out << '/' << r.den; }
  }
}
inline void show_out(std::ostream& out, const ZpField& F, char v, size_t i, const int64_t& e0) {
  const int64_t e = F.canon(e0), a = F.neg(e);  // Fsign / Fabs, plinopt_library.h:208-224
  out << (a < e ? '-' : '+') << v << i;
  if (!(F.is_one(e) || F.is_mone(e))) out << '*' << (a < e ? a : e);
}

template <class F>
struct DependencyHost {
  typedef typename F::Elt Elt;
  const F& f;
  explicit DependencyHost(const F& field) : f(field) {}

  // one output line of the reference for a hit (:84-90): [-W[pos] i<pos>] then the combination
  std::string line(const Dense<F>& M, const std::vector<Elt>& C, const plo_dep_hit& h) const {
    std::ostringstream out;
    if (h.pos >= 0) {
      Elt w = M.at((size_t)h.rows[0], (size_t)h.pos);
      for (int t = 1; t <= h.depth; ++t) w = f.add(w, f.mul(C[(size_t)h.coefs[t]], M.at((size_t)h.rows[t], (size_t)h.pos)));
      show_out(out, f, 'i', (size_t)h.pos, f.neg(w));
    }
    show_out(out, f, 'o', (size_t)h.rows[0], f.one());
    for (int t = 1; t <= h.depth; ++t) show_out(out, f, 'o', (size_t)h.rows[t], C[(size_t)h.coefs[t]]);
    out << ';';
    return out.str();
  }
};

}  // namespace host
}  // namespace plo
