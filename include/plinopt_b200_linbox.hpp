// plinopt_b200_linbox.hpp -- the reference-side adapter: what a PLinOpt maintainer includes to bind libplinopt_b200.so from
// the reference's template code (include/plinopt_sparsify.inl, src/orbiter.cpp, include/plinopt_library.inl).
//
// The reference works on LinBox::SparseMatrix<Field, SparseSeq> (rows = std::vector<std::pair<size_t, Element>>, accessed as
// .first/.second, include/plinopt_library.inl:44-47) over Givaro::QField<Rational> or Givaro::Modular<Integer>.  Everything here
// is written against that small surface only -- rowdim(), coldim(), field(), operator[](i) -- so it also compiles against the
// 50-line mock of tests/linbox_mock.hpp (which is how this repository compile-tests it: LinBox is not installed here).
// The two Givaro element conversions are the only LinBox-specific lines; they sit under PLINOPT_HAVE_LINBOX.
//
//   plo::adapter::scale_columns / scale_vector / scale_rows   exact images for plo_lincomb_search / plo_lincomb_quad
//   plo::adapter::quad_rows                                   replaces the num + i,j,k,l loops of localSparsifier (:288-326)
//   plo::adapter::scale_to_int32, orbit_sweep                 replaces the omp loop of Orbiter::operator() (src/orbiter.cpp:272-324)
//   plo::adapter::to_csr, mmcheck                             replaces PLinOpt::MMchecker's evaluation (plinopt_library.inl:504-528)
#ifndef PLINOPT_B200_LINBOX_HPP
#define PLINOPT_B200_LINBOX_HPP

#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "plinopt_b200.h"

namespace plo {
namespace adapter {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

// ---- field traits: characteristic (0 over Q) and element -> (numerator, denominator) on machine words ----------------------
// Specialise for other fields; a value that does not fit int64 must throw (never wrap).
template <class Field>
struct FieldTraits;  // { static uint32_t characteristic(const Field&); static void num_den(const Field&, const Element&, int64_t&, int64_t&); }

#ifdef PLINOPT_HAVE_LINBOX
}  // namespace adapter
}  // namespace plo
#include <givaro/modular-integer.h>
#include <givaro/qfield.h>
namespace plo {
namespace adapter {
inline int64_t to_i64(const Givaro::Integer& v) {
  if (v.bitsize() > 62) throw Error(PLO_E_RANGE, "entry exceeds 62 bits");
  return (int64_t)v;
}
template <>
struct FieldTraits<Givaro::QField<Givaro::Rational>> {
  static uint32_t characteristic(const Givaro::QField<Givaro::Rational>&) { return 0; }
  static void num_den(const Givaro::QField<Givaro::Rational>&, const Givaro::Rational& e, int64_t& n, int64_t& d) { n = to_i64(e.nume()); d = to_i64(e.deno()); }
};
template <>
struct FieldTraits<Givaro::Modular<Givaro::Integer>> {
  typedef Givaro::Modular<Givaro::Integer> F;
  static uint32_t characteristic(const F& f) {
    if (f.characteristic() >= (Givaro::Integer(1) << 32)) throw Error(PLO_E_RANGE, "modulus beyond 32 bits");
    return (uint32_t)(uint64_t)f.characteristic();
  }
  // un-reduced representatives (the raw -r of augment(), plinopt_sparsify.inl:20-35) are reduced here: the engine takes canonical residues
  static void num_den(const F& f, const Givaro::Integer& e, int64_t& n, int64_t& d) { Givaro::Integer r(e % f.characteristic()); if (r < 0) r += f.characteristic(); n = to_i64(r); d = 1; }
};
#endif

inline int64_t gcd_i64(int64_t a, int64_t b) { a = a < 0 ? -a : a; b = b < 0 ? -b : b; while (b) { const int64_t t = a % b; a = b; b = t; } return a; }
inline int64_t lcm_i64(int64_t a, int64_t b) {
  const __int128 l = (__int128)(a / gcd_i64(a, b)) * b;
  if (l > (__int128)INT64_MAX) throw Error(PLO_E_RANGE, "common denominator exceeds 64 bits");
  return (int64_t)l;
}
inline int64_t scaled_i64(int64_t n, int64_t d, int64_t lcd) {
  const __int128 v = (__int128)n * (lcd / d);
  if (v > (__int128)INT64_MAX || v < -(__int128)INT64_MAX) throw Error(PLO_E_RANGE, "scaled entry exceeds 64 bits");
  return (int64_t)v;
}

template <class Field>
uint32_t characteristic_or_0(const Field& F) { return FieldTraits<Field>::characteristic(F); }

// dense (num, den) image of a sparse matrix, row-major
template <class Mat>
void dense_num_den(const Mat& M, std::vector<int64_t>& num, std::vector<int64_t>& den) {
  typedef FieldTraits<typename Mat::Field> T;
  const size_t r = M.rowdim(), c = M.coldim();
  num.assign(r * c, 0); den.assign(r * c, 1);
  for (size_t i = 0; i < r; ++i)
    for (const auto& e : M[i]) T::num_den(M.field(), e.second, num[i * c + e.first], den[i * c + e.first]);
}

// TM (n x m): over Q every COLUMN times its LCD (the zero pattern of TM^T.w does not change; rows must not be scaled one by one);
// over Z/pZ the canonical residues.
template <class Mat>
std::vector<int64_t> scale_columns(const Mat& TM) {
  std::vector<int64_t> num, den;
  dense_num_den(TM, num, den);
  const size_t n = TM.rowdim(), m = TM.coldim();
  std::vector<int64_t> out(n * m, 0);
  for (size_t j = 0; j < m; ++j) {
    int64_t l = 1;
    for (size_t i = 0; i < n; ++i) l = lcm_i64(l, den[i * m + j]);
    for (size_t i = 0; i < n; ++i) out[i * m + j] = scaled_i64(num[i * m + j], den[i * m + j], l);
  }
  return out;
}
// a coefficient list / one vector times its common LCD
template <class Field, class Vec>
std::vector<int64_t> scale_vector(const Field& F, const Vec& v) {
  std::vector<int64_t> num(v.size()), den(v.size());
  for (size_t i = 0; i < v.size(); ++i) FieldTraits<Field>::num_den(F, v[i], num[i], den[i]);
  int64_t l = 1;
  for (int64_t d : den) l = lcm_i64(l, d);
  for (size_t i = 0; i < v.size(); ++i) num[i] = scaled_i64(num[i], den[i], l);
  return num;
}
// the first `nrows` rows of LCoB (n x n), each times its own LCD
template <class Mat>
std::vector<int64_t> scale_rows(const Mat& LCoB, size_t nrows) {
  std::vector<int64_t> num, den;
  dense_num_den(LCoB, num, den);
  const size_t n = LCoB.coldim();
  std::vector<int64_t> out(nrows * n, 0);
  for (size_t i = 0; i < nrows; ++i) {
    int64_t l = 1;
    for (size_t j = 0; j < n; ++j) l = lcm_i64(l, den[i * n + j]);
    for (size_t j = 0; j < n; ++j) out[i * n + j] = scaled_i64(num[i * n + j], den[i * n + j], l);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// localSparsifier, one inner block: replaces the `for num` loop with its quad loop (include/plinopt_sparsify.inl:288-314).
// On return rows[t] holds, for every row the device decided, the four coefficient indices (i,j,k,l) of the accepted candidate
// (index PLO_NO_INDEX: the nullspace vector keeps row 0) and the weight; the caller does setRow(LCoB, offsetblock+t, w) exactly as
// testLinComb would have (:190-193).  status tells how to go on (PLO_QUAD_MISS: canonical fallback :317-326 for row nrows, then call
// again; PLO_QUAD_SEED / PLO_QUAD_RANGE: see plinopt_b200.h).
// ---------------------------------------------------------------------------------------------------------------------
struct QuadRows {
  int nrows = 0, status = PLO_QUAD_DONE;
  int rl[4] = {-1, -1, -1, -1}, cl[4] = {-1, -1, -1, -1};
  uint64_t index[4] = {PLO_NO_INDEX, PLO_NO_INDEX, PLO_NO_INDEX, PLO_NO_INDEX};
  size_t coef[4][4] = {};  // coef[t] = (i, j, k, l) of row t
};
template <class Mat, class Vec>
QuadRows quad_rows(const Mat& TM, const Vec& Coeffs, const Mat& LCoB, size_t offsetblock, size_t rows_known, int init_rl, int init_cl) {
  const auto& F = TM.field();
  const std::vector<int64_t> tm = scale_columns(TM), cf = scale_vector(F, Coeffs), prev = scale_rows(LCoB, rows_known);
  std::vector<int64_t> seed;
  const bool has_seed = rows_known == 0 && !(init_rl == -1 && init_cl == -1);
  if (has_seed) seed = scale_rows(LCoB, 1);  // the nullspace vector sits in row 0 (:236-243)
  plo_quad_problem q;
  q.n = (int)TM.rowdim(); q.off = (int)offsetblock; q.c = (int)Coeffs.size(); q.nprev = (int)rows_known;
  q.TM = tm.data(); q.coeffs = cf.data(); q.prev_rows = rows_known ? prev.data() : nullptr; q.seed_vec = has_seed ? seed.data() : nullptr;
  q.init_rl = init_rl; q.init_cl = init_cl;
  const int rc = plo_lincomb_quad(characteristic_or_0(F), (int)TM.coldim(), 1, &q);
  if (rc != PLO_OK) throw Error(rc, plo_last_error());
  QuadRows out;
  out.nrows = q.nrows; out.status = q.status;
  const uint64_t c = (uint64_t)Coeffs.size();
  for (int t = 0; t < q.nrows; ++t) {
    out.rl[t] = q.rl[t]; out.cl[t] = q.cl[t]; out.index[t] = q.index[t];
    if (q.index[t] != PLO_NO_INDEX) {
      out.coef[t][0] = (size_t)(q.index[t] / (c * c * c)); out.coef[t][1] = (size_t)((q.index[t] / (c * c)) % c);
      out.coef[t][2] = (size_t)((q.index[t] / c) % c); out.coef[t][3] = (size_t)(q.index[t] % c);
    }
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// Orbiter: integer images of L, R, P (every entry times the matrix' common LCD) and the sweep (src/orbiter.cpp:272-324).
// ---------------------------------------------------------------------------------------------------------------------
template <class Mat>
std::vector<int32_t> scale_to_int32(const Mat& M, int32_t& lcd) {
  std::vector<int64_t> num, den;
  dense_num_den(M, num, den);
  int64_t l = 1;
  for (int64_t d : den) l = lcm_i64(l, d);
  if (l > INT32_MAX) throw Error(PLO_E_RANGE, "common denominator beyond the int32 interface (use plo_orbit_sweep64)");
  std::vector<int32_t> out(num.size());
  for (size_t e = 0; e < num.size(); ++e) {
    const int64_t v = scaled_i64(num[e], den[e], l);
    if (v > INT32_MAX || v < -INT32_MAX) throw Error(PLO_E_RANGE, "scaled entry beyond the int32 interface (use plo_orbit_sweep64)");
    out[e] = (int32_t)v;
  }
  lcd = (int32_t)l;
  return out;
}
// measure: PLO_MEASURE_NNZ (Orbiter<0>) or PLO_MEASURE_G2; candidates [0, randomloops) of the Philox enumeration of `seed`.
// The winner's (U, V, W) come back as {-1,0,1} matrices: apply them once with the reference's own Tensor + BMD.mul (:284-294).
template <class Mat>
plo_orbit_best orbit_sweep(const Mat& L, const Mat& R, const Mat& P, int measure, uint64_t seed, uint64_t randomloops,
                           std::vector<int32_t>& U, std::vector<int32_t>& V, std::vector<int32_t>& W) {
  int m, k, n;
  plo_LRP2MM((int)L.coldim(), (int)R.coldim(), (int)P.rowdim(), &m, &k, &n);
  int32_t dL, dR, dP;
  const std::vector<int32_t> Li = scale_to_int32(L, dL), Ri = scale_to_int32(R, dR), Pi = scale_to_int32(P, dP);
  plo_orbit_best best;
  const int rc = plo_orbit_sweep(0, m, k, n, (int)L.rowdim(), Li.data(), Ri.data(), Pi.data(), dL, dR, dP, measure, PLO_MODE_PHILOX, seed, 0, randomloops, &best);
  if (rc != PLO_OK) throw Error(rc, plo_last_error());
  U.assign((size_t)m * m, 0); V.assign((size_t)k * k, 0); W.assign((size_t)n * n, 0);
  if (best.index != PLO_NO_INDEX) plo_orbit_decode(m, k, n, PLO_MODE_PHILOX, seed, best.index, U.data(), V.data(), W.data());
  return best;
}

// ---------------------------------------------------------------------------------------------------------------------
// MMchecker over Z/pZ (after the rebind of src/MMchecker.cpp:61-63): CSR images and the batched check.
// ---------------------------------------------------------------------------------------------------------------------
struct Csr {
  std::vector<int64_t> ptr;
  std::vector<int32_t> col;
  std::vector<uint32_t> val;
  int rows = 0, cols = 0;
  plo_csr view() const { plo_csr c; c.rows = rows; c.cols = cols; c.ptr = ptr.data(); c.col = col.data(); c.val = val.data(); return c; }
};
template <class Mat>
Csr to_csr(const Mat& M) {
  typedef FieldTraits<typename Mat::Field> T;
  Csr out;
  out.rows = (int)M.rowdim(); out.cols = (int)M.coldim();
  out.ptr.assign(1, 0);
  for (size_t i = 0; i < M.rowdim(); ++i) {
    for (const auto& e : M[i]) {
      int64_t n, d;
      T::num_den(M.field(), e.second, n, d);
      if (n) { out.col.push_back((int32_t)e.first); out.val.push_back((uint32_t)n); }
    }
    out.ptr.push_back((int64_t)out.col.size());
  }
  return out;
}
// 0 correct / 1 not an MM algorithm (plinopt_library.inl:555) / 3 outer dimension mismatch (:494)
template <class Mat>
int mmcheck(const Mat& L, const Mat& R, const Mat& P, uint64_t seed, int batch) {
  int m, k, n;
  plo_LRP2MM((int)L.coldim(), (int)R.coldim(), (int)P.rowdim(), &m, &k, &n);
  const Csr cl = to_csr(L), cr = to_csr(R), cp = to_csr(P);
  const plo_csr vl = cl.view(), vr = cr.view(), vp = cp.view();
  std::vector<uint8_t> ok((size_t)batch);
  const int rc = plo_mmcheck_batch(characteristic_or_0(L.field()), m, k, n, (int)L.rowdim(), &vl, &vr, &vp, seed, batch, nullptr, nullptr, ok.data());
  if (rc < 0) throw Error(rc, plo_last_error());
  return rc;
}

}  // namespace adapter
}  // namespace plo
#endif  // PLINOPT_B200_LINBOX_HPP
