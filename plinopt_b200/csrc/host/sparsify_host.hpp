// sparsify_host.hpp -- host orchestration of the alternate-basis sparsifier around the GPU candidate search.
// Mirrors the reference's entry points of include/plinopt_sparsify.inl (names, argument meaning, log prefixes):
//   augment :20-35 · rank :38-45 · localSparsifier :205-347 · FactorDiagonals :354-375 · inverse/inverseTranspose :380-465 ·
//   SparseFactor :473-513 · sparseLU :523-568 · sparseILU :576-604 · sparseAlternate :609-661 · blockSparsifier :666-748 ·
//   consistency :871-907
//
// How the work is organised here (the reference is strictly sequential):
//   * the quad loop + testLinComb (:299-314, :166-197) runs on the GPU -- all the rows of an inner block in ONE launch sequence
//     (plo_lincomb_quad, lincomb_quad.cu); wide outputs (m > 64) keep the one-row entry point plo_lincomb_search;
//   * the independent column blocks of blockSparsifier (:710-723) advance in LOCK STEP: sparseAlternate is a resumable state
//     machine (AlternateRun) that stops whenever it needs a search; the driver collects the pending searches of every block and
//     submits them as one batch, so a whole run costs a handful of device round trips instead of one per (block, num) step;
//   * eliminations work on sparse rows (column-sorted entry lists with running column counts), so the initial LU of
//     blockSparsifier scales to wide inputs (32x32x32_15096: 1024 x 15096).
// Progress lines of a column block are buffered and emitted in block order: the log reads like the reference's.
//
// LinBox's GaussDomain::QLUPin / nullspacebasisin are not in the reference tree; the pivot rule used instead is documented in
// DESIGN.md ("pivot rule") -- parity unpinned upstream.
#pragma once
#include <algorithm>
#include <chrono>
#include <map>
#include <memory>
#include <numeric>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/plinopt_b200.h"
#include "exact.hpp"

#ifndef COEFFICIENT_SEARCH
#define COEFFICIENT_SEARCH 11u  // include/plinopt_sparsify.h:36-38
#endif

namespace plo {
namespace host {

struct EngineError : std::runtime_error {
  int code;
  EngineError(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};

// statistics of the GPU part of a run: candidate evaluations of the reference loop covered (c^4 per row decided),
// device round trips (batched quad calls + single-row searches), canonical fallbacks
struct SearchStats {
  unsigned long long candidates = 0, searches = 0, fallbacks = 0;
  double device_seconds = 0;  // wall time inside the device calls (copies, launches, synchronisation)
  double t_begin_local = 0, t_finish_local = 0, t_begin_alt = 0;  // host phases (PLO_TIMING)
};

template <class F>
struct Sparsifier {
  typedef typename F::Elt Elt;
  typedef Dense<F> Mat;
  typedef std::vector<std::pair<size_t, Elt>> SRow;  // one sparse row: (column, value), columns increasing
  const F& f;
  std::ostream* log;  // may be null
  SearchStats stats;

  Sparsifier(const F& field, std::ostream* logstream) : f(field), log(logstream) {}

  // ---- small helpers ------------------------------------------------------------------
  size_t rowSize(const Mat& M, size_t i) const {
    size_t s = 0;
    for (size_t j = 0; j < M.cols; ++j) s += !f.is_zero(M.at(i, j));
    return s;
  }
  size_t density(const Mat& M) const {  // plinopt_library.inl:238-245
    size_t s = 0;
    for (const Elt& e : M.v) s += !f.is_zero(e);
    return s;
  }
  Mat transpose(const Mat& A) const {
    Mat T(f, A.cols, A.rows);
    for (size_t i = 0; i < A.rows; ++i) for (size_t j = 0; j < A.cols; ++j) T.at(j, i) = A.at(i, j);
    return T;
  }
  Mat identity(size_t n) const {
    Mat I(f, n, n);
    for (size_t i = 0; i < n; ++i) I.at(i, i) = f.one();
    return I;
  }
  Mat mul(const Mat& A, const Mat& B) const {
    Mat C(f, A.rows, B.cols);
    std::vector<SRow> Bs(B.rows);
    for (size_t t = 0; t < B.rows; ++t) Bs[t] = sparse_row(B, t);
    for (size_t i = 0; i < A.rows; ++i)
      for (size_t t = 0; t < A.cols; ++t) {
        const Elt& a = A.at(i, t);
        if (f.is_zero(a)) continue;
        for (const auto& e : Bs[t]) C.at(i, e.first) = f.add(C.at(i, e.first), f.mul(a, e.second));
      }
    return C;
  }
  SRow sparse_row(const Mat& M, size_t i) const {
    SRow r;
    for (size_t j = 0; j < M.cols; ++j) if (!f.is_zero(M.at(i, j))) r.emplace_back(j, M.at(i, j));
    return r;
  }
  size_t rank(const Mat& A) const {  // :38-45
    Mat U = A;
    return rref(f, U).size();
  }
  Mat inverse(const Mat& A) const {  // :380-414 (exact, hence unique)
    const size_t n = A.rows;
    Mat W(f, n, 2 * n);
    for (size_t i = 0; i < n; ++i) {
      for (size_t j = 0; j < n; ++j) W.at(i, j) = A.at(i, j);
      W.at(i, n + i) = f.one();
    }
    const std::vector<size_t> piv = rref(f, W);
    if (piv.size() != n || piv.back() != n - 1) throw RangeError("singular matrix in inverse()");
    Mat I(f, n, n);
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) I.at(i, j) = W.at(i, n + j);
    return I;
  }
  Mat inverseTranspose(const Mat& A) const { return transpose(inverse(A)); }  // :431-465

  std::ostream& densityProfile(std::ostream& out, size_t& ss, const Mat& M) const {  // :118-126
    ss = 0;
    for (size_t i = 0; i < M.rows; ++i) { const size_t s = rowSize(M, i); ss += s; out << s << ' '; }
    return out << '=' << ss;
  }

  // ---- elimination with the documented pivot rule (stand-in for QLUPin), on sparse rows -----------------------------------
  // Step k: the sparsest non-empty remaining row comes first (lowest position among ties); its pivot is the entry whose column
  // holds the fewest non-zeroes among the remaining rows (leftmost among ties); rows below are reduced.  A = Pr^T . L . U with L
  // unit lower triangular (strict part in `L`, row-wise) and Pr the row order `rowperm`.
  struct Elim {
    size_t rank = 0, rows = 0, cols = 0;
    std::vector<size_t> rowperm, pivcol;
    std::vector<SRow> U, L;
  };
  // r <- r - mult . piv   (both column-sorted); `colcnt` follows the fill-in and the cancellations
  void axpy_row(SRow& r, const Elt& mult, const SRow& piv, std::vector<size_t>& colcnt) const {
    SRow out;
    out.reserve(r.size() + piv.size());
    size_t a = 0, b = 0;
    while (a < r.size() || b < piv.size()) {
      if (b == piv.size() || (a < r.size() && r[a].first < piv[b].first)) { out.push_back(r[a++]); continue; }
      if (a == r.size() || piv[b].first < r[a].first) {
        out.emplace_back(piv[b].first, f.neg(f.mul(mult, piv[b].second)));
        ++colcnt[piv[b].first];
        ++b;
        continue;
      }
      const Elt v = f.sub(r[a].second, f.mul(mult, piv[b].second));
      if (f.is_zero(v)) --colcnt[r[a].first]; else out.emplace_back(r[a].first, v);
      ++a; ++b;
    }
    r.swap(out);
  }
  Elim eliminate(const Mat& A) const {
    Elim e;
    const size_t m = A.rows, n = A.cols;
    e.rows = m; e.cols = n;
    e.U.resize(m); e.L.resize(m);
    e.rowperm.resize(m);
    std::iota(e.rowperm.begin(), e.rowperm.end(), 0);
    std::vector<size_t> colcnt(n, 0);
    for (size_t i = 0; i < m; ++i) {
      e.U[i] = sparse_row(A, i);
      for (const auto& x : e.U[i]) ++colcnt[x.first];
    }
    for (size_t k = 0; k < m; ++k) {
      size_t best = m, bestsz = n + 1;
      for (size_t i = k; i < m; ++i) { const size_t s = e.U[i].size(); if (s > 0 && s < bestsz) { bestsz = s; best = i; } }
      if (best == m) break;
      if (best != k) { e.U[k].swap(e.U[best]); e.L[k].swap(e.L[best]); std::swap(e.rowperm[k], e.rowperm[best]); }
      const SRow& prow = e.U[k];
      size_t pidx = 0, pcsz = m + 1;
      for (size_t t = 0; t < prow.size(); ++t) if (colcnt[prow[t].first] < pcsz) { pcsz = colcnt[prow[t].first]; pidx = t; }
      const size_t pc = prow[pidx].first;
      e.pivcol.push_back(pc);
      const Elt ip = f.inv(prow[pidx].second);
      for (size_t i = k + 1; i < m; ++i) {
        SRow& r = e.U[i];
        auto it = std::lower_bound(r.begin(), r.end(), pc, [](const std::pair<size_t, Elt>& x, size_t c) { return x.first < c; });
        if (it == r.end() || it->first != pc) continue;
        const Elt mult = f.mul(it->second, ip);
        e.L[i].emplace_back(k, mult);
        axpy_row(r, mult, e.U[k], colcnt);
      }
      for (const auto& x : e.U[k]) --colcnt[x.first];  // row k leaves the remaining set
      ++e.rank;
    }
    return e;
  }
  Mat dense_of(const std::vector<SRow>& R, size_t rows, size_t cols) const {
    Mat M(f, rows, cols);
    for (size_t i = 0; i < rows; ++i) for (const auto& x : R[i]) M.at(i, x.first) = x.second;
    return M;
  }
  // first nullspace vector (stand-in for nullspacebasisin column 0, :235-239): the leftmost non-pivot column is set to one, the
  // pivot coordinates follow by back substitution through U
  bool nullspaceVector(const Mat& N, std::vector<Elt>& x) const {
    const size_t n = N.cols;
    const Elim e = eliminate(N);
    std::vector<char> isp(n, 0);
    for (size_t c : e.pivcol) isp[c] = 1;
    x.assign(n, f.zero());
    const size_t fc = (size_t)(std::find(isp.begin(), isp.end(), 0) - isp.begin());
    if (fc == n) return false;
    x[fc] = f.one();
    for (size_t kk = e.rank; kk-- > 0;) {
      Elt s = f.zero(), pv = f.one();
      for (const auto& t : e.U[kk]) {
        if (t.first == e.pivcol[kk]) pv = t.second;
        else if (!f.is_zero(x[t.first])) s = f.add(s, f.mul(t.second, x[t.first]));
      }
      x[e.pivcol[kk]] = f.neg(f.div(s, pv));
    }
    return true;
  }

  // ---- coefficient list (:20-35, :256-268) ---------------------------------------------
  void augment(std::vector<Elt>& v, const Elt& r) const {
    for (const Elt& x : v) if (f.same_rep(x, r)) return;
    v.push_back(r);
    v.push_back(f.raw_neg(r));
    const Elt t = f.inv(r);
    v.push_back(t);
    v.push_back(f.neg(t));
  }
  std::vector<Elt> coefficients(const Mat& TM, size_t maxnumcoeff) const {
    std::vector<Elt> C{f.from_int(0), f.from_int(1), f.raw_neg(f.from_int(1))};
    for (size_t i = 0; i < TM.rows; ++i)
      for (size_t j = 0; j < TM.cols; ++j)
        if (!f.is_zero(TM.at(i, j))) augment(C, TM.at(i, j));
    for (size_t i = 2; C.size() < maxnumcoeff; ++i) augment(C, f.from_int((int64_t)i));
    if (C.size() > maxnumcoeff) C.resize(maxnumcoeff);
    return C;
  }

  // ---- integer images handed to the C ABI ----------------------------------------------
  // Q: columns of TM scaled by their LCD, Coeffs by their common LCD, previous rows by their own
  // LCD (zero patterns and linear (in)dependence are invariant under these scalings).
  static int64_t lcm64(int64_t a, int64_t b) {
    const wide l = (wide)a / wgcd(a, b) * b;
    if (l > (wide)INT64_MAX) throw RangeError("LCD exceeds 64 bits");
    return (int64_t)l;
  }
  static int64_t scaled(const Rat& r, int64_t lcd) {
    const wide v = (wide)r.num * (lcd / r.den);
    if (wabs(v) > (wide)INT64_MAX) throw RangeError("scaled entry exceeds 64 bits");
    return (int64_t)v;
  }
  void to_int_columns(const Dense<QField>& TM, std::vector<int64_t>& out) const {
    out.assign(TM.rows * TM.cols, 0);
    for (size_t j = 0; j < TM.cols; ++j) {
      int64_t l = 1;
      for (size_t i = 0; i < TM.rows; ++i) l = lcm64(l, TM.at(i, j).den);
      for (size_t i = 0; i < TM.rows; ++i) out[i * TM.cols + j] = scaled(TM.at(i, j), l);
    }
  }
  void to_int_columns(const Dense<ZpField>& TM, std::vector<int64_t>& out) const {
    out.resize(TM.rows * TM.cols);
    for (size_t e = 0; e < out.size(); ++e) out[e] = f.canon(TM.v[e]);
  }
  void to_int_vector(const QField&, const Rat* v, size_t n, int64_t* out) const {
    int64_t l = 1;
    for (size_t i = 0; i < n; ++i) l = lcm64(l, v[i].den);
    for (size_t i = 0; i < n; ++i) out[i] = scaled(v[i], l);
  }
  void to_int_vector(const ZpField& g, const int64_t* v, size_t n, int64_t* out) const {
    for (size_t i = 0; i < n; ++i) out[i] = g.canon(v[i]);
  }

  // ---- CPU testLinComb, only for the canonical-vector fallback (:317-326) ----------------
  bool testLinComb(std::pair<int, int>& weight, Mat& LCoB, Mat& Cand, size_t num, const std::vector<Elt>& w, const Mat& TM) const {
    for (size_t j = 0; j < Cand.cols; ++j) Cand.at(num, j) = w[j];
    if (rank(Cand) > num) {
      int rl = 0, cl = 0;
      for (size_t j = 0; j < TM.cols; ++j) {
        Elt s = f.zero();
        for (size_t i = 0; i < TM.rows; ++i)
          if (!f.is_zero(w[i]) && !f.is_zero(TM.at(i, j))) s = f.add(s, f.mul(w[i], TM.at(i, j)));
        rl += f.is_zero(s);
      }
      for (size_t i = 0; i < w.size(); ++i) cl += f.is_zero(w[i]);
      if (rl > weight.first || (rl == weight.first && cl > weight.second)) {
        weight = {rl, cl};
        for (size_t j = 0; j < LCoB.cols; ++j) LCoB.at(num, j) = w[j];
        return true;
      }
    }
    return false;
  }

  // ---- FactorDiagonals (:354-375; Q18: first maximum in std::map order) -----------------
  void FactorDiagonals(Mat& TCoB, Mat& TM) const {
    for (size_t i = 0; i < TM.rows; ++i) {
      if (rowSize(TM, i) == 0) continue;
      auto cmp = [this](const Elt& a, const Elt& b) { return f.less(a, b); };
      std::map<Elt, int, decltype(cmp)> count(cmp);
      for (size_t j = 0; j < TM.cols; ++j) if (!f.is_zero(TM.at(i, j))) ++count[TM.at(i, j)];
      auto best = count.begin();
      for (auto it = count.begin(); it != count.end(); ++it) if (best->second < it->second) best = it;
      const Elt r = best->first;
      if (!f.is_one(r)) {
        for (size_t j = 0; j < TM.cols; ++j) if (!f.is_zero(TM.at(i, j))) TM.at(i, j) = f.div(TM.at(i, j), r);
        for (size_t j = 0; j < TCoB.cols; ++j) if (!f.is_zero(TCoB.at(i, j))) TCoB.at(i, j) = f.div(TCoB.at(i, j), r);
      }
    }
  }

  // ---- sparseLU (:523-568) / sparseILU (:576-604) -----------------------------------------
  bool sparseLU(Mat& QL, Mat& A, size_t sparsity) const {
    const Elim e = eliminate(A);
    size_t du = 0;
    for (const SRow& r : e.U) du += r.size();
    if (du >= sparsity) return false;
    Mat C(f, A.rows, A.rows);  // Pr^T . L
    for (size_t k = 0; k < A.rows; ++k) {
      C.at(e.rowperm[k], k) = f.one();
      for (const auto& x : e.L[k]) C.at(e.rowperm[k], x.first) = x.second;
    }
    A = dense_of(e.U, A.rows, A.cols);
    QL = C;
    return true;
  }
  bool sparseILU(Mat& TC, Mat& A, size_t sparsity) const {
    Mat QL = identity(A.rows);
    const bool sparser = sparseLU(QL, A, sparsity);
    if (sparser) TC = mul(inverse(QL), TC);  // applyInverse :417-428
    return sparser;
  }

  // =====================================================================================================================
  // sparseAlternate (:609-661) as a resumable state machine.  One AlternateRun owns the transposed block TM, the accumulated
  // inverse change of basis TICoB and the position inside
  //   SparseFactor(3, +4, COEFFICIENT_SEARCH) ; SparseFactor(maxnumcoeff, +1, maxnumcoeff)      (:640-642, loops :497-510)
  //     localSparsifier (:205-347): prelude -> for every inner block: rows via the GPU (+ canonical fallbacks) -> apply
  // advance() runs host work until a device search is needed (returns true, `pending` filled) or the block is finished.
  // =====================================================================================================================
  struct Pending {  // one inner-block search waiting for the device
    size_t off = 0, nprev = 0, c = 0;
    std::vector<int64_t> prev, seed;
    bool has_seed = false;
    int init_rl = -1, init_cl = -1;
  };
  struct AlternateRun {
    Mat TM, TICoB, LCoB;
    std::ostringstream text;  // progress lines of this block
    bool logging = false;
    int phase = -1;           // -1 not started, 0/1 the two SparseFactor calls, 2 finished
    size_t numcoeffs = 0, increment = 0, threshold = 0, ss = 0, s2 = 0, maxnumcoeff = 0;
    // localSparsifier in flight
    std::vector<Elt> Coeffs;
    std::vector<int64_t> tm_int, cf_int;
    int rnHw = -1, cnHw = -1;
    size_t inner = 0, numblocks = 0, rows_done = 0;  // rows_done: rows of the current inner block already decided
    Pending pending;
  };

  std::ostream* out_of(AlternateRun& R) { return R.logging ? &R.text : nullptr; }

  void begin_alternate(AlternateRun& R, const Mat& M, size_t maxnumcoeff) {
    Tick tick(stats.t_begin_alt);
    R.logging = log != nullptr;
    R.maxnumcoeff = maxnumcoeff;
    R.TM = transpose(M);
    R.TICoB = identity(M.cols);
    FactorDiagonals(R.TICoB, R.TM);
    const bool reduced = sparseILU(R.TICoB, R.TM, density(R.TM));
    if (reduced && R.logging) {
      size_t sl, su;
      densityProfile(R.text << "# [sALT] GaussLo profile: ", sl, R.TICoB) << std::endl;
      densityProfile(R.text << "# [sALT] GaussUp profile: ", su, R.TM) << std::endl;
    }
    begin_factor(R, 0);
  }
  void begin_factor(AlternateRun& R, int phase) {  // SparseFactor entry (:473-496)
    R.phase = phase;
    if (phase == 0) { R.numcoeffs = 3u; R.increment = 4u; R.threshold = COEFFICIENT_SEARCH; }
    else { R.numcoeffs = R.maxnumcoeff; R.increment = 1u; R.threshold = R.maxnumcoeff; }
    if (R.logging) densityProfile(R.text << "# [SpFc] Columns profile: ", R.s2, R.TM) << std::endl; else R.s2 = density(R.TM);
    begin_local(R);
  }
  // localSparsifier up to the first search: nullspace prelude (:227-252), Coeffs (:256-270), integer images
  struct Tick {
    double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Tick(double& a) : acc(a) {}
    ~Tick() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  };
  void begin_local(AlternateRun& R) {
    Tick tick(stats.t_begin_local);
    R.ss = R.s2;  // do { ss = s2; localSparsifier(...)
    const Mat& TM = R.TM;
    const size_t n = TM.rows;
    R.LCoB = Mat(f, n, n);
    R.cnHw = -1; R.rnHw = -1;
    if (TM.rows > 1) {
      // rows of TM^T by decreasing size -- the same std::sort call on (size, row) objects as the reference's on its rows (Q7);
      // then the longest prefix of rank n-1: the prefix is extended row by row and an echelon basis tells when rank n is reached
      struct Line { size_t size, row; };
      std::vector<Line> order(TM.cols);
      for (size_t j = 0; j < TM.cols; ++j) { size_t s = 0; for (size_t i = 0; i < n; ++i) s += !f.is_zero(TM.at(i, j)); order[j] = Line{s, j}; }
      std::sort(order.begin(), order.end(), [](const Line& a, const Line& b) { return a.size > b.size; });
      Mat basis(f, n, n);
      std::vector<size_t> lead;  // pivot column of each basis row
      size_t keep = order.size();
      for (size_t t = 0; t < order.size(); ++t) {
        std::vector<Elt> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = TM.at(i, order[t].row);
        for (size_t b = 0; b < lead.size(); ++b)
          if (!f.is_zero(v[lead[b]])) {
            const Elt mlt = v[lead[b]];
            for (size_t i = 0; i < n; ++i) v[i] = f.sub(v[i], f.mul(mlt, basis.at(b, i)));
          }
        size_t pc = n;
        for (size_t i = 0; i < n; ++i) if (!f.is_zero(v[i])) { pc = i; break; }
        if (pc == n) continue;
        if (lead.size() + 1 == n) { keep = t; break; }  // this row would complete rank n: the prefix stops before it
        const Elt ip = f.inv(v[pc]);
        for (size_t i = 0; i < n; ++i) basis.at(lead.size(), i) = f.mul(v[i], ip);
        lead.push_back(pc);
      }
      if (keep > 0) {
        Mat N(f, keep, n);
        for (size_t t = 0; t < keep; ++t) for (size_t i = 0; i < n; ++i) N.at(t, i) = TM.at(i, order[t].row);
        std::vector<Elt> x;
        nullspaceVector(N, x);
        for (size_t i = 0; i < n; ++i) if (!f.is_zero(x[i])) R.LCoB.at(0, i) = x[i];
        R.cnHw = (int)rowSize(R.LCoB, 0);  // number of NON-zeroes (sic, :242)
        R.rnHw = 0;
        for (size_t j = 0; j < TM.cols; ++j) {
          Elt s = f.zero();
          for (size_t i = 0; i < n; ++i)
            if (!f.is_zero(R.LCoB.at(0, i)) && !f.is_zero(TM.at(i, j))) s = f.add(s, f.mul(R.LCoB.at(0, i), TM.at(i, j)));
          R.rnHw += f.is_zero(s);
        }
      }
    }
    R.Coeffs = coefficients(TM, R.numcoeffs);
    if (R.logging) {
      R.text << "# [SPRF] linear combination coefficients: [";
      for (size_t i = 0; i < R.Coeffs.size(); ++i) { if (i) R.text << ' '; print(R.text, R.Coeffs[i]); }
      R.text << ']' << std::endl;
    }
    R.cf_int.resize(R.Coeffs.size());
    to_int_columns(TM, R.tm_int);
    to_int_vector(f, R.Coeffs.data(), R.Coeffs.size(), R.cf_int.data());
    R.numblocks = (TM.rows + 3) >> 2;
    R.inner = 0;
    R.rows_done = 0;
    post_search(R);
  }
  // the search for the undecided rows of the current inner block
  void post_search(AlternateRun& R) {
    const size_t n = R.TM.rows, off = R.inner << 2;
    Pending& P = R.pending;
    P = Pending();
    P.off = off;
    P.nprev = off + R.rows_done;
    P.c = R.Coeffs.size();
    P.prev.assign(P.nprev * n, 0);
    for (size_t i = 0; i < P.nprev; ++i) to_int_vector(f, &R.LCoB.at(i, 0), n, &P.prev[i * n]);
    if (R.inner == 0 && R.rows_done == 0) {  // found starts true with the nullspace weight (:290-295)
      P.init_rl = R.rnHw; P.init_cl = R.cnHw;
      P.has_seed = true;
      P.seed.resize(n);
      to_int_vector(f, &R.LCoB.at(0, 0), n, P.seed.data());
    }
  }
  void set_row(AlternateRun& R, size_t row, uint64_t idx) {
    const size_t n = R.TM.rows, off = R.inner << 2, c = R.Coeffs.size();
    const size_t ids[4] = {(size_t)(idx / (c * c * c)), (size_t)((idx / (c * c)) % c), (size_t)((idx / c) % c), (size_t)(idx % c)};
    for (size_t j = 0; j < n; ++j) R.LCoB.at(row, j) = f.zero();
    for (size_t t = 0; t < 4; ++t) if (off + t < n) R.LCoB.at(row, off + t) = R.Coeffs[ids[t]];
  }
  void canonical_fallback(AlternateRun& R, size_t row) {  // :317-326
    Mat A = R.LCoB;
    std::pair<int, int> weight;
    bool found = false;
    for (size_t q = 0; !found; ++q) {
      if (q >= R.TM.rows) throw RangeError("localSparsifier: no independent canonical vector");
      weight = {-1, -1};
      std::vector<Elt> w(R.TM.rows, f.zero());
      w[q] = f.one();
      found = testLinComb(weight, R.LCoB, A, row, w, R.TM);
    }
    ++stats.fallbacks;
  }
  // Results of the pending search: `nrows` decided rows (index PLO_NO_INDEX: the seed vector keeps row 0) and the quad status.
  // Returns true when another search is pending, false when the block has finished.
  bool consume(AlternateRun& R, int nrows, int status, const uint64_t* index) {
    const size_t n = R.TM.rows, off = R.inner << 2;
    const size_t nact = std::min<size_t>(4u, n - off);
    const unsigned long long c4 = (unsigned long long)R.Coeffs.size() * R.Coeffs.size() * R.Coeffs.size() * R.Coeffs.size();
    for (int t = 0; t < nrows; ++t) {
      if (index[t] != PLO_NO_INDEX) set_row(R, off + R.rows_done, index[t]);
      ++R.rows_done;
      stats.candidates += c4;
    }
    if (status == PLO_QUAD_MISS) {
      stats.candidates += c4;
      // (block 0, num 0) starts with found == true even when the prelude produced no vector (weight (-1,-1), quirk Q5): the row
      // then stays as the prelude left it; every other row without an admissible candidate takes a canonical vector
      if (!(R.pending.has_seed && R.rows_done == 0)) canonical_fallback(R, off + R.rows_done);
      ++R.rows_done;
    } else if (status == PLO_QUAD_RANGE) {  // the device-side filter could not stay exact: this row through the one-row entry point
      single_row(R);
    }
    if (R.rows_done < nact) { post_search(R); return true; }
    R.rows_done = 0;
    if (++R.inner < R.numblocks) { post_search(R); return true; }
    return finish_local(R);
  }
  // one row through plo_lincomb_search (host-computed filter; wide outputs and the RANGE fallback)
  void single_row(AlternateRun& R) {
    const size_t n = R.TM.rows, off = R.inner << 2, nprev = off + R.rows_done;
    post_search(R);
    const Pending& P = R.pending;
    int brl = -1, bcl = -1;
    uint64_t bidx = PLO_NO_INDEX;
    const int rc = plo_lincomb_search((uint32_t)f.characteristic(), (int)n, (int)R.TM.cols, R.tm_int.data(), (int)off, (int)P.c, R.cf_int.data(), (int)nprev,
                                      nprev ? P.prev.data() : nullptr, P.init_rl, P.init_cl, &brl, &bcl, &bidx);
    if (rc != PLO_OK) throw EngineError(rc, plo_last_error());
    ++stats.searches;
    stats.candidates += (unsigned long long)P.c * P.c * P.c * P.c;
    if (bidx != PLO_NO_INDEX) set_row(R, nprev, bidx);
    else if (!P.has_seed) canonical_fallback(R, nprev);
    ++R.rows_done;
  }
  // end of localSparsifier (:335-346) and the do-while of SparseFactor (:497-510)
  bool finish_local(AlternateRun& R) {
    Tick tick(stats.t_finish_local);
    R.TM = mul(R.LCoB, R.TM);
    R.TICoB = mul(R.LCoB, R.TICoB);
    FactorDiagonals(R.TICoB, R.TM);
    if (R.logging) densityProfile(R.text << "# [SpFc] Density profile: ", R.s2, R.TM) << std::endl; else R.s2 = density(R.TM);
    if (R.numcoeffs < R.threshold) R.numcoeffs += R.increment;
    if (R.s2 < R.ss) { begin_local(R); return true; }
    if (R.phase == 0) { begin_factor(R, 1); return true; }
    R.phase = 2;
    return false;
  }
  void finish_alternate(AlternateRun& R, Mat& CoB, Mat& Res) {
    CoB = inverseTranspose(R.TICoB);
    if (R.logging) { size_t sc; densityProfile(R.text << "# [sALT] CoBasis profile: ", sc, CoB) << std::endl; }
    Res = transpose(R.TM);
    if (log) *log << R.text.str() << std::flush;
  }

  // ---- lock-step driver over any number of independent blocks ------------------------------------------------------------
  void run_blocks(std::vector<std::unique_ptr<AlternateRun>>& runs) {
    const uint32_t p = (uint32_t)f.characteristic();
    std::vector<AlternateRun*> active;
    for (auto& r : runs) if (r->phase != 2) active.push_back(r.get());
    while (!active.empty()) {
      const int m = (int)active[0]->TM.cols;
      std::vector<AlternateRun*> next;
      if (m > 64) {  // wide outputs: tiled kernels behind the one-row entry point
        for (AlternateRun* R : active) {
          single_row(*R);
          const uint64_t none = PLO_NO_INDEX;
          if (consume(*R, 0, PLO_QUAD_DONE, &none)) next.push_back(R);
        }
        active.swap(next);
        continue;
      }
      std::vector<plo_quad_problem> probs(active.size());
      for (size_t b = 0; b < active.size(); ++b) {
        AlternateRun& R = *active[b];
        const Pending& P = R.pending;
        plo_quad_problem& q = probs[b];
        q.n = (int)R.TM.rows; q.off = (int)P.off; q.c = (int)P.c; q.nprev = (int)P.nprev;
        q.TM = R.tm_int.data(); q.coeffs = R.cf_int.data();
        q.prev_rows = P.nprev ? P.prev.data() : nullptr;
        q.seed_vec = P.has_seed ? P.seed.data() : nullptr;
        q.init_rl = P.init_rl; q.init_cl = P.init_cl;
      }
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = plo_lincomb_quad(p, m, (int)probs.size(), probs.data());
      stats.device_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rc != PLO_OK) throw EngineError(rc, plo_last_error());
      ++stats.searches;
      for (size_t b = 0; b < active.size(); ++b)
        if (consume(*active[b], probs[b].nrows, probs[b].status, probs[b].index)) next.push_back(active[b]);
      active.swap(next);
    }
  }

  // ---- sparseAlternate (:609-661) -------------------------------------------------------------
  void sparseAlternate(Mat& CoB, Mat& Res, const Mat& M, size_t maxnumcoeff) {
    std::vector<std::unique_ptr<AlternateRun>> runs;
    runs.emplace_back(new AlternateRun());
    begin_alternate(*runs[0], M, maxnumcoeff);
    run_blocks(runs);
    finish_alternate(*runs[0], CoB, Res);
  }

  // ---- blockSparsifier (:666-748) --------------------------------------------------------------
  int blockSparsifier(Mat& CoB, Mat& Res, const Mat& M, size_t blocksize, size_t maxnumcoeff, bool initialElimination) {
    if (blocksize <= 1) { sparseAlternate(CoB, Res, M, maxnumcoeff); return 0; }
    const size_t m = M.rows, n = M.cols;
    Mat U(f, 0, 0), L = identity(n);
    bool reduced = initialElimination;
    if (initialElimination) {
      U = transpose(M);
      reduced = sparseLU(L, U, density(U));
      if (log) {
        size_t sl, su;
        densityProfile(*log << "# [bSpr] IGaussL profile: ", sl, L) << std::endl;
        densityProfile(*log << "# [bSpr] IGaussU profile: ", su, U) << std::endl;
      }
    }
    const Mat A = reduced ? transpose(U) : M;
    // separateColumnBlocks :88-114 ; every block is an independent sparseAlternate (:710-723): they advance in lock step
    std::vector<std::unique_ptr<AlternateRun>> runs;
    std::vector<size_t> widths;
    for (size_t c0 = 0; c0 < n; c0 += blocksize) {
      const size_t w = std::min(blocksize, n - c0);
      Mat B(f, m, w);
      for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < w; ++j) B.at(i, j) = A.at(i, c0 + j);
      runs.emplace_back(new AlternateRun());
      begin_alternate(*runs.back(), B, maxnumcoeff);
      widths.push_back(w);
    }
    run_blocks(runs);
    Res = Mat(f, m, n);
    CoB = Mat(f, n, n);
    Mat TCoB(f, n, n);
    size_t c0 = 0;
    for (size_t b = 0; b < runs.size(); ++b) {
      const size_t w = widths[b];
      Mat C(f, w, w), R(f, m, w);
      finish_alternate(*runs[b], C, R);
      for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < w; ++j) Res.at(i, c0 + j) = R.at(i, j);  // augmentedMatrix :726
      if (reduced) {  // :728-741
        Mat Lb(f, n, w);
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < w; ++j) Lb.at(i, j) = L.at(i, c0 + j);
        const Mat B = mul(Lb, transpose(C));
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < w; ++j) TCoB.at(i, c0 + j) = B.at(i, j);
      } else {  // diagonalMatrix :743
        for (size_t i = 0; i < w; ++i) for (size_t j = 0; j < w; ++j) CoB.at(c0 + i, c0 + j) = C.at(i, j);
      }
      c0 += w;
    }
    if (reduced) CoB = transpose(TCoB);
    return 0;
  }

  // ---- consistency (:871-907): M == R.C ? --------------------------------------------------------
  bool consistency(const Mat& M, const Mat& R, const Mat& C) const {
    const Mat A = mul(R, C);
    for (size_t e = 0; e < M.v.size(); ++e) if (!f.is_zero(f.sub(A.v[e], M.v[e]))) return false;
    return true;
  }

  static void print(std::ostream& o, const Rat& r) { o << r.num; if (r.den != 1) o << '/' << r.den; }
  static void print(std::ostream& o, const int64_t& r) { o << r; }
};

}  // namespace host
}  // namespace plo
