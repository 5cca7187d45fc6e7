// sparsify_host.hpp -- host orchestration of the alternate-basis sparsifier around the GPU
// candidate search.  Mirrors the reference's entry points (names, argument meaning, log
// prefixes) of include/plinopt_sparsify.inl so that a reference maintainer can map one onto
// the other:
//   augment :20-35 · rank :38-45 · localSparsifier :205-347 · FactorDiagonals :354-375 ·
//   inverse/inverseTranspose :380-465 · SparseFactor :473-513 · sparseLU :523-568 ·
//   sparseILU :576-604 · sparseAlternate :609-661 · blockSparsifier :666-748 · consistency :871-907
// The quad loop + testLinComb (:299-314, :166-197) is NOT here: it runs on the GPU through
// plo_lincomb_search (lincomb_search.cu).  Everything in this file is O(n^3) glue on
// matrices of a few dozen entries.
//
// LinBox's GaussDomain::QLUPin / nullspacebasisin are not in the reference tree; the pivot
// rule used instead is documented in DESIGN.md ("pivot rule") -- parity unpinned upstream.
#pragma once
#include <algorithm>
#include <map>
#include <numeric>
#include <ostream>
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

// statistics of the GPU part of a run (candidates scored, kernel calls)
struct SearchStats {
  unsigned long long candidates = 0, searches = 0, fallbacks = 0;
};

template <class F>
struct Sparsifier {
  typedef typename F::Elt Elt;
  typedef Dense<F> Mat;
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
    for (size_t i = 0; i < M.rows; ++i) s += rowSize(M, i);
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
    for (size_t i = 0; i < A.rows; ++i)
      for (size_t t = 0; t < A.cols; ++t) {
        if (f.is_zero(A.at(i, t))) continue;
        for (size_t j = 0; j < B.cols; ++j)
          if (!f.is_zero(B.at(t, j))) C.at(i, j) = f.add(C.at(i, j), f.mul(A.at(i, t), B.at(t, j)));
      }
    return C;
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

  // ---- elimination with the documented pivot rule (stand-in for QLUPin) ----------------
  struct Elim {
    size_t rank;
    std::vector<size_t> rowperm, pivcol;
    Mat U, L;
  };
  Elim eliminate(const Mat& A) const {
    const size_t m = A.rows, n = A.cols;
    Elim e;
    e.U = A; e.L = identity(m); e.rank = 0;
    e.rowperm.resize(m);
    std::iota(e.rowperm.begin(), e.rowperm.end(), 0);
    for (size_t k = 0; k < m; ++k) {
      size_t best = m, bestsz = n + 1;  // sparsest non-empty remaining row, first among ties
      for (size_t i = k; i < m; ++i) { const size_t s = rowSize(e.U, i); if (s > 0 && s < bestsz) { bestsz = s; best = i; } }
      if (best == m) break;
      if (best != k) {
        for (size_t j = 0; j < n; ++j) std::swap(e.U.at(k, j), e.U.at(best, j));
        for (size_t j = 0; j < k; ++j) std::swap(e.L.at(k, j), e.L.at(best, j));
        std::swap(e.rowperm[k], e.rowperm[best]);
      }
      size_t pc = n, pcsz = m + 1;  // pivot: entry of that row whose column is sparsest below, first among ties
      for (size_t j = 0; j < n; ++j) {
        if (f.is_zero(e.U.at(k, j))) continue;
        size_t s = 0;
        for (size_t i = k; i < m; ++i) s += !f.is_zero(e.U.at(i, j));
        if (s < pcsz) { pcsz = s; pc = j; }
      }
      e.pivcol.push_back(pc);
      const Elt ip = f.inv(e.U.at(k, pc));
      for (size_t i = k + 1; i < m; ++i) {
        if (f.is_zero(e.U.at(i, pc))) continue;
        const Elt mult = f.mul(e.U.at(i, pc), ip);
        e.L.at(i, k) = mult;
        for (size_t j = 0; j < n; ++j)
          if (!f.is_zero(e.U.at(k, j))) e.U.at(i, j) = f.sub(e.U.at(i, j), f.mul(mult, e.U.at(k, j)));
      }
      ++e.rank;
    }
    return e;
  }
  // first nullspace vector (stand-in for nullspacebasisin column 0, :235-239)
  bool nullspaceVector(const Mat& N, std::vector<Elt>& x) const {
    const size_t n = N.cols;
    const Elim e = eliminate(N);
    std::vector<char> isp(n, 0);
    for (size_t k = 0; k < e.rank; ++k) isp[e.pivcol[k]] = 1;
    size_t fc = n;
    for (size_t j = 0; j < n; ++j) if (!isp[j]) { fc = j; break; }
    x.assign(n, f.zero());
    if (fc == n) return false;
    x[fc] = f.one();
    for (size_t kk = e.rank; kk-- > 0;) {
      Elt s = f.zero();
      for (size_t j = 0; j < n; ++j)
        if (j != e.pivcol[kk] && !f.is_zero(e.U.at(kk, j)) && !f.is_zero(x[j])) s = f.add(s, f.mul(e.U.at(kk, j), x[j]));
      x[e.pivcol[kk]] = f.neg(f.div(s, e.U.at(kk, e.pivcol[kk])));
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

  // ---- localSparsifier (:205-347): GPU quad loop ------------------------------------------
  void localSparsifier(Mat& TCoB, Mat& TM, size_t maxnumcoeff) {
    const size_t n = TCoB.rows;
    Mat LCoB(f, n, n);
    int cnHw = -1, rnHw = -1;
    if (TM.rows > 1) {  // nullspace prelude :227-252
      Mat N = transpose(TM);
      {  // std::sort(N.rowBegin(), N.rowEnd(), sizeSup) :229
        std::vector<std::vector<std::pair<size_t, Elt>>> rows(N.rows);
        for (size_t i = 0; i < N.rows; ++i)
          for (size_t j = 0; j < N.cols; ++j)
            if (!f.is_zero(N.at(i, j))) rows[i].emplace_back(j, N.at(i, j));
        std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) { return a.size() > b.size(); });
        Mat S(f, N.rows, N.cols);
        for (size_t i = 0; i < N.rows; ++i) for (const auto& e : rows[i]) S.at(i, e.first) = e.second;
        N = S;
      }
      while (N.rows > 0 && rank(N) == N.cols) { N.v.resize((N.rows - 1) * N.cols); N.rows -= 1; }
      if (N.rows > 0) {
        std::vector<Elt> x;
        nullspaceVector(N, x);
        for (size_t i = 0; i < n; ++i) if (!f.is_zero(x[i])) LCoB.at(0, i) = x[i];
        cnHw = (int)rowSize(LCoB, 0);  // number of NON-zeroes (sic, :242)
        rnHw = 0;
        for (size_t j = 0; j < TM.cols; ++j) {
          Elt s = f.zero();
          for (size_t i = 0; i < TM.rows; ++i)
            if (!f.is_zero(LCoB.at(0, i)) && !f.is_zero(TM.at(i, j))) s = f.add(s, f.mul(LCoB.at(0, i), TM.at(i, j)));
          rnHw += f.is_zero(s);
        }
      }
    }
    const std::vector<Elt> Coeffs = coefficients(TM, maxnumcoeff);
    if (log) {
      *log << "# [SPRF] linear combination coefficients: [";
      for (size_t i = 0; i < Coeffs.size(); ++i) { if (i) *log << ' '; print(*log, Coeffs[i]); }
      *log << ']' << std::endl;
    }
    std::vector<int64_t> tm_int, cf_int(Coeffs.size());
    to_int_columns(TM, tm_int);
    to_int_vector(f, Coeffs.data(), Coeffs.size(), cf_int.data());

    const size_t numblocks = (TM.rows + 3) >> 2;
    const uint32_t p = (uint32_t)f.characteristic();
    std::vector<int64_t> prev;
    for (size_t block = 0; block < numblocks; ++block) {
      const size_t off = block << 2;
      const size_t firstcolumns = std::min<size_t>(4u, LCoB.rows - off);
      for (size_t num = 0; num < firstcolumns; ++num) {
        Mat A = LCoB;  // :289
        std::pair<int, int> weight{-1, -1};
        bool found = (block == 0) && (num == 0);  // :291
        if (found) weight = {rnHw, cnHw};
        const size_t nprev = off + num;
        prev.assign(nprev * n, 0);
        for (size_t i = 0; i < nprev; ++i) to_int_vector(f, &LCoB.at(i, 0), n, &prev[i * n]);
        int brl = -1, bcl = -1;
        uint64_t bidx = PLO_NO_INDEX;
        const int rc = plo_lincomb_search(p, (int)n, (int)TM.cols, tm_int.data(), (int)off, (int)Coeffs.size(), cf_int.data(), (int)nprev,
                                          nprev ? prev.data() : nullptr, weight.first, weight.second, &brl, &bcl, &bidx);
        if (rc != PLO_OK) throw EngineError(rc, plo_last_error());
        ++stats.searches;
        stats.candidates += (unsigned long long)Coeffs.size() * Coeffs.size() * Coeffs.size() * Coeffs.size();
        if (bidx != PLO_NO_INDEX) {
          const size_t c = Coeffs.size();
          const size_t ids[4] = {(size_t)(bidx / (c * c * c)), (size_t)((bidx / (c * c)) % c), (size_t)((bidx / c) % c), (size_t)(bidx % c)};
          for (size_t j = 0; j < n; ++j) LCoB.at(nprev, j) = f.zero();
          for (size_t t = 0; t < 4; ++t) if (off + t < n) LCoB.at(nprev, off + t) = Coeffs[ids[t]];
          weight = {brl, bcl};
          found = true;
        }
        for (size_t q = 0; !found; ++q) {  // canonical fallback :317-326
          if (q >= TM.rows) throw RangeError("localSparsifier: no independent canonical vector");
          weight = {-1, -1};
          std::vector<Elt> w(TM.rows, f.zero());
          w[q] = f.one();
          found |= testLinComb(weight, LCoB, A, nprev, w, TM);
          if (found) ++stats.fallbacks;
        }
      }
    }
    TM = mul(LCoB, TM);      // :339,343
    TCoB = mul(LCoB, TCoB);  // :340,344
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
    const bool sparser = density(e.U) < sparsity;
    if (sparser) {
      Mat C(f, A.rows, A.rows);
      for (size_t k = 0; k < A.rows; ++k) for (size_t j = 0; j < A.rows; ++j) C.at(e.rowperm[k], j) = e.L.at(k, j);
      A = e.U;
      QL = C;
    }
    return sparser;
  }
  bool sparseILU(Mat& TC, Mat& A, size_t sparsity) const {
    Mat QL = identity(A.rows);
    const bool sparser = sparseLU(QL, A, sparsity);
    if (sparser) TC = mul(inverse(QL), TC);  // applyInverse :417-428
    return sparser;
  }

  // ---- SparseFactor (:473-513) ---------------------------------------------------------------
  size_t SparseFactor(Mat& TICoB, Mat& TM, size_t start = 3u, size_t increment = 4u, size_t threshold = COEFFICIENT_SEARCH) {
    size_t s2;
    if (log) densityProfile(*log << "# [SpFc] Columns profile: ", s2, TM) << std::endl; else s2 = density(TM);
    size_t numcoeffs = start, ss;
    do {
      ss = s2;
      localSparsifier(TICoB, TM, numcoeffs);
      FactorDiagonals(TICoB, TM);
      if (log) densityProfile(*log << "# [SpFc] Density profile: ", s2, TM) << std::endl; else s2 = density(TM);
      if (numcoeffs < threshold) numcoeffs += increment;
    } while (s2 < ss);
    return s2;
  }

  // ---- sparseAlternate (:609-661) -------------------------------------------------------------
  void sparseAlternate(Mat& CoB, Mat& Res, const Mat& M, size_t maxnumcoeff) {
    Mat TM = transpose(M);
    Mat TICoB = identity(M.cols);
    FactorDiagonals(TICoB, TM);
    const bool reduced = sparseILU(TICoB, TM, density(TM));
    if (reduced && log) {
      size_t sl, su;
      densityProfile(*log << "# [sALT] GaussLo profile: ", sl, TICoB) << std::endl;
      densityProfile(*log << "# [sALT] GaussUp profile: ", su, TM) << std::endl;
    }
    SparseFactor(TICoB, TM);                                   // :640
    SparseFactor(TICoB, TM, maxnumcoeff, 1u, maxnumcoeff);      // :642
    CoB = inverseTranspose(TICoB);
    if (log) { size_t sc; densityProfile(*log << "# [sALT] CoBasis profile: ", sc, CoB) << std::endl; }
    Res = transpose(TM);
  }

  // ---- blockSparsifier (:666-748) --------------------------------------------------------------
  int blockSparsifier(Mat& CoB, Mat& Res, const Mat& M, size_t blocksize, size_t maxnumcoeff, bool initialElimination) {
    if (blocksize <= 1) { sparseAlternate(CoB, Res, M, maxnumcoeff); return 0; }
    const size_t m = M.rows, n = M.cols;
    Mat U(f, n, m), L = identity(n);
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
    std::vector<Mat> vC, vR;
    for (size_t c0 = 0; c0 < n; c0 += blocksize) {  // separateColumnBlocks :88-114
      const size_t w = std::min(blocksize, n - c0);
      Mat B(f, m, w);
      for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < w; ++j) B.at(i, j) = A.at(i, c0 + j);
      Mat C(f, w, w), R(f, m, w);
      sparseAlternate(C, R, B, maxnumcoeff);
      vC.push_back(C);
      vR.push_back(R);
    }
    Res = Mat(f, m, n);
    CoB = Mat(f, n, n);
    size_t c0 = 0;
    Mat TCoB(f, n, n);
    for (size_t b = 0; b < vC.size(); ++b) {
      const size_t w = vC[b].cols;
      for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < w; ++j) Res.at(i, c0 + j) = vR[b].at(i, j);  // augmentedMatrix :726
      if (reduced) {  // :728-741
        Mat Lb(f, n, w);
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < w; ++j) Lb.at(i, j) = L.at(i, c0 + j);
        const Mat B = mul(Lb, transpose(vC[b]));
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < w; ++j) TCoB.at(i, c0 + j) = B.at(i, j);
      } else {  // diagonalMatrix :743
        for (size_t i = 0; i < w; ++i) for (size_t j = 0; j < w; ++j) CoB.at(c0 + i, c0 + j) = vC[b].at(i, j);
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
