// negate_host.hpp -- host-only pre/post passes around a sweep:
//   * negater  (src/negater.cpp:117-209): per product i, move the common divisors of the numerators and of
//     the denominators of row i of L and of R into column i of P, then flip the signs of two of the three
//     (L_i, R_i, P^T_i) when that lowers the number of negative coefficients;
//   * rotater  (bin/rotater.sh:75-83 with src/columns-swap.cpp and matrix-transpose): the cyclic rotations
//     <m,k,n> -> <k,n,m> ("left": R ; (P^T)_s ; (L_s)^T) and -> <n,m,k> ("right": (P^T)_s ; L ; (R_s)^T),
//     where X_s re-vectorises every row of X, seen as a small matrix, into its transpose.
// Exact rational arithmetic; nothing here touches the GPU.
#pragma once
#include "sparsify_host.hpp"

namespace plo {
namespace host {

struct NegaterStats {
  uint64_t gcd_before = 0, gcd_after = 0;          // rows with a non-trivial common numerator / denominator divisor (:124-140)
  uint64_t neg_before[3] = {0, 0, 0}, neg_after[3] = {0, 0, 0};  // negative coefficients in L, R, P
  uint64_t entries[3] = {0, 0, 0};
  uint64_t swaps = 0;                              // rows where a pair of signs was flipped
};

class Negater {
  QField Q;
  typedef Dense<QField> Mat;

  static int64_t iabs(int64_t a) { return a < 0 ? -a : a; }
  static int64_t gcd64(int64_t a, int64_t b) { return (int64_t)wgcd(a, b); }
  // common divisor of the numerators and of the denominators of the non-zero entries of row i (ndGCD :29-53)
  size_t nd_gcd(const Mat& M, size_t i, int64_t& num, int64_t& den) const {
    num = 0; den = 0;
    for (size_t j = 0; j < M.cols; ++j) {
      const Rat& e = M.at(i, j);
      if (e.num == 0) continue;
      num = gcd64(num, iabs(e.num));
      den = gcd64(den, e.den);
    }
    return (size_t)(num > 1) + (size_t)(den > 1);
  }
  void scale_row(Mat& M, size_t i, const Rat& c) const {
    for (size_t j = 0; j < M.cols; ++j) if (M.at(i, j).num != 0) M.at(i, j) = Q.mul(M.at(i, j), c);
  }
  // divide one row by c, multiply the other (swapMultipliers :56-63)
  void swap_multipliers(Mat& divM, Mat& mulM, size_t i, int64_t c) const {
    if (c == 0 || c == 1) return;
    scale_row(divM, i, Rat::make(1, c));
    scale_row(mulM, i, Rat(c));
  }
  static size_t row_size(const Mat& M, size_t i) { size_t s = 0; for (size_t j = 0; j < M.cols; ++j) s += M.at(i, j).num != 0; return s; }
  static size_t row_negs(const Mat& M, size_t i) { size_t s = 0; for (size_t j = 0; j < M.cols; ++j) s += M.at(i, j).num < 0; return s; }
  void negate_row(Mat& M, size_t i) const { for (size_t j = 0; j < M.cols; ++j) M.at(i, j) = Q.neg(M.at(i, j)); }

 public:
  // L (r x mk), R (r x kn), Pt (r x mn = P transposed) are modified in place.
  NegaterStats run(Mat& L, Mat& R, Mat& Pt, bool only_sign = false) const {
    NegaterStats st;
    for (size_t i = 0; i < L.rows; ++i) {
      if (!only_sign) {
        int64_t ln, ld, rn, rd, pn, pd;
        st.gcd_before += nd_gcd(L, i, ln, ld) + nd_gcd(R, i, rn, rd) + nd_gcd(Pt, i, pn, pd);
        swap_multipliers(L, Pt, i, ln);   // :135-138
        swap_multipliers(Pt, L, i, ld);
        swap_multipliers(R, Pt, i, rn);
        swap_multipliers(Pt, R, i, rd);
        st.gcd_after += nd_gcd(L, i, ln, ld) + nd_gcd(R, i, rn, rd) + nd_gcd(Pt, i, pn, pd);
      }
      const size_t sl = row_size(L, i), sr = row_size(R, i), sp = row_size(Pt, i);
      const size_t nl = row_negs(L, i), nr = row_negs(R, i), np = row_negs(Pt, i);
      st.entries[0] += sl; st.entries[1] += sr; st.entries[2] += sp;
      st.neg_before[0] += nl; st.neg_before[1] += nr; st.neg_before[2] += np;
      const size_t none = nl + nr + np;
      const size_t nlr = (sl - nl) + (sr - nr) + np, nlp = (sl - nl) + nr + (sp - np), nrp = nl + (sr - nr) + (sp - np);
      if (nlr < none && nlr <= nlp && nlr <= nrp) {            // :168-176
        negate_row(L, i); negate_row(R, i); ++st.swaps;
        st.neg_after[0] += sl - nl; st.neg_after[1] += sr - nr; st.neg_after[2] += np;
      } else if (nlp < none && nlp < nlr && nlp <= nrp) {      // :177-186
        negate_row(L, i); negate_row(Pt, i); ++st.swaps;
        st.neg_after[0] += sl - nl; st.neg_after[1] += nr; st.neg_after[2] += sp - np;
      } else if (nrp < none && nrp < nlp && nrp < nlr) {       // :187-196
        negate_row(R, i); negate_row(Pt, i); ++st.swaps;
        st.neg_after[0] += nl; st.neg_after[1] += sr - nr; st.neg_after[2] += sp - np;
      } else {
        st.neg_after[0] += nl; st.neg_after[1] += nr; st.neg_after[2] += np;
      }
    }
    return st;
  }
};

// columns-swap (src/columns-swap.cpp:41-52): every row, seen as an (n x m) row-major matrix, is transposed and re-vectorised
inline Dense<QField> columns_swap(const Dense<QField>& A, size_t m) {
  QField Q;
  const size_t n = A.cols / m;
  Dense<QField> S(Q, A.rows, A.cols);
  for (size_t r = 0; r < A.rows; ++r)
    for (size_t col = 0; col < A.cols; ++col) {
      const size_t i = col % m, j = (col - i) / m;
      S.at(r, i * n + j) = A.at(r, col);
    }
  return S;
}
inline Dense<QField> transposed(const Dense<QField>& A) {
  QField Q;
  Dense<QField> T(Q, A.cols, A.rows);
  for (size_t i = 0; i < A.rows; ++i) for (size_t j = 0; j < A.cols; ++j) T.at(j, i) = A.at(i, j);
  return T;
}
// bin/rotater.sh:75-83.  left: <m,k,n> -> <k,n,m>; right: <m,k,n> -> <n,m,k>.
inline void rotate(bool right, size_t k, size_t n, const Dense<QField>& L, const Dense<QField>& R, const Dense<QField>& P,
                   Dense<QField>& Lr, Dense<QField>& Rr, Dense<QField>& Pr) {
  if (right) {
    Lr = columns_swap(transposed(P), n);
    Rr = L;
    Pr = transposed(columns_swap(R, n));
  } else {
    Lr = R;
    Rr = columns_swap(transposed(P), n);
    Pr = transposed(columns_swap(L, k));
  }
}

}  // namespace host
}  // namespace plo
