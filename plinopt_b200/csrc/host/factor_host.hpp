// factor_host.hpp -- host side of the Factorizer:
// PLinOpt::Factorizer  include/plinopt_sparsify.inl:924-990  around the GPU sweep
// (plo_factor_sweep, factor_sweep.cu), and an exact backSolver (:755-867) that rebuilds
// the winning candidate over the true field (Q or Z/qZ) from its row order.
#pragma once
#include <array>

#include "sparsify_host.hpp"

namespace plo {
namespace host {

typedef std::array<size_t, 3> Tricounter;  // (nnz(Alt), non-+-1 of Alt, nnz(CoB)), plinopt_sparsify.inl:914-921
inline bool tricOpCount(const Tricounter& a, const Tricounter& b) { return a < b; }

template <class F>
struct FactorHost {
  typedef typename F::Elt Elt;
  typedef Dense<F> Mat;
  const F& f;
  explicit FactorHost(const F& field) : f(field) {}

  std::pair<size_t, size_t> nonzeroes(const Mat& M) const {  // plinopt_library.inl:258-269
    size_t nnz = 0, nno = 0;
    for (const Elt& e : M.v)
      if (!f.is_zero(e)) { ++nnz; if (!(f.is_one(e) || f.is_mone(e))) ++nno; }
    return {nnz, nno};
  }

  // backSolver for the row order `order` (order[t] = original row at position t).  Incremental
  // reduced echelon basis instead of one rank() per trial (same accepted rows, same swaps).
  // Extra rows (positions n..k-1) get coordinate zero in the solved rows.
  bool backSolver(Mat& CoB, Mat& Alt, const Mat& M, size_t k, std::vector<int> order, Tricounter& ops) const {
    const size_t r = M.rows, n = M.cols;
    Mat E(f, n, 2 * n);  // [reduced basis | transformation], row i <-> i-th kept row
    std::vector<size_t> pc;
    size_t nb = 0;
    for (size_t t = 0; t < r && nb < n; ++t) {
      const size_t row = (size_t)order[t];
      std::vector<Elt> v(2 * n, f.zero());
      for (size_t j = 0; j < n; ++j) v[j] = M.at(row, j);
      v[n + nb] = f.one();
      for (size_t i = 0; i < nb; ++i) {
        const Elt c = v[pc[i]];
        if (f.is_zero(c)) continue;
        for (size_t j = 0; j < 2 * n; ++j) v[j] = f.sub(v[j], f.mul(c, E.at(i, j)));
      }
      size_t p = n;
      for (size_t j = 0; j < n; ++j) if (!f.is_zero(v[j])) { p = j; break; }
      if (p == n) continue;
      const Elt ip = f.inv(v[p]);
      for (size_t j = 0; j < 2 * n; ++j) v[j] = f.mul(v[j], ip);
      for (size_t i = 0; i < nb; ++i) {
        const Elt c = E.at(i, p);
        if (f.is_zero(c)) continue;
        for (size_t j = 0; j < 2 * n; ++j) E.at(i, j) = f.sub(E.at(i, j), f.mul(c, v[j]));
      }
      for (size_t j = 0; j < 2 * n; ++j) E.at(nb, j) = v[j];
      pc.push_back(p);
      std::swap(order[nb], order[t]);  // :793-796
      ++nb;
    }
    if (nb < n) return false;
    CoB = Mat(f, k, n);
    Alt = Mat(f, r, k);
    for (size_t t = 0; t < r; ++t) {
      const size_t row = (size_t)order[t];
      if (t < k) {
        for (size_t j = 0; j < n; ++j) CoB.at(t, j) = M.at(row, j);
        Alt.at(row, t) = f.one();
        continue;
      }
      for (size_t j = 0; j < n; ++j) {  // x_j = sum_i row[pc_i] T[i][j]
        Elt x = f.zero();
        for (size_t i = 0; i < n; ++i) {
          const Elt& c = M.at(row, pc[i]);
          if (!f.is_zero(c) && !f.is_zero(E.at(i, n + j))) x = f.add(x, f.mul(c, E.at(i, n + j)));
        }
        Alt.at(row, j) = x;
      }
    }
    const auto nz = nonzeroes(Alt);
    size_t dc = 0;
    for (const Elt& e : CoB.v) dc += !f.is_zero(e);
    ops = Tricounter{nz.first, nz.second, dc};
    return true;
  }
};

}  // namespace host
}  // namespace plo
