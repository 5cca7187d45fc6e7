// lincomb_common.cuh -- pieces shared by the sparsifier search kernels (lincomb_search.cu: one (block,num) step per launch;
// lincomb_quad.cu: all rows of an inner block in one launch sequence): the packed (rl, cl, -index) key, the lazy independence
// test against annihilator functionals, and the host code that derives those functionals from the previous rows.
#pragma once
#include <algorithm>
#include <vector>

#include "host/exact.hpp"
#include "plo_device.cuh"

namespace plo {

constexpr int kLcThreads = 128;
constexpr int kIdxBits = 36;
constexpr unsigned long long kIdxMask = (1ull << kIdxBits) - 1ull;

// key = (rl+1) << 48 | (cl+1) << 36 | (2^36 - 2 - index); the weight seed uses low bits 2^36-1
__host__ __device__ __forceinline__ unsigned long long pack_key(int rl, int cl, unsigned long long low) {
  return ((unsigned long long)(rl + 1) << 48) | ((unsigned long long)(cl + 1) << kIdxBits) | low;
}

template <bool MODP>
__device__ __forceinline__ bool independent(const long long* __restrict__ phi, int nphi, const long long* __restrict__ coef,
                                            unsigned int p, int i, int j, int k, int l) {
  const long long w[4] = {coef[i], coef[j], coef[k], coef[l]};
  for (int q = 0; q < nphi; ++q) {
    if (MODP) {
      unsigned long long s = 0;
#pragma unroll
      for (int t = 0; t < 4; ++t) s += ((unsigned long long)phi[q * 4 + t] * (unsigned long long)w[t]) % p;
      if (s % p) return true;
    } else {
      long long s = 0;
#pragma unroll
      for (int t = 0; t < 4; ++t) s += phi[q * 4 + t] * w[t];
      if (s) return true;
    }
  }
  return false;
}


// ---- host side ----------------------------------------------------------------------------------
inline int pad_m(int m) {
  const int opts[] = {8, 16, 32, 48, 64};
  for (int o : opts) if (m <= o) return o;
  return -1;
}

// One reduced functional -> integers: clear denominators over Q; residues are used as they are mod p.
inline void phi_row(const plo::host::QField&, const plo::host::Rat* row, long long* out) {
  // LCD and scaled numerators in 128 bits: a functional that leaves the int64 range throws (PLO_E_RANGE), never wraps
  using plo::host::wide;
  wide lcd = 1;
  for (int t = 0; t < 4; ++t) {
    lcd = lcd / plo::host::wgcd(lcd, row[t].den) * row[t].den;
    if (lcd > (wide)INT64_MAX) throw plo::host::RangeError("annihilator functional: common denominator exceeds 64 bits");
  }
  for (int t = 0; t < 4; ++t) {
    const wide v = (wide)row[t].num * (lcd / row[t].den);
    if (plo::host::wabs(v) > ((wide)1 << 62)) throw plo::host::RangeError("annihilator functional exceeds 62 bits");
    out[t] = (long long)v;
  }
}
inline void phi_row(const plo::host::ZpField&, const int64_t* row, long long* out) {
  for (int t = 0; t < 4; ++t) out[t] = row[t];
}

// Annihilator functionals of span(prev rows) restricted to the live positions off..off+nact-1,
// reduced to an independent set (<= 4 vectors of 4 entries).  Returns false if the previous rows
// are linearly dependent (then rank(Cand) can never exceed num: plinopt_sparsify.inl:173-175).
template <class F>
bool annihilators(const F& f, int n, int nprev, const int64_t* prev, int off, int nact, std::vector<long long>& phi, int& nphi) {
  using namespace plo::host;
  phi.assign(16, 0);
  nphi = 0;
  std::vector<std::vector<typename F::Elt>> basis;
  if (nprev == 0) {
    for (int t = 0; t < nact; ++t) { phi[nphi * 4 + t] = 1; ++nphi; }  // everything non-zero is independent
    return true;
  }
  Dense<F> A(f, (size_t)nprev, (size_t)n);
  for (int i = 0; i < nprev; ++i)
    for (int j = 0; j < n; ++j) A.at(i, j) = f.modular ? f.from_ratio(prev[(size_t)i * n + j], 1) : f.from_int(prev[(size_t)i * n + j]);
  size_t rk = 0;
  basis = nullspace(f, A, &rk);
  if ((int)rk < nprev) return false;
  Dense<F> R(f, basis.size(), 4);
  for (size_t q = 0; q < basis.size(); ++q)
    for (int t = 0; t < nact; ++t) R.at(q, t) = basis[q][off + t];
  const std::vector<size_t> piv = rref(f, R);
  for (size_t q = 0; q < piv.size(); ++q) {
    phi_row(f, &R.at(q, 0), &phi[(size_t)nphi * 4]);
    ++nphi;
  }
  return true;
}


}  // namespace plo
