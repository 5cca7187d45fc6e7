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


// ---- inverse lookup (mod p, many coefficients) ----------------------------------------------------------------------------------
// For a fixed prefix (i,j,k) coordinate e of v vanishes iff  C_l . A3_e = -(s_e),  s_e = C_i A0_e + C_j A1_e + C_k A2_e.  When A3_e is
// invertible that pins the VALUE of C_l: x_e = s_e . (-A3_e^-1); a hash table over the coefficient values turns it into the (usually
// one, possibly zero or several) positions l.  So a prefix costs one multiplication + one probe per coordinate plus a scan of c byte
// counters, instead of c compares per coordinate: ~5x fewer instructions at c = 128 (m = 48).  Exactly the same zero counts.
struct InvTables {           // one problem; lives in global memory, staged into shared memory by each block
  static constexpr unsigned kEmpty = 0xFFFFFFFFu;
  // layout (32-bit words): ninv[mpad] | htab[2 * hsize] (value, first l) | nextdup[c] (next l with the same value, kEmpty: none)
  static __host__ __device__ size_t words(int mpad, int hsize, int c) { return (size_t)mpad + 2 * (size_t)hsize + (size_t)c; }
};
__host__ __device__ __forceinline__ unsigned inv_hash(unsigned x, int hbits) { return (x * 0x9E3779B1u) >> (32 - hbits); }

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


// Builds the inverse-lookup tables of one problem (residues mod an ODD p): a3 = the fourth live row of TM (nullptr / zeros when the block has
// fewer than four live positions), coef = c canonical residues.  ninv[e] = -A3_e^-1 mod p, or kEmpty when A3_e = 0.  The product tables T0, T1,
// T2 of a plan that uses the lookup are PRE-MULTIPLIED by ninv[e] (fold_inv_into_tables on the host, quad_tables_kernel on the device), so
// the value C_l must have at coordinate e is just T0[i][e] + T1[j][e] + T2[e][k] mod p: no multiplication per (prefix, coordinate); the
// kernels read ninv[e] only to tell the coordinates that do not depend on l.  Returns false when some A3_e is not invertible (composite
// modulus).
inline bool build_inv_tables(uint32_t p, int m, int mpad, int c, int hbits, const int64_t* a3, const int64_t* coef, uint32_t* out) {
  const int hsize = 1 << hbits;
  plo::host::ZpField f((int64_t)p);
  uint32_t* ninv = out;
  uint32_t* htab = out + mpad;
  uint32_t* nextdup = htab + 2 * (size_t)hsize;
  for (int e = 0; e < mpad; ++e) {
    const int64_t v = (e < m && a3) ? a3[e] : 0;
    if (v == 0) { ninv[e] = InvTables::kEmpty; continue; }  // the coordinate does not depend on l
    int64_t iv;
    try { iv = f.inv(v); } catch (const plo::host::RangeError&) { return false; }
    ninv[e] = (uint32_t)f.neg(iv);
  }
  for (int h = 0; h < hsize; ++h) { htab[2 * h] = 0; htab[2 * h + 1] = InvTables::kEmpty; }
  for (int l = 0; l < c; ++l) nextdup[l] = InvTables::kEmpty;
  for (int l = c - 1; l >= 0; --l) {  // descending: the head of a chain is its smallest l, chains ascend
    const uint32_t x = (uint32_t)coef[l];
    unsigned h = inv_hash(x, hbits);
    for (;;) {
      if (htab[2 * h + 1] == InvTables::kEmpty) { htab[2 * h] = x; htab[2 * h + 1] = (uint32_t)l; break; }
      if (htab[2 * h] == x) { nextdup[l] = htab[2 * h + 1]; htab[2 * h + 1] = (uint32_t)l; break; }
      h = (h + 1) & (unsigned)(hsize - 1);
    }
  }
  return true;
}
// T0, T1 ([c][mpad]) and T2 ([m][c], transposed) of one problem times ninv[e] where the coordinate depends on l
inline void fold_inv_into_tables(uint32_t p, int m, int mpad, int c, const uint32_t* ninv, uint32_t* t0, uint32_t* t1, uint32_t* t2) {
  for (int e = 0; e < m; ++e) {
    const uint64_t ni = ninv[e];
    if (ni == InvTables::kEmpty) continue;
    for (int l = 0; l < c; ++l) {
      t0[(size_t)l * mpad + e] = (uint32_t)(t0[(size_t)l * mpad + e] * ni % p);
      t1[(size_t)l * mpad + e] = (uint32_t)(t1[(size_t)l * mpad + e] * ni % p);
      t2[(size_t)e * c + l] = (uint32_t)(t2[(size_t)e * c + l] * ni % p);
    }
  }
}
#ifdef __CUDACC__
// The counting loop of the inverse-lookup kernels for ONE prefix (i, j, k): p0 = T0 row i, p1 = T1 row j (m words each), p2 = &T2[0][k]
// (stride c words), all three pre-multiplied by ninv.  Every coordinate costs three loads, two modular additions, one hash probe; a hit increments the byte
// counter of every l with that coefficient value.  Returns the number of coordinates that vanish whatever l is.  The shared-memory
// tables are addressed through 32-bit shared addresses the caller made opaque (under register pressure ptxas re-derives generic
// shared pointers -- S2R + LEA -- inside the loop, ncu profiles/ncu_r02_ad_lincomb_inv_redc.md), the global rows through running
// pointers instead of 64-bit index arithmetic per load.
struct InvShared { uint32_t ninv, htab, nextdup, hist; };  // shared addresses: tables of the problem, byte counters of this thread
__device__ __forceinline__ unsigned inv_count_prefix(const unsigned int* __restrict__ p0, const unsigned int* __restrict__ p1, const unsigned int* __restrict__ p2,
                                                     int c, int m, unsigned int p, int hbits, const InvShared sh) {
  unsigned base = 0;
  const unsigned hmask = (1u << hbits) - 1u;
#pragma unroll 2
  for (int e = 0; e < m; ++e, p2 += c) {
    unsigned int sum = p0[e] + p1[e];  // p <= 2^31: no overflow
    sum -= sum >= p ? p : 0u;
    sum += *p2;
    sum -= sum >= p ? p : 0u;
    unsigned int ni;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ni) : "r"(sh.ninv + 4u * (unsigned)e));
    if (ni == InvTables::kEmpty) {
      base += (sum == 0u);
    } else {
      const unsigned int x = sum;  // the value C_l must have (the tables carry the factor -A3_e^-1)
      unsigned h = inv_hash(x, hbits);
      for (;;) {
        unsigned int ex, ey;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ex), "=r"(ey) : "r"(sh.htab + 8u * h));
        if (ey == InvTables::kEmpty) break;
        if (ex == x) {
          for (unsigned l = ey; l != InvTables::kEmpty;) {
            unsigned int v;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sh.hist + l));
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(sh.hist + l), "r"(v + 1u) : "memory");
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(l) : "r"(sh.nextdup + 4u * l));
          }
          break;
        }
        h = (h + 1u) & hmask;
      }
    }
  }
  return base;
}
__device__ __forceinline__ InvShared inv_shared_addresses(const void* ninv, const void* htab, const void* nextdup, const void* hist) {
  InvShared s;
  asm volatile("mov.u32 %0, %1;" : "=r"(s.ninv) : "r"((uint32_t)__cvta_generic_to_shared(ninv)));
  asm volatile("mov.u32 %0, %1;" : "=r"(s.htab) : "r"((uint32_t)__cvta_generic_to_shared(htab)));
  asm volatile("mov.u32 %0, %1;" : "=r"(s.nextdup) : "r"((uint32_t)__cvta_generic_to_shared(nextdup)));
  asm volatile("mov.u32 %0, %1;" : "=r"(s.hist) : "r"((uint32_t)__cvta_generic_to_shared(hist)));
  return s;
}
#endif

inline int inv_hash_bits(int c) { int b = 3; while ((1 << b) < 4 * c) ++b; return b; }

}  // namespace plo
