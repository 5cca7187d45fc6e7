// adapter_check.cpp -- drives include/plinopt_b200_linbox.hpp through the LinBox mock (tests/linbox_mock.hpp).
// Reads a triple in a trivial text form on stdin ("name rows cols nnz" then "i j num den" lines, three times: L R P), prints one JSON
// object.  Without arguments only the host-side conversions run (no device needed); with "device" the three shims run too:
// quad_rows on TM = first four columns of L transposed, orbit_sweep (sparsity, 4096 candidates), mmcheck mod 2^31-1.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>

#include "linbox_mock.hpp"

typedef mock::SparseMatrix<mock::QField> QMat;
typedef mock::SparseMatrix<mock::ZpField> ZMat;

static QMat read_q(std::istream& in) {
  std::string name; size_t r, c, nnz;
  in >> name >> r >> c >> nnz;
  QMat M(mock::QField(), r, c);
  for (size_t e = 0; e < nnz; ++e) { size_t i, j; mock::Rational v; in >> i >> j >> v.n >> v.d; M.setEntry(i, j, v); }
  return M;
}
static int64_t inv_mod(int64_t a, int64_t p) { int64_t r0 = p, r1 = ((a % p) + p) % p, t0 = 0, t1 = 1; while (r1) { int64_t q = r0 / r1, t = r0 - q * r1; r0 = r1; r1 = t; t = t0 - q * t1; t0 = t1; t1 = t; } return ((t0 % p) + p) % p; }
static ZMat mod_p(const QMat& M, int64_t p) {
  ZMat Z(mock::ZpField{p}, M.rowdim(), M.coldim());
  for (size_t i = 0; i < M.rowdim(); ++i)
    for (const auto& e : M[i]) Z.setEntry(i, e.first, (int64_t)((__int128)(((e.second.n % p) + p) % p) * inv_mod(e.second.d, p) % p));
  return Z;
}
template <class V> static void dump(const char* key, const V& v, bool last = false) {
  printf("\"%s\": [", key);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%lld", i ? "," : "", (long long)v[i]);
  printf("]%s", last ? "" : ", ");
}

int main(int argc, char** argv) {
  const bool device = argc > 1 && !strcmp(argv[1], "device");
  const QMat L = read_q(std::cin), R = read_q(std::cin), P = read_q(std::cin);
  try {
    // TM = transpose of the first four columns of L (a localSparsifier block), Coeffs = {0, 1, -1, 1/2, -1/2, 2, -2}
    QMat TM(mock::QField(), 4, L.rowdim());
    for (size_t i = 0; i < L.rowdim(); ++i) for (const auto& e : L[i]) if (e.first < 4) TM.setEntry(e.first, i, e.second);
    const std::vector<mock::Rational> Coeffs{{0, 1}, {1, 1}, {-1, 1}, {1, 2}, {-1, 2}, {2, 1}, {-2, 1}};
    QMat LCoB(mock::QField(), 4, 4);
    printf("{");
    dump("tm", plo::adapter::scale_columns(TM));
    dump("coeffs", plo::adapter::scale_vector(TM.field(), Coeffs));
    int32_t dL;
    dump("L32", plo::adapter::scale_to_int32(L, dL));
    printf("\"denL\": %d, ", dL);
    const plo::adapter::Csr c = plo::adapter::to_csr(mod_p(P, 2147483647));
    dump("csr_ptr", c.ptr); dump("csr_col", c.col); dump("csr_val", c.val, !device);
    if (device) {
      const plo::adapter::QuadRows q = plo::adapter::quad_rows(TM, Coeffs, LCoB, 0, 0, -1, -1);
      printf("\"quad_status\": %d, \"quad_nrows\": %d, ", q.status, q.nrows);
      dump("quad_rl", std::vector<int>(q.rl, q.rl + 4)); dump("quad_cl", std::vector<int>(q.cl, q.cl + 4));
      dump("quad_index", std::vector<long long>(q.index, q.index + 4));
      std::vector<int32_t> U, V, W;
      const plo_orbit_best b = plo::adapter::orbit_sweep(L, R, P, PLO_MEASURE_NNZ, 0x504C494E4F505431ull, 4096, U, V, W);
      printf("\"orbit\": [%u, %u, %llu], ", b.nnz, b.nno, (unsigned long long)b.index);
      dump("U", U);
      printf("\"mmcheck\": %d", plo::adapter::mmcheck(mod_p(L, 2147483647), mod_p(R, 2147483647), mod_p(P, 2147483647), 1, 64));
    }
    printf("}\n");
  } catch (const plo::adapter::Error& e) {
    fprintf(stderr, "adapter error %d: %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
