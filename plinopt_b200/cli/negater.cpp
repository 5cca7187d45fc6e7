// bin/negater -- drop-in for src/negater.cpp:68-242: L.sms R.sms P.sms -> L.neg.sms R.neg.sms P.neg.sms
// (common divisors pushed to P, pairs of signs flipped); summary lines "# GCDs:" / "# NEGs:" on stderr.
#include <cstdlib>

#include "cli_common.hpp"

int main(int argc, char** argv) {
  if (argc <= 3 || std::string(argv[1]) == "-h") { std::clog << "Usage:" << argv[0] << " L.sms R.sms P.sms\n"; exit(-1); }
  plo::host::QField Q;
  plo::host::Dense<plo::host::QField> L, R, P;
  if (!cli::read_file(argv[1], L) || !cli::read_file(argv[2], R) || !cli::read_file(argv[3], P)) return -1;
  if (L.rows != R.rows || L.rows != P.cols) {  // :89-95
    std::cerr << "# \033[1;31m****** ERROR, inner dimension mismatch: " << L.rows << "(.)" << R.rows << '|' << P.cols << " ******\033[0m" << std::endl;
    return 2;
  }
  const cli::NumDen l = cli::flatten(L), r = cli::flatten(R), p = cli::flatten(P);
  int m, k, n;
  plo_LRP2MM(l.cols, r.cols, p.rows, &m, &k, &n);
  std::clog << "# Optimizing negative values for " << m << 'x' << k << 'x' << n << " Matrix-Multiplication..." << std::endl;
  std::vector<int64_t> oLn(l.num.size()), oLd(l.num.size()), oRn(r.num.size()), oRd(r.num.size()), oPn(p.num.size()), oPd(p.num.size());
  uint64_t st[12];
  const int rc = plo_negater(0, l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(), p.den.data(),
                             oLn.data(), oLd.data(), oRn.data(), oRd.data(), oPn.data(), oPd.data(), st);
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  const auto Ln = cli::unflatten(l.rows, l.cols, oLn, oLd), Rn = cli::unflatten(r.rows, r.cols, oRn, oRd), Pn = cli::unflatten(p.rows, p.cols, oPn, oPd);
  std::ofstream ol(cli::replace_extension(argv[1], ".neg.sms")), orr(cli::replace_extension(argv[2], ".neg.sms")), op(cli::replace_extension(argv[3], ".neg.sms"));
  plo::host::write_matrix(ol, Q, Ln, plo::host::FF_SMS);
  plo::host::write_matrix(orr, Q, Rn, plo::host::FF_SMS);
  plo::host::write_matrix(op, Q, Pn, plo::host::FF_SMS);
  const uint64_t Gn = st[3] + st[4] + st[5], Nn = st[6] + st[7] + st[8], Sn = st[9] + st[10] + st[11];
  std::clog << "# GCDs: " << (st[1] < st[0] ? "\033[1;32m" : "\033[1;36m") << st[1] << " common divisors instead of " << st[0] << "\033[0m" << std::endl;
  std::clog << "# NEGs: " << (Nn < Gn ? "\033[1;32m" : "\033[1;36m") << Nn << '/' << Sn << ':' << '(' << st[6] << '+' << st[7] << '+' << st[8] << ')' << '/'
            << '(' << st[9] << '+' << st[10] << '+' << st[11] << ')' << " negative instead of " << Gn << '(' << st[3] << '+' << st[4] << '+' << st[5] << ')'
            << '/' << Sn << "\033[0m" << std::endl;
  return 0;
}
