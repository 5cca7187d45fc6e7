// bin/growthfactor -- drop-in for src/growthfactor.cpp:148-231: the growth / error factors of an <m x k x n> algorithm,
// printed in the reference's format ("## G2:\t\t<gamma>\t<log_k gamma>").
#include <cmath>
#include <cstdlib>
#include <iomanip>

#include "cli_common.hpp"

int main(int argc, char** argv) {
  if (argc <= 3 || std::string(argv[1]) == "-h") { std::clog << "Usage:" << argv[0] << " L.sms R.sms P.sms\n"; exit(-1); }
  plo::host::Dense<plo::host::QField> L, R, P;
  if (!cli::read_file(argv[1], L) || !cli::read_file(argv[2], R) || !cli::read_file(argv[3], P)) return -1;
  if (L.rows != R.rows || L.rows != P.cols) {  // :163-169
    std::cerr << "# \033[1;31m****** ERROR, inner dimension mismatch: " << L.rows << "(.)" << R.rows << '|' << P.cols << " ******\033[0m" << std::endl;
    return 2;
  }
  const cli::NumDen l = cli::flatten(L), r = cli::flatten(R), p = cli::flatten(P);
  int m, k, n;
  plo_LRP2MM(l.cols, r.cols, p.rows, &m, &k, &n);
  double g[11];
  const int rc = plo_growth_factors(l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(), p.den.data(), g);
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  std::clog << "# Norms of " << m << 'x' << k << 'x' << n << " Matrix-Multiplication:" << std::endl;
  const char* names[11] = {"## Ginfinf:\t", "## Ginf2:\t", "## G2inf:\t", "## G22:\t\t", "## G2:\t\t", "## Q0:\t\t", "## Qkinfinf:\t", "## Q1inf2:\t",
                           "## Qk12inf:\t", "## Qk2inf:\t", "## Q122:\t"};
  std::clog << std::fixed << std::setw(8) << "#  \t\tGamma \t\tlog_" << k << std::endl;
  for (int t = 0; t < 11; ++t) std::clog << names[t] << g[t] << '\t' << std::log(g[t]) / std::log((double)k) << std::endl;
  return 0;
}
