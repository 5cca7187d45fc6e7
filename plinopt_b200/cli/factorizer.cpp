// bin/factorizer -- drop-in for the reference driver src/factorizer.cpp:139-206 (same flags and
// streams: CoB on stdout, Alt and the consistency line on stderr, profileConsistency(...,Alt,-1,CoB,1)
// src/factorizer.cpp:95), with the random restarts of Factorizer running on the GPU.
// -V 1 (initial sparsification) chains plo_sparsifier first, as TFactorizer does (:50-86).
#include <cstdlib>
#include <unistd.h>

#include "cli_common.hpp"

int main(int argc, char** argv) {
  int matformat = plo::host::FF_PRETTY;
  std::string filename;
  size_t innerdim = 0, randomloops = 30, maxnumcoeff = 11, blocksize = 4;  // DEFAULT_RANDOM_LOOPS, COEFFICIENT_SEARCH
  bool initialSparsification = false, initialElimination = true;
  unsigned long long q = 0, seed = 0x504C494E4F505431ull;
  for (int i = 1; i < argc; ++i) {
    const std::string args(argv[i]);
    if (args == "-h") {
      std::clog << "Usage: " << argv[0] << " [-h|-M|-P|-S|-L|[-k|-O|-c|-b|-U|-V #]] [stdin|matrixfile.sms]\n"
                << "  -k #: inner dimension (default is column dimension)\n"
                << "  -M/-P/-S/-L: selects the ouput format\n"
                << "  -V [1|0]: initial sparsification or not (default 0)\n"
                << "  -b #: states the blocking dimension (default " << blocksize << ")\n"
                << "  -c #: max number of coefficients per iteration (default " << maxnumcoeff << ")\n"
                << "  -U [1|0]: initial LU factorization or not (default 1) \n"
                << "  -q #: search modulo (default is Rationals)\n"
                << "  -O #: search for reduced randomized sparsity (default " << randomloops << " loops)\n"
                << "  -s #: seed of the counter-based row orders (B200 engine; the reference is time-seeded)\n";
      exit(-1);
    } else if (args == "-M") matformat = plo::host::FF_MAPLE;
    else if (args == "-S") matformat = plo::host::FF_SMS;
    else if (args == "-P") matformat = plo::host::FF_PRETTY;
    else if (args == "-L") matformat = plo::host::FF_LINALG;
    else if (args == "-k" && i + 1 < argc) innerdim = (size_t)atoi(argv[++i]);
    else if (args == "-q" && i + 1 < argc) q = strtoull(argv[++i], nullptr, 10);
    else if (args == "-V" && i + 1 < argc) initialSparsification = atoi(argv[++i]) != 0;
    else if (args == "-b" && i + 1 < argc) blocksize = (size_t)atoi(argv[++i]);
    else if (args == "-c" && i + 1 < argc) maxnumcoeff = (size_t)atoi(argv[++i]);
    else if (args == "-U" && i + 1 < argc) initialElimination = atoi(argv[++i]) != 0;
    else if (args == "-O" && i + 1 < argc) randomloops = (size_t)strtoull(argv[++i], nullptr, 10);
    else if (args == "-s" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 0);
    else filename = args;
  }
  plo::host::Dense<plo::host::QField> A;
  plo::host::QField Q;
  try {
    if (filename.empty()) { if (!plo::host::read_sms(std::cin, A)) { std::cerr << "# ERROR, malformed SMS on stdin" << std::endl; return -1; } }
    else if (!cli::read_file(filename, A)) return -1;
  } catch (const std::exception& e) { std::cerr << "# ERROR, " << e.what() << std::endl; return -1; }

  std::clog << "# [FCTZ] Initial profile: ";
  const size_t sc = cli::profile(std::clog, A);
  std::clog << std::endl;
  cli::Timer timer;
  cli::NumDen in = cli::flatten(A);
  const int rows = in.rows, cols = in.cols;
  std::vector<int64_t> csn, csd;  // change of basis of the optional sparsification: A = M.Cs
  if (initialSparsification) {
    csn.assign((size_t)cols * cols, 0); csd.assign(csn.size(), 1);
    std::vector<int64_t> mn((size_t)rows * cols), md(mn.size(), 1);
    int consistent = 0;
    const int rc = plo_sparsifier(q, rows, cols, in.num.data(), in.den.data(), (int)blocksize, (int)maxnumcoeff, initialElimination ? 1 : 0,
                                  csn.data(), csd.data(), mn.data(), md.data(), &consistent, nullptr, STDERR_FILENO);
    if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
    std::clog << "# [SPRB] sparsified to " << [&] { size_t s = 0; for (auto v : mn) s += v != 0; return s; }() << " non-zeroes, "
              << (consistent ? "consistent" : "INCONSISTENT") << std::endl << std::string(30, '#') << std::endl;
    in.num = mn; in.den = md;
  }
  const size_t k = innerdim == 0 ? (size_t)cols : innerdim;
  std::vector<int64_t> an((size_t)rows * (k ? k : 1)), ad(an.size(), 1), cn((size_t)(k ? k : 1) * cols), cd(cn.size(), 1);
  uint64_t report[8] = {0};
  int rc = plo_factorizer(q, rows, cols, in.num.data(), in.den.data(), (int)innerdim, randomloops, seed, an.data(), ad.data(), cn.data(), cd.data(), report);
  if (rc == -1) { std::cerr << "# \033[1;36m" << plo_last_error() << "\033[0m\n"; return -1; }  // plinopt_sparsify.inl:937-942
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  if (rows == cols) std::clog << std::string(30, '#') << std::endl << "# \033[1;36mWARNING: identity factorization\033[0m\n";
  auto Alt = cli::unflatten(rows, (int)k, an, ad);
  auto CoB = cli::unflatten((int)k, cols, cn, cd);
  if (initialSparsification) {  // CoB = Ca.Cs  (src/factorizer.cpp:75-84); rationals only on this path
    const auto Cs = cli::unflatten(cols, cols, csn, csd);
    plo::host::Dense<plo::host::QField> Do(Q, k, (size_t)cols);
    for (size_t i = 0; i < k; ++i)
      for (size_t t = 0; t < (size_t)cols; ++t) {
        if (CoB.at(i, t).num == 0) continue;
        for (size_t j = 0; j < (size_t)cols; ++j) Do.at(i, j) = Q.add(Do.at(i, j), Q.mul(CoB.at(i, t), Cs.at(t, j)));
      }
    CoB = Do;
  }
  const double elapsed = timer.seconds();
  // profileConsistency(matformat, elapsed, A, sc, "[FCTZ]", Alt, -1, CoB, 1)
  size_t sa, sb;
  std::clog << "# [FCTZ] chgobase profile: \033[1;36m"; sb = cli::profile(std::clog, CoB); std::clog << "\033[0m" << std::endl;
  plo::host::write_matrix(std::cout, Q, CoB, matformat) << std::endl;
  std::clog << "# [FCTZ] residuum profile: \033[1;36m"; sa = cli::profile(std::clog, Alt); std::clog << "\033[0m" << std::endl;
  plo::host::write_matrix(std::clog, Q, Alt, matformat) << std::endl;
  bool consistent = report[7] != 0;
  if (q == 0) {  // re-check A == Alt.CoB on the final (possibly recombined) matrices
    consistent = true;
    for (size_t i = 0; i < (size_t)rows && consistent; ++i)
      for (size_t j = 0; j < (size_t)cols; ++j) {
        plo::host::Rat s;
        for (size_t t = 0; t < k; ++t) if (Alt.at(i, t).num != 0 && CoB.at(t, j).num != 0) s = Q.add(s, Q.mul(Alt.at(i, t), CoB.at(t, j)));
        if (s != A.at(i, j)) { consistent = false; break; }
      }
  }
  if (consistent) std::clog << "# \033[1;32mSUCCESS: consistent factorization!\033[0m";
  else std::cerr << "# \033[1;31m****** ERROR inconsistency ******\033[0m" << std::endl;
  std::clog << " \033[1;36m" << Alt.rows << 'x' << Alt.cols << " by " << CoB.rows << 'x' << CoB.cols << " with " << sa << " non-zeroes (" << sb
            << " alt.) instead of " << sc << "\033[0m:" << ' ' << elapsed << "s" << std::endl;
  std::clog << "# [B200] " << randomloops << " row orders scored; R/CB profile (" << report[3] << ',' << report[4] << ',' << report[5] << ")";
  if (report[6] != PLO_NO_INDEX) std::clog << " found at candidate " << report[6];
  std::clog << std::endl;
  return 0;
}
