// bin/sparsifier -- drop-in for the reference driver src/sparsifier.cpp:89-135 (same flags, same
// streams: CoB on stdout, progress + residue + the consistency line on stderr), with the
// candidate search of every localSparsifier step running on the GPU.
#include <cstdlib>
#include <unistd.h>

#include "cli_common.hpp"

int main(int argc, char** argv) {
  int matformat = plo::host::FF_PRETTY;
  std::string filename;
  size_t maxnumcoeff = 11;  // COEFFICIENT_SEARCH, include/plinopt_sparsify.h:36-38
  size_t blocksize = 4;
  bool initialElimination = true;
  unsigned long long q = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string args(argv[i]);
    if (args == "-h") {
      std::clog << "Usage: " << argv[0] << " [-h|-q|-M|-P|-S|-L|-c #|-U [1|0]] [stdin|matfile.sms]\n"
                << "  -c #: max number of coefficients per iteration (default " << maxnumcoeff << ")\n"
                << "  -b #: states the blocking dimension (default " << blocksize << ")\n"
                << "  -U [1|0]: initial LU factorization (default) or not\n"
                << "  -M/-P/-S/-L: selects the ouput format\n"
                << "  -q #: computes modulo q (default is over the rationals)\n";
      exit(-1);
    } else if (args == "-q" && i + 1 < argc) q = strtoull(argv[++i], nullptr, 10);
    else if (args == "-M") matformat = plo::host::FF_MAPLE;
    else if (args == "-S") matformat = plo::host::FF_SMS;
    else if (args == "-P") matformat = plo::host::FF_PRETTY;
    else if (args == "-L") matformat = plo::host::FF_LINALG;
    else if (args == "-c" && i + 1 < argc) maxnumcoeff = (size_t)atoi(argv[++i]);
    else if (args == "-b" && i + 1 < argc) blocksize = (size_t)atoi(argv[++i]);
    else if (args == "-U" && i + 1 < argc) initialElimination = atoi(argv[++i]) != 0;
    else filename = args;
  }
  plo::host::Dense<plo::host::QField> M;
  plo::host::QField Q;
  try {
    if (filename.empty()) { if (!plo::host::read_sms(std::cin, M)) { std::cerr << "# ERROR, malformed SMS on stdin" << std::endl; return -1; } }
    else if (!cli::read_file(filename, M)) return -1;
  } catch (const std::exception& e) { std::cerr << "# ERROR, " << e.what() << std::endl; return -1; }

  const cli::NumDen in = cli::flatten(M);
  const size_t sc = [&] { size_t s = 0; for (auto& e : M.v) s += e.num != 0; return s; }();
  std::vector<int64_t> cn((size_t)in.cols * in.cols), cd(cn.size(), 1), rn((size_t)in.rows * in.cols), rd(rn.size(), 1);
  int consistent = 0;
  uint64_t stats[3] = {0, 0, 0};
  cli::Timer timer;
  const int rc = plo_sparsifier(q, in.rows, in.cols, in.num.data(), in.den.data(), (int)blocksize, (int)maxnumcoeff, initialElimination ? 1 : 0,
                                cn.data(), cd.data(), rn.data(), rd.data(), &consistent, stats, STDERR_FILENO);
  const double elapsed = timer.seconds();
  if (rc != PLO_OK) {
    std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl;
    return rc;
  }
  const auto CoB = cli::unflatten(in.cols, in.cols, cn, cd), Res = cli::unflatten(in.rows, in.cols, rn, rd);
  // profileConsistency(matformat, elapsed, M, sc, "[SPRF]", Res, -1, CoB, 1)   plinopt_sparsify.inl:131-156
  std::clog << std::string(30, '#') << std::endl;
  size_t sb, sa;
  std::clog << "# [SPRF] chgobase profile: \033[1;36m"; sb = cli::profile(std::clog, CoB); std::clog << "\033[0m" << std::endl;
  plo::host::write_matrix(std::cout, Q, CoB, matformat) << std::endl;
  std::clog << "# [SPRF] residuum profile: \033[1;36m"; sa = cli::profile(std::clog, Res); std::clog << "\033[0m" << std::endl;
  plo::host::write_matrix(std::clog, Q, Res, matformat) << std::endl;
  if (consistent) std::clog << "# \033[1;32mSUCCESS: consistent factorization!\033[0m";
  else std::cerr << "# \033[1;31m****** ERROR inconsistency ******\033[0m" << std::endl;
  std::clog << " \033[1;36m" << Res.rows << 'x' << Res.cols << " by " << CoB.rows << 'x' << CoB.cols << " with " << sa << " non-zeroes (" << sb
            << " alt.) instead of " << sc << "\033[0m:" << ' ' << elapsed << "s" << std::endl;
  std::clog << "# [B200] " << stats[0] << " candidates scored in " << stats[1] << " GPU searches (" << stats[2] << " canonical fallbacks)" << std::endl;
  return 0;
}
