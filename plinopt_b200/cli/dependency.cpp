// bin/dependency -- drop-in for the reference driver src/dependency.cpp:212-251 (same flags; the hits
// on stdout in the reference's depth-first order, "# [DEPND] ..." progress on stderr), with the Explore
// enumeration running on the GPU.
#include <cstdlib>
#include <sstream>

#include "cli_common.hpp"

int main(int argc, char** argv) {
  std::string filename;
  unsigned long long q = 0;
  size_t maxnumcoeff = 11, level = 4;  // COEFFICIENT_SEARCH, src/dependency.cpp:218-219
  std::vector<int64_t> un, ud;
  for (int i = 1; i < argc; ++i) {
    const std::string args(argv[i]);
    if (args == "-h") {
      std::clog << "Usage: " << argv[0] << " [-h|[-c|-l|-q] #] [-v \"# ... #\"] [stdin|matfile.sms]\n"
                << "  -c #: max number of coefficients per iteration\n"
                << "  -v \"# ... #\": string of space separated coefficients\n"
                << "  -l #: maximal number of monomials in the combination\n"
                << "  -q #: modular generation/check (default is Rationals)\n";
      exit(-1);
    } else if (args == "-q" && i + 1 < argc) q = strtoull(argv[++i], nullptr, 10);
    else if (args == "-c" && i + 1 < argc) maxnumcoeff = (size_t)atoi(argv[++i]);
    else if (args == "-l" && i + 1 < argc) level = (size_t)atoi(argv[++i]);
    else if (args == "-v" && i + 1 < argc) {
      std::stringstream sin(argv[++i]);
      std::string tok;
      while (sin >> tok) {
        const size_t slash = tok.find('/');
        un.push_back(strtoll(tok.substr(0, slash).c_str(), nullptr, 10));
        ud.push_back(slash == std::string::npos ? 1 : strtoll(tok.substr(slash + 1).c_str(), nullptr, 10));
      }
    } else filename = args;
  }
  plo::host::Dense<plo::host::QField> M;
  try {
    if (filename.empty()) { if (!plo::host::read_sms(std::cin, M)) { std::cerr << "# ERROR, malformed SMS on stdin" << std::endl; return -1; } }
    else if (!cli::read_file(filename, M)) return -1;
  } catch (const std::exception& e) { std::cerr << "# ERROR, " << e.what() << std::endl; return -1; }
  const cli::NumDen in = cli::flatten(M);
  cli::Timer timer;
  uint64_t max_hits = 1 << 16, nhits = 0, ncand = 0, tlen = 0;
  std::vector<plo_dep_hit> hits;
  std::vector<char> text;
  std::vector<int64_t> cn(maxnumcoeff), cd(maxnumcoeff, 1);
  int ncoef = 0, rc = 0;
  for (int attempt = 0; attempt < 3; ++attempt) {  // further passes only if the hit list or its text was larger than the first guess
    hits.resize(max_hits);
    text.resize(max_hits * 64 + 64);
    rc = plo_depender(q, in.rows, in.cols, in.num.data(), in.den.data(), (int)un.size(), un.data(), ud.data(), (int)maxnumcoeff, (int)level,
                      max_hits, hits.data(), &nhits, &ncand, text.data(), text.size(), &tlen, cn.data(), cd.data(), &ncoef);
    if (rc == PLO_E_RANGE && nhits > max_hits) { max_hits = nhits; continue; }  // more hits than room: come back with room for all of them
    if (rc != PLO_OK || tlen < text.size()) break;
  }
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  std::clog << "# [DEPND] level " << level << ", coefficients: [";
  for (int v = 0; v < ncoef; ++v) { std::clog << (v ? "," : "") << cn[v]; if (cd[v] != 1) std::clog << '/' << cd[v]; }
  std::clog << ']' << std::endl;
  std::cout << text.data();
  std::clog << "# [DEPND]: " << timer.seconds() << "s" << std::endl;
  std::clog << "# [B200] " << ncand << " combinations tested, " << nhits << " reported" << std::endl;
  return 0;
}
