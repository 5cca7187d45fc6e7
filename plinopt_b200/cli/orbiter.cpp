// bin/orbiter -- drop-in for src/orbiter.cpp:365-441: random DeGroote-orbit search writing
// <stem>.nnz.sms next to every input when the winner improves on it (:338-348).
// Flags of the reference are kept; -z/-c (SLP-operation and canonical measures) and -P/-I
// (polynomial quotient) are out of scope and rejected.  Added: -g (minimise the growth factor G2
// instead of the sparsity), --seed N, --exhaustive.
#include <algorithm>
#include <cstdlib>

#include "cli_common.hpp"

static void usage(const char* prg, size_t loops) {
  std::clog << "Usage: " << prg << "  [-h|-O/-b #|-m/-q #|-r # # #|-s|-g] L.sms R.sms P.sms\n"
            << "  [-b b]: random check with values of size 'bitsize' (at most 32 here)\n"
            << "  [-m/-q m]: check is modulo (mod) or (mod/2^k) (default no)\n"
            << "  [-r r e s]: check is modulo (r^e-s) or ((r^e-s)/2^k) (default no)\n"
            << "  [-s|-g]: search sparser|lower growth factor (default is sparser)\n"
            << "  [-O #]: randomized search with that many loops (default " << loops << " loops)\n"
            << "  [--seed #] [--exhaustive]: candidate enumeration (default Philox, seed 0x504C494E4F505431)\n"
            << "  [--gpus #]: shard the sweep over that many GPUs of this box (default 1)\n";
  exit(-1);
}

int main(int argc, char** argv) {
  unsigned long long loops = 100, modulus = 0, seed = 0x504C494E4F505431ull, bitsize = 32;  // DEFAULT_RANDOM_LOOPS, plinopt_library.h:37-39
  int measure = PLO_MEASURE_NNZ, mode = PLO_MODE_PHILOX;
  std::vector<std::string> files;
  for (int i = 1; i < argc; ++i) {
    const std::string a(argv[i]);
    if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 0);
    else if (a == "--gpus" && i + 1 < argc) plo_set_sweep_devices(atoi(argv[++i]));
    else if (a == "--exhaustive") mode = PLO_MODE_EXHAUSTIVE;
    else if (a[0] == '-' && a.size() > 1) {
      if (a[1] == 'h') usage(argv[0], loops);
      else if (a[1] == 'b' && i + 1 < argc) {
        // -b is the bit size of the random integer inputs of the reference's over-Q check (plinopt_library.inl:497-500); it is used
        // the same way here, up to 32 bits: the input triple is checked at 32 random points with b-bit coordinates, exactly over Q
        // (plo_mmchecker_bits) or modulo -m/-q/-r.
        bitsize = strtoull(argv[++i], nullptr, 10);
        std::cerr << "# \033[1;33mNOTE: -b " << bitsize << ": 32 random points with " << (bitsize > 32 ? 32 : bitsize) << "-bit coordinates (at most 32 bits here)\033[0m" << std::endl;
      }
      else if ((a[1] == 'm' || a[1] == 'q') && i + 1 < argc) modulus = strtoull(argv[++i], nullptr, 10);
      else if (a[1] == 'r' && i + 3 < argc) {
        const unsigned long long r = strtoull(argv[++i], nullptr, 10); const int e = atoi(argv[++i]); const unsigned long long s = strtoull(argv[++i], nullptr, 10);
        unsigned long long pw = 1; for (int t = 0; t < e; ++t) pw *= r;
        modulus = pw - s;
      } else if (a[1] == 'O' && i + 1 < argc) loops = strtoull(argv[++i], nullptr, 10);
      else if (a[1] == 's') measure = PLO_MEASURE_NNZ;
      else if (a[1] == 'g') measure = PLO_MEASURE_G2;
      else { std::cerr << "# \033[1;31m****** ERROR, option " << a << " is out of scope of this engine ******\033[0m" << std::endl; return -1; }
    } else files.push_back(a);
  }
  if (files.size() < 3) usage(argv[0], loops);
  plo::host::QField Q;
  plo::host::Dense<plo::host::QField> L, R, P;
  if (!cli::read_file(files[0], L) || !cli::read_file(files[1], R) || !cli::read_file(files[2], P)) return -1;
  if (L.rows != R.rows || L.rows != P.cols) {
    // The reference only warns here (its `return 2` is commented out, src/orbiter.cpp:236-242) and then fails inside LinBox; this
    // engine takes flat r x cols buffers, so a mismatched triple is rejected with fMMchecker's code (src/MMchecker.cpp:65-71).
    std::cerr << "# \033[1;31m****** ERROR, inner dimension mismatch: " << L.rows << "(.)" << R.rows << '|' << P.cols << " ******\033[0m" << std::endl;
    return 2;
  }
  const cli::NumDen l = cli::flatten(L), r = cli::flatten(R), p = cli::flatten(P);
  uint32_t cnt[2];
  const int v0 = plo_mmchecker_bits(modulus, (int)(bitsize < 1 ? 1 : (bitsize > 32 ? 32 : bitsize)), seed, 32, l.rows, l.cols, r.rows, r.cols, p.rows, p.cols, l.num.data(), l.den.data(),
                                    r.num.data(), r.den.data(), p.num.data(), p.den.data(), cnt, nullptr);  // :251 (result ignored by the reference)
  int m, k, n;
  plo_LRP2MM(l.cols, r.cols, p.rows, &m, &k, &n);
  if (v0 == 0) std::clog << "# \033[1;32mSUCCESS: correct " << m << 'x' << k << 'x' << n << " {" << cnt[0] << ',' << cnt[1] << "} Matrix-Multiplication \033[0m" << std::endl;
  else std::cerr << "# \033[1;31m****** ERROR, not a " << m << 'x' << k << 'x' << n << " MM algorithm******\033[0m" << std::endl;
  std::vector<int64_t> oLn(l.num.size()), oLd(l.num.size()), oRn(r.num.size()), oRd(r.num.size()), oPn(p.num.size()), oPd(p.num.size());
  plo_orbiter_report rep;
  cli::Timer timer;
  int rc;
  if (modulus > 0) {  // Orbiter<0>()(QQ, F, ...): the search runs in Z/pZ (src/orbiter.cpp:419-426)
    if (measure != PLO_MEASURE_NNZ) { std::cerr << "# \033[1;31m****** ERROR, -g needs the rationals ******\033[0m" << std::endl; return -1; }
    std::fill(oLd.begin(), oLd.end(), 1); std::fill(oRd.begin(), oRd.end(), 1); std::fill(oPd.begin(), oPd.end(), 1);
    rc = plo_orbiter_modp(modulus, mode, seed, loops, l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(),
                          p.den.data(), oLn.data(), oRn.data(), oPn.data(), &rep);
  } else
    rc = plo_orbiter(measure, mode, seed, loops, l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(),
                     p.den.data(), oLn.data(), oLd.data(), oRn.data(), oRd.data(), oPn.data(), oPd.data(), &rep);
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  std::clog << "# Init. ops: " << rep.init_score << ", {" << rep.init_nnz << ',' << rep.init_nno << '}' << std::endl;
  if (modulus == 0 && rep.improved) {  // '# Found opt:' lines of src/orbiter.cpp:312-315, in index order (deterministic)
    std::vector<plo_orbit_best> recs(256);
    uint64_t nrec = 0;
    if (plo_orbiter_progress(measure, mode, seed, loops, l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(),
                             p.den.data(), recs.size(), recs.data(), &nrec) == PLO_OK) {
      double bestopt = rep.init_score;
      uint32_t bnnz = rep.init_nnz, bnno = rep.init_nno;
      for (uint64_t t = 0; t < nrec; ++t) {
        std::clog << "# Found opt: " << recs[t].score << (recs[t].score < bestopt ? '<' : '=') << bestopt << "\t{" << recs[t].nnz << ',' << recs[t].nno << "}&{" << bnnz << ','
                  << bnno << "}\t[" << recs[t].index << "/gpu]" << std::endl;
        bestopt = recs[t].score; bnnz = recs[t].nnz; bnno = recs[t].nno;
      }
    }
  }
  std::clog << "# Search(" << loops << "): " << timer.seconds() << "s" << std::endl;
  if (rep.improved) {
    std::clog << "# \033[1;36mRdcd. opt: " << rep.best.score << '<' << rep.init_score << "\t{" << rep.best.nnz << ',' << rep.best.nno << "}\033[0m\t[" << rep.best.index << ']' << std::endl;
    const auto Lj = cli::unflatten(l.rows, l.cols, oLn, oLd), Rg = cli::unflatten(r.rows, r.cols, oRn, oRd), hP = cli::unflatten(p.rows, p.cols, oPn, oPd);
    std::ofstream ol(cli::replace_extension(files[0], ".nnz.sms")), orr(cli::replace_extension(files[1], ".nnz.sms")), op(cli::replace_extension(files[2], ".nnz.sms"));
    plo::host::write_matrix(ol, Q, Lj, plo::host::FF_SMS);
    plo::host::write_matrix(orr, Q, Rg, plo::host::FF_SMS);
    plo::host::write_matrix(op, Q, hP, plo::host::FF_SMS);
    if (rep.mm_verdict == 0) std::clog << "# \033[1;32mSUCCESS: correct " << m << 'x' << k << 'x' << n << " {" << rep.best.nnz << ',' << rep.best.nno << "} Matrix-Multiplication \033[0m" << std::endl;
    else std::cerr << "# \033[1;31m****** ERROR, not a " << m << 'x' << k << 'x' << n << " MM algorithm******\033[0m" << std::endl;
  }
  return 0;
}
