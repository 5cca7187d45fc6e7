// bin/rotater -- bin/rotater.sh:14-85 as one executable: builds the rotated matrix-multiplication algorithm
//   left : R ; (P^T)_s ; (L_s)^T      right: (P^T)_s ; L ; (R_s)^T
// into <stem>_left_{L,R,P}.sms / <stem>_right_{L,R,P}.sms in the current directory (same naming as the script:
// the suffix goes before "_L." / "_R." / "_P."), then checks the result with the batched MMchecker.
#include <cstdlib>

#include "cli_common.hpp"

static std::string rotated_name(const std::string& path, const std::string& suffix) {
  std::string base = path.substr(path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
  for (const char* tag : {"_L.", "_R.", "_P."}) {
    const size_t at = base.find(tag);
    if (at != std::string::npos) { base.insert(at, "_" + suffix); break; }
  }
  return base;
}

int main(int argc, char** argv) {
  std::string suffix = "left";
  std::vector<std::string> files;
  for (int i = 1; i < argc; ++i) {
    const std::string a(argv[i]);
    if ((a == "-d" || a == "--direction") && i + 1 < argc) suffix = argv[++i];
    else if (a == "-l" || a == "--left") suffix = "left";
    else if (a == "-r" || a == "--right") suffix = "right";
    else if (a[0] == '-') { std::cout << "Usage: " << argv[0] << " [-d left/right] [-r|-l] L.sms R.sms P.sms" << std::endl; return 1; }
    else files.push_back(a);
  }
  if (files.size() < 3) { std::cout << "Usage: " << argv[0] << " [-d left/right] [-r|-l] L.sms R.sms P.sms" << std::endl; return 1; }
  plo::host::QField Q;
  plo::host::Dense<plo::host::QField> L, R, P;
  if (!cli::read_file(files[0], L) || !cli::read_file(files[1], R) || !cli::read_file(files[2], P)) return -1;
  const cli::NumDen l = cli::flatten(L), r = cli::flatten(R), p = cli::flatten(P);
  int m, k, n;
  plo_LRP2MM(l.cols, r.cols, p.rows, &m, &k, &n);
  const bool right = suffix == "right";
  const int lc = right ? m * n : k * n, rc_ = right ? m * k : m * n, pr = right ? n * k : m * k;
  std::vector<int64_t> oLn((size_t)l.rows * lc), oLd(oLn.size(), 1), oRn((size_t)l.rows * rc_), oRd(oRn.size(), 1), oPn((size_t)pr * l.rows), oPd(oPn.size(), 1);
  const int rc = plo_rotater(right ? 1 : 0, l.rows, l.cols, r.cols, p.rows, l.num.data(), l.den.data(), r.num.data(), r.den.data(), p.num.data(), p.den.data(),
                             oLn.data(), oLd.data(), oRn.data(), oRd.data(), oPn.data(), oPd.data());
  if (rc != PLO_OK) { std::cerr << "# \033[1;31m****** ERROR " << rc << ": " << plo_last_error() << " ******\033[0m" << std::endl; return rc; }
  const std::string nl = rotated_name(files[0], suffix), nr = rotated_name(files[1], suffix), np = rotated_name(files[2], suffix);
  {
    std::ofstream ol(nl), orr(nr), op(np);
    plo::host::write_matrix(ol, Q, cli::unflatten(l.rows, lc, oLn, oLd), plo::host::FF_SMS);
    plo::host::write_matrix(orr, Q, cli::unflatten(l.rows, rc_, oRn, oRd), plo::host::FF_SMS);
    plo::host::write_matrix(op, Q, cli::unflatten(pr, l.rows, oPn, oPd), plo::host::FF_SMS);
  }
  std::clog << "# rotated <" << m << 'x' << k << 'x' << n << "> " << suffix << " --> " << nl << ' ' << nr << ' ' << np << std::endl;
  uint32_t cnt[2];
  const int v = plo_mmchecker(0, 3, 32, l.rows, lc, l.rows, rc_, pr, l.rows, oLn.data(), oLd.data(), oRn.data(), oRd.data(), oPn.data(), oPd.data(), cnt);  // MMchecker -b 3 (:84)
  int a, b, c;
  plo_LRP2MM(lc, rc_, pr, &a, &b, &c);
  if (v == 0) std::clog << "# \033[1;32mSUCCESS: correct " << a << 'x' << b << 'x' << c << " {" << cnt[0] << ',' << cnt[1] << "} Matrix-Multiplication \033[0m" << std::endl;
  else std::cerr << "# \033[1;31m****** ERROR, not a " << a << 'x' << b << 'x' << c << " MM algorithm (" << v << ") ******\033[0m" << std::endl;
  return v;
}
