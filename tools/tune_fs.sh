# A/B of build variants of the Factorizer kernel on one box (tuning aid): bash tools/tune_fs.sh "<flags A>" "<flags B>" ...
for f in "$@"; do
  touch plinopt_b200/csrc/factor_sweep.cu
  PLO_NVCC_EXTRA="$f" python -c "from plinopt_b200 import build; build.build_library()" 2>&1 | tail -1
  echo "== $f"; grep -E "Used [0-9]+ registers" plinopt_b200/build/factor_sweep.o.log | sort | uniq -c | head -12
  python tools/bench_kernels.py --factor-only 2>&1 | grep candidates_per_s | cut -c1-170
done
