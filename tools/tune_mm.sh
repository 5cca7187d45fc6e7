# A/B of build variants of the SpMM kernel on one box (tuning aid): bash tools/tune_mm.sh "<flags A>" "<flags B>" ...
for f in "$@"; do
  touch plinopt_b200/csrc/mmcheck.cu
  PLO_NVCC_EXTRA="$f" python -c "from plinopt_b200 import build; build.build_library()" 2>&1 | tail -1
  echo "== $f"; PLO_TIMING=1 python tools/prof_mm.py 4096 2>&1 | tail -2; python tools/prof_mm.py 4096 2>&1 | tail -1
done
