#!/bin/bash
# Large orbit sweeps through the drop-in CLI (BASELINE metric: "best nnz/growth found"): writes the logs under gpurun_out/search/
# usage: tools/search_results.sh [LOG2_LOOPS]   (default 33: 2^33 candidates per run)
set -u
cd "$(dirname "$0")/.."
LG=${1:-33}
LOOPS=$((1 << LG))
OUT=gpurun_out/search
mkdir -p $OUT/data
python - <<'PY'
import sys, os
sys.path.insert(0, '.')
from plinopt_b200 import hm
for stem in ("2x2x2_7_Winograd", "3x3x3_23_58", "4x4x4_48_rational", "3x4x7_63_rational"):
    for x, M in zip("LRP", hm.load_fixture(stem)):
        hm.write_sms(M, f"gpurun_out/search/data/{stem}_{x}.sms")
PY
for stem in 2x2x2_7_Winograd 3x3x3_23_58 4x4x4_48_rational 3x4x7_63_rational; do
  for flag in -s -g; do
    echo "== $stem $flag -O $LOOPS" | tee -a $OUT/summary.txt
    ( time bin/orbiter $flag -O $LOOPS $OUT/data/${stem}_L.sms $OUT/data/${stem}_R.sms $OUT/data/${stem}_P.sms ) 2>&1 | grep -v "^$" | tee -a $OUT/summary.txt | tail -6
  done
done
