#!/bin/bash
# Exhaustive sweep of the whole zoi-parameterised orbit of a 3x3x3 algorithm (7776^3 = 470184984576 candidates) through bin/orbiter.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/search/data
python - <<'PY'
import sys
sys.path.insert(0, '.')
from plinopt_b200 import hm
for x, M in zip("LRP", hm.load_fixture("3x3x3_23_58")):
    hm.write_sms(M, f"gpurun_out/search/data/3x3x3_23_58_{x}.sms")
PY
for flag in -s -g; do
  echo "== 3x3x3_23_58 $flag --exhaustive -O 470184984576"
  ( time bin/orbiter $flag --exhaustive -O 470184984576 gpurun_out/search/data/3x3x3_23_58_L.sms gpurun_out/search/data/3x3x3_23_58_R.sms gpurun_out/search/data/3x3x3_23_58_P.sms ) 2>&1 | grep -v "^$"
done
