#!/bin/bash
# Larger Philox sweeps through bin/orbiter for the two biggest shipped shapes (best sparsity / growth factor found).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/search/data
python - <<'PY'
import sys
sys.path.insert(0, '.')
from plinopt_b200 import hm
for stem in ("4x4x4_48_rational", "3x4x7_63_rational", "4x4x4_49_156", "3x3x6_40"):
    for x, M in zip("LRP", hm.load_fixture(stem)):
        hm.write_sms(M, f"gpurun_out/search/data/{stem}_{x}.sms")
PY
run() { echo "== $1 $2 -O $3"; ( time bin/orbiter $2 -O $3 --seed 2026 gpurun_out/search/data/$1_L.sms gpurun_out/search/data/$1_R.sms gpurun_out/search/data/$1_P.sms ) 2>&1 | grep -v "^$"; }
run 4x4x4_48_rational -s $((1 << 37))
run 4x4x4_48_rational -g $((1 << 37))
run 3x4x7_63_rational -s $((1 << 36))
run 3x4x7_63_rational -g $((1 << 36))
run 4x4x4_49_156 -s $((1 << 35))
run 4x4x4_49_156 -g $((1 << 35))
run 3x3x6_40 -s $((1 << 35))
run 3x3x6_40 -g $((1 << 35))
