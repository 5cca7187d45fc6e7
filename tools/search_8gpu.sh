#!/bin/bash
# One-process multi-GPU sweeps through the drop-in CLI: bin/orbiter --gpus 8 on 4x4x4_48_rational, 2^39 candidates per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/search/data
python - <<'PY'
import sys
sys.path.insert(0, '.')
from plinopt_b200 import hm
for x, M in zip("LRP", hm.load_fixture("4x4x4_48_rational")):
    hm.write_sms(M, f"gpurun_out/search/data/4x4x4_48_rational_{x}.sms")
PY
G=${1:-8}
for flag in -s -g; do
  echo "== 4x4x4_48_rational $flag --gpus $G -O $((1 << 39))"
  ( time bin/orbiter $flag --gpus $G -O $((1 << 39)) --seed 7 gpurun_out/search/data/4x4x4_48_rational_L.sms gpurun_out/search/data/4x4x4_48_rational_R.sms gpurun_out/search/data/4x4x4_48_rational_P.sms ) 2>&1 | grep -v "^$" | sed 's/\x1b\[[0-9;]*m//g'
done
