#!/usr/bin/env python3
"""One batched MMchecker pass on the regenerated 32x32x32_15096 triple (for ncu captures):
  python tools/prof_mm.py BATCH [REPS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plinopt_b200 import capi, hm  # noqa: E402

P31 = 2147483647
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
capi.set_device(0)
mkn, r, (L, R, P) = hm.load_large_csr(P31)
plan = capi.MMcheckPlan(P31, mkn, r, L, R, P, B)
for i in range(reps):
    plan.run(1, i * B, None)
v, ok = plan.result()
print("verdict", v, "ok", int(ok.sum()), "of", B)
plan.close()
