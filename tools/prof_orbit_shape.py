#!/usr/bin/env python3
"""One sweep of a given shipped triple through the plan API (for ncu captures of the larger orbit shapes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plinopt_b200 import capi, hm  # noqa: E402

stem = sys.argv[1] if len(sys.argv) > 1 else "3x4x7_63_rational"
measure = capi.MEASURE_G2 if (len(sys.argv) > 2 and sys.argv[2] == "G2") else capi.MEASURE_NNZ
lg = int(sys.argv[3]) if len(sys.argv) > 3 else 20
capi.set_device(0)
L, R, P = hm.load_fixture(stem)
mkn = hm.LRP2MM(L, R, P)
(Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
plan = capi.OrbitPlan(mkn, Li, Ri, Pi, (dl, dr, dp), measure, capi.MODE_PHILOX, 0x504C494E4F505431)
for s in range(3):
    plan.run(s << lg, (s + 1) << lg, 0)
    print(plan.result())
plan.close()
