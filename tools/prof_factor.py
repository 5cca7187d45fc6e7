#!/usr/bin/env python3
"""One Factorizer sweep (for ncu captures): python tools/prof_factor.py [STEM_X] [LOG2_CANDIDATES]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plinopt_b200 import capi, hm  # noqa: E402

P31 = 2147483647
name = sys.argv[1] if len(sys.argv) > 1 else "4x4x4_48_rational_L"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 18
stem, x = name.rsplit("_", 1)
M = hm.load_fixture(stem)["LRP".index(x)]
A = np.array([[(v.numerator % P31) * pow(v.denominator % P31, -1, P31) % P31 for v in row] for row in M], dtype=np.uint32)
capi.set_device(0)
plan = capi.FactorPlan(P31, A, A.shape[1], 0x504C494E4F505431)
for _ in range(2):
    plan.run(0, 1 << lg)
print(plan.result())
plan.close()
