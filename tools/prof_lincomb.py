#!/usr/bin/env python3
"""One sparsifier candidate search (BASELINE config 3 shape: 4x4x4_48_rational_L mod 2^31-1, 4 blocks, c = 64) and one dependency
exploration (48x16, level 4, 11 coefficients), for ncu captures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from plinopt_b200 import capi, hm  # noqa: E402
from bench_kernels import coeff_list  # noqa: E402

P31 = 2147483647
capi.set_device(0)
L, _, _ = hm.load_fixture("4x4x4_48_rational")
tms, cfs = [], []
for blk in range(4):
    TM = [[L[i][4 * blk + t] for i in range(len(L))] for t in range(4)]
    tm = np.array([[(v.numerator % P31) * pow(v.denominator % P31, -1, P31) % P31 for v in row] for row in TM], dtype=np.int64)
    tms.append(tm); cfs.append(coeff_list(tm.tolist(), P31, int(sys.argv[1]) if len(sys.argv) > 1 else 64))
plan = capi.LincombPlan(P31, np.stack(tms), 0, np.stack(cfs))
for _ in range(2):
    plan.run()
print(plan.result())
plan.close()
d = capi.depender(L, 4, 11, q=P31, max_hits=1 << 16)
print(d["ncand"], d["nhits"])
