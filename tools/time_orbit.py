#!/usr/bin/env python3
"""Times the sweep kernel of one shipped triple, both measures: python tools/time_orbit.py [stem] [log2 candidates]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plinopt_b200 import capi, hm  # noqa: E402

stem = sys.argv[1] if len(sys.argv) > 1 else "3x4x7_63_rational"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 22
capi.set_device(0)
L, R, P = hm.load_fixture(stem)
mkn = hm.LRP2MM(L, R, P)
(Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
st = torch.cuda.current_stream()
B = 1 << lg
for meas, name in ((capi.MEASURE_NNZ, "nnz"), (capi.MEASURE_G2, "G2")):
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, (dl, dr, dp), meas, capi.MODE_PHILOX, 0x504C494E4F505431)
    for _ in range(3):
        plan.run(0, B, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for s in range(5):
        plan.run(s * B, (s + 1) * B, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(stem, name, plan.kernel, "%.3f ms" % ms, "%.4g candidates/s" % (B / ms * 1e3), plan.result())
    plan.close()
