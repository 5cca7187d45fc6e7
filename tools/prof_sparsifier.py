"""One plo_sparsifier call per case (for ncu launch lists): python tools/prof_sparsifier.py [q] [c]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import time_sparsifier_c as t  # noqa: E402
from plinopt_b200 import capi, hm  # noqa: E402

q = int(sys.argv[1]) if len(sys.argv) > 1 else 2147483647
c = int(sys.argv[2]) if len(sys.argv) > 2 else 11
capi.set_device(0)
M = hm.load_fixture("4x4x4_48_rational")[0]
print(t.time_case(M, q, c, reps=1))
