"""One plo_lincomb_quad call on the four column blocks of 4x4x4_48_rational_L mod 2^31-1: python tools/prof_quad.py [c=128]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from plinopt_b200 import capi  # noqa: E402

c = int(sys.argv[1]) if len(sys.argv) > 1 else 128
capi.set_device(0)
L, tms, cfs = bench.c3_problem(c)
probs = [dict(TM=tm, off=0, coeffs=cf) for tm, cf in zip(tms, cfs)]
capi.lincomb_quad(bench.P31, probs)
t0 = time.perf_counter()
res = capi.lincomb_quad(bench.P31, probs)
dt = time.perf_counter() - t0
print(c, "%.3f ms" % (dt * 1e3), "%.4g covered candidates/s" % (16 * c ** 4 / dt), res[0])
