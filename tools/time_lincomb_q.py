import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from plinopt_b200 import capi
capi.set_device(0)
for name in ("4x4x4_48_rational_L", "3x4x7_63_rational_R", "2x2x2_7_DPS-smallrat-12.2034_L"):
    M = O.dense_fractions(name)
    TM = [[M[i][t] for i in range(len(M))] for t in range(4)]
    tm_num, tm_den = O.numden(TM)
    cn, cd = O.coeffs(TM, 0, 11)
    lc = int(np.lcm.reduce(cd))
    cf = np.array([int(a) * (lc // int(d)) for a, d in zip(cn, cd)], dtype=np.int64)
    cols = []
    for j in range(len(TM[0])):
        l = int(np.lcm.reduce([TM[t][j].denominator for t in range(4)]))
        cols.append([int(TM[t][j] * l) for t in range(4)])
    tm_int = np.array(cols, dtype=np.int64).T.copy()
    print(name, "max|tm|", np.abs(tm_int).max(), "max|cf|", np.abs(cf).max())
    for rep in range(3):
        t0 = time.perf_counter(); plan = capi.LincombPlan(0, tm_int, 0, cf); t1 = time.perf_counter()
        plan.run(); r = plan.result(); t2 = time.perf_counter(); plan.close(); t3 = time.perf_counter()
        print("  create %.2f ms  run+result %.2f ms  close %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), r)
