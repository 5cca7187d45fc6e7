import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from plinopt_b200 import capi, hm
capi.set_device(0)
L, R, P = hm.load_fixture("2x2x2_7_Winograd")
mkn = hm.LRP2MM(L, R, P)
(Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int64) for M in (L, R, P))
SEED = 0x1234
nnz, nno, g2 = capi.orbit_table(mkn, Li, Ri, Pi, (dl, dr, dp), 1, SEED, 0, 200)
bad = 0
for i in range(200):
    a = capi.orbit_sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_NNZ, 1, SEED, i, i + 1)
    if (a["nnz"], a["nno"]) != (nnz[i], nno[i]): bad += 1
print("nnz sweep mismatches", bad)
os.environ["PLO_ORBIT_NOPACK8"] = "1"
bad = 0
for i in range(200):
    a = capi.orbit_sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_G2, 1, SEED, i, i + 1)
    if a["score"] != g2[i]: bad += 1
print("g2 two-lane sweep mismatches", bad)
del os.environ["PLO_ORBIT_NOPACK8"]
bad = 0
for i in range(200):
    a = capi.orbit_sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_G2, 1, SEED, i, i + 1)
    if a["score"] != g2[i]: bad += 1
print("g2 four-lane sweep mismatches", bad)
