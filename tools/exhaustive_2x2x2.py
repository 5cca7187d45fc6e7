#!/usr/bin/env python3
"""Exhaustive sweep (48^3 = 110592 candidates) of the zoi-parameterised orbit of every shipped 2x2x2 algorithm: initial and best
sparsity / growth factor.  One table, for profiles/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plinopt_b200 import capi, hm  # noqa: E402

capi.set_device(0)
stems = ["2x2x2_7_Strassen", "2x2x2_7_Winograd", "2x2x2_7_DPS-smallrat-12.2034", "2x2x2_7_DPS-evenpow-12.2034", "2x2x2_7_DPS-integral-12.0662",
         "2x2x2_7_DPS-intermediate-12.0695"]
print(f"{'algorithm':36s} {'nnz':>5s} {'G2':>10s} | {'best nnz':>8s} {'(nno)':>6s} {'@index':>7s} | {'best G2':>10s} {'@index':>7s}")
for stem in stems:
    L, R, P = hm.load_fixture(stem)
    mkn = hm.LRP2MM(L, R, P)
    (Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int64) for M in (L, R, P))
    nnz0 = sum(1 for M in (L, R, P) for row in M for v in row if v != 0)
    g0 = capi.growth_G2(np.array([[float(v) for v in row] for row in L]), np.array([[float(v) for v in row] for row in R]),
                        np.array([[float(v) for v in row] for row in P]))[0]
    space = capi.orbit_space(*mkn)
    wide = max(dl, dr, dp) >= 2 ** 31 or max(int(np.abs(A).max()) for A in (Li, Ri, Pi)) >= 2 ** 31  # beyond the int32 C ABI
    sweep = (lambda *a: capi.orbit_sweep64(*a)) if wide else (lambda *a: capi.orbit_sweep(*a))
    try:
        a = sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_NNZ, capi.MODE_EXHAUSTIVE, 0, 0, space)
        b = sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_G2, capi.MODE_EXHAUSTIVE, 0, 0, space)
    except capi.PloError as e:  # common denominators of ~10^9: outside the exact int32 path (PLO_E_RANGE), reported, never wrapped
        print(f"{stem:36s} {nnz0:5d} {g0:10.6f} | {str(e)[:70]}")
        continue
    print(f"{stem:36s} {nnz0:5d} {g0:10.6f} | {a['nnz']:8d} {a['nno']:6d} {a['index']:7d} | {b['score']:10.6f} {b['index']:7d}")
