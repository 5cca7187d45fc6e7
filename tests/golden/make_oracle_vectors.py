#!/usr/bin/env python3
"""Golden vectors of the hot path, generated ONCE from the CPU oracle (oracle/plo_oracle.cpp) and committed
(tests/golden/oracle_vectors.json): SURVEY.md section 8c asks for per-step (rl, cl, index) traces of the sparsifier for
C1 and C3 and the exhaustive 48^3 orbit table of C2.  tests/test_oracle_vectors.py (CPU) re-derives them from the
oracle on every run, so the oracle cannot drift silently; the GPU tests compare the engine with the oracle itself.
Large tables are stored as histograms + a SHA-256 of the raw little-endian arrays."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

SEED = 0x504C494E4F505431
OUT = os.path.join(HERE, "oracle_vectors.json")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def frac_rows(M):
    return [[str(v) for v in row] for row in M]


def build():
    out = {}
    # C1: sparsifier -c 4 on 2x2x2_7_DPS-smallrat-12.2034_L (BASELINE config 1), and FDT's -c 5 / -q 7 -c 5
    M = O.dense_fractions("2x2x2_7_DPS-smallrat-12.2034_L")
    for key, (q, c) in {"C1_c4": (0, 4), "C1_c5": (0, 5), "C1_q7_c5": (7, 5)}.items():
        CoB, Res, ok, tr = O.sparsifier(M, q, 4, c, True, trace=True)
        out[key] = dict(consistent=ok, trace=tr, CoB=frac_rows(CoB), Res=frac_rows(Res))
    # C3: sparsifier on 4x4x4_48_rational_L mod 2^31-1, -c 11 (trace only; matrices by digest)
    M = O.dense_fractions("4x4x4_48_rational_L")
    CoB, Res, ok, tr = O.sparsifier(M, 2147483647, 4, 11, True, trace=True)
    out["C3_L_p31_c11"] = dict(consistent=ok, trace=tr, digest=digest(np.array(CoB, dtype=np.int64), np.array(Res, dtype=np.int64)),
                               nnz_res=int(sum(1 for row in Res for v in row if v)))
    # C2: exhaustive 48^3 orbit table of 2x2x2_7_Winograd
    L, R, P = O.triple("2x2x2_7_Winograd")
    t = O.orbit_sweep(L, R, P, 3, 0, 0, 0, 110592)
    hist = {str(int(k)): int(v) for k, v in zip(*np.unique(t["nnz"], return_counts=True))}
    out["C2_exhaustive"] = dict(count=110592, nnz_hist=hist, g2_min=float(t["g2"].min()), g2_min_index=int(t["g2"].argmin()),
                                nnz_min=int(t["nnz"].min()), best_g2=list(O.orbit_sweep(L, R, P, 3, 0, 0, 0, 110592, table=False)["best"]),
                                best_nnz=list(O.orbit_sweep(L, R, P, 0, 0, 0, 0, 110592, table=False)["best"]),
                                digest=digest(t["nnz"], t["nno"], t["g2"]))
    # Philox decode of a few candidates (index -> U, V, W)
    dec = {}
    for (m, k, n), idx in (((2, 2, 2), 12345), ((3, 4, 7), 2 ** 40 + 17), ((4, 4, 4), 99)):
        U, V, W = O.orbit_decode(m, k, n, 1, SEED, idx)
        dec[f"{m}x{k}x{n}:{idx}"] = [np.asarray(U).reshape(-1).tolist(), np.asarray(V).reshape(-1).tolist(), np.asarray(W).reshape(-1).tolist()]
    out["orbit_decode_philox"] = dec
    # Factorizer: first 300 candidates of 4x4x4_48_rational_L, k = 16
    f = O.factor_sweep(M, 16, SEED, 0, 300)
    out["factor_4x4x4_L_k16"] = dict(best=list(f["best"]), digest=digest(f["table"]), first=f["table"][:5].tolist(),
                                     order_17=O.factor_decode(48, SEED, 17).tolist())
    # dependency: level 3, 5 coefficients on 3x3x3_23_58_L
    d = O.depender(O.dense_fractions("3x3x3_23_58_L"), 3, 5)
    out["dependency_3x3x3_L_l3_c5"] = dict(nhits=d["nhits"], ncand=d["ncand"], coeffs=[str(c) for c in d["coeffs"]],
                                           hits=[[h[0], h[1], list(h[2]), list(h[3])] for h in d["hits"][:20]])
    return out


if __name__ == "__main__":
    with open(OUT, "w") as fh:
        json.dump(build(), fh, indent=0, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
