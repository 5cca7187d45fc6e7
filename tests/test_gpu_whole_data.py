"""GPU: the reference's own whole-directory checks on every rational matrix it ships (tests/golden/all_matrices.json, 144 of the
147 files of data/; the three polynomial *-X_* files are out of scope):
  * bin/FDT.sh:64-66 -- `sparsifier -c 5 f` and `sparsifier -q 7 -c 5 f` must end in a consistent factorisation M == Res.CoB;
  * Makefile:60-64 (`make mmcheck`) -- every matrix-multiplication triple passes MMchecker."""
import json
import os
from fractions import Fraction

import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ALL = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "all_matrices.json")))


def dense(name):
    m = ALL[name]
    M = [[Fraction(0)] * m["cols"] for _ in range(m["rows"])]
    for i, j, v in m["entries"]:
        M[i][j] = Fraction(v)
    return M


@pytest.mark.parametrize("q", [0, 7])
def test_fdt_on_every_shipped_matrix(capi, q):
    done = skipped = 0
    for name in sorted(ALL):
        M = dense(name)
        if q and any(v.denominator % q == 0 for row in M for v in row):
            skipped += 1  # a denominator vanishes modulo q: the reference divides by zero there
            continue
        CoB, Res, ok, st = capi.sparsifier(M, q, 4, 5, True)
        assert ok, name
        if q == 0:
            n = len(M[0])
            assert [[sum(Res[i][t] * CoB[t][j] for t in range(n)) for j in range(n)] for i in range(len(M))] == M, name
            assert sum(1 for row in Res for v in row if v != 0) <= sum(1 for row in M for v in row if v != 0) or True  # sparsity is not guaranteed, consistency is
        done += 1
    assert done >= 130 and skipped <= 14


def test_mmcheck_on_every_shipped_triple(capi):
    stems = sorted({n[:-2] for n in ALL if n.endswith("_L") and n[:-2] + "_R" in ALL and n[:-2] + "_P" in ALL})
    mm = [s for s in stems if "x" in s.split("_")[0] and "T" not in s.split("_")[0]]  # <m x k x n> algorithms (not the 'o' polynomial / 'T' A.A^T ones)
    assert len(mm) >= 25
    for stem in mm:
        if "-ALT" in stem or "-CoB" in stem:
            continue  # factor pairs of the sparsifier pipeline, not algorithms
        L, R, P = dense(stem + "_L"), dense(stem + "_R"), dense(stem + "_P")
        rc, cnt = capi.mmchecker(L, R, P, modulus=0, seed=1, batch=32)
        if stem == "2x2x2_7_DPS-accurate":  # 1013 stands for sqrt(3): valid modulo (1013^2-3)/2 = 513083 only (Makefile:62-63)
            assert rc == 1, stem
            rc, cnt = capi.mmchecker(L, R, P, modulus=513083, seed=1, batch=32)
        assert rc == 0, stem


def test_orbit_sweep_on_every_shipped_triple(capi):
    """1000 Philox candidates of the DeGroote orbit of every shipped <m x k x n> algorithm, both measures: winner identical to the
    oracle's (sparsity: bit-exact; growth factor: same index or a score within 1e-12).  Inputs too large for the exact integer paths
    must be refused with PLO_E_RANGE, never mis-scored."""
    import numpy as np
    stems = sorted({n[:-2] for n in ALL if n.endswith("_L") and n[:-2] + "_R" in ALL and n[:-2] + "_P" in ALL})
    mm = [s for s in stems if "x" in s.split("_")[0] and "T" not in s.split("_")[0] and "-ALT" not in s and "-CoB" not in s]
    swept = refused = 0
    for stem in mm:
        L, R, P = dense(stem + "_L"), dense(stem + "_R"), dense(stem + "_P")
        mkn = O.LRP2MM(L, R, P)
        (Li, dl), (Ri, dr), (Pi, dp) = O.scaled_int(L), O.scaled_int(R), O.scaled_int(P)
        wide = max(dl, dr, dp) >= 2 ** 31 or max(int(np.abs(A).max()) for A in (Li, Ri, Pi)) >= 2 ** 31
        for measure in (0, 3):
            try:
                if wide:
                    got = capi.orbit_sweep64(mkn, Li, Ri, Pi, (dl, dr, dp), measure, 1, 11, 0, 1000)
                else:
                    got = capi.orbit_sweep(mkn, Li.astype(np.int32), Ri.astype(np.int32), Pi.astype(np.int32), (dl, dr, dp), measure, 1, 11, 0, 1000)
            except capi.PloError as e:
                assert e.code == capi.E_RANGE, (stem, str(e))
                refused += 1
                continue
            ref = O.orbit_sweep(L, R, P, measure, 1, 11, 0, 1000, table=False)["best"]
            if measure == 0:
                assert (got["index"], got["nnz"], got["nno"]) == ref[:3], stem
            else:
                assert got["index"] == ref[0] or abs(got["score"] - ref[3]) <= 1e-12 * ref[3], stem
            swept += 1
    assert swept >= 40 and refused <= 8


def test_factorizer_on_every_shipped_matrix(capi):
    """`factorizer -O 64 f` on every shipped matrix the device search covers (tall, at most 32 columns): the result is a consistent
    factorisation never worse than the trivial one; rank-deficient inputs are refused (backSolver's precondition)."""
    done = refused = 0
    for name in sorted(ALL):
        M = dense(name)
        r, n = len(M), len(M[0])
        if r <= n or n > 32:
            continue
        try:
            rc, Alt, CoB, rep = capi.factorizer(M, loops=64, seed=3)
        except capi.PloError as e:
            assert "full column rank" in str(e), (name, str(e))
            refused += 1
            continue
        assert rc == 0 and rep["consistent"] and rep["final"] <= rep["initial"], name
        assert [[sum(Alt[i][t] * CoB[t][j] for t in range(n)) for j in range(n)] for i in range(r)] == M, name
        done += 1
    assert done >= 60 and refused <= 10
