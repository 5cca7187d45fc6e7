"""GPU parity: Factorizer random restarts (include/plinopt_sparsify.inl:755-867, 924-990)
against the oracle's restatement of backSolver, candidate by candidate (bit-exact integer scores)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
P31 = 2147483647
SEED = 0x504C494E4F505431


def residues(M, p):
    return np.array([[(v.numerator % p) * pow(v.denominator % p, -1, p) % p for v in row] for row in M], dtype=np.uint32)


def matrix(stem, x):
    M = O.dense_fractions(f"{stem}_{x}")
    if len(M) < len(M[0]):  # P is stored (mn x r): the factorizer works on the tall orientation
        M = [list(c) for c in zip(*M)]
    return M


CASES = [("2x2x2_7_Winograd", "L", 0), ("2x2x2_7_Winograd", "P", 2), ("3x3x3_23_58", "R", 0), ("4x4x4_48_rational", "L", 0),
         ("4x4x4_48_rational", "R", 4), ("3x4x7_63_rational", "L", 0), ("3x4x7_63_rational", "R", 0), ("3x4x7_63_rational", "P", 5),
         ("3x3x6_40", "P", 0)]


@pytest.mark.parametrize("stem,x,extra", CASES)
def test_candidate_table_matches_oracle(capi, stem, x, extra):
    M = matrix(stem, x)
    r, n = len(M), len(M[0])
    k = n + extra
    cnt = 300 if r > 40 else 600
    best, tab = capi.factor_sweep(P31, residues(M, P31), k, SEED, 10, 10 + cnt, table=True)
    ref = O.factor_sweep(M, k, SEED, 10, 10 + cnt, p=P31)
    assert np.array_equal(tab, ref["table"])
    assert best == ref["best"]
    # scoring modulo a 31-bit prime sees the same zero / +-1 pattern as exact rational arithmetic
    exact = O.factor_sweep(M, k, SEED, 10, 10 + cnt)
    assert np.array_equal(tab, exact["table"]) and best == exact["best"]


@pytest.mark.parametrize("p", [513083, 101, 7])
def test_small_moduli(capi, p):
    """`factorizer -q p` (src/factorizer.cpp:116-128): the search itself runs in Z/pZ."""
    M = matrix("4x4x4_48_rational", "L")  # denominators 2, 4, 8: invertible mod every odd p
    best, tab = capi.factor_sweep(p, residues(M, p), 16, 7, 0, 400, table=True)
    ref = O.factor_sweep(M, 16, 7, 0, 400, p=p)
    assert np.array_equal(tab, ref["table"]) and best == ref["best"]


def test_row_order_decode_is_the_oracles(capi):
    for r in (2, 7, 23, 48, 63, 200):
        for idx in (0, 1, 12345, 2**40 + 17):
            assert capi.factor_decode(r, SEED, idx).tolist() == O.factor_decode(r, SEED, idx).tolist()
            assert sorted(capi.factor_decode(r, SEED, idx).tolist()) == list(range(r))


def test_edges(capi):
    M = matrix("2x2x2_7_Winograd", "L")
    A = residues(M, P31)
    assert capi.factor_sweep(P31, A, 4, SEED, 5, 5)[3] is None  # empty range
    one = capi.factor_sweep(P31, A, 4, SEED, 5, 6)
    assert one[3] == 5 and one[:3] == tuple(int(v) for v in O.factor_sweep(M, 4, SEED, 5, 6, p=P31)["table"][0])
    # k = r: every row is in CoB, Alt is the identity
    assert capi.factor_sweep(P31, A, 7, SEED, 0, 50)[:3] == (7, 0, 14)
    # rank-deficient input (column 3 = column 0): no candidate reaches rank n
    B = A.copy(); B[:, 3] = B[:, 0]
    assert capi.factor_sweep(P31, B, 4, SEED, 0, 64)[3] is None
    with pytest.raises(capi.PloError):
        capi.factor_sweep(P31 + 1, A, 4, SEED, 0, 8)  # even modulus
    with pytest.raises(capi.PloError):
        capi.factor_sweep(P31, A, 3, SEED, 0, 8)  # inner dimension below the column dimension (:937-942)


def test_large_sweep_winner_is_reproducible_and_exact(capi):
    """Size-independent properties on a sweep far beyond what the oracle can enumerate: any split of the index range
    gives the same winner; the winner's score is what the exact (rational) oracle computes for that one candidate;
    it is no worse than the best of the oracle's first 2000 candidates."""
    M = matrix("4x4x4_48_rational", "L")
    A = residues(M, P31)
    N = 1 << 20
    whole = capi.factor_sweep(P31, A, 16, SEED, 0, N)
    halves = [capi.factor_sweep(P31, A, 16, SEED, 0, N // 3), capi.factor_sweep(P31, A, 16, SEED, N // 3, N)]
    assert whole == min(halves, key=lambda b: (b[0], b[1], b[2], b[3]))
    exact = O.factor_sweep(M, 16, SEED, whole[3], whole[3] + 1, matrices=True)
    assert tuple(int(v) for v in exact["table"][0]) == whole[:3]
    k = 16
    prod = [[sum(exact["alt"][i][t] * exact["cob"][t][j] for t in range(k)) for j in range(len(M[0]))] for i in range(len(M))]
    assert prod == M  # consistency(), plinopt_sparsify.inl:871-907
    first = O.factor_sweep(M, 16, SEED, 0, 2000)["best"]
    assert whole[:3] <= first[:3]
