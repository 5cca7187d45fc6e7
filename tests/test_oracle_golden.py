"""CPU tests: the oracle against every known answer the reference's own tests / data hold for
the hot path: G2 values in data/ headers and file names, MMchecker verdicts of
`make mmcheck` (Makefile:60-64), the FDT.sh consistency invariant (bin/FDT.sh:64-66)."""
import numpy as np
import pytest

import oracle_lib as O

G2_KNOWN = {  # value, tolerance (digits published)
    "2x2x2_7_Strassen": (12 + 2 * 2 ** 0.5, 1e-12),
    "2x2x2_7_Winograd": (17.8530, 5e-5),
    "2x2x2_7_DPS-smallrat-12.2034": (12.2034, 5e-5),
    "2x2x2_7_DPS-evenpow-12.2034": (12.2034, 5e-5),
    "2x2x2_7_DPS-integral-12.0662": (12.06616423, 5e-9),      # data/2x2x2_7_DPS-integral-12.0662_L.sms:1
    "2x2x2_7_DPS-intermediate-12.0695": (12.06954148, 5e-9),  # data/2x2x2_7_DPS-intermediate-12.0695_L.sms:1
}


@pytest.mark.parametrize("stem", sorted(G2_KNOWN))
def test_G2_known_answers(stem):
    val, tol = G2_KNOWN[stem]
    L, R, P = O.triple(stem)
    assert abs(O.growth_G2(L, R, P) - val) <= tol


def test_header_comment_values_match():
    c = O.matrices()["2x2x2_7_DPS-integral-12.0662_L"]["comments"]
    assert any("12.06616423" in s for s in c)


TRIPLES = ["2x2x2_7_Strassen", "2x2x2_7_Winograd", "2x2x2_7_DPS-accurate", "2x2x2_7_DPS-smallrat-12.2034",
           "3x3x3_23_58", "3x3x3_23_Grey-221", "3x3x6_40", "3x6x3_40", "6x3x3_40", "4x4x4_48_rational",
           "4x4x4_48_accurate", "4x4x4_49_156", "3x4x7_63_rational"]


@pytest.mark.parametrize("stem", TRIPLES)
def test_mmchecker_verdicts(stem):
    """Every shipped triple is a valid algorithm (SUCCESS), over Q and mod 513083 / 2^31-1; a corrupted one is not."""
    L, R, P = O.triple(stem)
    m, k, n = O.LRP2MM(L, R, P)
    assert (f"{m}x{k}x{n}" == stem.split("_")[0])
    rng = np.random.default_rng(5)
    placeholder = stem == "2x2x2_7_DPS-accurate"  # 1013 stands for sqrt(3): valid only mod (1013^2-3)/2
    if not placeholder:
        assert O.mmcheck_q(L, R, P, rng.integers(-30, 30, m * k), rng.integers(-30, 30, k * n)) == 0
    else:
        assert O.mmcheck_q(L, R, P, rng.integers(-30, 30, m * k), rng.integers(-30, 30, k * n)) == 1
    for p in ((513083,) if placeholder else (513083, 2147483647)):  # Makefile:62-63: -m 513083 and -r 1013 2 3 -> (1013^2-3)/2 = 513083
        ua, ub = rng.integers(0, p, m * k), rng.integers(0, p, k * n)
        assert O.mmcheck_modp(p, L, R, P, ua, ub) == 0
        L2 = [row[:] for row in L]; L2[0][0] += 1
        assert O.mmcheck_modp(p, L2, R, P, ua, ub) == 1
    assert O.mmcheck_modp(7, L, R[:-1] + [R[-1]], [row[:-1] for row in P[:-1]], np.zeros(m * k, np.int64), np.zeros(k * n, np.int64)) in (1, 3)


def test_alt_cob_fixtures_are_exact_factorisations():
    """X == X-ALT . X-CoB exactly over Q, the consistency() definition."""
    for stem in ["4x4x4_48_rational", "4x4x4_48_accurate", "3x4x7_63_rational"]:
        for x in "LRP":
            X = O.dense_fractions(f"{stem}_{x}"); A = O.dense_fractions(f"{stem}-ALT_{x}"); C = O.dense_fractions(f"{stem}-CoB_{x}")
            if x == "P":  # P = CoB . ALT on the output side
                prod = [[sum(C[i][t] * A[t][j] for t in range(len(A))) for j in range(len(A[0]))] for i in range(len(C))]
            else:
                prod = [[sum(A[i][t] * C[t][j] for t in range(len(C))) for j in range(len(C[0]))] for i in range(len(A))]
            assert prod == X


FDT = ["2x2x2_7_DPS-smallrat-12.2034_L", "2x2x2_7_Winograd_R", "3x3x3_23_58_L", "3x3x6_40_R", "4x4x4_48_rational_L",
       "4x4x4_48_rational_R", "4x4x4_48_accurate_L", "3x4x7_63_rational_L", "3x4x7_63_rational_P", "4x4x4_49_156_P", "cyclic"]


@pytest.mark.parametrize("name", FDT)
@pytest.mark.parametrize("p", [0, 7])
def test_sparsifier_consistency_invariant(name, p):
    """bin/FDT.sh:64-66: `sparsifier -c 5` and `sparsifier -q 7 -c 5` must report a consistent factorisation."""
    M = O.dense_fractions(name)
    CoB, Res, ok, _ = O.sparsifier(M, p, 4, 5, True)
    assert ok
    nnz0 = sum(1 for r in M for v in r if v != 0); nnz1 = sum(1 for r in Res for v in r if v != 0)
    assert nnz1 <= nnz0


def test_c1_trace_shape():
    """Appendix A of SURVEY.md: -c 4 on the 7x4 matrix: steps come in groups of 4, c = 3,(7,11..) then 4."""
    M = O.dense_fractions("2x2x2_7_DPS-smallrat-12.2034_L")
    _, Res, ok, tr = O.sparsifier(M, 0, 4, 4, True, trace=True)
    assert ok and len(tr) % 4 == 0 and tr[0]["c"] == 3 and tr[-1]["c"] == 4
    assert all(t["index"] >= 0 or t["fallback"] >= 0 or (t["block"] == 0 and t["num"] == 0) for t in tr)


def test_coeffs_order_and_modp_quirk():
    """plinopt_sparsify.inl:256-268 + Q3: {0,1,-1}, then r,-r,1/r,-1/r per new entry; mod p the raw negations stay un-reduced."""
    TM = [[O.Fraction(2), O.Fraction(0), O.Fraction(1, 3)]]
    num, den = O.coeffs(TM, 0, 11)
    got = [O.Fraction(int(a), int(b)) for a, b in zip(num, den)]
    F = O.Fraction
    assert got == [F(0), F(1), F(-1), F(2), F(-2), F(1, 2), F(-1, 2), F(1, 3), F(-1, 3), F(3), F(-3)]
    num, _ = O.coeffs([[6, 0, 3]], 7, 11)
    assert num.tolist() == [0, 1, -1, 6, -6, 6, 1, 3, -3, 5, 2]


def test_orbit_decode_is_unimodular_and_bijective():
    """zoiRandomMatrix shape (src/orbiter.cpp:125-136): signed-permuted unit upper triangular, det +-1;
    the exhaustive decode enumerates 48 distinct matrices per 2x2 factor."""
    seen = set()
    for idx in range(48):
        U, V, W = O.orbit_decode(2, 1, 1, 0, 0, idx)
        assert abs(round(np.linalg.det(U))) == 1 and set(np.unique(U)) <= {-1, 0, 1}
        seen.add(U.tobytes())
    assert len(seen) == 40  # 32 with a non-zero off-diagonal trit + the 8 signed 2x2 permutation matrices
    for idx in [0, 1, 12345, 2 ** 40 + 3]:
        for mkn in [(3, 4, 7), (4, 4, 4), (6, 3, 3)]:
            for M in O.orbit_decode(*mkn, 1, 42, idx):
                assert abs(round(np.linalg.det(M))) == 1


def test_orbit_sweep_invariants_small():
    """Transformed triples stay valid algorithms (src/orbiter.cpp:355) and identity-like candidates keep nnz."""
    L, R, P = O.triple("2x2x2_7_Winograd")
    rng = np.random.default_rng(0)
    for idx in [0, 5, 1000, 99999]:
        U, V, W = O.orbit_decode(2, 2, 2, 1, 7, idx)
        Lj, Rg, hP = O.orbit_apply(L, R, P, U, V, W)
        assert O.mmcheck_q(Lj, Rg, hP, rng.integers(-9, 9, 4), rng.integers(-9, 9, 4)) == 0
    res = O.orbit_sweep(L, R, P, 0, 1, 7, 0, 2000)
    assert res["best"][1] == res["nnz"].min()
    first = int(np.flatnonzero((res["nnz"] == res["nnz"].min()))[0])
    cand = np.flatnonzero(res["nnz"] == res["nnz"].min())
    best_nno = res["nno"][cand].min()
    assert res["best"][0] == int(cand[np.flatnonzero(res["nno"][cand] == best_nno)[0]]) and first >= 0
