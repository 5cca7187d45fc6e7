"""CPU: the host-only passes around a sweep -- negater (src/negater.cpp:117-209) and rotater
(bin/rotater.sh:75-83) -- engine vs the oracle's literal restatement, plus the invariants the reference relies on: the
result is still a matrix-multiplication algorithm (Makefile:60-64 style check) and three rotations are the identity."""
import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import capi

TRIPLES = ["2x2x2_7_Strassen", "2x2x2_7_Winograd", "2x2x2_7_DPS-smallrat-12.2034", "3x3x3_23_58", "3x3x6_40", "4x4x4_48_rational",
           "4x4x4_48_accurate", "3x4x7_63_rational"]


def is_mm(L, R, P, seed=3):
    m, k, n = O.LRP2MM(L, R, P)
    rng = np.random.default_rng(seed)
    return O.mmcheck_q(L, R, P, rng.integers(-9, 9, m * k), rng.integers(-9, 9, k * n)) == 0


@pytest.mark.parametrize("stem", TRIPLES)
@pytest.mark.parametrize("only_sign", [False, True])
def test_negater_matches_oracle_and_keeps_the_algorithm(stem, only_sign):
    L, R, P = O.triple(stem)
    (gL, gR, gP), gst = capi.negater(L, R, P, only_sign)
    (eL, eR, eP), est = O.negater(L, R, P, only_sign)
    assert (gL, gR, gP) == (eL, eR, eP) and gst == est
    assert is_mm(gL, gR, gP)
    negs = lambda M: sum(1 for row in M for v in row if v < 0)
    assert [negs(gL), negs(gR), negs(gP)] == gst[6:9] and sum(gst[6:9]) <= sum(gst[3:6])
    assert gst[1] <= gst[0] and gst[9:12] == [sum(1 for row in M for v in row if v != 0) for M in (gL, gR, gP)]


@pytest.mark.parametrize("stem", TRIPLES)
def test_rotations(stem):
    L, R, P = O.triple(stem)
    m, k, n = O.LRP2MM(L, R, P)
    for right in (False, True):
        rc, (gL, gR, gP) = capi.rotater(L, R, P, right)
        assert rc == 0 and [gL, gR, gP] == O.rotater(L, R, P, right)
        assert O.LRP2MM(gL, gR, gP) == ((n, m, k) if right else (k, n, m)) or len({m, k, n}) < 3
        assert is_mm(gL, gR, gP)
    # three left rotations are the identity; left then right too
    t = (L, R, P)
    for _ in range(3):
        t = tuple(capi.rotater(*t, False)[1])
    assert list(t) == [L, R, P]
    back = capi.rotater(*capi.rotater(L, R, P, False)[1], True)[1]
    assert back == [L, R, P]


def test_rotater_dimension_mismatch():
    L, R, P = O.triple("2x2x2_7_Strassen")
    rc, _ = capi.rotater(L, [row[:3] for row in R], P)
    assert rc == 3
