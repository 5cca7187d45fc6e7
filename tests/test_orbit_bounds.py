"""CPU tests of the host-side magnitude bounds that select the lane packing of the orbit sweep kernels
(plo_orbit_magnitude_bounds; orbit_sweep.cu `magnitude_ok`): the bound must dominate every transformed entry of
L.(U^-1 (x) V), R.(V^-T (x) W), (U (x) W^-1).P (src/orbiter.cpp:284-294) -- checked here against numpy on candidates
decoded by the host decode (bit-identical to the device decode), adversarial extreme matrices included."""
import itertools

import numpy as np
import pytest

from plinopt_b200 import capi, hm

SEED = 0x504C494E4F505431


def transformed(mkn, Li, Ri, Pi, U, V, W):
    """Entries of the three products for one (U, V, W), rows as the kernels form them."""
    m, k, n = mkn
    r = Li.shape[0]
    Ui = np.rint(np.linalg.inv(U)).astype(np.int64)
    Vi = np.rint(np.linalg.inv(V)).astype(np.int64)
    Wi = np.rint(np.linalg.inv(W)).astype(np.int64)
    assert (Ui @ U == np.eye(m, dtype=np.int64)).all() and (Vi @ V == np.eye(k, dtype=np.int64)).all() and (Wi @ W == np.eye(n, dtype=np.int64)).all()
    A = Li.reshape(r, m, k).astype(np.int64)
    B = Ri.reshape(r, k, n).astype(np.int64)
    Cm = Pi.T.reshape(r, m, n).astype(np.int64)
    YL = np.einsum("ix,lij,jy->lxy", Ui, A, V)     # U^-T A V
    YR = np.einsum("xi,lij,jy->lxy", Vi, B, W)     # V^-1 B W
    YP = np.einsum("xi,lij,yj->lxy", U, Cm, Wi)    # U C W^-T
    return np.abs(YL).max(), np.abs(YR).max(), np.abs(YP).max()


def extreme_zoi(s, sign):
    """Unit upper triangular T with all strict-upper entries = sign: |T^-1| reaches the 2^(j-i-1) growth for sign = -1."""
    T = np.eye(s, dtype=np.int64)
    T[np.triu_indices(s, 1)] = sign
    return T


@pytest.mark.parametrize("stem", ["2x2x2_7_Winograd", "3x3x3_23_58", "4x4x4_49_156", "4x4x4_48_rational", "3x4x7_63_rational"])
def test_bounds_dominate_sampled_and_extreme_candidates(stem):
    L, R, P = hm.load_fixture(stem)
    mkn = hm.LRP2MM(L, R, P)
    m, k, n = mkn
    Li, Ri, Pi = (hm.scaled(M, np.int32)[0] for M in (L, R, P))
    lanes, (bl, br, bp) = capi.orbit_magnitude_bounds(mkn, Li, Ri, Pi)
    assert lanes in (1, 2, 4)
    worst = [0, 0, 0]
    for idx in range(300):
        U, V, W = capi.orbit_decode(m, k, n, capi.MODE_PHILOX, SEED, idx)
        got = transformed(mkn, Li, Ri, Pi, U.astype(np.int64), V.astype(np.int64), W.astype(np.int64))
        worst = [max(a, b) for a, b in zip(worst, got)]
    rng = np.random.default_rng(1)
    for su, sv, sw in itertools.product((-1, 1), repeat=3):  # extreme triangular factors under random permutations
        for _ in range(6):
            mats = []
            for s, sg in ((m, su), (k, sv), (n, sw)):
                T = extreme_zoi(s, sg)
                mats.append(np.eye(s, dtype=np.int64)[rng.permutation(s)] @ T @ np.eye(s, dtype=np.int64)[rng.permutation(s)])
            got = transformed(mkn, Li, Ri, Pi, *mats)
            worst = [max(a, b) for a, b in zip(worst, got)]
    assert worst[0] <= bl and worst[1] <= br and worst[2] <= bp, (worst, (bl, br, bp))


def test_rational_3x4x7_takes_the_four_lane_kernels():
    """BASELINE config 4: with the per-row bound every transformed entry of 3x4x7_63_rational provably fits a signed byte."""
    L, R, P = hm.load_fixture("3x4x7_63_rational")
    mkn = hm.LRP2MM(L, R, P)
    lanes, b = capi.orbit_magnitude_bounds(mkn, *(hm.scaled(M, np.int32)[0] for M in (L, R, P)))
    assert lanes == 4 and max(b) < 128, (lanes, b)
    L, R, P = hm.load_fixture("4x4x4_48_rational")  # P reaches 128 = one past a signed byte: sixteen-bit lanes
    lanes, b = capi.orbit_magnitude_bounds((4, 4, 4), *(hm.scaled(M, np.int32)[0] for M in (L, R, P)))
    assert lanes == 2 and b[2] == 128


def test_bound_values_on_a_hand_made_instance():
    """One full line of ones in A_0 against g(4) = (4, 2, 1, 1): 4 . 4; single entries in B_0, C_0: 1 . 4."""
    m = k = n = 4
    Li = np.zeros((1, m * k), np.int32); Li[0, :k] = 1
    Ri = np.zeros((1, k * n), np.int32); Ri[0, 0] = 1
    Pi = np.zeros((m * n, 1), np.int32); Pi[0, 0] = 1
    lanes, b = capi.orbit_magnitude_bounds((m, k, n), Li, Ri, Pi)
    assert b == (16, 4, 4) and lanes == 4
    I = np.eye(4, dtype=np.int64)
    T = extreme_zoi(4, -1)          # T^-1 has first row (1, 1, 2, 4)
    V = I.copy(); V[:, 0] = 1       # a {0,1} column of ones collects the whole row of A
    got = transformed((m, k, n), Li, Ri, Pi, T, V, I)
    assert got[0] <= 16 and got[1] <= 4 and got[2] <= 4
