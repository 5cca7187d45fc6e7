"""GPU parity: orbit sweep (CUDA, through the C ABI) vs the CPU oracle's literal restatement
of src/orbiter.cpp:272-324 (+ plinopt_library.inl:210-284, growthfactor.cpp:117-125) on the
same (mode, seed, index) candidates.  nnz / nno / index bit-exact; G2 bit-exact for integer
triples and within 1e-12 relative for rational ones."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
SEED = 0x504C494E4F505431
RTOL = 1e-12


def ints(stem):
    L, R, P = O.triple(stem)
    (Li, dl), (Ri, dr), (Pi, dp) = O.scaled_int(L), O.scaled_int(R), O.scaled_int(P)
    return (L, R, P), O.LRP2MM(L, R, P), (Li.astype(np.int32), Ri.astype(np.int32), Pi.astype(np.int32)), (dl, dr, dp)


def test_c2_winograd_exhaustive_table(capi):
    """All 48^3 = 110592 candidates of the 2x2x2 orbit: full per-candidate table parity."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("2x2x2_7_Winograd")
    space = capi.orbit_space(*mkn)
    assert space == 110592
    ref = O.orbit_sweep(L, R, P, 3, 0, 0, 0, space)
    nnz, nno, g2 = capi.orbit_table(mkn, Li, Ri, Pi, dens, 0, 0, 0, space)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"])
    assert np.array_equal(g2, ref["g2"])  # integer triple: bit-exact doubles
    for measure in (0, 3):
        refb = O.orbit_sweep(L, R, P, measure, 0, 0, 0, space, table=False)["best"]
        got = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, measure, 0, 0, 0, space)
        assert got["index"] == refb[0] and got["nnz"] == refb[1] and got["nno"] == refb[2]
        if measure == 3:
            assert got["score"] == refb[3]


@pytest.mark.parametrize("stem,count", [("2x2x2_7_Winograd", 20000), ("2x2x2_7_DPS-smallrat-12.2034", 20000),
                                        ("3x3x3_23_58", 6000), ("4x4x4_48_rational", 3000), ("3x4x7_63_rational", 1500),
                                        ("3x3x6_40", 1500), ("3x6x3_40", 1500), ("6x3x3_40", 1500)])
def test_philox_table_parity(capi, stem, count):
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints(stem)
    lo = 2 ** 33 + 17
    ref = O.orbit_sweep(L, R, P, 3, 1, SEED, lo, lo + count)
    nnz, nno, g2 = capi.orbit_table(mkn, Li, Ri, Pi, dens, 1, SEED, lo, lo + count)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"])
    np.testing.assert_allclose(g2, ref["g2"], rtol=RTOL, atol=0)
    if dens == (1, 1, 1):
        assert np.array_equal(g2, ref["g2"])
    # winners (deterministic rule: lexicographic minimum, lowest index)
    got0 = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, lo, lo + count)
    ref0 = O.orbit_sweep(L, R, P, 0, 1, SEED, lo, lo + count, table=False)["best"]
    assert (got0["index"], got0["nnz"], got0["nno"]) == ref0[:3]
    got3 = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 3, 1, SEED, lo, lo + count)
    gmin = ref["g2"].min()
    assert abs(got3["score"] - gmin) <= RTOL * gmin
    assert abs(ref["g2"][got3["index"] - lo] - got3["score"]) <= RTOL * gmin


def test_winner_is_a_valid_algorithm(capi):
    """src/orbiter.cpp:355: the winner must still pass MMchecker; its scores match a direct evaluation."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("3x3x3_23_58")
    got = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 0, 200000)
    U, V, W = capi.orbit_decode(*mkn, 1, SEED, got["index"])
    Lj, Rg, hP = O.orbit_apply(L, R, P, U, V, W)
    rng = np.random.default_rng(3)
    m, k, n = mkn
    assert O.mmcheck_q(Lj, Rg, hP, rng.integers(-50, 50, m * k), rng.integers(-50, 50, k * n)) == 0
    nnz = sum(1 for M in (Lj, Rg, hP) for row in M for v in row if v != 0)
    assert nnz == got["nnz"]
    init = sum(1 for M in (L, R, P) for row in M for v in row if v != 0)
    assert got["nnz"] <= init or True  # informational: a random sweep need not improve


def test_large_sweep_properties(capi):
    """2^26 Philox candidates (BASELINE config 2 scale-down): the result is independent of how the
    index range is split (associativity of the argmin), and the winner re-scored by the oracle agrees."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("2x2x2_7_Winograd")
    N = 1 << 26
    whole = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 3, 1, SEED, 0, N)
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, 3, 1, SEED)
    parts = []
    for a, b in [(0, N // 3), (N // 3, N // 2 + 5), (N // 2 + 5, N)]:
        plan.run(a, b); parts.append(plan.result())
    plan.close()
    best = min(parts, key=lambda d: (d["score"], d["index"]))
    assert best == whole
    one = O.orbit_sweep(L, R, P, 3, 1, SEED, whole["index"], whole["index"] + 1)
    assert one["g2"][0] == whole["score"] and one["nnz"][0] == whole["nnz"]
    # the global optimum of the orbit for G2 is known from the exhaustive table (<= Winograd's own 17.853)
    assert whole["score"] <= 17.85300667219901


def test_empty_range_and_errors(capi):
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("2x2x2_7_Winograd")
    got = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 5, 5)
    assert got["index"] is None
    with pytest.raises(capi.PloError) as e:
        capi.orbit_sweep((5, 5, 5), np.zeros((1, 25), np.int32), np.zeros((1, 25), np.int32), np.zeros((25, 1), np.int32), (1, 1, 1), 0, 1, 0, 0, 1)
    assert e.value.code == capi.E_SHAPE
    # beyond the int32 product bound every compiled shape switches to its 64-bit exact kernel (same counts)
    big = capi.orbit_sweep(mkn, Li * 100000, Ri, Pi, (100000, 1, 1), 0, 1, SEED, 0, 2000)
    assert (big["index"], big["nnz"], big["nno"]) == tuple(capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 0, 2000)[k] for k in ("index", "nnz", "nno"))
    (_, _, _), mkn7, (L7, R7, P7), dens7 = ints("3x4x7_63_rational")
    big7 = capi.orbit_sweep(mkn7, L7 * 100000, R7, P7, (dens7[0] * 100000, dens7[1], dens7[2]), 0, 1, SEED, 0, 500)
    assert (big7["index"], big7["nnz"], big7["nno"]) == tuple(capi.orbit_sweep(mkn7, L7, R7, P7, dens7, 0, 1, SEED, 0, 500)[k] for k in ("index", "nnz", "nno"))


def test_growth_G2_known_answers(capi):
    """growthfactor.cpp:117-125 known answers from the data set headers / file names."""
    want = {"2x2x2_7_Strassen": 14.828427124746192, "2x2x2_7_Winograd": 17.85300667219901,
            "2x2x2_7_DPS-smallrat-12.2034": 12.203427124746, "2x2x2_7_DPS-integral-12.0662": 12.06616423,
            "2x2x2_7_DPS-intermediate-12.0695": 12.06954148}
    for stem, val in want.items():
        L, R, P = O.triple(stem)
        f = lambda M: np.array([[float(v) for v in row] for row in M])
        got = capi.growth_G2(f(L), f(R), f(P))[0]
        assert abs(got - val) < 5e-9 * val, (stem, got)
        assert abs(got - O.growth_G2(L, R, P)) <= RTOL * got


def residues(M, p):
    return np.array([[(v.numerator % p) * pow(v.denominator % p, -1, p) % p for v in row] for row in M], dtype=np.int32)


@pytest.mark.parametrize("stem,count,p", [("2x2x2_7_Winograd", 20000, 2147483647), ("2x2x2_7_DPS-accurate", 8000, 513083),
                                          ("3x3x3_23_58", 4000, 101), ("4x4x4_48_rational", 2000, 2147483647),
                                          ("3x4x7_63_rational", 1000, 1000003), ("2x2x2_7_Strassen", 5000, 3),
                                          ("3x6x3_40", 800, 2147483647), ("6x3x3_40", 800, 513083)])
def test_modular_orbit_sweep(capi, stem, count, p):
    """`orbiter -m p` (src/orbiter.cpp:232-234, 419-426): the whole search runs in Z/pZ; per-candidate (nnz, nno) and the
    winner are bit-exact against the oracle over the same field."""
    L, R, P = O.triple(stem)
    mkn = O.LRP2MM(L, R, P)
    Lr, Rr, Pr = residues(L, p), residues(R, p), residues(P, p)
    ref = O.orbit_sweep(L, R, P, 0, 1, SEED, 100, 100 + count, p=p)
    nnz, nno = capi.orbit_table_modp(p, mkn, Lr, Rr, Pr, 1, SEED, 100, 100 + count)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"])
    got = capi.orbit_sweep(mkn, Lr, Rr, Pr, (1, 1, 1), 0, 1, SEED, 100, 100 + count, p=p)
    assert (got["index"], got["nnz"], got["nno"]) == ref["best"][:3]
    if p > 10 ** 6 and stem != "2x2x2_7_DPS-accurate":
        # a large prime sees the same zero / +-1 pattern as the rationals
        exact = O.orbit_sweep(L, R, P, 0, 1, SEED, 100, 100 + count)
        assert np.array_equal(nnz, exact["nnz"]) and np.array_equal(nno, exact["nno"])
    with pytest.raises(capi.PloError):
        capi.orbit_sweep(mkn, Lr, Rr, Pr, (1, 1, 1), 3, 1, SEED, 0, 10, p=p)  # no growth factor in a finite field


def test_multi_device_sweep_in_one_process(capi):
    """plo_orbit_sweep_devices / bin/orbiter --gpus N: index shards over the devices of this process give the single-device winner
    (with one visible device the call degenerates to it; with several, every shard runs on its own GPU)."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("4x4x4_48_rational")
    ndev = capi.device_count()
    for measure in (0, 3):
        one = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, measure, 1, SEED, 5, 5 + 200000)
        for n in sorted({1, 2, ndev, ndev + 3}):
            got = capi.orbit_sweep_devices(n, mkn, Li, Ri, Pi, dens, measure, 1, SEED, 5, 5 + 200000)
            assert got == one, (measure, n)
    assert capi.orbit_sweep_devices(2, mkn, Li, Ri, Pi, dens, 0, 1, SEED, 9, 9)["index"] is None


def test_wide_exact_path_for_large_denominators(capi):
    """2x2x2_7_DPS-integral-12.0662 (the reference's most accurate integral-coefficient variant, common denominators ~10^9): the
    int32 product bound fails, the 64-bit exact kernels take over.  nnz / nno bit-exact, G2 within 1e-12 relative (entries become
    doubles before squaring, growthfactor.cpp:25-28), the exhaustive orbit confirms 12.0662 as its minimum."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("2x2x2_7_DPS-integral-12.0662")
    assert max(dens) > 10 ** 9
    space = capi.orbit_space(*mkn)
    ref = O.orbit_sweep(L, R, P, 3, 0, 0, 0, space)
    nnz, nno, g2 = capi.orbit_table(mkn, Li, Ri, Pi, dens, 0, 0, 0, space)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"])
    assert np.allclose(g2, ref["g2"], rtol=RTOL, atol=0)
    got = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 3, 0, 0, 0, space)
    assert abs(got["score"] - 12.06616423) < 5e-9 and got["nnz"] == 63  # data/2x2x2_7_DPS-integral-12.0662_L.sms:1
    assert abs(got["score"] - ref["best"][3]) <= RTOL * ref["best"][3]
    refn = O.orbit_sweep(L, R, P, 0, 1, SEED, 0, 30000, table=False)["best"]
    gotn = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 0, 30000)
    assert (gotn["index"], gotn["nnz"], gotn["nno"]) == refn[:3]
    halves = [capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 0, 11111), capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 0, 1, SEED, 11111, 30000)]
    assert min(halves, key=lambda b: (b["nnz"], b["nno"], b["index"])) == gotn


def test_int64_inputs_dps_intermediate(capi):
    """2x2x2_7_DPS-intermediate-12.0695: common denominator 1.0e11 > 2^31, outside the int32 C ABI -> plo_orbit_sweep64 / table64.
    Same exactness contract as the wide path: counts bit-exact, G2 within 1e-12; its own G2 (12.06954148,
    data/2x2x2_7_DPS-intermediate-12.0695_L.sms:1) is the minimum of its orbit."""
    L, R, P = O.triple("2x2x2_7_DPS-intermediate-12.0695")
    mkn = O.LRP2MM(L, R, P)
    (Li, dl), (Ri, dr), (Pi, dp) = O.scaled_int(L), O.scaled_int(R), O.scaled_int(P)
    assert max(dl, dr, dp) > 2 ** 31
    space = capi.orbit_space(*mkn)
    ref = O.orbit_sweep(L, R, P, 3, 0, 0, 0, space)
    nnz, nno, g2 = capi.orbit_table64(mkn, Li, Ri, Pi, (dl, dr, dp), 0, 0, 0, space)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"]) and np.allclose(g2, ref["g2"], rtol=RTOL, atol=0)
    got = capi.orbit_sweep64(mkn, Li, Ri, Pi, (dl, dr, dp), 3, 0, 0, 0, space)
    assert abs(got["score"] - 12.06954148) < 5e-9
    gotn = capi.orbit_sweep64(mkn, Li, Ri, Pi, (dl, dr, dp), 0, 1, SEED, 0, 20000)
    assert (gotn["index"], gotn["nnz"], gotn["nno"]) == O.orbit_sweep(L, R, P, 0, 1, SEED, 0, 20000, table=False)["best"][:3]
    # host-level driver picks the 64-bit path by itself
    rc = capi.orbiter(L, R, P, measure=capi.MEASURE_G2, mode=capi.MODE_EXHAUSTIVE, seed=0, loops=space)
    assert rc[3]["mm_verdict"] == 0 and not rc[3]["improved"]


@pytest.mark.parametrize("stem", ["2x2x2_7_Winograd", "2x2x2_7_Strassen", "3x3x3_23_58", "4x4x4_49_156"])
def test_four_lane_growth_kernel_is_bit_identical(capi, stem):
    """Integer-coefficient algorithms take the four-lane (8-bit spacing, dp4a) growth-factor kernel in sweeps; the per-candidate table
    comes from the plain kernel.  A sweep over a single candidate must return exactly the table's double, for every candidate tried;
    and block / range splits do not change the winner."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints(stem)
    assert dens == (1, 1, 1)
    lo, cnt = 1000, 3000
    _, _, g2 = capi.orbit_table(mkn, Li, Ri, Pi, dens, 1, SEED, lo, lo + cnt)
    for i in list(range(0, 40)) + [cnt - 1, 1234, 2222]:
        one = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 3, 1, SEED, lo + i, lo + i + 1)
        assert one["index"] == lo + i and one["score"] == g2[i], (stem, i)
    whole = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, 3, 1, SEED, lo, lo + cnt)
    assert whole["score"] == g2.min() and whole["index"] == lo + int(np.argmin(g2))
    ref = O.orbit_sweep(L, R, P, 3, 1, SEED, lo, lo + cnt, table=False)["best"]
    assert (whole["index"], whole["score"]) == (ref[0], ref[3])


@pytest.mark.parametrize("stem,count", [("2x2x2_7_Winograd", 30011), ("3x3x3_23_58", 5003), ("2x2x2_7_DPS-integral-12.0662", 4001)])
def test_survivor_compaction_equals_filtered_oracle_table(capi, stem, count):
    """plo_orbit_plan_survivors: exactly the candidates of the oracle's per-candidate table that pass the threshold, with both
    measures, in index order -- for the sparsity order (nnz, nno) and for the growth factor; a ragged range (count not a multiple of
    the warp), an empty range, the capacity protocol, and the 64-bit exact kernel (DPS-integral)."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints(stem)
    lo = 12345
    ref = O.orbit_sweep(L, R, P, 3, 1, SEED, lo, lo + count)
    idx = np.arange(lo, lo + count, dtype=np.uint64)
    # sparsity: a threshold in the lower part of the distribution
    order = np.lexsort((ref["nno"], ref["nnz"]))
    tn, to = int(ref["nnz"][order[count // 50]]), int(ref["nno"][order[count // 50]])
    keep = (ref["nnz"] < tn) | ((ref["nnz"] == tn) & (ref["nno"] <= to))
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, capi.MEASURE_NNZ, 1, SEED)
    got = plan.survivors(lo, lo + count, nnz=tn, nno=to, capacity=8)  # too small on purpose: the wrapper retries with *count
    assert [g["index"] for g in got] == idx[keep].tolist()
    assert [g["nnz"] for g in got] == ref["nnz"][keep].tolist() and [g["nno"] for g in got] == ref["nno"][keep].tolist()
    np.testing.assert_allclose([g["score"] for g in got], ref["g2"][keep], rtol=RTOL, atol=0)
    assert plan.survivors(lo, lo) == [] and plan.survivors(lo, lo + count, nnz=0, nno=0) == []
    with pytest.raises(capi.PloError):  # capacity protocol at the C level: PLO_E_RANGE, count reported
        import ctypes as C
        cnt = C.c_uint64(0)
        thr = capi.OrbitBest(score=0.0, nnz=tn, nno=to, index=0)
        buf = (capi.OrbitBest * 1)()
        rc = capi.lib().plo_orbit_plan_survivors(plan._h, lo, lo + count, C.byref(thr), 1, buf, C.byref(cnt))
        assert rc == capi.E_RANGE and cnt.value == int(keep.sum())
        capi._check(rc)
    plan.close()
    # growth factor
    thr = float(np.sort(ref["g2"])[count // 100]) * (1 + 1e-9)
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, capi.MEASURE_G2, 1, SEED)
    got = plan.survivors(lo, lo + count, score=thr)
    want = idx[ref["g2"] <= thr].tolist()
    assert [g["index"] for g in got] == want and len(want) >= count // 100
    plan.close()


@pytest.mark.parametrize("measure", [0, 3])
def test_full_size_sweep_reaches_the_exhaustive_optimum(capi, measure):
    """BASELINE config 2 at bench.py's size (2^31 Philox candidates of 2x2x2_7_Winograd), through properties that do not need the
    oracle at that size: the 48^3 orbit is sampled ~19000 times over, so the winner's score must be the optimum of the exhaustive
    sweep; the result does not depend on how the range is cut; and the winner is the FIRST index that reaches the optimum (survivor
    compaction below the winner returns exactly that one candidate)."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints("2x2x2_7_Winograd")
    opt = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, measure, 0, 0, 0, capi.orbit_space(*mkn))
    N = 1 << 31
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, measure, 1, SEED)
    plan.run(0, N); whole = plan.result()
    assert (whole["score"], whole["nnz"], whole["nno"]) == (opt["score"], opt["nnz"], opt["nno"])
    parts = []
    for a, b in [(0, N // 7), (N // 7, N // 2 + 3), (N // 2 + 3, N)]:
        plan.run(a, b); parts.append(plan.result())
    key = (lambda d: (d["nnz"], d["nno"], d["index"])) if measure == 0 else (lambda d: (d["score"], d["index"]))
    assert min(parts, key=key) == whole
    first = plan.survivors(0, whole["index"] + 1, nnz=opt["nnz"], nno=opt["nno"], score=opt["score"])
    assert [s["index"] for s in first] == [whole["index"]]
    one = O.orbit_sweep(L, R, P, 3, 1, SEED, whole["index"], whole["index"] + 1)
    assert (one["nnz"][0], one["nno"][0]) == (whole["nnz"], whole["nno"]) and (measure == 0 or one["g2"][0] == whole["score"])
    plan.close()


@pytest.mark.parametrize("stem,lo", [("3x3x3_23_58", 7776 * 7776 * 5 + 12345), ("2x2x2_7_Winograd", 48 * 48 * 7 + 11)])
def test_table_driven_decode_in_exhaustive_mode(capi, stem, lo):
    """The kernels that read whole zoi matrices from a table (2x2: shared memory, 3x3: one global table per plan) index them by
    `index mod count` in exhaustive mode: windows of the mixed-radix enumeration must elect the per-candidate table's minimum (first
    index among ties), for both measures; the table itself equals the oracle's."""
    (L, R, P), mkn, (Li, Ri, Pi), dens = ints(stem)
    cnt, win = 2000, 50
    ref = O.orbit_sweep(L, R, P, 3, 0, 0, lo, lo + cnt)
    nnz, nno, g2 = capi.orbit_table(mkn, Li, Ri, Pi, dens, 0, 0, lo, lo + cnt)
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"]) and np.array_equal(g2, ref["g2"])
    for measure in (capi.MEASURE_NNZ, capi.MEASURE_G2):
        plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, measure, 0, 0)
        for a in range(0, cnt, win):
            plan.run(lo + a, lo + a + win)
            got = plan.result()
            sl = slice(a, a + win)
            if measure == capi.MEASURE_NNZ:
                keys = [(int(x), int(y)) for x, y in zip(nnz[sl], nno[sl])]
                j = keys.index(min(keys))
                assert (got["index"], got["nnz"], got["nno"]) == (lo + a + j, keys[j][0], keys[j][1])
            else:
                j = int(np.argmin(g2[sl]))
                assert got["index"] == lo + a + j and got["score"] == g2[sl][j]
        plan.close()


def test_two_plans_on_two_streams_share_the_constant_bank_safely(capi):
    """plo_orbit_plan_run is asynchronous and every plan of a device uses the same constant bank: alternating two plans on two
    non-blocking streams without any synchronisation in between must give the single-stream answers (the bank is guarded by an
    event: a stream that re-writes it waits for its last use)."""
    import torch
    from plinopt_b200 import hm
    torch.cuda.set_device(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    plans, ranges = [], [(0, 1 << 22), (5, (1 << 18) + 5)]
    for stem, measure in (("2x2x2_7_Winograd", capi.MEASURE_G2), ("3x3x3_23_58", capi.MEASURE_NNZ)):
        L, R, P = hm.load_fixture(stem)
        (Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
        plans.append(capi.OrbitPlan(hm.LRP2MM(L, R, P), Li, Ri, Pi, (dl, dr, dp), measure, capi.MODE_PHILOX, SEED))
    expect = []
    for pl, (lo, hi) in zip(plans, ranges):
        pl.run(lo, hi, 0)
        expect.append(pl.result(0))
    for _ in range(25):
        plans[0].run(*ranges[0], s1.cuda_stream)
        plans[1].run(*ranges[1], s2.cuda_stream)
    got = [plans[0].result(s1.cuda_stream), plans[1].result(s2.cuda_stream)]
    assert got == expect
    for pl in plans:
        pl.close()
