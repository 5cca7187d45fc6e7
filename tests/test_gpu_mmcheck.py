"""GPU parity: batched MMchecker mod p vs the oracle's restatement of plinopt_library.inl:472-558."""
import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import hm

pytestmark = pytest.mark.gpu
P31 = 2147483647


@pytest.mark.parametrize("stem", ["2x2x2_7_Strassen", "2x2x2_7_DPS-accurate", "4x4x4_48_rational", "3x4x7_63_rational", "3x3x6_40"])
@pytest.mark.parametrize("p", [P31, 513083, 7])
def test_valid_triples_pass_and_corruption_is_caught(capi, stem, p):
    """Makefile:60-64 (mmcheck): every shipped triple is SUCCESS, also mod 513083 = (1013^2-3)/2."""
    if stem == "2x2x2_7_DPS-accurate" and p != 513083:
        pytest.skip("1013 is a placeholder for sqrt(3): only valid mod (1013^2-3)/2 = 513083 (Makefile:62-63)")
    L, R, P = O.triple(stem)
    mkn = hm.LRP2MM(L, R, P)
    m, k, n = mkn
    B = 40
    rng = np.random.default_rng(11)
    ua = rng.integers(0, p, (B, m * k)).astype(np.uint32); ub = rng.integers(0, p, (B, k * n)).astype(np.uint32)
    v, ok = capi.mmcheck_batch(p, mkn, len(L), hm.csr_modp(L, p), hm.csr_modp(R, p), hm.csr_modp(P, p), batch=B, ua=ua, ub=ub)
    exp = [O.mmcheck_modp(p, L, R, P, ua[b].astype(np.int64), ub[b].astype(np.int64)) for b in range(B)]
    assert v == 0 and ok.tolist() == [1] * B and exp == [0] * B
    # corrupt one entry: per-sample verdicts must match the oracle exactly
    L2 = [row[:] for row in L]; L2[1][0] += 1
    v, ok = capi.mmcheck_batch(p, mkn, len(L), hm.csr_modp(L2, p), hm.csr_modp(R, p), hm.csr_modp(P, p), batch=B, ua=ua, ub=ub)
    exp = [O.mmcheck_modp(p, L2, R, P, ua[b].astype(np.int64), ub[b].astype(np.int64)) for b in range(B)]
    assert ok.tolist() == [1 - e for e in exp]
    assert v == (1 if any(exp) else 0)
    if p > 1000:
        assert v == 1


def test_philox_samples_and_dimension_errors(capi):
    L, R, P = O.triple("4x4x4_48_rational")
    mkn = hm.LRP2MM(L, R, P)
    csr = [hm.csr_modp(M, P31) for M in (L, R, P)]
    v, ok = capi.mmcheck_batch(P31, mkn, len(L), *csr, seed=99, batch=257)
    assert v == 0 and ok.sum() == 257
    v, _ = capi.mmcheck_batch(P31, (4, 4, 5), len(L), *csr, seed=1, batch=4)
    assert v == 3  # plinopt_library.inl:494 outer dimension mismatch
    v, _ = capi.mmcheck_batch(P31, mkn, len(L) - 1, *csr, seed=1, batch=4)
    assert v == 2  # MMchecker.cpp:65-71 inner dimension mismatch


def test_random_dense_triple_linearity(capi):
    """Size-independent property on a large synthetic instance: the trivial algorithm
    (r = m*k*n elementary products) is correct for every sample; swapping two rows of P breaks it."""
    m, k, n = 6, 5, 7
    r = m * k * n
    Lr, Rr, Pr = [], [], [[] for _ in range(m * n)]
    t = 0
    for i in range(m):
        for a in range(k):
            for j in range(n):
                Lr.append((t, i * k + a)); Rr.append((t, a * n + j)); Pr[i * n + j].append(t); t += 1
    ones = lambda cnt: np.ones(cnt, dtype=np.uint32)
    Lc = (r, m * k, np.arange(r + 1, dtype=np.int64), np.array([c for _, c in Lr], dtype=np.int32), ones(r))
    Rc = (r, k * n, np.arange(r + 1, dtype=np.int64), np.array([c for _, c in Rr], dtype=np.int32), ones(r))
    pp = np.array([0] + list(np.cumsum([len(x) for x in Pr])), dtype=np.int64)
    Pc = (m * n, r, pp, np.array([c for row in Pr for c in row], dtype=np.int32), ones(int(pp[-1])))
    v, ok = capi.mmcheck_batch(P31, (m, k, n), r, Lc, Rc, Pc, seed=5, batch=100)
    assert v == 0 and ok.all()
    pc2 = Pc[3].copy(); pc2[0], pc2[k] = pc2[k], pc2[0]
    v, ok = capi.mmcheck_batch(P31, (m, k, n), r, Lc, Rc, (Pc[0], Pc[1], Pc[2], pc2, Pc[4]), seed=5, batch=100)
    assert v == 1 and not ok.any()


def test_regenerated_32x32x32_15096_passes_and_corruption_is_caught(capi):
    """BASELINE config 5: the 32x32x32_15096 triple (regenerated from the reference's .slp)
    is a correct MM algorithm for every Philox sample mod 2^31-1; flipping one residue of P breaks every sample
    whose evaluation touches it (plinopt_library.inl:472-558)."""
    big = hm.load_large_csr(P31)
    assert big is not None, "tests/golden/large/32x32x32_15096.npz missing"
    mkn, r, (L, R, P) = big
    for B in (1, 33, 256):
        v, ok = capi.mmcheck_batch(P31, mkn, r, L, R, P, seed=7, batch=B)
        assert v == 0 and ok.sum() == B
    val = P[4].copy(); val[12345] = (int(val[12345]) + 1) % P31
    v, ok = capi.mmcheck_batch(P31, mkn, r, L, R, (P[0], P[1], P[2], P[3], val), seed=7, batch=64)
    assert v == 1 and ok.sum() == 0
    # given samples: agree with a numpy evaluation mod a 20-bit prime
    p = 1000003
    _, _, (L2, R2, P2) = hm.load_large_csr(p)
    rng = np.random.default_rng(5)
    ua = rng.integers(0, p, (3, 1024)).astype(np.uint32); ub = rng.integers(0, p, (3, 1024)).astype(np.uint32)
    v, ok = capi.mmcheck_batch(p, mkn, r, L2, R2, P2, batch=3, ua=ua, ub=ub)
    assert v == 0 and ok.all()
    # mod a prime above 2^31 (two-step reduction in the kernels); a flipped residue of L is caught as well
    pb = 4294967291
    _, _, (L3, R3, P3) = hm.load_large_csr(pb)
    v, ok = capi.mmcheck_batch(pb, mkn, r, L3, R3, P3, seed=11, batch=40)
    assert v == 0 and ok.all()
    val = L3[4].copy(); val[777] = (int(val[777]) + 1) % pb
    v, ok = capi.mmcheck_batch(pb, mkn, r, (L3[0], L3[1], L3[2], L3[3], val), R3, P3, seed=11, batch=40)
    assert v == 1 and ok.sum() < 40


def test_plan_reports_its_encoding(capi):
    """plo_mmcheck_plan_encoding: on 32x32x32_15096 the block sums cut the X loads per sample at least six-fold, L and R use column
    blocks, P row blocks of stride n = 32; what the plan computes is unchanged (previous test)."""
    big = hm.load_large_csr(P31)
    mkn, r, (L, R, P) = big
    plan = capi.MMcheckPlan(P31, mkn, r, L, R, P, 64)
    enc = plan.encoding
    nnz = [len(x[3]) for x in (L, R, P)]
    assert all(6 * a < b for a, b in zip(enc["loads"], nnz))
    assert enc["col_stride"][0] == 1 and enc["col_stride"][1] == 1 and enc["row_stride"] == [0, 0, 32]
    plan.run(3, 0)
    v, ok = plan.result()
    assert v == 0 and ok.all()
    plan.close()


def _trivial_algorithm(m, k, n, rng=None, p=P31):
    """r = m*k*n elementary products (optionally scaled by random units a, b, 1/(ab)) as CSR triples."""
    r = m * k * n
    t = np.arange(r)
    i, a, j = t // (k * n), (t // n) % k, t % n
    la = rng.integers(1, p, r) if rng is not None else np.ones(r, dtype=np.int64)
    lb = rng.integers(1, p, r) if rng is not None else np.ones(r, dtype=np.int64)
    lc = np.array([pow(int(x) * int(y) % p, -1, p) for x, y in zip(la, lb)], dtype=np.int64) if rng is not None else np.ones(r, dtype=np.int64)
    Lc = (r, m * k, np.arange(r + 1, dtype=np.int64), (i * k + a).astype(np.int32), la.astype(np.uint32))
    Rc = (r, k * n, np.arange(r + 1, dtype=np.int64), (a * n + j).astype(np.int32), lb.astype(np.uint32))
    order = np.lexsort((t, i * n + j))  # rows of P = outputs (i, j), k products each
    Pc = (m * n, r, np.arange(m * n + 1, dtype=np.int64) * k, t[order].astype(np.int32), lc[order].astype(np.uint32))
    return r, Lc, Rc, Pc


@pytest.mark.parametrize("mkn,scaled", [((40, 30, 2), False), ((40, 30, 2), True), ((3, 400, 3), True)])
def test_wide_matrices_span_several_column_slabs(capi, mkn, scaled):
    """Operands wider than one 1024-column shared-memory slab (L: 1200 columns, P: 2400/3600 columns) take the
    partial-sum path; all-distinct random values take the plain (column, value) format, all-ones the grouped one."""
    m, k, n = mkn
    r, Lc, Rc, Pc = _trivial_algorithm(m, k, n, np.random.default_rng(1) if scaled else None)
    v, ok = capi.mmcheck_batch(P31, mkn, r, Lc, Rc, Pc, seed=3, batch=70)
    assert v == 0 and ok.all()
    val = Pc[4].copy(); val[-1] = (int(val[-1]) + 1) % P31
    v, ok = capi.mmcheck_batch(P31, mkn, r, Lc, Rc, (Pc[0], Pc[1], Pc[2], Pc[3], val), seed=3, batch=70)
    assert v == 1 and not ok.any()


@pytest.mark.parametrize("p", [2, 3, 65537, 4294967291])
@pytest.mark.parametrize("stem", ["2x2x2_7_Strassen", "3x3x3_23_58", "4x4x4_49_156"])
def test_extreme_moduli(capi, stem, p):
    """Smallest moduli (factors of 2 are stripped down to 2, src/MMchecker.cpp:123-126) and the largest 32-bit prime:
    per-sample verdicts equal the oracle's, also for a corrupted triple."""
    L, R, P = O.triple(stem)  # integer coefficients: valid modulo every p
    mkn = hm.LRP2MM(L, R, P)
    m, k, n = mkn
    B = 33
    rng = np.random.default_rng(p % 1000)
    ua = rng.integers(0, p, (B, m * k)).astype(np.uint32); ub = rng.integers(0, p, (B, k * n)).astype(np.uint32)
    for LL in (L, [[v + (1 if (i, j) == (2, 1) else 0) for j, v in enumerate(row)] for i, row in enumerate(L)]):
        v, ok = capi.mmcheck_batch(p, mkn, len(L), hm.csr_modp(LL, p), hm.csr_modp(R, p), hm.csr_modp(P, p), batch=B, ua=ua, ub=ub)
        exp = [O.mmcheck_modp(p, LL, R, P, ua[b].astype(np.int64), ub[b].astype(np.int64)) for b in range(B)]
        assert ok.tolist() == [1 - e for e in exp] and v == (1 if any(exp) else 0)
