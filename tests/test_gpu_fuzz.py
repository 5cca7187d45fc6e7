"""GPU: randomised differential tests against the oracle (seeded; ragged shapes, tiny and degenerate inputs).
  * sparsifier search: random TM (n = 1..9 rows, m = 1..70 columns), random previous rows (sometimes dependent), every offset,
    over Q and modulo small / large primes;
  * batched MMchecker: random sparse "triples" (not algorithms: most samples must FAIL identically), few distinct values
    (grouped format) or all distinct (plain format), operands wider than one slab."""
import math
from fractions import Fraction

import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import hm

pytestmark = pytest.mark.gpu


def _residue(v, p):
    return (v.numerator % p) * pow(v.denominator % p, -1, p) % p


@pytest.mark.parametrize("seed", range(24))
def test_lincomb_random_instances(capi, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 10)); m = int(rng.choice([1, 2, 5, 8, 9, 16, 17, 33, 48, 64, 65, 70]))
    p = int(rng.choice([0, 0, 3, 7, 101, 2147483647]))
    dens = [1, 1, 1, 2, 3, 4]
    if p:
        dens = [d for d in dens if d % p]
    TM = [[Fraction(int(rng.choice([0, 0, 0, 1, -1, 2, -3])), int(rng.choice(dens))) for _ in range(m)] for _ in range(n)]
    c = int(rng.choice([2, 3, 4, 5, 7]))
    if p:
        c = min(c, p)
    nblocks = (n + 3) // 4
    off = 4 * int(rng.integers(0, nblocks))
    nprev = int(rng.integers(0, n))  # rows already chosen (may be linearly dependent: then nothing is admissible)
    prev = [[Fraction(int(rng.integers(-2, 3))) for _ in range(n)] for _ in range(nprev)]
    if nprev >= 2 and rng.random() < 0.3:
        prev[-1] = [2 * v for v in prev[0]]
    init = (-1, -1) if rng.random() < 0.6 else (int(rng.integers(0, m + 1)), int(rng.integers(0, n + 1)))
    if p == 0:
        tm_num, tm_den = O.numden(TM)
        cf_num, cf_den = O.coeffs(TM, 0, c)
        lc = 1
        for d in cf_den:
            lc = lc * int(d) // math.gcd(lc, int(d))
        cf_int = np.array([int(a) * (lc // int(d)) for a, d in zip(cf_num, cf_den)], dtype=np.int64)
        tm_int = np.zeros((n, m), dtype=np.int64)
        for j in range(m):
            l = 1
            for i in range(n):
                l = l * TM[i][j].denominator // math.gcd(l, TM[i][j].denominator)
            for i in range(n):
                tm_int[i, j] = int(TM[i][j] * l)
        prev_int = np.array([[int(v) for v in row] for row in prev], dtype=np.int64).reshape(nprev, n)
        lcob_num = np.zeros((n, n), dtype=np.int64); lcob_den = np.ones((n, n), dtype=np.int64)
        lcob_num[:nprev] = prev_int
    else:
        tm_num = np.array([[_residue(v, p) for v in row] for row in TM], dtype=np.int64); tm_den = np.ones_like(tm_num)
        cf_num, cf_den = O.coeffs(tm_num.tolist(), p, c)
        cf_int = cf_num.copy(); tm_int = tm_num
        prev_int = np.array([[int(v) % p for v in row] for row in prev], dtype=np.int64).reshape(nprev, n)
        lcob_num = np.zeros((n, n), dtype=np.int64); lcob_den = np.ones((n, n), dtype=np.int64)
        lcob_num[:nprev] = prev_int
    num = nprev - off
    if num < 0 or num > 3 or off + num >= n:
        # the reference only ever searches row `off + num` with exactly the rows before it chosen: re-align the instance
        nprev = min(off, n - 1); num = 0
        prev_int = prev_int[:nprev] if nprev <= prev_int.shape[0] else np.vstack([prev_int, np.eye(n, dtype=np.int64)[prev_int.shape[0]:nprev]])
        lcob_num = np.zeros((n, n), dtype=np.int64); lcob_num[:nprev] = prev_int
    exp = O.lincomb_search(p, tm_num, tm_den, off, num, cf_num, cf_den, lcob_num, lcob_den, init[0], init[1])
    got = capi.lincomb_search(p, tm_int, off, cf_int, prev_int if nprev else None, init[0], init[1])
    assert (got[0], got[1], got[2]) == (exp[0], exp[1], None if exp[2] < 0 else exp[2]), (n, m, p, c, off, nprev, init)


@pytest.mark.parametrize("seed", range(12))
def test_mmcheck_random_sparse_triples(capi, seed):
    rng = np.random.default_rng(2000 + seed)
    p = int(rng.choice([3, 101, 513083, 2147483647, 4294967291]))
    m, k, n = (int(v) for v in rng.choice([(2, 2, 2), (3, 5, 2), (4, 4, 4), (40, 30, 2), (1, 1, 7), (2, 600, 2)]))
    r = int(rng.integers(1, 40))
    few = rng.random() < 0.5  # few distinct values -> grouped format; all distinct -> plain format
    palette = rng.integers(1, p, 3)

    def rand(rows, cols, dens):
        M = []
        for _ in range(rows):
            row = [0] * cols
            for j in rng.choice(cols, size=max(1, int(cols * dens)), replace=False):
                row[int(j)] = int(rng.choice(palette)) if few else int(rng.integers(1, p))
            M.append(row)
        return M
    L, R, P = rand(r, m * k, 0.3), rand(r, k * n, 0.3), rand(m * n, r, 0.5)
    B = int(rng.choice([1, 31, 32, 33, 70]))
    ua = rng.integers(0, p, (B, m * k)).astype(np.uint32); ub = rng.integers(0, p, (B, k * n)).astype(np.uint32)
    fr = lambda M: [[Fraction(v) for v in row] for row in M]
    v, ok = capi.mmcheck_batch(p, (m, k, n), r, hm.csr_modp(fr(L), p), hm.csr_modp(fr(R), p), hm.csr_modp(fr(P), p), batch=B, ua=ua, ub=ub)
    exp = [O.mmcheck_modp(p, fr(L), fr(R), fr(P), ua[b].astype(np.int64), ub[b].astype(np.int64)) for b in range(B)]
    assert ok.tolist() == [1 - e for e in exp] and v == (1 if any(exp) else 0)


@pytest.mark.parametrize("seed", range(10))
def test_factor_sweep_random_matrices(capi, seed):
    """Random tall matrices (possibly rank-deficient), every lane-group width (n = 1..32), inner dimensions n..r."""
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 13, 16, 17, 24, 32])); r = n + int(rng.integers(1, 12))
    k = n + int(rng.integers(0, r - n + 1))
    p = int(rng.choice([3, 101, 2147483647]))
    dens = [d for d in (1, 1, 2, 4) if d % p]
    M = [[Fraction(int(rng.choice([0, 0, 1, -1, 2])), int(rng.choice(dens))) for _ in range(n)] for _ in range(r)]
    if rng.random() < 0.2 and n > 1:
        for row in M:
            row[-1] = row[0]  # rank-deficient: no candidate can reach rank n
    A = np.array([[_residue(v, p) for v in row] for row in M], dtype=np.uint32)
    best, tab = capi.factor_sweep(p, A, k, 77, 3, 3 + 150, table=True)
    ref = O.factor_sweep(M, k, 77, 3, 3 + 150, p=p)
    assert np.array_equal(tab, ref["table"]) and best == ref["best"], (r, n, k, p)


@pytest.mark.parametrize("seed", range(8))
def test_dependency_random_matrices(capi, seed):
    rng = np.random.default_rng(4000 + seed)
    r = int(rng.integers(2, 14)); n = int(rng.choice([1, 2, 3, 4, 7, 16, 17, 33, 64]))
    q = int(rng.choice([0, 0, 7, 101]))
    dens = [d for d in (1, 1, 2, 3) if not q or d % q]
    M = [[Fraction(int(rng.choice([0, 0, 0, 1, -1, 2])), int(rng.choice(dens))) for _ in range(n)] for _ in range(r)]
    level = int(rng.integers(1, 5)); c = int(rng.integers(1, 6))
    got = capi.depender(M, level, c, q=q)
    ref = O.depender(M, level, c, p=q)
    assert got["coeffs"] == ref["coeffs"] and got["ncand"] == ref["ncand"] and got["hits"] == ref["hits"], (r, n, q, level, c)


@pytest.mark.parametrize("seed", range(10))
def test_orbit_random_triples(capi, seed):
    """Random (L, R, P) of every compiled shape -- not algorithms, scoring does not care -- small or large entries (packed int32 /
    plain int32 / 64-bit paths), with denominators, over Q and modulo p: per-candidate tables against the oracle."""
    rng = np.random.default_rng(5000 + seed)
    m, k, n = (int(v) for v in rng.choice([(2, 2, 2), (3, 3, 3), (4, 4, 4), (3, 4, 7), (3, 3, 6), (3, 6, 3), (6, 3, 3)]))
    r = int(rng.integers(1, 9))
    scale = int(rng.choice([1, 3, 1000, 40000]))  # 1000+: beyond the 16-bit lanes; 40000: beyond the int32 product bound
    den = int(rng.choice([1, 2, 6]))
    mat = lambda rows, cols: [[Fraction(int(rng.integers(-2, 3)) * scale, den) for _ in range(cols)] for _ in range(rows)]
    L, R, P = mat(r, m * k), mat(r, k * n), mat(m * n, r)
    (Li, dl), (Ri, dr), (Pi, dp) = O.scaled_int(L), O.scaled_int(R), O.scaled_int(P)
    cnt = 400
    ref = O.orbit_sweep(L, R, P, 3, 1, 9, 50, 50 + cnt)
    try:
        nnz, nno, g2 = capi.orbit_table((m, k, n), Li.astype(np.int32), Ri.astype(np.int32), Pi.astype(np.int32), (dl, dr, dp), 1, 9, 50, 50 + cnt)
    except capi.PloError as e:  # every compiled shape has a 64-bit kernel: nothing of this size may be refused
        raise AssertionError((m, k, n, scale, str(e)))
    assert np.array_equal(nnz, ref["nnz"]) and np.array_equal(nno, ref["nno"])
    assert np.allclose(g2, ref["g2"], rtol=1e-12, atol=0)
    p = int(rng.choice([3, 101, 2147483647]))
    if den % p and scale % p:
        red = lambda M: np.array([[_residue(v, p) for v in row] for row in M], dtype=np.int32)
        refp = O.orbit_sweep(L, R, P, 0, 1, 9, 50, 50 + cnt, p=p)
        nz, no = capi.orbit_table_modp(p, (m, k, n), red(L), red(R), red(P), 1, 9, 50, 50 + cnt)
        assert np.array_equal(nz, refp["nnz"]) and np.array_equal(no, refp["nno"])


@pytest.mark.parametrize("seed", range(16))
def test_orbit_sweep_kernels_against_the_table_kernel(capi, seed):
    """The sweep kernels (two-lane / four-lane / table-driven / three-phase / 64-bit, chosen by shape, rank and magnitudes) against the
    plain per-candidate table kernel on random triples: every 16-candidate window must elect the table's lexicographic minimum with the
    lowest index, for both measures, and survivor compaction must return the table's sub-threshold set.  Seeds 0-5 force the
    2x2x2, r = 7, small-integer case of the headline kernel."""
    rng = np.random.default_rng(7000 + seed)
    if seed < 6:
        (m, k, n), r, scale, den = (2, 2, 2), 7, 1, 1
    else:
        m, k, n = (int(v) for v in rng.choice([(2, 2, 2), (3, 3, 3), (4, 4, 4), (3, 4, 7), (3, 3, 6), (3, 6, 3), (6, 3, 3)]))
        r = int(rng.integers(1, 12)); scale = int(rng.choice([1, 1, 3, 1000, 40000])); den = int(rng.choice([1, 1, 2, 6]))
    hi_v = 2 if seed % 2 else 3
    mat = lambda rows, cols: [[Fraction(int(rng.integers(-hi_v + 1, hi_v)) * scale, den) for _ in range(cols)] for _ in range(rows)]
    L, R, P = mat(r, m * k), mat(r, k * n), mat(m * n, r)
    (Li, dl), (Ri, dr), (Pi, dp) = O.scaled_int(L), O.scaled_int(R), O.scaled_int(P)
    Li, Ri, Pi = Li.astype(np.int32), Ri.astype(np.int32), Pi.astype(np.int32)
    lo, cnt, win = 1 << 20, 320, 16
    mode = int(seed % 3 != 2)  # mostly Philox, sometimes the mixed-radix enumeration (where the space fits 64 bits)
    if mode == 0 and capi.orbit_space(m, k, n) <= (1 << 20) + 320:
        mode = 1
    nnz, nno, g2 = capi.orbit_table((m, k, n), Li, Ri, Pi, (dl, dr, dp), mode, 77, lo, lo + cnt)
    for measure in (capi.MEASURE_NNZ, capi.MEASURE_G2):
        plan = capi.OrbitPlan((m, k, n), Li, Ri, Pi, (dl, dr, dp), measure, mode, 77)
        for a in range(0, cnt, win):
            plan.run(lo + a, lo + a + win)
            got = plan.result()
            sl = slice(a, a + win)
            if measure == capi.MEASURE_NNZ:
                keys = [(int(x), int(y)) for x, y in zip(nnz[sl], nno[sl])]
                j = keys.index(min(keys))
                assert (got["index"], got["nnz"], got["nno"]) == (lo + a + j, keys[j][0], keys[j][1]), (m, k, n, r, scale, den, a)
            else:
                j = int(np.argmin(g2[sl]))
                assert got["index"] == lo + a + j and got["score"] == g2[sl][j], (m, k, n, r, scale, den, a)
        thr_n = int(np.sort(nnz)[cnt // 10]); thr_g = float(np.sort(g2)[cnt // 10])
        sv = plan.survivors(lo, lo + cnt, nnz=thr_n, nno=0xFFFFFFFF, score=thr_g)
        keep = (nnz <= thr_n) if measure == capi.MEASURE_NNZ else (g2 <= thr_g)
        assert [s["index"] - lo for s in sv] == np.nonzero(keep)[0].tolist()
        plan.close()
