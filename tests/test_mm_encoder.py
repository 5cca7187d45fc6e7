"""The MMchecker plan's matrix encoder (column / row block sums, plain pairs, value groups; plinopt_b200/csrc/mmcheck.cu) without a
device: plo_mmcheck_encode_check encodes a CSR matrix exactly as plo_mmcheck_plan_create does, replays the encoded blobs on the CPU
for one sample and must return A.x mod p (include/plinopt_library.inl:504-509 is a plain sparse product)."""
import numpy as np
import pytest

from plinopt_b200 import capi, hm

P31 = 2147483647
PBIG = 4294967291  # the largest prime below 2^32: the two-step reduction path


def csr_from_dense(a):
    rows, cols = a.shape
    ptr, col, val = [0], [], []
    for i in range(rows):
        nz = np.nonzero(a[i])[0]
        col += nz.tolist()
        val += [int(a[i, c]) for c in nz]
        ptr.append(len(col))
    return (rows, cols, np.array(ptr, dtype=np.int64), np.array(col, dtype=np.int32), np.array(val, dtype=np.uint32))


def product(A, x, p):
    rows, cols, ptr, col, val = A
    out = np.zeros(rows, dtype=np.uint64)
    for i in range(rows):
        s = 0
        for t in range(ptr[i], ptr[i + 1]):
            s += int(val[t]) * int(x[col[t]])
        out[i] = s % p
    return out


def check(A, p, seed=0, **kw):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, p, A[1], dtype=np.uint64).astype(np.uint32)
    y, st = capi.mmcheck_encode_check(p, A, x, **kw)
    assert (y.astype(np.uint64) == product(A, x, p)).all()
    return st


@pytest.mark.parametrize("p", [7, P31, PBIG])
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (40, 64), (33, 1024), (17, 1030), (64, 2500)])
def test_random_matrices(p, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for density, nvals in ((0.05, None), (0.5, 3), (0.9, 1), (1.0, None)):
        a = rng.integers(1, p, shape, dtype=np.uint64) if nvals is None else rng.choice(rng.integers(1, p, nvals, dtype=np.uint64), shape)
        a = np.where(rng.random(shape) < density, a, 0)
        for rb in (False, True):
            st = check(csr_from_dense(a), p, row_blocks=rb)
            assert st["loads"] <= max(int((a != 0).sum()), 0) + 4  # plain padding never adds more than a pair per row part... and block sums only pay


def test_empty_and_zero_entries_and_duplicates():
    # an all-zero matrix, explicit zeros, and a CSR with a repeated column (the entries add up)
    A = (3, 5, np.array([0, 0, 0, 0], dtype=np.int64), np.array([], dtype=np.int32), np.array([], dtype=np.uint32))
    assert check(A, P31)["tasks"] == 0
    B = (2, 4, np.array([0, 3, 5], dtype=np.int64), np.array([1, 1, 3, 0, 2], dtype=np.int32), np.array([5, P31 - 5, 0, 7, 7], dtype=np.uint32))
    rng = np.random.default_rng(3)
    x = rng.integers(0, P31, 4).astype(np.uint32)
    y, st = capi.mmcheck_encode_check(P31, B, x)
    assert y[0] == 0 and y[1] == (7 * int(x[0]) + 7 * int(x[2])) % P31 and st["tasks"] == 1


@pytest.mark.parametrize("cs", [1, 2, 8, 64])
def test_column_block_structure_is_found_at_any_stride(cs):
    """rows that carry one value on blocks {hi*16*cs + t*cs + lo}: the encoder must pick that stride and read ~1/16 of the entries"""
    rng = np.random.default_rng(cs)
    rows, cols = 48, 1024 + 256
    a = np.zeros((rows, cols), dtype=np.uint64)
    for i in range(rows):
        for _ in range(6):
            base = int(rng.integers(0, cols // (16 * cs))) * 16 * cs + int(rng.integers(0, cs))
            v = int(rng.integers(1, P31))
            for t in range(16):
                if base + t * cs < cols:
                    a[i, base + t * cs] = v
        a[i, rng.integers(0, cols, 3)] = rng.integers(1, P31, 3)  # a few corrections
    st = check(csr_from_dense(a), P31)
    assert st["col_stride"] == cs
    assert st["loads"] * 4 < int((a != 0).sum())


@pytest.mark.parametrize("rs", [1, 4, 32])
def test_row_block_structure_is_found_at_any_stride(rs):
    """the transposed pattern: blocks of 16 rows {hi*16*rs + t*rs + lo} that share their entries (P of a recursive algorithm)"""
    rng = np.random.default_rng(rs)
    rows, cols = 16 * rs * 2, 300
    a = np.zeros((rows, cols), dtype=np.uint64)
    for hi in range(2):
        for lo in range(rs):
            shared = rng.integers(0, cols, 20)
            vals = rng.integers(1, P31, 20)
            for t in range(16):
                a[hi * 16 * rs + t * rs + lo, shared] = vals
    a[rng.integers(0, rows, 10), rng.integers(0, cols, 10)] = rng.integers(1, P31, 10)
    st = check(csr_from_dense(a), P31, row_blocks=True)
    assert st["row_stride"] == rs
    assert st["loads"] * 4 < int((a != 0).sum())
    assert check(csr_from_dense(a), P31, row_blocks=False)["row_stride"] == 0


def test_the_32x32x32_triple():
    """BASELINE config 5: the encoder turns the 3.8 M entries of 32x32x32_15096 into under 0.5 M loads per sample, same products"""
    big = hm.load_large_csr(P31)
    assert big is not None, "tests/golden/large/32x32x32_15096.npz missing"
    _, _, (L, R, P) = big
    rng = np.random.default_rng(5)
    total = 0
    for A, rb in ((L, False), (R, False), (P, True)):
        rows, cols, ptr, col, val = A
        x = rng.integers(0, P31, cols).astype(np.uint32)
        y, st = capi.mmcheck_encode_check(P31, A, x, groups=128, row_blocks=rb)
        prod = val.astype(object) * x[col].astype(object)
        cs = np.concatenate([[0], np.cumsum(prod)])
        ref = np.array([int(cs[ptr[i + 1]] - cs[ptr[i]]) % P31 for i in range(rows)], dtype=np.uint64)
        assert (ref == y).all()
        total += st["loads"]
    assert total < 500000
