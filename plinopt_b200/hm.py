"""Host-side data plumbing for Hopcroft-Musinski (HM) matrices: the SMS text
format of the reference (data/README.md:10-17: optional '#' lines, header
`rows cols R|M`, 1-based `i j value` with value = integer or a/b, terminator
`0 0 0`), LCD scaling to the integer arrays the C ABI takes, CSR mod p."""
import json
import math
import os
from fractions import Fraction

import numpy as np


def read_sms(path_or_lines):
    """-> list of rows of Fraction."""
    if isinstance(path_or_lines, str):
        with open(path_or_lines) as f:
            lines = f.readlines()
    else:
        lines = list(path_or_lines)
    rows = cols = None
    M = None
    for line in lines:
        s = line.strip()
        if not s or s.startswith("#"):
            continue
        t = s.split()
        if rows is None:
            rows, cols = int(t[0]), int(t[1])
            M = [[Fraction(0)] * cols for _ in range(rows)]
            continue
        i, j = int(t[0]), int(t[1])
        if i == 0 and j == 0:
            break
        M[i - 1][j - 1] = Fraction(t[2])
    if M is None:
        raise ValueError("empty SMS stream")
    return M


def write_sms(M, path=None):
    """LinBox FileFormat(5) shape: `rows cols M`, entries row-major, `0 0 0` (src/orbiter.cpp:346-348)."""
    out = [f"{len(M)} {len(M[0])} M"]
    for i, row in enumerate(M):
        for j, v in enumerate(row):
            if v != 0:
                out.append(f"{i + 1} {j + 1} {v}")
    out.append("0 0 0")
    text = "\n".join(out) + "\n"
    if path:
        with open(path, "w") as f:
            f.write(text)
    return text


def load_fixture(stem, json_path=None):
    """(L, R, P) Fraction matrices of a triple stored in tests/golden/hm_matrices.json (data fixture)."""
    if json_path is None:
        json_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hm_matrices.json")
    with open(json_path) as f:
        d = json.load(f)
    out = []
    for x in "LRP":
        e = d[f"{stem}_{x}"]
        M = [[Fraction(0)] * e["cols"] for _ in range(e["rows"])]
        for i, j, v in e["entries"]:
            M[i][j] = Fraction(v)
        out.append(M)
    return tuple(out)


def lcd(M):
    l = 1
    for row in M:
        for v in row:
            l = l * v.denominator // math.gcd(l, v.denominator)
    return l


def scaled(M, dtype=np.int64):
    """(integer array of M * lcd, lcd)."""
    d = lcd(M)
    return np.array([[int(v * d) for v in row] for row in M], dtype=dtype), d


def LRP2MM(L, R, P):
    """include/plinopt_library.h:177-181."""
    n = int(math.sqrt(len(R[0]) * len(P) // len(L[0])))
    return len(P) // n, len(R[0]) // n, n


def csr_modp(M, p):
    """(rows, cols, ptr, col, val) with val = a * b^-1 mod p (src/MMchecker.cpp rebind to Modular)."""
    ptr = [0]; col = []; val = []
    for row in M:
        for j, v in enumerate(row):
            if v != 0:
                x = (v.numerator % p) * pow(v.denominator % p, -1, p) % p
                if x:
                    col.append(j); val.append(x)
        ptr.append(len(col))
    return (len(M), len(M[0]), np.array(ptr, dtype=np.int64), np.array(col, dtype=np.int32), np.array(val, dtype=np.uint32))


def strip_modulus(q):
    """src/MMchecker.cpp:123-126, src/orbiter.cpp:419-422: drop the factors of 2, 1 -> 2."""
    while q % 2 == 0 and q > 0:
        q >>= 1
    return 2 if q == 1 else q


def load_large_csr(p, path=None):
    """The regenerated 32x32x32_15096 triple (tools/regen_32x32x32.py) reduced mod p:
    returns (mkn, r, (L, R, P) CSR tuples) or None when the cache file is absent."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "large", "32x32x32_15096.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    out = []
    for x in "LRP":
        rows, cols = (int(v) for v in z[f"{x}_shape"])
        num = z[f"{x}_num"].astype(np.int64) % p
        den = z[f"{x}_den"].astype(np.int64) % p
        inv = {int(d): pow(int(d), -1, p) for d in np.unique(den)}
        val = (num * np.array([inv[int(d)] for d in den], dtype=np.int64)) % p if p < (1 << 31) else np.array([(int(a) * inv[int(d)]) % p for a, d in zip(num, den)], dtype=np.int64)
        out.append((rows, cols, z[f"{x}_ptr"].astype(np.int64), z[f"{x}_col"].astype(np.int32), val.astype(np.uint32)))
    return (32, 32, 32), 15096, tuple(out)
