"""ctypes view of the CPU oracle (oracle/libplo_oracle.so) and of the committed
HM-matrix fixtures (tests/golden/hm_matrices.json).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py only.  The product package
(plinopt_b200/) never imports this module.
"""
import ctypes as C
import json
import math
import os
import subprocess
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libplo_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "plo_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return ORACLE_SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_growth_G2.restype = C.c_double
        _lib.orc_lincomb_bench.restype = C.c_int64
    return _lib


# --------------------------------------------------------------------------
# fixtures
# --------------------------------------------------------------------------
_mats = None


def matrices():
    global _mats
    if _mats is None:
        with open(os.path.join(GOLDEN, "hm_matrices.json")) as f:
            _mats = json.load(f)
    return _mats


def dense_fractions(name):
    """rows x cols list-of-lists of Fraction for fixture `name`."""
    d = matrices()[name]
    M = [[Fraction(0)] * d["cols"] for _ in range(d["rows"])]
    for i, j, v in d["entries"]:
        M[i][j] = Fraction(v)
    return M


def numden(M):
    """list-of-lists of Fraction -> (num, den) int64 arrays (row-major)."""
    r, c = len(M), len(M[0])
    num = np.zeros((r, c), dtype=np.int64)
    den = np.ones((r, c), dtype=np.int64)
    for i in range(r):
        for j in range(c):
            num[i, j] = M[i][j].numerator
            den[i, j] = M[i][j].denominator
    return num, den


def triple(stem):
    return tuple(dense_fractions(f"{stem}_{x}") for x in "LRP")


def lcm_den(M):
    l = 1
    for row in M:
        for v in row:
            l = l * v.denominator // math.gcd(l, v.denominator)
    return l


def scaled_int(M):
    """(int64 array of M*den, den) with den the LCD of the entries."""
    d = lcm_den(M)
    A = np.array([[int(v * d) for v in row] for row in M], dtype=np.int64)
    return A, d


def LRP2MM(L, R, P):
    m, k, n = C.c_int(), C.c_int(), C.c_int()
    lib().orc_LRP2MM(len(L[0]), len(R[0]), len(P), C.byref(m), C.byref(k), C.byref(n))
    return m.value, k.value, n.value


# --------------------------------------------------------------------------
# oracle entry points
# --------------------------------------------------------------------------
def growth_G2(L, R, P):
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    f = lib().orc_growth_G2
    f.argtypes = [C.c_int] * 4 + [_i64p] * 6
    return f(len(L), len(L[0]), len(R[0]), len(P), Ln, Ld, Rn, Rd, Pn, Pd)


def sparsifier(M, p=0, blocksize=4, maxnumcoeff=11, initial_elimination=True, trace=False):
    """Returns (CoB, Res, consistent, trace) with CoB/Res as Fraction (p=0) or int (mod p) lists."""
    num, den = numden(M)
    r, c = num.shape
    cn = np.zeros((c, c), dtype=np.int64); cd = np.ones((c, c), dtype=np.int64)
    rn = np.zeros((r, c), dtype=np.int64); rd = np.ones((r, c), dtype=np.int64)
    ok = C.c_int(0)
    L = lib()
    L.orc_trace_enable(1 if trace else 0)
    f = L.orc_sparsifier
    f.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p, C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p, C.POINTER(C.c_int)]
    rc = f(p, r, c, num, den, blocksize, maxnumcoeff, 1 if initial_elimination else 0, cn, cd, rn, rd, C.byref(ok))
    if rc != 0:
        raise RuntimeError(f"oracle sparsifier error {rc}")
    tr = []
    if trace:
        for i in range(L.orc_trace_size()):
            b, nu, rl, cl, fb, cc = (C.c_int32() for _ in range(6))
            idx = C.c_int64()
            L.orc_trace_get(i, C.byref(b), C.byref(nu), C.byref(rl), C.byref(cl), C.byref(idx), C.byref(fb), C.byref(cc))
            tr.append(dict(block=b.value, num=nu.value, rl=rl.value, cl=cl.value, index=idx.value, fallback=fb.value, c=cc.value))
        L.orc_trace_enable(0)
    if p == 0:
        CoB = [[Fraction(int(cn[i, j]), int(cd[i, j])) for j in range(c)] for i in range(c)]
        Res = [[Fraction(int(rn[i, j]), int(rd[i, j])) for j in range(c)] for i in range(r)]
    else:
        CoB = cn.tolist(); Res = rn.tolist()
    return CoB, Res, bool(ok.value), tr


def coeffs(TM, p, maxnumcoeff):
    """Coefficient list (plinopt_sparsify.inl:256-268) as (num, den) arrays."""
    num, den = numden(TM) if isinstance(TM[0][0], Fraction) else (np.asarray(TM, dtype=np.int64), None)
    if den is None:
        den = np.ones_like(num)
    n, m = num.shape
    on = np.zeros(maxnumcoeff + 8, dtype=np.int64); od = np.ones(maxnumcoeff + 8, dtype=np.int64)
    f = lib().orc_coeffs
    f.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p, C.c_int, _i64p, _i64p]
    k = f(p, n, m, num, den, maxnumcoeff, on, od)
    if k < 0:
        raise RuntimeError(f"oracle coeffs error {k}")
    return on[:k].copy(), od[:k].copy()


def lincomb_search(p, tm_num, tm_den, off, num, cf_num, cf_den, lcob_num, lcob_den, init_rl=-1, init_cl=-1):
    """One literal quad-loop step; returns (rl, cl, index)."""
    n, m = tm_num.shape
    rl, cl, idx = C.c_int(), C.c_int(), C.c_int64()
    f = lib().orc_lincomb_search
    f.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p, C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p,
                  C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    rc = f(p, n, m, np.ascontiguousarray(tm_num), np.ascontiguousarray(tm_den), off, num, len(cf_num),
           np.ascontiguousarray(cf_num), np.ascontiguousarray(cf_den), np.ascontiguousarray(lcob_num),
           np.ascontiguousarray(lcob_den), init_rl, init_cl, C.byref(rl), C.byref(cl), C.byref(idx))
    if rc != 0:
        raise RuntimeError(f"oracle lincomb error {rc}")
    return rl.value, cl.value, idx.value


def lincomb_bench(p, tm_num, tm_den, off, num, cf_num, cf_den, lcob_num, lcob_den, i_lo, i_hi, nthreads=0):
    n, m = tm_num.shape
    rl, cl, idx = C.c_int(), C.c_int(), C.c_int64()
    f = lib().orc_lincomb_bench
    f.argtypes = [C.c_int64, C.c_int, C.c_int, _i64p, _i64p, C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i64p, _i64p,
                  C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    tot = f(p, n, m, np.ascontiguousarray(tm_num), np.ascontiguousarray(tm_den), off, num, len(cf_num),
            np.ascontiguousarray(cf_num), np.ascontiguousarray(cf_den), np.ascontiguousarray(lcob_num),
            np.ascontiguousarray(lcob_den), i_lo, i_hi, nthreads, C.byref(rl), C.byref(cl), C.byref(idx))
    return tot, rl.value, cl.value, idx.value


def orbit_decode(m, k, n, mode, seed, index):
    U = np.zeros((m, m), dtype=np.int32); V = np.zeros((k, k), dtype=np.int32); W = np.zeros((n, n), dtype=np.int32)
    f = lib().orc_orbit_decode
    f.argtypes = [C.c_int] * 4 + [C.c_uint64, C.c_uint64, _i32p, _i32p, _i32p]
    f(m, k, n, mode, seed, index, U, V, W)
    return U, V, W


def orbit_sweep(L, R, P, measure, mode, seed, lo, hi, p=0, nthreads=0, table=True):
    """Scores candidates lo..hi-1.  Returns dict(best=(index,nnz,nno,g2), nnz=, nno=, g2=)."""
    m, k, n = LRP2MM(L, R, P)
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    cnt = hi - lo
    nnz = np.zeros(cnt if table else 1, dtype=np.uint32)
    nno = np.zeros(cnt if table else 1, dtype=np.uint32)
    g2 = np.zeros(cnt if table else 1, dtype=np.float64)
    bi, bz, bo, bg = C.c_uint64(), C.c_uint32(), C.c_uint32(), C.c_double()
    f = lib().orc_orbit_sweep
    f.argtypes = [C.c_int64] + [C.c_int] * 4 + [_i64p] * 6 + [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
    rc = f(p, m, k, n, len(L), Ln, Ld, Rn, Rd, Pn, Pd, measure, mode, seed, lo, hi, nthreads,
           nnz.ctypes.data if table else None, nno.ctypes.data if table else None, g2.ctypes.data if table else None,
           C.byref(bi), C.byref(bz), C.byref(bo), C.byref(bg))
    if rc != 0:
        raise RuntimeError(f"oracle orbit error {rc}")
    return dict(best=(bi.value, bz.value, bo.value, bg.value), nnz=nnz, nno=nno, g2=g2)


def orbit_apply(L, R, P, U, V, W):
    m, k, n = LRP2MM(L, R, P)
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    outs = [np.zeros_like(a) for a in (Ln, Ld, Rn, Rd, Pn, Pd)]
    f = lib().orc_orbit_apply
    f.argtypes = [C.c_int] * 4 + [_i64p] * 6 + [_i32p] * 3 + [_i64p] * 6
    rc = f(m, k, n, len(L), Ln, Ld, Rn, Rd, Pn, Pd, np.ascontiguousarray(U, dtype=np.int32),
           np.ascontiguousarray(V, dtype=np.int32), np.ascontiguousarray(W, dtype=np.int32), *outs)
    if rc != 0:
        raise RuntimeError(f"oracle orbit_apply error {rc}")

    def back(nu, de):
        return [[Fraction(int(nu[i, j]), int(de[i, j])) for j in range(nu.shape[1])] for i in range(nu.shape[0])]
    return back(outs[0], outs[1]), back(outs[2], outs[3]), back(outs[4], outs[5])


def csr_modp(M, p):
    """CSR (ptr int64, col int32, val int64 residues) of a Fraction matrix reduced mod p."""
    ptr = [0]; col = []; val = []
    for row in M:
        for j, v in enumerate(row):
            if v != 0:
                x = (v.numerator % p) * pow(v.denominator % p, -1, p) % p
                if x != 0:
                    col.append(j); val.append(x)
        ptr.append(len(col))
    return np.array(ptr, dtype=np.int64), np.array(col, dtype=np.int32), np.array(val, dtype=np.int64)


def mmcheck_modp(p, L, R, P, ua, ub):
    Lp, Lc, Lv = csr_modp(L, p); Rp, Rc, Rv = csr_modp(R, p); Pp, Pc, Pv = csr_modp(P, p)
    f = lib().orc_mmcheck_modp
    f.argtypes = [C.c_int64] + [C.c_int] * 4 + [_i64p, _i32p, _i64p] * 3 + [_i64p, _i64p]
    return f(p, len(L), len(L[0]), len(R[0]), len(P), Lp, Lc, Lv, Rp, Rc, Rv, Pp, Pc, Pv,
             np.ascontiguousarray(ua, dtype=np.int64), np.ascontiguousarray(ub, dtype=np.int64))


def mmcheck_q(L, R, P, ua, ub):
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    f = lib().orc_mmcheck_q
    f.argtypes = [C.c_int] * 4 + [_i64p] * 6 + [_i64p, _i64p]
    return f(len(L), len(L[0]), len(R[0]), len(P), Ln, Ld, Rn, Rd, Pn, Pd,
             np.ascontiguousarray(ua, dtype=np.int64), np.ascontiguousarray(ub, dtype=np.int64))


def factor_sweep(M, k, seed, lo, hi, p=0, nthreads=0, table=True, matrices=False):
    """Factorizer random restarts (plinopt_sparsify.inl:924-990) over candidates lo..hi-1; M as Fractions.
    Returns dict(best=(nnz_alt, nno_alt, nnz_cob, index), table=(hi-lo) x 3, alt=, cob=)."""
    L = lib()
    r, n = len(M), len(M[0])
    num, den = numden(M)
    cnt = hi - lo
    tab = np.zeros((cnt, 3), dtype=np.uint32) if table else None
    bidx = C.c_uint64(); bops = (C.c_uint32 * 3)()
    an = np.zeros((r, k), dtype=np.int64); ad = np.ones((r, k), dtype=np.int64)
    cn = np.zeros((k, n), dtype=np.int64); cd = np.ones((k, n), dtype=np.int64)
    f = L.orc_factor_sweep
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = f(p, r, n, k, num.ctypes.data, den.ctypes.data, seed, lo, hi, nthreads, tab.ctypes.data if table else None, C.byref(bidx), bops,
           an.ctypes.data if matrices else None, ad.ctypes.data if matrices else None, cn.ctypes.data if matrices else None,
           cd.ctypes.data if matrices else None)
    assert rc == 0, f"oracle overflow/invalid ({rc})"
    idx = None if bidx.value == 0xFFFFFFFFFFFFFFFF else bidx.value
    out = {"best": (bops[0], bops[1], bops[2], idx), "table": tab}
    if matrices:
        out["alt"] = [[Fraction(int(a), int(b)) for a, b in zip(ra, rb)] for ra, rb in zip(an, ad)]
        out["cob"] = [[Fraction(int(a), int(b)) for a, b in zip(ra, rb)] for ra, rb in zip(cn, cd)]
    return out


def factor_decode(r, seed, index):
    perm = np.zeros(r, dtype=np.int32)
    f = lib().orc_factor_decode
    f.restype = None
    f.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
    f(r, seed, index, perm.ctypes.data)
    return perm


def depender(M, level, maxnumcoeff=11, p=0, user=(), max_hits=1 << 20):
    """dependency (src/dependency.cpp:106-169): returns dict(hits=[(depth, pos, rows, coefs)], nhits, ncand, coeffs)."""
    r, n = len(M), len(M[0])
    num, den = numden(M)
    un = np.array([Fraction(u).numerator for u in user], dtype=np.int64); ud = np.array([Fraction(u).denominator for u in user], dtype=np.int64)
    hits = np.zeros((max_hits, 12), dtype=np.int32)
    nh, nc, ncoef = C.c_uint64(), C.c_uint64(), C.c_int()
    cn = np.zeros(maxnumcoeff, dtype=np.int64); cd = np.ones(maxnumcoeff, dtype=np.int64)
    f = lib().orc_depender
    f.restype = C.c_int
    f.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    rc = f(p, r, n, num.ctypes.data, den.ctypes.data, len(user), un.ctypes.data if len(user) else None, ud.ctypes.data if len(user) else None,
           maxnumcoeff, level, max_hits, hits.ctypes.data, C.byref(nh), C.byref(nc), cn.ctypes.data, cd.ctypes.data, C.byref(ncoef))
    assert rc == 0, f"oracle overflow ({rc})"
    k = min(nh.value, max_hits)
    out = [(int(h[0]), int(h[1]), tuple(int(x) for x in h[2:7]), tuple(int(x) for x in h[7:12])) for h in hits[:k]]
    coeffs = [Fraction(int(a), int(b)) for a, b in zip(cn[:ncoef.value], cd[:ncoef.value])]
    return dict(hits=out, nhits=nh.value, ncand=nc.value, coeffs=coeffs)


def _triple_call(fname, head_args, L, R, P, out_shapes, stats_words=0):
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    outs = []
    for shp in out_shapes:
        outs += [np.zeros(shp, dtype=np.int64), np.ones(shp, dtype=np.int64)]
    st = np.zeros(max(stats_words, 1), dtype=np.uint64)
    f = getattr(lib(), fname)
    f.restype = C.c_int
    f.argtypes = [C.c_int] * len(head_args) + [C.c_void_p] * (12 + (1 if stats_words else 0))
    args = list(head_args) + [a.ctypes.data for a in (Ln, Ld, Rn, Rd, Pn, Pd)] + [o.ctypes.data for o in outs] + ([st.ctypes.data] if stats_words else [])
    rc = f(*args)
    assert rc == 0, f"oracle error {rc}"
    mats = [[[Fraction(int(a), int(b)) for a, b in zip(ra, rb)] for ra, rb in zip(outs[2 * t], outs[2 * t + 1])] for t in range(3)]
    return mats, [int(v) for v in st]


def negater(L, R, P, only_sign=False):
    """src/negater.cpp:117-209: returns ((L', R', P'), stats[12])."""
    return _triple_call("orc_negater", [1 if only_sign else 0, len(L), len(L[0]), len(R[0]), len(P)], L, R, P,
                        [(len(L), len(L[0])), (len(R), len(R[0])), (len(P), len(P[0]))], 12)


def rotater(L, R, P, right=False):
    """bin/rotater.sh:75-83: the rotated triple."""
    m, k, n = LRP2MM(L, R, P)
    r = len(L)
    shapes = [(r, m * n), (r, m * k), (n * k, r)] if right else [(r, k * n), (r, m * n), (m * k, r)]
    return _triple_call("orc_rotater", [1 if right else 0, m, k, n, r], L, R, P, shapes)[0]


def growth_factors(L, R, P):
    """src/growthfactor.cpp:199-229: the eleven printed factors, in print order."""
    Ln, Ld = numden(L); Rn, Rd = numden(R); Pn, Pd = numden(P)
    out = np.zeros(11, dtype=np.float64)
    f = lib().orc_growth_factors
    f.restype = C.c_int
    f.argtypes = [C.c_int] * 4 + [C.c_void_p] * 7
    rc = f(len(L), len(L[0]), len(R[0]), len(P), *[a.ctypes.data for a in (Ln, Ld, Rn, Rd, Pn, Pd)], out.ctypes.data)
    assert rc == 0
    return out.tolist()
