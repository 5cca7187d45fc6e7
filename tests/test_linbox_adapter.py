"""include/plinopt_b200_linbox.hpp -- the reference-side adapter -- compiled and run against a 50-line mock of
LinBox::SparseMatrix<Field, SparseSeq> (tests/linbox_mock.hpp): LinBox/Givaro are not installed here, so this is how the shims of
INTEGRATION.md are kept compiling.  CPU: the exact conversions; GPU: the three shims end to end against the ctypes binding."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from plinopt_b200 import capi, hm

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
P31 = 2147483647


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("adapter") / "adapter_check")
    cmd = ["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "adapter_check.cpp"), "-o", out,
           "-L" + os.path.join(ROOT, "plinopt_b200"), "-lplinopt_b200", "-Wl,-rpath," + os.path.join(ROOT, "plinopt_b200")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return out


def text_triple(stem):
    out = []
    for M, x in zip(hm.load_fixture(stem), "LRP"):
        ent = [(i, j, v) for i, row in enumerate(M) for j, v in enumerate(row) if v != 0]
        out.append(f"{x} {len(M)} {len(M[0])} {len(ent)}")
        out += [f"{i} {j} {v.numerator} {v.denominator}" for i, j, v in ent]
    return "\n".join(out) + "\n"


def run(exe, stem, *args):
    p = subprocess.run([exe, *args], input=text_triple(stem), capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    return json.loads(p.stdout)


def col_scaled(TM):
    A = np.zeros((len(TM), len(TM[0])), dtype=np.int64)
    for j in range(len(TM[0])):
        l = 1
        for i in range(len(TM)):
            l = l * TM[i][j].denominator // math.gcd(l, TM[i][j].denominator)
        for i in range(len(TM)):
            A[i, j] = int(TM[i][j] * l)
    return A


def test_adapter_conversions(exe):
    stem = "4x4x4_48_rational"
    d = run(exe, stem)
    L, R, P = hm.load_fixture(stem)
    TM = [[L[i][t] for i in range(len(L))] for t in range(4)]
    assert d["tm"] == col_scaled(TM).ravel().tolist()
    assert d["coeffs"] == [0, 2, -2, 1, -1, 4, -4]
    Li, dl = hm.scaled(L, np.int32)
    assert d["L32"] == Li.ravel().tolist() and d["denL"] == dl
    rows, cols, ptr, col, val = hm.csr_modp(P, P31)
    assert d["csr_ptr"] == ptr.tolist() and d["csr_col"] == col.tolist() and d["csr_val"] == val.tolist()


@pytest.mark.gpu
def test_adapter_shims_on_the_device(exe):
    stem = "4x4x4_48_rational"
    d = run(exe, stem, "device")
    L, R, P = hm.load_fixture(stem)
    TM = col_scaled([[L[i][t] for i in range(len(L))] for t in range(4)])
    (status, rows), = capi.lincomb_quad(0, [dict(TM=TM, off=0, coeffs=np.array(d["coeffs"], dtype=np.int64))])
    assert d["quad_status"] == status and d["quad_nrows"] == len(rows)
    assert [(d["quad_rl"][t], d["quad_cl"][t], d["quad_index"][t]) for t in range(len(rows))] == rows
    mkn = hm.LRP2MM(L, R, P)
    (Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
    best = capi.orbit_sweep(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_NNZ, capi.MODE_PHILOX, 0x504C494E4F505431, 0, 4096)
    assert d["orbit"] == [best["nnz"], best["nno"], best["index"]]
    assert d["U"] == capi.orbit_decode(*mkn, capi.MODE_PHILOX, 0x504C494E4F505431, best["index"])[0].ravel().tolist()
    assert d["mmcheck"] == 0
