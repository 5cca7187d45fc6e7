"""GPU parity: dependency `Explore` (src/dependency.cpp:73-169) against the oracle's literal
recursion: same coefficient list, same hits, same (depth-first) order, same count of combinations."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import hm

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bin")

CASES = [("2x2x2_7_Winograd_L", 4, 5, 0), ("2x2x2_7_Winograd_P", 4, 7, 0), ("2x2x2_7_DPS-smallrat-12.2034_L", 4, 11, 0),
         ("3x3x3_23_58_L", 3, 5, 0), ("3x3x3_23_58_R", 4, 3, 0), ("4x4x4_48_rational_L", 3, 5, 0), ("3x4x7_63_rational_R", 3, 4, 0),
         ("3x4x7_63_rational_L", 3, 7, 101), ("2x2x2_7_Winograd_R", 5, 4, 7), ("3x3x6_40_P", 3, 5, 513083), ("cyclic", 3, 5, 0)]


@pytest.mark.parametrize("name,level,c,q", CASES)
def test_hits_match_oracle(capi, name, level, c, q):
    M = O.dense_fractions(name)
    got = capi.depender(M, level, c, q=q)
    ref = O.depender(M, level, c, p=q)
    assert got["coeffs"] == ref["coeffs"]
    assert got["ncand"] == ref["ncand"] and got["nhits"] == ref["nhits"]
    assert got["hits"] == ref["hits"]
    assert got["text"].count("\n") == ref["nhits"]


def test_user_coefficients_and_text(capi):
    """-v "3 1/3" (src/dependency.cpp:241-245); output format of showOut/showLC (:47-71)."""
    M = O.dense_fractions("2x2x2_7_Winograd_L")
    got = capi.depender(M, 3, 6, user=("3", "1/3"))
    ref = O.depender(M, 3, 6, user=("3", "1/3"))
    assert got["coeffs"] == ref["coeffs"] and got["coeffs"][:4] == [1, -1, 3, O.Fraction(1, 3)]
    assert got["hits"] == ref["hits"]
    lines = got["text"].splitlines()
    assert all(l.endswith(";") and l[0] in "+-" for l in lines)
    # every line is an identity: re-evaluate it exactly
    for (depth, pos, rows, coefs), line in zip(got["hits"], lines):
        w = [M[rows[0]][j] + sum(got["coeffs"][coefs[t]] * M[rows[t]][j] for t in range(1, depth + 1)) for j in range(len(M[0]))]
        assert sum(1 for v in w if v != 0) == (0 if pos < 0 else 1)
        if pos >= 0:
            assert w[pos] != 0 and line.startswith(("-" if w[pos] > 0 else "+") + f"i{pos}")
        assert f"+o{rows[0]}" in line


def test_hit_list_truncation_and_errors(capi):
    M = O.dense_fractions("2x2x2_7_Winograd_L")
    full = capi.depender(M, 4, 5)
    # too little room: the C call reports PLO_E_RANGE with the total (a truncated list would be an arbitrary subset, not the first lines of the
    # reference's depth-first output); the binding comes back with room and returns the complete list
    import ctypes as C
    num, den = capi._numden(M)
    hits = (capi.DepHit * 7)()
    nh, nc, tl, ncoef = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
    cn = np.zeros(5, dtype=np.int64); cd = np.ones(5, dtype=np.int64)
    f = capi.lib().plo_depender
    rc = f(0, len(M), len(M[0]), capi._ptr(num), capi._ptr(den), 0, None, None, 5, 4, 7, C.cast(hits, C.c_void_p), C.byref(nh), C.byref(nc), None, 0, C.byref(tl),
           capi._ptr(cn), capi._ptr(cd), C.byref(ncoef))
    assert rc == capi.E_RANGE and nh.value == full["nhits"] > 7
    part = capi.depender(M, 4, 5, max_hits=7)
    assert part["nhits"] == full["nhits"] and part["hits"] == full["hits"] and part["text"] == full["text"]
    assert capi.depender(M, 1, 5)["ncand"] == 0  # level 1: nothing to add (:153-160)
    with pytest.raises(capi.PloError):
        capi.depender(M, 7, 5)  # more than 4 added rows is not supported
    with pytest.raises(capi.PloError):
        capi.depender(M, 3, 5, q=15)  # composite modulus


def test_dependency_cli(capi, tmp_path):
    M = O.dense_fractions("3x3x3_23_58_L")
    f = tmp_path / "m.sms"
    hm.write_sms(M, str(f))
    p = subprocess.run([os.path.join(BIN, "dependency"), "-l", "3", "-c", "5", str(f)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    ref = O.depender(M, 3, 5)
    assert "# [DEPND] level 3, coefficients: [1,-1,2,-2,1/2]" in p.stderr
    assert len(p.stdout.splitlines()) == ref["nhits"] and p.stdout == capi.depender(M, 3, 5)["text"]
