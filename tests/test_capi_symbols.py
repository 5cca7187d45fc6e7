"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/plinopt_b200.h
declares, its host-only entry points agree with the oracle, and compute calls FAIL LOUDLY without a
CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import capi, hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "plinopt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plo_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/plinopt_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == syms
    assert L.plo_version() >= 100


def test_host_decode_matches_oracle_decode():
    for mkn in [(2, 2, 2), (3, 3, 3), (4, 4, 4), (3, 4, 7), (6, 3, 3), (8, 2, 5)]:
        for mode in (0, 1):
            if mode == 0 and capi.orbit_space(*mkn) == 0:
                continue
            for idx in list(range(40)) + [2 ** 40 + 12345, 2 ** 63 + 5]:
                a = capi.orbit_decode(*mkn, mode, 0x504C494E4F505431, idx)
                b = O.orbit_decode(*mkn, mode, 0x504C494E4F505431, idx)
                for x, y in zip(a, b):
                    assert np.array_equal(x, y)


def test_orbit_space():
    assert capi.orbit_space(2, 2, 2) == 48 ** 3
    assert capi.orbit_space(3, 3, 3) == (36 * 8 * 27) ** 3
    assert capi.orbit_space(3, 4, 7) == 0  # does not fit in 64 bits


def test_sms_roundtrip_and_helpers(tmp_path):
    L, R, P = hm.load_fixture("2x2x2_7_DPS-smallrat-12.2034")
    p = tmp_path / "x.sms"
    hm.write_sms(L, str(p))
    assert hm.read_sms(str(p)) == L
    assert hm.read_sms(["# comment", "2 2 R", "2 1 -3/4", "1 2 5", "0 0 0"]) == [[0, 5], [O.Fraction(-3, 4), 0]]
    assert hm.LRP2MM(L, R, P) == (2, 2, 2) == O.LRP2MM(L, R, P)
    A, d = hm.scaled(L)
    assert d == 9 and A.dtype == np.int64
    assert hm.strip_modulus(1026166) == 513083 and hm.strip_modulus(8) == 2  # MMchecker.cpp:123-126


@pytest.mark.skipif(capi.device_count() > 0, reason="checks the no-device behaviour")
def test_compute_calls_fail_loudly_without_a_device():
    L, R, P = hm.load_fixture("2x2x2_7_Winograd")
    Li, Ri, Pi = (hm.scaled(M, np.int32)[0] for M in (L, R, P))
    with pytest.raises(capi.PloError) as e:
        capi.orbit_sweep((2, 2, 2), Li, Ri, Pi, (1, 1, 1), 0, 1, 0, 0, 10)
    assert e.value.code == capi.E_NODEVICE
    with pytest.raises(capi.PloError) as e:
        capi.lincomb_search(0, np.ones((4, 7), np.int64), 0, np.array([0, 1, -1], np.int64))
    assert e.value.code == capi.E_NODEVICE
    with pytest.raises(capi.PloError) as e:
        capi.measure_peaks(1)
    assert e.value.code == capi.E_NODEVICE
    # the MMchecker drivers (one prime, and over Q through several primes) and the plans: no CPU path either.  The encoder self-test
    # (plo_mmcheck_encode_check, tests/test_mm_encoder.py) is host-only by design and is not a way around this.
    for call in (lambda: capi.mmchecker(L, R, P), lambda: capi.mmchecker_bits(L, R, P, bitsize=8),
                 lambda: capi.mmcheck_batch(7, (2, 2, 2), 7, hm.csr_modp(L, 7), hm.csr_modp(R, 7), hm.csr_modp(P, 7), batch=4)):
        with pytest.raises(capi.PloError) as e:
            call()
        assert e.value.code == capi.E_NODEVICE


def test_argument_errors():
    with pytest.raises(capi.PloError) as e:
        capi.orbit_decode(9, 2, 2, 1, 0, 0)
    assert e.value.code == capi.E_ARG


def test_whole_matrix_decode_selftest():
    """Host-only: the table-driven kernels index a 2x2 (48) or 3x3 (7776) zoi matrix by ONE number per factor; the library checks that
    this numbering reproduces the digit-by-digit decode in both enumeration modes (no device needed)."""
    import ctypes
    from plinopt_b200 import capi
    f = capi.lib().plo_selftest_matrix_index
    f.restype = ctypes.c_int
    assert f() == 0
