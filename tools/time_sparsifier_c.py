"""plo_sparsifier timed at the C boundary (inputs already marshalled into int64 arrays: no Python conversion inside the timed region):
whole blockSparsifier pipeline, host buffers in, CoB/Res out.  One line per case: seconds per call, reference-loop candidate
evaluations covered per second, device round trips."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from plinopt_b200 import capi, hm  # noqa: E402


def time_case(M, q, c, reps=50, blocksize=4):
    num, den = capi._numden(M)
    r, cols = num.shape
    cn = np.zeros((cols, cols), dtype=np.int64); cd = np.ones((cols, cols), dtype=np.int64)
    rn = np.zeros((r, cols), dtype=np.int64); rd = np.ones((r, cols), dtype=np.int64)
    ok = C.c_int(0)
    stats = np.zeros(3, dtype=np.uint64)
    f = capi.lib().plo_sparsifier
    f.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.POINTER(C.c_int), C.c_void_p, C.c_int]
    args = (q, r, cols, capi._ptr(num), capi._ptr(den), blocksize, c, 1, capi._ptr(cn), capi._ptr(cd), capi._ptr(rn), capi._ptr(rd), C.byref(ok), capi._ptr(stats), -1)
    for _ in range(3):
        capi._check(f(*args))
    t = time.perf_counter()
    for _ in range(reps):
        f(*args)
    dt = (time.perf_counter() - t) / reps
    return dict(seconds=dt, candidates=int(stats[0]), round_trips=int(stats[1]), fallbacks=int(stats[2]), candidates_per_s=int(stats[0]) / dt,
                consistent=bool(ok.value), nnz_res=int(np.count_nonzero(rn)))


if __name__ == "__main__":
    capi.set_device(0)
    for stem, x, q, c in [("2x2x2_7_DPS-smallrat-12.2034", 0, 0, 4), ("4x4x4_48_rational", 0, 2147483647, 11), ("4x4x4_48_rational", 0, 0, 11),
                          ("3x4x7_63_rational", 1, 0, 11), ("4x4x4_48_rational", 0, 2147483647, 40), ("4x4x4_48_rational", 0, 2147483647, 128),
                          ("4x4x4_48_rational", 0, 0, 40)]:
        M = hm.load_fixture(stem)[x]
        print(json.dumps(dict(case=f"{stem}_{'LRP'[x]} -q {q} -c {c}", **time_case(M, q, c, reps=50 if c < 100 else 5))))
