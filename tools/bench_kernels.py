#!/usr/bin/env python3
"""Side benchmarks of the other two kernels of the hot path (not the bench.py headline):
  * sparsifier candidate search, BASELINE config 3 shape: 4x4x4_48_rational_L mod 2^31-1, 4 column
    blocks batched, c in {20, 64, 128}
  * batched MMchecker mod 2^31-1 on a synthetic CSR triple with the size/density of 32x32x32_15096
Prints one JSON line per case (CUDA-event timing on the launch stream, 3 warm-up runs)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from plinopt_b200 import capi, hm  # noqa: E402

P31 = 2147483647
BATCHES = (32, 256, 1024, 4096)


def coeff_list(tm, p, c):
    """Reference enumeration order (plinopt_sparsify.inl:256-268) restated for residues mod p."""
    C = [0, 1, -1]

    def aug(r):
        if r in C:
            return
        inv = pow(r % p, -1, p)
        C.extend([r, -r, inv, (p - inv) % p])
    for row in tm:
        for v in row:
            if v % p:
                aug(int(v))
    i = 2
    while len(C) < c:
        aug(i); i += 1
    return np.array(C[:c], dtype=np.int64)


def time_plan(run, reps):
    st = torch.cuda.current_stream()
    for _ in range(3):
        run(st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        run(st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lincomb_cases(peaks):
    L, _, _ = hm.load_fixture("4x4x4_48_rational")
    for c in (20, 64, 128):
        tms, cfs = [], []
        for blk in range(4):
            TM = [[L[i][4 * blk + t] for i in range(len(L))] for t in range(4)]
            tm = np.array([[(v.numerator % P31) * pow(v.denominator % P31, -1, P31) % P31 for v in row] for row in TM], dtype=np.int64)
            tms.append(tm); cfs.append(coeff_list(tm.tolist(), P31, c))
        plan = capi.LincombPlan(P31, np.stack(tms), 0, np.stack(cfs))
        ms = time_plan(lambda s: plan.run(s), 5 if c >= 64 else 50)
        rl, cl, idx = plan.result()
        cand = plan.candidates
        m = 48
        print(json.dumps({"kernel": "lincomb_kernel<u32,48,modp>", "case": f"4x4x4_48_rational_L, 4 blocks, c={c}", "candidates": cand,
                          "ms": ms, "candidates_per_s": cand / ms * 1e3, "compare_add_pairs_per_s": cand * m / ms * 1e3,
                          "ialu_pair_peak": peaks["ialu_pairs_per_s"], "frac_of_ialu_pair_peak": cand * m / ms * 1e3 / peaks["ialu_pairs_per_s"],
                          "best": [int(rl[0]), int(cl[0]), int(idx[0])]}))
        plan.close()


def mmcheck_case():
    rng = np.random.default_rng(0)
    m = k = n = 32
    r = 15096

    def rand_csr(rows, cols, nnz_row):
        ptr = np.arange(rows + 1, dtype=np.int64) * nnz_row
        col = np.concatenate([np.sort(rng.choice(cols, nnz_row, replace=False)) for _ in range(rows)]).astype(np.int32)
        val = rng.integers(1, P31, rows * nnz_row).astype(np.uint32)
        return (rows, cols, ptr, col, val)
    big = hm.load_large_csr(P31)
    if big is not None:  # the real triple regenerated from the reference's .slp (tools/regen_32x32x32.py)
        _, _, (Lc, Rc, Pc) = big
        name = "32x32x32_15096_{L,R,P} mod 2^31-1"
    else:
        Lc, Rc, Pc = rand_csr(r, 1024, 83), rand_csr(r, 1024, 84), rand_csr(1024, r, 1230)
        name = "synthetic 32x32x32_15096-like CSR"
    nnz = sum(len(c[3]) for c in (Lc, Rc, Pc))
    for B in BATCHES:
        plan = capi.MMcheckPlan(P31, (m, k, n), r, Lc, Rc, Pc, B)
        ms = time_plan(lambda s: plan.run(1, 0, s), 10)
        v, ok = plan.result()
        bytes_csr = nnz * 8
        print(json.dumps({"kernel": "mm_spmm_kernel (x3) + hadamard + verify", "case": f"{name}, batch={B}",
                          "ms": ms, "samples_per_s": B / ms * 1e3, "modmac_per_s": (nnz + r + m * k * n) * B / ms * 1e3,
                          "csr_stream_GBs_if_read_once": bytes_csr / ms * 1e-6, "verdict": v, "samples_ok": int(ok.sum())}))
        plan.close()


if __name__ == "__main__":
    capi.set_device(0)
    peaks = capi.measure_peaks(5)
    print(json.dumps({"peaks": peaks}))
    if "--mm-only" not in sys.argv:
        lincomb_cases(peaks)
    mmcheck_case()
