"""Per-kernel times of the batched MMchecker on the 32x32x32_15096 triple (PLO_TIMING=1 prints them on stderr).
usage: PLO_TIMING=1 python tools/prof_mm.py [batch ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plinopt_b200 import capi, hm  # noqa: E402

P31 = 2147483647

if __name__ == "__main__":
    capi.set_device(0)
    batches = [int(x) for x in sys.argv[1:]] or [4096, 1024, 32]
    _, _, (Lc, Rc, Pc) = hm.load_large_csr(P31)
    nnz = sum(len(c[3]) for c in (Lc, Rc, Pc))
    for B in batches:
        plan = capi.MMcheckPlan(P31, (32, 32, 32), 15096, Lc, Rc, Pc, B)
        for _ in range(3):
            plan.run(1, 0)
        v, ok = plan.result()
        t0 = time.perf_counter()
        reps = 10
        for i in range(reps):
            plan.run(1 + i, 0)
        v, ok = plan.result()
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"batch": B, "ms_per_pass_wall": dt * 1e3, "samples_per_s": B / dt, "modmac_per_s": (nnz + 15096 + 32768) * B / dt,
                          "verdict": v, "ok": int(ok.sum())}))
        plan.close()
