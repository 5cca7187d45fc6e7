#!/usr/bin/env python3
"""Multi-GPU runs of the other BASELINE configs (bench.py is configs[1]); one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_multi.py
  * configs[2]: sparsifier candidate search on 4x4x4_48_rational_L mod 2^31-1, 4 column blocks, c = 128: the prefix range of the
    search is sharded over the ranks (strong scaling), one all_reduce(MAX) merges the winners;
  * configs[3]: orbit + sparsity search on 3x4x7_63_rational, 2^21 candidates per GPU (weak scaling), one all_reduce(MIN);
  * configs[4]: batched MMchecker mod 2^31-1 of 32x32x32_15096, 4096 samples per GPU (weak scaling), verdicts AND-ed.
CUDA-event timing on the launch stream, max over ranks; rank 0 prints one JSON line per config."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from plinopt_b200 import capi, hm, sharding  # noqa: E402

P31 = 2147483647
SEED = 0x504C494E4F505431


def timed(run, reps, stream, world, dev):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    capi.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_kernels import coeff_list

    # ---- configs[2] ----
    L, R, P = hm.load_fixture("4x4x4_48_rational")
    c = 128
    tms, cfs = [], []
    for blk in range(4):
        TM = [[L[i][4 * blk + t] for i in range(len(L))] for t in range(4)]
        tm = np.array([[(v.numerator % P31) * pow(v.denominator % P31, -1, P31) % P31 for v in row] for row in TM], dtype=np.int64)
        tms.append(tm); cfs.append(coeff_list(tm.tolist(), P31, c))
    plan = capi.LincombPlan(P31, np.stack(tms), 0, np.stack(cfs))
    lo, hi = sharding.shard_range(0, c ** 3, rank, world)
    ms = timed(lambda: plan.run_range(lo, hi, sp), 5, stream, world, dev)
    mine = [(int(a), int(b), None if int(i) == capi.NO_INDEX else int(i)) for a, b, i in zip(*plan.result(sp))]
    best = sharding.allreduce_lincomb(mine, device=dev)
    if rank == 0:
        print(json.dumps({"config": "sparsifier CoB search, 4x4x4_48_rational_L mod 2^31-1, 4 blocks, c=128, prefix range sharded", "n_gpus": world,
                          "scaling": "strong", "candidates": plan.candidates, "ms": ms, "candidates_per_s": plan.candidates / ms * 1e3,
                          "best_per_block": best}))
    plan.close()

    # ---- configs[3] ----
    L, R, P = hm.load_fixture("3x4x7_63_rational")
    mkn = hm.LRP2MM(L, R, P)
    (Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
    oplan = capi.OrbitPlan(mkn, Li, Ri, Pi, (dl, dr, dp), capi.MEASURE_NNZ, capi.MODE_PHILOX, SEED)
    B = 1 << 21
    ms = timed(lambda: oplan.run(rank * B, (rank + 1) * B, sp), 5, stream, world, dev)
    g = sharding.allreduce_best(oplan.result(sp), measure_nnz=True, device=dev) if world > 1 else oplan.result(sp)
    if rank == 0:
        print(json.dumps({"config": "orbit + sparsity search, 3x4x7_63_rational, 2^21 Philox candidates per GPU", "n_gpus": world, "scaling": "weak",
                          "candidates": B * world, "ms": ms, "candidates_per_s": B * world / ms * 1e3, "best": g}))
    oplan.close()

    # ---- configs[4] ----
    big = hm.load_large_csr(P31)
    if big is not None:
        mkn, r, (Lc, Rc, Pc) = big
        batch = 4096
        mplan = capi.MMcheckPlan(P31, mkn, r, Lc, Rc, Pc, batch)
        ms = timed(lambda: mplan.run(SEED, rank * batch, sp), 5, stream, world, dev)
        v, ok = mplan.result(sp)
        allok = torch.tensor([int(v == 0 and ok.all())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(allok, op=dist.ReduceOp.MIN)
        nnz = sum(len(x[3]) for x in (Lc, Rc, Pc))
        if rank == 0:
            print(json.dumps({"config": "batched MMchecker mod 2^31-1, 32x32x32_15096, 4096 samples per GPU", "n_gpus": world, "scaling": "weak",
                              "samples": batch * world, "ms": ms, "samples_per_s": batch * world / ms * 1e3,
                              "modmac_per_s": (nnz + r + 32 ** 3) * batch * world / ms * 1e3, "all_samples_agree": bool(allok.item())}))
        mplan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
