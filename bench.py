#!/usr/bin/env python3
"""bench.py -- candidates scored / s on the DeGroote-orbit sweep of 2x2x2_7_Winograd minimising the
growth factor G2 (BASELINE.json configs[1]), 1..8 B200, one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (engine arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle restatement of
                                                            src/orbiter.cpp:272-324 on host cores)

A step = one sweep of `2^batch_log2` Philox candidates per GPU (weak scaling: every rank owns a
disjoint contiguous index range) + the single min-allreduce that picks the global winner.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x504C494E4F505431
STEM = "2x2x2_7_Winograd"
WORKLOAD = "orbit sweep of 2x2x2_7_Winograd_{L,R,P}, measure G2 (growthfactor.cpp:117-125), Philox4x32-10 candidates"
# algorithmic work per candidate: r*(m^2k+mk^2+k^2n+kn^2+m^2n+mn^2) = 336 int32 MAC for the three
# transforms + r*(mk+kn+mn) = 84 int32 square-accumulates; 21 DSQRT + 14 DMUL + 7 DADD in FP64.
INT_OPS_PER_CAND = 336 + 84
FP64_OPS_PER_CAND = 21 + 14 + 7
METRIC = "candidates scored/sec"
NCU_DRAM_BYTES_PER_LAUNCH = 22784  # ncu --set full, profiles/ncu_r01_orbit_sweep.md: 22.8 KB read + 0 B written per launch
# issued thread instructions per candidate of orbit_sweep8x_kernel<philox>: smsp__inst_executed.sum x 32 / candidates of the
# same capture (2 434 142 576 warp instructions for 2^28 candidates)
NCU_INST_PER_CAND = 2434142576 * 32 / float(1 << 28)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (nvidia-smi, 50 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        nv = self.nv
        sms = sorted(s for s, _ in self.samples)
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        seen = set()
        for _, r in self.samples:
            for k, bit in names.items():
                if r & bit:
                    seen.add(k)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": sms[len(sms) // 2], "sm_max_mhz": mx, "reasons": sorted(seen), "samples": len(sms)}


def load_problem():
    import numpy as np
    from plinopt_b200 import hm
    L, R, P = hm.load_fixture(STEM)
    mkn = hm.LRP2MM(L, R, P)
    (Li, dl), (Ri, dr), (Pi, dp) = (hm.scaled(M, np.int32) for M in (L, R, P))
    return (L, R, P), mkn, (Li, Ri, Pi), (dl, dr, dp)


def host_threads():
    """Every host core this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arms override it explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(fr, target_s, threads=0):
    """Times the oracle (literal restatement of the orbiter loop body, OpenMP over candidates, all host
    threads) on a bounded sample sized for ~target_s seconds.  Returns (candidates/s, cores, sample count)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    L, R, P = fr
    if threads == 0:
        threads = host_threads()
    cores = threads
    t0 = time.perf_counter()
    O.orbit_sweep(L, R, P, 3, 1, SEED, 0, 20000, nthreads=threads, table=False)
    dt = time.perf_counter() - t0
    n = max(20000, int(20000 / max(dt, 1e-6) * target_s))
    t0 = time.perf_counter()
    res = O.orbit_sweep(L, R, P, 3, 1, SEED, 0, n, nthreads=threads, table=False)
    dt = time.perf_counter() - t0
    return n / dt, cores, n, res["best"]


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    fr, mkn, _, _ = load_problem()
    # each "step" = a bounded sample of the same workload (about 4 s of CPU work with every host thread)
    rate0, cores, _, _ = cpu_reference_rate(fr, 1.0)
    n = max(1000, int(rate0 * 4.0))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    for _ in range(args.warmup):
        O.orbit_sweep(*fr, 3, 1, SEED, 0, max(1000, n // 8), nthreads=cores, table=False)
    t0 = time.perf_counter()
    for s in range(args.steps):
        O.orbit_sweep(*fr, 3, 1, SEED, s * n, (s + 1) * n, nthreads=cores, table=False)
    dt = time.perf_counter() - t0
    val = args.steps * n / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "candidates/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64 rationals + f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "candidates_per_step": n, "note": "CPU oracle (port of src/orbiter.cpp:272-324; the reference itself needs LinBox/Givaro and cannot be built here)"},
        "cpu_baseline": {"value": val, "unit": "candidates/s", "cores": cores, "kind": "port", "sample": f"{args.steps} x {n} Philox candidates, OpenMP over candidates"},
        "e2e": {"value": val, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch-log2", type=int, default=31, help="candidates per GPU per step = 2^this")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from plinopt_b200 import capi, sharding

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if capi.device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    capi.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    fr, mkn, (Li, Ri, Pi), dens = load_problem()
    B = 1 << args.batch_log2
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    plan = capi.OrbitPlan(mkn, Li, Ri, Pi, dens, capi.MEASURE_G2, capi.MODE_PHILOX, SEED)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(s, timed_events=None):
        lo = (s * world + rank) * B
        flush.zero_()  # L2 flush between iterations (outside the timed events)
        if timed_events is not None:
            timed_events[0].record(stream)
        plan.run(lo, lo + B, sp)
        best = plan.result(sp)  # 24 B device->host
        g = sharding.allreduce_best(best, device=dev) if world > 1 else best
        if timed_events is not None:
            timed_events[1].record(stream)
        return g

    peaks = capi.measure_peaks(5) if rank == 0 else None
    for s in range(args.warmup):
        step(s)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_evs = []
    overall = None
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        g = step(args.warmup + s, evs[s])
        if g is not None and (overall is None or (g["score"], g["index"]) < (overall["score"], overall["index"])):
            overall = g
    barrier()
    t_wall = time.perf_counter() - t_wall0
    sampler.stop_flag = True
    sampler.join()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total = args.steps * world * B
    value = total / (ms * 1e-3)

    # dominant kernel alone (sweep + final launches, no host sync inside): live roofline numerator
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    k0.record(stream)
    for s in range(reps):
        plan.run((1000 + s) * B, (1001 + s) * B, sp)
    k1.record(stream)
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / reps

    # end to end through the host-buffer C-ABI call (upload of L/R/P, sweep, result download inside)
    h2d = int(Li.nbytes + Ri.nbytes + Pi.nbytes)
    d2h = 24
    e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for s in range(e_steps):
        lo = ((5000 + s) * world + rank) * B
        best = capi.orbit_sweep(mkn, Li, Ri, Pi, dens, capi.MEASURE_G2, capi.MODE_PHILOX, SEED, lo, lo + B)
        if world > 1:
            sharding.allreduce_best(best, device=dev)
    barrier()
    e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.item())
    e2e_val = e_steps * world * B / e_dt

    if rank == 0:
        clocks = sampler.summary()
        # The kernel carries four 8-bit lanes per IMAD and reads the 2x2 matrices and the first product stage from tables, so it does the
        # 420 algorithmic int32 operations of a candidate in ~290 issued instructions: the algorithmic rate is ABOVE the scalar IMAD peak, and the limit that
        # binds is the scheduler's issue rate (IMAD/IDP on the fma-heavy pipe + LOP3 on the alu pipe).  frac = issued instructions per
        # second over the live-measured issue peak; the algorithmic view is kept beside it.
        cand_per_s = B / (kern_ms * 1e-3)
        achieved = NCU_INST_PER_CAND * cand_per_s / 1e12
        peak_measured = peaks["issue_inst_per_s"] / 1e12
        # the microbenchmark reaches 88-96 % of the schedulers' nominal rate depending on the box; the denominator is the LARGER of
        # it and SMs x 4 schedulers x 32 lanes x the SM clock sampled under load, so that frac never flatters the kernel
        peak_clock = None
        try:
            if clocks.get("sm_mhz"):
                peak_clock = torch.cuda.get_device_properties(dev).multi_processor_count * 4 * 32 * float(clocks["sm_mhz"]) * 1e6 / 1e12
        except Exception:
            peak_clock = None
        peak = max(peak_measured, peak_clock) if peak_clock else peak_measured
        alg = INT_OPS_PER_CAND * cand_per_s / 1e12
        roof = {"bound": "int32", "achieved": achieved, "peak": peak, "unit": "T thread-instructions/s (issue slots)", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": "profiles/ncu_r01_orbit_sweep.md (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch; "
                "independent of the candidate count: constants + one 16 B key per block)",
                "kernel": "orbit_sweep8x_kernel<philox> (2x2x2, r = 7: four lanes per IMAD, first product stage from shared-memory tables)", "kernel_ms": kern_ms,
                "instructions_per_candidate": NCU_INST_PER_CAND,
                "instructions_source": "profiles/ncu_r01_orbit_sweep.md (smsp__inst_executed.sum x 32 / candidates of the captured launch)",
                "peak_measured": peak_measured, "peak_from_clock": peak_clock,
                "peak_source": "max(plo_measure_issue_peak: IMAD and LOP3 chains interleaved 1:1, all SMs, best of 5, measured in this run; "
                               "SMs x 4 schedulers x 32 lanes x SM clock sampled under load); MEASURED_PEAKS.json has no int32/fp64 entry",
                "algorithmic": {"ops_per_candidate": {"int32": INT_OPS_PER_CAND, "fp64": FP64_OPS_PER_CAND}, "achieved_tiops": alg,
                                "scalar_imad_peak_tiops": peaks["imad_per_s"] / 1e12, "vs_scalar_imad_peak": alg / (peaks["imad_per_s"] / 1e12),
                                "note": "above 1: one IMAD carries four 8-bit lanes (two rows of the left factor x two Hopcroft-Musinski rows)"},
                "shared_memory": {"wavefronts_per_warp_candidate": 78, "ncu_pct_of_peak": 78.9,
                                  "note": "9 LDS.128 + 21 LDS.64 per candidate, bank-conflict free (profiles/ncu_r01_orbit_sweep.md): the busiest unit after the schedulers"},
                "fp64": {"achieved_tflop": FP64_OPS_PER_CAND * cand_per_s / 1e12, "peak_dfma_tflop": 2 * peaks["dfma_per_s"] / 1e12},
                "hbm_bytes_per_candidate": 16.0 * plan_grid_bytes(B),
                "hbm": hbm_view(kern_ms),
                "bound_note": "the contract's bound classes are hbm|tensor; this kernel is neither: it is bound by instruction issue on the INT32 pipes "
                              "(fma-heavy + alu), so frac is issued instructions over the live-measured issue peak; the hbm view shows how far it is "
                              "from the memory roof"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the contract times the CPU baseline on rank 0 at N = 1 only
            rate, cores, n, _ = cpu_reference_rate(fr, args.cpu_seconds)
            cpu = {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "port", "sample": f"{n} Philox candidates of the same workload (oracle, OpenMP over candidates)"}
        line = {
            "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 + f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "candidates_per_gpu_per_step": B, "seed": hex(SEED), "parallelism": f"index-range sharding x{world} + one min-allreduce",
                       "l2": "256 MiB buffer rewritten between timed iterations (outside the event pairs); the path reads 84 ints from constant memory"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "candidates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                    "api": "plo_orbit_sweep (host buffers)"},
            "gpu_launches": args.steps * plan.launches,
            "roofline": roof, "cpu_baseline": cpu,
            "best": overall, "wall_s": t_wall,
        }
        print(json.dumps(line))
    plan.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def hbm_view(kern_ms):
    """The same launch against the HBM roof of MEASURED_PEAKS.json (ncu dram bytes per launch / kernel time)."""
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        pass
    achieved = NCU_DRAM_BYTES_PER_LAUNCH / (kern_ms * 1e-3) / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if peak else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peak else "MEASURED_PEAKS.json missing"}


def plan_grid_bytes(B):
    """16 B per block key + 24 B result over B candidates (negligible): HBM bytes per candidate / 16."""
    return (148 * 8 + 2) / B


if __name__ == "__main__":
    sys.exit(main())
