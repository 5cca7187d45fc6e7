"""The C-level multi-device entry points (one process, first ndev devices, plinopt_b200/csrc/multi_device.cu): same answers as the
single-device calls for any ndev.  On a one-GPU box ndev is clamped to 1 (the sharding logic still runs); `gpurun --gpus 2` (or the
driver's multi-GPU boxes) exercises real shards."""
import numpy as np
import pytest

import oracle_lib as O
from plinopt_b200 import hm

pytestmark = pytest.mark.gpu
P31 = 2147483647
SEED = 0x504C494E4F505431


@pytest.mark.parametrize("ndev", [1, 2, 8])
def test_lincomb_search_devices(capi, ndev):
    M = O.dense_fractions("4x4x4_48_rational_L")
    TM = np.array([[(v.numerator % P31) * pow(v.denominator % P31, -1, P31) % P31 for v in col] for col in zip(*[row[:4] for row in M])], dtype=np.int64)
    cf = O.coeffs(TM.tolist(), P31, 24)[0].copy()
    one = capi.lincomb_search(P31, TM, 0, cf)
    assert capi.lincomb_search_devices(ndev, P31, TM, 0, cf) == one
    prev = np.zeros((1, 4), dtype=np.int64); prev[0, 1] = 1
    assert capi.lincomb_search_devices(ndev, P31, TM, 0, cf, prev, -1, -1) == capi.lincomb_search(P31, TM, 0, cf, prev)
    # a seed nobody beats: every shard reports "none"
    assert capi.lincomb_search_devices(ndev, P31, TM, 0, cf, None, 48, 4) == (48, 4, None)


@pytest.mark.parametrize("ndev", [1, 2, 8])
def test_mmcheck_batch_devices(capi, ndev):
    L, R, P = hm.load_fixture("3x3x3_23_58")
    mkn = hm.LRP2MM(L, R, P)
    csr = [hm.csr_modp(M, P31) for M in (L, R, P)]
    v, ok = capi.mmcheck_batch_devices(ndev, P31, mkn, len(L), *csr, seed=3, batch=100)
    v1, ok1 = capi.mmcheck_batch(P31, mkn, len(L), *csr, seed=3, batch=100)
    assert v == v1 == 0 and ok.all() and ok1.all()
    bad = [list(r) for r in P]; bad[2][5] += 1
    csr[2] = hm.csr_modp(bad, P31)
    v, ok = capi.mmcheck_batch_devices(ndev, P31, mkn, len(L), *csr, seed=3, batch=100)
    v1, ok1 = capi.mmcheck_batch(P31, mkn, len(L), *csr, seed=3, batch=100)
    assert v == v1 == 1 and np.array_equal(ok, ok1)  # sample s draws the same Philox inputs whichever device checks it


@pytest.mark.parametrize("ndev", [1, 2, 8])
def test_factor_sweep_devices(capi, ndev):
    M = O.dense_fractions("3x4x7_63_rational_L")
    A = np.array([[(x.numerator % P31) * pow(x.denominator % P31, -1, P31) % P31 for x in row] for row in M], dtype=np.uint32)
    assert capi.factor_sweep_devices(ndev, P31, A, 12, SEED, 0, 5000) == capi.factor_sweep(P31, A, 12, SEED, 0, 5000)
    assert capi.factor_sweep_devices(ndev, P31, A, 12, SEED, 7, 7)[3] is None


def test_engine_comm_single_rank(capi):
    """plo_comm_* (NCCL bound at run time): a one-rank communicator reduces a device table to itself; the N-rank case runs in
    bench.py under torchrun."""
    import torch
    assert capi.lib().plo_comm_nccl_version() >= 20000
    comm = capi.Comm(0, 1, lambda ident: ident)
    t = torch.tensor([5, -7, 2 ** 62, 0], dtype=torch.int64, device="cuda:0")
    comm.allreduce_i64(t.data_ptr(), 4, capi.REDUCE_MIN, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert t.tolist() == [5, -7, 2 ** 62, 0]
    comm.close()
