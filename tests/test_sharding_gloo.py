"""CPU tests of the N>1 host logic: shard arithmetic and the single min-allreduce over gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plinopt_b200 import sharding as S


def test_shard_ranges_partition():
    for lo, hi, world in [(0, 10, 3), (5, 5, 4), (7, 1 << 40, 8), (0, 3, 8)]:
        parts = [S.shard_range(lo, hi, r, world) for r in range(world)]
        assert parts[0][0] == lo and parts[-1][1] == max(lo, hi)
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0] and a[0] <= a[1]


def test_score_key_is_order_preserving():
    xs = [0.0, 1e-300, 0.5, 12.066164230573415, 12.069541477224684, 17.85300667219901, 1e300]
    ks = [S.score_key(x) for x in xs]
    assert ks == sorted(ks) and [S.key_score(k) for k in ks] == xs


def test_pick_global_tie_goes_to_lowest_index():
    w = S.pack_local(dict(score=12.5, index=900, nnz=40, nno=3), 1, 3)
    w2 = S.pack_local(dict(score=12.5, index=100, nnz=41, nno=2), 0, 3)
    w3 = S.pack_local(None, 2, 3)
    merged = [min(a, b, c) for a, b, c in zip(w, w2, w3)]
    assert S.pick_global(merged, 3)["index"] == 100


def test_lincomb_key_roundtrip_and_order():
    assert S.lincomb_unkey(S.lincomb_key(24, 2, 130)) == (24, 2, 130) and S.lincomb_unkey(S.lincomb_key(-1, -1, None)) == (-1, -1, None)
    assert S.lincomb_key(24, 2, 130) > S.lincomb_key(24, 2, 131) > S.lincomb_key(24, 1, 0) > S.lincomb_key(23, 4, 0)
    assert S.lincomb_key(5, 5, 2 ** 36 - 2) > S.lincomb_key(5, 5, None)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = S.shard_range(0, 1000, rank, world)
    # fake local sweeps: the true minimum 12.0 sits at index 700 (rank 1); rank 0 also has a 12.0 tie at 123
    local = dict(score=12.0, index=123 if rank == 0 else 700, nnz=40 + rank, nno=rank)
    assert lo <= local["index"] < hi
    g = S.allreduce_best(local)
    e = S.allreduce_best(None if rank == 0 else dict(score=3.0, index=999, nnz=1, nno=0))
    f = S.allreduce_factor_best((196, 100, 168, 40 + rank) if rank == 1 else (196, 100, 168, 77))
    f0 = S.allreduce_factor_best((1, 2, 3, None))
    lc = S.allreduce_lincomb([(24, 2, 500 + rank), (10, 1, None) if rank == 0 else (10, 1, 7), (-1, -1, None)])
    sv = S.gather_survivors([dict(index=(1 << 63) + 5 + 10 * rank + j, nnz=36 + j, nno=rank, score=14.5 + j) for j in range(2 - rank)])
    sv0 = S.gather_survivors([])
    q.put((rank, g, e, f, f0, lc, sv, sv0))
    dist.destroy_process_group()


def test_allreduce_min_with_index_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    out = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    for rank, g, e, f, f0, lc, sv, sv0 in out:
        assert sv0 == [] and [(d["index"], d["nnz"], d["nno"], d["score"]) for d in sv] == [
            ((1 << 63) + 5, 36, 0, 14.5), ((1 << 63) + 6, 37, 0, 15.5), ((1 << 63) + 15, 36, 1, 14.5)]
        assert lc == [(24, 2, 500), (10, 1, 7), (-1, -1, None)]  # first maximiser in enumeration order; seed kept if nobody beat it
        assert g["index"] == 123 and g["rank"] == 0 and g["score"] == 12.0 and g["nnz"] == 40
        assert e["index"] == 999 and e["rank"] == 1
        assert f == (196, 100, 168, 41) and f0[3] is None  # tie on the score: lowest index wins; no candidate anywhere
