"""Multi-GPU plumbing of a sweep: disjoint contiguous index ranges per rank and ONE tiny
min-allreduce that picks the global winner (the reference analogue is the `omp parallel for`
+ `omp critical` of src/orbiter.cpp:272,298).  torch.distributed is plumbing only; no other
inter-GPU traffic exists on this path."""
import struct

INT64_MAX = 2 ** 63 - 1
SLOT = 4  # int64 words per rank: score key, index, nnz, nno


def shard_range(lo, hi, rank, world):
    """Contiguous, disjoint, ascending shards of [lo,hi): shard r precedes shard r+1, so the
    'lowest index among ties' rule equals 'lowest rank among ties'."""
    total = max(0, hi - lo)
    base, rem = divmod(total, world)
    a = lo + rank * base + min(rank, rem)
    return a, a + base + (1 if rank < rem else 0)


def score_key(score):
    """Order-preserving int64 image of a non-negative double (IEEE bits of x >= 0 sort like integers)."""
    if score != score or score < 0:
        raise ValueError("score must be a non-negative number")
    return struct.unpack("<q", struct.pack("<d", float(score)))[0]


def key_score(key):
    return struct.unpack("<d", struct.pack("<q", int(key)))[0]


def pack_local(best, rank, world):
    """world*SLOT int64 words: INT64_MAX everywhere except this rank's slot (score key, index, nnz, nno)."""
    words = [INT64_MAX] * (world * SLOT)
    if best is not None and best.get("index") is not None:
        idx = best["index"]
        if idx > INT64_MAX:
            raise ValueError("candidate index exceeds 63 bits")
        words[rank * SLOT:(rank + 1) * SLOT] = [score_key(best["score"]), idx, best["nnz"], best["nno"]]
    return words


def pick_global(words, world, measure_nnz=False):
    """Deterministic winner of the gathered table: minimum (score[, nno], index)."""
    best = None
    for r in range(world):
        k, idx, nnz, nno = words[r * SLOT:(r + 1) * SLOT]
        if idx == INT64_MAX:
            continue
        key = (k, nno, idx) if measure_nnz else (k, idx)
        if best is None or key < best[0]:
            best = (key, dict(score=key_score(k), index=idx, nnz=nnz, nno=nno, rank=r))
    return None if best is None else best[1]


def allreduce_best(best, measure_nnz=False, device=None):
    """One all_reduce(MIN) over world*SLOT int64 words (256 B at 8 ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None if best is None or best.get("index") is None else dict(best, rank=0)
    rank, world = dist.get_rank(), dist.get_world_size()
    t = torch.tensor(pack_local(best, rank, world), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return pick_global(t.tolist(), world, measure_nnz)


def allreduce_factor_best(best, device=None):
    """Factorizer sweep (plo_factor_sweep) over sharded index ranges: best = (nnz_alt, nno_alt, nnz_cob, index|None).
    Same single all_reduce(MIN) over world*4 int64 words; lexicographic minimum, lowest index among ties."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return best
    rank, world = dist.get_rank(), dist.get_world_size()
    words = [INT64_MAX] * (world * SLOT)
    if best is not None and best[3] is not None:
        words[rank * SLOT:(rank + 1) * SLOT] = [int(v) for v in best]
    t = torch.tensor(words, dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    rows = [tuple(t[r * SLOT:(r + 1) * SLOT].tolist()) for r in range(world)]
    rows = [r for r in rows if r[3] != INT64_MAX]
    return min(rows) if rows else (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, None)


LC_IDX_BITS = 36


def lincomb_key(rl, cl, idx):
    """Order-preserving int64 image of a sparsifier search result: maximum of (rl, cl, -index)
    (strict '>' acceptance of plinopt_sparsify.inl:183-184 = first maximiser in enumeration order).
    idx None (no candidate beat the seed) maps below every real candidate of the same weight."""
    if idx is not None and int(idx) >= (1 << LC_IDX_BITS) - 1:
        raise ValueError("candidate index does not fit the 36-bit key field (c <= 511)")
    low = 0 if idx is None else (1 << LC_IDX_BITS) - 1 - int(idx)
    return ((int(rl) + 1) << 48) | ((int(cl) + 1) << LC_IDX_BITS) | low


def lincomb_unkey(key):
    low = key & ((1 << LC_IDX_BITS) - 1)
    return (key >> 48) - 1, ((key >> LC_IDX_BITS) & 0xFFF) - 1, (None if low == 0 else (1 << LC_IDX_BITS) - 1 - low)


def allreduce_lincomb(results, device=None):
    """results: list of (rl, cl, idx|None), one per independent problem of the batch, from this rank's prefix shard
    (LincombPlan.run_range).  One all_reduce(MAX) over nbatch int64 words returns the global winners."""
    import torch
    import torch.distributed as dist
    keys = [lincomb_key(*r) for r in results]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor(keys, dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        keys = t.tolist()
    return [lincomb_unkey(k) for k in keys]


def gather_survivors(local, device=None):
    """Survivor lists (OrbitPlan.survivors: dicts with index, nnz, nno, score) of disjoint index ranges -> the global list on every
    rank, sorted by index.  Two collectives: an all_gather of the counts and one of the padded (index, nnz<<32|nno, score bits)
    triples; the ranges are disjoint and ascending with the rank, so concatenation in rank order is already sorted."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    n = torch.tensor([len(local)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    buf = torch.zeros((width, 3), dtype=torch.int64)
    for i, sv in enumerate(local):
        buf[i, 0] = _to_i64(sv["index"])
        buf[i, 1] = _to_i64((sv["nnz"] << 32) | sv["nno"])
        buf[i, 2] = struct.unpack("<q", struct.pack("<d", sv["score"]))[0]
    buf = buf.to(device) if device is not None else buf
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for rk in range(world):
        rows = parts[rk][:counts[rk]].cpu().tolist()
        for idx, cnt, sb in rows:
            cnt &= (1 << 64) - 1
            out.append(dict(index=idx & ((1 << 64) - 1), nnz=cnt >> 32, nno=cnt & 0xFFFFFFFF, score=struct.unpack("<d", struct.pack("<q", sb))[0]))
    out.sort(key=lambda d: d["index"])
    return out


def _to_i64(u):
    """uint64 -> the int64 with the same bits (torch has no uint64 collectives)."""
    u &= (1 << 64) - 1
    return u - (1 << 64) if u >= (1 << 63) else u
