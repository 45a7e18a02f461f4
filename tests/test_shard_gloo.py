"""CPU, world_size 2 over gloo: contig sharding + the summary all-gather reproduce the single-rank table.
The per-contig triples are computed with the oracle here (no GPU); on the GPU box bench.py runs the same
shard code with the CUDA engine over NCCL."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from pandepth_b200.shard import assign_contigs, gather_triples, owned_mask


def test_lpt_assignment_balances_and_is_deterministic():
    lens = [333_333_333] * 9
    own = assign_contigs(lens, 8)
    loads = [sum(l for l, o in zip(lens, own) if o == r) for r in range(8)]
    assert max(loads) == 2 * 333_333_333 and min(loads) == 333_333_333
    assert own == assign_contigs(lens, 8)
    own = assign_contigs([5, 100, 7, 60, 60, 1], 3)
    assert sorted(set(own)) == [0, 1, 2]
    assert assign_contigs([10, 20, 30], 1) == [0, 0, 0]


def _triples(lens, seed):
    from oracle import model
    rng = np.random.default_rng(seed)
    out = np.zeros((len(lens), 3), np.uint64)
    for t, ln in enumerate(lens):
        n = ln // 7
        pos = np.sort(rng.integers(0, ln - 150, n)).astype(np.int32)
        rec = model.Records(np.zeros(n, np.int32), pos, np.zeros(n, np.uint16), np.zeros(n, np.uint8),
                            np.arange(n + 1, dtype=np.uint32), np.full(n, (150 << 4), np.uint32))
        d = model.accumulate(np.array([ln], np.uint32), rec)[0]
        cov, tot = model.reduce_intervals(d, [0], [ln])
        out[t] = (ln, cov[0], tot[0])
    return out


def _worker(rank, world, port, lens, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owner = assign_contigs(lens, world)
    full = _triples(lens, 3)
    local = full * owned_mask(owner, rank)[:, None].astype(np.uint64)      # a rank only has its own contigs
    got = gather_triples(local, owner, rank, world)
    q.put((rank, got.tolist()))
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank():
    lens = [50_000, 20_000, 80_000, 1_000, 33_333]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    want = _triples(lens, 3).tolist()
    assert res[0] == want and res[1] == want


# ------------------------------------------------------------------------------------------------------
# range sharding: equal cut of the concatenated genome, boundary reads to both owners, all-reduce of (cov, sum)
# ------------------------------------------------------------------------------------------------------
def test_range_assignment_is_an_equal_partition():
    from pandepth_b200.shard import assign_ranges
    lens = [333_333_333] * 8 + [333_333_336]
    beg, end = assign_ranges(lens, 8)
    owned = (end.astype(np.int64) - beg.astype(np.int64)).clip(min=0)
    assert owned.sum() == sum(lens)
    per_rank = owned.sum(axis=1)
    assert per_rank.max() - per_rank.min() <= 1                       # 9 contigs on 8 ranks: 12.5 % each, not 2:1
    for t, ln in enumerate(lens):                                    # every base of every contig has exactly one owner
        iv = sorted((int(beg[r, t]), int(end[r, t])) for r in range(8) if end[r, t] > beg[r, t])
        assert iv[0][0] == 0 and iv[-1][1] == ln and all(a[1] == b[0] for a, b in zip(iv[:-1], iv[1:]))
    b1, e1 = assign_ranges([10, 0, 5], 1)
    assert b1.tolist() == [[0, 0, 0]] and e1.tolist() == [[10, 0, 5]]


def _range_worker(rank, world, port, lens, q):
    from oracle import model
    from pandepth_b200.shard import assign_ranges, allreduce_sum_u64
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    beg, end = assign_ranges(lens, world)
    rng = np.random.default_rng(11)
    tlen = np.array(lens, np.uint32)
    # the same reads on every rank (boundary reads included); a rank keeps depth only inside its ranges -- what an engine
    # created with pd_create_ranges does -- and reduces ALL intervals clipped to them
    cov = np.zeros(0, np.uint64); tot = np.zeros(0, np.uint64)
    for t, ln in enumerate(lens):
        n = ln // 5
        pos = np.sort(rng.integers(0, ln - 150, n)).astype(np.int32)
        rec = model.Records(np.zeros(n, np.int32), pos, np.zeros(n, np.uint16), np.zeros(n, np.uint8),
                            np.arange(n + 1, dtype=np.uint32), np.full(n, (150 << 4), np.uint32))
        d = model.accumulate(np.array([ln], np.uint32), rec)[0]
        mine = np.zeros_like(d)
        mine[int(beg[rank, t]):int(end[rank, t])] = d[int(beg[rank, t]):int(end[rank, t])]
        a, b = model.tile_contig(ln, 7000)
        c, s = model.reduce_intervals(mine, a - 1, b, 2)
        cov = np.concatenate([cov, c]); tot = np.concatenate([tot, s])
    both = allreduce_sum_u64(np.stack([cov, tot]), world)
    q.put((rank, both.tolist()))
    dist.destroy_process_group()


def test_two_rank_range_sharding_adds_up_to_the_single_rank_result():
    from oracle import model
    lens = [50_000, 20_001, 80_000]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_range_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    rng = np.random.default_rng(11)
    cov = np.zeros(0, np.uint64); tot = np.zeros(0, np.uint64)
    for t, ln in enumerate(lens):
        n = ln // 5
        pos = np.sort(rng.integers(0, ln - 150, n)).astype(np.int32)
        rec = model.Records(np.zeros(n, np.int32), pos, np.zeros(n, np.uint16), np.zeros(n, np.uint8),
                            np.arange(n + 1, dtype=np.uint32), np.full(n, (150 << 4), np.uint32))
        d = model.accumulate(np.array([ln], np.uint32), rec)[0]
        a, b = model.tile_contig(ln, 7000)
        c, s2 = model.reduce_intervals(d, a - 1, b, 2)
        cov = np.concatenate([cov, c]); tot = np.concatenate([tot, s2])
    want = np.stack([cov, tot]).tolist()
    assert res[0] == want and res[1] == want
