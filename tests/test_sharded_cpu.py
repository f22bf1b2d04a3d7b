"""The N>1 host logic on CPU: the exchange plan (C-ABI osb200_sharded_plan, no GPU needed) driven by two gloo ranks
that exchange real keys and finish with the oracle's sort -- the same steps the GPU path takes (histogram ->
all-gather -> plan -> exchange into bucket-major/source-minor slots -> local sort)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_properties_single_process():
    from gpusorting_b200 import sharded

    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        hist = rng.integers(0, 1000, size=(world, 256)).astype(np.uint64)
        hist[:, 40:50] = 0  # empty buckets
        dests = []
        for r in range(world):
            dest, recv_count, recv_off = sharded.plan(hist, r)
            dests.append(dest)
            assert (np.diff(dest) >= 0).all() and dest.min() >= 0 and dest.max() < world
            assert int(recv_count.sum()) == int(hist.sum())
            # slots of (bucket, source) pairs tile every destination exactly
            for q in range(world):
                mine = [d for d in range(256) if dest[d] == q]
                assert int(recv_count[q]) == int(hist[:, mine].sum())
        for r in range(1, world):
            assert np.array_equal(dests[0], dests[r])  # every rank derives the same bucket -> rank map
        # balance: no rank gets more than its fair share plus one bucket
        dest, recv_count, _ = sharded.plan(hist, 0)
        assert recv_count.max() <= hist.sum() / world + hist.sum(0).max()
    # all keys in one bucket: everything goes to a single rank, nothing is lost
    hist = np.zeros((4, 256), np.uint64)
    hist[:, 7] = 100
    dest, recv_count, recv_off = sharded.plan(hist, 2)
    assert int(recv_count.sum()) == 400 and int(recv_off[7]) == 200
    # empty input
    dest, recv_count, _ = sharded.plan(np.zeros((2, 256), np.uint64), 1)
    assert int(recv_count.sum()) == 0


def _worker(rank, world, port, n, q):
    import sys

    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gpusorting_b200 import sharded
        from tests import oraclelib

        orc = oraclelib.load_oracle()
        keys = orc.init_random_u32(n, 0, 10 + rank)
        if rank == 1:
            keys[: n // 2] &= np.uint32(0x0FFFFFFF)  # skew: rank 1 is heavy in the low buckets
        hist = np.bincount(keys >> 24, minlength=256).astype(np.int64)
        allh = [torch.zeros(256, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allh, torch.from_numpy(hist))
        hist_all = np.stack([h.numpy() for h in allh]).astype(np.uint64)
        dest, recv_count, recv_off = sharded.plan(hist_all, rank)
        # "exchange pass": stable partition by the top digit, every (bucket) run goes to its slot at the destination
        part = orc.binning_pass(keys, 24)
        starts = np.concatenate([[0], np.cumsum(hist)])
        outbox = [[] for _ in range(world)]
        for d in range(256):
            if hist[d]:
                outbox[dest[d]].append((int(recv_off[d]), part[starts[d]:starts[d + 1]]))
        inbox = [None] * world
        dist.all_to_all_object = None
        gathered = [None] * world
        dist.all_gather_object(gathered, outbox)
        recv = np.zeros(int(recv_count[rank]), np.uint32)
        filled = np.zeros(int(recv_count[rank]), bool)
        for src in range(world):
            for off, chunk in gathered[src][rank]:
                assert not filled[off:off + chunk.size].any()  # slots never overlap
                recv[off:off + chunk.size] = chunk
                filled[off:off + chunk.size] = True
        assert filled.all()
        # the received layout is the globally stable MSD partition restricted to this rank's buckets
        top = recv >> 24
        assert (np.diff(top.astype(np.int64)) >= 0).all()
        mine = orc.sort_keys(recv)
        allk = [None] * world
        dist.all_gather_object(allk, keys)
        want = np.sort(np.concatenate(allk))
        lo = int(recv_count[:rank].sum())
        ok = np.array_equal(mine, want[lo:lo + mine.size])
        q.put((rank, bool(ok), int(mine.size)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_matches_global_sort():
    world, n = 2, 1 << 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert all(ok for _, ok, _ in res)
    assert sum(sz for _, _, sz in res) == world * n
