"""Host-side logic of the multi-GPU layer on CPU tensors: two gloo processes (world_size 2).
Covers fragment sharding, the work-table merge per class (int64 SUM | fp64 SUM | MIN | MAX), the row
all-to-all and the broadcast.  The device kernels are exercised by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def run2(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _merge(rank, world):
    from hdk_b200 import distributed as D
    E = 37
    # layout: 2 int64-sum accs, 1 fp64-sum acc, 2 min accs, 1 max acc
    rng = np.random.default_rng(100 + rank)
    si = rng.integers(-1000, 1000, 2 * E)
    sf = rng.normal(0, 10, E)
    mn = rng.integers(-10**12, 10**12, 2 * E)
    mx = rng.integers(-10**12, 10**12, E)
    mn[::5] = np.iinfo(np.int64).max     # identities (group not seen on this rank)
    mx[::7] = np.iinfo(np.int64).min
    w = torch.from_numpy(np.concatenate([si, sf.view(np.int64), mn, mx]).astype(np.int64)).view(torch.uint8)
    D.allreduce_work_table(w, 6 * E, 2 * E, 3 * E, 2 * E, E)
    return w.view(torch.int64).numpy().copy(), (si, sf, mn, mx)


def test_work_table_merge_two_ranks():
    out = run2(_merge)
    (w0, a), (w1, b) = out
    assert np.array_equal(w0, w1)
    E = 37
    assert np.array_equal(w0[:2 * E], a[0] + b[0])
    assert np.allclose(w0[2 * E:3 * E].view(np.float64), a[1] + b[1], rtol=0, atol=0) or np.allclose(w0[2 * E:3 * E].view(np.float64), a[1] + b[1])
    assert np.array_equal(w0[3 * E:5 * E], np.minimum(a[2], b[2]))
    assert np.array_equal(w0[5 * E:], np.maximum(a[3], b[3]))


def _exchange(rank, world):
    from hdk_b200 import distributed as D
    # rank r holds rows with key = 10*r + i; rows are grouped by destination (key % world)
    keys = np.arange(10, dtype=np.int64) + 100 * rank
    vals = (keys * 3).astype(np.int32)
    dest = keys % world
    order = np.argsort(dest, kind="stable")
    counts = torch.tensor([int((dest == p).sum()) for p in range(world)])
    cols = [torch.from_numpy(keys[order].copy()).view(torch.uint8), torch.from_numpy(vals[order].copy()).view(torch.uint8)]
    out, n = D.all_to_all_rows(cols, counts, [8, 4])
    k = out[0][: n * 8].view(torch.int64).numpy().copy()
    v = out[1][: n * 4].view(torch.int32).numpy().copy()
    frs = D.shard_fragments(7)
    b = D.broadcast_tensor(torch.arange(16, dtype=torch.uint8) if rank == 0 else None, 16, "cpu")
    return k, v, frs, b.numpy().copy()


def test_row_all_to_all_and_sharding_and_broadcast():
    out = run2(_exchange)
    allk = np.concatenate([np.arange(10) + 100 * r for r in range(2)])
    for r, (k, v, frs, b) in enumerate(out):
        assert sorted(k.tolist()) == sorted(allk[allk % 2 == r].tolist())     # every key on exactly one rank
        assert np.array_equal(v, (k * 3).astype(np.int32))                    # rows stay intact across columns
        assert frs == [i for i in range(7) if i % 2 == r]
        assert b.tolist() == list(range(16))


def test_single_process_is_a_no_op():
    from hdk_b200 import distributed as D
    w = torch.arange(12, dtype=torch.int64)
    D.allreduce_work_table(w.view(torch.uint8), 12, 4, 8, 2, 2)
    assert w.tolist() == list(range(12))
    assert D.shard_fragments(5) == [0, 1, 2, 3, 4] and D.world() == 1 and D.rank() == 0
