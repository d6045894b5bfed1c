"""world_size-2 (and 3) gloo runs of the sharding logic on the CPU: frame shards partition the batch, and the
train-sharded top-2 + all-gather + ordered merge equals the unsharded scan (oracle used as the per-shard checker)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rumi_slam_b200.sharding import frame_shard, merge_top2_host, train_shard


def test_frame_shard_partitions():
    for n in (0, 1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [frame_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q_np, t_np, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orb_oracle
    b, e = train_shard(len(t_np), rank, world)
    i1, d1, d2 = orb_oracle.hamming_top2(q_np, t_np[b:e])
    i1 = np.where(i1 >= 0, i1 + b, -1)                          # global train indices
    packed = (d1.astype(np.int64) << 48) | (d2.astype(np.int64) << 32) | (i1.astype(np.int64) & 0xFFFFFFFF)
    gathered = [torch.zeros(len(q_np), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(packed))
    parts = []
    for g in gathered:                                         # rank order == ascending train ranges
        g = g.numpy()
        idx = (g & 0xFFFFFFFF).astype(np.uint32).astype(np.int64)
        idx = np.where(idx == 0xFFFFFFFF, -1, idx).astype(np.int32)
        parts.append((idx, ((g >> 48) & 0xFFFF).astype(np.uint16), ((g >> 32) & 0xFFFF).astype(np.uint16)))
    mi, m1, m2 = merge_top2_host(parts)
    if rank == 0:
        ret.put((mi, m1, m2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_full_scan(oracle, world):
    rng = np.random.default_rng(7)
    T = rng.integers(0, 256, (701, 32), dtype=np.uint8)
    Q = T[rng.integers(0, 701, 300)] ^ np.packbits(rng.random((300, 256)) < 0.08, axis=1, bitorder="little")
    T[500] = T[20]                                             # exact duplicate in a later shard: earliest index wins
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Q, T, ret)) for r in range(world)]
    for p in procs:
        p.start()
    mi, m1, m2 = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ri, r1, r2 = oracle.hamming_top2(Q, T)
    assert np.array_equal(mi, ri) and np.array_equal(m1, r1) and np.array_equal(m2, r2)
