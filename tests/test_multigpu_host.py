"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of VFO sharding + IQ broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdrpp_b200 import shard


def test_shard_partition_properties(cuda_lib):
    vf = [(48e3, 12.5e3) if i % 2 == 0 else (24e3, 12e3) for i in range(512)]
    costs = [shard.vfo_cost(122.88e6, o, b, cuda_lib.design_resampler, cuda_lib.design_decim_plan) for o, b in vf]
    assert 13.0 < costs[0] < 16.0 and 15.0 < costs[1] < 18.0
    for world in (1, 2, 4, 8, 3):
        sh = shard.shard_vfos(costs, world)
        allv = sorted(i for s in sh for i in s)
        assert allv == list(range(512))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(costs) + 1e-9
    sh = shard.shard_vfos(costs, 8, base_load=[770.0])
    assert sorted(i for s in sh for i in s) == list(range(512)) and len(sh[0]) < len(sh[1])
    assert shard.shard_vfos([], 2) == [[], []]
    assert shard.shard_vfos([1.0], 4) == [[0], [], [], []]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [14.3 if i % 2 == 0 else 16.5 for i in range(40)]
        mine = shard.shard_vfos(costs, world)[rank]
        # every rank sees the same block after the broadcast
        g = torch.Generator().manual_seed(1234)
        blk = torch.randn(7936, 2, generator=g) if rank == 0 else torch.zeros(7936, 2)
        shard.broadcast_block(blk, src=0)
        ref = torch.randn(7936, 2, generator=torch.Generator().manual_seed(1234))
        same = bool(torch.equal(blk, ref))
        # the union of the shards is the full set, gathered without any data-path collective
        owned = torch.zeros(40, dtype=torch.int32)
        owned[mine] = 1
        dist.all_reduce(owned)
        q.put((rank, same, owned.tolist(), len(mine)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shard_and_broadcast():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, owned, n in res:
        assert same, f"rank {rank}: broadcast block differs"
        assert owned == [1] * 40
        assert n == 20


def test_bucket_indexing():
    assert shard.fit_bucket(4, 32) == 4 and shard.fit_bucket(8, 20) == 5 and shard.fit_bucket(99, 6) == 6 and shard.fit_bucket(0, 6) == 1
    seen = []
    for i in range(40):
        k, j, base = shard.bucket_slot(i, 4, 8)
        assert k == (i // 4) % 2 and 0 <= j < 4 and base % 4 == 0 and base < 8
        seen.append(base + j)
    assert seen == [i % 8 for i in range(40)]  # blocks are consumed in stream order


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        NB, KB, n = 8, 4, 64
        src = torch.arange(NB * n * 2, dtype=torch.float32).reshape(NB, n, 2) if rank == 0 else None
        stage = [torch.zeros(KB, n, 2) for _ in range(2)]
        got = []
        for i in range(20):  # the loop of bench.py's step_device for N > 1, without the streams
            k, j, base = shard.bucket_slot(i, KB, NB)
            if j == 0:
                shard.broadcast_block(src[base:base + KB] if rank == 0 else stage[k], src=0)
            blk = src[base + j] if rank == 0 else stage[k][j]
            got.append(float(blk[0, 0]))
        q.put((rank, got))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bucketed_broadcast():
    """Every rank submits the same block sequence when the stream is broadcast in 4-block buckets."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [float((i % 8) * 64 * 2) for i in range(20)]
    assert res[0] == want and res[1] == want
