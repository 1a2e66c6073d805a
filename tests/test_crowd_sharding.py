"""Multi-GPU host logic on CPU: instance sharding and the gather to the rendering rank, exercised
with world_size 2 over the gloo backend (the data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from panda3d_b200 import crowd


def test_round_robin_shards_partition_the_crowd():
    for n, w in ((10_000, 8), (7, 4), (3, 8), (0, 2)):
        shards = [list(crowd.shard_instances(n, w, r)) for r in range(w)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
        assert crowd.shard_sizes(n, w) == [len(s) for s in shards]


def test_mixed_crowd_is_balanced_by_vertices():
    """BASELINE configs[4]: 4 character types, 2 M vertices per frame."""
    verts = [5_000] * 100 + [20_000] * 25 + [50_000] * 10 + [125_000] * 4
    assert sum(verts) == 2_000_000
    bins = crowd.balance_mixed_crowd(verts, 8)
    assert sorted(i for b in bins for i in b) == list(range(len(verts)))
    loads = [sum(verts[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= 125_000 and max(loads) <= 1.3 * (sum(verts) / 8)


def test_sink_aware_split_balances_compute_against_the_root_link():
    """All rows end on the rendering GPU: it keeps the share that makes its compute time equal to the time the others'
    rows need on its inbound NVLink (an even deal was 4x slower than one GPU)."""
    n = 10_000
    # one GPU: 1.25 ms for the crowd; 4.8 GB over ~720 GB/s of inbound NVLink: 6.7 ms
    counts = crowd.sink_aware_split(n, 8, 1.25, 6.7)
    assert sum(counts) == n and len(counts) == 8
    f = counts[0] / n
    assert abs(f - 6.7 / (1.25 + 6.7)) < 1e-3
    assert max(counts[1:]) - min(counts[1:]) <= 1
    t_root, t_remote = f * 1.25, (1 - f) * 6.7
    assert abs(t_root - t_remote) < 0.01 and t_root < 1.25          # faster than not sharding
    assert crowd.split_firsts(counts)[0] == 0 and crowd.split_firsts(counts)[-1] == n - counts[-1]
    # a free link degenerates into the even deal, a dead link keeps everything at home
    even = crowd.sink_aware_split(n, 4, 4.0, 0.0)
    assert max(even) - min(even) <= 1
    assert crowd.sink_aware_split(n, 4, 1.0, 1e9)[0] == n
    assert crowd.sink_aware_split(7, 1, 1.0, 1.0) == [7]
    # any root
    c2 = crowd.sink_aware_split(1000, 4, 1.0, 3.0, root=2)
    assert c2[2] == max(c2) and sum(c2) == 1000


def test_global_order_inverts_the_deal():
    n, w = 11, 4
    per = -(-n // w)
    gathered = np.full(w * per, -1)
    for r in range(w):
        for j, i in enumerate(crowd.shard_instances(n, w, r)):
            gathered[r * per + j] = i
    assert np.array_equal(gathered[crowd.global_order(n, w)], np.arange(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, row_bytes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = list(crowd.shard_instances(n, world, rank))
        # every instance's "animated rows" are filled with its instance id
        local = torch.tensor([[i % 251] * row_bytes for i in mine], dtype=torch.uint8).reshape(len(mine), row_bytes)
        got = crowd.gather_outputs(local, n, root=0)
        if rank == 0:
            expect = torch.tensor([[i % 251] * row_bytes for i in range(n)], dtype=torch.uint8)
            out.put(bool(torch.equal(got, expect)))
        else:
            assert got is None
            out.put(True)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_gather_to_rendering_rank_gloo_world2(n):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 48, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)
