#!/usr/bin/env python
"""2-rank probe: can a kernel running on GPU r store its output rows into a buffer that lives on GPU 0?
Run under torchrun with 2 ranks; P3CS_PATH=wide uses plain global stores, the default path TMA bulk stores."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panda3d_b200 import _capi, crowd, synth
from oracle import oracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
flat = synth.make_mesh(6000, 32, 600, seed=9, index_order="runs").flat
n_total = 10
# PEER_SPLIT=sink: the sink-aware split (crowd.sink_aware_split): consecutive instance ranges, most of them on rank 0
weighted = os.environ.get("PEER_SPLIT") == "sink"
counts = crowd.sink_aware_split(n_total, world, 1.0, 4.0, root=0) if weighted else None
firsts = crowd.split_firsts(counts) if weighted else None
mine = list(range(firsts[rank], firsts[rank] + counts[rank])) if weighted else list(crowd.shard_instances(n_total, world, rank))
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = _capi.Context(rank, stream.cuda_stream)
    dm = _capi.Mesh(ctx, flat)
    pitch = (flat.num_rows * 24 + 255) // 256 * 256
    per = -(-n_total // world)
    target = crowd.RenderTarget(ctx, world, rank, per, pitch, root=0, counts=counts)
    pal = synth.make_palettes(flat.num_transforms, n_total, 5)
    g = _capi.Group(dm, max(len(mine), 1), outputs_dev=[target.slab_ptr]) if mine else None
    if g is not None:
        g.set_palettes(pal[mine])
        g.animate()
    ctx.sync()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        om = oracle.OracleMesh(flat)
        ok = True
        for i in range(n_total):
            r, j = (max(k for k in range(world) if firsts[k] <= i and counts[k] > 0), None) if weighted else (i % world, i // world)
            if weighted:
                j = i - firsts[r]
            got = target.read_row(r, j, flat.num_rows * 24).reshape(flat.num_rows, 24)
            ok &= bool(np.array_equal(got, om.animate(pal[i])[0][0]))
        print("PEER_PROBE path=%s split=%s ok=%s" % (os.environ.get("P3CS_PATH", "default"), "sink" if weighted else "even", ok), flush=True)
    if g is not None:
        g.close()
    dm.close()
    dist.barrier()
    target.close()
dist.barrier()
dist.destroy_process_group()
