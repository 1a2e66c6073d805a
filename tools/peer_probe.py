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
n_total = 6
mine = list(crowd.shard_instances(n_total, world, rank))
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = _capi.Context(rank, stream.cuda_stream)
    dm = _capi.Mesh(ctx, flat)
    pitch = (flat.num_rows * 24 + 255) // 256 * 256
    per = -(-n_total // world)
    target = crowd.RenderTarget(ctx, world, rank, per, pitch, root=0)
    pal = synth.make_palettes(flat.num_transforms, n_total, 5)
    g = _capi.Group(dm, len(mine), outputs_dev=[target.slab_ptr])
    g.set_palettes(pal[mine])
    g.animate()
    ctx.sync()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        om = oracle.OracleMesh(flat)
        ok = True
        for i in range(n_total):
            r, j = i % world, i // world
            got = target.read_row(r, j, flat.num_rows * 24).reshape(flat.num_rows, 24)
            ok &= bool(np.array_equal(got, om.animate(pal[i])[0][0]))
        print("PEER_PROBE path=%s ok=%s" % (os.environ.get("P3CS_PATH", "strip"), ok), flush=True)
    g.close(); dm.close()
    dist.barrier()
    target.close()
dist.barrier()
dist.destroy_process_group()
