"""Crowds of independent Character instances over the GPUs of one box.

The path shards naturally: an instance (its palette, slider values and output arrays) depends on
nothing outside itself, so instances are dealt to ranks with no data-path collective.  The only
exchange is the gather of finished vertex arrays to the rendering GPU, done with NCCL through
`torch.distributed` (plumbing) on the very buffers the kernel writes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def shard_instances(num_instances: int, world_size: int, rank: int) -> range:
    """Round-robin deal: instance i lives on rank i % world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside the world")
    return range(rank, num_instances, world_size)


def shard_sizes(num_instances: int, world_size: int) -> List[int]:
    return [len(shard_instances(num_instances, world_size, r)) for r in range(world_size)]


def balance_mixed_crowd(vertices_per_instance: Sequence[int], world_size: int) -> List[List[int]]:
    """Mixed crowds (BASELINE configs[4]): longest-processing-time bin packing on vertex counts, so
    every GPU gets about the same number of vertices per frame.  Returns instance ids per rank."""
    order = sorted(range(len(vertices_per_instance)), key=lambda i: (-vertices_per_instance[i], i))
    loads = [0] * world_size
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += vertices_per_instance[i]
    for b in bins:
        b.sort()
    return bins


def sink_aware_split(num_instances: int, world_size: int, compute_ms_all: float, gather_ms_all: float, root: int = 0) -> List[int]:
    """Instances per rank when every rank's finished rows must end up on ONE GPU (`root`, the rendering GPU).

    The other ranks' rows cross NVLink into the root, whose inbound link carries at most total_bytes / gather_ms_all;
    the root's own rows cost it only compute.  With a share f on the root, the root needs f * compute_ms_all and the
    remote ranks together (1 - f) * max(gather_ms_all, compute_ms_all / (world_size - 1)); the two finish together at
    f = G / (C + G).  An even deal (f = 1 / world_size) makes the root's link the bottleneck: on 8 B200 it was 4x
    SLOWER than not sharding at all.  `compute_ms_all`: one GPU animating the whole crowd;
    `gather_ms_all`: the whole crowd's rows crossing the root's inbound link."""
    if world_size <= 1:
        return [num_instances]
    remote_ms_all = max(gather_ms_all, compute_ms_all / (world_size - 1))
    f = remote_ms_all / (compute_ms_all + remote_ms_all)
    f = min(1.0, max(1.0 / world_size, f))
    n_root = min(num_instances, int(round(f * num_instances)))
    rest = num_instances - n_root
    others = [rest // (world_size - 1) + (1 if k < rest % (world_size - 1) else 0) for k in range(world_size - 1)]
    counts = others[:root] + [n_root] + others[root:]
    assert sum(counts) == num_instances and len(counts) == world_size
    return counts


def split_firsts(counts: Sequence[int]) -> List[int]:
    """First instance of every rank when the ranks hold consecutive ranges of `counts` instances."""
    firsts, acc = [], 0
    for c in counts:
        firsts.append(acc)
        acc += c
    return firsts


def global_order(num_instances: int, world_size: int) -> np.ndarray:
    """Position of every instance in the rank-major gathered buffer: gathered[r, j] holds instance
    r + j * world_size.  Returns an index array `pos` with gathered.reshape(-1)[pos[i]] = instance i
    when every rank's block is padded to ceil(num_instances / world_size) rows."""
    per = -(-num_instances // world_size)
    i = np.arange(num_instances)
    return (i % world_size) * per + i // world_size


def gather_outputs(local, num_instances: int, root: int = 0, group=None, recv=None, reorder: bool = True):
    """Gathers every rank's finished output rows ([n_local, row_bytes] uint8 tensor, device or CPU)
    to `root` with one collective.  Shards are padded to equal length because the collective wants
    equal contributions.  On `root` returns the rows in instance order ([num_instances, row_bytes]),
    or with reorder=False the raw rank-major buffer [world, ceil(n/world), row_bytes] (instance i
    is at [i % world, i // world]: a renderer can index that directly and skip the extra copy).
    `recv` may pass a preallocated rank-major buffer.  Other ranks get None."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-num_instances // world)
    n_local = len(shard_instances(num_instances, world, rank))
    if local.shape[0] != n_local:
        raise ValueError(f"rank {rank} holds {local.shape[0]} instances, expected {n_local}")
    if n_local < per:
        pad = torch.zeros((per - n_local,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send = torch.cat([local, pad], 0)
    else:
        send = local.contiguous()
    if rank == root:
        if recv is None:
            recv = torch.empty((world, per) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.gather(send, [recv[r] for r in range(world)], dst=root, group=group)
        if not reorder:
            return recv
        pos = torch.as_tensor(global_order(num_instances, world), device=local.device)
        return recv.reshape((world * per,) + tuple(local.shape[1:])).index_select(0, pos)
    dist.gather(send, None, dst=root, group=group)
    return None


class RenderTarget:
    """The fused gather: ONE buffer [world, per_rank, row_bytes] in the rendering GPU's HBM, mapped into
    every rank's address space with CUDA IPC (cudaIpcMemLazyEnablePeerAccess).  Each rank hands its slab
    (`slab_ptr`) to p3cs_group_create as the output buffer, so the strip kernel's TMA bulk stores travel
    over NVLink straight into the renderer's memory while the next tiles are being computed; no collective
    follows.  The 64-byte IPC handle is exchanged through torch.distributed (plumbing)."""

    def __init__(self, ctx, world_size: int, rank: int, per_rank: int, row_bytes: int, root: int = 0, group=None,
                 counts: Optional[Sequence[int]] = None):
        """`counts`: instances per rank for uneven (sink-aware) splits -- rank r's slab then starts at instance
        sum(counts[:r]) and the buffer holds sum(counts) rows; else every rank gets `per_rank` rows."""
        import torch
        import torch.distributed as dist

        self.ctx, self.rank, self.root = ctx, rank, root
        self.per_rank, self.row_bytes, self.world_size = per_rank, row_bytes, world_size
        self.slab_bytes = per_rank * row_bytes
        self.firsts = split_firsts(counts) if counts is not None else [r * per_rank for r in range(world_size)]
        total_rows = sum(counts) if counts is not None else world_size * per_rank
        handle = torch.zeros(64, dtype=torch.uint8)
        self.base = 0
        if rank == root:
            self.base = ctx.device_alloc(max(1, total_rows) * row_bytes)
            handle = torch.frombuffer(bytearray(ctx.ipc_export(self.base)), dtype=torch.uint8).clone()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        h = handle.to(dev)
        dist.broadcast(h, src=root, group=group)
        if rank != root:
            self.base = ctx.ipc_open(bytes(h.cpu().numpy().tobytes()))
        dist.barrier(group)

    @property
    def slab_ptr(self) -> int:
        return self.base + self.firsts[self.rank] * self.row_bytes

    def read_row(self, slab: int, index: int, nbytes: int) -> np.ndarray:
        return self.ctx.device_read(self.base + (self.firsts[slab] + index) * self.row_bytes, nbytes)

    def close(self) -> None:
        if self.base:
            if self.rank == self.root:
                self.ctx.device_free(self.base)
            else:
                self.ctx.ipc_close(self.base)
            self.base = 0


class CrowdShard:
    """The instances of one mesh that live on this rank's GPU: device buffers owned by torch
    (so NCCL can send them), animation through the C-ABI group."""

    def __init__(self, ctx, flat_mesh, num_instances: int, world_size: int = 1, rank: int = 0):
        import torch
        from . import _capi

        self.num_instances = num_instances
        self.world_size, self.rank = world_size, rank
        self.instances = shard_instances(num_instances, world_size, rank)
        self.n_local = len(self.instances)
        self.mesh = _capi.Mesh(ctx, flat_mesh)
        self.flat = flat_mesh
        dev = torch.device("cuda", ctx.device)
        T = flat_mesh.num_transforms
        self.palettes = torch.empty((max(self.n_local, 1), T * 16), dtype=torch.float32, device=dev)
        self.outputs = []
        for o, stride in enumerate(self.mesh.out_strides):
            pitch = (flat_mesh.num_rows * stride + 255) // 256 * 256
            self.outputs.append(torch.empty((max(self.n_local, 1), pitch), dtype=torch.uint8, device=dev))
        self.group = _capi.Group(self.mesh, max(self.n_local, 1), palettes_dev=self.palettes.data_ptr(),
                                 outputs_dev=[t.data_ptr() for t in self.outputs]) if self.n_local else None

    def animate(self) -> None:
        if self.group is not None:
            self.group.animate()

    def gather(self, output: int = 0, root: int = 0):
        """NCCL gather of this output array to the rendering GPU (instance order, padded rows)."""
        return gather_outputs(self.outputs[output][:self.n_local], self.num_instances, root)

    def close(self) -> None:
        if self.group is not None:
            self.group.close()
        self.mesh.close()
