#!/usr/bin/env python
"""Kernel-level timing of every BASELINE.json configuration (development aid; bench.py is the contract).

Prints one JSON line per configuration: device time per animate step (CUDA events on the context
stream), vertices/s and the algorithmic / logical HBM rates.  Inputs are resident in HBM.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def time_groups(torch, stream, groups, steps, warmup, flush=None, together=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if together:   # one p3cs_animate_groups call per step (independent groups overlap on side streams)
        from panda3d_b200 import _capi

        class _All:
            def animate(self):
                _capi.animate_groups(groups)
        groups_iter = [_All()]
        return time_groups(torch, stream, groups_iter, steps, warmup, flush)
    for _ in range(warmup):
        for g in groups:
            g.animate()
    stream.synchronize()
    if flush is None:
        e0.record(stream)
        for _ in range(steps):
            for g in groups:
                g.animate()
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / steps
    total = 0.0
    for _ in range(steps):
        flush.add_(1)          # touches a buffer larger than L2 between timed steps
        e0.record(stream)
        for g in groups:
            g.animate()
        e1.record(stream)
        stream.synchronize()
        total += e0.elapsed_time(e1)
    return total / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c2r,c2s,c4,c5,c1")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--instances", type=int, default=10000)
    args = ap.parse_args()
    import torch
    from panda3d_b200 import _capi, synth
    from panda3d_b200.casefile import load_case
    from panda3d_b200.flat import mesh_from_case

    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = _capi.Context(0, stream.cuda_stream)
        flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
        for name in args.configs.split(","):
            specs = []  # (flat mesh, instances)
            small = False
            if name == "c3":
                specs = [(synth.config_c3_mesh().flat, args.instances)]
            elif name == "c3rand":   # the C3 character with a random blend index per vertex (no runs)
                specs = [(synth.config_c3_mesh(index_order="random").flat, args.instances)]
            elif name.startswith("c3fixed"):   # runs of exactly N rows, e.g. c3fixed5
                specs = [(synth.config_c3_mesh(index_order=name[2:]).flat, args.instances)]
            elif name == "c3uv":   # the C3 character with a texcoord column: stride 32, the egg loader's usual layout
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, texcoord=True, index_order="runs").flat, args.instances)]
            elif name == "c3tb":   # ... with tangent + binormal, no texcoord: stride 48
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, tbn=True, index_order="runs").flat, args.instances)]
            elif name == "c3tbuv":  # ... with texcoord, tangent, binormal: stride 56 (the C5 layout)
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, tbn=True, texcoord=True, index_order="runs").flat, args.instances)]
            elif name == "c3a16":  # vertex-animation-align-16 layout (Eigen builds): 4-component columns on 16-byte boundaries
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, align16=True, index_order="runs").flat, args.instances)]
            elif name == "c3a16uv":
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, align16=True, texcoord=True, index_order="runs").flat, args.instances)]
            elif name == "c3a16tbuv":
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, align16=True, tbn=True, texcoord=True, index_order="runs").flat, args.instances)]
            elif name == "c3a16tb":
                specs = [(synth.make_mesh(20_000, 128, 4096, seed=33, align16=True, tbn=True, index_order="runs").flat, args.instances)]
            elif name == "c2r":
                specs, small = [(synth.config_c2("random").flat, 1)], True
            elif name == "c2s":
                specs, small = [(synth.config_c2("sorted").flat, 1)], True
            elif name == "c2x":   # the C2 mesh as a crowd (bandwidth-bound variant of the same kernel work)
                specs = [(synth.config_c2("random").flat, 1000)]
            elif name == "c4":
                specs, small = [(synth.config_c4().flat, 1)], True
            elif name == "c4x":   # the morph mesh as a crowd of 200 faces (bandwidth-bound variant)
                specs = [(synth.config_c4().flat, 200)]
            elif name == "c5":
                specs = [(sm.flat, count) for sm, count in synth.config_c5_meshes()]
            elif name == "c5x":   # the same 4 character types, 100x the crowd
                specs = [(sm.flat, count * 100) for sm, count in synth.config_c5_meshes()]
            elif name == "c1":
                here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
                specs, small = [(mesh_from_case(load_case(os.path.join(here, "tests", "golden", "panda_walk4.p3cs"))), 1)], True
            else:
                continue
            meshes, groups, keep = [], [], []
            verts = out_bytes = logical = 0
            for flat, n in specs:
                dm = _capi.Mesh(ctx, flat)
                g = _capi.Group(dm, n)
                pal = synth.make_palettes(flat.num_transforms, min(n, 64), 77).reshape(min(n, 64), -1)
                reps = (n + pal.shape[0] - 1) // pal.shape[0]
                g.set_palettes(np.tile(pal, (reps, 1))[:n])
                if flat.num_sliders:
                    g.set_sliders(np.tile(synth.make_sliders(flat.num_sliders, 1, 5), (n, 1)))
                meshes.append(dm)
                groups.append(g)
                stride = sum(dm.out_strides)
                verts += n * flat.skinned_rows()
                out_bytes += n * flat.num_rows * stride
                dyn = sum(4 * c.num_values for c in flat.skin_columns)
                logical += n * flat.num_rows * (2 * dyn + 2)
            ctx.sync()
            ms = time_groups(torch, stream, groups, args.steps, args.warmup, flush if small else None)
            if len(groups) > 1:
                ms_together = time_groups(torch, stream, groups, args.steps, args.warmup, flush if small else None, together=True)
                print(json.dumps({"config": name + "+animate_groups", "ms_per_step": ms_together, "gverts_per_s": verts / ms_together / 1e6}), flush=True)
            print(json.dumps({"config": name, "ms_per_step": ms, "vertices": verts, "gverts_per_s": verts / ms / 1e6,
                              "out_gbs": out_bytes / ms / 1e6, "logical_gbs": logical / ms / 1e6,
                              "l2_flush_between_steps": small, "launches": sum(g.timings().launches for g in groups)}), flush=True)
            for g in groups:
                g.close()
            for dm in meshes:
                dm.close()
        ctx.close()


if __name__ == "__main__":
    main()
