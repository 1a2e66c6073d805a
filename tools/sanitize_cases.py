#!/usr/bin/env python
"""A handful of small animate calls through every kernel family, for compute-sanitizer (memcheck / racecheck):
the run kernel (mbarrier ring, warp-specialised TMA producer, cp.async item staging, shared-memory slot queue), the strip
kernel in its row-layout / aligned / generic modes (mbarrier + TMA tiles), the two-kernel path and the pose kernel --
with morphs, tight bounds, the selection read-back and an instance list.  Every result is checked against the oracle.

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from panda3d_b200 import _capi, synth

CASES = [
    ("run", dict(num_rows=3_000, num_transforms=24, num_blends=400, index_order="runs")),
    ("run", dict(num_rows=2_000, num_transforms=24, num_blends=300, index_order="random", texcoord=True)),
    ("run", dict(num_rows=2_500, num_transforms=24, num_blends=300, index_order="runs", tbn=True, num_sliders=2, slider_band=0.4)),
    ("strip", dict(num_rows=3_000, num_transforms=24, num_blends=400, index_order="runs")),
    ("strip", dict(num_rows=2_000, num_transforms=16, num_blends=200, index_order="runs", tbn=True, texcoord=True, num_sliders=2, slider_band=0.4)),
    ("strip", dict(num_rows=1_500, num_transforms=16, num_blends=200, align16=True, tbn=True)),
    ("strip", dict(num_rows=1_200, num_transforms=16, num_blends=150, float64=True, rows=[(3, 1100)])),
    ("wide", dict(num_rows=1_500, num_transforms=16, num_blends=200, index_order="runs")),
]


def main():
    ctx = _capi.Context(0)
    n = 5
    for k, (path, kw) in enumerate(CASES):
        os.environ["P3CS_PATH"] = path
        flat = synth.make_mesh(seed=40 + k, **kw).flat
        pal = synth.make_palettes(flat.num_transforms, n, 50 + k)
        sl = synth.make_sliders(flat.num_sliders, n, 60 + k) if flat.num_sliders else None
        om = oracle.OracleMesh(flat)
        dm = _capi.Mesh(ctx, flat)
        g = _capi.Group(dm, n)
        g.read_selection()
        g.set_palettes(pal.reshape(n, -1))
        if sl is not None:
            g.set_sliders(sl)
        bounds = not kw.get("align16") and not kw.get("float64")   # tight bounds want a float32 x 3 vertex column
        if bounds:
            g.enable_bounds()
        g.animate()
        out = [g.read_output(o).copy() for o in range(len(dm.out_strides))]
        if bounds:
            g.read_bounds()
            g.enable_bounds(enable=False)
        g.animate_instances([3, 1])
        for i in range(n):
            explained, unexplained, _ = oracle.unexplained_differences(flat, [o[i] for o in out], pal[i], sl[i] if sl is not None else None, om)
            assert unexplained == 0, (path, kw, i)
        g.close()
        dm.close()
        print(f"case {k} ({path}) ok", flush=True)
    os.environ.pop("P3CS_PATH", None)
    # joint hierarchy on the GPU
    flat = synth.make_mesh(1_000, 32, 100, seed=9, index_order="runs").flat
    skel = synth.make_skeleton(32, 16, seed=3)
    dm = _capi.Mesh(ctx, flat)
    g = _capi.Group(dm, 4)
    ds = _capi.Skeleton(ctx, skel)
    g.pose(ds, np.asarray([0, 3, 7, 15], np.int32))
    g.animate()
    g.read_output(0)
    ds.close()
    g.close()
    dm.close()
    ctx.close()
    print("sanitize cases ok", flush=True)


if __name__ == "__main__":
    main()
