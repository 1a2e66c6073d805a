#!/usr/bin/env python
"""p3cs_group_animate_host on the bench workload for several chunk sizes (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panda3d_b200 import _capi, synth
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = _capi.Context(0, stream.cuda_stream)
    flat = synth.config_c3_mesh().flat
    n, T = 10000, flat.num_transforms
    dm = _capi.Mesh(ctx, flat)
    g = _capi.Group(dm, n)
    pal = torch.from_numpy(np.tile(synth.make_palettes(T, 100, 1).reshape(100, -1), (100, 1))).pin_memory()
    out = torch.empty((n, flat.num_rows * 24), dtype=torch.uint8).pin_memory()
    for chunk in (0, 250, 500, 1000, 2000, 5000, 10000):
        for _ in range(2):
            g.animate_host_ptr(pal.data_ptr(), None, [out.data_ptr()], chunk); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(4):
            g.animate_host_ptr(pal.data_ptr(), None, [out.data_ptr()], chunk); ctx.sync()
        ms = (time.perf_counter() - t0) / 4 * 1e3
        print("chunk", chunk, "ms/step %.2f" % ms, "GB/s out %.1f" % (out.numel() / ms / 1e6), flush=True)
    # kernel storing straight into page-locked host memory (no download step at all)
    g2 = _capi.Group(dm, n, outputs_dev=[out.data_ptr()])
    out.zero_()
    for _ in range(2):
        g2.set_palettes_ptr(pal.data_ptr(), 0, n); g2.animate(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(4):
        g2.set_palettes_ptr(pal.data_ptr(), 0, n); g2.animate(); ctx.sync()
    ms = (time.perf_counter() - t0) / 4 * 1e3
    print("direct stores to host: ms/step %.2f" % ms, "GB/s out %.1f" % (out.numel() / ms / 1e6), "nonzero", int(out[n // 2].count_nonzero()), flush=True)
