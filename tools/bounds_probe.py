import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from panda3d_b200 import _capi, synth
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = _capi.Context(0, stream.cuda_stream)
    flat = synth.config_c3_mesh().flat
    dm = _capi.Mesh(ctx, flat)
    for n in (10000, 2500, 1250):
        g = _capi.Group(dm, n)
        pal = synth.make_palettes(flat.num_transforms, 64, 77).reshape(64, -1)
        g.set_palettes(np.tile(pal, ((n + 63) // 64, 1))[:n])
        for mode in ("plain", "bounds", "plain"):
            g.enable_bounds(enable=(mode == "bounds"))
            for _ in range(5): g.animate()
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(20): g.animate()
            e1.record(stream)
            th = (time.perf_counter() - t0) / 20 * 1e3
            stream.synchronize()
            print(n, mode, "gpu ms/step %.4f" % (e0.elapsed_time(e1) / 20), "host ms/call %.4f" % th, "launches", g.timings().launches, flush=True)
        g.close()
