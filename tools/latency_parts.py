#!/usr/bin/env python
"""Where the microseconds of one small p3cs_animate_vertices call go (development aid): the same call with the
output left on the device, into registered host memory, the bare launch + wait of an empty kernel, and the kernel's own
device time."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from panda3d_b200 import _capi, synth
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case

def best(fn, n=300, reps=7):
    b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        for it in range(n):
            fn(it)
        b = min(b, (time.perf_counter() - t0) / n * 1e6)
    return round(b, 2)

ctx = _capi.Context(0)
L = _capi.lib()
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the floor: launch of an empty kernel + polling wait on a stream
s = torch.cuda.Stream()
x = torch.zeros(1, device="cuda")
with torch.cuda.stream(s):
    def empty(it):
        x.add_(1)
        while not s.query():
            pass
    print("torch 1-element kernel + stream.query() poll us", best(empty), flush=True)
cases = {"c1_panda": mesh_from_case(load_case(os.path.join(here, "tests", "golden", "panda_walk4.p3cs"))),
         "c2_random": synth.config_c2("random").flat}
for name, flat in cases.items():
    dm = _capi.Mesh(ctx, flat); g = _capi.Group(dm, 1)
    pal = np.ascontiguousarray(synth.make_palettes(flat.num_transforms, 8, 3), np.float32)
    pp = [pal[i].ctypes.data for i in range(8)]
    out = np.empty((flat.num_rows, dm.out_strides[0]), np.uint8)
    ptrs = (C.c_void_p * 1)(out.ctypes.data)
    fn, gh, ch = L.p3cs_animate_vertices, g.handle, ctx.handle
    for it in range(30): fn(gh, 0, pp[it & 7], None, None)
    print(name, "animate_vertices, output stays on the device us", best(lambda it: fn(gh, 0, pp[it & 7], None, None)), flush=True)
    ctx.host_register(out)
    for it in range(30): fn(gh, 0, pp[it & 7], None, ptrs)
    print(name, "animate_vertices, registered host output us", best(lambda it: fn(gh, 0, pp[it & 7], None, ptrs)), flush=True)
    ctx.host_unregister(out)
    an, sy = L.p3cs_group_animate, L.p3cs_context_sync
    def plain(it):
        an(gh); sy(ch)
    print(name, "group_animate (2 event records) + blocking sync us", best(plain), flush=True)
    g.animate(); ctx.sync()
    print(name, "launch-to-end device ms (event pair)", g.timings().total_ms, flush=True)
    def launch_only(it):
        an(gh)
    t = best(launch_only, n=200, reps=3); ctx.sync()
    print(name, "group_animate enqueue only (back to back) us", t, flush=True)
    g.close(); dm.close()
