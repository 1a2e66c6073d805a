#!/usr/bin/env python
"""Where one small-mesh p3cs_animate_vertices call spends its host time (development aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panda3d_b200 import _capi, synth
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case

ctx = _capi.Context(0)
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = {"c1_panda": mesh_from_case(load_case(os.path.join(here, "tests", "golden", "panda_walk4.p3cs"))),
         "c2_random": synth.config_c2("random").flat}
L = _capi.lib()
for name, flat in cases.items():
    dm = _capi.Mesh(ctx, flat); g = _capi.Group(dm, 1)
    pal = np.ascontiguousarray(synth.make_palettes(flat.num_transforms, 8, 3), np.float32)
    out = np.empty((flat.num_rows, dm.out_strides[0]), np.uint8)
    T = {"set_palettes": 0.0, "animate": 0.0, "read_output": 0.0, "sync": 0.0}
    n = 300
    for it in range(n + 20):
        p = pal[it % 8]
        t0 = time.perf_counter(); L.p3cs_group_set_palettes(g.handle, 0, 1, p.ctypes.data)
        t1 = time.perf_counter(); L.p3cs_group_animate(g.handle)
        t2 = time.perf_counter(); L.p3cs_group_read_output(g.handle, 0, 0, 1, out.ctypes.data)
        t3 = time.perf_counter(); L.p3cs_context_sync(ctx.handle)
        t4 = time.perf_counter()
        if it >= 20:
            T["set_palettes"] += t1 - t0; T["animate"] += t2 - t1; T["read_output"] += t3 - t2; T["sync"] += t4 - t3
    print(name, {k: round(v / n * 1e6, 2) for k, v in T.items()}, "total us", round(sum(T.values()) / n * 1e6, 2), flush=True)
    # the one-call entry point: pageable destination vs registered (direct stores)
    import ctypes as C
    ptrs = (C.c_void_p * 1)(out.ctypes.data)
    pp = [pal[i].ctypes.data for i in range(8)]   # resolved once: numpy's .ctypes costs a microsecond
    fn, gh = L.p3cs_animate_vertices, g.handle
    for label in ("pageable", "registered"):
        if label == "registered":
            ctx.host_register(out)
        for it in range(20):
            L.p3cs_animate_vertices(g.handle, 0, pal[it % 8].ctypes.data, None, ptrs)
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            for it in range(n):
                fn(gh, 0, pp[it & 7], None, ptrs)
            best = min(best, (time.perf_counter() - t0) / n * 1e6)
        print(name, "p3cs_animate_vertices", label, "us/call (best of 5 x %d)" % n, round(best, 2), flush=True)
    ctx.host_unregister(out)
    g.close(); dm.close()
