#!/usr/bin/env python
"""End-to-end latency of ONE p3cs_animate_vertices call (host buffers in, host buffers out) for the
single-character configurations, beside the oracle port on one core."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from panda3d_b200 import _capi, synth
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case
from oracle import oracle

ctx = _capi.Context(0)
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = {"c1_panda": mesh_from_case(load_case(os.path.join(here, "tests", "golden", "panda_walk4.p3cs"))),
         "c2_random": synth.config_c2("random").flat, "c2_sorted": synth.config_c2("sorted").flat,
         "c4_morph": synth.config_c4().flat}
for name, flat in cases.items():
    dm = _capi.Mesh(ctx, flat); g = _capi.Group(dm, 1)
    pal = synth.make_palettes(flat.num_transforms, 8, 3)
    sl = synth.make_sliders(flat.num_sliders, 8, 4) if flat.num_sliders else [None] * 8
    for i in range(5): g.animate_vertices(pal[i % 8], sl[i % 8])
    n = 200; t0 = time.perf_counter()
    for i in range(n): g.animate_vertices(pal[i % 8], sl[i % 8])
    us = (time.perf_counter() - t0) / n * 1e6
    om = oracle.OracleMesh(flat); t0 = time.perf_counter(); m = 5
    for i in range(m): om.animate(pal[i % 8], sl[i % 8])
    cpu_us = (time.perf_counter() - t0) / m * 1e6
    print(json.dumps({"config": name, "rows": flat.num_rows, "cuda_e2e_us_per_call": us, "oracle_port_1core_us": cpu_us,
                      "mverts_per_s_e2e": flat.num_rows / us}), flush=True)
    g.close(); dm.close()
