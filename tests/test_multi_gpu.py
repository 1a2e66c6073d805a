"""Multi-GPU path on real hardware (skipped with fewer than 2 GPUs): the fused skin + gather -- every
rank's strip kernel stores its finished tiles into ONE buffer on the rendering GPU (CUDA IPC + NVLink peer
access, TMA bulk stores), checked byte for byte against the oracle on rank 0 (tools/peer_probe.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("split", ["even", "sink"])
@pytest.mark.parametrize("path", ["run", "strip", "wide"])
def test_fused_gather_over_peer_memory(path, split):
    """`split` = sink: the sink-aware split (crowd.sink_aware_split) -- consecutive instance ranges, most on the rendering GPU."""
    from panda3d_b200 import _capi
    if _capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, P3CS_PATH=path, PEER_SPLIT=split)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "peer_probe.py")],
                         capture_output=True, text=True, timeout=300, env=env)
    assert f"PEER_PROBE path={path} split={split} ok=True" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("num_devices", [1, 2])
def test_crowd_over_the_devices_of_one_process(num_devices):
    """p3cs_crowd_*: the C face of the multi-GPU crowd -- one process, one context per device, every device's kernel
    stores into the rendering GPU's buffer; the deal is sink-aware and re-split from measured times (p3cs_crowd_tune).
    Every instance equals the oracle whatever device animated it."""
    import numpy as np
    from oracle import oracle
    from panda3d_b200 import _capi, synth
    if _capi.device_count() < num_devices:
        pytest.skip(f"needs {num_devices} GPUs")
    flat = synth.make_mesh(9_000, 48, 900, seed=12, index_order="runs", tbn=True).flat
    n = 23
    crowd = _capi.Crowd(list(range(num_devices)), flat, n, root_share=0.6 if num_devices > 1 else 0.0)
    try:
        ranges = crowd.ranges()
        assert sum(c for _, c in ranges) == n and ranges[0][0] == 0
        if num_devices > 1:
            assert ranges[0][1] == 14 and ranges[1] == (14, 9)
        pal = synth.make_palettes(flat.num_transforms, n, 3).reshape(n, -1)
        om = oracle.OracleMesh(flat)
        for tuned in (False, True):
            if tuned:
                crowd.tune(2)      # re-splits; palettes have to be sent again
                assert sum(c for _, c in crowd.ranges()) == n
            crowd.set_palettes(pal[:5])
            crowd.set_palettes(pal[5:], first=5)     # a span that crosses the device boundary
            crowd.animate()
            crowd.sync()
            assert all(ms >= 0 for ms in crowd.last_ms())
            out = crowd.read_output(0)
            for i in range(n):
                explained, unexplained, _ = oracle.unexplained_differences(flat, [out[i]], pal[i].reshape(-1, 16), None, om)
                assert unexplained == 0, f"instance {i} (tuned={tuned})"
    finally:
        crowd.close()


@pytest.mark.gpu
def test_c_host_drives_a_crowd_over_the_devices(tmp_path):
    """examples/crowd_host.c: a plain C99 host (no Python, no torch in its process) builds a mesh, deals a crowd to every
    GPU of the box, tunes the split and times frames through p3cs_crowd_*."""
    import json
    from panda3d_b200 import _capi
    lib_dir = os.path.dirname(_capi.LIB_PATH)
    exe = tmp_path / "crowd_host"
    subprocess.run(["gcc", "-std=gnu99", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "crowd_host.c"),
                    "-L", lib_dir, "-lp3cudaskin", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)], check=True)
    devices = [str(d) for d in range(min(_capi.device_count(), 8))]
    res = subprocess.run([str(exe), "300", "20000", "128", "5"] + devices, capture_output=True, text=True, timeout=300)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"] and out["devices"] == len(devices) and sum(out["split"]) == 300
    assert out["split"][0] == max(out["split"])      # the rendering GPU keeps the largest share
