"""Multi-GPU path on real hardware (skipped with fewer than 2 GPUs): the fused skin + gather -- every
rank's strip kernel stores its finished tiles into ONE buffer on the rendering GPU (CUDA IPC + NVLink peer
access, TMA bulk stores), checked byte for byte against the oracle on rank 0 (tools/peer_probe.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["strip", "wide"])
def test_fused_gather_over_peer_memory(path):
    from panda3d_b200 import _capi
    if _capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, P3CS_PATH=path)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "peer_probe.py")],
                         capture_output=True, text=True, timeout=300, env=env)
    assert f"PEER_PROBE path={path} ok=True" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
