"""The drop-in inside real Panda3D: integration/_build/panda_cuda_check (built here by
integration/build.sh against the unmodified reference, shipped to the GPU box) loads
models/panda-model.egg + panda-walk4.egg (BASELINE configs[0]) through the reference's own egg
loader, poses frames, and compares the reference's GeomVertexData::animate_vertices with the shim
(integration/cudaSkinBackend.cxx -> C-ABI -> B200) byte for byte in the same process."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "integration", "_build", "panda_cuda_check")


@pytest.mark.gpu
def test_real_panda_actor_through_the_shim():
    if not os.path.exists(CHECK):
        pytest.skip("integration/_build/panda_cuda_check not built (needs the reference tree: integration/build.sh)")
    models = os.path.join(ROOT, "integration", "_build", "models")
    frames = [str(f) for f in (0, 7, 10, 20, 33, 45, 60)]
    res = subprocess.run([CHECK, os.path.join(ROOT, "panda3d_b200"), os.path.join(models, "panda-model.egg"),
                          os.path.join(models, "panda-walk4.egg")] + frames, capture_output=True, text=True, timeout=300)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"], out
    assert out["frames"] == len(frames) and out["rows"] == 1338
    assert out["same_result_object"]            # same output object reused / cached (geomVertexData.cxx:895-960)
    assert out["max_rel_err"] <= 1e-5           # north_star tolerance; in practice every byte matches
    assert out["differing_bytes"] <= out["bytes_compared"] * 1e-3


@pytest.mark.gpu
def test_real_morph_character_through_the_shim():
    """models/ripple.egg: 7 CharacterVertexSliders over <Dxyz> / <DNormal> morph columns -- the shim's SliderTable /
    morph flattening against loader-made formats, compared in-process with the reference's CPU result."""
    if not os.path.exists(CHECK):
        pytest.skip("integration/_build/panda_cuda_check not built (needs the reference tree: integration/build.sh)")
    ripple = os.path.join(ROOT, "integration", "_build", "models", "ripple.egg")
    if not os.path.exists(ripple):
        pytest.skip("integration/_build/models/ripple.egg missing (re-run integration/build.sh)")
    frames = [str(f) for f in (0, 3, 7, 20, 31, 44)]
    res = subprocess.run([CHECK, os.path.join(ROOT, "panda3d_b200"), ripple, ripple] + frames, capture_output=True, text=True, timeout=300)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"], out
    assert out["frames"] == len(frames) and out["rows"] == 641
    assert out["same_result_object"]
    assert out["differing_bytes"] == 0          # morphs + one rigid joint: no libm on this path


def test_shim_sources_cite_the_reference():
    src = open(os.path.join(ROOT, "integration", "cudaSkinBackend.cxx")).read()
    for needle in ("geomVertexData.cxx", "audioManager.cxx", "p3cs_get_api", "vertex-animation-backend"):
        assert needle in src
