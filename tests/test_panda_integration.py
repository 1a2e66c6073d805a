"""The drop-in inside real Panda3D: integration/_build/panda_cuda_check (built here by
integration/build.sh against the reference + the 27-line hook of integration/geomVertexData.patch, shipped to the GPU
box) loads
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
    assert out["frames"] == len(frames) + 1 and out["rows"] == 1338   # + the frame AnimateVerticesRequest animated
    assert out["same_result_object"]            # same output object reused / cached (geomVertexData.cxx:895-960)
    assert out["stock_callers_agree"]           # Geom::get_animated_vertex_data lands on the same object (gobj/geom.cxx:287-288)
    assert out["animate_vertices_request"]      # ... and so does AnimateVerticesRequest (gobj/animateVerticesRequest.cxx:28)
    assert out["entries_before_clear"] == 1 and out["entries_after_clear"] == 0   # clear_animated_vertices releases the device copy
    assert out["max_rel_err"] <= 1e-5           # north_star tolerance
    assert out["differing_bytes"] <= 16         # a blend of frame 13 takes the sinf / cosf normal branch (last-place differences)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["panda", "ripple"])
def test_characters_loaded_from_bam_through_the_shim(which, tmp_path):
    """SURVEY 8 f3, the on-disk half: the character goes through the reference's own .bam writer and reader
    (TransformBlendTable::write_datagram / fillin, gobj/transformBlendTable.cxx:210-284, and the SliderTable /
    JointVertexTransform / Character equivalents) before the shim flattens it -- what a shipped application loads."""
    if not os.path.exists(CHECK):
        pytest.skip("integration/_build/panda_cuda_check not built (needs the reference tree: integration/build.sh)")
    models = os.path.join(ROOT, "integration", "_build", "models")
    if which == "panda":
        files, rows = [os.path.join(models, "panda-model.egg"), os.path.join(models, "panda-walk4.egg")], 1338
    else:
        files, rows = [os.path.join(models, "ripple.egg")] * 2, 641
    frames = [str(f) for f in (0, 7, 20, 33)]
    env = {**os.environ, "P3CS_CHECK_BAM": str(tmp_path)}
    res = subprocess.run([CHECK, os.path.join(ROOT, "panda3d_b200")] + files + frames, capture_output=True, text=True, timeout=300, env=env)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"], out
    assert os.path.getsize(os.path.join(str(tmp_path), "p3cs_check_model.bam")) > 10000     # the round trip really happened
    assert out["frames"] == len(frames) + 1 and out["rows"] == rows
    assert out["same_result_object"] and out["stock_callers_agree"] and out["animate_vertices_request"]
    assert out["max_rel_err"] <= 1e-5
    assert out["differing_bytes"] <= 16


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
    assert out["frames"] == len(frames) + 1 and out["rows"] == 641
    assert out["same_result_object"]
    assert out["differing_bytes"] == 0          # morphs + one rigid joint: no libm on this path


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [False, True], ids=["rgba_uint8", "argb_packed"])
def test_loader_made_colour_and_texcoord_morphs_through_the_shim(packed):
    """ripple.egg with <DRGBA> / <Duv> morphs added (integration/make_morph_egg.py): the egg loader makes the colour
    column (4 x uint8; packed_dabc with vertex-colors-prefer-packed) and the texcoord column morph bases.  Inside real
    Panda3D the hook's result equals the CPU loops' byte for byte -- the colour packers truncate after every morph."""
    if not os.path.exists(CHECK):
        pytest.skip("integration/_build/panda_cuda_check not built (needs the reference tree: integration/build.sh)")
    egg = os.path.join(ROOT, "integration", "_build", "models", "ripple_rgba_uv.egg")
    if not os.path.exists(egg):
        pytest.skip("integration/_build/models/ripple_rgba_uv.egg missing (re-run integration/build.sh)")
    frames = [str(f) for f in (0, 3, 7, 20, 31, 44)]
    env = {**os.environ, "P3CS_CHECK_PRC": "vertex-colors-prefer-packed 1\n"} if packed else dict(os.environ)
    res = subprocess.run([CHECK, os.path.join(ROOT, "panda3d_b200"), egg, egg] + frames, capture_output=True, text=True, timeout=300, env=env)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"], out
    assert ("color:4x4" if packed else "color:4x0") in out["columns"] and "texcoord:2x5" in out["columns"], out["columns"]
    assert out["morphs"] == 28                  # 7 sliders x (vertex, normal, color, texcoord)
    assert out["frames"] == len(frames) + 1 and out["rows"] == 641
    assert out["same_result_object"] and out["stock_callers_agree"]
    assert out["integer_column_differing_bytes"] == 0 and out["differing_bytes"] == 0


@pytest.mark.gpu
def test_crowd_of_panda_actors_is_one_launch_per_frame():
    """1000 copies of the panda Actor (NodePath::copy_to: own joints, own vertex data and TransformBlendTable, shared
    arrays), each on its own frame: CudaSkinBackend::animate_frame animates all of them with ONE kernel launch into
    their own result objects; the stock Geom::get_animated_vertex_data calls then hit the cache.  Bytes equal to the
    reference's CPU loops run in the same process on the same objects."""
    if not os.path.exists(CHECK):
        pytest.skip("integration/_build/panda_cuda_check not built (needs the reference tree: integration/build.sh)")
    models = os.path.join(ROOT, "integration", "_build", "models")
    res = subprocess.run([CHECK, os.path.join(ROOT, "panda3d_b200"), os.path.join(models, "panda-model.egg"),
                          os.path.join(models, "panda-walk4.egg"), "--crowd", "1000", "8"], capture_output=True, text=True, timeout=600)
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert line, res.stdout + res.stderr
    out = json.loads(line[-1])
    print(out)
    assert out["ok"], out
    assert out["crowd"] == 1000 and out["rows"] == 1338
    assert out["device_meshes"] == 1            # every copy is an instance of ONE device mesh
    assert out["launches_last_frame"] == 1 and out["batched_last_frame"] == 1000
    assert out["max_rel_err"] <= 1e-5 and out["differing_bytes"] <= out["bytes_compared"] * 1e-5   # (sinf / cosf normal branch)
    assert out["bytes_compared"] >= 1000 * 1338 * 32
    assert out["cuda_ms_per_frame_e2e"] < out["cpu_ms_per_frame"]


def test_shim_sources_cite_the_reference():
    src = open(os.path.join(ROOT, "integration", "cudaSkinBackend.cxx")).read()
    for needle in ("geomVertexData.cxx", "audioManager.cxx", "p3cs_get_api", "vertex-animation-backend"):
        assert needle in src
