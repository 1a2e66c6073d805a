"""Randomised parity sweep (seeded): formats, index types and orders, rows subsets, blend sizes, morph set-ups,
coordinate systems, palette scales and crowd sizes drawn at random, every case checked against the oracle through
the C-ABI.  Catches interactions the hand-written cases do not enumerate (which kernel mode a layout picks, how
the record groups fall, ragged last tiles)."""
import numpy as np
import pytest

from oracle import oracle
from panda3d_b200 import _capi, synth
from test_gpu_parity import compare_outputs, expected_selection

pytestmark = pytest.mark.gpu


def draw_case(seed: int):
    rng = np.random.default_rng(10_000 + seed)
    rows = int(rng.choice([1, 7, 31, 255, 256, 257, 511, 513, 1000, 2049, 5003]))
    transforms = int(rng.choice([1, 2, 5, 16, 33, 64, 130]))
    k = int(rng.choice([1, 2, 3, 4, 4, 4, 6]))
    k = min(k, transforms)
    max_blends = transforms if k == 1 else 4000
    blends = int(min(max_blends, rng.choice([1, 2, 17, 256, 300, 1500])))
    align16 = bool(rng.random() < 0.3)
    float64 = bool(rng.random() < 0.15)
    tbn = bool(rng.random() < 0.4)
    texcoord = bool(rng.random() < 0.5)
    index_type = str(rng.choice(["u8", "u16", "u32"])) if blends <= 256 else str(rng.choice(["u16", "u32"]))
    index_order = str(rng.choice(["random", "sorted", "runs"]))
    sliders = int(rng.choice([0, 0, 1, 3, 9]))
    morph_columns = [("vertex",), ("vertex", "normal"), ("normal",)][int(rng.integers(0, 3))]
    if tbn and rng.random() < 0.3:
        morph_columns = morph_columns + ("tangent",)
    sub = None
    if rows > 40 and rng.random() < 0.4:
        a = int(rng.integers(0, rows // 3))
        b = int(rng.integers(rows // 2, rows))
        sub = [(a, a + 1), (a + 3, b)] if a + 3 < b else [(a, b)]
    cs = int(rng.choice([1, 1, 2, 3, 4]))
    kw = dict(num_rows=rows, num_transforms=transforms, num_blends=blends, seed=int(rng.integers(1, 1 << 30)),
              weights_per_blend=k, align16=align16, float64=float64, tbn=tbn, texcoord=texcoord, index_type=index_type,
              index_order=index_order, num_sliders=sliders, morph_columns=morph_columns, slider_band=float(rng.choice([0.05, 0.3, 1.0])),
              rows=sub, coordinate_system=cs)
    scale = [1.0, 1.0, 0.6, (1.0, 1.3, 0.7), 2.5][int(rng.integers(0, 5))]
    jitter = float(rng.choice([0.0, 0.0, 0.001, 0.2]))
    instances = int(rng.choice([1, 1, 2, 5]))
    return kw, scale, jitter, instances, int(rng.integers(1, 1 << 30))


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("P3CS_FUZZ_CASES", "48"))))
def test_random_case_matches_oracle(gpu_ctx, seed):
    kw, scale, jitter, n, pseed = draw_case(seed)
    sm = synth.make_mesh(**kw)
    mesh = sm.flat
    pal = synth.make_palettes(mesh.num_transforms, n, pseed, scale=scale, jitter=jitter)
    sl = synth.make_sliders(mesh.num_sliders, n, pseed + 1) if mesh.num_sliders else None
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        assert np.array_equal(grp.read_selection(), expected_selection(mesh)), f"selection differs: {kw}"
        grp.set_palettes(pal.reshape(n, -1))
        if sl is not None:
            grp.set_sliders(sl)
        grp.animate()
        outs = [grp.read_output(o) for o in range(dmesh.num_outputs)]
        om = oracle.OracleMesh(mesh)
        for i in range(n):
            expect, _, _ = om.animate(pal[i], sl[i] if sl is not None else None)
            compare_outputs(mesh, [o[i] for o in outs], expect, f"seed {seed} instance {i}: {kw}")
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.parametrize("seed", range(100, 116))
def test_random_case_matches_oracle_on_the_two_kernel_path(gpu_ctx, monkeypatch, seed):
    """The same sweep through blend_kernel + skin_kernel (P3CS_PATH=wide), the path meshes with palettes too large for
    shared memory take."""
    monkeypatch.setenv("P3CS_PATH", "wide")
    test_random_case_matches_oracle(gpu_ctx, seed)
