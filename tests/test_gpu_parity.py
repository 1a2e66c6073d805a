"""GPU parity tests: the CUDA path, called through the C-ABI (libp3cudaskin.so), against
  (1) the committed golden fixtures = bytes the UNMODIFIED reference produced, and
  (2) the oracle (oracle/skin_oracle.c, itself pinned bit-exact to the reference) on seeded inputs.

Bars (BASELINE.json north_star): blend-index / row selection bit-exact; float32 outputs within
1e-5 relative (tolerance written below); static bytes identical.  The kernels are built without
FMA contraction, so in practice every byte matches except through the sinf/cosf/atan2f of the
uniform-scale normal branch; the tests assert the contract and report the max-abs error.
"""
import os

import numpy as np
import pytest

from conftest import golden_ids, golden_paths
from oracle import oracle
from panda3d_b200 import _capi, flat as F, synth
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star: float32 positions/normals agree within 1e-5 relative


def dynamic_float_mask(mesh, numeric_type=None):
    """Byte mask per output array of the float columns the path may modify (all of them, or those of one numeric type)."""
    masks = []
    for a in mesh.output_arrays:
        m = np.zeros(mesh.strides[a], bool)
        cols = list(mesh.skin_columns) + [mo.base for mo in mesh.morphs]
        for c in cols:
            if c.array == a and (numeric_type is None or c.numeric_type == numeric_type):
                m[c.start:c.start + c.num_bytes] = True
        masks.append(m)
    return masks


def compare_outputs(mesh, got, expect, what=""):
    """Static bytes identical; every dynamic column within REL_TOL, judged PER COLUMN (SURVEY 8d "per component
    class"): a component's error is taken relative to max(|component|, largest |component| of the SAME column in
    that row) -- a normal is never measured against the position's magnitude.  Returns (max_abs_error,
    number_of_differing_bytes)."""
    from panda3d_b200.flat import NT_float32, NT_float64
    max_abs, nbytes = 0.0, 0
    all_masks = dynamic_float_mask(mesh)
    columns = []
    for c in list(mesh.skin_columns) + [mo.base for mo in mesh.morphs]:
        key = (c.array, c.start, c.num_values, c.numeric_type)
        if key not in [k for k, _ in columns]:
            columns.append((key, c))
    for o, (g, e, mask) in enumerate(zip(got, expect, all_masks)):
        g = np.asarray(g).reshape(mesh.num_rows, -1)
        e = np.asarray(e).reshape(mesh.num_rows, -1)
        assert g.shape == e.shape
        assert np.array_equal(g[:, ~mask], e[:, ~mask]), f"{what}: static bytes of output {o} differ"
        nbytes += int((g != e).sum())
        for _, c in columns:
            if c.array != mesh.output_arrays[o]:
                continue
            width = c.num_bytes
            if c.numeric_type not in (NT_float32, NT_float64):   # integer and packed columns: index / byte work, bit-exact
                assert np.array_equal(g[:, c.start:c.start + width], e[:, c.start:c.start + width]), \
                    f"{what}: output {o} integer/packed column @{c.start} (type {c.numeric_type}) differs"
                continue
            dt = np.float64 if c.numeric_type == NT_float64 else np.float32
            gf = np.ascontiguousarray(g[:, c.start:c.start + width]).view(dt).astype(np.float64)
            ef = np.ascontiguousarray(e[:, c.start:c.start + width]).view(dt).astype(np.float64)
            if not gf.size:
                continue
            same = np.ascontiguousarray(g[:, c.start:c.start + width]) == np.ascontiguousarray(e[:, c.start:c.start + width])
            with np.errstate(invalid="ignore"):
                err = np.abs(gf - ef)
            err[same.reshape(gf.shape[0], c.num_values, -1).all(axis=2)] = 0.0   # identical bit patterns (inf, NaN included)
            assert np.isfinite(err).all(), f"{what}: output {o} column @{c.start}: non-finite values differ"
            max_abs = max(max_abs, float(err.max()))
            col_scale = np.maximum(np.abs(np.where(np.isfinite(ef), ef, 0.0)).max(axis=1, keepdims=True), 1e-30)
            rel = (err / np.maximum(np.abs(ef), col_scale)).max()
            assert rel <= REL_TOL, f"{what}: output {o} column @{c.start} relative error {rel:.3e} > {REL_TOL}"
    return max_abs, nbytes


def expected_selection(mesh):
    bi = mesh.blend_indices()
    sel = np.full(mesh.num_rows, -1, np.int64)
    if mesh.blend_column is not None:
        for b, e in np.asarray(mesh.row_ranges).reshape(-1, 2):
            sel[b:e] = bi[b:e]
    return sel


# Fixtures built to hit the uniform-scale normal branch (decompose_matrix / compose_matrix, geomVertexData.cxx:1825-1827):
# sinf / cosf / atan2f differ in the last place between libm and CUDA, so these agree within REL_TOL only.  EVERY other
# fixture must come out byte for byte as the reference wrote it -- asserted below.
NOT_BYTE_EXACT = {"rigid_uniform_scale", "align16_uniform_scale", "f64_uniform_scale", "cs_zup_left", "cs_yup_right", "cs_yup_left"}


@pytest.mark.parametrize("path", golden_paths(), ids=golden_ids())
def test_golden_fixture_parity(gpu_ctx, path):
    """CUDA output vs the bytes the reference itself produced (tests/golden/, frames of each case)."""
    case = load_case(path)
    mesh = mesh_from_case(case)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        assert np.array_equal(grp.read_selection(), expected_selection(mesh)), "blend-index / row selection differs"
        total_diff = 0
        for fr in range(int(case["meta"][7])):
            sliders = case["sliders"][fr] if mesh.num_sliders else None
            got = grp.animate_vertices(case["palette"][fr], sliders)
            expect = [case[f"expect{o}"][fr] for o in range(len(got))]
            max_abs, nb = compare_outputs(mesh, got, expect, f"{path} frame {fr}")
            total_diff += nb
            blends = grp.read_blends(0)
            eb = case["expect_blends"][fr]
            cols = slice(None) if dmesh_full(mesh) else [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14]
            assert np.array_equal(blends[:, cols].view(np.uint32), eb.reshape(-1, 16)[:, cols].view(np.uint32)), \
                "blended matrices are not bit-identical to TransformBlend::get_blend"
        print(f"{path.rsplit('/', 1)[1]}: differing bytes vs reference = {total_diff}")
        name = os.path.splitext(os.path.basename(path))[0]
        if name not in NOT_BYTE_EXACT:
            assert total_diff == 0, f"{name}: {total_diff} bytes differ from the reference's output (must be byte-identical)"
    finally:
        grp.close()
        dmesh.close()


def dmesh_full(mesh):
    return any(c.num_values == 4 for c in mesh.skin_columns)


SYNTH_CASES = {
    "c2_small_random": dict(num_rows=20_000, num_transforms=64, num_blends=4096, index_order="random"),
    "c2_small_sorted": dict(num_rows=20_000, num_transforms=64, num_blends=4096, index_order="sorted"),
    "c5_layout": dict(num_rows=9_000, num_transforms=128, num_blends=2000, tbn=True, texcoord=True, index_order="runs"),
    "c4_morphs": dict(num_rows=12_000, num_transforms=64, num_blends=2048, num_sliders=16, slider_band=0.1,
                      morph_columns=("vertex", "normal"), index_order="runs"),
    "ragged_rows": dict(num_rows=5_003, num_transforms=16, num_blends=300, rows=[(1, 2), (5, 1000), (1003, 5003)]),
    "u32_index": dict(num_rows=70_001, num_transforms=32, num_blends=66_000, index_type="u32"),
    "one_row": dict(num_rows=1, num_transforms=4, num_blends=1),
    "align16": dict(num_rows=4_099, num_transforms=32, num_blends=700, align16=True, tbn=True),
    "row32_layout": dict(num_rows=7_013, num_transforms=48, num_blends=1500, texcoord=True, index_order="runs"),
    "row48_layout": dict(num_rows=7_013, num_transforms=48, num_blends=1500, tbn=True, index_order="runs",
                         rows=[(9, 7000)]),
    "row32_morphs": dict(num_rows=3_000, num_transforms=16, num_blends=300, texcoord=True, num_sliders=4, slider_band=0.3,
                         morph_columns=("vertex", "normal")),
    "align16_morphs": dict(num_rows=5_001, num_transforms=24, num_blends=600, align16=True, texcoord=True, index_order="runs",
                           num_sliders=5, slider_band=0.3, morph_columns=("vertex", "normal")),
    "align16_uv_tbn": dict(num_rows=6_007, num_transforms=40, num_blends=900, align16=True, texcoord=True, tbn=True,
                           index_order="runs", rows=[(2, 6000)]),
    "float64_columns": dict(num_rows=6_001, num_transforms=32, num_blends=900, float64=True, tbn=True, index_order="runs",
                            num_sliders=6, slider_band=0.2, morph_columns=("vertex", "normal")),
    "float64_align16": dict(num_rows=3_001, num_transforms=16, num_blends=500, float64=True, align16=True,
                            rows=[(3, 2900)], num_sliders=2, slider_band=0.6),
}


@pytest.mark.parametrize("name", sorted(SYNTH_CASES))
def test_synthetic_parity_vs_oracle(gpu_ctx, name):
    kw = dict(SYNTH_CASES[name])
    sm = synth.make_mesh(seed=sum(map(ord, name)) % 1000, **kw)
    mesh = sm.flat
    frames = 2
    pal = synth.make_palettes(mesh.num_transforms, frames, 99, jitter=0.3 if "align16" in name else 0.0)
    sl = synth.make_sliders(mesh.num_sliders, frames, 98) if mesh.num_sliders else None
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        assert np.array_equal(grp.read_selection(), expected_selection(mesh))
        for fr in range(frames):
            expect, _, sel = om.animate(pal[fr], sl[fr] if sl is not None else None, want_selection=True)
            got = grp.animate_vertices(pal[fr], sl[fr] if sl is not None else None)
            max_abs, nb = compare_outputs(mesh, got, expect, f"{name} frame {fr}")
            print(f"{name} frame {fr}: max-abs error {max_abs:.3e}, differing bytes {nb}")
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.parametrize("texcoord", [False, True])
def test_vector_columns_listed_binormal_first(gpu_ctx, texcoord):
    """Skin columns handed over as vertex, normal, BINORMAL, TANGENT (the order of a caller's column list is free),
    with morphs on both vector columns: the row-layout modes (strides 48 / 56) must still add every delta to the
    column it belongs to (the kRow48 / kRow56 morph loop keyed columns by dyn index)."""
    sm = synth.make_mesh(4_000, 24, 500, seed=61, tbn=True, texcoord=texcoord, index_order="runs", num_sliders=3, slider_band=0.4,
                         morph_columns=("tangent", "binormal", "vertex"))
    mesh = sm.flat
    v, n, t, b = mesh.skin_columns
    assert t.start < b.start
    mesh.skin_columns = [v, n, b, t]
    pal = synth.make_palettes(mesh.num_transforms, 2, 62)
    sl = synth.make_sliders(3, 2, 63)
    sl[:, 0] = 0.75   # never all zero
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        for fr in range(2):
            expect, _, _ = om.animate(pal[fr], sl[fr])
            got = grp.animate_vertices(pal[fr], sl[fr])
            max_abs, nb = compare_outputs(mesh, got, expect, f"binormal-first frame {fr}")
            assert nb == 0, f"{nb} bytes differ from the oracle"
    finally:
        grp.close()
        dmesh.close()


def test_crowd_instances_are_independent(gpu_ctx):
    """A group of n instances == n single-instance calls (crowd sharding unit)."""
    sm = synth.make_mesh(3_000, 32, 500, seed=5, index_order="runs", num_sliders=4, slider_band=0.3)
    mesh = sm.flat
    n = 37
    pal = synth.make_palettes(mesh.num_transforms, n, 7)
    sl = synth.make_sliders(mesh.num_sliders, n, 8)
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        grp.set_palettes(pal)
        grp.set_sliders(sl)
        grp.animate()
        out = grp.read_output(0)
        for i in (0, 1, n // 2, n - 1):
            expect, _, _ = om.animate(pal[i], sl[i])
            compare_outputs(mesh, [out[i]], expect, f"instance {i}")
        t = grp.timings()
        assert t.launches >= 1 and t.total_ms > 0
    finally:
        grp.close()
        dmesh.close()


def test_full_size_c2_properties(gpu_ctx):
    """BASELINE configs[1] at full size (100k vertices, 64 joints, 16 384 blends): size-independent
    properties -- identity palette is a byte-exact copy; a rigid global translation moves every point
    by it and leaves normals; linearity of blending in the palette (M -> 2M doubles positions)."""
    sm = synth.config_c2("random")
    mesh = sm.flat
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        T = mesh.num_transforms
        ident = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (T, 1))
        out = grp.animate_vertices(ident)[0]
        src = mesh.arrays[0]
        f_out, f_src = out.view(np.float32), src.view(np.float32)
        # weights sum to 1 within an ulp, so identity blends reproduce positions to ~1e-7 relative
        assert np.allclose(f_out, f_src, rtol=5e-7, atol=1e-7)
        tr = ident.copy().reshape(T, 4, 4)
        tr[:, 3, :3] = (1.0, -2.0, 0.5)
        out_t = grp.animate_vertices(tr.reshape(T, 16))[0].view(np.float32)
        assert np.allclose(out_t[:, :3] - f_out[:, :3], (1.0, -2.0, 0.5), atol=2e-6)
        assert np.array_equal(out_t[:, 3:6], f_out[:, 3:6])
        pal = synth.make_palettes(T, 1, 3)[0]
        a = grp.animate_vertices(pal)[0].view(np.float32)
        b = grp.animate_vertices(pal * 2.0)[0].view(np.float32)
        assert np.array_equal(b[:, :3], a[:, :3] * 2.0)  # scaling by 2 is exact in floating point
        # and the full-size result agrees with the oracle
        expect, _, _ = oracle.OracleMesh(mesh).animate(pal)
        compare_outputs(mesh, [grp.animate_vertices(pal)[0]], expect, "C2 full size")
    finally:
        grp.close()
        dmesh.close()


def test_unsupported_and_invalid_inputs_fail_loudly(gpu_ctx):
    sm = synth.make_mesh(100, 4, 8, seed=1)
    mesh = sm.flat
    bad = F.FlatMesh(**{**mesh.__dict__})
    # combinations the reference's own packers refuse (nassert): a packed colour word of three values, a signed colour
    for column in (F.Column(0, 0, 3, F.NT_packed_dcba, F.C_point), F.Column(0, 0, 4, F.NT_packed_ufloat, F.C_vector),
                   F.Column(0, 0, 3, F.NT_stdfloat, F.C_point), F.Column(0, 2, 3, F.NT_float32, F.C_point)):
        bad.skin_columns = [column]
        with pytest.raises(_capi.P3csError) as e:
            _capi.Mesh(gpu_ctx, bad)
        assert e.value.code == _capi.P3CS_ERR_UNSUPPORTED
    sm_color = synth.make_typed_mesh(64, 4, 8, seed=2, num_sliders=1, columns=[("vertex", 3, F.NT_float32, F.C_point), ("color", 4, F.NT_int8, F.C_color)],
                                     morphs=[("color", 4, F.NT_float32)])
    with pytest.raises(_capi.P3csError) as e:
        _capi.Mesh(gpu_ctx, sm_color.flat)
    assert e.value.code == _capi.P3CS_ERR_UNSUPPORTED
    bad2 = F.FlatMesh(**{**mesh.__dict__})
    bad2.blend_offsets = mesh.blend_offsets[:3].copy()  # fewer blends than the indices select
    bad2.blend_transform = mesh.blend_transform[:int(bad2.blend_offsets[-1])]
    bad2.blend_weight = mesh.blend_weight[:int(bad2.blend_offsets[-1])]
    with pytest.raises(_capi.P3csError) as e:
        _capi.Mesh(gpu_ctx, bad2)
    assert e.value.code == _capi.P3CS_ERR_INVALID


def test_mirror_animate_vertices_semantics(gpu_ctx):
    """The host-side mirror of GeomVertexData::animate_vertices (panda3d_b200/gobj.py): written the way a
    test against panda3d.core would be (cf. the survey's probe): morph first, then skin;
    rows outside TransformBlendTable::get_rows() untouched; the SAME output object is returned, unchanged
    while nothing is modified and mutated in place when a slider moves (geomVertexData.cxx:895-960)."""
    from panda3d_b200 import gobj
    from panda3d_b200.gobj import (GeomEnums, GeomVertexAnimationSpec, GeomVertexArrayFormat, GeomVertexData,
                                   GeomVertexFormat, InternalName, SliderTable, SparseArray, TransformBlend,
                                   TransformBlendTable, UserVertexSlider, UserVertexTransform)
    gobj._Backend.ctx = gpu_ctx
    af = GeomVertexArrayFormat()
    af.add_column(InternalName.get_vertex(), 3, GeomEnums.NT_float32, GeomEnums.C_point)
    af.add_column(InternalName.get_normal(), 3, GeomEnums.NT_float32, GeomEnums.C_normal)
    af.add_column(InternalName.get_texcoord(), 2, GeomEnums.NT_float32, GeomEnums.C_texcoord)
    bf = GeomVertexArrayFormat()
    bf.add_column(InternalName.get_transform_blend(), 1, GeomEnums.NT_uint16, GeomEnums.C_index, 0, 2)
    dname = InternalName.get_morph(InternalName.get_vertex(), "smile")
    bf.add_column(dname, 3, GeomEnums.NT_float32, GeomEnums.C_morph_delta)
    f = GeomVertexFormat(af)
    f.add_array(bf)
    spec = GeomVertexAnimationSpec()
    spec.set_panda()
    f.set_animation(spec)
    fmt = GeomVertexFormat.register_format(f)
    vd = GeomVertexData("probe", fmt, GeomEnums.UH_static)
    vd.set_num_rows(4)
    t0 = UserVertexTransform("j0")
    scale_translate = np.diag([1, 2, 3, 1]).astype(np.float32)
    scale_translate[3, 0] = 1.0                                  # scale(1,2,3) * translate(1,0,0)
    t0.set_matrix(scale_translate)
    tbt = TransformBlendTable()
    b0 = tbt.add_blend(TransformBlend(t0, 1.0))
    tbt.set_rows(SparseArray.range(1, 3))                        # rows 1..3 only
    vd.set_transform_blend_table(tbt)
    sl = UserVertexSlider("smile")
    st = SliderTable()
    st.add_slider(sl, SparseArray.range(2, 2))                   # rows 2, 3
    vd.set_slider_table(SliderTable.register_table(st))
    vd.set_column("vertex", [[i, 1, 1] for i in range(4)])
    vd.set_column("normal", [[0, 0, 1]] * 4)
    vd.set_column("transform_blend", [[b0]] * 4)
    vd.set_column(dname, [[0, 10, 0]] * 4)
    sl.set_slider(0.5)
    out = vd.animate_vertices(True)
    assert out is not vd and out.get_format().get_num_arrays() == 1
    v, n = out.get_column("vertex"), out.get_column("normal")
    # the reference's own results for this probe
    assert np.array_equal(v, [[0, 1, 1], [2, 2, 3], [3, 12, 3], [4, 12, 3]])
    assert np.allclose(n[1:], [[0, 0, 1]] * 3) and np.array_equal(n[0], [0, 0, 1])
    assert vd.animate_vertices(True) is out                      # nothing modified: cached
    sl.set_slider(1.0)
    out2 = vd.animate_vertices(True)
    assert out2 is out                                           # same object, mutated in place
    assert np.array_equal(out2.get_column("vertex")[3], [4, 22, 3])


# ---- code paths the BASELINE shapes do not reach ---------------------------------------------------

def _check_against_oracle(gpu_ctx, mesh, pal, sl=None, what=""):
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        assert np.array_equal(grp.read_selection(), expected_selection(mesh))
        expect, eb, _ = om.animate(pal, sl, want_blends=True)
        got = grp.animate_vertices(pal, sl)
        max_abs, nb = compare_outputs(mesh, got, expect, what)
        print(f"{what}: max-abs error {max_abs:.3e}, differing bytes {nb}")
        return nb
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.parametrize("k", [1, 2, 3, 6, 9])
def test_blends_with_other_than_four_weights(gpu_ctx, k):
    """TransformBlend is not limited to 4 entries on the CPU path: 1..3 use the
    fixed entry block partially, > 4 take the CSR fallback of the strip kernel."""
    sm = synth.make_mesh(5_000, 24, 24 if k == 1 else 400, seed=40 + k, weights_per_blend=k, index_order="runs")
    pal = synth.make_palettes(24, 1, 50 + k, jitter=0.2)[0]
    nb = _check_against_oracle(gpu_ctx, sm.flat, pal, what=f"{k}-weight blends")
    # rigid blends of scaled joints take the decompose_matrix branch (libm sinf/cosf/atan2f: last-ulp
    # differences, bounded by REL_TOL inside the check); every other case must match byte for byte
    assert nb == 0 or k == 1


def test_non_affine_palette_takes_the_general_inverse(gpu_ctx):
    """Column 3 != (0,0,0,1): LMatrix4f::invert_from falls to the LU branch (lmatrix4_src.I:1225-1246)."""
    sm = synth.make_mesh(3_000, 16, 300, seed=61, index_order="runs")
    pal = synth.make_palettes(16, 1, 62, scale=(1.0, 1.3, 0.7))[0].reshape(16, 4, 4).copy()
    pal[:, 0, 3] = 0.01
    pal[:, 2, 3] = -0.02
    pal[:, 3, 3] = 1.05
    _check_against_oracle(gpu_ctx, sm.flat, pal.reshape(16, 16), what="non-affine palette")


@pytest.mark.parametrize("layout", ["packed", "texcoord", "tbn", "align16"])
def test_normals_of_extreme_length_normalise_like_the_cpu(gpu_ctx, layout):
    """LVecBase3f::normalize over the whole float range (lvecBase3_src.I:347-361): the kernel's short form
    (one range test, then the in-range sqrt / reciprocal sequences) and its fallback must both reproduce the
    CPU's correctly rounded sqrtf and 1/x -- zero, subnormal, tiny, huge, infinite and NaN lengths included.
    Non-uniform joint scales select the normalising branch; every byte has to match the oracle."""
    kw = dict(texcoord=layout == "texcoord", tbn=layout == "tbn", align16=layout == "align16")
    sm = synth.make_mesh(4_096, 12, 200, seed=71, index_order="runs", **kw)
    mesh = sm.flat
    ncol = [c for c in mesh.skin_columns if c.contents == F.C_normal][0]
    arr = mesh.arrays[ncol.array]
    nrm = np.ascontiguousarray(arr[:, ncol.start:ncol.start + 12]).view(np.float32).reshape(-1, 3).copy()
    rng = np.random.default_rng(72)
    mags = np.array([0.0, 1e-45, 1e-40, 1e-38, 3e-31, 1e-25, 1e-19, 9.9e-16, 1e-10, 1.0, 1e10, 3e15, 1e18, 1e19, 3e19,
                     1e25, 1e38, np.inf, np.nan], np.float32)
    with np.errstate(all="ignore"):
        nrm = (nrm * mags[rng.integers(0, mags.size, size=(mesh.num_rows, 1))]).astype(np.float32)
    nrm[::97] = 0.0
    nrm[5::131, 1] = -0.0
    arr[:, ncol.start:ncol.start + 12] = nrm.view(np.uint8).reshape(-1, 12)
    pal = synth.make_palettes(12, 1, 73, scale=(1.0, 1.6, 0.55))[0]
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        expect = om.animate(pal)[0]
        got = grp.animate_vertices(pal)
        for g, e in zip(got, expect):
            # all columns of these layouts are float32; NaN payloads are the one thing x86 and the GPU encode differently
            g = np.ascontiguousarray(g).reshape(mesh.num_rows, -1).view(np.uint32)
            e = np.ascontiguousarray(e).reshape(mesh.num_rows, -1).view(np.uint32)
            both_nan = np.isnan(g.view(np.float32)) & np.isnan(e.view(np.float32))
            bad = np.flatnonzero(((g != e) & ~both_nan).any(axis=1))
            assert bad.size == 0, f"{layout}: {bad.size} rows differ, first {bad[:5]}"
            assert np.isfinite(e.view(np.float32)).mean() > 0.8
    finally:
        grp.close()
        dmesh.close()


def test_sliders_without_blend_table(gpu_ctx):
    """A vertex data with a SliderTable only: morphs apply, nothing is skinned (geomVertexData.cxx:947-948)."""
    sm = synth.make_mesh(2_000, 4, 8, seed=63, num_sliders=5, slider_band=0.4, morph_columns=("vertex", "normal"))
    mesh = sm.flat
    mesh = F.FlatMesh(**{**mesh.__dict__})
    mesh.blend_column = None
    mesh.skin_columns = []
    mesh.num_transforms = 0
    mesh.blend_offsets = np.zeros(1, np.int32)
    mesh.blend_transform = np.zeros(0, np.int32)
    mesh.blend_weight = np.zeros(0, np.float32)
    mesh.row_ranges = np.zeros((0, 2), np.int32)
    sl = synth.make_sliders(5, 1, 64)[0]
    assert _check_against_oracle(gpu_ctx, mesh, np.zeros((0, 16), np.float32), sl, what="sliders only") == 0


def test_singular_and_zero_matrices_do_not_poison_neighbours(gpu_ctx):
    """A zero joint matrix makes its blends singular; rows using other blends must still be exact and
    nothing may read out of bounds (the reference's own result for the singular rows is undefined)."""
    sm = synth.make_mesh(2_000, 8, 8, seed=65, weights_per_blend=1, index_order="sorted")
    mesh = sm.flat
    pal = synth.make_palettes(8, 1, 66, scale=(1.0, 1.4, 0.6))[0].copy()
    pal[3] = 0.0
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        got = grp.animate_vertices(pal)[0]
        expect = om.animate(pal)[0][0]
        ok_rows = mesh.blend_indices() != 3
        assert np.array_equal(got[ok_rows], expect[ok_rows])
        assert np.array_equal(got[~ok_rows][:, :12], expect[~ok_rows][:, :12])   # positions are defined (all zero)
    finally:
        grp.close()
        dmesh.close()


def test_wide_path_matches_strip_path(gpu_ctx, monkeypatch):
    """The two-kernel fallback (P3CS_PATH=wide) and the fused strip kernel agree byte for byte."""
    sm = synth.make_mesh(9_000, 48, 900, seed=67, tbn=True, texcoord=True, index_order="runs", num_sliders=3, slider_band=0.2)
    mesh = sm.flat
    pal = synth.make_palettes(48, 1, 68)[0]
    sl = synth.make_sliders(3, 1, 69)[0]
    outs = []
    for path in ("strip", "wide"):
        monkeypatch.setenv("P3CS_PATH", path)
        dmesh = _capi.Mesh(gpu_ctx, mesh)
        grp = _capi.Group(dmesh, 1)
        outs.append(grp.animate_vertices(pal, sl)[0])
        grp.close()
        dmesh.close()
    assert np.array_equal(outs[0], outs[1])


def test_joint_hierarchy_on_gpu_panda(gpu_ctx):
    """f1: p3cs_group_pose evaluates panda-walk4's 42-joint hierarchy on the GPU; the palette agrees with the
    reference's CharacterJoint::_skinning_matrix (1e-5 of the matrix scale: compose_matrix goes through
    sincosf), and pose -> animate reproduces the reference's animated vertices within the same bound."""
    from panda3d_b200.flat import skeleton_from_case
    case = load_case([p for p in golden_paths() if p.endswith("panda_walk4.p3cs")][0])
    mesh, skel = mesh_from_case(case), skeleton_from_case(case)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    frames = case["anim_frames"].astype(np.int32)
    grp = _capi.Group(dmesh, len(frames))
    try:
        grp.pose(dskel, frames)
        pal = grp.read_palettes()
        ref = case["palette"]
        scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
        assert (np.abs(pal - ref) / scale).max() <= REL_TOL
        print("pose: max-abs palette error", float(np.abs(pal - ref).max()))
        grp.animate()
        out = grp.read_output(0)
        for fi in range(len(frames)):
            compare_outputs(mesh, [out[fi]], [case["expect0"][fi]], f"pose+animate frame {frames[fi]}")
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


def test_joint_hierarchy_crowd_vs_oracle(gpu_ctx):
    """f1 at crowd scale: 300 instances x 128 joints, each instance on its own frame, against oracle.pose."""
    sm = synth.make_mesh(2_000, 128, 400, seed=71, index_order="runs")
    skel = synth.make_skeleton(128, 48, seed=72)
    n = 300
    frames = (np.arange(n) * 7 % 48).astype(np.int32)
    dmesh = _capi.Mesh(gpu_ctx, sm.flat)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    grp = _capi.Group(dmesh, n)
    try:
        grp.pose(dskel, frames)
        pal = grp.read_palettes()
        for i in (0, 1, 57, n - 1):
            ref = oracle.pose(skel, int(frames[i]))[skel.palette_joint]
            assert (np.abs(pal[i] - ref) / np.abs(ref).max()).max() <= REL_TOL
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


def _expected_bounds(out_array, col):
    """GeomPrimitive::calc_tight_bounds (gobj/geomPrimitive.cxx:1646-1700, non-indexed, no matrix) over the animated
    vertex column: component-wise min / max and max of length_squared() = (x*x + y*y) + z*z in float32."""
    v = np.ascontiguousarray(out_array[:, col.start:col.start + 12]).view(np.float32).reshape(-1, 3)
    with np.errstate(invalid="ignore"):
        sq = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
        return np.concatenate([np.nanmin(v, axis=0), np.nanmax(v, axis=0), [np.nanmax(sq)], [1.0]]).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["strip", "wide"])
def test_animated_tight_bounds(gpu_ctx, monkeypatch, path):
    """f4: the bounds the kernel reduces while skinning equal a CPU walk over its own animated vertices exactly
    (min / max are exact), for a crowd (one CTA per instance), a single mesh (several CTAs, atomics), a
    morph-only mesh with an explicit vertex column, and rows holding NaN."""
    monkeypatch.setenv("P3CS_PATH", path)
    # (1) crowd: 300 instances, generic layout with morphs
    sm = synth.make_mesh(3_000, 24, 400, seed=81, tbn=True, index_order="runs", num_sliders=3, slider_band=0.3)
    mesh = sm.flat
    n = 300
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        grp.set_palettes(synth.make_palettes(24, n, 82).reshape(n, -1))
        grp.set_sliders(synth.make_sliders(3, n, 83))
        grp.enable_bounds()
        grp.animate()
        got = grp.read_bounds()
        out = grp.read_output(0)
        for i in (0, 1, 17, n - 1):
            assert np.array_equal(got[i], _expected_bounds(out[i], mesh.skin_columns[0])), f"instance {i}"
        # against the oracle's vertices: within the float contract
        expect, _, _ = oracle.animate(mesh, grp.read_palettes(0, 1)[0], synth.make_sliders(3, n, 83)[0])
        eb = _expected_bounds(expect[0], mesh.skin_columns[0])
        assert np.allclose(got[0], eb, rtol=REL_TOL, atol=1e-6)
        # a second animate with other palettes replaces (does not accumulate into) the bounds
        grp.set_palettes(synth.make_palettes(24, n, 84, scale=0.25).reshape(n, -1))
        grp.animate()
        got2 = grp.read_bounds()
        out2 = grp.read_output(0)
        assert np.array_equal(got2[5], _expected_bounds(out2[5], mesh.skin_columns[0]))
        grp.enable_bounds(enable=False)
        grp.animate()
        with pytest.raises(_capi.P3csError):
            grp.read_bounds()
    finally:
        grp.close()
        dmesh.close()
    # (2) one big packed mesh (specialised mode, many CTAs for the one instance), with NaN rows
    sm = synth.make_mesh(70_000, 32, 3000, seed=85, index_order="runs")
    mesh = sm.flat
    f = mesh.arrays[0].view(np.float32)
    f[123, 0] = np.nan
    f[50_000, 0:3] = np.nan
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        grp.set_palettes(synth.make_palettes(32, 1, 86).reshape(1, -1))
        grp.enable_bounds()
        grp.animate()
        assert np.array_equal(grp.read_bounds()[0], _expected_bounds(grp.read_output(0)[0], mesh.skin_columns[0]))
    finally:
        grp.close()
        dmesh.close()
    # (3) sliders only (nothing skinned): the vertex column is named explicitly
    sm = synth.make_mesh(2_000, 4, 8, seed=87, num_sliders=5, slider_band=0.4)
    mesh = F.FlatMesh(**{**sm.flat.__dict__})
    vertex = mesh.skin_columns[0]
    mesh.skin_columns, mesh.blend_column = [], None
    mesh.blend_offsets = np.zeros(1, np.int32)
    mesh.blend_transform, mesh.blend_weight = np.zeros(0, np.int32), np.zeros(0, np.float32)
    mesh.num_transforms = 0
    mesh.row_ranges = np.zeros((0, 2), np.int32)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 2)
    try:
        grp.set_sliders(synth.make_sliders(5, 2, 88))
        grp.enable_bounds(vertex)
        grp.animate()
        out = grp.read_output(0)
        got = grp.read_bounds()
        for i in range(2):
            assert np.array_equal(got[i], _expected_bounds(out[i], vertex))
        with pytest.raises(_capi.P3csError):   # a float64 / 4-component column is refused, not misread
            grp.enable_bounds(F.Column(0, 0, 4, F.NT_float32, F.C_point))
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["strip", "wide"])
@pytest.mark.parametrize("layout", ["packed", "tbn_morph"])
def test_tight_bounds_over_referenced_rows_only(gpu_ctx, monkeypatch, path, layout):
    """f4, indexed primitives: GeomPrimitive::calc_tight_bounds walks the vertices the primitive references
    (gobj/geomPrimitive.cxx:1646-1830), so rows no primitive indexes must not widen the bounds --
    p3cs_group_set_bounds_rows; exact against a CPU walk over the referenced rows of the kernel's own output."""
    monkeypatch.setenv("P3CS_PATH", path)
    kw = dict(tbn=True, num_sliders=3, slider_band=0.3) if layout == "tbn_morph" else {}
    sm = synth.make_mesh(5_003, 24, 400, seed=91, index_order="runs", **kw)
    mesh = sm.flat
    rng = np.random.default_rng(92)
    referenced = rng.uniform(size=mesh.num_rows) < 0.4
    referenced[4990:] = False
    # the extreme vertices are exactly the unreferenced ones
    f = mesh.arrays[0][:, :12].copy().view(np.float32).reshape(-1, 3)
    f[~referenced] *= 50.0
    mesh.arrays[0][:, :12] = f.view(np.uint8).reshape(-1, 12)
    n = 40
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        grp.set_palettes(synth.make_palettes(24, n, 93).reshape(n, -1))
        if mesh.num_sliders:
            grp.set_sliders(synth.make_sliders(3, n, 94))
        grp.enable_bounds()
        grp.set_bounds_rows(referenced)
        grp.animate()
        got, out = grp.read_bounds(), grp.read_output(0)
        for i in (0, 7, n - 1):
            assert np.array_equal(got[i], _expected_bounds(out[i][referenced], mesh.skin_columns[0])), f"instance {i}"
            assert not np.array_equal(got[i], _expected_bounds(out[i], mesh.skin_columns[0]))
        grp.set_bounds_rows(None)
        grp.animate()
        assert np.array_equal(grp.read_bounds()[3], _expected_bounds(grp.read_output(0)[3], mesh.skin_columns[0]))
        grp.set_bounds_rows(np.zeros(mesh.num_rows, bool))   # nothing referenced: found_any stays 0
        grp.animate()
        assert grp.read_bounds()[0][7] == 0.0
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("plan", ["1,3,256", "1,2,256", "2,2,256", "2,2,168", "1,2,136", "2,2,64", "1,3,8"])
@pytest.mark.parametrize("layout", ["packed", "texcoord", "tbn_uv", "align16"])
def test_every_shared_memory_plan_gives_the_same_bytes(gpu_ctx, monkeypatch, plan, layout):
    """The strip kernel's shared-memory plan (rows per thread, tile buffers, record-group size; p3cs_mesh.cu) only
    changes how the work is cut: forced through the P3CS_STRIP_PLAN tuning aid, every plan must give the very bytes of
    the default plan -- 256- and 512-row tiles, two and three buffers, record groups from 256 blends down to 8 -- and
    those agree with the oracle (one blend of this seed is nearly uniformly scaled: libm branch, within the contract)."""
    kw = {"packed": {}, "texcoord": dict(texcoord=True), "tbn_uv": dict(tbn=True, texcoord=True), "align16": dict(align16=True)}[layout]
    sm = synth.make_mesh(6_011, 40, 900, seed=101, index_order="runs", num_sliders=2, slider_band=0.2, **kw)
    mesh = sm.flat
    pal = synth.make_palettes(40, 3, 102, scale=(1.0, 1.2, 0.8))
    sl = synth.make_sliders(2, 3, 103)

    def run():
        dmesh = _capi.Mesh(gpu_ctx, mesh)
        grp = _capi.Group(dmesh, 3)
        try:
            assert np.array_equal(grp.read_selection(), expected_selection(mesh))
            grp.set_palettes(pal.reshape(3, -1))
            grp.set_sliders(sl)
            grp.animate()
            return grp.read_output(0).copy()
        finally:
            grp.close()
            dmesh.close()

    monkeypatch.delenv("P3CS_STRIP_PLAN", raising=False)
    reference = run()
    monkeypatch.setenv("P3CS_STRIP_PLAN", plan)
    forced = run()
    assert np.array_equal(forced, reference), f"plan {plan} layout {layout}"
    om = oracle.OracleMesh(mesh)
    for i in range(3):
        compare_outputs(mesh, [forced[i]], [om.animate(pal[i], sl[i])[0][0]], f"plan {plan} layout {layout} instance {i}")


@pytest.mark.gpu
def test_animate_groups_mixed_crowd(gpu_ctx):
    """p3cs_animate_groups (independent groups forked over side streams) == animating the groups one by one,
    byte for byte, on a C5-style mixed crowd of 6 character types (more groups than side streams)."""
    shapes = [(900, 8, 100, 7), (2_000, 16, 300, 5), (5_000, 32, 700, 3), (300, 4, 20, 11), (1_200, 12, 200, 2), (4_099, 24, 512, 1)]
    meshes, groups, expect = [], [], []
    try:
        for i, (rows, t, b, n) in enumerate(shapes):
            sm = synth.make_mesh(rows, t, b, seed=900 + i, tbn=(i % 2 == 0), index_order="runs",
                                 num_sliders=2 if i == 2 else 0, slider_band=0.3)
            dm = _capi.Mesh(gpu_ctx, sm.flat)
            g = _capi.Group(dm, n)
            g.set_palettes(synth.make_palettes(t, n, 950 + i).reshape(n, -1))
            if sm.flat.num_sliders:
                g.set_sliders(synth.make_sliders(2, n, 960 + i))
            meshes.append(dm)
            groups.append(g)
        for g in groups:
            g.animate()
            expect.append(g.read_output(0).copy())
        for _ in range(3):   # repeated frames: the join must order every side stream before the read-back
            for g in groups:  # scribble over the outputs so a skipped group would show
                g.set_palettes(np.zeros((g.n, g.mesh.flat.num_transforms * 16), np.float32))
                g.animate()
            for i, g in enumerate(groups):
                g.set_palettes(synth.make_palettes(shapes[i][1], shapes[i][3], 950 + i).reshape(shapes[i][3], -1))
            _capi.animate_groups(groups)
            for g, e in zip(groups, expect):
                assert np.array_equal(g.read_output(0), e)
        assert all(g.timings().launches >= 1 for g in groups)
    finally:
        for g in groups:
            g.close()
        for dm in meshes:
            dm.close()


@pytest.mark.gpu
def test_registered_host_outputs_are_written_directly(gpu_ctx):
    """p3cs_animate_vertices into page-locked, mapped host arrays (p3cs_host_register): same bytes as the copying
    path, for the packed specialised mode, a several-array generic mesh and a mesh whose size is not a multiple of
    16 bytes (which must quietly take the copying path)."""
    cases = [synth.make_mesh(1_338, 26, 197, seed=91, index_order="runs"),
             synth.make_mesh(4_000, 16, 300, seed=92, tbn=True, texcoord=True, num_sliders=3, slider_band=0.3),
             synth.make_mesh(1_001, 8, 50, seed=93, tbn=True)]   # 1001 x 48 bytes: multiple of 16; see below
    for k, sm in enumerate(cases):
        mesh = sm.flat
        pal = synth.make_palettes(mesh.num_transforms, 2, 94 + k)
        sl = synth.make_sliders(mesh.num_sliders, 2, 97) if mesh.num_sliders else [None, None]
        dmesh = _capi.Mesh(gpu_ctx, mesh)
        grp = _capi.Group(dmesh, 3)
        outs = [np.full((mesh.num_rows, s), 0xEE, np.uint8) for s in dmesh.out_strides]
        try:
            expect = [grp.animate_vertices(pal[f], sl[f], instance=1) for f in range(2)]
            for o in outs:
                gpu_ctx.host_register(o)
            for f in range(2):
                got = grp.animate_vertices(pal[f], sl[f], instance=1, outs=outs)
                for g, e in zip(got, expect[f]):
                    assert np.array_equal(g, e), f"case {k} frame {f}"
        finally:
            for o in outs:
                gpu_ctx.host_unregister(o)
            grp.close()
            dmesh.close()
    # rows * stride not a multiple of 16 (1001 x 24): the registered buffer is still filled, through the copy
    sm = synth.make_mesh(1_001, 8, 50, seed=95)
    mesh = sm.flat
    assert (mesh.num_rows * mesh.strides[0]) % 16 != 0
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    guard = np.full(mesh.num_rows * mesh.strides[0] + 64, 0xAB, np.uint8)
    out = guard[:mesh.num_rows * mesh.strides[0]].reshape(mesh.num_rows, -1)
    try:
        pal = synth.make_palettes(8, 1, 96)[0]
        expect = grp.animate_vertices(pal)[0]
        gpu_ctx.host_register(guard)
        got = grp.animate_vertices(pal, outs=[out])[0]
        assert np.array_equal(got, expect)
        assert (guard[mesh.num_rows * mesh.strides[0]:] == 0xAB).all(), "bytes past the array were written"
    finally:
        gpu_ctx.host_unregister(guard)
        grp.close()
        dmesh.close()


@pytest.mark.gpu
def test_one_call_entry_point_keeps_the_group_up_to_date(gpu_ctx):
    """p3cs_animate_vertices back to back on different instances of a group: same bytes as the crowd path
    (set_palettes / set_sliders / animate), and the call's palette and sliders stay behind as the instance's device
    copies, so a later p3cs_group_animate of the whole group sees, per instance, the inputs of the last call that named it."""
    sm = synth.make_mesh(1_338, 26, 197, seed=191, texcoord=True, index_order="runs", num_sliders=5, slider_band=0.3)
    mesh = sm.flat
    pal = synth.make_palettes(mesh.num_transforms, 7, 192)
    sl = synth.make_sliders(mesh.num_sliders, 7, 193)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 4)
    ref = _capi.Group(dmesh, 1)
    try:
        for f in range(7):
            inst = f % 4
            got = grp.animate_vertices(pal[f], sl[f], instance=inst)
            ref.set_palettes(pal[f].reshape(1, -1))
            ref.set_sliders(sl[f].reshape(1, -1))
            ref.animate()
            assert np.array_equal(got[0], ref.read_output(0)[0]), f"call {f}"
            assert np.array_equal(grp.read_palettes()[inst].ravel(), pal[f].ravel()), "device palette not refreshed"
        grp.animate()   # every instance from its device copies: instance i holds the inputs of the last call that named it
        out = grp.read_output(0)
        for inst, f in ((0, 4), (1, 5), (2, 6), (3, 3)):
            ref.set_palettes(pal[f].reshape(1, -1))
            ref.set_sliders(sl[f].reshape(1, -1))
            ref.animate()
            assert np.array_equal(out[inst], ref.read_output(0)[0]), f"instance {inst}"
    finally:
        ref.close()
        grp.close()
        dmesh.close()


@pytest.mark.gpu
def test_edge_sizes(gpu_ctx):
    """Sizes at the edges of the launch geometry: an empty vertex data, a single instance of a mesh smaller than a
    warp, a palette too large for shared memory (automatic two-kernel path), more instances than one grid
    dimension holds (70 000 > 65 535), and TransformBlendTable::get_rows() empty."""
    # (a) zero rows
    sm = synth.make_mesh(8, 2, 2, seed=1)
    empty = F.FlatMesh(**{**sm.flat.__dict__})
    empty.num_rows = 0
    empty.arrays = [a[:0] for a in sm.flat.arrays]
    empty.row_ranges = np.zeros((0, 2), np.int32)
    dmesh = _capi.Mesh(gpu_ctx, empty)
    grp = _capi.Group(dmesh, 2)
    grp.set_palettes(synth.make_palettes(2, 2, 3).reshape(2, -1))
    grp.animate()
    assert grp.read_output(0).shape == (2, 0, empty.strides[0])
    grp.close(); dmesh.close()
    # (b) 5 rows
    sm = synth.make_mesh(5, 3, 4, seed=2)
    assert _check_against_oracle(gpu_ctx, sm.flat, synth.make_palettes(3, 1, 4)[0], None, what="5 rows") >= 0
    # (c) 4000 joints: 256 KB of palette cannot be staged in shared memory
    sm = synth.make_mesh(6_000, 4000, 5000, seed=3, index_order="runs")
    assert _check_against_oracle(gpu_ctx, sm.flat, synth.make_palettes(4000, 1, 5)[0], None, what="4000 joints") >= 0
    # (d) 70 000 instances of a 40-row mesh: two launches (grid.y <= 65535); spot-check both halves
    sm = synth.make_mesh(40, 6, 10, seed=6)
    mesh = sm.flat
    n = 70_000
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        base = synth.make_palettes(6, 50, 7).reshape(50, -1)
        grp.set_palettes(np.tile(base, (n // 50, 1)))
        grp.enable_bounds()
        grp.animate()
        om = oracle.OracleMesh(mesh)
        for i in (0, 49, 65_534, 65_535, 65_536, n - 1):
            expect, _, _ = om.animate(base[i % 50])
            got = grp.read_output(0, i, 1)[0]
            compare_outputs(mesh, [got], expect, f"instance {i}")
            assert np.array_equal(grp.read_bounds(i, 1)[0], _expected_bounds(got, mesh.skin_columns[0]))
        assert grp.timings().launches >= 2
    finally:
        grp.close()
        dmesh.close()
    # (e) a blend table whose rows set is empty: plain copy
    sm = synth.make_mesh(300, 4, 8, seed=8)
    mesh = F.FlatMesh(**{**sm.flat.__dict__})
    mesh.row_ranges = np.zeros((0, 2), np.int32)
    assert _check_against_oracle(gpu_ctx, mesh, synth.make_palettes(4, 1, 9)[0], None, what="empty rows") == 0


@pytest.mark.gpu
def test_shareable_output_reaches_another_process(gpu_ctx):
    """f2: the group animates straight into memory exported as a POSIX fd; a second process (the "renderer",
    tests/renderer_import_helper.py) imports the fd and finds the same bytes the oracle-checked copy path returns."""
    import hashlib
    import subprocess
    import sys
    sm = synth.make_mesh(5_000, 24, 600, seed=111, texcoord=True, index_order="runs")
    mesh = sm.flat
    n = 3
    pal = synth.make_palettes(24, n, 112).reshape(n, -1)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    plain = _capi.Group(dmesh, n)
    pitch = plain.output_pitch(0)
    ptr, fd, allocated = gpu_ctx.shareable_alloc(n * pitch)
    shared = _capi.Group(dmesh, n, outputs_dev=[ptr])
    try:
        assert allocated >= n * pitch and fd >= 0
        for g in (plain, shared):
            g.set_palettes(pal)
            g.animate()
        expect = plain.read_output(0)
        got = shared.read_output(0)
        assert np.array_equal(got, expect)
        oracle_out, _, _ = oracle.animate(mesh, pal[1].reshape(-1, 16))
        compare_outputs(mesh, [got[1]], oracle_out, "shareable output, instance 1")
        gpu_ctx.sync()
        mine = hashlib.sha256(bytes(gpu_ctx.device_read(ptr, n * pitch))).hexdigest()
        helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "renderer_import_helper.py")
        res = subprocess.run([sys.executable, helper, str(fd), str(allocated), str(n * pitch)], pass_fds=[fd],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        assert res.stdout.strip().splitlines()[-1] == mine, "the importing process sees different bytes"
    finally:
        shared.close()
        plain.close()
        dmesh.close()
        gpu_ctx.shareable_free(ptr)
        os.close(fd)


@pytest.mark.gpu
def test_animate_host_pipelined_crowd(gpu_ctx):
    """p3cs_group_animate_host (chunks pipelined over upload / kernel / download streams) returns, for every
    instance, the bytes of the one-instance call -- with chunk sizes that do and do not divide the crowd."""
    sm = synth.make_mesh(2_500, 16, 300, seed=121, tbn=True, num_sliders=3, slider_band=0.3, index_order="runs")
    mesh = sm.flat
    n = 37
    pal = synth.make_palettes(16, n, 122)
    sl = synth.make_sliders(3, n, 123)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        single = [grp.animate_vertices(pal[i], sl[i], instance=i)[0] for i in (0, 17, n - 1)]
        for chunk in (0, 1, 5, 37, 100):
            outs = grp.animate_host(pal.reshape(n, -1), sl, chunk=chunk)
            for k, i in enumerate((0, 17, n - 1)):
                assert np.array_equal(outs[0][i], single[k]), f"chunk {chunk}, instance {i}"
        expect, _, _ = oracle.animate(mesh, pal[17], sl[17])
        compare_outputs(mesh, [outs[0][17]], expect, "animate_host instance 17")
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["strip", "wide"])
def test_animate_only_the_listed_instances(gpu_ctx, monkeypatch, path):
    """p3cs_group_animate_instances recomputes exactly the listed instances (any order) and leaves the arrays and
    bounds of the others as they were."""
    monkeypatch.setenv("P3CS_PATH", path)
    sm = synth.make_mesh(1_500, 12, 200, seed=131, texcoord=True, index_order="runs")
    mesh = sm.flat
    n = 40
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        pal0 = synth.make_palettes(12, n, 132).reshape(n, -1)
        pal1 = synth.make_palettes(12, n, 133).reshape(n, -1)
        grp.set_palettes(pal0)
        grp.enable_bounds()
        grp.animate()
        before, bounds_before = grp.read_output(0).copy(), grp.read_bounds().copy()
        grp.set_palettes(pal1)           # every palette changes on the device ...
        listed = [31, 2, 17, 39, 0]
        grp.animate_instances(listed)    # ... but only these are re-animated
        after, bounds_after = grp.read_output(0), grp.read_bounds()
        grp.animate()
        full = grp.read_output(0)
        for i in range(n):
            if i in listed:
                assert np.array_equal(after[i], full[i]), f"listed instance {i} was not recomputed"
                assert np.array_equal(bounds_after[i], _expected_bounds(after[i], mesh.skin_columns[0]))
            else:
                assert np.array_equal(after[i], before[i]), f"instance {i} changed although it was not listed"
                assert np.array_equal(bounds_after[i], bounds_before[i])
        with pytest.raises(_capi.P3csError):
            grp.animate_instances([0, n])
    finally:
        grp.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["panda_walk4_frame_blend", "panda_walk4_frame_blend_linear", "panda_walk4_frame_blend_componentwise_quat"])
def test_joint_hierarchy_with_frame_blending(gpu_ctx, fixture):
    """f1 with PartBundle::set_frame_blend_flag(true): p3cs_group_pose_blend reproduces the reference's skinning
    matrices between frames (BT_normalized_linear, the PRC default, and BT_linear) within the float contract
    (compose / decompose go through libm), and pose -> animate the reference's animated vertices."""
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    mesh, skel = mesh_from_case(case), skeleton_from_case(case)
    bt = int(case["anim_blend_type"][0])
    frames, nxt, fracs = case["anim_frames"], case["anim_next_frames"], case["anim_fracs"]
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    grp = _capi.Group(dmesh, len(frames))
    try:
        grp.pose_blend(dskel, frames, nxt, fracs, bt)
        pal = grp.read_palettes()
        ref = case["palette"]
        scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
        assert (np.abs(pal - ref) / scale).max() <= REL_TOL
        grp.animate()
        out = grp.read_output(0)
        for fi in range(len(frames)):
            compare_outputs(mesh, [out[fi]], [case["expect0"][fi]], f"{fixture} frame {frames[fi]}+{fracs[fi]}")
        # against the oracle on a random rig, both blend types
        sk = synth.make_skeleton(40, 16, seed=141)
        dsk = _capi.Skeleton(gpu_ctx, sk)
        sm = synth.make_mesh(500, 40, 60, seed=142)
        dm = _capi.Mesh(gpu_ctx, sm.flat)
        g2 = _capi.Group(dm, 6)
        f = np.arange(6, dtype=np.int32) * 3
        n = (f + 1) % 16
        fr = np.linspace(0.0, 0.9, 6).astype(np.float32)
        g2.pose_blend(dsk, f, n, fr, bt)
        got = g2.read_palettes()
        for i in range(6):
            want = oracle.pose(sk, int(f[i]), int(n[i]), float(fr[i]), bt)[sk.palette_joint]
            assert (np.abs(got[i] - want) / np.abs(want).max()).max() <= REL_TOL
        g2.close(); dm.close(); dsk.close()
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cs", [F.CS_zup_right, F.CS_yup_right, F.CS_zup_left, F.CS_yup_left])
@pytest.mark.parametrize("bt", [2, 3])
def test_componentwise_blends_in_every_coordinate_system(gpu_ctx, cs, bt):
    """BT_componentwise and BT_componentwise_quat (chan/movingPartMatrix.cxx:200-357) on a random rig in each of the
    four coordinate systems (LQuaternionf::set_hpr multiplies in the opposite order and inverts on the left-handed
    ones; set_scale_shear_mat moves and negates the shear terms): p3cs_group_pose_blend against the oracle."""
    sk = synth.make_skeleton(48, 12, seed=151 + cs, coordinate_system=cs)
    sm = synth.make_mesh(400, 48, 50, seed=152, coordinate_system=cs)
    dsk = _capi.Skeleton(gpu_ctx, sk)
    dm = _capi.Mesh(gpu_ctx, sm.flat)
    grp = _capi.Group(dm, 8)
    try:
        f = (np.arange(8, dtype=np.int32) * 5) % 12
        n = (f + 1) % 12
        fr = np.linspace(0.0, 0.95, 8).astype(np.float32)
        grp.pose_blend(dsk, f, n, fr, bt)
        got = grp.read_palettes()
        for i in range(8):
            want = oracle.pose(sk, int(f[i]), int(n[i]), float(fr[i]), bt)[sk.palette_joint]
            assert (np.abs(got[i] - want) / np.abs(want).max()).max() <= REL_TOL, f"cs {cs} blend type {bt} instance {i}"
    finally:
        grp.close()
        dm.close()
        dsk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["panda_walk4_forced", "panda_walk4_forced_frame_blend"])
def test_joint_hierarchy_with_forced_channels(gpu_ctx, fixture):
    """f1 with PartBundle::freeze_joint (two joints) and control_joint (a third, its node moving every frame):
    p3cs_group_set_forced_joints / p3cs_group_set_forced_values + the pose calls reproduce the reference's skinning
    matrices and animated vertices; releasing the joints gives the plain animation back."""
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    plain = load_case(os.path.join(GOLDEN_DIR, "panda_walk4.p3cs"))
    mesh, skel = mesh_from_case(case), skeleton_from_case(case)
    frames = case["anim_frames"]
    fb = "anim_next_frames" in case
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    grp = _capi.Group(dmesh, len(frames))
    try:
        grp.set_forced_joints(dskel, case["forced_joints"])
        grp.set_forced_values(case["forced_values"])
        if fb:
            grp.pose_blend(dskel, frames, case["anim_next_frames"], case["anim_fracs"], int(case["anim_blend_type"][0]))
        else:
            grp.pose(dskel, frames)
        pal = grp.read_palettes()
        ref = case["palette"]
        scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
        assert (np.abs(pal - ref) / scale).max() <= REL_TOL
        grp.animate()
        out = grp.read_output(0)
        for fi in range(len(frames)):
            compare_outputs(mesh, [out[fi]], [case["expect0"][fi]], f"{fixture} frame {frames[fi]}")
        with pytest.raises(_capi.P3csError):
            grp.set_forced_joints(dskel, [3, 3])
        with pytest.raises(_capi.P3csError):
            grp.set_forced_joints(dskel, [skel.num_joints])
        if not fb:   # release_joint: frames 0 / 10 / 37 of the plain fixture
            grp.set_forced_joints(dskel, [])
            grp.pose(dskel, frames)
            pal = grp.read_palettes()
            ref = plain["palette"][:3]
            assert list(plain["anim_frames"][:3]) == list(frames)
            assert (np.abs(pal - ref) / np.abs(ref).max(axis=(1, 2), keepdims=True)).max() <= REL_TOL
            with pytest.raises(_capi.P3csError):
                grp.set_forced_values(case["forced_values"])
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


@pytest.mark.gpu
def test_plain_c_host_animates_through_the_plugin_table(tmp_path):
    """examples/plugin_host.c --run: a C99 host (dlopen + p3cs_get_api, no Python, no torch) skins the survey's
    four-vertex probe: (2,4,1) under 0.25*T(1,2,3) + 0.75*S(2) -> (3.75, 7.5, 2.5)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "plugin_host"
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "plugin_host.c"),
                    "-ldl", "-o", str(exe)], check=True)
    res = subprocess.run([str(exe), _capi.LIB_PATH, "--run"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "vertex 0 -> (3.75, 7.5, 2.5)" in res.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["panda_walk4_two_controls_hold", "panda_walk4_two_controls_hold_linear",
                                     "panda_walk4_two_controls_hold_componentwise",
                                     "panda_walk4_two_controls_frame_blend", "panda_walk4_two_controls_frame_blend_linear",
                                     "panda_walk4_two_controls_frame_blend_componentwise",
                                     "panda_walk4_two_controls_hold_componentwise_quat",
                                     "panda_walk4_two_controls_frame_blend_componentwise_quat"])
def test_joint_hierarchy_with_two_blended_controls(gpu_ctx, fixture):
    """f1 with several AnimControls (PartBundle::set_anim_blend_flag / set_control_effect): p3cs_group_pose_controls
    against the reference's Character posed with two controls of different frames and effects -- the second control
    plays a second, separately uploaded copy of the animation (p3cs_skeleton_add_animation)."""
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    mesh, skel = mesh_from_case(case), skeleton_from_case(case)
    fb, bt = bool(case["anim_frame_blend"][0]), int(case["anim_blend_type"][0])
    n = case["ctrl_frames"].shape[0]
    states = np.zeros((n, 2, 5))
    states[:, :, 1] = case["ctrl_frames"]
    states[:, :, 2] = case["ctrl_next_frames"]
    states[:, :, 3] = case["ctrl_fracs"]
    states[:, :, 4] = case["ctrl_effects"]
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    states[:, 1, 0] = dskel.add_animation(skel)
    assert states[0, 1, 0] == 1
    grp = _capi.Group(dmesh, n)
    try:
        grp.pose_controls(dskel, states, fb, bt)
        pal = grp.read_palettes()
        ref = case["palette"]
        scale = np.abs(ref).max(axis=(1, 2), keepdims=True)
        assert (np.abs(pal - ref) / scale).max() <= REL_TOL
        grp.animate()
        out = grp.read_output(0)
        for i in range(n):
            compare_outputs(mesh, [out[i]], [case["expect0"][i]], f"{fixture} sample {i}")
        bad = states.copy()
        bad[0, 0, 0] = 7
        with pytest.raises(_capi.P3csError):
            grp.pose_controls(dskel, bad, fb, bt)
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["ripple", "skeleton/ripple_frame_blend", "skeleton/ripple_two_controls_hold",
                                     "skeleton/ripple_two_controls_frame_blend"])
def test_slider_channels_on_gpu_ripple(gpu_ctx, fixture):
    """models/ripple.egg (7 morph sliders driven by <S$Anim> scalar channels): the pose calls also evaluate
    CharacterSlider's values on the device -- bit for bit the reference's (table look-ups, products and one
    division, no libm) -- and pose -> animate reproduces the reference's animated vertices."""
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    from test_oracle_golden import _control_arrays
    case = load_case(os.path.join(GOLDEN_DIR, fixture + ".p3cs"))
    mesh, skel = mesh_from_case(case), skeleton_from_case(case)
    frames, nxt, fracs, effects, fb = _control_arrays(case)
    n, K = frames.shape
    bt = int(case["anim_blend_type"][0]) if "anim_blend_type" in case else 1
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    dskel = _capi.Skeleton(gpu_ctx, skel)
    grp = _capi.Group(dmesh, n)
    try:
        if K == 1 and not fb:
            grp.pose(dskel, frames[:, 0])
        elif K == 1:
            grp.pose_blend(dskel, frames[:, 0], nxt[:, 0], fracs[:, 0], bt)
        else:
            states = np.zeros((n, K, 5))
            states[:, :, 1], states[:, :, 2], states[:, :, 3], states[:, :, 4] = frames, nxt, fracs, effects
            states[:, 1, 0] = dskel.add_animation(skel)
            grp.pose_controls(dskel, states, fb, bt)
        gpu_ctx.sync()
        got = gpu_ctx.device_read(_capi.lib().p3cs_group_sliders_device(grp.handle), n * 7 * 4).view(np.float32).reshape(n, 7)
        assert np.array_equal(got.view(np.uint32), case["sliders"].view(np.uint32)), "slider values differ from CharacterSlider's"
        pal = grp.read_palettes()
        assert (np.abs(pal - case["palette"]) / np.abs(case["palette"]).max()).max() <= REL_TOL
        grp.animate()
        out = grp.read_output(0)
        for i in range(n):
            compare_outputs(mesh, [out[i]], [case["expect0"][i]], f"{fixture} sample {i}")
    finally:
        grp.close()
        dskel.close()
        dmesh.close()


RUN_KERNEL_CASES = {
    # layout -> make_mesh keywords; every case also runs with sparse rows, blends of 1..6 entries and a non-affine palette
    "packed": dict(),
    "texcoord": dict(texcoord=True),
    "tbn": dict(tbn=True),
    "tbn_uv": dict(tbn=True, texcoord=True),
    "packed_morph": dict(num_sliders=3, slider_band=0.4, morph_columns=("vertex", "normal")),
    "tbn_uv_morph": dict(tbn=True, texcoord=True, num_sliders=3, slider_band=0.4, morph_columns=("vertex", "tangent")),
}


@pytest.mark.gpu
@pytest.mark.parametrize("order", ["runs", "random", "fixed4", "sorted"])
@pytest.mark.parametrize("layout", sorted(RUN_KERNEL_CASES))
def test_run_kernel_and_strip_kernel_agree(gpu_ctx, monkeypatch, layout, order):
    """The two kernels of the row-layout modes (p3cs_run.cuh: one thread per run of rows; p3cs_strip.cuh: one thread per
    row) give the same bytes, and the oracle's, on crowds of every layout: runs / random / aligned runs / long runs of
    the blend index, rows outside get_rows(), blends of 1..6 entries, selection read-back, tight bounds, instance lists."""
    kw = dict(RUN_KERNEL_CASES[layout])
    rows = 4_571
    sm = synth.make_mesh(rows, 40, 700, seed=300 + len(layout), index_order=order, rows=[(3, 1500), (1507, 4000), (4100, rows)], **kw)
    mesh = sm.flat
    # blends of 1..6 entries: re-cut the CSR (entries stay in ascending transform order)
    rng = np.random.default_rng(7)
    offs, xf, w = [0], [], []
    for b in range(mesh.num_blends):
        k = int(rng.integers(1, 7))
        j = np.sort(rng.choice(mesh.num_transforms, size=k, replace=False))
        ww = rng.uniform(0.05, 1.0, k).astype(np.float32)
        ww = (ww / ww.sum(dtype=np.float32)).astype(np.float32)
        xf += list(j)
        w += list(ww)
        offs.append(len(xf))
    mesh.blend_offsets = np.asarray(offs, np.int32)
    mesh.blend_transform = np.asarray(xf, np.int32)
    mesh.blend_weight = np.asarray(w, np.float32)
    n = 9
    pal = synth.make_palettes(mesh.num_transforms, n, 41).reshape(n, mesh.num_transforms, 16).copy()
    pal[4, :, 15] = 1.0
    pal[4, 5, 3] = 0.25      # instance 4: a palette that is not affine (general 4x4 inverse branch)
    sl = synth.make_sliders(mesh.num_sliders, n, 42) if mesh.num_sliders else None
    om = oracle.OracleMesh(mesh)
    results = {}
    # "run2": the run kernel with two palette copies in shared memory (items carry palette-area indices the planner
    # chose, p3cs_runplan.h); "run": one copy; by default the library picks per mesh
    for path in ("run", "run2", "strip"):
        monkeypatch.setenv("P3CS_PATH", path.rstrip("2"))
        monkeypatch.setenv("P3CS_RUN_COPIES", "2" if path == "run2" else "1")
        dmesh = _capi.Mesh(gpu_ctx, mesh)
        grp = _capi.Group(dmesh, n)
        try:
            assert np.array_equal(grp.read_selection(), expected_selection(mesh)), f"{path}: selection differs"
            grp.set_palettes(pal.reshape(n, -1))
            if sl is not None:
                grp.set_sliders(sl)
            grp.enable_bounds()
            grp.animate()
            out = grp.read_output(0).copy()
            bounds = grp.read_bounds().copy()
            grp.enable_bounds(enable=False)
            # scribble, then recompute only three instances
            grp.set_palettes(np.zeros((n, mesh.num_transforms * 16), np.float32))
            grp.animate()
            grp.set_palettes(pal.reshape(n, -1))
            grp.animate_instances([7, 0, 4])
            listed = grp.read_output(0).copy()
            results[path] = (out, bounds, listed)
        finally:
            grp.close()
            dmesh.close()
    out_r, bounds_r, listed_r = results["run"]
    out_s, bounds_s, listed_s = results["strip"]
    assert np.array_equal(out_r, results["run2"][0]), "two palette copies change the run kernel's bytes"
    assert np.array_equal(bounds_r.view(np.uint32), results["run2"][1].view(np.uint32)) and np.array_equal(listed_r, results["run2"][2])
    assert np.array_equal(out_r, out_s), "run kernel and strip kernel differ"
    assert np.array_equal(bounds_r.view(np.uint32), bounds_s.view(np.uint32)), "tight bounds differ between the kernels"
    assert np.array_equal(listed_r, listed_s)
    for i in (0, 4, 7):
        assert np.array_equal(listed_r[i], out_r[i])
    for i in range(n):
        expect, _, _ = om.animate(pal[i].reshape(-1, 16), sl[i] if sl is not None else None)
        compare_outputs(mesh, [out_r[i]], expect, f"{layout}/{order} instance {i}")
        v = out_r[i].reshape(rows, -1)[:, :12].copy().view(np.float32)
        assert np.array_equal(bounds_r[i][:3], v.min(axis=0)) and np.array_equal(bounds_r[i][3:6], v.max(axis=0))


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["packed", "tbn_uv", "align16"])
def test_per_instance_output_pointers(gpu_ctx, layout):
    """p3cs_group_set_output_pointers: one launch animates a crowd straight into per-instance destinations -- separately
    page-locked host arrays (the `_animated_vertices` arrays of a frame's Characters) -- byte for byte what the group's
    own buffer receives; NULL restores that buffer."""
    kw = dict(packed={}, tbn_uv=dict(tbn=True, texcoord=True), align16=dict(align16=True))[layout]
    sm = synth.make_mesh(2_501, 24, 300, seed=77, index_order="runs", **kw)
    mesh = sm.flat
    n = 13
    pal = synth.make_palettes(mesh.num_transforms, n, 78).reshape(n, -1)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    stride = dmesh.out_strides[0]
    nbytes = (mesh.num_rows * stride + 15) // 16 * 16   # the kernel's bulk stores move whole 16-byte units
    hosts = [np.zeros(nbytes + 64, np.uint8) for _ in range(n)]
    views = []
    try:
        grp.set_palettes(pal)
        grp.animate()
        expect = grp.read_output(0).copy()
        for h in hosts:
            off = (-h.ctypes.data) % 16
            v = h[off:off + nbytes]
            gpu_ctx.host_register(v)
            views.append(v)
        grp.set_output_pointers([v.ctypes.data for v in views])
        grp.animate()
        gpu_ctx.sync()
        for i, v in enumerate(views):
            assert np.array_equal(v[:mesh.num_rows * stride].reshape(mesh.num_rows, stride), expect[i]), f"instance {i}"
        with pytest.raises(_capi.P3csError):
            grp.read_output(0)
        grp.set_output_pointers(None)
        grp.set_palettes(pal[::-1].copy())
        grp.animate()
        assert np.array_equal(grp.read_output(0), expect[::-1])
    finally:
        for v in views:
            gpu_ctx.host_unregister(v)
        grp.close()
        dmesh.close()


# ---- columns of every numeric type: GeomVertexRewriter branches + the GeomVertexColumn packers

_TYPED_NT = [F.NT_uint8, F.NT_uint16, F.NT_uint32, F.NT_int8, F.NT_int16, F.NT_int32, F.NT_float32, F.NT_float64]


def _random_typed_mesh(seed, num_rows=4000):
    """A random format: 3..7 skinned columns of random numeric type / width / alignment, colour and texcoord columns
    with morphs, integer deltas."""
    rng = np.random.default_rng(seed)
    cols, morphs = [], []

    def pick(contents, name):
        r = int(rng.integers(0, 11))
        if r == 8:
            return (name, 4, F.NT_packed_dcba, contents, 1)
        if r == 9:
            return (name, 4, F.NT_packed_dabc, contents, int(rng.choice([1, 4])))
        if r == 10:
            return (name, 3, F.NT_packed_ufloat, contents, 4)
        nt = _TYPED_NT[r]
        nv = int(rng.integers(1, 5))
        if contents == F.C_normal and nv < 3:
            nv = 3
        cb = F.NUMERIC_BYTES[nt]
        align = max(cb, 4) if nt in (F.NT_float32, F.NT_float64) else int(rng.choice([1, cb, 4]))
        return (name, nv, nt, contents, align)

    cols.append(pick(F.C_point, "vertex"))
    cols.append(pick(F.C_normal, "normal"))
    for i in range(int(rng.integers(1, 5))):
        cols.append(pick(F.C_vector if rng.random() < 0.6 else F.C_point, f"extra{i}"))
    color_nt = [(4, F.NT_uint8), (4, F.NT_packed_dabc), (4, F.NT_packed_dcba), (4, F.NT_uint16), (3, F.NT_uint8), (3, F.NT_uint32), (4, F.NT_float32)]
    nv, nt = color_nt[int(rng.integers(0, len(color_nt)))]
    cols.append(("color", nv, nt, F.C_color, 4))
    cols.append(("texcoord", int(rng.integers(2, 5)), F.NT_float32 if rng.random() < 0.5 else F.NT_int16, F.C_texcoord, 4))
    delta_nt = [F.NT_float32, F.NT_int8, F.NT_int16, F.NT_float64]
    for name, nv, nt, contents, _ in list(cols):
        if nt == F.NT_packed_ufloat and rng.random() < 0.5:
            continue
        if rng.random() < 0.6:
            dn = F.NT_float32 if contents == F.C_color else delta_nt[int(rng.integers(0, 4))]
            morphs.append((name, int(rng.integers(1, 5)) if contents != F.C_color else nv, dn))
    rows = None if rng.random() < 0.5 else [(5, num_rows // 3), (num_rows // 2, num_rows - 7)]
    return synth.make_typed_mesh(num_rows, 24, 300, seed=1000 + seed, columns=cols, morphs=morphs, num_sliders=3 if morphs else 0,
                                 slider_band=0.5, rows=rows)


@pytest.mark.parametrize("seed", range(12))
def test_columns_of_every_numeric_type_vs_oracle(gpu_ctx, monkeypatch, seed):
    """Random formats of integer / packed / short columns, a crowd of three instances, both kernels that take them (the
    strip kernel's generic mode and the two-kernel path): every output byte equals the oracle's (palettes without the
    uniform-scale libm branch on even seeds, non-uniformly scaled on odd ones)."""
    from oracle import oracle
    sm = _random_typed_mesh(seed)
    mesh = sm.flat
    scale = 1.0 if seed % 2 == 0 else (1.3, 0.8, 1.1)
    pal = synth.make_palettes(24, 3, 2000 + seed, scale=scale)
    sl = synth.make_sliders(mesh.num_sliders, 3, 3000 + seed) if mesh.num_sliders else None
    om = oracle.OracleMesh(mesh)
    expect = [om.animate(pal[i], None if sl is None else sl[i])[0] for i in range(3)]
    for path in ("strip", "wide"):
        monkeypatch.setenv("P3CS_PATH", path)
        dmesh = _capi.Mesh(gpu_ctx, mesh)
        grp = _capi.Group(dmesh, 3)
        try:
            grp.set_palettes(pal)
            if sl is not None:
                grp.set_sliders(sl)
            grp.animate()
            for i in range(3):
                got = grp.read_output(0, i, 1)
                assert np.array_equal(np.asarray(got).reshape(mesh.num_rows, -1), expect[i][0]), \
                    f"seed {seed} path {path} instance {i}: {(np.asarray(got).reshape(mesh.num_rows, -1) != expect[i][0]).sum()} bytes differ"
        finally:
            grp.close()
            dmesh.close()


def test_integer_conversions_at_the_edges(gpu_ctx):
    """Values that leave the column type's range wrap exactly as the reference's x86-64 casts do (cvttss2si: low bits
    kept for unsigned, 0x80000000 for signed overflow): large palettes push int8 / uint8 / int16 columns far out."""
    from oracle import oracle
    sm = synth.make_typed_mesh(2000, 8, 40, seed=77, columns=[
        ("vertex", 3, F.NT_int8, F.C_point, 1), ("normal", 3, F.NT_int16, F.C_normal, 2), ("t", 3, F.NT_uint8, F.C_vector, 1),
        ("b", 4, F.NT_uint16, F.C_vector, 2), ("p4", 4, F.NT_int16, F.C_point, 2), ("w", 3, F.NT_int32, F.C_point), ("u", 3, F.NT_uint32, F.C_vector)])
    mesh = sm.flat
    pal = synth.make_palettes(8, 1, 78, scale=(37.0, -11.5, 3.25))[0]
    pal[:, 12:15] *= 500.0   # translations in the hundreds
    expect = oracle.OracleMesh(mesh).animate(pal, None)[0]
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, 1)
    try:
        got = grp.animate_vertices(pal, None)
        assert np.array_equal(np.asarray(got[0]).reshape(mesh.num_rows, -1), expect[0])
    finally:
        grp.close()
        dmesh.close()
