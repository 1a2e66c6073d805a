"""Host-side mirror of the reference interface (panda3d_b200/gobj.py): the ingredient tests the
reference itself carries for this path, re-expressed against the mirror:
tests/putil/test_updateseq.py, tests/putil/test_sparsearray.py, tests/gobj/test_geom_vertex_format.py,
plus the flattening that hands a GeomVertexData to the C-ABI."""
import numpy as np
import pytest

from panda3d_b200 import synth
from panda3d_b200.gobj import (GeomEnums, GeomVertexAnimationSpec, GeomVertexArrayFormat, GeomVertexData,
                               GeomVertexFormat, InternalName, SliderTable, SparseArray, TransformBlend,
                               TransformBlendTable, UpdateSeq, UserVertexSlider, UserVertexTransform, load_prc_file_data)


# ---- UpdateSeq (reference: tests/putil/test_updateseq.py:4-108)
def test_updateseq_initial_old_fresh_ordering():
    seq = UpdateSeq()
    assert seq == UpdateSeq.initial() and seq.is_special() and seq.is_initial() and seq.seq == 0
    fresh, old = UpdateSeq.fresh(), UpdateSeq.old()
    assert seq < fresh and seq <= fresh and seq != fresh and not seq > fresh
    assert seq < old and not seq >= old
    assert old.seq == 1 and old > UpdateSeq.initial() and old < fresh
    assert fresh > old and fresh >= old and not fresh < UpdateSeq.initial()
    s = UpdateSeq(UpdateSeq.fresh())
    s.clear()
    assert s.is_initial()


def test_updateseq_increment_skips_special_values():
    s = UpdateSeq.old()
    s.increment()
    assert s.seq == 2 and not s.is_special()
    w = UpdateSeq._make(0xFFFFFFFE)
    w.increment()                      # wraps past fresh/initial/old
    assert w.seq == 2
    a, b = UpdateSeq._make(10), UpdateSeq._make(0x80000005)
    assert (a < b) != (b < a)          # circular comparison is antisymmetric


# ---- SparseArray (reference: tests/putil/test_sparsearray.py:9-330)
def test_sparsearray_ranges():
    s = SparseArray()
    assert s.is_zero() and not s.is_inverse() and s.get_num_subranges() == 0
    s.set_range(3, 4)
    assert [s.get_bit(i) for i in range(2, 8)] == [False, True, True, True, True, False]
    s.set_range(7, 2)                  # adjacent ranges merge
    assert s.get_num_subranges() == 1 and (s.get_subrange_begin(0), s.get_subrange_end(0)) == (3, 9)
    s.clear_range(4, 2)
    assert s.get_num_subranges() == 2 and s.get_num_on_bits() == 4
    assert s.get_lowest_on_bit() == 3 and s.get_highest_on_bit() == 8
    assert SparseArray.lower_on(5) == SparseArray.range(0, 5)
    assert SparseArray.bit(7).get_bit(7) and not SparseArray.bit(7).get_bit(6)


def test_sparsearray_ops_and_inverse():
    a, b = SparseArray.range(0, 10), SparseArray.range(5, 10)
    assert (a & b) == SparseArray.range(5, 5)
    assert (a | b) == SparseArray.range(0, 15)
    assert (a ^ b).get_num_on_bits() == 10
    inv = ~a
    assert inv.is_inverse() and not inv.get_bit(3) and inv.get_bit(100)
    assert SparseArray.all_on().is_all_on() and (~SparseArray.all_on()).is_zero()
    assert a.has_all_of(2, 5) and a.has_any_of(8, 10) and not a.has_any_of(10, 3)


# ---- formats (reference: tests/gobj/test_geom_vertex_format.py, geomVertexColumn.cxx:154-234)
def loader_format(align16=False, morph=False):
    af = GeomVertexArrayFormat()
    af.add_column(InternalName.get_vertex(), 3, GeomEnums.NT_float32, GeomEnums.C_point)
    af.add_column(InternalName.get_normal(), 3, GeomEnums.NT_float32, GeomEnums.C_normal)
    af.add_column(InternalName.get_texcoord(), 2, GeomEnums.NT_float32, GeomEnums.C_texcoord)
    bf = GeomVertexArrayFormat()
    bf.add_column(InternalName.get_transform_blend(), 1, GeomEnums.NT_uint16, GeomEnums.C_index, 0, 2)
    if morph:
        bf.add_column(InternalName.get_morph("vertex", "smile"), 3, GeomEnums.NT_float32, GeomEnums.C_morph_delta)
    f = GeomVertexFormat(af)
    f.add_array(bf)
    spec = GeomVertexAnimationSpec()
    spec.set_panda()
    f.set_animation(spec)
    if align16:
        f.align_columns_for_animation()
    return GeomVertexFormat.register_format(f)


def test_format_strides_match_the_reference_probe():
    """strides 32 / 2; with a morph column 32 / 16; align-16 layout stride 48."""
    f = loader_format()
    assert (f.get_array(0).get_stride(), f.get_array(1).get_stride()) == (32, 2)
    f = loader_format(morph=True)
    assert (f.get_array(0).get_stride(), f.get_array(1).get_stride()) == (32, 16)
    assert f.get_num_morphs() == 1 and f.get_morph_slider(0) == "smile" and f.get_morph_base(0) == "vertex"
    f = loader_format(align16=True)
    a0 = f.get_array(0)
    assert a0.get_stride() == 48
    assert [(a0.get_column(n).get_start(), a0.get_column(n).get_num_values()) for n in ("vertex", "normal", "texcoord")] == \
        [(0, 4), (16, 4), (32, 2)]


def test_post_animated_format_drops_blend_and_morph_columns():
    f = loader_format(morph=True)
    post = f.get_post_animated_format()
    assert post.get_num_arrays() == 1 and post.get_array(0).get_stride() == 32   # stride is kept
    assert post.get_animation().get_animation_type() == GeomEnums.AT_none
    assert post.get_num_points() == 1 and post.get_num_vectors() == 1 and not post.has_column("transform_blend")


# ---- blend tables
def test_transform_blend_semantics():
    t0, t1, t2 = (UserVertexTransform(n) for n in "abc")
    b = TransformBlend(t0, 0.3)                # two-argument constructor ignores the weight (transformBlend.I:24-27)
    assert b.get_num_transforms() == 1 and b.get_weight(0) == 1.0
    b = TransformBlend()
    b.add_transform(t2, 0.5)
    b.add_transform(t0, 0.25)
    b.add_transform(t2, 0.25)                  # same transform: weights merge
    b.add_transform(t1, 1e-9)                  # nearly-zero weight is dropped
    assert [b.get_transform(i) for i in range(b.get_num_transforms())] == [t0, t2]   # address order
    assert b.get_weight(t2) == pytest.approx(0.75)
    b.normalize_weights()
    assert sum(b.get_weight(i) for i in range(2)) == pytest.approx(1.0)
    b.limit_transforms(1)
    assert b.get_num_transforms() == 1 and b.get_transform(0) is t2
    tbt = TransformBlendTable()
    i0 = tbt.add_blend(TransformBlend(t0, 1.0))
    assert tbt.add_blend(TransformBlend(t0, 1.0)) == i0     # identical blends are shared
    assert tbt.add_blend(TransformBlend(t1, 1.0)) == i0 + 1


def test_staleness_chain():
    t = UserVertexTransform("j")
    tbt = TransformBlendTable()
    tbt.add_blend(TransformBlend(t, 1.0))
    m0 = tbt.get_modified()
    t.set_matrix(np.eye(4))
    assert tbt.get_modified() > m0


def test_flatten_matches_synth_flat_mesh():
    """A GeomVertexData built through the mirror flattens to the same p3cs_mesh_desc content as
    the synthetic generator's direct FlatMesh (formats, blend CSR in stored order, rows)."""
    sm = synth.make_mesh(200, 6, 20, seed=3, texcoord=True, num_sliders=2, slider_band=0.3)
    fm = sm.flat
    af = GeomVertexArrayFormat()
    bf = GeomVertexArrayFormat()
    for name, c in sm.columns:
        (af if c.array == 0 else bf).add_column(name, c.num_values, c.numeric_type, c.contents, c.start,
                                                 2 if name == "transform_blend" else 0)
    f = GeomVertexFormat(af)
    f.add_array(bf)
    spec = GeomVertexAnimationSpec()
    spec.set_panda()
    f.set_animation(spec)
    fmt = GeomVertexFormat.register_format(f)
    assert [fmt.get_array(i).get_stride() for i in range(2)] == fm.strides
    vd = GeomVertexData("m", fmt)
    vd.set_num_rows(fm.num_rows)
    vd.set_array(0, fm.arrays[0])
    vd.set_array(1, fm.arrays[1])
    joints = [UserVertexTransform(f"j{i}") for i in range(fm.num_transforms)]
    tbt = TransformBlendTable()
    for b in range(fm.num_blends):
        bl = TransformBlend()
        for e in range(fm.blend_offsets[b], fm.blend_offsets[b + 1]):
            bl.add_transform(joints[fm.blend_transform[e]], float(fm.blend_weight[e]))
        assert tbt.add_blend(bl) == b
    tbt.set_rows(SparseArray.lower_on(fm.num_rows))
    vd.set_transform_blend_table(tbt)
    st = SliderTable()
    for name, rows in zip(sm.slider_names, sm.slider_rows):
        sa = SparseArray()
        for b, e in rows:
            sa.set_range(int(b), int(e - b))
        st.add_slider(UserVertexSlider(name), sa)
    vd.set_slider_table(SliderTable.register_table(st))
    got, palette = vd.flatten()
    assert got.num_rows == fm.num_rows and got.is_output == fm.is_output and got.num_transforms == fm.num_transforms
    assert [c.as_tuple() for c in got.skin_columns] == [c.as_tuple() for c in fm.skin_columns]
    assert got.blend_column.as_tuple() == fm.blend_column.as_tuple()
    assert np.array_equal(got.blend_offsets, fm.blend_offsets) and np.array_equal(got.blend_transform, fm.blend_transform)
    assert np.array_equal(got.blend_weight, fm.blend_weight) and np.array_equal(got.row_ranges, fm.row_ranges)
    assert sorted((m.slider, m.base.as_tuple(), m.delta.as_tuple()) for m in got.morphs) == \
        sorted((m.slider, m.base.as_tuple(), m.delta.as_tuple()) for m in fm.morphs)
    got.validate()


def test_cpu_backend_is_not_a_fallback():
    fmt = loader_format()
    vd = GeomVertexData("m", fmt)
    vd.set_num_rows(4)
    t = UserVertexTransform("j")
    tbt = TransformBlendTable()
    tbt.add_blend(TransformBlend(t, 1.0))
    tbt.set_rows(SparseArray.lower_on(4))
    vd.set_transform_blend_table(tbt)
    t.set_matrix(np.eye(4))
    load_prc_file_data("", "vertex-animation-backend cpu")
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            vd.animate_vertices(True)
    finally:
        load_prc_file_data("", "vertex-animation-backend cuda")


def test_unanimated_data_is_returned_as_is():
    af = GeomVertexArrayFormat(InternalName.get_vertex(), 3, GeomEnums.NT_float32, GeomEnums.C_point)
    fmt = GeomVertexFormat.register_format(GeomVertexFormat(af))
    vd = GeomVertexData("static", fmt)
    assert vd.animate_vertices(True) is vd          # not AT_panda (geomVertexData.cxx:919-921)
