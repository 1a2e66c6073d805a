#!/usr/bin/env python
"""Regenerates tests/golden/*.p3cs by running the UNMODIFIED reference.

Needs oracle/_ref (bash oracle/build_ref.sh, which needs /root/reference): every
golden file holds the flattened inputs as the reference's own objects report them and
the byte image of what GeomVertexData::animate_vertices(true, thread) returned
(panda/src/gobj/geomVertexData.cxx:912), per frame.  Run from the repo root:
    python tests/golden/make_golden.py
The fixtures are small on purpose (a few hundred rows); they pin oracle/skin_oracle.c
bit for bit (tests/test_oracle_golden.py) and are the first parity target of the CUDA path.
"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from panda3d_b200 import synth  # noqa: E402
from panda3d_b200.casefile import save_case  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_harness")
REF = os.environ.get("P3D_REFERENCE", "/root/reference")


def cases():
    P, S = synth.make_palettes, synth.make_sliders
    mk = synth.make_mesh
    yield "basic", mk(600, 16, 120, seed=1), P(16, 2, 7), None
    yield "sorted", mk(600, 16, 120, seed=21, index_order="sorted"), P(16, 2, 27), None
    yield "tbn_uv", mk(500, 16, 100, seed=2, tbn=True, texcoord=True, index_order="runs"), P(16, 2, 8), None
    yield "rigid", mk(400, 16, 16, seed=3, weights_per_blend=1, index_order="sorted"), P(16, 2, 9), None
    yield "rigid_uniform_scale", mk(400, 16, 16, seed=3, weights_per_blend=1), P(16, 2, 10, scale=0.687), None
    yield "rigid_nonuniform_scale", mk(400, 16, 16, seed=3, weights_per_blend=1), P(16, 2, 11, scale=(1, 1.2, 0.8)), None
    yield "two_weights_near_uniform", mk(400, 8, 60, seed=31, weights_per_blend=2), P(8, 2, 32, scale=1.0, jitter=0.002), None
    yield "align16", mk(400, 12, 80, seed=4, align16=True, texcoord=True), P(12, 2, 12), None
    yield "align16_uniform_scale", mk(300, 12, 12, seed=4, align16=True, weights_per_blend=1), P(12, 2, 12, scale=1.3), None
    yield "align16_tbn", mk(300, 12, 60, seed=41, align16=True, tbn=True), P(12, 1, 42), None
    yield "morph", mk(800, 12, 100, seed=5, num_sliders=8, morph_columns=("vertex", "normal"), slider_band=0.2,
                      index_order="runs"), P(12, 3, 13), S(8, 3, 14)
    yield "rows_u32", mk(900, 12, 100, seed=6, rows=[(10, 300), (450, 451), (600, 899)], index_type="u32"), P(12, 2, 15), None
    yield "u8_align16_morph", mk(300, 12, 100, seed=7, index_type="u8", align16=True, num_sliders=3, slider_band=0.5), \
        P(12, 2, 16), S(3, 2, 17)
    # NT_float64 columns: the reference's GeomVertexRewriter branches (geomVertexData.cxx:1782-1799, 1862-1879)
    yield "f64", mk(300, 12, 60, seed=71, float64=True, tbn=True, index_order="runs"), P(12, 2, 72), None
    yield "f64_uniform_scale", mk(200, 8, 8, seed=73, float64=True, weights_per_blend=1), P(8, 1, 74, scale=0.8), None
    yield "f64_align16_morph", mk(300, 12, 60, seed=75, float64=True, align16=True, tbn=True, num_sliders=4, slider_band=0.4,
                                  morph_columns=("vertex", "normal"), rows=[(5, 250)]), P(12, 2, 76), S(4, 2, 77)
    # columns of every other numeric type, 1..4 values: the GeomVertexColumn::Packer family behind GeomVertexRewriter
    # (geomVertexColumn.cxx:457-1975), and morphs of colour / texcoord / integer base columns (geomVertexData.cxx:1489-1537)
    from panda3d_b200 import flat as F
    tm = synth.make_typed_mesh
    pt, vec, nrm, col, tex, oth = F.C_point, F.C_vector, F.C_normal, F.C_color, F.C_texcoord, F.C_other
    yield "typed_signed", tm(300, 12, 60, seed=81, columns=[
        ("vertex", 3, F.NT_int16, pt), ("normal", 4, F.NT_int8, nrm), ("tangent", 3, F.NT_int32, vec),
        ("binormal", 3, F.NT_uint16, vec), ("texcoord", 2, F.NT_float32, tex)]), P(12, 2, 82), None
    yield "typed_unsigned_point4", tm(300, 12, 60, seed=83, columns=[
        ("vertex", 4, F.NT_uint8, pt), ("normal", 3, F.NT_uint32, nrm), ("tangent", 4, F.NT_packed_dcba, vec),
        ("binormal", 4, F.NT_packed_dabc, vec), ("aux", 4, F.NT_int32, pt)]), P(12, 2, 84, scale=(1, 1.2, 0.8)), None
    yield "typed_short_columns", tm(300, 10, 50, seed=85, columns=[
        ("vertex", 2, F.NT_float32, pt), ("normal", 3, F.NT_int16, nrm, 2), ("tangent2", 2, F.NT_float64, vec, 8), ("tangent", 1, F.NT_int16, vec, 2),
        ("binormal", 1, F.NT_float32, vec), ("aux", 2, F.NT_int16, pt, 2), ("aux2", 1, F.NT_uint32, pt)]), P(10, 2, 86), None
    yield "typed_ufloat", tm(300, 10, 50, seed=87, columns=[
        ("vertex", 3, F.NT_packed_ufloat, pt), ("normal", 3, F.NT_packed_ufloat, nrm), ("tangent", 3, F.NT_packed_ufloat, vec)]), \
        P(10, 2, 88, scale=(1.5, 0.7, 1.1)), None
    yield "typed_unaligned", tm(300, 10, 50, seed=89, columns=[
        ("pad", 1, F.NT_uint8, oth, 1), ("vertex", 3, F.NT_uint8, pt, 1), ("normal", 3, F.NT_int16, nrm, 1),
        ("tangent", 4, F.NT_packed_dabc, vec, 1), ("binormal", 3, F.NT_uint16, vec, 1), ("aux", 4, F.NT_int16, pt, 1)],
        rows=[(7, 120), (150, 299)]), P(10, 2, 90), None
    yield "typed_morph_colors", tm(300, 10, 50, seed=91, num_sliders=3, columns=[
        ("vertex", 3, F.NT_float32, pt), ("normal", 3, F.NT_float32, nrm), ("color", 4, F.NT_uint8, col),
        ("color.b", 4, F.NT_packed_dabc, col), ("color.c", 4, F.NT_uint16, col), ("color.d", 3, F.NT_uint32, col),
        ("color.e", 4, F.NT_packed_dcba, col), ("color.f", 3, F.NT_packed_ufloat, col), ("color.g", 3, F.NT_uint8, col),
        ("color.h", 4, F.NT_float32, col), ("color.i", 3, F.NT_uint16, col)],
        morphs=[("color", 4, F.NT_float32), ("color.b", 4, F.NT_float32), ("color.c", 4, F.NT_float32), ("color.d", 3, F.NT_float32),
                ("color.e", 4, F.NT_float32), ("color.f", 3, F.NT_float32), ("color.g", 3, F.NT_float32), ("color.h", 4, F.NT_float32),
                ("color.i", 3, F.NT_float64)]), P(10, 3, 92), S(3, 3, 93)
    yield "typed_morph_points", tm(300, 10, 50, seed=94, num_sliders=4, columns=[
        ("vertex", 3, F.NT_int16, pt), ("normal", 3, F.NT_int8, nrm), ("texcoord", 2, F.NT_float32, tex),
        ("texcoord.b", 4, F.NT_int16, tex), ("texcoord.c", 3, F.NT_uint16, tex), ("aux", 4, F.NT_int32, pt), ("aux2", 4, F.NT_int16, oth)],
        morphs=[("vertex", 3, F.NT_int8), ("normal", 3, F.NT_float32), ("texcoord", 2, F.NT_float32), ("texcoord.b", 3, F.NT_int8),
                ("texcoord.c", 2, F.NT_uint8), ("aux", 3, F.NT_int16), ("aux2", 4, F.NT_float32)]), P(10, 3, 95), S(4, 3, 96)
    for cs, csname in ((2, "yup_right"), (3, "zup_left"), (4, "yup_left")):
        yield "cs_" + csname, mk(300, 8, 8, seed=8, weights_per_blend=1, coordinate_system=cs), P(8, 1, 18, scale=0.5), None


def hardware_cases():
    """Inputs of the AT_hardware conversion fixtures (tests/golden/hardware/): <= 4 entries per blend, more than 4
    (limit_transforms + normalize_weights), 1..3 entries with a u8 index and a rows subset."""
    mk = synth.make_mesh
    yield "hw_four", mk(700, 20, 150, seed=201, index_order="runs")
    yield "hw_six", mk(500, 16, 90, seed=202, weights_per_blend=6)
    yield "hw_two_u8", mk(300, 9, 40, seed=203, weights_per_blend=2, index_type="u8", rows=[(3, 200)])


def make_hardware_fixtures(tmp):
    from panda3d_b200.casefile import load_case
    os.makedirs(os.path.join(HERE, "hardware"), exist_ok=True)
    for name, sm in hardware_cases():
        recipe = os.path.join(tmp, name + ".recipe")
        case = sm.recipe(synth.make_palettes(sm.flat.num_transforms, 1, 7))
        save_case(recipe, case)
        raw = os.path.join(tmp, name + ".hw")
        subprocess.run([HARNESS, "hw", recipe, raw], check=True, timeout=120, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        case.update(load_case(raw))   # the recipe (inputs) + hw_weight / hw_index / hw_table (what the reference wrote)
        out = os.path.join(HERE, "hardware", name + ".p3cs")
        save_case(out, case)
        print("wrote", out, os.path.getsize(out))


def main():
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/bin/ref_harness missing: run `bash oracle/build_ref.sh` first")
    with tempfile.TemporaryDirectory() as tmp:
        make_hardware_fixtures(tmp)
        for name, sm, pal, sl in cases():
            recipe = os.path.join(tmp, name + ".recipe")
            save_case(recipe, sm.recipe(pal, sl))
            out = os.path.join(HERE, name + ".p3cs")
            subprocess.run([HARNESS, "run", recipe, out], check=True, timeout=120)
            print("wrote", out, os.path.getsize(out))
    # config C1: the panda Actor (models/panda-model.egg + panda-walk4.egg), frames 0/10/37
    model = os.path.join(REF, "models", "panda-model.egg")
    anim = os.path.join(REF, "models", "panda-walk4.egg")
    out = os.path.join(HERE, "panda_walk4.p3cs")
    subprocess.run([HARNESS, "egg", model, anim, out, "0", "10", "37", "59"], check=True, timeout=120,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("wrote", out, os.path.getsize(out))
    # models/ripple.egg: the reference's own morph-animated character (7 <S$Anim> sliders driving <Dxyz> / <DNormal>
    # deltas, model and animation in one egg) -- the morph path on loader-made formats
    ripple = os.path.join(REF, "models", "ripple.egg")
    out = os.path.join(HERE, "ripple.p3cs")
    subprocess.run([HARNESS, "egg", ripple, ripple, out, "0", "3", "7", "20", "44"], check=True, timeout=120,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("wrote", out, os.path.getsize(out))
    # ripple.egg with <DRGBA> / <Duv> morphs added (integration/make_morph_egg.py): the egg loader's colour column
    # (4 x uint8, or packed_dabc with vertex-colors-prefer-packed) and texcoord column as morph bases
    with tempfile.TemporaryDirectory() as tmp:
        derived = os.path.join(tmp, "ripple_rgba_uv.egg")
        subprocess.run([sys.executable, os.path.join(ROOT, "integration", "make_morph_egg.py"), ripple, derived], check=True)
        for suffix, env in (("", {}), ("_packed", {"P3CS_PRC": "vertex-colors-prefer-packed 1"})):
            out = os.path.join(HERE, f"ripple_rgba_uv{suffix}.p3cs")
            subprocess.run([HARNESS, "egg", derived, derived, out, "0", "3", "7", "20", "44"], check=True, timeout=120,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env={**os.environ, **env})
            print("wrote", out, os.path.getsize(out))
    # the same Actor with PartBundle::set_frame_blend_flag(true), posed between frames (joint hierarchy fixture, f1)
    os.makedirs(os.path.join(HERE, "skeleton"), exist_ok=True)
    for suffix, env in (("", {}), ("_linear", {"P3CS_BLEND_TYPE": "linear"}),   # anim-blend-type default, then BT_linear
                        ("_componentwise_quat", {"P3CS_BLEND_TYPE": "componentwise_quat"})):
        out = os.path.join(HERE, "skeleton", f"panda_walk4_frame_blend{suffix}.p3cs")
        subprocess.run([HARNESS, "egg", model, anim, out, "0.5", "10.25", "37.0", "58.75", "3.999"], check=True, timeout=120,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env={**os.environ, **env})
        print("wrote", out, os.path.getsize(out))
    # forced channels: PartBundle::freeze_joint on joints 6 and 24, control_joint on joint 29 (its node moves every sample)
    for suffix, frames_ in (("", ["0", "10", "37"]), ("_frame_blend", ["0.5", "10.25", "37.0"])):
        out = os.path.join(HERE, "skeleton", f"panda_walk4_forced{suffix}.p3cs")
        subprocess.run([HARNESS, "egg", model, anim, out] + frames_, check=True, timeout=120, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, env={**os.environ, "P3CS_FREEZE": "6,24", "P3CS_CONTROL": "29"})
        print("wrote", out, os.path.getsize(out))
    # two AnimControls blended by their effects (PartBundle::set_anim_blend_flag): hold-frame and frame-blend, both blend types
    samples_hold = ["3:0.7,20:0.3", "10:1.0,40:0.0", "0:0.5,59:0.5", "17:0.2,18:1.3"]
    samples_blend = ["3.5:0.7,20.25:0.3", "10.75:1.0,40:0.0", "5.5:0.25,33.25:0.75"]
    for samples, tag in ((samples_hold, "hold"), (samples_blend, "frame_blend")):
        for suffix, env in (("", {}), ("_linear", {"P3CS_BLEND_TYPE": "linear"}), ("_componentwise", {"P3CS_BLEND_TYPE": "componentwise"}),
                            ("_componentwise_quat", {"P3CS_BLEND_TYPE": "componentwise_quat"})):
            out = os.path.join(HERE, "skeleton", f"panda_walk4_two_controls_{tag}{suffix}.p3cs")
            subprocess.run([HARNESS, "eggmulti", model, anim, out] + samples, check=True, timeout=120,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env={**os.environ, **env})
            print("wrote", out, os.path.getsize(out))
    # ripple.egg between frames and with two controls: the morph sliders' scalar channels (MovingPartScalar)
    out = os.path.join(HERE, "skeleton", "ripple_frame_blend.p3cs")
    subprocess.run([HARNESS, "egg", ripple, ripple, out, "0.5", "3.25", "20.0", "43.75"], check=True, timeout=120,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("wrote", out, os.path.getsize(out))
    for samples, tag in ((["3:0.7,20:0.3", "10:1.0,40:0.0", "17:0.2,18:1.3"], "hold"), (["3.5:0.7,20.25:0.3", "5.5:0.25,33.25:0.75"], "frame_blend")):
        out = os.path.join(HERE, "skeleton", f"ripple_two_controls_{tag}.p3cs")
        subprocess.run([HARNESS, "eggmulti", ripple, ripple, out] + samples, check=True, timeout=120,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
