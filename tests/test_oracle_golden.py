"""Pins oracle/skin_oracle.c (the CPU restatement) against the UNMODIFIED reference.

tests/golden/*.p3cs were produced by oracle/ref_harness.cxx linked against Panda3D 1.11.0
built from /root/reference (tests/golden/make_golden.py): inputs as the reference's objects
report them, outputs = the bytes GeomVertexData::animate_vertices returned.  The bar is
bit-exactness of every output byte and of every blended matrix.
"""
import numpy as np
import pytest

from conftest import golden_ids, golden_paths
from oracle import oracle
from panda3d_b200.casefile import load_case
from panda3d_b200.flat import mesh_from_case


def test_golden_fixtures_present():
    ids = golden_ids()
    assert "panda_walk4" in ids and "morph" in ids and len(ids) >= 15


@pytest.mark.parametrize("path", golden_paths(), ids=golden_ids())
def test_oracle_matches_reference_bit_exact(path):
    case = load_case(path)
    mesh = mesh_from_case(case)
    mesh.validate()
    om = oracle.OracleMesh(mesh)
    frames = int(case["meta"][7])
    assert frames >= 1
    for fr in range(frames):
        sliders = case["sliders"][fr] if mesh.num_sliders else None
        outs, blends, sel = om.animate(case["palette"][fr], sliders, want_blends=True, want_selection=True)
        assert len(outs) == int(case["meta"][2])
        for o, out in enumerate(outs):
            expect = case[f"expect{o}"][fr].reshape(mesh.num_rows, -1)
            assert np.array_equal(out, expect), f"frame {fr} output {o}: {(out != expect).sum()} bytes differ"
        assert np.array_equal(blends.view(np.uint32), case["expect_blends"][fr].view(np.uint32))
        # selection: rows inside TransformBlendTable::get_rows() pick blendt[row], others none
        bi = mesh.blend_indices()
        mask = np.zeros(mesh.num_rows, bool)
        for b, e in mesh.row_ranges:
            mask[b:e] = True
        assert np.array_equal(sel[mask], bi[mask]) and np.all(sel[~mask] == -1)


def test_panda_walk4_shape_matches_survey():
    """C1 = 1338 rows, 197 blends, 26 transforms, stride 32."""
    case = load_case([p for p in golden_paths() if p.endswith("panda_walk4.p3cs")][0])
    mesh = mesh_from_case(case)
    assert (mesh.num_rows, mesh.num_blends, mesh.num_transforms) == (1338, 197, 26)
    assert mesh.strides == [32, 2] and mesh.is_output == [True, False]
    # real data exercises the identity-scale and the inverse-transpose normal branches
    # (blended rotations are not orthonormal)
    branches = {oracle.normal_xform(b)[0] for fr in range(3) for b in case["expect_blends"][fr]}
    assert {0, 3} <= branches


def test_golden_branch_coverage():
    """The synthetic fixtures isolate each normal-matrix branch of geomVertexData.cxx:1811-1840."""
    want = {"rigid": {0}, "rigid_uniform_scale": {1}, "rigid_nonuniform_scale": {3}}
    for p in golden_paths():
        name = p.rsplit("/", 1)[1][:-5]
        if name in want:
            case = load_case(p)
            got = {oracle.normal_xform(b)[0] for b in case["expect_blends"][0]}
            assert got == want[name], (name, got)


def test_oracle_pose_matches_reference_skinning_matrices():
    """"Next" row f1: the joint hierarchy.  oracle.pose (channel tables -> compose_matrix -> net transform
    -> skinning matrix) reproduces bit for bit the CharacterJoint::_skinning_matrix values the reference
    computed for panda-walk4 at every exported frame (the golden palette)."""
    from panda3d_b200.flat import skeleton_from_case
    case = load_case([p for p in golden_paths() if p.endswith("panda_walk4.p3cs")][0])
    skel = skeleton_from_case(case)
    assert skel.num_joints == 42 and len(skel.palette_joint) == 26     # 42 <Joint>, 26 referenced
    for fi, frame in enumerate(case["anim_frames"]):
        skin = oracle.pose(skel, int(frame))
        assert np.array_equal(skin[skel.palette_joint].view(np.uint32), case["palette"][fi].view(np.uint32))


@pytest.mark.parametrize("fixture", ["panda_walk4_frame_blend", "panda_walk4_frame_blend_linear", "panda_walk4_frame_blend_componentwise_quat"])
def test_oracle_pose_with_frame_blending_matches_reference(fixture):
    """The same with PartBundle::set_frame_blend_flag(true) and the Actor posed BETWEEN frames
    (tests/golden/skeleton/panda_walk4_frame_blend.p3cs: MovingPartMatrix::get_blend_value, BT_linear):
    bit for bit the reference's skinning matrices, and the animated vertices that follow from them."""
    import os
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import mesh_from_case, skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    blend_type = int(case["anim_blend_type"][0])
    skel, mesh = skeleton_from_case(case), mesh_from_case(case)
    fracs = case["anim_fracs"]
    assert (fracs > 0).any() and (fracs == 0).any()
    om = oracle.OracleMesh(mesh)
    for fi, (frame, nxt, frac) in enumerate(zip(case["anim_frames"], case["anim_next_frames"], fracs)):
        skin = oracle.pose(skel, int(frame), int(nxt), float(frac), blend_type)
        pal = skin[skel.palette_joint]
        assert np.array_equal(pal.view(np.uint32), case["palette"][fi].view(np.uint32)), f"frame {frame}+{frac}"
        outs, _, _ = om.animate(pal)
        assert np.array_equal(outs[0].reshape(-1), np.asarray(case["expect0"][fi]).reshape(-1))


@pytest.mark.parametrize("fixture", ["panda_walk4_forced", "panda_walk4_forced_frame_blend"])
def test_oracle_pose_with_forced_channels_matches_reference(fixture):
    """PartBundle::freeze_joint on two joints and control_joint on a third (oracle/ref_harness.cxx, P3CS_FREEZE /
    P3CS_CONTROL): a joint with a forced channel takes that channel's value whatever is playing
    (chan/movingPartMatrix.cxx:53-60); the oracle reproduces the reference's skinning matrices bit for bit."""
    import os
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    skel = skeleton_from_case(case)
    fb = "anim_next_frames" in case
    bt = int(case["anim_blend_type"][0]) if fb else 1
    assert case["forced_joints"].size == 3
    for fi, frame in enumerate(case["anim_frames"]):
        nxt = case["anim_next_frames"][fi] if fb else frame
        frac = case["anim_fracs"][fi] if fb else 0.0
        skin = oracle.pose_controls([skel], [frame], [nxt], [frac], [1.0], fb, bt, case["forced_joints"], case["forced_values"][fi])
        assert np.array_equal(skin[skel.palette_joint].view(np.uint32), case["palette"][fi].view(np.uint32)), f"frame {frame}"
        plain = oracle.pose_controls([skel], [frame], [nxt], [frac], [1.0], fb, bt)
        assert not np.array_equal(plain[skel.palette_joint], case["palette"][fi])


MULTI = ["panda_walk4_two_controls_hold", "panda_walk4_two_controls_hold_linear", "panda_walk4_two_controls_hold_componentwise",
         "panda_walk4_two_controls_frame_blend", "panda_walk4_two_controls_frame_blend_linear",
         "panda_walk4_two_controls_frame_blend_componentwise", "panda_walk4_two_controls_hold_componentwise_quat",
         "panda_walk4_two_controls_frame_blend_componentwise_quat"]


@pytest.mark.parametrize("fixture", MULTI)
def test_oracle_pose_with_two_controls_matches_reference(fixture):
    """Two AnimControls blended by their effects (PartBundle::set_anim_blend_flag / set_control_effect), with the
    frames held or blended, BT_normalized_linear and BT_linear: oracle.pose_controls reproduces the reference's
    skinning matrices bit for bit (tests/golden/skeleton/, oracle/ref_harness.cxx `eggmulti`)."""
    import os
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, "skeleton", fixture + ".p3cs"))
    skel = skeleton_from_case(case)
    fb, bt = bool(case["anim_frame_blend"][0]), int(case["anim_blend_type"][0])
    assert (case["ctrl_effects"] == 0).any(), "one sample must drop a control from the blend"
    for i in range(case["ctrl_frames"].shape[0]):
        skin = oracle.pose_controls([skel, skel], case["ctrl_frames"][i], case["ctrl_next_frames"][i], case["ctrl_fracs"][i],
                                    case["ctrl_effects"][i], fb, bt)
        assert np.array_equal(skin[skel.palette_joint].view(np.uint32), case["palette"][i].view(np.uint32)), f"sample {i}"


def _control_arrays(case):
    """(frames, next_frames, fracs, effects, frame_blend) per sample of a skeleton fixture, one or two controls."""
    if "ctrl_frames" in case:
        return case["ctrl_frames"], case["ctrl_next_frames"], case["ctrl_fracs"], case["ctrl_effects"], bool(case["anim_frame_blend"][0])
    f = case["anim_frames"].reshape(-1, 1)
    if "anim_next_frames" in case:
        return f, case["anim_next_frames"].reshape(-1, 1), case["anim_fracs"].reshape(-1, 1), np.ones_like(f, np.float32), True
    return f, f, np.zeros_like(f, np.float32), np.ones_like(f, np.float32), False


@pytest.mark.parametrize("fixture", ["ripple", "skeleton/ripple_frame_blend", "skeleton/ripple_two_controls_hold",
                                     "skeleton/ripple_two_controls_frame_blend"])
def test_oracle_slider_channels_match_reference(fixture):
    """models/ripple.egg, the reference's morph-animated character: the slider values CharacterSlider takes from its
    scalar channels (MovingPartScalar::get_blend_value, chan/movingPartScalar.cxx:43-105) -- plain, between frames
    and with two blended controls -- bit for bit, then the animated vertices that follow from palette + sliders."""
    import os
    from conftest import GOLDEN_DIR
    from panda3d_b200.flat import mesh_from_case, skeleton_from_case
    case = load_case(os.path.join(GOLDEN_DIR, fixture + ".p3cs"))
    skel, mesh = skeleton_from_case(case), mesh_from_case(case)
    assert skel.num_sliders == mesh.num_sliders == 7
    frames, nxt, fracs, effects, fb = _control_arrays(case)
    K = frames.shape[1]
    bt = int(case["anim_blend_type"][0]) if "anim_blend_type" in case else 1
    om = oracle.OracleMesh(mesh)
    for i in range(frames.shape[0]):
        sl = oracle.pose_sliders([skel] * K, frames[i], nxt[i], fracs[i], effects[i], fb)
        assert np.array_equal(sl.view(np.uint32), case["sliders"][i].view(np.uint32)), f"sample {i}: {sl} vs {case['sliders'][i]}"
        pal = oracle.pose_controls([skel] * K, frames[i], nxt[i], fracs[i], effects[i], fb, bt)[skel.palette_joint]
        assert np.array_equal(pal.view(np.uint32), case["palette"][i].view(np.uint32))
        outs, _, _ = om.animate(pal, sl)
        assert np.array_equal(outs[0].reshape(-1), np.asarray(case["expect0"][i]).reshape(-1))
