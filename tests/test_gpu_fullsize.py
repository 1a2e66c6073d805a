"""GPU parity at the FULL sizes of BASELINE.json configs[2..4] (C3, C4, C5), through the C-ABI, against the oracle
(oracle/skin_oracle.c, itself pinned bit-exact to the reference by tests/test_oracle_golden.py).

Every byte must equal the oracle's -- asserted -- except the normals of the rare blends that land in the uniform-scale
branch, where libm and CUDA differ in the last place of sinf / cosf / atan2f (assert_bytes_equal_except_libm checks
that every differing row really is such a row).  Tolerances (REL_TOL, per column) come from compare_outputs.
"""
import numpy as np
import pytest

from oracle import oracle
from panda3d_b200 import _capi, synth
from test_gpu_parity import compare_outputs, expected_selection

pytestmark = pytest.mark.gpu

C3_INSTANCES = 10_000


def assert_bytes_equal_except_libm(mesh, om, palette, sliders, got, what):
    """Every byte equals the oracle's, except inside C_normal columns of rows whose blended matrix takes the
    uniform-scale branch (oracle.unexplained_differences; a random 4-weight blend lands there about once in 10^4).
    REL_TOL is enforced on all of it by compare_outputs.  Returns the number of such bytes."""
    explained, unexplained, expect = oracle.unexplained_differences(mesh, got, palette, sliders, om)
    compare_outputs(mesh, got, expect, what)
    assert unexplained == 0, f"{what}: {unexplained} bytes differ from the oracle outside the libm-dependent normals"
    return explained


def _c3_sample(n):
    """>= 32 instances spread over the launch: first, last, both sides of power-of-two and SM-count multiples (the
    launch deals instances to CTAs in waves of 148 SMs x resident CTAs) and a pseudo-random handful."""
    picks = {0, 1, n - 1, n - 2, n // 2, 147, 148, 149, 148 * 4 - 1, 148 * 4, 148 * 8, 148 * 16 + 1, 255, 256, 4095, 4096,
             8191, 8192, 65535 % n}
    rng = np.random.default_rng(77)
    picks.update(int(i) for i in rng.integers(0, n, 16))
    return sorted(p for p in picks if 0 <= p < n)


@pytest.mark.parametrize("index_order", ["runs", "random"])
def test_full_size_c3_crowd(gpu_ctx, index_order):
    """configs[2]: 10 000 instances x (20 000 vertices, 128 joints, 4096 blends) in ONE launch; >= 32 instances of
    it byte for byte against the oracle.  `runs` is the bench workload, `random` the per-vertex random index order."""
    sm = synth.config_c3_mesh(index_order=index_order)
    mesh = sm.flat
    n = C3_INSTANCES
    T = mesh.num_transforms
    sample = _c3_sample(n)
    assert len(sample) >= 32
    # palettes: unique for the sampled instances, a cycle of 64 for the others (generation time)
    cyc = synth.make_palettes(T, 64, 2024).reshape(64, T * 16)
    pal = np.ascontiguousarray(cyc[np.arange(n) % 64])
    uniq = synth.make_palettes(T, len(sample), 2025).reshape(len(sample), T * 16)
    pal[sample] = uniq
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    grp = _capi.Group(dmesh, n)
    try:
        assert np.array_equal(grp.read_selection(), expected_selection(mesh)), "blend-index / row selection differs"
        grp.set_palettes(pal)
        grp.animate()
        assert grp.timings().launches == 1
        libm_bytes = 0
        for i in sample:
            got = grp.read_output(0, i, 1)[0]
            libm_bytes += assert_bytes_equal_except_libm(mesh, om, pal[i].reshape(T, 16), None, [got], f"C3/{index_order} instance {i}")
        # instances sharing a palette must share their bytes (covers every instance of the launch cheaply):
        # a checksum per instance, equal within each palette class
        out = grp.read_output(0)
        sums = out.reshape(n, -1).view(np.uint64).sum(axis=1)
        cls = np.arange(n) % 64
        plain = np.ones(n, bool)
        plain[sample] = False
        for c in range(64):
            s = sums[plain & (cls == c)]
            assert (s == s[0]).all(), f"instances of palette class {c} differ among themselves"
        assert libm_bytes <= 64 * len(sample)
        print(f"C3/{index_order}: {len(sample)} instances equal to the oracle but for {libm_bytes} bytes in uniform-scale normals")
    finally:
        grp.close()
        dmesh.close()


def test_full_size_c4_morph_mesh_and_crowd(gpu_ctx):
    """configs[3]: 250 000 vertices, 64 sliders (25 % exactly zero: the `!= 0` skip, geomVertexData.cxx:1483) +
    4-weight skinning -- as a single mesh (animate_vertices, host buffers) and as a crowd of 24 faces."""
    sm = synth.config_c4()
    mesh = sm.flat
    assert mesh.num_rows == 250_000 and mesh.num_sliders == 64
    T, S = mesh.num_transforms, mesh.num_sliders
    n = 24
    pal = synth.make_palettes(T, n, 3031)
    sl = synth.make_sliders(S, n, 3032)
    assert (sl == 0.0).any() and (sl != 0.0).any()
    om = oracle.OracleMesh(mesh)
    dmesh = _capi.Mesh(gpu_ctx, mesh)
    single = _capi.Group(dmesh, 1)
    crowd = _capi.Group(dmesh, n)
    try:
        assert np.array_equal(single.read_selection(), expected_selection(mesh))
        for fr in range(2):
            got = single.animate_vertices(pal[fr], sl[fr])
            assert_bytes_equal_except_libm(mesh, om, pal[fr], sl[fr], got, f"C4 single frame {fr}")
        crowd.set_palettes(pal.reshape(n, -1))
        crowd.set_sliders(sl)
        crowd.animate()
        for i in (0, 1, n // 2, n - 1):
            got = crowd.read_output(0, i, 1)[0]
            assert_bytes_equal_except_libm(mesh, om, pal[i], sl[i], [got], f"C4 crowd instance {i}")
        # all sliders zero: the morph pass must be a no-op (bytes equal to the same palette without sliders)
        zero = np.zeros(S, np.float32)
        a = single.animate_vertices(pal[0], zero)[0].copy()
        expect, _, _ = om.animate(pal[0], zero)
        assert np.array_equal(a.reshape(expect[0].shape), expect[0])
    finally:
        crowd.close()
        single.close()
        dmesh.close()


def test_full_size_c5_mixed_crowd(gpu_ctx):
    """configs[4]: the four character types (5k/20k/50k/125k vertices; 32/64/128/256 joints; vertex, normal, texcoord,
    tangent, binormal = stride 56) at full size, 2 M vertices per frame through ONE p3cs_animate_groups call."""
    spec = synth.config_c5_meshes()
    assert sum(m.flat.num_rows * c for m, c in spec) == 2_000_000
    meshes, groups, pals = [], [], []
    try:
        for i, (sm, count) in enumerate(spec):
            dm = _capi.Mesh(gpu_ctx, sm.flat)
            g = _capi.Group(dm, count)
            p = synth.make_palettes(sm.flat.num_transforms, count, 5100 + i)
            g.set_palettes(p.reshape(count, -1))
            meshes.append(dm)
            groups.append(g)
            pals.append(p)
        for g, (sm, _) in zip(groups, spec):
            assert np.array_equal(g.read_selection(), expected_selection(sm.flat))
        for rep in range(2):   # second call replays the captured graph
            _capi.animate_groups(groups)
        for g, (sm, count), p in zip(groups, spec, pals):
            om = oracle.OracleMesh(sm.flat)
            for i in sorted({0, count // 2, count - 1}):
                got = g.read_output(0, i, 1)[0]
                assert_bytes_equal_except_libm(sm.flat, om, p[i], None, [got], f"{sm.name} instance {i}")
    finally:
        for g in groups:
            g.close()
        for dm in meshes:
            dm.close()
