"""f3: the AT_hardware upload format.

tests/golden/hardware/*.p3cs hold a recipe (a mesh with its TransformBlendTable) and what the UNMODIFIED
reference wrote when converting it to an indexed hardware format with GeomVertexData::convert_to()
(gobj/geomVertexData.cxx:574-648; oracle/ref_harness.cxx `hw`, tests/golden/make_golden.py): the
transform_weight / transform_index columns and the TransformTable order.  p3cs_hardware_from_blends must
reproduce them bit for bit (host code: runs without a GPU); on the GPU, a mesh rebuilt from the hardware
columns animates to the same bytes as the original.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from panda3d_b200 import _capi, flat as F, synth
from panda3d_b200.casefile import load_case

HW = sorted(glob.glob(os.path.join(GOLDEN_DIR, "hardware", "*.p3cs")))


def mesh_from_recipe(case) -> F.FlatMesh:
    meta = case["meta"]
    rows, na = int(meta[0]), int(meta[1])
    cols = [F.Column(*map(int, r)) for r in case["columns"].reshape(-1, 5)]
    names = bytes(case["column_names"]).decode().split("\n")
    arrays = [np.ascontiguousarray(case[f"src_array{a}"]).reshape(rows, -1) for a in range(na)]
    skin = [c for c, n in zip(cols, names) if c.contents in (F.C_point, F.C_vector, F.C_normal) and ".morph." not in n]
    bc = F.Column(*map(int, case["blend_column"]))
    return F.FlatMesh(num_rows=rows, arrays=arrays, is_output=[a != bc.array for a in range(na)], skin_columns=skin,
                      blend_column=bc, num_transforms=int(meta[3]), blend_offsets=case["blend_offsets"].astype(np.int32),
                      blend_transform=case["blend_transform"].astype(np.int32), blend_weight=case["blend_weight"].astype(np.float32),
                      row_ranges=case["rows"].reshape(-1, 2).astype(np.int32), coordinate_system=int(meta[6]) or F.CS_zup_right)


@pytest.mark.parametrize("path", HW, ids=[os.path.basename(p)[:-5] for p in HW])
def test_hardware_columns_match_the_reference_conversion(path):
    case = load_case(path)
    mesh = mesh_from_recipe(case)
    w, ix, table = _capi.hardware_from_blends(mesh)
    assert np.array_equal(ix, case["hw_index"].reshape(-1, 4)), "transform_index differs from the reference's"
    assert np.array_equal(w.view(np.uint32), case["hw_weight"].reshape(-1, 4).view(np.uint32)), "transform_weight is not bit-identical"
    assert np.array_equal(table, case["hw_table"]), "TransformTable order differs"


def test_fixture_set_covers_the_over_four_branch():
    assert HW, "tests/golden/hardware is empty"
    counts = [np.diff(load_case(p)["blend_offsets"]).max() for p in HW]
    assert max(counts) > 4 and min(counts) <= 4


def test_round_trip_rebuilds_the_blends():
    """hardware columns -> p3cs_blends_from_hardware gives back, row by row, the entries of the row's blend."""
    mesh = synth.make_mesh(2_000, 24, 300, seed=5, weights_per_blend=3, index_order="runs").flat
    w, ix, table = _capi.hardware_from_blends(mesh)
    row_blend, off, xf, bw = _capi.blends_from_hardware(w, ix, table)
    assert len(off) - 1 <= mesh.num_blends          # identical rows share a blend
    bi = mesh.blend_indices()
    for r in (0, 1, 777, 1999):
        b0, b1 = int(bi[r]), int(row_blend[r])
        assert np.array_equal(mesh.blend_transform[mesh.blend_offsets[b0]:mesh.blend_offsets[b0 + 1]], xf[off[b1]:off[b1 + 1]])
        assert np.array_equal(mesh.blend_weight[mesh.blend_offsets[b0]:mesh.blend_offsets[b0 + 1]], bw[off[b1]:off[b1 + 1]])


@pytest.mark.gpu
def test_mesh_rebuilt_from_hardware_columns_animates_identically(gpu_ctx):
    sm = synth.make_mesh(6_000, 32, 900, seed=6, tbn=True, index_order="runs")
    mesh = sm.flat
    w, ix, table = _capi.hardware_from_blends(mesh)
    row_blend, off, xf, bw = _capi.blends_from_hardware(w, ix, table)
    rebuilt = F.FlatMesh(**{**mesh.__dict__})
    rebuilt.arrays = [mesh.arrays[0], np.ascontiguousarray(row_blend.astype(np.uint32).reshape(-1, 1)).view(np.uint8).reshape(mesh.num_rows, 4)]
    rebuilt.blend_column = F.Column(1, 0, 1, F.NT_uint32, F.C_index)
    rebuilt.blend_offsets, rebuilt.blend_transform, rebuilt.blend_weight = off, xf, bw
    pal = synth.make_palettes(32, 1, 7)[0]
    outs = []
    for m in (mesh, rebuilt):
        dm = _capi.Mesh(gpu_ctx, m)
        g = _capi.Group(dm, 1)
        outs.append(g.animate_vertices(pal)[0])
        g.close()
        dm.close()
    assert np.array_equal(outs[0], outs[1])
