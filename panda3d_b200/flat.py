"""Flattened description of one animated GeomVertexData -- what crosses the C-ABI.

Field for field this is `p3cs_mesh_desc` of include/p3cudaskin.h: the source arrays,
the columns the blend matrices transform, the `transform_blend` index column, the
TransformBlendTable as CSR in stored entry order, the rows mask and the morph
applications in the reference's iteration order (panda/src/gobj/geomVertexData.cxx:
1462-1751).  Integer codes are GeomEnums values (panda/src/gobj/geomEnums.h:177-212).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# GeomEnums::NumericType
NT_uint8, NT_uint16, NT_uint32, NT_packed_dcba, NT_packed_dabc, NT_float32, NT_float64, \
    NT_stdfloat, NT_int8, NT_int16, NT_int32, NT_packed_ufloat = range(12)
# GeomEnums::Contents
C_other, C_point, C_clip_point, C_vector, C_texcoord, C_color, C_index, C_morph_delta, \
    C_matrix, C_normal = range(10)
# CoordinateSystem (panda/src/linmath/coordinateSystem.h)
CS_default, CS_zup_right, CS_yup_right, CS_zup_left, CS_yup_left, CS_invalid = range(6)

NUMERIC_BYTES = {NT_uint8: 1, NT_uint16: 2, NT_uint32: 4, NT_packed_dcba: 4, NT_packed_dabc: 4,
                 NT_float32: 4, NT_float64: 8, NT_stdfloat: 4, NT_int8: 1, NT_int16: 2, NT_int32: 4,
                 NT_packed_ufloat: 4}


PACKED_VALUES = {NT_packed_dcba: 4, NT_packed_dabc: 4, NT_packed_ufloat: 3}   # values held by ONE 32-bit component


def column_bytes(numeric_type: int, num_values: int) -> int:
    """Bytes of a column in its row (GeomVertexColumn::setup, geomVertexColumn.cxx:154-234): the packed types keep
    all their values in one 32-bit word."""
    if numeric_type in PACKED_VALUES:
        return 4 * (num_values // PACKED_VALUES[numeric_type])
    return num_values * NUMERIC_BYTES[numeric_type]


@dataclass
class Column:
    """One GeomVertexColumn: (array, start, num_values, numeric_type, contents)."""
    array: int
    start: int
    num_values: int
    numeric_type: int
    contents: int

    def as_tuple(self) -> Tuple[int, int, int, int, int]:
        return (self.array, self.start, self.num_values, self.numeric_type, self.contents)

    @property
    def num_bytes(self) -> int:
        return column_bytes(self.numeric_type, self.num_values)


@dataclass
class Morph:
    """One (C_morph_delta column, slider) application with the slider's rows."""
    slider: int
    base: Column
    delta: Column
    row_ranges: np.ndarray  # int32 [k, 2]


@dataclass
class FlatMesh:
    num_rows: int
    arrays: List[np.ndarray]            # uint8 [num_rows, stride] per SOURCE array
    is_output: List[bool]
    skin_columns: List[Column]          # points first, then vectors
    blend_column: Optional[Column]
    num_transforms: int
    blend_offsets: np.ndarray           # int32 [B+1]
    blend_transform: np.ndarray         # int32 [nnz]
    blend_weight: np.ndarray            # float32 [nnz]
    row_ranges: np.ndarray              # int32 [k, 2]
    num_sliders: int = 0
    morphs: List[Morph] = field(default_factory=list)
    coordinate_system: int = CS_zup_right

    @property
    def num_blends(self) -> int:
        return int(len(self.blend_offsets) - 1)

    @property
    def strides(self) -> List[int]:
        return [int(a.shape[1]) for a in self.arrays]

    @property
    def output_arrays(self) -> List[int]:
        return [i for i, o in enumerate(self.is_output) if o]

    def skinned_rows(self) -> int:
        """Rows inside TransformBlendTable::get_rows() (what `vertices/s` counts)."""
        if self.blend_column is None:
            return 0
        r = np.asarray(self.row_ranges).reshape(-1, 2)
        return int((r[:, 1] - r[:, 0]).sum())

    def blend_indices(self) -> np.ndarray:
        """The per-row transform_blend value, as GeomVertexReader::get_data1i reads it."""
        c = self.blend_column
        if c is None:
            return np.zeros(0, np.int64)
        dt = {NT_uint8: np.uint8, NT_uint16: np.uint16, NT_uint32: np.uint32}[c.numeric_type]
        nb = np.dtype(dt).itemsize
        raw = np.ascontiguousarray(self.arrays[c.array][:, c.start:c.start + nb])
        return raw.view(dt).reshape(-1).astype(np.int64)

    def validate(self) -> None:
        n = self.num_rows
        for a in self.arrays:
            if a.dtype != np.uint8 or a.ndim != 2 or a.shape[0] != n:
                raise ValueError("arrays must be uint8 [num_rows, stride]")
        for c in self.skin_columns:
            if not self.is_output[c.array]:
                raise ValueError("skin column in a non-output array")
            if c.start + c.num_bytes > self.arrays[c.array].shape[1]:
                raise ValueError("skin column exceeds the stride")
        if self.blend_column is not None:
            bi = self.blend_indices()
            r = np.asarray(self.row_ranges).reshape(-1, 2)
            for b, e in r:
                if not (0 <= b < e <= n):
                    raise ValueError("bad row range")
                if bi[b:e].size and int(bi[b:e].max()) >= self.num_blends:
                    raise ValueError("transform_blend index out of range")
            if len(self.blend_transform) and int(self.blend_transform.max()) >= self.num_transforms:
                raise ValueError("blend refers to a transform outside the palette")


def mesh_from_case(case: Dict[str, np.ndarray]) -> FlatMesh:
    """Builds a FlatMesh from a golden case written by oracle/ref_harness.cxx."""
    meta = case["meta"]
    num_rows, na = int(meta[0]), int(meta[1])
    arrays = [np.ascontiguousarray(case[f"src_array{a}"]).reshape(num_rows, -1).astype(np.uint8) for a in range(na)]
    is_output = [bool(x) for x in case["src_is_output"]]
    skin = [Column(*map(int, r)) for r in case["skin_columns"].reshape(-1, 5)]
    bc = case["blend_column"]
    blend_column = Column(*map(int, bc)) if int(bc[0]) >= 0 else None
    morphs = []
    mtab = case["morphs"].reshape(-1, 11)
    mro = case["morph_row_offsets"]
    mrows = case["morph_rows"].reshape(-1, 2)
    for i, r in enumerate(mtab):
        morphs.append(Morph(int(r[0]), Column(*map(int, r[1:6])), Column(*map(int, r[6:11])),
                            mrows[int(mro[i]):int(mro[i + 1])].astype(np.int32)))
    return FlatMesh(
        num_rows=num_rows, arrays=arrays, is_output=is_output, skin_columns=skin,
        blend_column=blend_column, num_transforms=int(meta[3]),
        blend_offsets=case["blend_offsets"].astype(np.int32),
        blend_transform=case["blend_transform"].astype(np.int32),
        blend_weight=case["blend_weight"].astype(np.float32),
        row_ranges=case["rows"].reshape(-1, 2).astype(np.int32),
        num_sliders=int(meta[5]), morphs=morphs, coordinate_system=int(meta[6]) or CS_zup_right)


@dataclass
class FlatSkeleton:
    """The joint hierarchy that produces a palette ("next" row f1): what
    MovingPartBase::do_update / CharacterJoint::update_internals read, joints listed parents-first
    (`p3cs_skeleton_desc` of include/p3cudaskin.h)."""
    parent: np.ndarray            # int32 [J], -1 = top-level joint
    has_channel: np.ndarray       # int32 [J]
    table_length: np.ndarray      # int32 [J, 12]  (i j k a b c h p r x y z; 0 = default component)
    table_offset: np.ndarray      # int32 [J, 12]
    tables: np.ndarray            # float32 [*]
    default_value: np.ndarray     # float32 [J, 16]
    initial_inverse: np.ndarray   # float32 [J, 16]
    root_xform: np.ndarray        # float32 [16]
    palette_joint: np.ndarray     # int32 [T]: joint feeding each palette slot
    coordinate_system: int = CS_zup_right
    # morph sliders driven by scalar channels (CharacterSlider / AnimChannelScalarTable), in SliderTable order
    slider_has_channel: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    slider_table_length: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # 0: empty table (value 0)
    slider_table_offset: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))   # into `tables`
    slider_default: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))

    @property
    def num_sliders(self) -> int:
        return int(len(self.slider_has_channel))

    @property
    def num_joints(self) -> int:
        return int(len(self.parent))


def skeleton_from_case(case: Dict[str, np.ndarray]) -> FlatSkeleton:
    return FlatSkeleton(
        parent=case["skel_parent"].astype(np.int32), has_channel=case["skel_has_channel"].astype(np.int32),
        table_length=case["skel_table_length"].astype(np.int32).reshape(-1, 12),
        table_offset=case["skel_table_offset"].astype(np.int32).reshape(-1, 12),
        tables=case["skel_tables"].astype(np.float32), default_value=case["skel_default_value"].astype(np.float32).reshape(-1, 16),
        initial_inverse=case["skel_initial_inverse"].astype(np.float32).reshape(-1, 16),
        root_xform=case["skel_root_xform"].astype(np.float32).reshape(16),
        palette_joint=case["skel_palette_joint"].astype(np.int32),
        coordinate_system=int(case["meta"][6]) or CS_zup_right,
        slider_has_channel=case.get("skel_slider_has_channel", np.zeros(0, np.int32)).astype(np.int32),
        slider_table_length=case.get("skel_slider_table_length", np.zeros(0, np.int32)).astype(np.int32),
        slider_table_offset=case.get("skel_slider_table_offset", np.zeros(0, np.int32)).astype(np.int32),
        slider_default=case.get("skel_slider_default", np.zeros(0, np.float32)).astype(np.float32))
